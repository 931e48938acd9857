set -x
for w in embed rerank pipeline; do timeout 300 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench_${w}_r1b.log 2>&1; tail -c 600 gpurun_out/bench_${w}_r1b.log; done
timeout 200 python scripts/one_embed.py 2 > gpurun_out/one_embed.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_embed_r1b.csv python scripts/one_embed.py 2 > gpurun_out/ncu_embed.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 13 -c 1 -f -o gpurun_out/prof_r1b_attn python scripts/one_embed.py 2 > gpurun_out/ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 50 -c 4 -f -o gpurun_out/prof_r1b_gemm python scripts/one_embed.py 2 > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
