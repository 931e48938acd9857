#!/bin/bash
# Round-2 profiler captures of the bge-small ingest pass (128 chunks x 512 tokens, one GPU), final round-2 kernels
# (programmatic dependent launch, shifted LayerNorm statistics, fp32 [CLS] rows):
#   1. the command exits 0 without ncu first;
#   2. launch list -> gpurun_out/r2_launches_embed_128x512.csv;
#   3. `--set full` captures of one attention launch and of four consecutive GEMM launches (QKV, out-proj+LN, FFN-up,
#      FFN-down+LN) -> gpurun_out/prof_r2_attn.ncu-rep / prof_r2_gemm.ncu-rep, summarised by scripts/summarise_ncu.py.
set -x
CMD="python scripts/one_embed.py 2"
timeout 200 $CMD > gpurun_out/r2_one_embed.log 2>&1 || { tail -20 gpurun_out/r2_one_embed.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_embed_128x512.csv $CMD > gpurun_out/r2_ncu_embed.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 13 -c 1 -f -o gpurun_out/prof_r2_attn $CMD > gpurun_out/r2_ncu_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 50 -c 4 -f -o gpurun_out/prof_r2_gemm $CMD > gpurun_out/r2_ncu_gemm.log 2>&1
ls -la gpurun_out/prof_r2_*.ncu-rep
