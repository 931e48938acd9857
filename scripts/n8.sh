#!/bin/bash
# usage: scripts/n8.sh <nranks> <steps> [extra env...]
N=$1; STEPS=$2; shift 2
show='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "1by1", round(d["e2e"]["one_at_a_time_qps"]), "scan", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],3), "host_us", d.get("host_enqueue_us_per_step"), d["breakdown_ms"], d["parity_checked"], d.get("per_rank"), d.get("timed_region_ms_per_rank"))'
env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps $STEPS --warmup 5 --no-secondary > gpurun_out/n8_last.log 2>&1
tail -1 gpurun_out/n8_last.log | python -c "$show" "N=$N steps=$STEPS $*" || tail -20 gpurun_out/n8_last.log
