#!/bin/bash
# proxy for the per-GPU work of the 8-GPU run: 1.25M rows on one GPU, pipelined; sweeps the scheduling knobs
show='import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "scan", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],3), d["breakdown_ms"])'
for rows in ${ROWS_LIST:-1250000}; do
for st in ${STREAMS_LIST:-1 2}; do
for r in ${RES_LIST:-4 8 12}; do
  FRS_BENCH_NO_CPU=1 FRS_BENCH_ROWS=$rows FRS_SCAN_STREAMS=$st FRS_PIPE_RESERVE=$r python bench.py --steps ${STEPS:-20} --warmup 5 --no-secondary 2>&1 | tail -1 | python -c "$show" "rows=$rows streams=$st reserve=$r"
done; done; done
