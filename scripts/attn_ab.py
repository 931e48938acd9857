"""One workload of encoder_perf.py (bge 128 x 512) for A/B runs of differently-flagged library builds:
FRS_B200_LIB=build/libfrs_X.so python scripts/attn_ab.py [n_seqs seq_len]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from encoder_perf import run, BGE_SMALL
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
s = int(sys.argv[2]) if len(sys.argv) > 2 else 512
print("lib:", os.environ.get("FRS_B200_LIB", "default"))
run(f"bge {n}x{s}", BGE_SMALL, 1234, [s] * n)
