"""Pretty-print gpurun_out/gemm_trace_<epi>.txt (FRS_GEMM_TRACE builds): merged timeline of CTA 0."""
import sys
names = {1: "I  tile start", 2: "I  tempty ok", 3: "I  full ok (k-step)", 4: "I  tile committed", 10: "E tile start", 11: "E tfull ok",
         12: "E chunk loaded", 13: "E chunk staged", 14: "E handed to store warp", 16: "E box free", 17: "E LN pass 1 done", 18: "E LN statistics exchanged", 15: "E store issued", 30: "P empty ok", 40: "I  uses acc 0", 41: "I  uses acc 1", 50: "E released acc 0", 51: "E released acc 1"}
ev = [tuple(map(int, l.split())) for l in open(sys.argv[1])]
t0 = min(e[1] for e in ev)
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 20000)
last = {0: None, 1: None, 2: None}
for r, t, i in sorted(ev, key=lambda e: e[1]):
    t -= t0
    d = t - last[r] if last[r] is not None else 0
    last[r] = t
    if lo <= t <= hi:
        col = {0: 0, 1: 36, 2: 72}[r]
        print(f"{t:8d} " + " " * col + f"{names[i]} (+{d})")
