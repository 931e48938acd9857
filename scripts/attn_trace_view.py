"""Pretty-print gpurun_out/attn_trace*.txt (FRS_ATTN_TRACE builds): merged timeline of CTA 0."""
import sys
names = {1: "I  s_free0 ok", 2: "I  s_free1 ok", 3: "I  S h0 issued", 4: "I  S h1 issued", 5: "I  p_full0 ok", 6: "I  p_full1 ok",
         7: "I  PV h0 issued", 8: "I  PV h1 issued", 10: "blk start", 11: "s_full ok", 12: "S loaded", 13: "max done", 14: "o_full ok",
         15: "o updated", 16: "turn ok", 17: "exp issued", 18: "p_full arrive", 19: "last o_full ok", 20: "item done"}
ev = [tuple(map(int, l.split())) for l in open(sys.argv[1])]
t0 = min(e[1] for e in ev)
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 20000)
last = {0: None, 1: None, 2: None}
for r, t, i in sorted(ev, key=lambda e: e[1]):
    t -= t0
    d = t - last[r] if last[r] is not None else 0
    last[r] = t
    if lo <= t <= hi:
        col = {0: 0, 1: 34, 2: 68}[r]
        print(f"{t:8d} " + " " * col + f"{'I' if r == 0 else 'AB'[r - 1]}:{names[i]} (+{d})")
