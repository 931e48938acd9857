import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from financial_rag_system_b200.index import VectorIndex
from financial_rag_system_b200.sharded import PeerExchange

def _i32(x):
    return torch.as_tensor(np.asarray(x, dtype=np.uint32).astype(np.int64)).to(torch.int32).cuda()

n, cuts = 45_001, [0, 9_000, 21_345, 45_001]
g = torch.Generator(device="cuda").manual_seed(11)
x = torch.randn((n, 384), generator=g, device="cuda")
codes = torch.randint(0, 5, (n,), generator=g, device="cuda", dtype=torch.int32)
whole = VectorIndex(n); whole.add(x, codes)
shards = []
for r in range(3):
    ix = VectorIndex(cuts[r + 1] - cuts[r], base=cuts[r]); ix.add(x[cuts[r]:cuts[r + 1]], codes[cuts[r]:cuts[r + 1]]); shards.append(ix)
dev = torch.device("cuda", 0)
exs = [PeerExchange(dev, 3, r, connect=False, timeout_ms=3000) for r in range(3)]
PeerExchange.link(exs)
q = x[:32] + 0.1 * torch.randn((32, 384), generator=g, device="cuda")
qc, qm = _i32(codes[:32].cpu().numpy()), _i32(np.full(32, 0x80FFFFFF, np.uint32))
wi, ws = whole.search(q, qc, qm, 15)
torch.cuda.synchronize()
mode = sys.argv[1] if len(sys.argv) > 1 else "async"
t0 = time.time()
if mode == "async":
    pend = [shards[r].search_async(q, qc, qm, 15, exchange=exs[r]) for r in range(3)]
    print("enqueued", time.time() - t0)
    for r, p in enumerate(pend):
        gi, gs = p.wait()
        print(r, "waited", round(time.time() - t0, 3), "equal", torch.equal(gi, wi), gi[0, :4].tolist())
else:
    for r in range(3):
        shards[r].search_push(q, qc, qm, 15, exs[r])
    print("pushed", time.time() - t0)
    for r in range(3):
        gi, gs = exs[r].wait_merge(32, 15)
        torch.cuda.synchronize()
        print(r, "waited", round(time.time() - t0, 3), "equal", torch.equal(gi, wi), gi[0, :4].tolist())
for r in range(3):
    try:
        exs[r].status(); print(r, "status ok")
    except Exception as e:
        print(r, "status", e)
