"""Run a few searches on a synthetic corpus (for ncu launch lists / captures).
usage: python scripts/one_search.py <bf16|f32> <rows> <iters> [nomatch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from financial_rag_system_b200.index import VectorIndex  # noqa: E402

dtype, n, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
nomatch = len(sys.argv) > 4 and sys.argv[4] == "nomatch"
g = torch.Generator(device="cuda").manual_seed(9)
ix = VectorIndex(n, dtype=dtype)
chunk = 1 << 18
for s in range(0, n, chunk):
    m = min(chunk, n - s)
    ix.add(torch.randn((m, 384), generator=g, device="cuda"),
           torch.randint(0, 500, (m,), generator=g, device="cuda", dtype=torch.int32))
q = torch.randn((32, 384), generator=g, device="cuda")
qc = torch.zeros(32, dtype=torch.int32, device="cuda")
qm = torch.full((32,), 0x80000000 - (1 << 32), dtype=torch.int64).to(torch.int32).cuda()
if nomatch:
    qc = torch.full((32,), 0xFFFFFF, dtype=torch.int32, device="cuda")
    qm = torch.full((32,), 0x80FFFFFF - (1 << 32), dtype=torch.int64).to(torch.int32).cuda()
for _ in range(iters):
    ids, sc = ix.search(q, qc, qm, 15)
torch.cuda.synchronize()
print("ok", ix.last_stats(), sc[0, :3].tolist())
