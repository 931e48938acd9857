"""A few encoder passes over 128 chunks x 512 tokens (the `--workload embed` step), for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from financial_rag_system_b200.checkpoint import BGE_SMALL, synthetic_checkpoint
from financial_rag_system_b200.encoder import BertEncoder

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n, s = 128, 512
enc = BertEncoder(BGE_SMALL, synthetic_checkpoint(BGE_SMALL, 1234), device=0, max_tokens=n * s)
cu = (np.arange(n + 1) * s).astype(np.int32)
ids = torch.from_numpy(np.random.default_rng(0).integers(1000, 30522, size=n * s).astype(np.int32)).cuda()
for _ in range(passes):
    out = enc.embed_device(ids, cu)
torch.cuda.synchronize()
print("ok", float(out.abs().sum()))
