"""Per-kernel-class timing of the encoders (CUDA events inside the library) on synthetic token ids."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from financial_rag_system_b200.checkpoint import BGE_SMALL, MINILM_L6_CE, synthetic_checkpoint
from financial_rag_system_b200.encoder import BertEncoder

def flops(shape, lens):
    lens = np.asarray(lens, dtype=np.float64)
    lin = 2 * 384 * (1152 + 384 + 1536 + 1536) * lens.sum()
    att = 4 * 384 * (lens ** 2).sum()
    return shape.layers * (lin + att)

def run(name, shape, seed, lens, pairs=False, iters=5):
    enc = BertEncoder(shape, synthetic_checkpoint(shape, seed), device=0, max_tokens=max(4096, int(sum((l + 7) // 8 * 8 for l in lens)) + 128))
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    rng = np.random.default_rng(0)
    ids = torch.from_numpy(rng.integers(1000, 30000, size=int(cu[-1])).astype(np.int32)).cuda()
    tts = torch.zeros_like(ids)
    f = (lambda: enc.score_device(ids, tts, cu)) if pairs else (lambda: enc.embed_device(ids, cu))
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    enc.set_profiling(True)
    f()
    torch.cuda.synchronize()
    prof = enc.read_profile()
    enc.set_profiling(False)
    fl = flops(shape, lens)
    print(f"{name}: {len(lens)} seqs, {int(cu[-1])} tokens: {ms:.3f} ms/pass, {len(lens) / ms * 1e3:.0f} seq/s, "
          f"{fl / ms / 1e9:.1f} TFLOP/s ({fl / ms / 1e9 / 1389.5 * 100:.1f}% of sustained bf16 peak)")
    print("   ", {k: round(v, 3) for k, v in prof.items()})
    enc.close()

if __name__ == "__main__":
    rng = np.random.default_rng(1)
    run("bge 128x512", BGE_SMALL, 1234, [512] * 128)
    run("bge 256 x N(230,40)", BGE_SMALL, 1234, np.clip(rng.normal(230, 40, 256).astype(int), 16, 512).tolist())
    run("bge 32 queries x 16", BGE_SMALL, 1234, [16] * 32, iters=20)
    run("ce 480 pairs x ~256", MINILM_L6_CE, 4321, np.clip(rng.normal(256, 40, 480).astype(int), 32, 512).tolist(), pairs=True)
