"""The two BERT encoders behind the reference's duck-typed seam:

    get_embedder().encode(texts)          main.py:80-84, 148, 213; main2.py:88-96, 171
    get_reranker().predict(pairs)         main.py:86-90, 245;      main2.py:98-103, 166

`Embedder.encode` / `Reranker.predict` keep the call surface of `SentenceTransformer.encode` and
`CrossEncoder.predict` (list[str] -> float32 ndarray [n,384], L2-normalised, input order preserved;
str -> [384]; list[[q, d]] -> float32 ndarray [n] of raw logits).  Tokenisation is on the host
(tokenizer.py); everything after it is one call into libfrs_b200.so — hand-written sm_100a kernels
over PACKED token ids, no padding.  There is no CPU or PyTorch fallback: without the library or a
CUDA device construction raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from .checkpoint import BGE_SMALL, MINILM_L6_CE, BertShape, load_hf_directory, synthetic_checkpoint, weight_table
from .tokenizer import WordPiece

POOL_CLS, POOL_MEAN = 0, 1


class _BertCfg(C.Structure):
    _fields_ = [("vocab_size", C.c_int32), ("hidden", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32),
                ("intermediate", C.c_int32), ("max_pos", C.c_int32), ("type_vocab", C.c_int32),
                ("has_head", C.c_int32), ("ln_eps", C.c_float), ("precision", C.c_int32)]


PRECISIONS = {"bf16": 0, "fp32": 1, "f32": 1}


def _np_ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


class BertEncoder:
    """Handle on one `frs_encoder` (weights + workspace on one GPU)."""

    def __init__(self, shape: BertShape, weights: dict, device: int = 0, max_tokens: int = 65536,
                 precision: str = "bf16"):
        """precision "bf16": tensor-core path (bf16 operands/activations, fp32 accumulation);
        "fp32": fp32 FFMA kernels end to end (~20x slower; error < 1e-5, the parity mode)."""
        self._lib = _lib.lib()
        self.shape = shape
        self.device = int(device)
        self.precision = "fp32" if PRECISIONS[precision] else "bf16"
        table = weight_table(shape, weights)
        cfg = _BertCfg(shape.vocab_size, shape.hidden, shape.layers, shape.heads, shape.intermediate, shape.max_pos,
                       shape.type_vocab, int(shape.has_head), shape.ln_eps, PRECISIONS[precision])
        ptrs = (C.c_void_p * len(table))(*[a.ctypes.data for a in table])
        h = C.c_void_p()
        check(self._lib.frs_encoder_create(self.device, C.byref(cfg), ptrs, len(table), 0, int(max_tokens), C.byref(h)))
        self._h = h
        self.max_tokens = int(self._lib.frs_encoder_max_tokens(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.frs_encoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- host buffers in, host buffers out (copies inside the call) -------------------------------
    def embed_packed(self, ids: np.ndarray, cu_seqlens: np.ndarray, pool: int = POOL_CLS) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        cu = np.ascontiguousarray(cu_seqlens, dtype=np.int32)
        n = cu.shape[0] - 1
        out = np.empty((n, self.shape.hidden), dtype=np.float32)
        check(self._lib.frs_encoder_embed_host(self._h, _np_ptr(ids), _np_ptr(cu), n, int(pool), _np_ptr(out)))
        return out

    def score_packed(self, ids: np.ndarray, type_ids: np.ndarray, cu_seqlens: np.ndarray) -> np.ndarray:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        tts = np.ascontiguousarray(type_ids, dtype=np.int32)
        cu = np.ascontiguousarray(cu_seqlens, dtype=np.int32)
        n = cu.shape[0] - 1
        out = np.empty((n,), dtype=np.float32)
        check(self._lib.frs_encoder_score_pairs_host(self._h, _np_ptr(ids), _np_ptr(tts), _np_ptr(cu), n, _np_ptr(out)))
        return out

    # -- device tensors (torch) in/out, asynchronous on the current stream -----------------------
    def embed_device(self, ids, cu_seqlens: np.ndarray, pool: int = POOL_CLS):
        import torch

        dev = torch.device("cuda", self.device)
        cu = np.ascontiguousarray(cu_seqlens, dtype=np.int32)
        n = cu.shape[0] - 1
        ids = ids.to(device=dev, dtype=torch.int32).contiguous()
        out = torch.empty((n, self.shape.hidden), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev)
        check(self._lib.frs_encoder_embed(self._h, C.c_void_p(ids.data_ptr()), _np_ptr(cu), n, int(pool),
                                          C.c_void_p(out.data_ptr()), C.c_void_p(st.cuda_stream)))
        ids.record_stream(st)
        return out

    def score_device(self, ids, type_ids, cu_seqlens: np.ndarray):
        import torch

        dev = torch.device("cuda", self.device)
        cu = np.ascontiguousarray(cu_seqlens, dtype=np.int32)
        n = cu.shape[0] - 1
        ids = ids.to(device=dev, dtype=torch.int32).contiguous()
        tts = type_ids.to(device=dev, dtype=torch.int32).contiguous()
        out = torch.empty((n,), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev)
        check(self._lib.frs_encoder_score_pairs(self._h, C.c_void_p(ids.data_ptr()), C.c_void_p(tts.data_ptr()),
                                                _np_ptr(cu), n, C.c_void_p(out.data_ptr()), C.c_void_p(st.cuda_stream)))
        ids.record_stream(st)
        tts.record_stream(st)
        return out

    def last_hidden(self, n_tokens: int):
        """last_hidden_state [n_tokens, 384] of the most recent forward pass (test aid)."""
        import torch

        dev = torch.device("cuda", self.device)
        out = torch.empty((n_tokens, self.shape.hidden), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev)
        check(self._lib.frs_encoder_last_hidden(self._h, C.c_void_p(out.data_ptr()), int(n_tokens), C.c_void_p(st.cuda_stream)))
        return out

    def debug_read(self, which: int, n_elems: int):
        """First n_elems values of workspace buffer `which` (see frs_encoder_debug_read) as float32."""
        import torch

        dev = torch.device("cuda", self.device)
        out = torch.empty((int(n_elems),), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream(dev)
        check(self._lib.frs_encoder_debug_read(self._h, int(which), C.c_void_p(out.data_ptr()), int(n_elems),
                                               C.c_void_p(st.cuda_stream)))
        return out

    def set_profiling(self, on: bool) -> None:
        check(self._lib.frs_encoder_set_profiling(self._h, int(bool(on))))

    def read_profile(self) -> dict:
        buf = (C.c_double * 8)()
        check(self._lib.frs_encoder_read_profile(self._h, buf))
        keys = ("embed_ms", "qkv_ms", "attn_ms", "outproj_ms", "ffn_up_ms", "ffn_down_ms", "head_ms", "launches")
        return dict(zip(keys, [float(x) for x in buf]))


def _load(model, default_shape: BertShape, seed: int):
    """model: None -> seeded synthetic weights + synthetic vocab; str -> local Hugging Face directory."""
    if model is None:
        return default_shape, synthetic_checkpoint(default_shape, seed), WordPiece.synthetic()
    import os

    shape, weights = load_hf_directory(model)
    return shape, weights, WordPiece.from_vocab_file(os.path.join(model, "vocab.txt"))


class Embedder:
    """Drop-in for `SentenceTransformer("BAAI/bge-small-en-v1.5", device=...)` as the reference uses
    it: `.encode(texts)`.  bge-small-en-v1.5 pools the [CLS] token and L2-normalises (its
    sentence-transformers modules.json: Transformer -> Pooling(cls) -> Normalize)."""

    SYNTHETIC_SEED = 1234
    PIPELINE_TEXTS = 512  # texts per tokenise/encode slice of a large encode() call

    def __init__(self, model: str | None = None, device: int = 0, pool: str = "cls", max_tokens: int = 65536,
                 tokenizer: WordPiece | None = None, precision: str = "bf16"):
        shape, weights, tok = _load(model, BGE_SMALL, self.SYNTHETIC_SEED)
        if shape.has_head:
            raise ValueError("an embedding model must not carry a classifier head")
        self.tokenizer = tokenizer or tok
        self.pool = {"cls": POOL_CLS, "mean": POOL_MEAN}[pool]
        self.bert = BertEncoder(shape, weights, device=device, max_tokens=max_tokens, precision=precision)

    def encode(self, texts, **_ignored) -> np.ndarray:
        single = isinstance(texts, str)
        batch = [texts] if single else list(texts)
        if not batch:
            return np.zeros((0, self.bert.shape.hidden), dtype=np.float32)
        if len(batch) <= self.PIPELINE_TEXTS:
            ids, cu = self.tokenizer.pack_texts(batch)
            out = self.bert.embed_packed(ids, cu, self.pool)
            return out[0] if single else out
        # ingest-sized input (ingest.py:52-66 posts 64 chunks at a time; a drop-in caller can pass them
        # all): the host tokenises slice i+1 (Rust, GIL released) while the GPU encodes slice i
        from concurrent.futures import ThreadPoolExecutor

        slices = [batch[s:s + self.PIPELINE_TEXTS] for s in range(0, len(batch), self.PIPELINE_TEXTS)]
        outs = []
        with ThreadPoolExecutor(max_workers=1) as pool:
            nxt = pool.submit(self.tokenizer.pack_texts, slices[0])
            for i in range(len(slices)):
                ids, cu = nxt.result()
                if i + 1 < len(slices):
                    nxt = pool.submit(self.tokenizer.pack_texts, slices[i + 1])
                outs.append(self.bert.embed_packed(ids, cu, self.pool))
        return np.concatenate(outs)

    def close(self) -> None:
        self.bert.close()


class Reranker:
    """Drop-in for `CrossEncoder("cross-encoder/ms-marco-MiniLM-L-6-v2")` as the reference uses it:
    `.predict([[query, text], ...])` -> raw logits (this model's default activation is Identity;
    frontend.py:112-117 applies its own sigmoid)."""

    SYNTHETIC_SEED = 4321

    def __init__(self, model: str | None = None, device: int = 0, max_tokens: int = 65536,
                 tokenizer: WordPiece | None = None, precision: str = "bf16"):
        shape, weights, tok = _load(model, MINILM_L6_CE, self.SYNTHETIC_SEED)
        if not shape.has_head:
            raise ValueError("a cross-encoder needs the pooler + classifier head")
        self.tokenizer = tokenizer or tok
        self.bert = BertEncoder(shape, weights, device=device, max_tokens=max_tokens, precision=precision)

    def predict(self, pairs, **_ignored) -> np.ndarray:
        pairs = list(pairs)
        if not pairs:
            return np.zeros((0,), dtype=np.float32)
        ids, tts, cu = self.tokenizer.pack_pairs(pairs)
        return self.bert.score_packed(ids, tts, cu)

    def close(self) -> None:
        self.bert.close()
