"""BERT checkpoints for the two encoders of the reference (main.py:84 `BAAI/bge-small-en-v1.5`,
main.py:90 `cross-encoder/ms-marco-MiniLM-L-6-v2`).

A checkpoint is a plain dict {HF state_dict name: float32 numpy array} plus a `BertShape`.  It comes
either from a local Hugging Face directory (`config.json` + `model.safetensors` [+ `vocab.txt`]) —
which is how real weights drop in — or from `synthetic_checkpoint`, a seeded random set with the
exact shapes of the real models (this image has no network and no weights on disk).  The names are
the ones `transformers.BertModel` / `BertForSequenceClassification` use, so the same dict loads
into the Hugging Face modules (that is what the oracle does) and into libfrs_b200.so
(`weight_table`, order documented in include/frs_b200.h).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class BertShape:
    vocab_size: int = 30522
    hidden: int = 384
    layers: int = 12
    heads: int = 12
    intermediate: int = 1536
    max_pos: int = 512
    type_vocab: int = 2
    has_head: bool = False
    ln_eps: float = 1e-12

    def hf_config(self) -> dict:
        """kwargs of transformers.BertConfig for this shape."""
        cfg = dict(vocab_size=self.vocab_size, hidden_size=self.hidden, num_hidden_layers=self.layers,
                   num_attention_heads=self.heads, intermediate_size=self.intermediate,
                   max_position_embeddings=self.max_pos, type_vocab_size=self.type_vocab, hidden_act="gelu",
                   layer_norm_eps=self.ln_eps, pad_token_id=0, hidden_dropout_prob=0.0,
                   attention_probs_dropout_prob=0.0)
        if self.has_head:
            cfg["num_labels"] = 1
        return cfg


BGE_SMALL = BertShape(layers=12, has_head=False)       # BAAI/bge-small-en-v1.5 (33.4 M parameters)
MINILM_L6_CE = BertShape(layers=6, has_head=True)       # cross-encoder/ms-marco-MiniLM-L-6-v2 (22.7 M)

_LAYER_TENSORS = (
    "attention.self.query.weight", "attention.self.query.bias",
    "attention.self.key.weight", "attention.self.key.bias",
    "attention.self.value.weight", "attention.self.value.bias",
    "attention.output.dense.weight", "attention.output.dense.bias",
    "attention.output.LayerNorm.weight", "attention.output.LayerNorm.bias",
    "intermediate.dense.weight", "intermediate.dense.bias",
    "output.dense.weight", "output.dense.bias",
    "output.LayerNorm.weight", "output.LayerNorm.bias",
)


def tensor_names(shape: BertShape) -> list[str]:
    """State-dict names (BertModel naming, no `bert.` prefix) in the order of the C ABI weight table."""
    names = ["embeddings.word_embeddings.weight", "embeddings.position_embeddings.weight",
             "embeddings.token_type_embeddings.weight", "embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias"]
    for l in range(shape.layers):
        names += [f"encoder.layer.{l}.{t}" for t in _LAYER_TENSORS]
    if shape.has_head:
        names += ["pooler.dense.weight", "pooler.dense.bias", "classifier.weight", "classifier.bias"]
    return names


def tensor_shape(shape: BertShape, name: str) -> tuple[int, ...]:
    H, F = shape.hidden, shape.intermediate
    if name == "embeddings.word_embeddings.weight":
        return (shape.vocab_size, H)
    if name == "embeddings.position_embeddings.weight":
        return (shape.max_pos, H)
    if name == "embeddings.token_type_embeddings.weight":
        return (shape.type_vocab, H)
    if name == "classifier.weight":
        return (1, H)
    if name == "classifier.bias":
        return (1,)
    if name.endswith("intermediate.dense.weight"):
        return (F, H)
    if name.endswith("intermediate.dense.bias"):
        return (F,)
    if name.endswith(".output.dense.weight") and "attention" not in name:
        return (H, F)
    if name.endswith(".weight") and "LayerNorm" not in name:
        return (H, H)
    return (H,)


def synthetic_checkpoint(shape: BertShape, seed: int) -> dict[str, np.ndarray]:
    """Seeded random weights with the real shapes.  Scales are chosen so that the network is not
    degenerate: attention logits have a standard deviation of ~2.5 (peaked, like a trained model),
    LayerNorm gains/biases are perturbed, every bias is non-zero — a kernel that drops a bias, a
    residual or a scale does not pass parity.  numpy's PCG64 stream is identical on every machine."""
    rng = np.random.default_rng(seed)
    out: dict[str, np.ndarray] = {}
    for name in tensor_names(shape):
        shp = tensor_shape(shape, name)
        if "LayerNorm.weight" in name:
            w = 1.0 + 0.1 * rng.standard_normal(shp)
        elif "LayerNorm.bias" in name:
            w = 0.05 * rng.standard_normal(shp)
        elif name.endswith(".bias"):
            w = 0.05 * rng.standard_normal(shp)
        elif "word_embeddings" in name:
            w = 0.6 * rng.standard_normal(shp)
        elif "position_embeddings" in name or "token_type" in name:
            w = 0.3 * rng.standard_normal(shp)
        elif ".query." in name or ".key." in name:
            w = 0.08 * rng.standard_normal(shp)
        elif name == "classifier.weight":
            w = 0.25 * rng.standard_normal(shp)
        elif name == "pooler.dense.weight":
            w = 0.06 * rng.standard_normal(shp)
        else:
            w = 0.04 * rng.standard_normal(shp)
        out[name] = np.ascontiguousarray(w, dtype=np.float32)
    return out


def weight_table(shape: BertShape, weights: dict[str, np.ndarray]) -> list[np.ndarray]:
    """float32 C-contiguous arrays in the order frs_encoder_create expects."""
    table = []
    for name in tensor_names(shape):
        if name not in weights:
            raise KeyError(f"checkpoint has no tensor {name!r}")
        a = np.ascontiguousarray(weights[name], dtype=np.float32)
        if tuple(a.shape) != tensor_shape(shape, name):
            raise ValueError(f"{name}: shape {a.shape}, expected {tensor_shape(shape, name)}")
        table.append(a)
    return table


def load_hf_directory(path: str) -> tuple[BertShape, dict[str, np.ndarray]]:
    """Read `config.json` + `model.safetensors` of a local Hugging Face BERT directory (what
    SentenceTransformer(...)/CrossEncoder(...) download in the reference's Dockerfile:32-34)."""
    from safetensors.numpy import load_file

    with open(os.path.join(path, "config.json")) as f:
        c = json.load(f)
    raw = load_file(os.path.join(path, "model.safetensors"))
    weights = {}
    for k, v in raw.items():
        k = k[5:] if k.startswith("bert.") else k
        weights[k] = np.asarray(v, dtype=np.float32)
    shape = BertShape(vocab_size=c["vocab_size"], hidden=c["hidden_size"], layers=c["num_hidden_layers"],
                      heads=c["num_attention_heads"], intermediate=c["intermediate_size"],
                      max_pos=c["max_position_embeddings"], type_vocab=c.get("type_vocab_size", 2),
                      has_head="classifier.weight" in weights, ln_eps=c.get("layer_norm_eps", 1e-12))
    return shape, weights


def save_hf_directory(path: str, shape: BertShape, weights: dict[str, np.ndarray], vocab: list[str] | None = None):
    """Write the checkpoint in Hugging Face layout (so that `from_pretrained(path)` reads it)."""
    from safetensors.numpy import save_file

    os.makedirs(path, exist_ok=True)
    cfg = shape.hf_config()
    cfg["model_type"] = "bert"
    cfg["architectures"] = ["BertForSequenceClassification" if shape.has_head else "BertModel"]
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f, indent=1)
    pre = "bert." if shape.has_head else ""
    tensors = {(k if k.startswith("classifier.") else pre + k): np.ascontiguousarray(v) for k, v in weights.items()}
    save_file(tensors, os.path.join(path, "model.safetensors"))
    if vocab is not None:
        with open(os.path.join(path, "vocab.txt"), "w") as f:
            f.write("\n".join(vocab) + "\n")
