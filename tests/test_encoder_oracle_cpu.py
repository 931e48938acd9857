"""Pins the encoder oracle: the float32 restatement (oracle/encoder_oracle.py) against the real
transformers modules and against the committed golden vectors (tests/golden/encoder_golden.npz,
generated from transformers' BertModel / BertForSequenceClassification by make_encoder_golden.py);
the host-side tokenizer against transformers.BertTokenizer; checkpoint plumbing."""
import os

import numpy as np
import pytest

from financial_rag_system_b200.checkpoint import (BGE_SMALL, MINILM_L6_CE, BertShape, load_hf_directory,
                                                  save_hf_directory, synthetic_checkpoint, tensor_names, weight_table)
from financial_rag_system_b200.tokenizer import CLS, SEP, WordPiece, synthetic_vocab
from oracle import encoder_oracle as eo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encoder_golden.npz")


@pytest.fixture(scope="module")
def tok():
    return WordPiece.synthetic()


def test_restatement_matches_transformers_bert_model():
    shape = BertShape(layers=3)
    w = synthetic_checkpoint(shape, 9)
    rng = np.random.default_rng(0)
    lens = [7, 64, 1, 130]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    ids = rng.integers(1000, 30522, size=int(cu[-1])).astype(np.int32)
    hf = eo.hf_model(shape, w)
    for pool in ("cls", "mean"):
        assert np.abs(eo.embed(shape, w, ids, cu, pool) - eo.hf_embed(hf, ids, cu, pool)).max() < 2e-6


def test_restatement_matches_transformers_sequence_classification():
    shape = BertShape(layers=2, has_head=True)
    w = synthetic_checkpoint(shape, 10)
    rng = np.random.default_rng(1)
    lens = [12, 90, 33]
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    ids = rng.integers(1000, 30522, size=int(cu[-1])).astype(np.int32)
    tts = (rng.random(int(cu[-1])) < 0.5).astype(np.int32)
    got = eo.score_pairs(shape, w, ids, tts, cu)
    ref = eo.hf_score_pairs(eo.hf_model(shape, w), ids, tts, cu)
    assert np.abs(got - ref).max() < 1e-5


def test_golden_embeddings_and_logits():
    g = np.load(GOLD)
    w = synthetic_checkpoint(BGE_SMALL, 1234)
    e = eo.embed(BGE_SMALL, w, g["ids"], g["cu"], "cls")
    assert np.abs(e - g["emb_cls"]).max() < 5e-6
    assert np.allclose(np.linalg.norm(e, axis=1), 1.0, atol=1e-6)
    assert np.abs(eo.embed(BGE_SMALL, w, g["ids"], g["cu"], "mean") - g["emb_mean"]).max() < 5e-6
    wc = synthetic_checkpoint(MINILM_L6_CE, 4321)
    logits = eo.score_pairs(MINILM_L6_CE, wc, g["pair_ids"], g["pair_types"], g["pair_cu"])
    assert np.abs(logits - g["logits"]).max() < 2e-5
    assert np.array_equal(eo.rerank(logits[:8], 5), g["top5"])


def test_golden_token_ids_are_what_the_tokenizer_produces(tok):
    import importlib.util

    spec = importlib.util.spec_from_file_location("mk", os.path.join(os.path.dirname(GOLD), "make_encoder_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    g = np.load(GOLD)
    ids, cu = tok.pack_texts(mk.TEXTS)
    assert np.array_equal(ids, g["ids"]) and np.array_equal(cu, g["cu"])


def test_tokenizer_matches_transformers_bert_tokenizer(tok):
    from transformers import BertTokenizer

    hf = BertTokenizer(vocab={t: i for i, t in enumerate(tok.vocab)}, do_lower_case=True)
    texts = ["What was Apple's total revenue in fiscal 2023?", "Net sales increased 8% to $394.3 billion; été naïve café.",
             "x" * 150 + " ok", "", "  multiple   spaces\tand\nnewlines 中文 test", "revenue " * 700]
    ids, cu = tok.pack_texts(texts)
    for i, t in enumerate(texts):
        assert ids[cu[i]:cu[i + 1]].tolist() == hf(t, truncation=True, max_length=512)["input_ids"], t[:30]
    assert cu[-1] - cu[-2] == 512 and ids[cu[-2]] == CLS and ids[cu[-1] - 1] == SEP
    rng = np.random.default_rng(3)
    pairs = [["q " * int(rng.integers(0, 600)), "d " * int(rng.integers(0, 700))] for _ in range(25)] + [["a", "b"], ["q", "d " * 600], ["q " * 600, "d"]]
    pi, pt, pc = tok.pack_pairs(pairs)
    for i, (a, b) in enumerate(pairs):
        e = hf(a, b, truncation="longest_first", max_length=512)
        assert pi[pc[i]:pc[i + 1]].tolist() == e["input_ids"]
        assert pt[pc[i]:pc[i + 1]].tolist() == e["token_type_ids"]


def test_synthetic_vocab_layout():
    v = synthetic_vocab()
    assert len(v) == 30522 == len(set(v))
    assert (v[0], v[100], v[101], v[102], v[103]) == ("[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]")
    assert v == synthetic_vocab()  # deterministic


def test_rerank_order_is_the_references():
    """main.py:246 — np.argsort(scores)[::-1][:top_k]."""
    s = np.array([0.1, 2.0, -1.0, 2.5, 0.3], dtype=np.float32)
    assert eo.rerank(s, 3).tolist() == [3, 1, 4]
    assert eo.rerank(s, 10).tolist() == [3, 1, 4, 0, 2]


def test_weight_table_order_and_hf_directory_round_trip(tmp_path):
    shape = BertShape(layers=2, has_head=True)
    w = synthetic_checkpoint(shape, 3)
    names = tensor_names(shape)
    assert len(names) == 5 + 16 * 2 + 4  # FRS_BERT_WEIGHTS(layers, has_head) of include/frs_b200.h
    table = weight_table(shape, w)
    assert table[0].shape == (30522, 384) and table[5 + 10].shape == (1536, 384) and table[5 + 12].shape == (384, 1536)
    assert table[-2].shape == (1, 384) and table[-1].shape == (1,)
    save_hf_directory(str(tmp_path), shape, w, vocab=["[PAD]", "a"])
    shape2, w2 = load_hf_directory(str(tmp_path))
    assert shape2 == shape
    assert all(np.array_equal(w[k], w2[k]) for k in w)
    # the directory is what transformers itself reads
    from transformers import BertForSequenceClassification

    m = BertForSequenceClassification.from_pretrained(str(tmp_path))
    assert np.array_equal(m.classifier.weight.detach().numpy(), w["classifier.weight"])
    with pytest.raises(KeyError):
        weight_table(shape, {k: v for k, v in w.items() if k != "pooler.dense.bias"})
