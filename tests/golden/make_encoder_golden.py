"""Generates tests/golden/encoder_golden.npz from the REAL transformers modules (BertModel /
BertForSequenceClassification, transformers 5.5.0, CPU fp32) loaded with the seeded synthetic
checkpoints — the arithmetic SentenceTransformer.encode / CrossEncoder.predict run in the reference
(main.py:148,213,245).  Run from the repo root:  python tests/golden/make_encoder_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from financial_rag_system_b200.checkpoint import BGE_SMALL, MINILM_L6_CE, synthetic_checkpoint  # noqa: E402
from financial_rag_system_b200.tokenizer import WordPiece  # noqa: E402
from oracle import encoder_oracle as eo  # noqa: E402

TEXTS = [
    "What was Apple's total revenue in fiscal 2023?",
    "Net sales increased 8% to $394.3 billion, primarily due to higher iPhone and Services net sales.",
    "The Company's gross margin percentage was 43.3 percent.",
    "Risk factors: the Company's business can be adversely affected by supply chain disruption. " * 10,
    "x",
    "Research and development expense grew 14% year over year as the Company continued to invest.",
    "Item 7. Management's Discussion and Analysis of Financial Condition and Results of Operations. " * 30,
    "How much cash did the company return to shareholders through dividends and share repurchases?",
]
QUERY = "How did iPhone revenue change compared to the prior year?"


def main():
    tok = WordPiece.synthetic()
    ids, cu = tok.pack_texts(TEXTS)
    w = synthetic_checkpoint(BGE_SMALL, 1234)
    hf = eo.hf_model(BGE_SMALL, w)
    emb_cls = eo.hf_embed(hf, ids, cu, "cls")
    emb_mean = eo.hf_embed(hf, ids, cu, "mean")
    pairs = [[QUERY, t] for t in TEXTS] + [[QUERY, TEXTS[1]]] * 7  # 15 candidates, as limit=15
    pi, pt, pc = tok.pack_pairs(pairs)
    wc = synthetic_checkpoint(MINILM_L6_CE, 4321)
    logits = eo.hf_score_pairs(eo.hf_model(MINILM_L6_CE, wc), pi, pt, pc)
    out = os.path.join(ROOT, "tests", "golden", "encoder_golden.npz")
    np.savez_compressed(out, ids=ids, cu=cu, emb_cls=emb_cls, emb_mean=emb_mean, pair_ids=pi, pair_types=pt,
                        pair_cu=pc, logits=logits, top5=eo.rerank(logits[:8], 5))
    print("wrote", out, emb_cls.shape, logits)


if __name__ == "__main__":
    main()
