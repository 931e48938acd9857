"""Host-side WordPiece tokenisation (stays on the CPU, as in the reference, where
`SentenceTransformer.encode` / `CrossEncoder.predict` call the Rust `tokenizers` library).

What differs from the reference is the OUTPUT layout: instead of a padded [batch, max_len] id matrix
plus attention mask, `pack_texts` / `pack_pairs` emit PACKED ids with prefix sums (`cu_seqlens`), the
layout the varlen CUDA encoder consumes — no padding tokens are ever computed on.

Semantics follow `transformers.BertTokenizer(do_lower_case=True)` (tokenization_bert.py:79-136):
BertNormalizer (clean text, lower-case, strip accents, split CJK), whitespace + punctuation
pre-tokenisation, greedy longest-match WordPiece with `##` continuation pieces and `[UNK]` for words
over 100 characters, `[CLS] a [SEP]` / `[CLS] a [SEP] b [SEP]` with token types 0/1, truncation to 512
(`longest_first` for pairs).  tests/test_encoder_oracle_cpu.py checks the ids against BertTokenizer.

This image has no `vocab.txt` of the real models (no network), so `synthetic_vocab` builds a seeded
30 522-entry vocabulary with the special tokens at their real BERT ids; a real `vocab.txt` drops in
through `WordPiece.from_vocab_file`.
"""
from __future__ import annotations

import numpy as np

PAD, UNK, CLS, SEP, MASK = 0, 100, 101, 102, 103
MAX_LEN = 512

_FIN_WORDS = """the of and to in a for is on that by this with as are was at from or an be its which
company we our revenue net income sales fiscal year quarter ended total operating expenses cash
million billion increase decrease compared prior period primarily due higher lower products services
segment margin gross cost costs research development selling general administrative tax taxes rate
effective interest expense debt notes senior term credit facility shares common stock repurchase
dividend dividends per share diluted basic earnings risk risks factors could adversely affect
business results operations financial condition market markets customers demand supply chain
competition competitive regulatory laws regulations compliance litigation legal proceedings
intellectual property cybersecurity data privacy international foreign currency exchange rates
inflation economic conditions assets liabilities equity balance sheet statement statements flows
investing financing activities capital expenditures liquidity resources obligations commitments
lease leases goodwill intangible impairment acquisition acquisitions fair value securities
marketable investments deferred recognized recognition performance contract contracts management
discussion analysis item annual report form filed commission exchange act section pursuant
iphone mac ipad wearables home accessories services americas europe greater china japan rest asia
pacific cloud software hardware platform advertising subscription subscribers users growth
employees employee compensation benefits stock based awards units restricted options plan plans
forward looking statements believe expect anticipate intend estimate may will should
not no any all other such these those than more less also including include includes
new first second third fourth three six nine twelve months weeks ended december september june march
january february april july august october november 2019 2020 2021 2022 2023 2024 2025
how what why did does much many who when where are there discuss explain describe main key"""


def synthetic_vocab(size: int = 30522, seed: int = 20240607) -> list[str]:
    """Seeded WordPiece vocabulary with BERT's id layout: [PAD]=0, [unused*], [UNK]=100, [CLS]=101,
    [SEP]=102, [MASK]=103, single characters, then words and `##` continuation pieces."""
    vocab: list[str] = ["[PAD]"] + [f"[unused{i}]" for i in range(99)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    vocab += [f"[unused{i}]" for i in range(99, 99 + 895)]  # ids 104..998, as in bert-base-uncased
    seen = set(vocab)

    def add(tok: str) -> None:
        if tok not in seen and len(vocab) < size:
            seen.add(tok)
            vocab.append(tok)

    chars = "!\"#$%&'()*+,-./0123456789:;<=>?@[\\]^_`abcdefghijklmnopqrstuvwxyz{|}~"
    for ch in chars:
        add(ch)
    for ch in "abcdefghijklmnopqrstuvwxyz0123456789":
        add("##" + ch)
    for w in _FIN_WORDS.split():
        add(w)
    rng = np.random.default_rng(seed)
    onsets = ["", "b", "c", "d", "f", "g", "h", "l", "m", "n", "p", "r", "s", "t", "v", "w", "st", "tr", "pr", "ch", "sh"]
    nuclei = ["a", "e", "i", "o", "u", "ea", "io", "ou", "ai"]
    codas = ["", "n", "r", "s", "t", "l", "m", "nt", "st", "ng", "rs"]
    suffixes = ["s", "ed", "ing", "ion", "ions", "ly", "er", "al", "ment", "ity", "ive", "able", "ized", "ance"]
    for s in suffixes:
        add("##" + s)
    while len(vocab) < size:
        nsyl = int(rng.integers(1, 4))
        w = "".join(onsets[int(rng.integers(len(onsets)))] + nuclei[int(rng.integers(len(nuclei)))] +
                    codas[int(rng.integers(len(codas)))] for _ in range(nsyl))
        if len(w) < 2:
            continue
        add(w if rng.random() < 0.7 else "##" + w)
    return vocab


class WordPiece:
    """BERT uncased WordPiece tokenizer emitting packed ids."""

    def __init__(self, vocab: list[str]):
        from tokenizers import Tokenizer, models, normalizers, pre_tokenizers

        self.vocab = list(vocab)
        self.token_to_id = {t: i for i, t in enumerate(self.vocab)}
        for tok, idx in (("[PAD]", PAD), ("[UNK]", UNK), ("[CLS]", CLS), ("[SEP]", SEP)):
            if self.token_to_id.get(tok) != idx:
                raise ValueError(f"{tok} must have id {idx}")
        tk = Tokenizer(models.WordPiece(vocab=self.token_to_id, unk_token="[UNK]", max_input_chars_per_word=100))
        tk.normalizer = normalizers.BertNormalizer(clean_text=True, handle_chinese_chars=True, strip_accents=None,
                                                   lowercase=True)
        tk.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
        self._tk = tk
        self._cache: dict[str, np.ndarray] = {}
        self.cache_entries = 262144

    @classmethod
    def synthetic(cls) -> "WordPiece":
        return cls(synthetic_vocab())

    @classmethod
    def from_vocab_file(cls, path: str) -> "WordPiece":
        with open(path, encoding="utf-8") as f:
            return cls([line.rstrip("\n") for line in f])

    def _word_ids(self, texts: list[str]) -> list[list[int]]:
        enc = self._tk.encode_batch_fast if hasattr(self._tk, "encode_batch_fast") else self._tk.encode_batch
        return [e.ids for e in enc(list(texts), add_special_tokens=False)]

    def _cached_ids(self, texts: list[str]) -> list[np.ndarray]:
        """WordPiece ids (no specials) as int32 arrays, through a bounded text -> ids cache: the rerank
        path sees the same chunk texts again and again (they come out of the chunk store), only the
        query side is new."""
        cache = self._cache
        missing = [t for t in dict.fromkeys(texts) if t not in cache]
        if missing:
            if len(cache) + len(missing) > self.cache_entries:
                cache.clear()
            for t, ids in zip(missing, self._word_ids(missing)):
                cache[t] = np.asarray(ids[:MAX_LEN], dtype=np.int32)
        return [cache[t] for t in texts]

    def pack_texts(self, texts: list[str], max_len: int = MAX_LEN):
        """`[CLS] t [SEP]` per text, truncated to max_len.  Returns (ids int32 [total], cu_seqlens
        int32 [n+1]).  The Rust tokenizer runs on all host threads and releases the GIL; the packing
        below is vectorised numpy (no per-token Python work)."""
        body = self._word_ids(texts)
        n = len(body)
        blen = np.fromiter((min(len(b), max_len - 2) for b in body), dtype=np.int64, count=n)
        cu = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(blen + 2, out=cu[1:])
        ids = np.empty(int(cu[-1]), dtype=np.int32)
        ids[cu[:-1]] = CLS
        ids[cu[1:] - 1] = SEP
        if int(blen.sum()):
            flat = np.concatenate([np.asarray(b[:k], dtype=np.int32) for b, k in zip(body, blen) if k])
            pos = np.arange(flat.size, dtype=np.int64) + np.repeat(2 * np.arange(n, dtype=np.int64) + 1, blen)
            ids[pos] = flat
        return ids, cu

    def pack_pairs(self, pairs: list, max_len: int = MAX_LEN):
        """`[CLS] a [SEP] b [SEP]` with token types 0/1; `longest_first` truncation as the Rust `tokenizers`
        library implements it (TruncationStrategy::LongestFirst, what AutoTokenizer's fast tokenizer
        does): only the longer side is cut while the shorter fits, else both go to budget/2 (+1 for the
        longer one when the budget is odd).  Returns (ids, type_ids, cu_seqlens)."""
        a_ids = self._cached_ids([p[0] for p in pairs])
        b_ids = self._cached_ids([p[1] for p in pairs])
        n = len(pairs)
        budget = max_len - 3
        la = np.fromiter((len(a) for a in a_ids), dtype=np.int64, count=n)
        lb = np.fromiter((len(b) for b in b_ids), dtype=np.int64, count=n)
        over = la + lb > budget
        if over.any():
            n1, n2 = np.minimum(la, lb), np.maximum(la, lb)
            swap = la > lb
            n2 = np.where(n1 > budget, n1, np.maximum(n1, budget - n1))
            both = n1 + n2 > budget
            n1 = np.where(both, budget // 2, n1)
            n2 = np.where(both, budget // 2 + budget % 2, n2)
            ta, tb = np.where(swap, n2, n1), np.where(swap, n1, n2)
            la, lb = np.where(over, ta, la), np.where(over, tb, lb)
        cu = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(la + lb + 3, out=cu[1:])
        ids = np.empty(int(cu[-1]), dtype=np.int32)
        tts = np.zeros(int(cu[-1]), dtype=np.int32)
        for i in range(n):
            o, x, y = int(cu[i]), int(la[i]), int(lb[i])
            ids[o] = CLS
            ids[o + 1:o + 1 + x] = a_ids[i][:x]
            ids[o + 1 + x] = SEP
            ids[o + 2 + x:o + 2 + x + y] = b_ids[i][:y]
            ids[o + 2 + x + y] = SEP
            tts[o + 2 + x:o + 3 + x + y] = 1
        return ids, tts, cu
