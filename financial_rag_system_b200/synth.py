"""Seeded synthetic SEC-style corpus (BASELINE.json configs[0]: "10k synthetic SEC chunks").

The reference ingests real EDGAR filings (ingest.py:101-145, network) and cuts them into
~1000-character chunks with 200 characters of overlap (ingest.py:25-26, 71-81), each stored with the
payload {ticker, document_type, text, source_file, ingested_at} under an md5 point id
(ingest.py:148-168).  This module produces chunks of the same shape from a template grammar, plus
analyst-style questions about them, so that benchmarks and tests run without the network.
"""
from __future__ import annotations

import hashlib

import numpy as np

CHUNK_SIZE, CHUNK_OVERLAP = 1000, 200  # ingest.py:25-26
DOC_TYPES = ("10-K", "10-Q")

_SEGMENTS = ["Products", "Services", "Cloud", "Devices", "Advertising", "Subscriptions", "Licensing", "Payments",
             "Enterprise", "Consumer", "Americas", "Europe", "Greater China", "Japan", "Rest of Asia Pacific"]
_DRIVERS = ["higher unit sales", "favorable product mix", "foreign currency headwinds", "increased component costs",
            "growth in paid subscribers", "pricing actions taken during the period", "supply chain constraints",
            "weaker consumer demand", "expansion of the installed base", "higher advertising spend",
            "new product introductions", "the timing of customer deployments", "lower manufacturing yields"]
_RISKS = ["global and regional economic conditions", "competition in highly volatile markets", "reliance on single-source suppliers",
          "changes in tax law and the interpretation of tax rules", "cybersecurity incidents and data breaches",
          "legal and regulatory compliance obligations", "fluctuations in foreign exchange rates",
          "the ability to retain key personnel", "interruptions of information technology systems",
          "political events, trade disputes and tariffs"]
_TEMPLATES = [
    "{seg} net sales {dir} {pct}% or ${amt} million during fiscal {year} compared to fiscal {prev} due primarily to {drv}. ",
    "Gross margin percentage for {seg} was {pct2}% compared to {pct3}% in the prior period, reflecting {drv} partially offset by {drv2}. ",
    "Research and development expense {dir} {pct}% to ${amt} million, driven by increases in headcount-related expenses and {drv}. ",
    "The Company's business, results of operations and financial condition could be materially adversely affected by {risk} and {risk2}. ",
    "Selling, general and administrative expense was ${amt} million in the {q} quarter of {year}, a change of {pct}% that reflects {drv}. ",
    "As of the end of the period the Company had ${amt} million in cash, cash equivalents and marketable securities and ${amt2} million of term debt outstanding. ",
    "During {year} the Company repurchased ${amt} million of its common stock and paid dividends and dividend equivalents of ${amt2} million. ",
    "The Company's effective tax rate for {year} was {pct2}% compared to {pct3}% for {prev}, the difference being due to {drv} and {risk}. ",
    "Management believes that {seg} demand will continue to be influenced by {risk}; actual results may differ materially from these forward-looking statements. ",
    "Item {item}. {ticker} operating income for {seg} {dir} to ${amt} million as {drv} more than offset {drv2}. ",
]
_QUESTIONS = [
    "What was {ticker}'s total revenue in fiscal {year}?", "How did {ticker}'s {seg} segment perform in {year}?",
    "What are the main risk factors {ticker} disclosed related to {riskshort}?", "How much did {ticker} spend on research and development?",
    "Why did {ticker}'s gross margin change compared to the prior year?", "How much cash and marketable securities does {ticker} hold?",
    "What did {ticker} return to shareholders through buybacks and dividends?", "What was {ticker}'s effective tax rate and why did it change?",
    "Compare {ticker}'s {seg} growth with the impact of foreign currency.", "What trends does {ticker}'s management expect for {seg} demand?",
]


def tickers(n: int = 50, seed: int = 1234) -> list[str]:
    rng = np.random.default_rng(seed)
    out, seen = [], set()
    while len(out) < n:
        t = "".join(chr(65 + int(c)) for c in rng.integers(0, 26, int(rng.integers(3, 5))))
        if t not in seen:
            seen.add(t)
            out.append(t)
    return out


def _sentence(rng, ticker: str) -> str:
    year = int(rng.integers(2019, 2025))
    f = dict(seg=_SEGMENTS[int(rng.integers(len(_SEGMENTS)))], dir=("increased", "decreased")[int(rng.integers(2))],
             pct=int(rng.integers(1, 40)), pct2=round(float(rng.uniform(10, 60)), 1), pct3=round(float(rng.uniform(10, 60)), 1),
             amt=f"{int(rng.integers(50, 99000)):,}", amt2=f"{int(rng.integers(50, 99000)):,}", year=year, prev=year - 1,
             drv=_DRIVERS[int(rng.integers(len(_DRIVERS)))], drv2=_DRIVERS[int(rng.integers(len(_DRIVERS)))],
             risk=_RISKS[int(rng.integers(len(_RISKS)))], risk2=_RISKS[int(rng.integers(len(_RISKS)))],
             q=("first", "second", "third", "fourth")[int(rng.integers(4))], item=int(rng.integers(1, 16)), ticker=ticker)
    return _TEMPLATES[int(rng.integers(len(_TEMPLATES)))].format(**f)


def make_chunks(n_chunks: int, n_tickers: int = 50, seed: int = 1234):
    """-> (ids, texts, payloads): n_chunks chunks of ~CHUNK_SIZE characters, consecutive chunks of a
    filing overlapping by ~CHUNK_OVERLAP characters; ids are md5 hex digests as in ingest.py:151-158."""
    rng = np.random.default_rng(seed)
    tk = tickers(n_tickers, seed)
    ids, texts, payloads = [], [], []
    filing = 0
    while len(texts) < n_chunks:
        t = tk[int(rng.integers(len(tk)))]
        d = DOC_TYPES[int(rng.integers(len(DOC_TYPES)))]
        per_filing = int(rng.integers(8, 40))
        body = ""
        while len(body) < CHUNK_SIZE + per_filing * (CHUNK_SIZE - CHUNK_OVERLAP):
            body += _sentence(rng, t)
        for c in range(per_filing):
            if len(texts) >= n_chunks:
                break
            s = c * (CHUNK_SIZE - CHUNK_OVERLAP)
            text = body[s:s + CHUNK_SIZE]
            src = f"filing_{filing}.txt"
            ids.append(hashlib.md5(f"{t}_{d}_{src}_{c}".encode()).hexdigest())
            texts.append(text)
            payloads.append({"ticker": t, "document_type": d, "text": text, "source_file": src, "ingested_at": "synthetic"})
        filing += 1
    return ids, texts, payloads


def make_queries(n: int, n_tickers: int = 50, seed: int = 99):
    """-> (queries, tickers): analyst-style questions, each about one of the corpus's tickers."""
    rng = np.random.default_rng(seed)
    tk = tickers(n_tickers, 1234)
    qs, ts = [], []
    for _ in range(n):
        t = tk[int(rng.integers(len(tk)))]
        q = _QUESTIONS[int(rng.integers(len(_QUESTIONS)))].format(
            ticker=t, year=int(rng.integers(2019, 2025)), seg=_SEGMENTS[int(rng.integers(len(_SEGMENTS)))],
            riskshort=_RISKS[int(rng.integers(len(_RISKS)))].split(" and ")[0])
        qs.append(q)
        ts.append(t)
    return qs, ts
