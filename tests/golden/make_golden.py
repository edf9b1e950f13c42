"""Regenerates tests/golden/*.json from the ORACLE (oracle/ref_tokenize.cpp over the oracle-built
IPADIC).  The reference is Rust and cannot run in this image (no cargo/rustc), so these vectors pin
the oracle against drift; the oracle itself is pinned against the reference's own known-answer tests
and README outputs in tests/test_oracle_pins.py.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import oracle  # noqa: E402
from kanpyo_b200 import corpus  # noqa: E402

SENTENCES = [
    "すもももももももものうち",          # BASELINE.json configs[0]; README.md:73-83
    "自然言語処理", "形態素解析",          # README.md:88-97
    "", "あ", "。", " ", "a", "Tシャツを3枚買ったABC", "\U0001F600の犬", "\U00020BB7野家で牛丼を食べた",
    "東京都に住んでいます。", "ﾊﾝｶｸｶﾀｶﾅ", "１２３４５円", "ＡＢＣ", "ΑΒΓαβγ", "Привет мир",
    "　", "a b", "漢字かなカナ交じり文を解析する", "カタカナカタカナカタカナカタカナ",
    "あ" * 40, "９" * 12, "http://example.com/path?q=1&r=2", "\t\n", "一二三四五六七八九十百千万億兆",
    "\x00", "a\x00b", "ー" * 1030, "亜" * 50,
]


def main():
    d = oracle.load_ipadic()
    tk = oracle.OracleTokenizer(d)
    out = {"sentences": []}
    for s in SENTENCES:
        toks, cost = tk.tokenize(s)
        out["sentences"].append({"text": s, "cost": cost,
                                 "tokens": [[t[0], t[1], t[2], t[3], t[4], t[5], tk.features(t)] for t in toks]})
    with open(os.path.join(HERE, "ipadic_sentences.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=True, indent=0)
    # a seeded slice of the cfg2 corpus: checksums of the packed oracle output
    v = corpus.Vocabulary(d.keywords, d.morphs)
    text, off = corpus.synth_corpus(v, 512, "cfg2")
    tok_off, tokens, cost, ctr = tk.tokenize_batch(text, off)
    gold = {"n_sent": 512, "kind": "cfg2", "seed": corpus.SEED, "text_sha256": corpus.sha256(text),
            "n_tokens": int(tok_off[-1]), "counters": ctr,
            "tokens_sha256": hashlib.sha256(np.ascontiguousarray(tokens[:, :5]).tobytes()).hexdigest(),
            "cost_sha256": hashlib.sha256(np.ascontiguousarray(cost).tobytes()).hexdigest(),
            "first_costs": [int(x) for x in cost[:16]]}
    with open(os.path.join(HERE, "cfg2_512.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", len(SENTENCES), "sentences;", gold["n_tokens"], "tokens in the cfg2 slice")


if __name__ == "__main__":
    main()
