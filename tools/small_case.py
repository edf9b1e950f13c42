"""Tiny end-to-end run used under compute-sanitizer on the GPU box (debug aid, not a test)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from oracle import oracle  # noqa: E402  (debug script: compares against the checker)
import kanpyo_b200  # noqa: E402
from kanpyo_b200 import corpus  # noqa: E402
from helpers import to_product_dict, assert_batch_equal  # noqa: E402

od = oracle.load_ipadic()
orc = oracle.OracleTokenizer(od)
tk = kanpyo_b200.Tokenizer(to_product_dict(od), device=0)
for s in ["すもももももももものうち", "", "Tシャツを3枚買ったABC"]:
    toks, cost = tk.tokenize_with_cost(s)
    print(repr(s), cost, [(t.id, t.surface) for t in toks], flush=True)
    print("  oracle:", orc.tokenize(s)[1], [(t[0], t[5]) for t in orc.tokenize(s)[0]], flush=True)
v = corpus.Vocabulary(od.keywords, od.morphs)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
edge = ["", "あ", "ー" * 300, "a" * 500, "ア" * 1030, "\x00あ\x00い", "𠮷野家で𩸽を食べた", "1" * 60 + "犬" + "ア" * 40 + "abc" * 30]
for kind, m in (("cfg2", n), ("cfg3", n + 70)):          # n <= 64: the one-round-trip path; above: classify + class kernels
    text, off = corpus.synth_corpus(v, m, kind)
    for path in ("fused", "pipeline"):
        tk.set_path(path)
        res = tk.tokenize_batch_bytes(text, off)
        o_off, o_tok, o_cost, ctr = orc.tokenize_batch(text, off, threads=4)
        print(kind, path, "gpu tokens", len(res.tokens), "oracle tokens", len(o_tok), "counters", tk.counters(), ctr, flush=True)
        print("profile", tk.profile(), flush=True)
        assert_batch_equal(res, o_off, o_tok, o_cost)
for path in ("fused", "pipeline"):
    tk.set_path(path)
    blobs = [e.encode("utf-8") for e in edge]
    eoff = np.zeros(len(blobs) + 1, np.uint64)
    eoff[1:] = np.cumsum([len(b) for b in blobs])
    etext = np.frombuffer(b"".join(blobs), np.uint8)
    res = tk.tokenize_batch_bytes(etext, eoff)
    o_off, o_tok, o_cost, _ = orc.tokenize_batch(etext, eoff)
    assert_batch_equal(res, o_off, o_tok, o_cost)
print("PARITY OK on", n, "sentences, both paths")
