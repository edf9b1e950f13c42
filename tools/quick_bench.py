"""Provisional device-time probe (debug aid; the contract benchmark is bench.py)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import oracle  # noqa: E402
import kanpyo_b200  # noqa: E402
from kanpyo_b200 import corpus  # noqa: E402
from helpers import to_product_dict  # noqa: E402

od = oracle.load_ipadic()
tk = kanpyo_b200.Tokenizer(to_product_dict(od), device=0)
v = corpus.Vocabulary(od.keywords, od.morphs)
kind = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
text, off = corpus.synth_corpus(v, n, kind)
print(kind, n, "bytes", text.size, flush=True)
for it in range(5):
    t0 = time.perf_counter()
    res = tk.tokenize_batch_bytes(text, off)
    dt = time.perf_counter() - t0
    p = tk.profile()
    print("iter", it, "wall %.2f ms" % (dt * 1e3), {k: round(v, 3) if isinstance(v, float) else v for k, v in p.items()},
          "B/s(device total) %.3e" % (text.size / (p["total_ms"] * 1e-3)), flush=True)
print(tk.counters())
