"""Time one blocking kp_tokenize_batch8 call at several batch sizes, both device paths, beside the CPU port.
    python tools/latency_sweep.py > gpurun_out/latency_sweep.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import kanpyo_b200
from kanpyo_b200 import builder, corpus
from oracle import oracle

d = builder.ipadic()
vocab = corpus.Vocabulary(d.keywords, d.morphs)
text, off = corpus.synth_corpus(vocab, 65536, "cfg2")
otk = oracle.OracleTokenizer(oracle.load_ipadic())
tk = kanpyo_b200.Tokenizer(d, device=0)
out = {"sizes": {}}
cases = [("cfg1", None)] + [(str(n), n) for n in (1, 8, 64, 256, 1024, 2048, 4096, 8192, 16384, 65536)]
for name, n in cases:
    if n is None:
        b = "すもももももももものうち".encode(); t_, o_ = np.frombuffer(b, np.uint8), np.array([0, len(b)], np.uint64)
    else:
        t_, o_ = text[:int(off[n])], off[:n + 1]
    row = {"sentences": len(o_) - 1, "bytes": int(o_[-1])}
    reps = 200 if (n or 1) <= 64 else 30 if (n or 1) <= 4096 else 10
    ref = otk.tokenize_batch(t_, o_, threads=os.cpu_count())
    for path in ("pipeline", "fused"):
        tk.set_path(path)
        r8 = tk.tokenize_batch8_bytes(t_, o_)
        assert np.array_equal(r8[0].astype(np.uint64), ref[0]) and np.array_equal(r8[2], ref[2]), (name, path)
        for _ in range(5):
            tk.tokenize_batch8_ptr(t_.ctypes.data, o_.ctypes.data, len(o_) - 1)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            tk.tokenize_batch8_ptr(t_.ctypes.data, o_.ctypes.data, len(o_) - 1)
            ts.append((time.perf_counter() - t0) * 1e6)
        ts.sort()
        row[path + "_us_p50"] = ts[len(ts) // 2]
        row[path + "_us_p99"] = ts[min(len(ts) - 1, int(len(ts) * 0.99))]
        row[path + "_profile"] = {k: v for k, v in tk.profile().items() if k in ("total_ms", "fused_ms", "kernel_launches", "fused_sentences")}
    for threads in (1, os.cpu_count()):
        t0 = time.perf_counter()
        k = 20 if (n or 1) <= 64 else 3
        for _ in range(k):
            otk.tokenize_batch(t_, o_, threads=threads, collect=False)
        row["cpu_port_us_%dthr" % threads] = (time.perf_counter() - t0) * 1e6 / k
    out["sizes"][name] = row
    print(name, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in row.items() if not k.endswith("_profile")}, file=sys.stderr)
print(json.dumps(out))
