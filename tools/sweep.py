"""Tuning sweep over library variants (debug aid; the contract benchmark is bench.py).

    python tools/sweep.py                # every kanpyo_b200/_variants/libkanpyo_b200.*.so + the in-tree library
    python tools/sweep.py --one <path>   # (internal) one library in this process

Per variant, in a fresh process: bit-exact parity against the oracle on 30 000 / 14 000 / 3 000-sentence cfg2 samples, the
edge-case sentences and a 64-sentence cfg4 sample, then CUDA-event stage times of cfg2 (65 536 sentences)
through kp_tokenize_batch, median of the timed passes.  One line per variant in gpurun_out/sweep.txt.
"""
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

EDGE = ["すもももももももものうち", "", "Tシャツを3枚買ったABC", "😀の犬", "あ" * 1030, "ア" * 1100 + "です", "a", "。",
        "ｶﾀｶﾅとカタカナと12345と hello world", "𠮷野家で𩸽を食べた", "\x00あ\x00", "東京都に住んでいます。" * 40]


def one(path, steps):
    import numpy as np
    from oracle import oracle
    import kanpyo_b200
    from kanpyo_b200 import corpus
    from helpers import to_product_dict, assert_batch_equal

    od = oracle.load_ipadic()
    orc = oracle.OracleTokenizer(od)
    tk = kanpyo_b200.Tokenizer(to_product_dict(od), device=0)
    v = corpus.Vocabulary(od.keywords, od.morphs)
    out = {"lib": os.path.basename(path)}
    try:
        text, off = corpus.synth_corpus(v, 65536, "cfg2")
        extra = [s.encode("utf-8") for s in EDGE]
        for n in (30000, 14000, 3000):      # the sweep picks its lanes per sentence from the batch size
            blob = text[:int(off[n])].tobytes() + b"".join(extra)
            offs = np.concatenate([off[:n + 1], off[n] + np.cumsum([len(e) for e in extra], dtype=np.uint64)])
            res = tk.tokenize_batch_bytes(blob, offs)
            assert_batch_equal(res, *orc.tokenize_batch(blob, offs, threads=os.cpu_count())[:3])
        t4, o4 = corpus.synth_corpus(v, 64, "cfg4")
        res = tk.tokenize_batch_bytes(t4, o4)
        assert_batch_equal(res, *orc.tokenize_batch(t4, o4, threads=os.cpu_count())[:3])
        out["parity"] = "ok"
    except AssertionError as e:
        out["parity"] = "FAIL: %s" % e
    keys = ("prep_ms", "lattice_ms", "bucket_ms", "viterbi_ms", "backtrace_ms")
    runs = []
    for it in range(3 + steps):
        tk.tokenize_batch_bytes(text, off)
        p = tk.profile()
        if it >= 3:
            runs.append([p[k] for k in keys])
    med = [statistics.median(r[i] for r in runs) for i in range(len(keys))]
    out.update({k[:-3]: round(m, 4) for k, m in zip(keys, med)})
    out["sum"] = round(sum(med), 4)
    print("SWEEP " + json.dumps(out), flush=True)


def main():
    if "--one" in sys.argv:
        path = sys.argv[sys.argv.index("--one") + 1]
        steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else 10
        return one(path, steps)
    vdir = os.path.join(ROOT, "kanpyo_b200", "_variants")
    libs = [os.path.join(ROOT, "kanpyo_b200", "libkanpyo_b200.so")]
    if os.path.isdir(vdir):
        libs += sorted(os.path.join(vdir, f) for f in os.listdir(vdir) if f.endswith(".so"))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "sweep.txt"), "a") as log:
        for lib in libs:
            env = dict(os.environ, KANPYO_B200_LIB=lib)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", lib] + sys.argv[1:], env=env,
                                   capture_output=True, text=True, timeout=240)
                lines = [ln for ln in r.stdout.splitlines() if ln.startswith("SWEEP ")]
                msg = lines[-1] if lines else "SWEEP %s crashed rc=%d: %s" % (os.path.basename(lib), r.returncode,
                                                                              (r.stderr or r.stdout)[-400:].replace("\n", " | "))
            except subprocess.TimeoutExpired:
                msg = "SWEEP %s timed out" % os.path.basename(lib)
            print(msg, flush=True)
            log.write(msg + "\n")
            log.flush()


if __name__ == "__main__":
    main()
