"""Tuning sweep on the GPU box (debug aid, not a test): the in-tree library and every variant under
kanpyo_b200/_variants (tools/build_variants.sh), each in its own process:
  * bit-exact parity against the oracle on 30 000 / 14 000 / 3 000-sentence cfg2 samples, a cfg4 sample and the edge-case
    sentences;
  * CUDA-event stage times of the 65 536-sentence cfg2 batch (text resident in HBM), median of --steps passes;
    with --cfg4 also the 4096 x 4096-char batch.
KP_SWEEP_SIZES=16384,8192 adds cfg2 batches of those sizes.
usage: python tools/sweep.py [--steps 10] [--cfg4] [--full-parity]  (without --full-parity only the 3 000-sentence cfg2 sample)"""
import argparse
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one(lib, steps, cfg4, full):
    if lib != "default":
        os.environ["KANPYO_B200_LIB"] = lib
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import torch
    from oracle import oracle          # debug script: compares against the checker
    import kanpyo_b200
    from kanpyo_b200 import corpus
    from helpers import to_product_dict, assert_batch_equal

    od = oracle.load_ipadic()
    orc = oracle.OracleTokenizer(od)
    tk = kanpyo_b200.Tokenizer(to_product_dict(od), device=0)
    tk.set_path("pipeline")
    v = corpus.Vocabulary(od.keywords, od.morphs)
    edge = ["", "あ", "ー" * 300, "a" * 500, "ア" * 1030, "\x00あ\x00い", "𠮷野家で𩸽を食べた", "1" * 60 + "犬" + "ア" * 40 + "abc" * 30]
    blobs = [e.encode("utf-8") for e in edge]
    eoff = np.zeros(len(blobs) + 1, np.uint64)
    eoff[1:] = np.cumsum([len(b) for b in blobs])
    # 30 000 / 14 000 / 3 000 sentences: the sweep picks its lanes per sentence (8 / 16 / 32) from the batch size
    sizes = (30000, 14000, 3000) if full else (3000,)
    samples = [corpus.synth_corpus(v, m, "cfg2") for m in sizes] + [corpus.synth_corpus(v, 24, "cfg4"),
                                                                    (np.frombuffer(b"".join(blobs), np.uint8), eoff)]
    ok = True
    for text, off in samples:
        res = tk.tokenize_batch_bytes(text, off)
        o_off, o_tok, o_cost, _ = orc.tokenize_batch(text, off, threads=8)
        try:
            assert_batch_equal(res, o_off, o_tok, o_cost)
        except AssertionError as e:
            ok = False
            print("PARITY FAILED", str(e)[:200], flush=True)
    rows = []
    extra = tuple(("cfg2", int(x)) for x in os.environ.get("KP_SWEEP_SIZES", "").split(",") if x)
    for kind, n in (("cfg2", 65536),) + extra + ((("cfg4", 4096),) if cfg4 else ()):
        text, off = corpus.synth_corpus(v, n, kind)
        t_ = torch.from_numpy(text.copy()).cuda()
        o_ = torch.from_numpy(off.astype(np.int64)).cuda()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        keys = ("prep_ms", "lattice_ms", "bucket_ms", "viterbi_ms", "backtrace_ms", "total_ms")
        acc = {k: [] for k in keys}
        for i in range(3 + steps):
            flush.zero_()
            torch.cuda.synchronize()
            tk.tokenize_batch_device8(t_.data_ptr(), o_.data_ptr(), n, 0, int(off[n]))
            p = tk.profile()
            if i >= 3:
                for k in keys:
                    acc[k].append(p[k])
        med = {k: sorted(x)[len(x) // 2] for k, x in acc.items()}
        rows.append("%s/%d %s" % (kind, n, " ".join("%s=%.3f" % (k[:-3], med[k]) for k in keys)))
    print("%-28s parity=%s | %s" % (os.path.basename(lib), "ok" if ok else "FAILED", " | ".join(rows)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--cfg4", action="store_true")
    ap.add_argument("--full-parity", action="store_true")
    ap.add_argument("--one")
    a = ap.parse_args()
    if a.one:
        return one(a.one, a.steps, a.cfg4, a.full_parity)
    libs = ["default"] + sorted(glob.glob(os.path.join(ROOT, "kanpyo_b200", "_variants", "*.so")))
    for lib in libs:
        cmd = [sys.executable, os.path.abspath(__file__), "--one", lib, "--steps", str(a.steps)] + (["--cfg4"] if a.cfg4 else []) + (["--full-parity"] if a.full_parity else [])
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        out = [l for l in r.stdout.splitlines() if "parity=" in l or "PARITY" in l]
        print("\n".join(out) if out else "%s: no result\n%s" % (lib, (r.stdout + r.stderr)[-800:]), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "sweep.txt"), "a") as f:
            f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    main()
