"""Deterministic synthetic Japanese corpora for the BASELINE.json configs.

The reference ships no corpus and no benchmark; the workloads are defined here once so the CUDA
path, the CPU oracle and the benchmark all consume identical bytes.

  cfg2 / cfg5  n sentences, target length max(8, round(N(80, 25))) chars, Zipf(s=1) over the
               dictionary's unique surfaces ranked by (min word cost, surface bytes), closed by '。'
  cfg3         "Wikipedia-shape": log-normal lengths (median 60, sigma 0.6, clipped to [5, 400]) and
               8 % of word slots replaced by non-dictionary runs (ASCII words, digits, katakana, space)
  cfg4         long-line stress: every sentence has a 4096-char target
"""
from __future__ import annotations

import hashlib

import numpy as np

SEED = 20261017
_PERIOD = "。".encode("utf-8")


class Vocabulary:
    """Unique dictionary surfaces ranked by (min word cost asc, surface bytes asc)."""

    def __init__(self, keywords, morphs):
        cost = np.asarray(morphs)[:, 2].astype(np.int64)
        best = {}
        for k, c in zip(keywords, cost.tolist()):
            p = best.get(k)
            if p is None or c < p:
                best[k] = c
        ranked = sorted(best.items(), key=lambda kv: (kv[1], kv[0]))
        self.words = [k for k, _ in ranked]
        self.byte_len = np.array([len(w) for w in self.words], np.int64)
        self.char_len = np.array([sum((b & 0xC0) != 0x80 for b in w) for w in self.words], np.int64)
        self.byte_off = np.zeros(len(self.words) + 1, np.int64)
        self.byte_off[1:] = np.cumsum(self.byte_len)
        self.blob = np.frombuffer(b"".join(self.words), np.uint8)
        w = 1.0 / np.arange(1, len(self.words) + 1, dtype=np.float64)     # Zipf, s = 1
        self.cdf = np.cumsum(w / w.sum())

    def sample(self, rng, n):
        return np.minimum(np.searchsorted(self.cdf, rng.random(n), side="right"), len(self.words) - 1)


def _extra_words(rng, n):
    """Non-dictionary runs for cfg3: ASCII word 2-10 letters, 1-6 digits, katakana run 2-8, one space."""
    out = []
    kinds = rng.integers(0, 4, n)
    for k in kinds.tolist():
        if k == 0:
            m = int(rng.integers(2, 11))
            out.append(bytes(rng.integers(97, 123, m).astype(np.uint8)))
        elif k == 1:
            m = int(rng.integers(1, 7))
            out.append(bytes(rng.integers(48, 58, m).astype(np.uint8)))
        elif k == 2:
            m = int(rng.integers(2, 9))
            out.append("".join(chr(c) for c in rng.integers(0x30A1, 0x30FB, m).tolist()).encode("utf-8"))
        else:
            out.append(b" ")
    return out


def _targets(rng, kind, n):
    if kind in ("cfg2", "cfg5"):
        return np.maximum(8, np.rint(rng.normal(80.0, 25.0, n))).astype(np.int64)
    if kind == "cfg3":
        return np.clip(np.rint(np.exp(rng.normal(np.log(60.0), 0.6, n))), 5, 400).astype(np.int64)
    if kind == "cfg4":
        return np.full(n, 4096, np.int64)
    raise ValueError("unknown corpus kind %r" % kind)


def synth_corpus(vocab: Vocabulary, n_sent: int, kind: str = "cfg2", seed: int = SEED):
    """-> (text uint8 [n_bytes], offsets uint64 [n_sent+1])."""
    rng = np.random.default_rng(seed)
    targets = _targets(rng, kind, n_sent)
    words = list(vocab.words)
    byte_len, char_len = vocab.byte_len, vocab.char_len
    period_id = len(words)
    words.append(_PERIOD)
    byte_len = np.append(byte_len, len(_PERIOD))
    char_len = np.append(char_len, 1)
    mean_chars = float((char_len[:-1] * np.diff(np.concatenate([[0.0], vocab.cdf]))).sum())
    parts, offsets = [], np.zeros(n_sent + 1, np.uint64)
    done = 0
    pos = 0
    BLOCK = 16384
    while done < n_sent:
        nb = min(BLOCK, n_sent - done)
        tg = targets[done:done + nb]
        need = int(tg.sum() / max(mean_chars, 1.0) * 1.3) + 64 * nb
        ids = vocab.sample(rng, need)
        if kind == "cfg3":
            repl = np.nonzero(rng.random(need) < 0.08)[0]
            extra = _extra_words(rng, len(repl))
            base = len(words)
            words.extend(extra)
            byte_len = np.append(byte_len, [len(e) for e in extra])
            char_len = np.append(char_len, [sum((b & 0xC0) != 0x80 for b in e) for e in extra])
            ids = ids.copy()
            ids[repl] = base + np.arange(len(repl))
        cl = char_len[ids]
        cum = np.concatenate([[0], np.cumsum(cl)])
        seq = []
        start = 0
        for s in range(nb):
            end = int(np.searchsorted(cum, cum[start] + tg[s], side="left"))
            if end > need:
                raise RuntimeError("corpus generator under-sampled; raise the 1.3 factor")
            seq.append(ids[start:end])
            seq.append(np.array([period_id]))
            start = end
            offsets[done + s + 1] = 0   # filled below
        seq_ids = np.concatenate(seq)
        bl = byte_len[seq_ids]
        # sentence byte lengths
        is_end = seq_ids == period_id
        sent_of = np.cumsum(is_end) - is_end
        sbytes = np.bincount(sent_of, weights=bl, minlength=nb).astype(np.int64)
        offsets[done + 1:done + nb + 1] = pos + np.cumsum(sbytes)
        pos += int(sbytes.sum())
        parts.append(b"".join(words[i] for i in seq_ids.tolist()))
        if kind == "cfg3":
            del words[period_id + 1:]
            byte_len = byte_len[:period_id + 1]
            char_len = char_len[:period_id + 1]
        done += nb
    text = np.frombuffer(b"".join(parts), np.uint8)
    assert int(offsets[-1]) == text.size
    return text, offsets


def sha256(text: np.ndarray) -> str:
    return hashlib.sha256(text.tobytes()).hexdigest()


def shard_by_bytes(offsets: np.ndarray, n_shards: int):
    """Contiguous sentence ranges balanced by cumulative bytes -> [(s0, s1)] * n_shards."""
    off = np.asarray(offsets, np.uint64)
    n = len(off) - 1
    total = int(off[-1] - off[0])
    cuts = [0]
    for k in range(1, n_shards):
        target = int(off[0]) + total * k // n_shards
        cuts.append(int(np.searchsorted(off, target, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.clip(cuts, 0, n))
    return [(int(cuts[k]), int(cuts[k + 1])) for k in range(n_shards)]
