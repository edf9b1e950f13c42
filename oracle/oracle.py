"""TEST INFRASTRUCTURE (oracle) — ctypes front-end for the CPU checker.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (kanpyo_b200/) never does.

Wraps `_build/libkanpyo_oracle.so` (da_build.c + ref_tokenize.cpp, built by `make -C oracle`) and
`dictbuild.py`, and caches the IPADIC dictionary built by the *oracle's* builder under
`oracle/_build/` so the GPU box does not rebuild it for every test process.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tarfile
import tempfile

import numpy as np

from . import dictbuild

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB_PATH = os.path.join(_BUILD, "libkanpyo_oracle.so")
IPADIC_TARBALL = os.path.join(os.path.dirname(_HERE), "third_party", "mecab-ipadic",
                              "mecab-ipadic-2.7.0-20070801.tar.gz")
IPADIC_SHA256 = "b62f527d881c504576baed9c6ef6561554658b175ce6ae0096a60307e49e3523"

_lib = None


def build_native(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("da_build.c", "ref_tokenize.cpp", "Makefile")]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Result(C.Structure):
    _fields_ = [("n_sent", C.c_uint64), ("tok_off", C.POINTER(C.c_uint64)), ("tokens", C.POINTER(C.c_int64)),
                ("eos_cost", C.POINTER(C.c_int32)), ("counters", C.c_uint64 * 7)]


class _Lattice(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("nodes", C.POINTER(C.c_int64)), ("dp", C.POINTER(C.c_int64)),
                ("pre", C.POINTER(C.c_int64)), ("n_path", C.c_uint64), ("path", C.POINTER(C.c_uint64)),
                ("n_buckets", C.c_uint64), ("edge_off", C.POINTER(C.c_uint64)), ("edge_idx", C.POINTER(C.c_uint64))]


def lib():
    global _lib
    if _lib is None:
        build_native()
        L = C.CDLL(_LIB_PATH)
        L.ko_da_build.restype = C.c_int
        L.ko_da_build.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_uint64)]
        L.ko_da_free.argtypes = [C.c_void_p]
        L.ko_dict_create.restype = C.c_void_p
        L.ko_dict_create.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                     C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64,
                                     C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.ko_dict_destroy.argtypes = [C.c_void_p]
        L.ko_da_search.restype = C.c_int64
        L.ko_da_search.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        L.ko_common_prefix.restype = C.c_int64
        L.ko_common_prefix.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_uint64]
        L.ko_conn_get.restype = C.c_int16
        L.ko_conn_get.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.ko_tokenize_batch.restype = C.POINTER(_Result)
        L.ko_tokenize_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
        L.ko_result_free.argtypes = [C.POINTER(_Result)]
        L.ko_lattice_dump.restype = C.POINTER(_Lattice)
        L.ko_lattice_dump.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        L.ko_lattice_free.argtypes = [C.POINTER(_Lattice)]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def da_build(keys, ids) -> np.ndarray:
    """da::build_with_ids (da.rs:206-217) through da_build.c.  keys: sorted unique bytes objects."""
    L = lib()
    blob = np.frombuffer(b"".join(keys), dtype=np.uint8) if keys else np.zeros(0, np.uint8)
    off = np.zeros(len(keys) + 1, dtype=np.uint64)
    if keys:
        off[1:] = np.cumsum([len(k) for k in keys], dtype=np.uint64)
    ids_a = np.asarray(ids, dtype=np.int64)
    out = C.c_void_p()
    out_len = C.c_uint64()
    rc = L.ko_da_build(_ptr(blob) if blob.size else None, _ptr(off), len(keys), _ptr(ids_a) if len(keys) else None,
                       C.byref(out), C.byref(out_len))
    if rc != 0:
        raise MemoryError("ko_da_build failed")
    n = out_len.value
    arr = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int32)), shape=(n, 2)).copy()
    L.ko_da_free(out)
    return arr


def da_build_plain(keys) -> np.ndarray:
    """da::build (da.rs:191-204): ids = 1..=n."""
    return da_build(keys, list(range(1, len(keys) + 1)))


_NPZ_FIELDS = ("da", "dup_ids", "dup_counts", "morphs", "conn", "char_category", "invoke_list", "group_list",
               "unk_cat", "unk_first_id", "unk_count", "unk_morphs")


def _save(d: dictbuild.OracleDict, path: str):
    feats_rows, feats_names = d.features
    unk_rows, unk_names = d.unk_features
    kw_blob = b"".join(d.keywords)
    kw_off = np.zeros(len(d.keywords) + 1, np.uint64)
    kw_off[1:] = np.cumsum([len(k) for k in d.keywords], dtype=np.uint64)

    def rows(r):
        off = np.zeros(len(r) + 1, np.uint64)
        off[1:] = np.cumsum([len(x) for x in r], dtype=np.uint64)
        flat = np.fromiter((v for x in r for v in x), dtype=np.uint32, count=int(off[-1]))
        return off, flat

    f_off, f_flat = rows(feats_rows)
    u_off, u_flat = rows(unk_rows)
    np.savez(path, **{k: getattr(d, k) for k in _NPZ_FIELDS},
             conn_shape=np.array([d.conn_row, d.conn_col], np.uint64),
             char_class=np.array(d.char_class), kw_blob=np.frombuffer(kw_blob, np.uint8), kw_off=kw_off,
             f_off=f_off, f_flat=f_flat, f_names=np.array(feats_names),
             u_off=u_off, u_flat=u_flat, u_names=np.array(unk_names))


def _load(path: str) -> dictbuild.OracleDict:
    z = np.load(path, allow_pickle=False)
    kw_blob = z["kw_blob"].tobytes()
    kw_off = z["kw_off"]
    keywords = [kw_blob[int(kw_off[i]):int(kw_off[i + 1])] for i in range(len(kw_off) - 1)]

    def rows(off, flat):
        return [flat[int(off[i]):int(off[i + 1])].tolist() for i in range(len(off) - 1)]

    return dictbuild.OracleDict(
        **{k: z[k] for k in _NPZ_FIELDS}, conn_row=int(z["conn_shape"][0]), conn_col=int(z["conn_shape"][1]),
        char_class=[str(x) for x in z["char_class"]], keywords=keywords,
        features=(rows(z["f_off"], z["f_flat"]), [str(x) for x in z["f_names"]]),
        unk_features=(rows(z["u_off"], z["u_flat"]), [str(x) for x in z["u_names"]]))


def extract_ipadic(dst: str) -> str:
    """Extract the vendored tarball; returns the directory holding the CSV/def files."""
    with tarfile.open(IPADIC_TARBALL, "r:gz") as t:
        t.extractall(dst, filter="data")
    return os.path.join(dst, "mecab-ipadic-2.7.0-20070801")


_IPADIC = None


def load_ipadic(rebuild: bool = False) -> dictbuild.OracleDict:
    """IPADIC built by the ORACLE's restatement of the reference builder (cached as .npz)."""
    global _IPADIC
    if _IPADIC is not None and not rebuild:
        return _IPADIC
    cache = os.path.join(_BUILD, "ipadic_oracle.npz")
    if os.path.exists(cache) and not rebuild:
        _IPADIC = _load(cache)
        return _IPADIC
    os.makedirs(_BUILD, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        root = extract_ipadic(tmp)
        d = dictbuild.from_dir(root, "euc-jp", da_build)
    _save(d, cache)
    _IPADIC = d
    return d


TOKEN_FIELDS = ("id", "class", "position", "start", "end", "byte_len")
NODE_FIELDS = ("kind", "id", "byte_pos", "char_pos", "end_char_pos", "left_id", "right_id", "cost", "byte_len")
COUNTER_FIELDS = ("B", "C", "P", "P_ok", "N", "E", "T")
DUMMY, KNOWN, UNKNOWN = 0, 1, 2


class OracleTokenizer:
    """Tokenizer (src/tokenizer.rs:7-45) over an OracleDict, executed by ref_tokenize.cpp."""

    def __init__(self, d: dictbuild.OracleDict):
        self.d = d
        L = lib()
        c = np.ascontiguousarray
        self._keep = [c(d.da, np.int32), c(d.dup_ids, np.int64), c(d.dup_counts, np.uint64), c(d.morphs, np.int16),
                      c(d.conn, np.int16), c(d.char_category, np.uint8), c(d.invoke_list, np.uint8),
                      c(d.group_list, np.uint8), c(d.unk_cat, np.uint8), c(d.unk_first_id, np.int64),
                      c(d.unk_count, np.uint64), c(d.unk_morphs, np.int16)]
        k = self._keep
        self.h = L.ko_dict_create(_ptr(k[0]), len(k[0]), _ptr(k[1]), _ptr(k[2]), len(k[1]), _ptr(k[3]), len(k[3]),
                                  d.conn_row, d.conn_col, _ptr(k[4]), _ptr(k[5]), len(k[5]), _ptr(k[6]), len(k[6]),
                                  _ptr(k[7]), len(k[7]), _ptr(k[8]), _ptr(k[9]), _ptr(k[10]), len(k[8]),
                                  _ptr(k[11]), len(k[11]))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ko_dict_destroy(self.h)
            self.h = None

    # --- dictionary lookups -------------------------------------------------------------------
    def da_search(self, s: str):
        b = s.encode("utf-8")
        r = lib().ko_da_search(self.h, b, len(b))
        return r if r != 0 else None

    def common_prefix(self, s: str, use_dup: bool = True):
        b = s.encode("utf-8")
        cap = 4096
        ids = np.zeros(cap, np.int64)
        lens = np.zeros(cap, np.uint64)
        n = lib().ko_common_prefix(self.h, b, len(b), 1 if use_dup else 0, _ptr(ids), _ptr(lens), cap)
        if n < 0:
            return None
        return [(int(ids[i]), int(lens[i])) for i in range(n)]

    def conn_get(self, row: int, col: int) -> int:
        return int(lib().ko_conn_get(self.h, row, col))

    # --- tokenize -----------------------------------------------------------------------------
    def tokenize_batch(self, blob: bytes | np.ndarray, offsets: np.ndarray, threads: int = 1, collect: bool = True):
        """-> (tok_off uint64[n+1], tokens int64[n_tok,6], eos_cost int32[n], counters dict)"""
        blob_a = np.frombuffer(blob, np.uint8) if isinstance(blob, (bytes, bytearray)) else np.ascontiguousarray(blob, np.uint8)
        off = np.ascontiguousarray(offsets, np.uint64)
        n = len(off) - 1
        r = lib().ko_tokenize_batch(self.h, _ptr(blob_a) if blob_a.size else None, _ptr(off), n, threads,
                                    1 if collect else 0)
        try:
            res = r.contents
            tok_off = np.ctypeslib.as_array(res.tok_off, shape=(n + 1,)).copy()
            nt = int(tok_off[-1])
            tokens = (np.ctypeslib.as_array(res.tokens, shape=(nt, 6)).copy() if nt and collect
                      else np.zeros((0, 6), np.int64))
            cost = np.ctypeslib.as_array(res.eos_cost, shape=(n,)).copy() if n else np.zeros(0, np.int32)
            counters = {k: int(res.counters[i]) for i, k in enumerate(COUNTER_FIELDS)}
        finally:
            lib().ko_result_free(r)
        return tok_off, tokens, cost, counters

    def tokenize(self, s: str):
        """Tokenizer::tokenize -> list of (id, class, position, start, end, surface), plus dp[EOS]."""
        b = s.encode("utf-8")
        _, toks, cost, _ = self.tokenize_batch(b, np.array([0, len(b)], np.uint64))
        out = []
        for t in toks:
            surface = "EOS" if t[1] == DUMMY else b[int(t[2]):int(t[2] + t[5])].decode("utf-8")
            out.append((int(t[0]), int(t[1]), int(t[2]), int(t[3]), int(t[4]), surface))
        return out, int(cost[0])

    def lattice(self, s: str):
        """Lattice::build + viterbi internals: dict(nodes int64[n,9], dp, pre, path, edge_off, edge_idx)."""
        b = s.encode("utf-8")
        r = lib().ko_lattice_dump(self.h, b, len(b))
        try:
            la = r.contents
            n = int(la.n_nodes)
            nb = int(la.n_buckets)
            edge_off = np.ctypeslib.as_array(la.edge_off, shape=(nb + 1,)).copy()
            ne = int(edge_off[-1])
            return dict(
                nodes=np.ctypeslib.as_array(la.nodes, shape=(n, 9)).copy(),
                dp=np.ctypeslib.as_array(la.dp, shape=(n,)).copy(),
                pre=np.ctypeslib.as_array(la.pre, shape=(n,)).copy(),
                path=(np.ctypeslib.as_array(la.path, shape=(int(la.n_path),)).copy() if la.n_path
                      else np.zeros(0, np.uint64)),
                edge_off=edge_off,
                edge_idx=(np.ctypeslib.as_array(la.edge_idx, shape=(ne,)).copy() if ne else np.zeros(0, np.uint64)))
        finally:
            lib().ko_lattice_free(r)

    def features(self, token) -> str:
        """print_tokens' feature join (src/bin/kanpyo.rs:174-197)."""
        tid, cls = token[0], token[1]
        if cls == KNOWN:
            rows, names = self.d.features
        elif cls == UNKNOWN:
            rows, names = self.d.unk_features
        else:
            return ""
        return ",".join(names[i] for i in rows[tid - 1])


def dict_from_keywords(sorted_keywords, morphs=None, conn=None, conn_shape=(1, 1), char_class=("DEFAULT",),
                       category=None, invoke=(False,), group=(False,), unk_map=None, unk_morphs=None):
    """Small hand-built dictionaries (mirrors src/tests.rs:8-108) for pins and edge-case tests."""
    kws = [k.encode("utf-8") if isinstance(k, str) else k for k in sorted_keywords]
    da, dup_ids, dup_counts = dictbuild.build_index(kws, da_build)
    n = len(kws)
    morphs = np.zeros((n, 3), np.int16) if morphs is None else np.asarray(morphs, np.int16).reshape(-1, 3)
    conn = np.zeros(conn_shape[0] * conn_shape[1], np.int16) if conn is None else np.asarray(conn, np.int16)
    category = np.zeros(1 << 16, np.uint8) if category is None else np.asarray(category, np.uint8)
    unk_map = unk_map or {}
    cats = sorted(unk_map)
    return dictbuild.OracleDict(
        da=da, dup_ids=dup_ids, dup_counts=dup_counts, morphs=morphs, conn_row=conn_shape[0], conn_col=conn_shape[1],
        conn=conn, char_class=list(char_class), char_category=category,
        invoke_list=np.asarray(invoke, np.uint8), group_list=np.asarray(group, np.uint8),
        unk_cat=np.asarray(cats, np.uint8), unk_first_id=np.asarray([unk_map[c][0] for c in cats], np.int64),
        unk_count=np.asarray([unk_map[c][1] for c in cats], np.uint64),
        unk_morphs=(np.zeros((0, 3), np.int16) if unk_morphs is None else np.asarray(unk_morphs, np.int16).reshape(-1, 3)),
        keywords=kws)
