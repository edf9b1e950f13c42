"""Dict — the hot-path members of the reference's `kanpyo_dict::dict::Dict` (kanpyo-dict/src/dict.rs:21-30)
as flat little-endian numpy arrays, plus the staging of those arrays to HBM through the C ABI."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib


@dataclass
class Dict:
    da: np.ndarray                 # int32 [da_len, 2] (base, check)            trie/da.rs:14-20
    dup_ids: np.ndarray            # int64 [n_dup]                               index.rs:12
    dup_counts: np.ndarray         # uint64 [n_dup]
    morphs: np.ndarray             # int16 [n, 3] (left_id, right_id, cost)      morph.rs:7-11
    conn_row: int                  # connection.rs:5-9
    conn_col: int
    conn: np.ndarray               # int16 [row*col]
    char_category: np.ndarray      # uint8 [65536]                               char_category_def.rs:15-20
    invoke_list: np.ndarray        # uint8 (bool)
    group_list: np.ndarray         # uint8 (bool)
    unk_cat: np.ndarray            # uint8 [n_map]                               unk_dict.rs:12-16
    unk_first_id: np.ndarray       # int64 [n_map]
    unk_count: np.ndarray          # uint64 [n_map]
    unk_morphs: np.ndarray         # int16 [n_unk, 3]
    # not read by the hot path
    char_class: list = field(default_factory=list)
    keywords: list = field(default_factory=list)        # sorted surfaces incl. duplicates (bytes)
    features: tuple = field(default_factory=tuple)      # (rows, names)   morph_feature.rs:7-10
    unk_features: tuple = field(default_factory=tuple)

    # kp_dict* per device.  A handle is shared by every Tokenizer built on that device and lives as
    # long as the Dict: nothing replaces or destroys it while a tokenizer may still read it.
    _handles: dict = field(default_factory=dict, repr=False, compare=False)

    # ---- C ABI ---------------------------------------------------------------------------------
    def _arrays(self):
        c = np.ascontiguousarray
        keep = dict(
            da=c(self.da, np.int32), dup_ids=c(self.dup_ids, np.int64), dup_counts=c(self.dup_counts, np.uint64),
            morphs=c(self.morphs, np.int16), conn=c(self.conn, np.int16), cat=c(self.char_category, np.uint8),
            invoke=c(self.invoke_list, np.uint8), group=c(self.group_list, np.uint8), ucat=c(self.unk_cat, np.uint8),
            ufirst=c(self.unk_first_id, np.int64), ucount=c(self.unk_count, np.uint64),
            umorphs=c(self.unk_morphs, np.int16))
        p = lambda a: a.ctypes.data_as(C.c_void_p) if a.size else None
        a = _lib.DictArrays(
            da=p(keep["da"]), da_len=keep["da"].size // 2,
            dup_ids=p(keep["dup_ids"]), dup_counts=p(keep["dup_counts"]), n_dup=keep["dup_ids"].size,
            morphs=p(keep["morphs"]), n_morphs=keep["morphs"].size // 3,
            conn_row=int(self.conn_row), conn_col=int(self.conn_col), conn=p(keep["conn"]),
            char_category=p(keep["cat"]), n_char_category=keep["cat"].size,
            invoke_list=p(keep["invoke"]), n_invoke=keep["invoke"].size,
            group_list=p(keep["group"]), n_group=keep["group"].size,
            unk_cat=p(keep["ucat"]), unk_first_id=p(keep["ufirst"]), unk_count=p(keep["ucount"]),
            n_unk_map=keep["ucat"].size,
            unk_morphs=p(keep["umorphs"]), n_unk_morphs=keep["umorphs"].size // 3)
        if keep["conn"].size != int(self.conn_row) * int(self.conn_col):
            raise ValueError("conn has %d entries, expected %d x %d" % (keep["conn"].size, self.conn_row, self.conn_col))
        return a, keep

    def pack(self) -> np.ndarray:
        """Validated packed blob (host only; what rank 0 broadcasts to the other GPUs)."""
        L = _lib.load()
        a, _keep = self._arrays()
        size = C.c_uint64()
        _lib.check(L.kp_dict_pack(C.byref(a), None, 0, C.byref(size)))
        blob = np.empty(size.value, np.uint8)
        _lib.check(L.kp_dict_pack(C.byref(a), blob.ctypes.data_as(C.c_void_p), size.value, C.byref(size)))
        return blob

    def device_handle(self, device: int = 0):
        """kp_dict* staged on `device` (created once per Dict and device, kept until close())."""
        h = self._handles.get(device)
        if h is None:
            L = _lib.load()
            a, _keep = self._arrays()
            h = C.c_void_p()
            _lib.check(L.kp_dict_create(C.byref(a), device, C.byref(h)))
            self._handles[device] = h
        return h

    def attach_device_blob(self, device_ptr: int, size: int, device: int):
        """Stage `device`'s handle from a packed blob already in HBM there (e.g. the receive buffer of the
        NCCL broadcast from rank 0) instead of from the host arrays (kp_dict_create_from_device_blob:
        validated, then copied into memory the handle owns -- the caller's buffer is not kept).  Only
        allowed while this Dict has no handle on that device yet: a live Tokenizer holds the raw handle."""
        if device in self._handles:
            raise RuntimeError("Dict already staged on device %d; attach_device_blob must come first" % device)
        h = C.c_void_p()
        _lib.check(_lib.load().kp_dict_create_from_device_blob(device_ptr, size, device, C.byref(h)))
        self._handles[device] = h
        return h

    def close(self):
        """Destroys every device handle.  Only call once no Tokenizer built on this Dict is in use."""
        handles, self._handles = self._handles, {}
        for h in handles.values():
            _lib.load().kp_dict_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the reference's `.dict` zip container (kanpyo-dict/src/dict.rs:51-116) ---------------------
    def build(self, f):
        """`Dict::build(&self, f)`: write the six-member zip the reference's tools read."""
        from . import dictfile
        dictfile.save_dict(self, f)

    @classmethod
    def load(cls, f) -> "Dict":
        """`Dict::load(r)`: read a dictionary written by the reference's `ipa-dict-builder` (or by build())."""
        from . import dictfile
        return dictfile.load_dict(f)

    # ---- persistence of the flat form (a cache, not the reference's .dict zip) -------------------
    _NPZ = ("da", "dup_ids", "dup_counts", "morphs", "conn", "char_category", "invoke_list", "group_list", "unk_cat",
            "unk_first_id", "unk_count", "unk_morphs")

    def save_npz(self, path: str):
        kw_blob = np.frombuffer(b"".join(self.keywords), np.uint8)
        kw_off = np.zeros(len(self.keywords) + 1, np.uint64)
        if self.keywords:
            kw_off[1:] = np.cumsum([len(k) for k in self.keywords], dtype=np.uint64)
        extra = {}
        for tag, table in (("f", self.features), ("u", self.unk_features)):
            if table:
                rows, names = table
                off = np.zeros(len(rows) + 1, np.uint64)
                if rows:
                    off[1:] = np.cumsum([len(r) for r in rows], dtype=np.uint64)
                extra[tag + "_off"] = off
                extra[tag + "_flat"] = np.fromiter((v for r in rows for v in r), np.uint32, int(off[-1]))
                blob = "\x00".join(names).encode("utf-8")          # feature strings never contain NUL
                extra[tag + "_names"] = np.frombuffer(blob, np.uint8)
        np.savez(path, **{k: getattr(self, k) for k in self._NPZ},
                 conn_shape=np.array([self.conn_row, self.conn_col], np.uint64),
                 char_class=np.array(self.char_class if self.char_class else [""]), kw_blob=kw_blob, kw_off=kw_off,
                 **extra)

    @classmethod
    def load_npz(cls, path: str) -> "Dict":
        z = np.load(path, allow_pickle=False)
        kw_blob = z["kw_blob"].tobytes()
        kw_off = z["kw_off"]
        keywords = [kw_blob[int(kw_off[i]):int(kw_off[i + 1])] for i in range(len(kw_off) - 1)]

        def table(tag):
            if tag + "_off" not in z:
                return ()
            off, flat = z[tag + "_off"], z[tag + "_flat"]
            rows = [flat[int(off[i]):int(off[i + 1])].tolist() for i in range(len(off) - 1)]
            return rows, z[tag + "_names"].tobytes().decode("utf-8").split("\x00")

        return cls(**{k: z[k] for k in cls._NPZ}, conn_row=int(z["conn_shape"][0]), conn_col=int(z["conn_shape"][1]),
                   char_class=[str(x) for x in z["char_class"]], keywords=keywords, features=table("f"),
                   unk_features=table("u"))

    # ---- feature strings (src/bin/kanpyo.rs:174-197; not read by the hot path) ---------------------
    def token_features(self, token) -> list:
        """`morph_feature_table.morph_features[id-1]` mapped through `name_list` for a Known token, the
        unknown dictionary's table for an Unknown one, [] for EOS."""
        cls, tid = int(token.cls), int(token.id)
        if tid == 0 or cls == 0:
            return []
        rows, names = self.features if cls == 1 else self.unk_features
        return [names[i] for i in rows[tid - 1]]
