"""Reader / writer for the reference's `.dict` container, so the CUDA path can consume dictionaries
produced by the reference's `ipa-dict-builder` (and hand its own to the reference).

    reference (Rust)                                         here
    Dict::load / Dict::build      kanpyo-dict/src/dict.rs:51-116        load_dict / save_dict
    Morphs                        kanpyo-dict/src/morph.rs:61-92        i64 count, then (i16,i16,i16) little endian
    ConnectionTable               kanpyo-dict/src/connection.rs:28-51   usize row, usize col, i16 data
    DoubleArray                   kanpyo-dict/src/trie/da.rs:219-246    usize len, then (i32 base, i32 check)
    IndexTable                    kanpyo-dict/src/index.rs:57-84        DoubleArray, u64 n, then (isize id, usize count)
    UnkDict                       kanpyo-dict/src/unk_dict.rs:61-99     usize n, (u8, isize, usize)*, Morphs, MorphFeatureTable
    CharCategoryDef               kanpyo-dict/src/char_category_def.rs:41-58   bincode 2, standard config
    MorphFeatureTable             kanpyo-dict/src/morph_feature.rs:20-38       bincode 2, standard config

The container is a zip archive with the six members `morph.dict, morph_feature.dict, connection.dict,
index.dict, chardef.dict, unk.dict` (deflate).  bincode's standard configuration is little endian
with variable-length integers: a value below 251 is one byte; otherwise a marker byte 251 / 252 /
253 followed by the value as u16 / u32 / u64; `Vec<T>` and `String` are a length followed by the
elements; `u8` and `bool` are single bytes.

Format parity is pinned only by round trips and hand-derived byte vectors (the reference has no
fixture file and its released dictionary needs the network); SURVEY.md 8f marks this "parity
unpinned" against real files.
"""
from __future__ import annotations

import io
import struct
import zipfile

import numpy as np

from .dict import Dict

MEMBERS = ("morph.dict", "morph_feature.dict", "connection.dict", "index.dict", "chardef.dict", "unk.dict")


class DictFormatError(ValueError):
    pass


# ---- bincode 2, config::standard() ------------------------------------------------------------------
def _put_varint(out: bytearray, v: int):
    if v < 251:
        out.append(v)
    elif v < 1 << 16:
        out.append(251)
        out += struct.pack("<H", v)
    elif v < 1 << 32:
        out.append(252)
        out += struct.pack("<I", v)
    else:
        out.append(253)
        out += struct.pack("<Q", v)


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.p = data, 0

    def take(self, n: int) -> bytes:
        if self.p + n > len(self.d):
            raise DictFormatError("truncated dictionary member")
        b = self.d[self.p:self.p + n]
        self.p += n
        return b

    def varint(self) -> int:
        m = self.take(1)[0]
        if m < 251:
            return m
        if m == 251:
            return struct.unpack("<H", self.take(2))[0]
        if m == 252:
            return struct.unpack("<I", self.take(4))[0]
        if m == 253:
            return struct.unpack("<Q", self.take(8))[0]
        raise DictFormatError("unsupported bincode integer marker %d" % m)

    def u64(self) -> int:
        return struct.unpack("<Q", self.take(8))[0]

    def i64(self) -> int:
        return struct.unpack("<q", self.take(8))[0]


def encode_feature_table(table) -> bytes:
    """MorphFeatureTable { morph_features: Vec<Vec<u32>>, name_list: Vec<String> }."""
    rows, names = table if table else ([], [""])
    out = bytearray()
    _put_varint(out, len(rows))
    for r in rows:
        _put_varint(out, len(r))
        for v in r:
            _put_varint(out, int(v))
    _put_varint(out, len(names))
    for s in names:
        b = s.encode("utf-8")
        _put_varint(out, len(b))
        out += b
    return bytes(out)


def decode_feature_table(r: _Reader):
    rows = []
    for _ in range(r.varint()):
        rows.append([r.varint() for _ in range(r.varint())])
    names = [r.take(r.varint()).decode("utf-8") for _ in range(r.varint())]
    return rows, names


def encode_chardef(char_class, category, invoke, group) -> bytes:
    """CharCategoryDef { char_class: Vec<String>, char_category: Vec<u8>, invoke_list: Vec<bool>, group_list: Vec<bool> }."""
    out = bytearray()
    _put_varint(out, len(char_class))
    for s in char_class:
        b = s.encode("utf-8")
        _put_varint(out, len(b))
        out += b
    cat = np.ascontiguousarray(category, np.uint8)
    _put_varint(out, cat.size)
    out += cat.tobytes()
    for flags in (invoke, group):
        f = np.ascontiguousarray(flags, np.uint8)
        _put_varint(out, f.size)
        out += (f != 0).astype(np.uint8).tobytes()
    return bytes(out)


def decode_chardef(data: bytes):
    r = _Reader(data)
    names = [r.take(r.varint()).decode("utf-8") for _ in range(r.varint())]
    cat = np.frombuffer(r.take(r.varint()), np.uint8).copy()
    flags = []
    for _ in range(2):
        f = np.frombuffer(r.take(r.varint()), np.uint8).copy()
        if f.size and f.max() > 1:
            raise DictFormatError("chardef.dict: bool out of range")
        flags.append(f)
    return names, cat, flags[0], flags[1]


# ---- fixed little-endian sections -----------------------------------------------------------------------
def _encode_morphs(m) -> bytes:
    m = np.ascontiguousarray(m, "<i2").reshape(-1, 3)
    return struct.pack("<q", len(m)) + m.tobytes()


def _decode_morphs(r: _Reader) -> np.ndarray:
    n = r.i64()
    if n < 0:
        raise DictFormatError("negative morph count")
    return np.frombuffer(r.take(6 * n), "<i2").reshape(n, 3).astype(np.int16)


def save_dict(d: Dict, f):
    """Dict::build (dict.rs:51-69).  `f`: path or binary file object."""
    with zipfile.ZipFile(f, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr("morph.dict", _encode_morphs(d.morphs))
        z.writestr("morph_feature.dict", encode_feature_table(d.features))
        conn = np.ascontiguousarray(d.conn, "<i2")
        z.writestr("connection.dict", struct.pack("<QQ", int(d.conn_row), int(d.conn_col)) + conn.tobytes())
        da = np.ascontiguousarray(d.da, "<i4").reshape(-1, 2)
        dup = np.empty((len(d.dup_ids), 2), "<i8")
        dup[:, 0] = d.dup_ids
        dup[:, 1] = np.asarray(d.dup_counts).astype(np.int64)
        z.writestr("index.dict", struct.pack("<Q", len(da)) + da.tobytes() + struct.pack("<Q", len(dup)) + dup.tobytes())
        z.writestr("chardef.dict", encode_chardef(d.char_class, d.char_category, d.invoke_list, d.group_list))
        unk = bytearray(struct.pack("<Q", len(d.unk_cat)))
        for c, first, count in zip(np.asarray(d.unk_cat).tolist(), np.asarray(d.unk_first_id).tolist(),
                                   np.asarray(d.unk_count).tolist()):
            unk += struct.pack("<BqQ", int(c), int(first), int(count))
        unk += _encode_morphs(d.unk_morphs)
        unk += encode_feature_table(d.unk_features)
        z.writestr("unk.dict", bytes(unk))


def load_dict(f) -> Dict:
    """Dict::load (dict.rs:70-116).  `f`: path, bytes or binary file object."""
    if isinstance(f, (bytes, bytearray)):
        f = io.BytesIO(f)
    try:
        z = zipfile.ZipFile(f, "r")
    except zipfile.BadZipFile as e:
        raise DictFormatError("not a zip archive: %s" % e) from None
    with z:
        missing = [m for m in MEMBERS if m not in z.namelist()]
        if missing:
            raise DictFormatError("missing members: %s" % ", ".join(missing))
        morphs = _decode_morphs(_Reader(z.read("morph.dict")))
        features = decode_feature_table(_Reader(z.read("morph_feature.dict")))
        r = _Reader(z.read("connection.dict"))
        row, col = r.u64(), r.u64()
        conn = np.frombuffer(r.take(2 * row * col), "<i2").astype(np.int16)
        r = _Reader(z.read("index.dict"))
        n = r.u64()
        da = np.frombuffer(r.take(8 * n), "<i4").reshape(n, 2).astype(np.int32)
        n = r.u64()
        dup = np.frombuffer(r.take(16 * n), "<i8").reshape(n, 2)
        names, cat, invoke, group = decode_chardef(z.read("chardef.dict"))
        r = _Reader(z.read("unk.dict"))
        n = r.u64()
        ucat, ufirst, ucount = [], [], []
        for _ in range(n):
            c, first, count = struct.unpack("<BqQ", r.take(17))
            ucat.append(c)
            ufirst.append(first)
            ucount.append(count)
        unk_morphs = _decode_morphs(r)
        unk_features = decode_feature_table(r)
    return Dict(da=da, dup_ids=dup[:, 0].astype(np.int64), dup_counts=dup[:, 1].astype(np.uint64), morphs=morphs,
                conn_row=row, conn_col=col, conn=conn, char_category=cat, invoke_list=invoke, group_list=group,
                unk_cat=np.array(ucat, np.uint8), unk_first_id=np.array(ufirst, np.int64),
                unk_count=np.array(ucount, np.uint64), unk_morphs=unk_morphs, char_class=names, features=features,
                unk_features=unk_features)
