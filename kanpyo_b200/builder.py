"""DictionaryBuilder — host-side mirror of the reference's IPADIC builder, producing the flat `Dict`
the CUDA path stages to HBM.

    reference (Rust)                                              here
    -----------------------------------------------------------   ------------------------------------
    DictionaryBuilder::from_config   kanpyo-dict/src/builder.rs:46-116      DictionaryBuilder.build()
    Config{root_path,encoding,...}   kanpyo-dict/src/builder/config.rs:6-27  DictionaryBuilder(root, encoding)
    parse_csv / Record: Ord          kanpyo-dict/src/builder/record.rs:5-42  _read_records / _sort_records
    parse_unk_def, UnkDict::build    builder/unk.rs:8-42, unk_dict.rs:19-57  _build_unk
    parse_char_def                   builder/char_def.rs:20-99               _parse_char_def
    parse_matrix_def                 builder/matrix_def.rs:17-64             _parse_matrix_def
    IndexTable::build                kanpyo-dict/src/index.rs:16-38          _build_index
    da::build_with_ids               kanpyo-dict/src/trie/da.rs:206-217      kp_da_build (C ABI, kp_dictbuild.cpp)

Token ids returned by `Tokenizer::tokenize` are 1-based ranks in this builder's global sort, so the
sort keys and the decoder below are part of the tokenizer's observable behaviour.

The source may be a directory or the MeCab IPADIC tarball (read in place, nothing is extracted).
`ipadic()` builds the vendored `third_party/mecab-ipadic` tarball once and caches the flat arrays
under `kanpyo_b200/_cache/`.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import csv
import io
import os
import re
import tarfile

import numpy as np

from . import _lib
from .dict import Dict

_HERE = os.path.dirname(os.path.abspath(__file__))
IPADIC_TARBALL = os.path.join(os.path.dirname(_HERE), "third_party", "mecab-ipadic",
                              "mecab-ipadic-2.7.0-20070801.tar.gz")
IPADIC_SHA256 = "b62f527d881c504576baed9c6ef6561554658b175ce6ae0096a60307e49e3523"
CACHE_DIR = os.path.join(_HERE, "_cache")
CACHE_VERSION = 3


class BuilderError(ValueError):
    """Counterpart of KanpyoError for the build path (kanpyo-dict/src/error.rs:6-54)."""


# encoding_rs::EUC_JP (record.rs:23) is the WHATWG decoder: its jis0208 index is the Windows-31J
# table, which differs from the JIS X 0208 table behind Python's `euc_jp` codec in these cells.
# Every left-hand code point is reachable from exactly one EUC-JP two-byte cell, so patching after
# the decode is exact.
_WHATWG_FIXUPS = {0x301C: 0xFF5E,   # A1C1 WAVE DASH          -> FULLWIDTH TILDE
                  0x2016: 0x2225,   # A1C2 DOUBLE VERTICAL    -> PARALLEL TO
                  0x2212: 0xFF0D,   # A1DD MINUS SIGN         -> FULLWIDTH HYPHEN-MINUS
                  0x00A2: 0xFFE0,   # A1F1 CENT SIGN          -> FULLWIDTH CENT SIGN
                  0x00A3: 0xFFE1,   # A1F2 POUND SIGN         -> FULLWIDTH POUND SIGN
                  0x00AC: 0xFFE2}   # A2CC NOT SIGN           -> FULLWIDTH NOT SIGN


def _decode(data: bytes, encoding: str, what: str) -> str:
    enc = encoding.lower().replace("_", "-")
    try:
        if enc in ("euc-jp", "eucjp"):
            return data.decode("euc_jp").translate(_WHATWG_FIXUPS)
        if enc in ("utf8", "utf-8"):
            return data.decode("utf-8")
    except UnicodeDecodeError as e:   # KanpyoError::EncodingError (record.rs:24-26)
        raise BuilderError("failed to decode %s as %s: %s" % (what, encoding, e)) from None
    raise BuilderError("unsupported encoding %r (euc-jp | utf8, ipa_dict_builder.rs:13-17)" % encoding)


def _wrap_i16(v: np.ndarray) -> np.ndarray:
    """`x as i16` (builder.rs:64-68): two's-complement truncation."""
    return v.astype(np.int64).astype(np.uint16).view(np.int16) if v.size else np.zeros(0, np.int16)


class _Source:
    """Files of a dictionary source tree: a directory or a .tar(.gz) read in place."""

    def __init__(self, root: str):
        self.files = {}
        if os.path.isdir(root):
            for name in os.listdir(root):
                p = os.path.join(root, name)
                if os.path.isfile(p):
                    self.files[name] = p
            self._tar = None
        else:
            self._tar = tarfile.open(root, "r:*")
            for m in self._tar.getmembers():
                if m.isfile():
                    self.files[os.path.basename(m.name)] = m

    def read(self, name: str) -> bytes:
        if name not in self.files:
            raise BuilderError("missing dictionary source file %s" % name)
        ent = self.files[name]
        if self._tar is None:
            with open(ent, "rb") as f:
                return f.read()
        return self._tar.extractfile(ent).read()

    def csv_names(self):
        return sorted(n for n in self.files if n.endswith(".csv"))   # glob *.csv (builder.rs:27-44)


_USIZE = re.compile(r"^\+?[0-9]+$")      # what `str::parse::<usize>` accepts (a leading '+' is legal, '-', blanks, '_' are not)
_I64 = re.compile(r"^[+-]?[0-9]+$")       # `str::parse::<i64>`


def _parse_usize(tok: str) -> int:
    if not _USIZE.match(tok) or int(tok) >= 1 << 64:
        raise ValueError(tok)
    return int(tok)


def _parse_i64(tok: str) -> int:
    if not _I64.match(tok) or not -(1 << 63) <= int(tok) < 1 << 63:
        raise ValueError(tok)
    return int(tok)


def _read_records(text: str, what: str):
    """parse_csv (record.rs:21-42) with the `csv` crate's default reader: no header row, '"' quoting with '""'
    escapes (a quoted field may hold commas and line breaks), empty lines skipped, and every record must have
    as many fields as the first one; surface, left_id (usize), right_id (usize), cost (i64), then user data."""
    rows, width = [], None
    reader = csv.reader(io.StringIO(text, newline=""), strict=True)
    try:
        for f in reader:
            if not f:
                continue
            if width is None:
                width = len(f)
            elif len(f) != width:
                raise BuilderError("%s:%d: found record with %d fields, but the previous record has %d fields"
                                   % (what, reader.line_num, len(f), width))
            if len(f) < 4:
                raise BuilderError("%s:%d: expected at least 4 fields" % (what, reader.line_num))
            try:
                rows.append((f[0].encode("utf-8"), _parse_usize(f[1]), _parse_usize(f[2]), _parse_i64(f[3]),
                             tuple(x.encode("utf-8") for x in f[4:])))
            except ValueError:
                raise BuilderError("%s:%d: left_id / right_id must parse as usize and cost as i64 (record.rs:35-37)"
                                   % (what, reader.line_num)) from None
    except csv.Error as e:
        raise BuilderError("%s:%d: malformed CSV: %s" % (what, reader.line_num, e)) from None
    return rows


class _Interner:
    """MorphFeatureTableBuilder (morph_feature.rs:40-92): ids from 1 in first-seen order, name 0 = ''."""

    def __init__(self):
        self.ids = {}
        self.rows = []

    def push(self, feats):
        ids = self.ids
        row = []
        for f in feats:
            i = ids.get(f)
            if i is None:
                i = ids[f] = len(ids) + 1
            row.append(i)
        self.rows.append(row)

    def finish(self):
        names = [b""] * (len(self.ids) + 1)
        for k, v in self.ids.items():
            names[v] = k
        return self.rows, [n.decode("utf-8") for n in names]


def _parse_matrix_def(data: bytes):
    """parse_matrix_def (matrix_def.rs:17-64): `row col` header, then `r c v` -> data[c*row + r]."""
    head, _, body = data.partition(b"\n")
    dims = head.split()
    if len(dims) != 2:
        raise BuilderError("matrix.def: first line must be `row col`")
    row, col = int(dims[0]), int(dims[1])
    vals = np.array(body.split(), dtype=np.int64)
    if vals.size % 3:
        raise BuilderError("matrix.def: every line must be `row col value`")
    r, c, v = vals[0::3], vals[1::3], vals[2::3]
    if r.size and (r.min() < 0 or c.min() < 0 or r.max() >= row or c.max() >= col):
        raise BuilderError("matrix.def: index outside %d x %d" % (row, col))
    if v.size and (v.min() < -32768 or v.max() > 32767):
        raise BuilderError("matrix.def: value outside the i16 range (matrix_def.rs:54)")
    conn = np.zeros(row * col, np.int16)
    conn[c * row + r] = v          # numpy keeps the LAST assignment for repeated cells, like the loop
    return row, col, conn


def _hex_upper(tok: str):
    """`0x[0-9A-F]+` (char_def.rs:39-47): upper-case hex only; -> (value, chars consumed) or None."""
    if not tok.startswith("0x"):
        return None
    j = 2
    while j < len(tok) and tok[j] in "0123456789ABCDEF":
        j += 1
    return (int(tok[2:j], 16), j) if j > 2 else None


def _parse_char_def(text: str):
    """parse_char_def (char_def.rs:20-99).  Class line `NAME invoke group length`; mapping line
    `0xXXXX[..0xYYYY] CLASS [CLASS2 ...]` of which only the first class is used; later lines win."""
    names, invoke, group = [], [], []
    ident = {}
    table = np.zeros(1 << 16, np.uint8)
    for raw in text.split("\n"):
        line = raw.strip()
        if not line or line.startswith("#"):
            continue
        tok = line.split()
        if (len(tok) >= 4 and tok[0].replace("_", "a").isalnum() and tok[1].isdigit() and tok[2].isdigit()
                and tok[3][:1].isdigit()):
            ident[tok[0]] = len(names) & 0xFF
            names.append(tok[0])
            invoke.append(tok[1] == "1")
            group.append(tok[2] == "1")
            continue
        h = _hex_upper(tok[0])
        if h is None or len(tok) < 2 or tok[1].startswith("#"):
            raise BuilderError("char.def: cannot parse %r" % line)
        lo, used = h
        hi = lo
        rest = tok[0][used:]
        if rest:
            h2 = _hex_upper(rest[2:]) if len(rest) > 2 else None     # `..` is two arbitrary chars in the regex
            if h2 is None or h2[1] != len(rest) - 2:
                raise BuilderError("char.def: cannot parse %r" % line)
            hi = h2[0]
        cls = tok[1].split("#")[0]
        if cls not in ident:
            raise BuilderError("char.def: unknown class %s" % cls)
        if hi >= table.size:
            raise BuilderError("char.def: code point 0x%X outside the 65536-entry table" % hi)
        table[lo:hi + 1] = ident[cls]
    return names, table, np.array(invoke, np.uint8), np.array(group, np.uint8)


def _build_index(keywords):
    """IndexTable::build (index.rs:16-38) over the sorted surfaces: one trie key per distinct surface,
    id = 1-based position of its first record, dup[id] = further records with the same surface."""
    n = len(keywords)
    if n == 0:
        firsts = np.zeros(0, np.int64)
    else:
        new = np.ones(n, bool)
        new[1:] = [keywords[i] != keywords[i - 1] for i in range(1, n)]
        firsts = np.nonzero(new)[0].astype(np.int64)
    ends = np.append(firsts[1:], n)
    extra = ends - firsts - 1
    keys = [keywords[i] for i in firsts.tolist()]
    ids = firsts + 1
    has = extra > 0
    return keys, ids, ids[has].astype(np.int64), extra[has].astype(np.uint64)


def da_build(keys, ids) -> np.ndarray:
    """da::build_with_ids (da.rs:206-217) -> int32 [len, 2] (base, check), through kp_da_build."""
    L = _lib.load()
    blob = np.frombuffer(b"".join(keys), np.uint8) if keys else np.zeros(0, np.uint8)
    off = np.zeros(len(keys) + 1, np.uint64)
    if keys:
        off[1:] = np.cumsum(np.fromiter((len(k) for k in keys), np.uint64, len(keys)))
    ids = np.ascontiguousarray(ids, np.int64)
    out = C.c_void_p()
    n = C.c_uint64()
    _lib.check(L.kp_da_build(blob.ctypes.data_as(C.c_void_p) if blob.size else None, off.ctypes.data_as(C.c_void_p),
                             len(keys), ids.ctypes.data_as(C.c_void_p) if len(keys) else None, C.byref(out), C.byref(n)))
    try:
        return np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int32)), shape=(n.value, 2)).copy()
    finally:
        L.kp_da_free(out)


class DictionaryBuilder:
    """`DictionaryBuilder::from_config(Config::new(root, encoding))` (builder.rs:46, config.rs:19-27)."""

    MATRIX_DEF, CHAR_DEF, UNK_DEF = "matrix.def", "char.def", "unk.def"     # config.rs:22-26

    def __init__(self, root: str, encoding: str = "euc-jp"):
        self.root = root
        self.encoding = encoding

    def build(self) -> Dict:
        src = _Source(self.root)
        records = []
        for name in src.csv_names():
            records.extend(_read_records(_decode(src.read(name), self.encoding, name), name))
        # `.sorted()` on Record's derived Ord (builder.rs:49-53, record.rs:5-19): surface bytes, then the
        # three integers, then the user-data strings; tuples of (bytes, int, int, int, tuple[bytes]) order the same
        records.sort()
        n = len(records)
        cost = np.fromiter((r[3] for r in records), np.int64, n)
        if n and cost.max() > 32767:
            raise BuilderError("Cost is too large: %d (builder.rs:59-61)" % int(cost.max()))
        morphs = np.empty((n, 3), np.int16)
        morphs[:, 0] = _wrap_i16(np.fromiter((r[1] for r in records), np.int64, n))
        morphs[:, 1] = _wrap_i16(np.fromiter((r[2] for r in records), np.int64, n))
        morphs[:, 2] = _wrap_i16(cost)
        keywords = [r[0] for r in records]
        feats = _Interner()
        for r in records:
            feats.push(r[4])
        row, col, conn = _parse_matrix_def(src.read(self.MATRIX_DEF))
        keys, ids, dup_ids, dup_counts = _build_index(keywords)
        da = da_build(keys, ids)
        names, table, invoke, group = _parse_char_def(_decode(src.read(self.CHAR_DEF), self.encoding, self.CHAR_DEF))
        unk = self._build_unk(_read_records(_decode(src.read(self.UNK_DEF), self.encoding, self.UNK_DEF),
                                            self.UNK_DEF), names)
        return Dict(da=da, dup_ids=dup_ids, dup_counts=dup_counts, morphs=morphs, conn_row=row, conn_col=col, conn=conn,
                    char_category=table, invoke_list=invoke, group_list=group, unk_cat=unk[0], unk_first_id=unk[1],
                    unk_count=unk[2], unk_morphs=unk[3], char_class=names, keywords=keywords, features=feats.finish(),
                    unk_features=unk[4])

    @staticmethod
    def _build_unk(records, class_names):
        """UnkDict::build (unk_dict.rs:19-57): records sorted by (category NAME, ids, cost, features);
        1-based ids in that order; per class the first id and the number of consecutive ids."""
        records = sorted(records)
        first, count = {}, {}
        feats = _Interner()
        morphs = np.zeros((len(records), 3), np.int16)
        for i, r in enumerate(records):
            if r[3] > 32767:
                raise BuilderError("unk.def: cost %d is out of range" % r[3])
            name = r[0].decode("utf-8")
            if name not in class_names:
                raise BuilderError("unk.def: class %s is not defined in char.def" % name)
            c = class_names.index(name) & 0xFF
            first.setdefault(c, i + 1)
            count[c] = count.get(c, 0) + 1
            morphs[i] = _wrap_i16(np.array(r[1:4], np.int64))
            feats.push(r[4])
        cats = sorted(first)
        return (np.array(cats, np.uint8), np.array([first[c] for c in cats], np.int64),
                np.array([count[c] for c in cats], np.uint64), morphs, feats.finish())


def _file_sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


_IPADIC = None


def ipadic(rebuild: bool = False) -> Dict:
    """The vendored MeCab IPADIC 2.7.0-20070801 built by DictionaryBuilder (cached flat arrays)."""
    global _IPADIC
    if _IPADIC is not None and not rebuild:
        return _IPADIC
    cache = os.path.join(CACHE_DIR, "ipadic_v%d.npz" % CACHE_VERSION)
    if os.path.exists(cache) and not rebuild:
        try:
            _IPADIC = Dict.load_npz(cache)
            return _IPADIC
        except Exception:
            pass
    digest = _file_sha256(IPADIC_TARBALL)
    if digest != IPADIC_SHA256:
        raise BuilderError("vendored IPADIC tarball has sha256 %s, expected %s" % (digest, IPADIC_SHA256))
    d = DictionaryBuilder(IPADIC_TARBALL, "euc-jp").build()
    os.makedirs(CACHE_DIR, exist_ok=True)
    tmp = cache + ".tmp.npz"
    d.save_npz(tmp)
    os.replace(tmp, cache)
    _IPADIC = d
    return d
