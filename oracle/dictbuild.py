"""TEST INFRASTRUCTURE (oracle) — Python restatement of the reference's dictionary builder.

Token ids returned by the hot path are 1-based ranks in the builder's global sort, so the oracle
has to restate the builder as well as the tokenizer:

  DictionaryBuilder::from_config        kanpyo-dict/src/builder.rs:46-116
  parse_csv / Record (derived Ord)      kanpyo-dict/src/builder/record.rs:5-42
  parse_unk_def / UnkDefRecord          kanpyo-dict/src/builder/unk.rs:8-42
  parse_char_def                        kanpyo-dict/src/builder/char_def.rs:20-99
  parse_matrix_def                      kanpyo-dict/src/builder/matrix_def.rs:17-64
  IndexTable::build                     kanpyo-dict/src/index.rs:16-38
  UnkDict::build                        kanpyo-dict/src/unk_dict.rs:19-57
  MorphFeatureTableBuilder              kanpyo-dict/src/morph_feature.rs:40-92
  da::build_with_ids                    kanpyo-dict/src/trie/da.rs:206-217  (-> da_build.c)

Output: a plain dict of numpy arrays (`OracleDict`), the flat little-endian form of the
reference's `Dict` (kanpyo-dict/src/dict.rs:21-30).
"""
from __future__ import annotations

import csv
import io
import os
import re
from dataclasses import dataclass, field

import numpy as np

from . import eucjp


def _decode(data: bytes, encoding: str) -> str:
    if encoding in ("euc-jp", "eucjp", "euc_jp"):
        return eucjp.decode(data)
    if encoding in ("utf8", "utf-8"):
        return data.decode("utf-8")
    raise ValueError("unsupported encoding %r (reference supports euc-jp|utf8, ipa_dict_builder.rs:13-17)" % encoding)


def _i16_wrap(v: int) -> int:
    """Rust `x as i16` on an integer: two's-complement truncation (builder.rs:64-68)."""
    v &= 0xFFFF
    return v - 0x10000 if v >= 0x8000 else v


def parse_csv(text: str):
    """record.rs:21-42 — csv crate, no headers; fields 0..3 = surface,left,right,cost; rest = features."""
    rows = []
    for rec in csv.reader(io.StringIO(text, newline="")):
        if not rec:
            continue
        rows.append((rec[0], int(rec[1]), int(rec[2]), int(rec[3]), rec[4:]))
    return rows


def _record_key(r):
    # derived Ord on Record (record.rs:5-19): String compares as bytes; Vec<String> lexicographic
    return (r[0].encode("utf-8"), r[1], r[2], r[3], [f.encode("utf-8") for f in r[4]])


_RE_CLASS = re.compile(r"^(\w+)\s+(\d+)\s+(\d+)\s+(\d+)")
_RE_SINGLE = re.compile(r"^(0x[0-9A-F]+)(?:\s+([^#\s]+))(?:\s+([^#\s]+))?")
_RE_RANGE = re.compile(r"^(0x[0-9A-F]+)..(0x[0-9A-F]+)(?:\s+([^#\s]+))(?:\s+([^#\s]+))?")


def parse_char_def(text: str):
    """char_def.rs:31-99: class lines `NAME invoke group length`; single / range code point lines;
    only the FIRST category of a mapping line is used; later lines override earlier ones."""
    char_class: list[str] = []
    category = np.zeros(1 << 16, dtype=np.uint8)
    invoke: list[bool] = []
    group: list[bool] = []
    cc2id: dict[str, int] = {}
    for raw in text.split("\n"):      # BufRead::lines splits on \n and strips a trailing \r
        line = raw.rstrip("\r").strip()
        if line.startswith("#") or not line:
            continue
        m = _RE_CLASS.match(line)
        if m:
            invoke.append(m.group(2) == "1")
            group.append(m.group(3) == "1")
            cc2id[m.group(1)] = len(char_class) & 0xFF
            char_class.append(m.group(1))
            continue
        m = _RE_SINGLE.match(line)
        if m:
            ch = int(m.group(1)[2:], 16)
            category[ch] = cc2id[m.group(2)]
            continue
        m = _RE_RANGE.match(line)
        if m:
            start = int(m.group(1)[2:], 16)
            end = int(m.group(2)[2:], 16)
            category[start:end + 1] = cc2id[m.group(3)]
            continue
        raise ValueError("Invalid char.def format: %s" % line)
    return char_class, category, invoke, group


def parse_matrix_def(text: str):
    """matrix_def.rs:23-64: header `row col`; each line `r c v` -> data[c*row + r] = v."""
    lines = text.split("\n")
    hdr = lines[0].split()
    if len(hdr) != 2:
        raise ValueError("Invalid row and col")
    row, col = int(hdr[0]), int(hdr[1])
    body = np.array(" ".join(lines[1:]).split(), dtype=np.int64)
    if body.size % 3:
        raise ValueError("Invalid matrix value")
    body = body.reshape(-1, 3)
    r, c, v = body[:, 0], body[:, 1], body[:, 2]
    if (r < 0).any() or (c < 0).any() or (r >= row).any() or (c >= col).any():
        raise ValueError("Invalid matrix index")
    if (v < -32768).any() or (v > 32767).any():
        raise ValueError("matrix value out of i16 range")
    data = np.zeros(row * col, dtype=np.int16)
    data[c * row + r] = v.astype(np.int16)      # later duplicates override, as in the loop
    return row, col, data


class FeatureTableBuilder:
    """morph_feature.rs:40-92: interned strings, ids start at 1, name_list[0] = ''."""

    def __init__(self):
        self.ids: dict[str, int] = {}
        self.rows: list[list[int]] = []

    def push(self, feats):
        row = []
        for name in feats:
            i = self.ids.get(name)
            if i is None:
                i = len(self.ids) + 1
                self.ids[name] = i
            row.append(i)
        self.rows.append(row)

    def build(self):
        names = [""] * (len(self.ids) + 1)
        for k, v in self.ids.items():
            names[v] = k
        return self.rows, names


@dataclass
class OracleDict:
    """Flat form of kanpyo_dict::dict::Dict (dict.rs:21-30)."""
    da: np.ndarray                 # int32 [da_len, 2] = (base, check)          da.rs:14-20
    dup_ids: np.ndarray            # int64 [n_dup]   BTreeMap keys (ascending)   index.rs:12
    dup_counts: np.ndarray         # uint64 [n_dup]  BTreeMap values
    morphs: np.ndarray             # int16 [n, 3] = (left_id, right_id, cost)    morph.rs:7-11
    conn_row: int                  # connection.rs:5-9
    conn_col: int
    conn: np.ndarray               # int16 [row*col], data[col_arg*row + row_arg]
    char_class: list               # char_category_def.rs:15-20
    char_category: np.ndarray      # uint8 [65536]
    invoke_list: np.ndarray        # uint8 (bool)
    group_list: np.ndarray         # uint8 (bool)
    unk_cat: np.ndarray            # uint8 [n_map]  BTreeMap<u8,(id,count)> keys  unk_dict.rs:12-16
    unk_first_id: np.ndarray       # int64 [n_map]
    unk_count: np.ndarray          # uint64 [n_map]
    unk_morphs: np.ndarray         # int16 [n_unk, 3]
    # not on the hot path (features are only printed, kanpyo.rs:174-197); kept for text goldens
    keywords: list = field(default_factory=list)       # sorted surfaces incl. duplicates (bytes)
    features: tuple = field(default_factory=tuple)     # (rows, names)
    unk_features: tuple = field(default_factory=tuple)


def build_index(sorted_keywords_bytes, da_build):
    """index.rs:16-38: dedup consecutive equal keywords; id of a key = 1-based position of its
    FIRST occurrence; dup[id] = number of additional occurrences."""
    keys, ids, dup = [], [], {}
    prev_key, prev_no = None, None
    for i, key in enumerate(sorted_keywords_bytes):
        if prev_key is not None and prev_key == key:
            dup[prev_no] = dup.get(prev_no, 0) + 1
            continue
        prev_key, prev_no = key, i + 1
        keys.append(key)
        ids.append(i + 1)
    da = da_build(keys, ids)
    dup_ids = np.array(sorted(dup), dtype=np.int64)
    dup_counts = np.array([dup[k] for k in sorted(dup)], dtype=np.uint64)
    return da, dup_ids, dup_counts


def build_unk(records, char_class):
    """unk_dict.rs:19-57: sort records (derived Ord: category string, ids, cost, features);
    1-based ids in that order; per category (first id, count)."""
    records = sorted(records, key=_record_key)
    morphs = []
    ftb = FeatureTableBuilder()
    cat_map: dict[int, list] = {}
    for morph_id, r in enumerate(records):
        if r[3] > 32767:
            raise ValueError("CostOutOfRange %d" % r[3])
        morphs.append((_i16_wrap(r[1]), _i16_wrap(r[2]), _i16_wrap(r[3])))
        if r[0] not in char_class:
            raise ValueError("CharCategoryNotFound %s" % r[0])
        cat = char_class.index(r[0]) & 0xFF
        ent = cat_map.setdefault(cat, [morph_id + 1, 0])
        ent[1] += 1
        ftb.push(r[4])
    cats = sorted(cat_map)
    return (np.array(morphs, dtype=np.int16).reshape(-1, 3),
            np.array(cats, dtype=np.uint8),
            np.array([cat_map[c][0] for c in cats], dtype=np.int64),
            np.array([cat_map[c][1] for c in cats], dtype=np.uint64),
            ftb.build())


def from_dir(root: str, encoding: str, da_build) -> OracleDict:
    """DictionaryBuilder::from_config (builder.rs:46-116)."""
    records = []
    for name in sorted(os.listdir(root)):           # read_dir order is irrelevant: global sort below
        if os.path.splitext(name)[1] == ".csv":
            with open(os.path.join(root, name), "rb") as f:
                records.extend(parse_csv(_decode(f.read(), encoding)))
    records.sort(key=_record_key)                    # .sorted(), builder.rs:49-53
    morphs = np.empty((len(records), 3), dtype=np.int16)
    keywords = []
    ftb = FeatureTableBuilder()
    for i, r in enumerate(records):
        if r[3] > 32767:
            raise ValueError("Cost is too large: %d" % r[3])   # panic!, builder.rs:59-61
        keywords.append(r[0].encode("utf-8"))
        morphs[i] = (_i16_wrap(r[1]), _i16_wrap(r[2]), _i16_wrap(r[3]))
        ftb.push(r[4])
    with open(os.path.join(root, "matrix.def"), "rb") as f:
        row, col, conn = parse_matrix_def(f.read().decode("utf-8"))   # read as-is (File + lines())
    da, dup_ids, dup_counts = build_index(keywords, da_build)
    with open(os.path.join(root, "char.def"), "rb") as f:
        char_class, category, invoke, group = parse_char_def(_decode(f.read(), encoding))
    with open(os.path.join(root, "unk.def"), "rb") as f:
        unk_records = parse_csv(_decode(f.read(), encoding))
    unk_morphs, unk_cat, unk_first, unk_count, unk_feats = build_unk(unk_records, char_class)
    return OracleDict(
        da=da, dup_ids=dup_ids, dup_counts=dup_counts, morphs=morphs,
        conn_row=row, conn_col=col, conn=conn,
        char_class=char_class, char_category=category,
        invoke_list=np.array(invoke, dtype=np.uint8), group_list=np.array(group, dtype=np.uint8),
        unk_cat=unk_cat, unk_first_id=unk_first, unk_count=unk_count, unk_morphs=unk_morphs,
        keywords=keywords, features=ftb.build(), unk_features=unk_feats)
