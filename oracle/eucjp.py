"""TEST INFRASTRUCTURE (oracle) — WHATWG EUC-JP decoder.

The reference decodes the IPADIC sources with `encoding_rs::EUC_JP`
(kanpyo-dict/src/builder/record.rs:23, char_def.rs:22, unk.rs:19; selected in
kanpyo-dict/src/bin/ipa_dict_builder.rs:41-44).  encoding_rs implements the WHATWG Encoding
Standard, whose `index jis0208` is the CP932 (Windows-31J) mapping, *not* the JIS X 0208 mapping
behind Python's `euc_jp` codec.  The two differ on IPADIC for four characters
(0xA1C1 -> U+FF5E not U+301C, 0xA1C2 -> U+2225 not U+2016, 0xA1DD -> U+FF0D not U+2212,
0xA1F2 -> U+FFE1 not U+00A3; 4908 occurrences), and because token ids are ranks in the byte-wise
sort of the decoded surfaces the difference changes ids.  The jis0208 index is reproduced here by
re-expressing each EUC-JP (lead, trail) pair as Shift_JIS and decoding it with Python's `cp932`.
"""


def _sjis_pair(lead: int, trail: int) -> bytes:
    row = lead - 0xA0          # 1..94
    cell = trail - 0xA0        # 1..94
    s1 = ((row - 1) >> 1) + (0x81 if row <= 62 else 0xC1)
    if row & 1:
        s2 = cell + 0x3F + (1 if cell >= 64 else 0)
    else:
        s2 = cell + 0x9E
    return bytes((s1, s2))


_CACHE: dict = {}


def _pair(lead: int, trail: int) -> str:
    key = (lead << 8) | trail
    ch = _CACHE.get(key)
    if ch is None:
        ch = _sjis_pair(lead, trail).decode("cp932")
        _CACHE[key] = ch
    return ch


def decode(data: bytes) -> str:
    """Decode EUC-JP bytes the way encoding_rs does; raises ValueError on malformed input
    (the reference returns KanpyoError::EncodingError when `had_errors`, record.rs:24-26)."""
    out = []
    i = 0
    n = len(data)
    while i < n:
        c = data[i]
        if c < 0x80:
            # fast path: run of ASCII
            j = i + 1
            while j < n and data[j] < 0x80:
                j += 1
            out.append(data[i:j].decode("ascii"))
            i = j
            continue
        if i + 1 >= n:
            raise ValueError("truncated EUC-JP sequence")
        t = data[i + 1]
        if c == 0x8E:
            if not 0xA1 <= t <= 0xDF:
                raise ValueError("bad half-width katakana trail byte")
            out.append(chr(0xFF61 + t - 0xA1))
            i += 2
            continue
        if c == 0x8F:
            # JIS X 0212 three-byte form; IPADIC never uses it and Python has no WHATWG jis0212
            # index, so refuse rather than guess.
            raise ValueError("JIS X 0212 (0x8F) sequences are not supported by the oracle")
        if not (0xA1 <= c <= 0xFE and 0xA1 <= t <= 0xFE):
            raise ValueError("bad EUC-JP byte pair %02x %02x" % (c, t))
        out.append(_pair(c, t))
        i += 2
    return "".join(out)


def table_94x94():
    """The (lead-0xA1)*94 + (trail-0xA1) -> code point table (0 = unmapped); used to generate the
    product's C++ table (`tools/gen_eucjp_table.py`)."""
    tab = []
    for lead in range(0xA1, 0xFF):
        for trail in range(0xA1, 0xFF):
            try:
                s = _sjis_pair(lead, trail).decode("cp932")
                tab.append(ord(s) if len(s) == 1 else 0)
            except UnicodeDecodeError:
                tab.append(0)
    return tab
