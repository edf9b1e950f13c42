"""Shared test helpers: the reference's test fixture dictionary and oracle <-> product conversions."""
import numpy as np


def to_product_dict(od):
    """OracleDict -> kanpyo_b200.Dict over the same arrays."""
    import kanpyo_b200
    return kanpyo_b200.Dict(
        da=od.da, dup_ids=od.dup_ids, dup_counts=od.dup_counts, morphs=od.morphs, conn_row=od.conn_row,
        conn_col=od.conn_col, conn=od.conn, char_category=od.char_category, invoke_list=od.invoke_list,
        group_list=od.group_list, unk_cat=od.unk_cat, unk_first_id=od.unk_first_id, unk_count=od.unk_count,
        unk_morphs=od.unk_morphs, char_class=list(od.char_class), keywords=list(od.keywords),
        features=tuple(od.features) if od.features else (), unk_features=tuple(od.unk_features) if od.unk_features else ())


def reference_fixture_dict(oracle_mod):
    """create_test_dict() of the reference's src/tests.rs:8-108."""
    cat = np.zeros(1 << 16, np.uint8)
    cat[ord("あ"):ord("ん") + 1] = 2      # 'あ'..='ん' -> HIRAGANA
    cat[ord("一"):ord("龥") + 1] = 1      # '一'..='龥' -> KANJI
    return oracle_mod.dict_from_keywords(
        ["テスト", "辞書", "形態素"], morphs=[(0, 0, 1000), (1, 1, 1200), (2, 2, 1100)],
        conn=[0, 100, 200, 100, 0, 100, 200, 100, 0], conn_shape=(3, 3), char_class=("DEFAULT", "KANJI", "HIRAGANA"),
        category=cat, invoke=(False, True, True), group=(False, True, True), unk_map={1: (1, 1), 2: (2, 1)},
        unk_morphs=[(0, 0, 5000), (1, 1, 5000)])


def pack(sentences):
    """list[str] -> (uint8 text, uint64 offsets)."""
    blobs = [s.encode("utf-8") for s in sentences]
    off = np.zeros(len(blobs) + 1, np.uint64)
    if blobs:
        off[1:] = np.cumsum([len(b) for b in blobs], dtype=np.uint64)
    return np.frombuffer(b"".join(blobs), np.uint8), off


def assert_batch_equal(res, o_tok_off, o_tokens, o_cost):
    """Bit-exact comparison of a product BatchResult with the oracle's batch output
    (oracle tokens: int64[n,6] = id, class, position, start, end, byte_len)."""
    assert np.array_equal(res.tok_off, o_tok_off), "token offsets differ"
    assert np.array_equal(res.eos_cost, o_cost), "EOS path costs differ"
    t = res.tokens
    assert len(t) == len(o_tokens)
    if len(t) == 0:
        return
    assert np.array_equal(t["id"].astype(np.int64), o_tokens[:, 0]), "ids differ"
    assert np.array_equal(t["cls"].astype(np.int64), o_tokens[:, 1]), "classes differ"
    assert np.array_equal(t["position"].astype(np.int64), o_tokens[:, 2]), "byte positions differ"
    assert np.array_equal(t["start"].astype(np.int64), o_tokens[:, 3]), "char starts differ"
    assert np.array_equal(t["start"].astype(np.int64) + t["char_len"], o_tokens[:, 4]), "char ends differ"
    # surface byte length: distance to the next token of the same sentence (EOS: len("EOS") = 3)
    nxt = np.empty(len(t), np.int64)
    nxt[:-1] = t["position"][1:]
    nxt[-1] = 0
    bl = np.where(t["cls"] == 0, 3, nxt - t["position"].astype(np.int64))
    assert np.array_equal(bl, o_tokens[:, 5]), "surface byte lengths differ"


def oracle_to_token8(o_tokens):
    """Oracle tokens (int64[n,6] = id, class, position, start, end, byte_len) -> kp_token8 records as the
    device writes them (EOS carries the sentence's char count in its two length fields)."""
    from kanpyo_b200.tokenizer import TOKEN8_DTYPE
    t = np.zeros(len(o_tokens), TOKEN8_DTYPE)
    if len(o_tokens) == 0:
        return t
    cls = o_tokens[:, 1]
    t["id_cls"] = (o_tokens[:, 0] | (cls << 30)).astype(np.uint32)
    eos = cls == 0
    t["byte_len"] = np.where(eos, o_tokens[:, 3] & 0xFFFF, o_tokens[:, 5]).astype(np.uint16)
    t["char_len"] = np.where(eos, o_tokens[:, 3] >> 16, o_tokens[:, 4] - o_tokens[:, 3]).astype(np.uint16)
    return t
