"""Pins the ORACLE (oracle/) against every known-answer test the reference holds for the hot path and
against the README's end-to-end outputs.  CPU only.

    reference test                                         here
    kanpyo-dict/src/trie/da.rs:253-286  test_build_and_search            test_da_build_and_search
    kanpyo-dict/src/trie/da.rs:289-323  test_search_common_prefix        test_da_search_common_prefix
    kanpyo-dict/src/trie/da.rs:326-351  test_build_and_search_multibyte  test_da_build_and_search_multibyte
    kanpyo-dict/src/index.rs:92-150     IndexTable tests                 test_index_*
    kanpyo-dict/src/connection.rs:58-72 test_get                         test_connection_get
    kanpyo-dict/src/builder/matrix_def.rs:70-85 test_parse               test_matrix_def_parse
    src/tests.rs:111-202 (fixture :8-108)                                test_fixture_*
    README.md:73-107                                                     test_readme_*
"""
import json
import os

import numpy as np
import pytest

from helpers import reference_fixture_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _tok(oracle_mod, keywords, **kw):
    return oracle_mod.OracleTokenizer(oracle_mod.dict_from_keywords(keywords, **kw))


def test_da_build_and_search(oracle_mod):
    kws = ["a", "ab", "abc", "abcd", "abcde", "abcdef", "abcdefg", "abcdefgh", "abcdefghi", "abcdefghij"]
    t = _tok(oracle_mod, kws)
    for i, k in enumerate(kws):
        assert t.da_search(k) == i + 1
    for k in ["", "b", "abcdeh", "abcdefghijj"]:
        assert t.da_search(k) is None


def test_da_search_common_prefix(oracle_mod):
    kws = ["早稲田", "早稲田大学", "東京", "東京大学", "東京大学大学院", "東京大学大学院情報理工学研究科",
           "東京大学大学院情報理工学研究科創造情報学専攻", "東京工業大学"]
    t = _tok(oracle_mod, kws)
    assert t.common_prefix("東京大学大学院情報理工学研究科創造情報学専攻", use_dup=False) == [
        (3, 6), (4, 12), (5, 21), (6, 45), (7, 66)]
    assert t.common_prefix("早稲田大学", use_dup=False) == [(1, 9), (2, 15)]
    assert t.common_prefix("大学", use_dup=False) is None


def test_da_build_and_search_multibyte(oracle_mod):
    kws = sorted(["12345", "2345", "１２３", "abc", "ABCD", "あいう", "Ａ"], key=lambda s: s.encode("utf-8"))
    t = _tok(oracle_mod, kws)
    for i, k in enumerate(kws):
        assert t.da_search(k) == i + 1
    for k in ["", "b", "ab", "abcdeh", "abcdefghijj", "あい", "あいうえお"]:
        assert t.da_search(k) is None


def test_index_build_empty(oracle_mod):
    # index.rs:92-95 only asserts the build succeeds: the truncated array has a single node, so a
    # search would index self.0[ROOT_ID] out of bounds (a panic in the reference).
    d = oracle_mod.dict_from_keywords([])
    assert d.da.shape == (1, 2) and len(d.dup_ids) == 0


def test_index_duplicates(oracle_mod):
    t = _tok(oracle_mod, ["apple", "apple", "banana", "banana", "banana", "cherry"],
             morphs=np.zeros((6, 3), np.int16))
    assert t.common_prefix("apple") == [(1, 5), (2, 5)]
    assert t.common_prefix("banana") == [(3, 6), (4, 6), (5, 6)]
    assert t.common_prefix("cherry") == [(6, 6)]


def test_index_not_found(oracle_mod):
    t = _tok(oracle_mod, ["apple", "banana"])
    assert t.common_prefix("cherry") is None


def test_index_common_prefix(oracle_mod):
    t = _tok(oracle_mod, ["東京", "東京大学", "東京大学大学院"])
    assert t.common_prefix("東京大学大学院情報学") == [(1, 6), (2, 12), (3, 21)]


def test_connection_get(oracle_mod):
    t = _tok(oracle_mod, [], conn=[0, 1, 2, 3], conn_shape=(2, 2))
    for i in range(2):
        for j in range(2):
            assert t.conn_get(i, j) == j * 2 + i


def test_matrix_def_parse():
    from oracle import dictbuild
    row, col, data = dictbuild.parse_matrix_def("2 2\n0 0 1\n0 1 2\n1 0 3\n1 1 4\n")
    assert (row, col) == (2, 2)
    assert data.tolist() == [1, 3, 2, 4]


# ---- src/tests.rs fixture -----------------------------------------------------------------------------
@pytest.fixture(scope="module")
def fixture_tok(oracle_mod):
    return oracle_mod.OracleTokenizer(reference_fixture_dict(oracle_mod))


def test_fixture_basic(fixture_tok):                      # src/tests.rs:111-129
    toks, cost = fixture_tok.tokenize("テスト")
    assert toks == [(1, 1, 0, 0, 3, "テスト"), (0, 0, 9, 3, 6, "EOS")]
    assert cost == 1000


def test_fixture_empty_input(fixture_tok):                # src/tests.rs:131-143
    toks, cost = fixture_tok.tokenize("")
    assert toks == [(0, 0, 0, 0, 3, "EOS")]
    assert cost == 0


def test_fixture_unknown_word(fixture_tok):               # src/tests.rs:145-154
    toks, cost = fixture_tok.tokenize("あいうえお")
    assert toks == [(2, 2, 0, 0, 5, "あいうえお"), (0, 0, 15, 5, 8, "EOS")]
    assert cost == 5200
    la = fixture_tok.lattice("あいうえお")
    # four unknown nodes start mid-run at positions no node ends on: dp stays INF, no predecessor
    dead = [i for i in range(len(la["dp"])) if la["dp"][i] == 1 << 30]
    assert len(dead) == 4 and all(la["pre"][i] == -1 for i in dead)


def test_fixture_positions(fixture_tok):                  # src/tests.rs:156-176
    toks, _ = fixture_tok.tokenize("テスト")
    for t in toks:
        if t[1] != 0:
            assert t[3] <= t[4] <= 3


# ---- README end-to-end outputs -------------------------------------------------------------------------
README = {
    "すもももももももものうち": [
        ("すもも", "名詞,一般,*,*,*,*,すもも,スモモ,スモモ"), ("も", "助詞,係助詞,*,*,*,*,も,モ,モ"),
        ("もも", "名詞,一般,*,*,*,*,もも,モモ,モモ"), ("も", "助詞,係助詞,*,*,*,*,も,モ,モ"),
        ("もも", "名詞,一般,*,*,*,*,もも,モモ,モモ"), ("の", "助詞,連体化,*,*,*,*,の,ノ,ノ"),
        ("うち", "名詞,非自立,副詞可能,*,*,*,うち,ウチ,ウチ"), ("EOS", "")],
    "自然言語処理": [
        ("自然", "名詞,形容動詞語幹,*,*,*,*,自然,シゼン,シゼン"), ("言語", "名詞,一般,*,*,*,*,言語,ゲンゴ,ゲンゴ"),
        ("処理", "名詞,サ変接続,*,*,*,*,処理,ショリ,ショリ"), ("EOS", "")],
    "形態素解析": [
        ("形態素", "名詞,一般,*,*,*,*,形態素,ケイタイソ,ケイタイソ"), ("解析", "名詞,サ変接続,*,*,*,*,解析,カイセキ,カイセキ"),
        ("EOS", "")],
}


@pytest.mark.parametrize("text", list(README))
def test_readme_outputs(oracle_tok, text):                # README.md:73-107
    toks, _ = oracle_tok.tokenize(text)
    assert [(t[5], oracle_tok.features(t)) for t in toks] == README[text]


def test_ipadic_shape(oracle_ipadic):
    d = oracle_ipadic
    assert d.morphs.shape == (392126, 3) and len(set(d.keywords)) == 325871
    assert (d.conn_row, d.conn_col) == (1316, 1316)
    assert list(d.char_class) == ["DEFAULT", "SPACE", "KANJI", "SYMBOL", "NUMERIC", "ALPHA", "HIRAGANA", "KATAKANA",
                                  "KANJINUMERIC", "GREEK", "CYRILLIC"]
    assert len(d.unk_morphs) == 40


def test_cfg1_vector(oracle_tok):
    """BASELINE.json configs[0]: 36 bytes, 12 chars, 63 lattice nodes, dp[EOS] = 21245."""
    toks, cost = oracle_tok.tokenize("すもももももももものうち")
    assert cost == 21245
    assert [(t[2], t[3], t[4]) for t in toks] == [(0, 0, 3), (9, 3, 4), (12, 4, 6), (18, 6, 7), (21, 7, 9),
                                                  (27, 9, 10), (30, 10, 12), (36, 12, 15)]
    assert len(oracle_tok.lattice("すもももももももものうち")["nodes"]) == 63


def test_golden_sentences(oracle_tok):
    g = json.load(open(os.path.join(GOLDEN, "ipadic_sentences.json"), encoding="utf-8"))
    for s in g["sentences"]:
        toks, cost = oracle_tok.tokenize(s["text"])
        assert cost == s["cost"], s["text"]
        assert [[t[0], t[1], t[2], t[3], t[4], t[5], oracle_tok.features(t)] for t in toks] == s["tokens"], s["text"]


def test_golden_cfg2_slice(oracle_tok, vocab):
    import hashlib
    from kanpyo_b200 import corpus
    g = json.load(open(os.path.join(GOLDEN, "cfg2_512.json")))
    text, off = corpus.synth_corpus(vocab, g["n_sent"], g["kind"], g["seed"])
    assert corpus.sha256(text) == g["text_sha256"]
    tok_off, tokens, cost, ctr = oracle_tok.tokenize_batch(text, off)
    assert int(tok_off[-1]) == g["n_tokens"] and ctr == g["counters"]
    assert hashlib.sha256(np.ascontiguousarray(tokens[:, :5]).tobytes()).hexdigest() == g["tokens_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(cost).tobytes()).hexdigest() == g["cost_sha256"]
    # multi-threaded oracle (the cpu_baseline leg) gives the same answer
    tok_off2, tokens2, cost2, ctr2 = oracle_tok.tokenize_batch(text, off, threads=4)
    assert np.array_equal(tok_off, tok_off2) and np.array_equal(tokens, tokens2) and np.array_equal(cost, cost2)
    assert ctr == ctr2


def test_token_ids_follow_the_whatwg_euc_jp_decoder():
    """Token ids are ranks in the byte-wise sort of the DECODED surfaces (builder.rs:46-57, record.rs derived Ord), so
    the decoder matters.  The reference decodes with `encoding_rs::EUC_JP` (record.rs:23), which is the WHATWG decoder:
    its jis0208 index is the Windows-31J table, not the JIS X 0208 table behind Python's `euc_jp`.  Five IPADIC records
    (Symbol.csv) move from below every kana surface to above it under the WHATWG table -- '£' (A1F2), '−' (A1DD),
    '−−', and the two '〜' (A1C1) rows -- which lowers the id of every surface in between by 5: すもも is 36164 here,
    36169 under Python's codec (the figure SURVEY.md section 8c derived with a throwaway script).  Both decoders are
    pinned here on the rank of すもも and on the records that differ."""
    import tarfile
    from oracle import eucjp, oracle
    surfaces_jis, surfaces_whatwg = [], []
    with tarfile.open(oracle.IPADIC_TARBALL, "r:*") as tar:
        for m in tar.getmembers():
            if not m.name.endswith(".csv"):
                continue
            raw = tar.extractfile(m).read()
            for a, b in ((surfaces_jis, raw.decode("euc_jp")), (surfaces_whatwg, eucjp.decode(raw))):
                a.extend(line.split(",", 1)[0].encode("utf-8") for line in b.split("\n") if line)
    assert len(surfaces_jis) == len(surfaces_whatwg) == 392126
    moved = sorted((a.decode(), b.decode()) for a, b in zip(surfaces_jis, surfaces_whatwg) if a != b)
    assert len(moved) == 21                                       # 21 records hold one of the four re-mapped characters
    target = "すもも".encode("utf-8")
    rank = lambda surfaces: sum(s < target for s in surfaces) + 1   # noqa: E731  (1-based id of the first すもも record)
    assert rank(surfaces_jis) == 36169 and rank(surfaces_whatwg) == 36164
    crossing = sorted(a for a, b in moved if a.encode() < target <= b.encode())
    assert crossing == sorted(["£", "−", "−−", "〜", "〜"])
    od = oracle.load_ipadic()
    assert od.keywords[36164 - 1] == target and od.keywords[36164 - 2] != target
