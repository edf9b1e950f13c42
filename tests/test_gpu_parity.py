"""Parity of the CUDA path (through the C ABI) with the oracle: bit-exact tokens (id, class, byte position,
char start/end, surface bytes) and dp[EOS], plus node-level lattice parity.  Needs a GPU: -m gpu."""
import json
import os

import numpy as np
import pytest

from helpers import assert_batch_equal, pack, reference_fixture_dict, to_product_dict

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _check_sentences(gpu, orc, sentences):
    text, off = pack(sentences)
    res = gpu.tokenize_batch_bytes(text, off)
    o_off, o_tok, o_cost, _ = orc.tokenize_batch(text, off)
    assert_batch_equal(res, o_off, o_tok, o_cost)
    return res


# ---- the reference's own known-answer vectors, on the device trie ---------------------------------------
def _small(oracle_mod, keywords, **kw):
    import kanpyo_b200
    od = oracle_mod.dict_from_keywords(keywords, **kw)
    return kanpyo_b200.Tokenizer(to_product_dict(od), device=0), oracle_mod.OracleTokenizer(od)


def test_da_search_common_prefix_vectors(oracle_mod):     # da.rs:289-323
    kws = ["早稲田", "早稲田大学", "東京", "東京大学", "東京大学大学院", "東京大学大学院情報理工学研究科",
           "東京大学大学院情報理工学研究科創造情報学専攻", "東京工業大学"]
    g, _ = _small(oracle_mod, kws)
    assert g.common_prefix("東京大学大学院情報理工学研究科創造情報学専攻", expand_dup=False) == [
        (3, 6), (4, 12), (5, 21), (6, 45), (7, 66)]
    assert g.common_prefix("早稲田大学", expand_dup=False) == [(1, 9), (2, 15)]
    assert g.common_prefix("大学", expand_dup=False) is None


def test_index_vectors(oracle_mod):                       # index.rs:98-150
    g, _ = _small(oracle_mod, ["apple", "apple", "banana", "banana", "banana", "cherry"],
                  morphs=np.zeros((6, 3), np.int16))
    assert g.common_prefix("apple") == [(1, 5), (2, 5)]
    assert g.common_prefix("banana") == [(3, 6), (4, 6), (5, 6)]
    g, _ = _small(oracle_mod, ["apple", "banana"])
    assert g.common_prefix("cherry") is None
    g, _ = _small(oracle_mod, ["東京", "東京大学", "東京大学大学院"])
    assert g.common_prefix("東京大学大学院情報学") == [(1, 6), (2, 12), (3, 21)]


def test_exact_keys_found(oracle_mod):                    # da.rs:253-286, 326-351 via common-prefix hits
    kws = sorted(["12345", "2345", "１２３", "abc", "ABCD", "あいう", "Ａ"], key=lambda s: s.encode("utf-8"))
    g, o = _small(oracle_mod, kws)
    for i, k in enumerate(kws):
        hits = g.common_prefix(k, expand_dup=False)
        assert hits is not None and hits[-1] == (i + 1, len(k.encode("utf-8")))
        assert hits == o.common_prefix(k, use_dup=False)
    for k in ["", "b", "あい"]:
        assert g.common_prefix(k) == o.common_prefix(k)


def test_empty_dictionary_is_graceful(oracle_mod):
    """index.rs:92-95 builds an empty index; the reference would panic on search, the device path finds nothing."""
    g, _ = _small(oracle_mod, [], unk_map={0: (1, 1)}, unk_morphs=[(0, 0, 7)])
    assert g.common_prefix("anything") is None
    toks = g.tokenize("ab")
    assert [(t.id, int(t.cls), t.surface) for t in toks] == [(1, 2, "a"), (1, 2, "b"), (0, 0, "EOS")]


# ---- src/tests.rs fixture --------------------------------------------------------------------------------
def test_reference_fixture(oracle_mod):
    import kanpyo_b200
    from kanpyo_b200 import TokenClass
    od = reference_fixture_dict(oracle_mod)
    g = kanpyo_b200.Tokenizer(to_product_dict(od), device=0)
    o = oracle_mod.OracleTokenizer(od)
    toks, cost = g.tokenize_with_cost("テスト")
    assert [(t.id, t.cls, t.position, t.start, t.end, t.surface) for t in toks] == [
        (1, TokenClass.Known, 0, 0, 3, "テスト"), (0, TokenClass.Dummy, 9, 3, 6, "EOS")]
    assert cost == 1000
    toks, cost = g.tokenize_with_cost("")
    assert [(t.id, t.cls, t.position, t.start, t.end, t.surface) for t in toks] == [(0, TokenClass.Dummy, 0, 0, 3, "EOS")]
    assert cost == 0
    toks, cost = g.tokenize_with_cost("あいうえお")
    assert [(t.id, t.cls, t.position, t.start, t.end, t.surface) for t in toks] == [
        (2, TokenClass.Unknown, 0, 0, 5, "あいうえお"), (0, TokenClass.Dummy, 15, 5, 8, "EOS")]
    assert cost == 5200
    _check_sentences(g, o, ["テスト", "", "あいうえお", "辞書テスト形態素", "テスト辞書あい形態素うえ", "xyz", "漢字テスト"])
    # no unknown entry for DEFAULT: 'x' has no node at all, the path is cut (dead nodes, truncated back-trace)
    for s in ["xテスト", "テストx", "xx", "テxスト"]:
        _check_sentences(g, o, [s])


# ---- IPADIC ----------------------------------------------------------------------------------------------
def test_cfg1(gpu_tok, oracle_tok):
    """BASELINE.json configs[0]."""
    s = "すもももももももものうち"
    toks, cost = gpu_tok.tokenize_with_cost(s)
    otoks, ocost = oracle_tok.tokenize(s)
    assert cost == ocost == 21245
    assert [(t.id, int(t.cls), t.position, t.start, t.end, t.surface) for t in toks] == otoks
    assert [t.surface for t in toks] == ["すもも", "も", "もも", "も", "もも", "の", "うち", "EOS"]


def test_golden_sentences(gpu_tok):
    g = json.load(open(os.path.join(GOLDEN, "ipadic_sentences.json"), encoding="utf-8"))
    for s in g["sentences"]:
        toks, cost = gpu_tok.tokenize_with_cost(s["text"])
        assert cost == s["cost"], s["text"]
        got = [[t.id, int(t.cls), t.position, t.start, t.end, t.surface] for t in toks]
        assert got == [t[:6] for t in s["tokens"]], s["text"]
    # and as one batch
    res = gpu_tok.tokenize_batch([s["text"] for s in g["sentences"]])
    for toks, s in zip(res, g["sentences"]):
        assert [[t.id, int(t.cls), t.position, t.start, t.end, t.surface] for t in toks] == [t[:6] for t in s["tokens"]]


def test_readme_cli_output(gpu_tok):                      # README.md:73-107 through print_tokens' format
    from test_oracle_pins import README
    for text, expect in README.items():
        out = gpu_tok.format_tokens(gpu_tok.tokenize(text))
        assert out == "".join("%s\t%s\n" % (s, f) for s, f in expect)


def test_graphviz_dump(gpu_tok):                          # src/graphviz.rs:30-163 over kp_lattice_dump
    from kanpyo_b200.graphviz import graphviz
    text = "すもももももももものうち"
    dot = graphviz(gpu_tok, text)
    lines = dot.splitlines()
    assert lines[0] == "graph lattice {" and lines[1] == "dpi=48;" and lines[-1] == "}"
    assert any(ln.startswith('0 [label="BOS", shape=ellipse, color=blue, peripheries=2]') for ln in lines)
    assert any(ln.startswith('1 [label="EOS", shape=ellipse, color=blue, peripheries=2]') for ln in lines)
    bold = [ln for ln in lines if "style=bold" in ln]
    assert len(bold) == len(gpu_tok.tokenize(text))            # BOS->t1, t1->t2, ..., t7->EOS
    assert not any("shape=diamond" in ln for ln in lines)      # off-path unknown nodes are hidden
    assert 'label="すもも\n名詞/一般/すもも/スモモ/スモモ\n' in dot
    full = graphviz(gpu_tok, text, dpi=96, full_state=True)
    n_nodes = len(gpu_tok.lattice(text))
    assert full.splitlines()[1] == "dpi=96;"
    assert sum(1 for ln in full.splitlines() if " [label=" in ln and " -- " not in ln) == n_nodes
    text2 = "Tシャツを3枚買ったABC"                             # unknown nodes on (3, ABC) and off (BC, C) the path
    hidden, shown = graphviz(gpu_tok, text2), graphviz(gpu_tok, text2, full_state=True)
    assert "shape=diamond" in shown and "shape=diamond" not in hidden
    assert 'color=red, peripheries=2]' in hidden               # best-path unknown nodes stay visible


def test_lattice_node_parity(gpu_tok, oracle_tok):
    """Lattice{nodes} order + dp/pre of every node (lattice.rs:101-154), including dead nodes."""
    for s in ["すもももももももものうち", "Tシャツを3枚買ったABC", "\U0001F600の犬", "", "カタカナカタカナ", "あ" * 40,
              "東京都に住んでいます。"]:
        la = gpu_tok.lattice(s)
        ol = oracle_tok.lattice(s)
        on = ol["nodes"]          # kind, id, byte_pos, char_pos, end_char_pos, left, right, cost, byte_len
        assert len(la) == len(on), s
        assert np.array_equal(la["cls"], on[:, 0]) and np.array_equal(la["id"], on[:, 1]), s
        assert np.array_equal(la["byte_pos"], on[:, 2]) and np.array_equal(la["char_pos"], on[:, 3]), s
        assert np.array_equal(la["end_char"], on[:, 4]), s
        assert np.array_equal(la["left_id"], on[:, 5]) and np.array_equal(la["right_id"], on[:, 6]), s
        assert np.array_equal(la["cost"], on[:, 7]), s
        odp = np.where(ol["dp"] == np.iinfo(np.int64).min, np.iinfo(np.int32).min, ol["dp"])
        assert np.array_equal(la["dp"].astype(np.int64), odp), s
        assert np.array_equal(la["pre"].astype(np.int64), ol["pre"]), s


@pytest.mark.parametrize("kind,n", [("cfg2", 4096), ("cfg3", 4096), ("cfg4", 24)])
def test_corpus_parity(gpu_tok, oracle_tok, vocab, kind, n):
    from kanpyo_b200 import corpus
    text, off = corpus.synth_corpus(vocab, n, kind)
    res = gpu_tok.tokenize_batch_bytes(text, off)
    o_off, o_tok, o_cost, ctr = oracle_tok.tokenize_batch(text, off, threads=8)
    assert_batch_equal(res, o_off, o_tok, o_cost)
    c = gpu_tok.counters()
    assert (c["bytes"], c["chars"], c["nodes"], c["tokens"]) == (ctr["B"], ctr["C"], ctr["N"], ctr["T"])


@pytest.mark.parametrize("n", [13000, 26000, 34000])
def test_batch_size_does_not_change_results(gpu_tok, oracle_tok, vocab, n):
    """kp_launch_viterbi and kp_launch_backtrace_count pick their lanes per sentence from the batch size (sweep: 32
    below 12 000 sentences, 16 below 30 000, 8 above; back-trace: 32 up to 25 000, 16 up to 50 000, 8 above): every
    combination must give the reference's tokens and costs (the small batches and the 65 536-sentence batch of the
    other tests cover 32 / 32 and 8 / 8; these cover 16 / 32, 16 / 16 and 8 / 16)."""
    from kanpyo_b200 import corpus
    text, off = corpus.synth_corpus(vocab, n, "cfg2")
    res = gpu_tok.tokenize_batch_bytes(text, off)
    o_off, o_tok, o_cost, _ = oracle_tok.tokenize_batch(text, off, threads=os.cpu_count() or 8)
    assert_batch_equal(res, o_off, o_tok, o_cost)


def test_first_character_paths(gpu_tok, oracle_tok):
    """The trie walk starts from a first-character table (1-3 byte characters); 4-byte characters, NUL bytes
    and characters no key starts with take its other paths (kp_dict.cu, kp_lattice_count)."""
    sents = ["𠮷野家で𩸽を食べた", "\x00あ\x00い", "\x00", "😀😀😀", "aあ𠮷\x7f\u0080\u07ff\u0800\uffff", "\ud7ff\ue000",
             "ヴァイオリンとヴィオラ", "ゟ", "東京都" + "\U0010ffff" + "に住む"]
    _check_sentences(gpu_tok, oracle_tok, sents)


def test_work_counters(gpu_tok, oracle_tok, vocab):
    """P, P_ok, E of the counting kernels equal the oracle's exact counts (the roofline's inputs)."""
    from kanpyo_b200 import corpus
    text, off = corpus.synth_corpus(vocab, 1024, "cfg2")
    gpu_tok.set_count_work(True)
    try:
        gpu_tok.tokenize_batch_bytes(text, off)
        c = gpu_tok.counters()
    finally:
        gpu_tok.set_count_work(False)
    _, _, _, ctr = oracle_tok.tokenize_batch(text, off, threads=8)
    assert (c["probes"], c["probes_ok"], c["pairs"]) == (ctr["P"], ctr["P_ok"], ctr["E"])


def test_golden_cfg2_checksum(gpu_tok, vocab):
    import hashlib
    from kanpyo_b200 import corpus
    g = json.load(open(os.path.join(GOLDEN, "cfg2_512.json")))
    text, off = corpus.synth_corpus(vocab, g["n_sent"], g["kind"], g["seed"])
    res = gpu_tok.tokenize_batch_bytes(text, off)
    t = res.tokens
    packed = np.stack([t["id"].astype(np.int64), t["cls"].astype(np.int64), t["position"].astype(np.int64),
                       t["start"].astype(np.int64), t["start"].astype(np.int64) + t["char_len"]], axis=1)
    assert hashlib.sha256(np.ascontiguousarray(packed).tobytes()).hexdigest() == g["tokens_sha256"]
    assert hashlib.sha256(np.ascontiguousarray(res.eos_cost).tobytes()).hexdigest() == g["cost_sha256"]


def test_chunking_is_invisible(gpu_ipadic, oracle_tok, vocab):
    """Small chunk size (many device passes) gives the same packed result as one pass."""
    import kanpyo_b200
    from kanpyo_b200 import corpus
    text, off = corpus.synth_corpus(vocab, 600, "cfg2")
    t = kanpyo_b200.Tokenizer(gpu_ipadic, device=0)
    t.set_chunk_bytes(10_000)
    res = t.tokenize_batch_bytes(text, off)
    assert t.profile()["chunks"] > 5
    o_off, o_tok, o_cost, _ = oracle_tok.tokenize_batch(text, off, threads=4)
    assert_batch_equal(res, o_off, o_tok, o_cost)
    t.close()


def test_ragged_and_empty(gpu_tok, oracle_tok):
    sents = ["", "", "あ", "", "犬" * 300, "", "a", ""]
    _check_sentences(gpu_tok, oracle_tok, sents)
    _check_sentences(gpu_tok, oracle_tok, [""])
    res = gpu_tok.tokenize_batch_bytes(b"", np.zeros(1, np.uint64))
    assert len(res.tokens) == 0 and res.tok_off.tolist() == [0]


def test_long_unknown_runs(gpu_tok, oracle_tok):
    """Unknown-word grouping cap of 1024 chars (lattice.rs:55,79-82) and buckets wider than a warp."""
    sents = ["ー" * 1030, "ア" * 2100, "a" * 1500 + "犬" + "1" * 40, "ｱ" * 1024, "ｱ" * 1025, "9" * 1023 + "a"]
    _check_sentences(gpu_tok, oracle_tok, sents)


def test_invalid_utf8_rejected(gpu_tok):
    import kanpyo_b200
    for bad in [b"\xff", b"\xe3\x81", b"a\x80b", b"\xc0\xaf", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\xe3\x81\x82\xe3",
                b"\x80", b"\xe3\x81\x82\x80", b"\xf0\x9f\x98", b"\xe0\x9f\xbf", b"\xf0\x8f\xbf\xbf", b"ab\xc2", b"\xbf\xe3\x81\x82"]:
        with pytest.raises(kanpyo_b200.KanpyoB200Error) as e:
            gpu_tok.tokenize_batch_bytes(bad, np.array([0, len(bad)], np.uint64))
        assert e.value.status == -4
    # a multi-byte char split by a sentence boundary is invalid in both halves
    b = "あ".encode("utf-8")
    with pytest.raises(kanpyo_b200.KanpyoB200Error):
        gpu_tok.tokenize_batch_bytes(b, np.array([0, 1, 3], np.uint64))


def test_utf8_validation_matches_python(gpu_tok):
    """The ABI's validity decision equals Python's strict UTF-8 decoder on random byte strings."""
    import kanpyo_b200
    rng = np.random.default_rng(11)
    pool = [b"a", b"\xe3\x81\x82", b"\xc3\xa9", b"\xf0\x9f\x98\x80", b"\x80", b"\xe3", b"\xf0\x9f", b"\xc0", b"\xed\xa0\x80",
            b"\xef\xbf\xbf", b"\xf4\x8f\xbf\xbf", b"\xf4\x90\x80\x80", b"\xe0\xa0\x80", b"\xe0\x9f\x80", b" "]
    n_bad = 0
    for _ in range(400):
        b = b"".join(pool[k] for k in rng.integers(0, len(pool), rng.integers(1, 8)))
        try:
            b.decode("utf-8")
            valid = True
        except UnicodeDecodeError:
            valid = False
        try:
            gpu_tok.tokenize_batch_bytes(b, np.array([0, len(b)], np.uint64))
            accepted = True
        except kanpyo_b200.KanpyoB200Error as e:
            assert e.status == -4
            accepted = False
        assert accepted == valid, b
        n_bad += not valid
    assert 50 < n_bad < 390


def test_device_resident_batch(gpu_tok, oracle_tok, vocab):
    """kp_tokenize_batch_device: text/offsets in HBM, result left in HBM."""
    import torch
    from kanpyo_b200 import corpus
    text, off = corpus.synth_corpus(vocab, 2000, "cfg2")
    d_text = torch.from_numpy(text.copy()).cuda()
    d_off = torch.from_numpy(off.astype(np.int64)).cuda()
    torch.cuda.synchronize()
    r = gpu_tok.tokenize_batch_device(d_text.data_ptr(), d_off.data_ptr(), len(off) - 1, 0, int(text.size))
    got = gpu_tok.copy_device_result(r)
    o_off, o_tok, o_cost, _ = oracle_tok.tokenize_batch(text, off, threads=8)
    assert_batch_equal(got, o_off, o_tok, o_cost)


def test_idempotent_and_sharding_invariant(gpu_tok, vocab):
    """Size-independent properties at a larger size: same answer twice; same answer when the batch is split."""
    from kanpyo_b200 import corpus
    text, off = corpus.synth_corpus(vocab, 20000, "cfg2", seed=7)
    a = gpu_tok.tokenize_batch_bytes(text, off)
    b = gpu_tok.tokenize_batch_bytes(text, off)
    assert np.array_equal(a.tokens, b.tokens) and np.array_equal(a.eos_cost, b.eos_cost)
    parts = corpus.shard_by_bytes(off, 3)
    toks, costs = [], []
    for s0, s1 in parts:
        sub = gpu_tok.tokenize_batch_bytes(text[int(off[s0]):int(off[s1])], off[s0:s1 + 1] - off[s0])
        toks.append(sub.tokens)
        costs.append(sub.eos_cost)
    assert np.array_equal(np.concatenate(toks), a.tokens) and np.array_equal(np.concatenate(costs), a.eos_cost)
    # every sentence's tokens tile its bytes exactly and end with EOS
    t = a.tokens
    last = a.tok_off[1:].astype(np.int64) - 1
    assert (t["cls"][last] == 0).all()
    assert np.array_equal(t["position"][last].astype(np.uint64), np.diff(off))


def test_sharded_tokenizer_single_rank(gpu_ipadic, oracle_tok, vocab):
    """kanpyo_b200.sharded on one rank over NCCL: dictionary blob round trip through the broadcast
    path (device blob -> kp_dict_create_from_device_blob, checksum-validated) and the gather's
    world-size-1 path.  The caller's Dict (and the session tokenizer built on it) stay usable."""
    import torch
    import torch.distributed as dist
    import kanpyo_b200
    from kanpyo_b200 import corpus, sharded
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda:0"))
    try:
        st = sharded.ShardedTokenizer(gpu_ipadic, device=0)
        text, off = corpus.synth_corpus(vocab, 512, "cfg3")
        res = st.tokenize_global(text, off)
        o_off, o_tok, o_cost, _ = oracle_tok.tokenize_batch(text, off)
        assert_batch_equal(res, o_off, o_tok, o_cost)
        assert st.dict is not gpu_ipadic
        # kp_gather_* (the gather behind the C ABI) with a world of one: block assembly + compaction kernel
        from kanpyo_b200.tokenizer import result8_to_batch
        d_text = torch.from_numpy(text.copy()).cuda()
        d_off = torch.from_numpy(off.astype(np.int64)).cuda()
        torch.cuda.synchronize()
        r = st.tokenizer.tokenize_batch_device8(d_text.data_ptr(), d_off.data_ptr(), len(off) - 1, 0, int(text.size))
        ng = sharded.NcclGather(0, len(off) - 1, int(r.n_tokens) + 5)
        g = ng.gather(r)
        assert int(g.n_sent) == len(off) - 1 and ng.last_ms() > 0
        assert_batch_equal(result8_to_batch(ng.to_host(g), off), o_off, o_tok, o_cost)
        small = sharded.NcclGather(0, len(off) - 1, 10)
        with pytest.raises(kanpyo_b200.KanpyoB200Error) as e:      # a shard beyond the capacity is refused, never truncated
            small.gather(r)
        assert e.value.status == -6
        ng.close()
        small.close()
    finally:
        dist.destroy_process_group()


# ---- full-size parity: BASELINE.json configs[1..3] at (or near) their stated sizes ------------------------
def _full_check(tok, oracle_tok, text, off):
    res = tok.tokenize_batch_bytes(text, off)
    o_off, o_tok, o_cost, ctr = oracle_tok.tokenize_batch(text, off, threads=os.cpu_count() or 8)
    assert_batch_equal(res, o_off, o_tok, o_cost)
    c = tok.counters()
    assert (c["bytes"], c["chars"], c["nodes"], c["tokens"]) == (ctr["B"], ctr["C"], ctr["N"], ctr["T"])
    return res


def test_full_cfg2_65536(gpu_tok, oracle_tok, vocab):
    """configs[1]: the benched batch itself, all 65 536 sentences, both device paths."""
    from kanpyo_b200 import corpus
    text, off = corpus.synth_corpus(vocab, 65536, "cfg2")
    for path in ("auto", "pipeline"):
        gpu_tok.set_path(path)
        try:
            _full_check(gpu_tok, oracle_tok, text, off)
        finally:
            gpu_tok.set_path("auto")


def test_full_cfg4_4096x4096(gpu_tok, oracle_tok, vocab):
    """configs[3]: 4096 sentences of 4096 chars: 4096 concurrent long chains, 32 lanes per sentence in the sweep,
    sentence lengths beyond the last bin of the length histogram."""
    from kanpyo_b200 import corpus
    text, off = corpus.synth_corpus(vocab, 4096, "cfg4")
    _full_check(gpu_tok, oracle_tok, text, off)


def test_large_cfg3_multi_chunk(gpu_ipadic, oracle_tok, vocab):
    """configs[2] shape at 327 680 sentences (69.6 MB): two device passes at the default 64 MiB chunk size."""
    import kanpyo_b200
    from kanpyo_b200 import corpus
    text, off = corpus.synth_corpus(vocab, 327680, "cfg3")
    t = kanpyo_b200.Tokenizer(gpu_ipadic, device=0)
    try:
        _full_check(t, oracle_tok, text, off)
        assert t.profile()["chunks"] == 2
    finally:
        t.close()


# ---- compact records, the asynchronous queue, single-process shards ----------------------------------------
def test_compact_records_match(gpu_tok, oracle_tok, oracle_mod, vocab):
    """kp_tokenize_batch8 + host expansion == kp_tokenize_batch == oracle, including truncated and empty paths."""
    import kanpyo_b200
    from kanpyo_b200 import corpus
    from kanpyo_b200.tokenizer import result8_to_batch
    from helpers import oracle_to_token8
    text, off = corpus.synth_corpus(vocab, 3000, "cfg3")
    r8 = gpu_tok.tokenize_batch8_bytes(text, off)
    o_off, o_tok, o_cost, _ = oracle_tok.tokenize_batch(text, off, threads=8)
    assert np.array_equal(r8[1], oracle_to_token8(o_tok)), "kp_token8 records differ from the oracle's"
    assert_batch_equal(result8_to_batch(r8, off), o_off, o_tok, o_cost)
    od = reference_fixture_dict(oracle_mod)
    g = kanpyo_b200.Tokenizer(to_product_dict(od), device=0)
    o = oracle_mod.OracleTokenizer(od)
    text, off = pack(["テスト", "", "xテスト", "テストx", "xx", "テxスト", "あいうえお", "x", "辞書テスト形態素"])
    r8 = g.tokenize_batch8_bytes(text, off)
    o_off, o_tok, o_cost, _ = o.tokenize_batch(text, off)
    assert np.array_equal(r8[1], oracle_to_token8(o_tok))
    assert_batch_equal(result8_to_batch(r8, off), o_off, o_tok, o_cost)
    g.close()


def test_queue_overlapped_batches(gpu_ipadic, oracle_tok, vocab):
    """kp_queue_*: several batches in flight on two contexts give each batch's own result."""
    from kanpyo_b200 import corpus
    from kanpyo_b200.tokenizer import Queue
    q = Queue(gpu_ipadic, device=0, depth=2)
    try:
        batches = [corpus.synth_corpus(vocab, n, kind, seed=100 + i)
                   for i, (n, kind) in enumerate([(3000, "cfg2"), (10, "cfg4"), (1, "cfg2"), (2500, "cfg3"), (700, "cfg2")])]
        tickets = []
        results = {}
        for i, (text, off) in enumerate(batches):
            tickets.append(q.submit(text, off))
            if i >= 1:                           # depth 2: wait for the batch before the previous one
                results[i - 1] = q.wait(tickets[i - 1])
        results[len(batches) - 1] = q.wait(tickets[-1])
        for i, (text, off) in enumerate(batches):
            o_off, o_tok, o_cost, _ = oracle_tok.tokenize_batch(text, off, threads=8)
            assert_batch_equal(results[i], o_off, o_tok, o_cost)
        import kanpyo_b200
        with pytest.raises(kanpyo_b200.KanpyoB200Error):      # a ticket never issued
            q.wait_raw(99)
    finally:
        q.close()


@pytest.mark.parametrize("all_devices", [False, True])
def test_shards_single_process(gpu_ipadic, oracle_tok, vocab, all_devices):
    """kp_shards_*: byte-balanced sentence ranges over the GPUs of one process; the host result lands at global
    offsets, the gathered result on devices[0].  With one device the same code runs without NCCL; with all
    devices (skipped below two) the dictionary travels by ncclBroadcast and the tokens by ncclSend / ncclRecv."""
    import torch
    from kanpyo_b200 import corpus, sharded
    n = torch.cuda.device_count()
    if all_devices and n < 2:
        pytest.skip("needs at least two GPUs")
    sh = sharded.Shards(gpu_ipadic, devices=list(range(n)) if all_devices else [0])
    try:
        for kind, ns in (("cfg2", 5000), ("cfg3", 37), ("cfg2", 1)):
            text, off = corpus.synth_corpus(vocab, ns, kind, seed=5)
            o_off, o_tok, o_cost, _ = oracle_tok.tokenize_batch(text, off, threads=8)
            assert_batch_equal(sh.tokenize(text, off), o_off, o_tok, o_cost)
            assert_batch_equal(sh.tokenize_gather(text, off), o_off, o_tok, o_cost)
        tm = sh.times()
        assert tm["call_ms"] > 0 and (tm["dict_broadcast_ms"] > 0) == (all_devices and n > 1)
        res = sh.tokenize(b"", np.zeros(1, np.uint64))
        assert len(res.tokens) == 0 and res.tok_off.tolist() == [0]
    finally:
        sh.close()


def test_two_tokenizers_share_one_dict(gpu_ipadic, oracle_tok, vocab):
    """A Dict keeps one handle per device for its whole life: building further tokenizers, queues or a
    sharded group on it never invalidates a live tokenizer."""
    import kanpyo_b200
    from kanpyo_b200 import corpus
    from kanpyo_b200.tokenizer import Queue
    text, off = corpus.synth_corpus(vocab, 400, "cfg2", seed=3)
    o = oracle_tok.tokenize_batch(text, off)
    a = kanpyo_b200.Tokenizer(gpu_ipadic, device=0)
    b = kanpyo_b200.Tokenizer(gpu_ipadic, device=0)
    q = Queue(gpu_ipadic, device=0, depth=1)
    assert_batch_equal(a.tokenize_batch_bytes(text, off), o[0], o[1], o[2])
    b.close()
    q.close()
    assert_batch_equal(a.tokenize_batch_bytes(text, off), o[0], o[1], o[2])
    a.close()


# ---- SURVEY 8f rows through the CUDA path -------------------------------------------------------------------
def test_dict_container_to_gpu_tokens(gpu_ipadic, oracle_tok, vocab, tmp_path):
    """`.dict` zip container (kanpyo-dict/src/dict.rs:51-116): Dict.build -> Dict.load -> Tokenizer on the GPU ->
    the oracle's tokens and feature strings."""
    import kanpyo_b200
    from kanpyo_b200 import corpus
    from test_oracle_pins import README
    path = str(tmp_path / "ipadic.dict")
    with open(path, "wb") as f:
        gpu_ipadic.build(f)
    with open(path, "rb") as f:
        loaded = kanpyo_b200.Dict.load(f)
    t = kanpyo_b200.Tokenizer(loaded, device=0)
    try:
        text, off = corpus.synth_corpus(vocab, 2000, "cfg3", seed=9)
        o_off, o_tok, o_cost, _ = oracle_tok.tokenize_batch(text, off, threads=8)
        assert_batch_equal(t.tokenize_batch_bytes(text, off), o_off, o_tok, o_cost)
        for s, expect in README.items():
            assert t.format_tokens(t.tokenize(s)) == "".join("%s\t%s\n" % (a, b) for a, b in expect)
    finally:
        t.close()


def test_graphviz_matches_oracle(gpu_tok, oracle_tok):
    """Dot text of the device lattice == the oracle's restatement of src/graphviz.rs over the oracle's lattice,
    byte for byte, in both views."""
    from kanpyo_b200.graphviz import graphviz
    from oracle import graphviz as ograph
    for s in ["すもももももももものうち", "Tシャツを3枚買ったABC", "", "東京都に住んでいます。", "\U0001F600の犬", "カタカナ語"]:
        for full in (False, True):
            assert graphviz(gpu_tok, s, dpi=72, full_state=full) == ograph.graphviz(oracle_tok, s, 72, full), (s, full)


# ---- the two device paths: fused per-sentence kernel (default where a sentence fits) and the pipeline -----
EDGE_SENTENCES = ["", "あ", "すもももももももものうち", "Tシャツを3枚買ったABC", "\U0001F600の犬", "ｶﾀｶﾅとカタカナと12345と hello world",
                  "東京都に住んでいます。", "ー" * 300, "a" * 500, "犬" * 200, "ア" * 1030, "9" * 1023 + "a", "\x00あ\x00い",
                  "𠮷野家で𩸽を食べた", "ヴァイオリンとヴィオラ", " ", "。。。", "1" * 60 + "犬" + "ア" * 40 + "abc" * 30]


@pytest.mark.parametrize("path", ["fused", "pipeline"])
def test_both_paths_on_edge_cases(gpu_ipadic, oracle_tok, oracle_mod, path):
    import kanpyo_b200
    t = kanpyo_b200.Tokenizer(gpu_ipadic, device=0)
    t.set_path(path)
    try:
        _check_sentences(t, oracle_tok, EDGE_SENTENCES)
        fused = t.profile()["fused_sentences"]
        # short sentences run in the fused kernel; the 1030-char run and the 1024-char one exceed every class
        assert (fused > 10 and fused < len(EDGE_SENTENCES)) if path == "fused" else fused == 0
        for s in EDGE_SENTENCES:
            _check_sentences(t, oracle_tok, [s])
    finally:
        t.close()
    # the reference's fixture: no unknown entry for DEFAULT -> dead nodes, cut and empty paths
    od = reference_fixture_dict(oracle_mod)
    g = kanpyo_b200.Tokenizer(to_product_dict(od), device=0)
    g.set_path(path)
    try:
        _check_sentences(g, oracle_mod.OracleTokenizer(od),
                         ["テスト", "", "あいうえお", "辞書テスト形態素", "テスト辞書あい形態素うえ", "xyz", "漢字テスト", "xテスト", "テストx",
                          "xx", "テxスト", "x"])
    finally:
        g.close()


def test_fused_path_takes_the_short_sentences(gpu_tok, oracle_tok, vocab):
    """On the cfg2 / cfg3 corpora nearly every sentence fits a size class of the fused kernel; the rest goes
    through the pipeline in the same call, and the packed result is the oracle's either way.  `auto` takes
    the fused kernel for batches of up to 3584 sentences (one round trip up to 64), the pipeline above."""
    from kanpyo_b200 import corpus
    for kind, n, path in (("cfg2", 8000, "fused"), ("cfg3", 8000, "fused"), ("cfg2", 3000, "auto"), ("cfg3", 40, "auto")):
        text, off = corpus.synth_corpus(vocab, n, kind, seed=21)
        gpu_tok.set_path(path)
        try:
            res = gpu_tok.tokenize_batch_bytes(text, off)
        finally:
            gpu_tok.set_path("auto")
        p = gpu_tok.profile()
        o_off, o_tok, o_cost, ctr = oracle_tok.tokenize_batch(text, off, threads=os.cpu_count() or 8)
        assert_batch_equal(res, o_off, o_tok, o_cost)
        c = gpu_tok.counters()
        assert (c["bytes"], c["chars"], c["nodes"], c["tokens"]) == (ctr["B"], ctr["C"], ctr["N"], ctr["T"])
        assert p["fused_sentences"] > 0.95 * n, p


def test_fuzz_both_paths(gpu_ipadic, oracle_tok, vocab):
    """Seeded random sentences glued from dictionary words and awkward pieces (long same-class runs, digits, ASCII,
    4-byte characters, NUL, half-width kana, symbols), with byte lengths spread around the fused kernel's class limits
    (192 ... 1536) so that size classes, overflow into the largest class and the hand-over to the pipeline all occur;
    both device paths must give the oracle's tokens and costs, as one batch and in slices of 1, 7 and 64 sentences
    (the one-round-trip path)."""
    import kanpyo_b200
    rng = np.random.default_rng(20261017)
    words = [w.decode("utf-8") for w in vocab.words[:4000]]
    pieces = ["ー" * 40, "ア" * 25, "123456789" * 3, "abc def ", "𠮷", "😀", "\x00", "ｶﾀｶﾅ", "。", "、", "・・・", "一二三", "ＡＢＣ", "々",
              "する" * 12, "こと" * 10, "の" * 30, "東京都" * 8, "http://a.b/c?d=1", "　", "\t", "ゟ", "Привет", "αβγ"]
    sents = []
    for k in range(1500):
        target = int(rng.choice([0, 3, 30, 150, 190, 200, 250, 262, 315, 330, 440, 460, 760, 780, 1500, 1560, 2500]))
        s = ""
        while len(s.encode("utf-8")) < target:
            s += pieces[int(rng.integers(len(pieces)))] if rng.random() < 0.35 else words[int(rng.integers(len(words)))]
        sents.append(s)
    text, off = pack(sents)
    o_off, o_tok, o_cost, _ = oracle_tok.tokenize_batch(text, off, threads=os.cpu_count() or 8)
    t = kanpyo_b200.Tokenizer(gpu_ipadic, device=0)
    try:
        for path in ("fused", "pipeline", "auto"):
            t.set_path(path)
            assert_batch_equal(t.tokenize_batch_bytes(text, off), o_off, o_tok, o_cost)
            if path == "fused":
                p = t.profile()
                assert 0.5 * len(sents) < p["fused_sentences"] < len(sents), p    # the > 1536-byte ones go to the pipeline
        t.set_path("auto")
        for width in (1, 7, 64):
            for s0 in range(0, 448, width):
                s1 = s0 + width
                sub = t.tokenize_batch_bytes(text[int(off[s0]):int(off[s1])], off[s0:s1 + 1] - off[s0])
                a, b = int(o_off[s0]), int(o_off[s1])
                assert_batch_equal(sub, o_off[s0:s1 + 1] - o_off[s0], o_tok[a:b], o_cost[s0:s1])
    finally:
        t.close()
