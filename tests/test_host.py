"""CPU-only tests of the host side: the C-ABI library loads and exports what include/kanpyo_b200.h
declares, the product's dictionary builder reproduces the oracle's arrays, the packed blob, and the
multi-GPU host logic (sharding + gather) under gloo with world_size 2.  No compute call is made here
(the compute entry points have no CPU fallback)."""
import ctypes as C
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

from helpers import reference_fixture_dict, to_product_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "kanpyo_b200.h"), encoding="utf-8").read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from kanpyo_b200 import _lib
    L = _lib.load()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libkanpyo_b200.so does not export %s" % n
        assert n in _lib.SYMBOLS, "ctypes binding misses %s" % n
    assert sorted(_lib.SYMBOLS) == names, "binding declares symbols the header does not"
    # ... with as many parameters as the header's prototypes
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "kanpyo_b200.h"), encoding="utf-8").read(), flags=re.S)
    for m in re.finditer(r"\b(kp_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", text):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else args.count(",") + 1
        assert len(_lib.SYMBOLS[m.group(1)][1]) == n, "%s: ctypes binding has %d parameters, the header %d" % (
            m.group(1), len(_lib.SYMBOLS[m.group(1)][1]), n)
    assert L.kp_abi_version() == 2
    assert L.kp_strerror(-2).decode().startswith("CUDA error or no usable device")


def test_library_does_not_depend_on_torch_or_oracle():
    from kanpyo_b200 import _lib
    out = subprocess.run(["ldd", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "libcudart" not in out, out


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "kanpyo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src and "libkanpyo_oracle" not in src, f


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import kanpyo_b200
    from oracle import oracle
    d = to_product_dict(reference_fixture_dict(oracle))
    with pytest.raises(kanpyo_b200.KanpyoB200Error) as e:
        kanpyo_b200.Tokenizer(d, device=0)
    assert e.value.status == -2      # KP_ERR_CUDA: no fallback


def test_da_build_matches_oracle(oracle_mod):
    from kanpyo_b200 import builder
    rng = np.random.default_rng(7)
    keysets = [[b"a", b"ab", b"abc"], [], [b"\xe3\x81\x82"], [b"hello", b"help", b"world", b"wor"]]
    alphabet = [bytes([c]) for c in range(1, 256)]
    keysets.append(sorted({b"".join(rng.choice(alphabet, rng.integers(1, 9))) for _ in range(3000)}))
    for keys in keysets:
        keys = sorted(keys)
        ids = list(range(1, len(keys) + 1))
        a = builder.da_build(keys, ids)
        b = oracle_mod.da_build(keys, ids)
        assert np.array_equal(a, b)


def test_da_build_reference_vectors():
    """da.rs:253-286 / 326-351: every key is found with id i+1 by walking the product-built array."""
    from kanpyo_b200 import builder
    for kws in (["a", "ab", "abc", "abcd", "abcde", "abcdef", "abcdefg", "abcdefgh", "abcdefghi", "abcdefghij"],
                sorted(["こんにちは", "世界", "すもも", "もも", "電気通信大学", "東京都"], key=lambda s: s.encode())):
        keys = [k.encode("utf-8") for k in kws]
        da = builder.da_build(keys, list(range(1, len(keys) + 1)))

        def search(key):       # DoubleArray::search, da.rs:133-153
            p = 1
            for c in key + b"\x00":
                q = int(da[p, 0]) + c
                if not (0 <= q < len(da)) or da[q, 1] != p:
                    return None
                p = q
            return -int(da[p, 0])

        for i, k in enumerate(keys):
            assert search(k) == i + 1
        assert search(b"zzz") is None and search(b"") is None


def test_builder_matches_oracle_on_ipadic(oracle_ipadic):
    """The product's DictionaryBuilder and the oracle's restatement of builder.rs are written
    independently (different EUC-JP handling, parsers and index construction): equal arrays."""
    from kanpyo_b200 import builder
    d = builder.ipadic()
    o = oracle_ipadic
    for k in ("da", "dup_ids", "dup_counts", "morphs", "conn", "char_category", "invoke_list", "group_list", "unk_cat",
              "unk_first_id", "unk_count", "unk_morphs"):
        assert np.array_equal(getattr(d, k), getattr(o, k)), k
    assert (d.conn_row, d.conn_col) == (o.conn_row, o.conn_col) == (1316, 1316)
    assert d.keywords == o.keywords and d.char_class == o.char_class
    assert len(d.keywords) == 392126 and len(set(d.keywords)) == 325871


def test_builder_small_directory(tmp_path, oracle_mod):
    """UTF-8 source tree with a quoted field, duplicates and a char.def override."""
    from kanpyo_b200 import builder
    (tmp_path / "a.csv").write_text('辞書,1,1,100,名詞\nテスト,0,0,50,"名,詞"\n辞書,1,1,90,動詞\n', encoding="utf-8")
    (tmp_path / "b.csv").write_text("あ,2,2,10,感動詞\n", encoding="utf-8")
    (tmp_path / "matrix.def").write_text("3 3\n0 0 1\n0 1 2\n1 0 3\n2 2 -7\n", encoding="utf-8")
    (tmp_path / "char.def").write_text("DEFAULT 0 1 0\nKANJI 0 0 2\nHIRAGANA 1 1 0 # c\n\n0x3041..0x309F HIRAGANA\n"
                                       "0x4E00..0x9FA5 KANJI\n0x3042 KANJI HIRAGANA\n", encoding="utf-8")
    (tmp_path / "unk.def").write_text("KANJI,1,1,500,名詞\nDEFAULT,0,0,700,記号\nKANJI,1,1,400,名詞\n", encoding="utf-8")
    d = builder.DictionaryBuilder(str(tmp_path), "utf8").build()
    assert d.keywords == sorted(k.encode() for k in ["辞書", "テスト", "辞書", "あ"])
    kws = [k.decode() for k in d.keywords]
    assert d.morphs[kws.index("テスト")].tolist() == [0, 0, 50]
    i = kws.index("辞書")
    assert d.morphs[i].tolist() == [1, 1, 90] and d.morphs[i + 1].tolist() == [1, 1, 100]   # cost orders duplicates
    assert d.dup_ids.tolist() == [i + 1] and d.dup_counts.tolist() == [1]
    assert d.features[1][d.features[0][kws.index("テスト")][0]] == "名,詞"
    conn = d.conn.reshape(3, 3)          # data[c*row + r]
    assert conn[0, 0] == 1 and conn[1, 0] == 2 and conn[0, 1] == 3 and conn[2, 2] == -7
    assert d.char_class == ["DEFAULT", "KANJI", "HIRAGANA"]
    assert d.char_category[0x3042] == 1 and d.char_category[0x3041] == 2 and d.char_category[0x4E00] == 1
    assert d.invoke_list.tolist() == [0, 0, 1] and d.group_list.tolist() == [1, 0, 1]
    # unk: sorted by class NAME -> DEFAULT(700) id1, KANJI(400) id2, KANJI(500) id3
    assert d.unk_cat.tolist() == [0, 1] and d.unk_first_id.tolist() == [1, 2] and d.unk_count.tolist() == [1, 2]
    assert d.unk_morphs.tolist() == [[0, 0, 700], [1, 1, 400], [1, 1, 500]]
    od = oracle_mod.dictbuild.from_dir(str(tmp_path), "utf8", oracle_mod.da_build)
    assert np.array_equal(d.da, od.da) and np.array_equal(d.char_category, od.char_category)


def test_builder_errors(tmp_path):
    from kanpyo_b200 import builder
    (tmp_path / "a.csv").write_text("あ,0,0,40000,x\n", encoding="utf-8")
    (tmp_path / "matrix.def").write_text("1 1\n0 0 0\n")
    (tmp_path / "char.def").write_text("DEFAULT 0 0 0\n")
    (tmp_path / "unk.def").write_text("DEFAULT,0,0,0,x\n")
    with pytest.raises(builder.BuilderError):      # builder.rs:59-61
        builder.DictionaryBuilder(str(tmp_path), "utf8").build()
    (tmp_path / "a.csv").write_text("あ,0,0,4,x\n", encoding="utf-8")
    (tmp_path / "char.def").write_text("DEFAULT 0 0 0\n0x00e0 DEFAULT\n")      # lower-case hex: no regex matches
    with pytest.raises(builder.BuilderError):
        builder.DictionaryBuilder(str(tmp_path), "utf8").build()
    (tmp_path / "char.def").write_text("DEFAULT 0 0 0\n")
    (tmp_path / "matrix.def").write_text("1 1\n0 0 99999\n")                  # matrix_def.rs:54
    with pytest.raises(builder.BuilderError):
        builder.DictionaryBuilder(str(tmp_path), "utf8").build()


def test_dict_pack_blob_is_validated(oracle_mod):
    from kanpyo_b200 import _lib
    d = to_product_dict(reference_fixture_dict(oracle_mod))
    blob = d.pack()
    assert blob[:8].tobytes() == b"PKB200D1" and blob.size % 256 == 0
    L = _lib.load()
    bad = to_product_dict(reference_fixture_dict(oracle_mod))
    bad.morphs = bad.morphs.copy()
    bad.morphs[0, 0] = 7                   # left_id outside the 3x3 matrix: the reference would panic
    a, _keep = bad._arrays()
    size = C.c_uint64()
    assert L.kp_dict_pack(C.byref(a), None, 0, C.byref(size)) == _lib.KP_ERR_DICT
    assert b"" != L.kp_last_error()


def test_first_character_table_in_blob(oracle_mod):
    """kp_dict_pack appends a first-character table {state, base[state]} per code point < 0x10000: it must equal a
    byte-by-byte walk of the double array (da.rs:155-182) from the root, with -1 where a transition fails."""
    import struct
    from helpers import to_product_dict
    od = oracle_mod.load_ipadic()
    blob = np.asarray(to_product_dict(od).pack()).tobytes()
    hdr = struct.unpack_from("<QIIQ6Q7Q4Q", blob, 0)
    da_len, off_da, off_first = hdr[4], hdr[10], hdr[19]
    da = np.frombuffer(blob, dtype=np.int32, count=da_len * 2, offset=off_da).reshape(-1, 2)
    assert np.array_equal(da, np.asarray(od.da, dtype=np.int32).reshape(-1, 2))
    first = np.frombuffer(blob, dtype=np.int32, count=65536 * 2, offset=off_first).reshape(-1, 2)

    def walk(bs):
        prev, base, q = 1, int(da[1, 0]), 0
        for c in bs:
            q = base + c
            if q < 0 or q >= da_len or da[q, 1] != prev:
                return (-1, 0)
            prev, base = q, int(da[q, 0])
        return (q, base)

    assert (first[0xD800:0xE000, 0] == -2).all()            # surrogates never occur in valid UTF-8
    rng = np.random.default_rng(7)
    cps = set(rng.integers(0, 0x10000, 3000).tolist()) | set(range(0x3040, 0x3100)) | {0, 0x7F, 0x80, 0x7FF, 0x800, 0xFFFF}
    for cp in sorted(cps):
        if 0xD800 <= cp < 0xE000:
            continue
        f = (int(first[cp, 0]), int(first[cp, 1]))
        assert f[0] != -2, "IPADIC keys are whole UTF-8 strings: no entry needs the byte-wise path"
        assert f == walk(chr(cp).encode("utf-8")), hex(cp)


def test_blob_flags_keys_ending_inside_a_character(oracle_mod):
    """The counting walk skips terminator probes in the middle of a character when every key of the dictionary
    ends on a character boundary (true for anything the reference can build: its keywords are Strings).
    kp_dict_pack checks the arrays it is given and records the answer in the blob header."""
    import struct
    from kanpyo_b200 import builder

    def flag(d):
        return struct.unpack_from("<QIIQ6Q7Q4Q", np.asarray(d.pack()).tobytes(), 0)[20]

    assert flag(to_product_dict(oracle_mod.load_ipadic())) == 0
    d = to_product_dict(reference_fixture_dict(oracle_mod))
    assert flag(d) == 0
    d.da = builder.da_build(["テスト".encode(), "形".encode()[:2], "辞書".encode()], [1, 2, 3])   # a key cut inside 形
    assert flag(d) == 1
    d.da = builder.da_build([b"a", "あa".encode(), "\U0001f600".encode()], [1, 2, 3])
    assert flag(d) == 0
    # a word's length travels in 16 bits on the device: longer keys are refused, not truncated
    from kanpyo_b200 import _lib
    d.da = builder.da_build([b"a" * 70000, "テスト".encode(), "辞書".encode()], [1, 2, 3])
    with pytest.raises(_lib.KanpyoB200Error):
        d.pack()


def test_sass_invariants_of_the_hot_kernels():
    """What the profiles rest on, checked on the built library without a GPU: the sweep's pair loop is one DPX
    add-min per pair (VIADDMNMX), and neither it nor the counting walk spills registers to local memory."""
    import shutil
    from kanpyo_b200 import build as kbuild
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", kbuild.build()], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for ln in sass.splitlines():
        if "Function :" in ln:
            cur = ln.split("Function :")[1].strip()
            funcs[cur] = []
        elif cur:
            funcs[cur].append(ln)
    vit = [k for k in funcs if "kp_viterbi" in k]
    assert len(vit) >= 3, vit                       # 8, 16 and 32 lanes per sentence (and 4 for sweeps)
    for k in vit:
        body = "\n".join(funcs[k])
        assert "VIADDMNMX" in body, k
        if "ILi8E" in k:
            assert "STL" not in body and "LDL" not in body, "%s spills" % k
    cnt = [k for k in funcs if "kp_lattice_count" in k]
    assert cnt and all("STL" not in "\n".join(funcs[k]) for k in cnt)


def test_shard_by_bytes_balances_and_covers():
    from kanpyo_b200.corpus import shard_by_bytes
    rng = np.random.default_rng(3)
    lens = rng.integers(0, 500, 1000)
    lens[10] = 40000
    off = np.zeros(1001, np.uint64)
    off[1:] = np.cumsum(lens)
    for n in (1, 2, 3, 8):
        sh = shard_by_bytes(off, n)
        assert sh[0][0] == 0 and sh[-1][1] == 1000
        assert all(sh[i][1] == sh[i + 1][0] for i in range(n - 1))
        sizes = [int(off[b] - off[a]) for a, b in sh]
        assert max(sizes) - min(sizes) <= 40000 + 500
    assert shard_by_bytes(np.zeros(1, np.uint64), 4) == [(0, 0)] * 4


_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from kanpyo_b200 import sharded, corpus
from oracle import oracle
from helpers import oracle_to_token8, reference_fixture_dict, to_product_dict

dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
od = reference_fixture_dict(oracle)
# 1. dictionary blob broadcast: only rank 0 packs
blob = to_product_dict(od).pack() if rank == 0 else None
t = sharded.broadcast_dict_blob(blob, 0, "cpu")
expect = to_product_dict(od).pack()
assert np.array_equal(t.numpy(), expect), "blob differs after broadcast"
# 2. shard, tokenize locally (CPU oracle stands in for the device pass), gather on rank 0
sents = ["テスト", "", "辞書テスト形態素", "あいうえお", "漢字", "テストテストテスト" * 5, "x", "形態素"] * 3
blobs = [s.encode() for s in sents]
off = np.zeros(len(blobs) + 1, np.uint64); off[1:] = np.cumsum([len(b) for b in blobs])
text = np.frombuffer(b"".join(blobs), np.uint8)
otk = oracle.OracleTokenizer(od)
s0, s1 = corpus.shard_by_bytes(off, world)[rank]
l_off, l_tok, l_cost, _ = otk.tokenize_batch(text, off[s0:s1 + 1])
rec = oracle_to_token8(l_tok)
ranges = corpus.shard_by_bytes(off, world)
tg = sharded.TokenGather("cpu", max(b - a for a, b in ranges), 64, 0)
def u8(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).copy())
for _ in range(2):      # the buffers are reused across calls
    g = tg.gather(u8(l_off.astype(np.uint32)), u8(rec), u8(l_cost))
if rank == 0:
    res = sharded.to_batch_result(*g, off)
    f_off, f_tok, f_cost, _ = otk.tokenize_batch(text, off)
    assert np.array_equal(res.tok_off, f_off), (res.tok_off, f_off)
    assert np.array_equal(res.eos_cost, f_cost)
    assert np.array_equal(res.tokens["id"].astype(np.int64), f_tok[:, 0])
    assert np.array_equal(res.tokens["cls"].astype(np.int64), f_tok[:, 1])
    assert np.array_equal(res.tokens["position"].astype(np.int64), f_tok[:, 2])
    assert np.array_equal(res.tokens["start"].astype(np.int64), f_tok[:, 3])
    assert np.array_equal(res.tokens["start"].astype(np.int64) + res.tokens["char_len"], f_tok[:, 4])
else:
    assert g is None
# 3. an empty shard still takes part; a shard beyond the capacity raises instead of truncating
e = tg.gather(u8(np.zeros(1, np.uint32)), torch.zeros(0, dtype=torch.uint8), torch.zeros(0, dtype=torch.uint8))
if rank == 0:
    assert e[0].tolist() == [0] and e[1].numel() == 0 and e[2].numel() == 0
try:
    tg.gather(u8(np.zeros(2, np.uint32)), torch.zeros(8 * 65, dtype=torch.uint8), torch.zeros(4, dtype=torch.uint8))
    raise SystemExit("capacity overflow not detected")
except ValueError:
    pass
dist.barrier()
dist.destroy_process_group()
print("worker", rank, "ok")
"""


def test_sharded_gather_gloo_world2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   GLOO_SOCKET_IFNAME="lo")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o)
        assert "worker %d ok" % r in o


def test_product_feature_tables_match_goldens():
    """print_tokens' feature join (src/bin/kanpyo.rs:174-197) from the PRODUCT builder's tables, checked
    on the ids of the committed golden tokens (no compute involved)."""
    import json
    from types import SimpleNamespace
    from kanpyo_b200 import builder
    d = builder.ipadic()
    with open(os.path.join(ROOT, "tests", "golden", "ipadic_sentences.json"), encoding="utf-8") as f:
        golden = json.load(f)
    n = 0
    for sent in golden["sentences"]:
        for tid, cls, _pos, _start, _end, _surface, feats in sent["tokens"]:
            assert ",".join(d.token_features(SimpleNamespace(id=tid, cls=cls))) == feats
            n += 1
    assert n > 100


def test_dict_container_round_trip_and_known_bytes(tmp_path, oracle_mod):
    """The reference's `.dict` zip container (dict.rs:51-116): hand-derived byte vectors for the fixed
    little-endian sections and the bincode-standard sections, then full round trips."""
    import io
    import struct
    import zipfile
    from kanpyo_b200 import builder, dictfile
    # bincode config::standard(): varint lengths, u8 / bool one byte each
    assert dictfile.encode_chardef(["A"], [0, 1], [1], [0]) == bytes([1, 1, 0x41, 2, 0, 1, 1, 1, 1, 0])
    big = dictfile.encode_chardef([], np.zeros(65536, np.uint8), [], [])
    assert big[:6] == bytes([0, 252]) + struct.pack("<I", 65536) and len(big) == 6 + 65536 + 2
    assert dictfile.encode_feature_table(([[1, 300]], ["", "x"])) == bytes([1, 2, 1, 251, 0x2C, 0x01, 2, 0, 1, 0x78])
    d = to_product_dict(reference_fixture_dict(oracle_mod))
    d.features = ([[1, 2], [1, 2], [1, 3]], ["", "名詞", "一般", "固有"])
    d.unk_features = ([[1], [1]], ["", "未知"])
    buf = io.BytesIO()
    dictfile.save_dict(d, buf)
    with zipfile.ZipFile(io.BytesIO(buf.getvalue())) as z:
        assert sorted(z.namelist()) == sorted(dictfile.MEMBERS)
        m = z.read("morph.dict")                                         # morph.rs:61-71
        assert m == struct.pack("<q", 3) + struct.pack("<9h", 0, 0, 1000, 1, 1, 1200, 2, 2, 1100)
        c = z.read("connection.dict")                                    # connection.rs:43-50
        assert c == struct.pack("<QQ", 3, 3) + struct.pack("<9h", 0, 100, 200, 100, 0, 100, 200, 100, 0)
        ix = z.read("index.dict")                                        # da.rs:236-245, index.rs:74-83
        n = struct.unpack_from("<Q", ix)[0]
        assert n == len(d.da) and ix[8 + 8 * n:] == struct.pack("<Q", 0)
        u = z.read("unk.dict")                                           # unk_dict.rs:62-72
        assert u[:8 + 2 * 17] == struct.pack("<Q", 2) + struct.pack("<BqQ", 1, 1, 1) + struct.pack("<BqQ", 2, 2, 1)
    e = dictfile.load_dict(buf.getvalue())
    for k in Dict_FIELDS:
        assert np.array_equal(getattr(e, k), getattr(d, k)), k
    assert (e.conn_row, e.conn_col, e.char_class) == (3, 3, d.char_class)
    assert e.features == d.features and e.unk_features == d.unk_features
    # IPADIC: file -> Dict -> identical packed device blob
    full = builder.ipadic()
    path = tmp_path / "ipa.dict"
    dictfile.save_dict(full, str(path))
    back = dictfile.load_dict(str(path))
    assert np.array_equal(back.pack(), full.pack())
    assert back.features == full.features and back.unk_features == full.unk_features
    with pytest.raises(dictfile.DictFormatError):
        dictfile.load_dict(b"not a zip")


Dict_FIELDS = ("da", "dup_ids", "dup_counts", "morphs", "conn", "char_category", "invoke_list", "group_list", "unk_cat",
               "unk_first_id", "unk_count", "unk_morphs")


def test_expand_tokens8_rebuilds_positions(oracle_mod, oracle_tok, vocab):
    """kp_token8 -> kp_token on the host (C helper and its numpy mirror): positions / starts are rebuilt
    backwards from each sentence's EOS record, also for truncated paths (first token not at byte 0) and
    empty paths (no token at all)."""
    from kanpyo_b200 import _lib, corpus
    from kanpyo_b200.tokenizer import TOKEN_DTYPE, expand_tokens8
    from helpers import oracle_to_token8, pack
    cases = []
    text, off = corpus.synth_corpus(vocab, 300, "cfg3")
    cases.append((oracle_tok, text, off))
    fx = oracle_mod.OracleTokenizer(reference_fixture_dict(oracle_mod))
    text, off = pack(["テスト", "", "xテスト", "テストx", "xx", "テxスト", "あいうえお", "x"])
    cases.append((fx, text, off))
    L = _lib.load()
    for tk, text, off in cases:
        o_off, o_tok, _, _ = tk.tokenize_batch(text, off)
        t8 = oracle_to_token8(o_tok)
        full = expand_tokens8(o_off, t8, off)
        assert np.array_equal(full["id"].astype(np.int64), o_tok[:, 0])
        assert np.array_equal(full["cls"].astype(np.int64), o_tok[:, 1])
        assert np.array_equal(full["position"].astype(np.int64), o_tok[:, 2])
        assert np.array_equal(full["start"].astype(np.int64), o_tok[:, 3])
        assert np.array_equal(full["start"].astype(np.int64) + full["char_len"], o_tok[:, 4])
        off32 = np.ascontiguousarray(o_off, np.uint32)
        off64 = np.ascontiguousarray(off, np.uint64)
        r = _lib.Result8(n_sent=len(off) - 1, n_tokens=len(t8), tok_off=off32.ctypes.data, tokens=t8.ctypes.data,
                         eos_cost=None)
        out = np.zeros(len(t8), TOKEN_DTYPE)
        _lib.check(L.kp_expand_tokens8(C.byref(r), off64.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        assert np.array_equal(out, full)
    assert any((np.diff(o_off) == 0).any() for _ in [0])      # the fixture batch holds empty paths


def test_expand_tokens8_large_result_uses_threads_and_agrees():
    """Above 2^18 tokens kp_expand_tokens8 splits the sentences over host threads: same records as the numpy
    mirror on a synthetic result (ragged sentences, empty paths in between), and a sentence whose last record
    is not EOS is still reported by number."""
    from kanpyo_b200 import _lib
    from kanpyo_b200.tokenizer import TOKEN8_DTYPE, TOKEN_DTYPE, expand_tokens8
    rng = np.random.default_rng(7)
    n_sent = 40_000
    counts = rng.integers(0, 40, n_sent)                       # tokens before EOS; 0 with no EOS = empty path
    empty = rng.random(n_sent) < 0.02
    per = np.where(empty, 0, counts + 1)
    tok_off = np.zeros(n_sent + 1, np.uint32)
    tok_off[1:] = np.cumsum(per)
    nt = int(tok_off[-1])
    assert nt > (1 << 19)
    t8 = np.zeros(nt, TOKEN8_DTYPE)
    t8["id_cls"] = rng.integers(1, 1 << 20, nt).astype(np.uint32) | (np.uint32(1) << 30)
    t8["byte_len"] = rng.integers(1, 13, nt)
    t8["char_len"] = rng.integers(1, 5, nt)
    last = tok_off[1:][per > 0] - 1
    sent_of = np.repeat(np.arange(n_sent), per)
    body = np.ones(nt, bool); body[last] = False
    nbytes = np.bincount(sent_of[body], t8["byte_len"][body], n_sent).astype(np.uint64)
    nchars = np.bincount(sent_of[body], t8["char_len"][body], n_sent).astype(np.uint32)
    t8["id_cls"][last] = 0                                     # EOS: class Dummy, n_chars in the two length fields
    t8["byte_len"][last] = nchars[per > 0] & 0xFFFF
    t8["char_len"][last] = nchars[per > 0] >> 16
    off = np.zeros(n_sent + 1, np.uint64)
    off[1:] = np.cumsum(nbytes + rng.integers(0, 3, n_sent).astype(np.uint64))   # some paths start after byte 0
    full = expand_tokens8(tok_off, t8, off)
    L = _lib.load()
    r = _lib.Result8(n_sent=n_sent, n_tokens=nt, tok_off=tok_off.ctypes.data, tokens=t8.ctypes.data, eos_cost=None)
    out = np.zeros(nt, TOKEN_DTYPE)
    _lib.check(L.kp_expand_tokens8(C.byref(r), off.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
    assert np.array_equal(out, full)
    victim = int(np.flatnonzero(per > 0)[-3])
    t8["id_cls"][tok_off[victim + 1] - 1] = np.uint32(1) << 30
    assert L.kp_expand_tokens8(C.byref(r), off.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == -1
    assert ("sentence %d" % victim) in L.kp_last_error().decode()


def test_blob_is_refused_unless_it_is_what_pack_wrote(oracle_mod):
    """kp_dict_create_from_blob validates before it touches a device: a flipped payload byte,
    a truncated blob, a section pointing outside the blob or a stale layout version all give KP_ERR_BLOB."""
    from kanpyo_b200 import _lib
    L = _lib.load()
    blob = to_product_dict(reference_fixture_dict(oracle_mod)).pack()

    def status(b):
        h = C.c_void_p()
        buf = np.ascontiguousarray(b)
        return L.kp_dict_create_from_blob(buf.ctypes.data_as(C.c_void_p), buf.size, 0, C.byref(h))

    good = status(blob)
    assert good in (0, -2), good                  # OK on a GPU box; "no device" here -- but never KP_ERR_BLOB
    bad = blob.copy(); bad[len(bad) // 2] ^= 1
    assert status(bad) == -7
    assert status(blob[:len(blob) - 256]) == -7
    hdr = blob.copy().view(np.uint64)
    hdr[9] = len(blob) + 256                      # off_da beyond the blob
    assert status(hdr.view(np.uint8)) == -7
    ver = blob.copy(); ver[8] = 1                 # layout version 1 (round 1) is not accepted
    assert status(ver) == -7


def test_builder_csv_is_as_strict_as_the_reference():
    """parse_csv (kanpyo-dict/src/builder/record.rs:21-42) reads with the `csv` crate and `str::parse`: quoted fields
    may hold commas and line breaks, every record has the first record's field count, ids parse as usize and the cost
    as i64 (a leading '+' is legal; '-' on an id, blanks, '_' are not).  What the reference refuses is refused here."""
    from kanpyo_b200.builder import BuilderError, _read_records
    rows = _read_records('東京,1,2,300,名詞,"a,b"\n"京\n都",+3,4,-5,名詞,"say ""hi"""\n\n', "t.csv")
    assert rows == [("東京".encode(), 1, 2, 300, ("名詞".encode(), b"a,b")),
                    ("京\n都".encode(), 3, 4, -5, ("名詞".encode(), b'say "hi"'))]
    for bad in ["a,1,2,3,x\nb,1,2,3\n",          # differing field counts
                "a,-1,2,3\n", "a,1,-2,3\n",      # negative ids
                "a,1_000,2,3\n", "a, 1,2,3\n", "a,1,2,3.0\n", "a,1,2,\n", "a,0x10,2,3\n",
                "a,1,2\n",                        # fewer than four fields
                "a,1,2,99999999999999999999\n"]:  # cost beyond i64
        with pytest.raises(BuilderError):
            _read_records(bad, "t.csv")


def test_rust_shim_declarations_match_the_header():
    """integration/rust/ cannot be compiled here (no Rust toolchain), so at least its extern "C" block is held
    against include/kanpyo_b200.h: every function it declares exists there with the same number of parameters,
    and the record structs it mirrors have the header's fields in the header's order."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "kanpyo_b200.h"), encoding="utf-8").read()
    rs = open(os.path.join(root, "integration", "rust", "src", "ffi.rs"), encoding="utf-8").read()
    hdr_nc = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)

    def n_params(arglist):
        arglist = arglist.strip()
        return 0 if arglist in ("", "void") else arglist.count(",") + 1

    c_protos = {m.group(1): n_params(m.group(2)) for m in re.finditer(r"\b(kp_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", hdr_nc)}
    rs_nc = re.sub(r"//[^\n]*", " ", rs)
    rs_protos = {m.group(1): n_params(m.group(2).rstrip().rstrip(","))
                 for m in re.finditer(r"pub fn (kp_[a-z0-9_]+)\s*\(([^()]*)\)", rs_nc)}
    assert len(rs_protos) >= 20
    for name, n in rs_protos.items():
        assert name in c_protos, "%s is declared in ffi.rs but not in the header" % name
        assert c_protos[name] == n, "%s: %d parameters in ffi.rs, %d in the header" % (name, n, c_protos[name])

    def c_fields(struct):
        body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s\s*;" % (struct, struct), hdr_nc, flags=re.S).group(1)
        return [re.search(r"(\w+)\s*(\[[^\]]*\])?\s*$", d.strip()).group(1) for d in body.split(";") if d.strip()]

    def rs_fields(struct):
        body = re.search(r"pub struct %s\s*\{(.*?)\}" % struct, rs_nc, flags=re.S).group(1)
        return re.findall(r"pub (\w+)\s*:", body)

    for struct in ("kp_token", "kp_token8", "kp_result", "kp_result8"):
        assert rs_fields(struct) == c_fields(struct), struct
