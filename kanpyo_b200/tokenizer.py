"""Tokenizer / Token / TokenClass — host-side mirror of src/tokenizer.rs:7-45 and src/token.rs:4-55,
calling the CUDA path through the C ABI (include/kanpyo_b200.h)."""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass

import numpy as np

from . import _lib
from .dict import Dict

TOKEN_DTYPE = np.dtype([("id", "<i4"), ("position", "<u4"), ("start", "<u4"), ("char_len", "<u2"), ("cls", "u1"),
                        ("reserved", "u1")])
LATTICE_NODE_DTYPE = np.dtype([("id", "<i4"), ("cls", "u1"), ("r0", "u1", (3,)), ("byte_pos", "<u4"),
                               ("char_pos", "<u4"), ("end_char", "<u4"), ("left_id", "<i2"), ("right_id", "<i2"),
                               ("cost", "<i2"), ("r1", "<i2"), ("dp", "<i4"), ("pre", "<i4")])
TOKEN8_DTYPE = np.dtype([("id_cls", "<u4"), ("byte_len", "<u2"), ("char_len", "<u2")])     # kp_token8
assert TOKEN_DTYPE.itemsize == 16 and LATTICE_NODE_DTYPE.itemsize == 36 and TOKEN8_DTYPE.itemsize == 8


def expand_tokens8(tok_off: np.ndarray, tokens8: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    """kp_token8 -> kp_token records on the host (numpy mirror of kp_expand_tokens8): every sentence's
    tokens are walked backwards from its EOS record, which carries the sentence's char count; position
    and start are the running differences (consecutive path nodes are adjacent in the input)."""
    nt = len(tokens8)
    out = np.zeros(nt, TOKEN_DTYPE)
    if nt == 0:
        return out
    cls = (tokens8["id_cls"] >> 30).astype(np.uint8)
    out["id"] = (tokens8["id_cls"] & 0x3FFFFFFF).astype(np.int32)
    out["cls"] = cls
    tok_off = np.asarray(tok_off, np.int64)
    counts = np.diff(tok_off)
    sent_of = np.repeat(np.arange(len(counts)), counts)
    is_eos = cls == 0
    last = tok_off[1:][counts > 0] - 1
    if not is_eos[last].all():
        raise ValueError("a sentence's last token is not EOS")
    n_chars = tokens8["byte_len"].astype(np.int64) | (tokens8["char_len"].astype(np.int64) << 16)
    bl = np.where(is_eos, 0, tokens8["byte_len"].astype(np.int64))
    cl = np.where(is_eos, 0, tokens8["char_len"].astype(np.int64))
    # suffix sums inside each sentence: S[k] = sum_{j >= k, same sentence} len_j
    def suffix(v):
        c = np.cumsum(v)
        total_to_end = c[tok_off[1:][sent_of] - 1]          # cumsum at the sentence's last token
        return total_to_end - c + v
    sent_bytes = np.diff(np.asarray(offsets, np.uint64)).astype(np.int64)
    eos_start = np.zeros(len(counts), np.int64)
    eos_start[counts > 0] = n_chars[last]
    out["position"] = (sent_bytes[sent_of] - suffix(bl)).astype(np.uint32)
    out["start"] = (eos_start[sent_of] - suffix(cl)).astype(np.uint32)
    out["char_len"] = np.where(is_eos, 3, tokens8["char_len"]).astype(np.uint16)
    return out


class TokenClass(enum.IntEnum):   # src/token.rs:4-8
    Dummy = 0
    Known = 1
    Unknown = 2


@dataclass(frozen=True)
class Token:                      # src/token.rs:11-18
    id: int
    cls: TokenClass
    position: int                 # byte position
    start: int                    # char position
    end: int                      # char position
    surface: str

    def length(self) -> int:      # src/token.rs:39-41
        return self.end - self.start


@dataclass
class BatchResult:
    """Packed result of a batch call (copies; independent of the tokenizer's buffers)."""
    tok_off: np.ndarray           # uint64 [n_sent+1]
    tokens: np.ndarray            # TOKEN_DTYPE [n_tokens]
    eos_cost: np.ndarray          # int32 [n_sent]

    def sentence(self, s: int) -> np.ndarray:
        return self.tokens[int(self.tok_off[s]):int(self.tok_off[s + 1])]


class Tokenizer:
    """`Tokenizer::new(dict)` (src/tokenizer.rs:12-14) on CUDA device `device`."""

    def __init__(self, dict: Dict, device: int = 0):
        self.dict = dict
        self.device = device
        self._L = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._L.kp_tokenizer_create(dict.device_handle(device), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.kp_tokenizer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Tokenizer::tokenize (src/tokenizer.rs:16-45) ---------------------------------------------
    def tokenize(self, input: str) -> list:
        b = input.encode("utf-8")
        res = self.tokenize_batch_bytes(b, np.array([0, len(b)], np.uint64))
        return self._materialize(b, 0, res.sentence(0))

    def tokenize_with_cost(self, input: str):
        b = input.encode("utf-8")
        res = self.tokenize_batch_bytes(b, np.array([0, len(b)], np.uint64))
        return self._materialize(b, 0, res.sentence(0)), int(res.eos_cost[0])

    def tokenize_batch(self, inputs) -> list:
        blobs = [s.encode("utf-8") for s in inputs]
        off = np.zeros(len(blobs) + 1, np.uint64)
        if blobs:
            off[1:] = np.cumsum([len(x) for x in blobs], dtype=np.uint64)
        text = b"".join(blobs)
        res = self.tokenize_batch_bytes(text, off)
        return [self._materialize(text, int(off[i]), res.sentence(i)) for i in range(len(blobs))]

    @staticmethod
    def _materialize(text: bytes, base: int, toks: np.ndarray) -> list:
        out = []
        n = len(toks)
        for k in range(n):
            t = toks[k]
            cls = TokenClass(int(t["cls"]))
            pos = int(t["position"])
            if cls == TokenClass.Dummy:
                surface = "EOS"                                   # src/tokenizer.rs:28-29
            else:
                nxt = int(toks[k + 1]["position"])                # path nodes are adjacent; EOS closes the path
                surface = text[base + pos:base + nxt].decode("utf-8")
            start = int(t["start"])
            out.append(Token(int(t["id"]), cls, pos, start, start + int(t["char_len"]), surface))
        return out

    def format_tokens(self, tokens) -> str:
        """The `kanpyo tokenize` output, `print_tokens` of src/bin/kanpyo.rs:174-197: one line per token,
        `surface<TAB>feature,feature,...`."""
        return "".join("%s\t%s\n" % (t.surface, ",".join(self.dict.token_features(t))) for t in tokens)

    # ---- packed batch entry points -----------------------------------------------------------------
    def tokenize_batch_bytes(self, text, offsets: np.ndarray) -> BatchResult:
        """Host text (bytes / uint8 array) + uint64 offsets [n+1] -> BatchResult (kp_tokenize_batch)."""
        buf = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else np.ascontiguousarray(text, np.uint8)
        off = np.ascontiguousarray(offsets, np.uint64)
        n = len(off) - 1
        r = _lib.Result()
        _lib.check(self._L.kp_tokenize_batch(self._h, buf.ctypes.data_as(C.c_void_p) if buf.size else None,
                                             off.ctypes.data_as(C.c_void_p), n, C.byref(r)))
        return self._copy_host_result(r)

    def tokenize_batch8_bytes(self, text, offsets: np.ndarray):
        """kp_tokenize_batch8: compact records -> (tok_off uint32 [n+1], tokens TOKEN8_DTYPE, eos_cost int32) copies."""
        buf = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else np.ascontiguousarray(text, np.uint8)
        off = np.ascontiguousarray(offsets, np.uint64)
        r = _lib.Result8()
        _lib.check(self._L.kp_tokenize_batch8(self._h, buf.ctypes.data_as(C.c_void_p) if buf.size else None,
                                              off.ctypes.data_as(C.c_void_p), len(off) - 1, C.byref(r)))
        return copy_result8(r)

    def tokenize_batch8_ptr(self, text_ptr: int, offsets_ptr: int, n_sent: int) -> "_lib.Result8":
        r = _lib.Result8()
        _lib.check(self._L.kp_tokenize_batch8(self._h, text_ptr, offsets_ptr, n_sent, C.byref(r)))
        return r

    def set_path(self, path: str):
        """'auto' (fused kernel where a sentence fits, pipeline otherwise) | 'pipeline' | 'fused'."""
        _lib.check(self._L.kp_tokenizer_set_path(self._h, {"auto": 0, "pipeline": 1, "fused": 2}[path]))

    def tokenize_batch_ptr(self, text_ptr: int, offsets_ptr: int, n_sent: int) -> "_lib.Result":
        """Raw host-pointer call (e.g. pinned torch tensors); the returned views alias tokenizer memory."""
        r = _lib.Result()
        _lib.check(self._L.kp_tokenize_batch(self._h, text_ptr, offsets_ptr, n_sent, C.byref(r)))
        return r

    def tokenize_batch_device(self, d_text_ptr: int, d_offsets_ptr: int, n_sent: int, first_offset: int,
                              n_bytes: int) -> "_lib.Result":
        """Text and offsets already in HBM; the result stays in HBM (kp_tokenize_batch_device)."""
        r = _lib.Result()
        _lib.check(self._L.kp_tokenize_batch_device(self._h, d_text_ptr, d_offsets_ptr, n_sent, first_offset, n_bytes,
                                                    C.byref(r)))
        return r

    def tokenize_batch_device8(self, d_text_ptr: int, d_offsets_ptr: int, n_sent: int, first_offset: int,
                               n_bytes: int) -> "_lib.Result8":
        """kp_tokenize_batch_device8: as above with the compact result (kp_token8, u32 offsets) in HBM."""
        r = _lib.Result8()
        _lib.check(self._L.kp_tokenize_batch_device8(self._h, d_text_ptr, d_offsets_ptr, n_sent, first_offset, n_bytes,
                                                     C.byref(r)))
        return r

    def copy_device_result8(self, r):
        """Device-resident compact result -> (tok_off u32, tokens TOKEN8_DTYPE, eos i32) host copies."""
        n, nt = int(r.n_sent), int(r.n_tokens)
        tok_off = np.empty(n + 1, np.uint32)
        tokens = np.empty(nt, TOKEN8_DTYPE)
        eos = np.empty(n, np.int32)
        for dst, src in ((tok_off, r.tok_off), (tokens, r.tokens), (eos, r.eos_cost)):
            if dst.nbytes:
                _lib.check(self._L.kp_copy_to_host(self._h, dst.ctypes.data_as(C.c_void_p), src, dst.nbytes))
        return tok_off, tokens, eos

    def copy_device_result(self, r) -> BatchResult:
        """Bring a kp_tokenize_batch_device result back to host memory (kp_copy_to_host)."""
        n, nt = int(r.n_sent), int(r.n_tokens)
        tok_off = np.empty(n + 1, np.uint64)
        tokens = np.empty(nt, TOKEN_DTYPE)
        eos = np.empty(n, np.int32)
        for dst, src in ((tok_off, r.tok_off), (tokens, r.tokens), (eos, r.eos_cost)):
            _lib.check(self._L.kp_copy_to_host(self._h, dst.ctypes.data_as(C.c_void_p), src, dst.nbytes))
        return BatchResult(tok_off, tokens, eos)

    @staticmethod
    def _copy_host_result(r) -> BatchResult:
        n, nt = int(r.n_sent), int(r.n_tokens)
        tok_off = np.ctypeslib.as_array(C.cast(r.tok_off, C.POINTER(C.c_uint64)), shape=(n + 1,)).copy()
        if nt:
            raw = np.ctypeslib.as_array(C.cast(r.tokens, C.POINTER(C.c_uint8)), shape=(nt * 16,))
            tokens = raw.view(TOKEN_DTYPE).copy()
        else:
            tokens = np.zeros(0, TOKEN_DTYPE)
        eos = (np.ctypeslib.as_array(C.cast(r.eos_cost, C.POINTER(C.c_int32)), shape=(n,)).copy() if n
               else np.zeros(0, np.int32))
        return BatchResult(tok_off, tokens, eos)

    # ---- introspection ------------------------------------------------------------------------------
    def set_chunk_bytes(self, n: int):
        _lib.check(self._L.kp_tokenizer_set_chunk_bytes(self._h, n))

    def set_count_work(self, on: bool):
        _lib.check(self._L.kp_tokenizer_set_count_work(self._h, 1 if on else 0))

    def counters(self) -> dict:
        c = _lib.Counters()
        _lib.check(self._L.kp_last_counters(self._h, C.byref(c)))
        return {k: int(getattr(c, k)) for k, _ in _lib.Counters._fields_}

    def profile(self) -> dict:
        p = _lib.Profile()
        _lib.check(self._L.kp_last_profile(self._h, C.byref(p)))
        return {k: (float(getattr(p, k)) if t is C.c_float else int(getattr(p, k))) for k, t in _lib.Profile._fields_}

    def lattice(self, input: str) -> np.ndarray:
        """Lattice::build + viterbi internals (src/lattice.rs:101-154): LATTICE_NODE_DTYPE array in the
        reference's node order (BOS first, EOS last)."""
        b = input.encode("utf-8")
        la = _lib.Lattice()
        buf = np.frombuffer(b, np.uint8)
        _lib.check(self._L.kp_lattice_dump(self._h, buf.ctypes.data_as(C.c_void_p) if buf.size else None, len(b),
                                           C.byref(la)))
        raw = np.ctypeslib.as_array(C.cast(la.nodes, C.POINTER(C.c_uint8)), shape=(int(la.n_nodes) * 36,))
        return raw.view(LATTICE_NODE_DTYPE).copy()

    def common_prefix(self, input: str, expand_dup: bool = True):
        """IndexTable::search_common_prefix_of (index.rs:40-53) on the device trie -> [(id, byte_len)] or None."""
        b = input.encode("utf-8")
        cap = 4096
        ids = np.zeros(cap, np.int64)
        lens = np.zeros(cap, np.uint64)
        n = C.c_uint64()
        buf = np.frombuffer(b, np.uint8)
        _lib.check(self._L.kp_da_common_prefix(self._h, buf.ctypes.data_as(C.c_void_p) if buf.size else None, len(b),
                                               1 if expand_dup else 0, ids.ctypes.data_as(C.c_void_p),
                                               lens.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
        if n.value == 0:
            return None
        return [(int(ids[i]), int(lens[i])) for i in range(min(n.value, cap))]


def copy_result8(r):
    """kp_result8 with HOST pointers -> (tok_off uint32, tokens TOKEN8_DTYPE, eos_cost int32) numpy copies."""
    n, nt = int(r.n_sent), int(r.n_tokens)
    tok_off = np.ctypeslib.as_array(C.cast(r.tok_off, C.POINTER(C.c_uint32)), shape=(n + 1,)).copy()
    tokens = (np.ctypeslib.as_array(C.cast(r.tokens, C.POINTER(C.c_uint8)), shape=(nt * 8,)).view(TOKEN8_DTYPE).copy()
              if nt else np.zeros(0, TOKEN8_DTYPE))
    eos = (np.ctypeslib.as_array(C.cast(r.eos_cost, C.POINTER(C.c_int32)), shape=(n,)).copy() if n
           else np.zeros(0, np.int32))
    return tok_off, tokens, eos


def result8_to_batch(res8, offsets) -> BatchResult:
    tok_off, tokens8, eos = res8
    return BatchResult(tok_off.astype(np.uint64), expand_tokens8(tok_off, tokens8, offsets), eos)


class Queue:
    """kp_queue_*: `depth` tokenizer contexts on one device, so that the copies of one batch hide behind
    the kernels of its neighbours.  submit() returns a ticket at once; wait(ticket) blocks for its
    result (compact form).  The caller's buffers must stay alive until wait() returns."""

    def __init__(self, dict: Dict, device: int = 0, depth: int = 2):
        self.dict, self.device, self.depth = dict, device, depth
        self._L = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._L.kp_queue_create(dict.device_handle(device), depth, C.byref(self._h)))
        self._keep = {}

    def set_path(self, path: str):
        _lib.check(self._L.kp_queue_set_path(self._h, {"auto": 0, "pipeline": 1, "fused": 2}[path]))

    def set_blocking_sync(self, on: bool):
        """Worker threads sleep while the device works instead of spinning (many contexts per host core)."""
        _lib.check(self._L.kp_queue_set_blocking_sync(self._h, 1 if on else 0))

    def submit_ptr(self, text_ptr: int, offsets_ptr: int, n_sent: int) -> int:
        tk = C.c_uint64()
        _lib.check(self._L.kp_queue_submit(self._h, text_ptr, offsets_ptr, n_sent, C.byref(tk)))
        return int(tk.value)

    def submit(self, text, offsets) -> int:
        buf = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else np.ascontiguousarray(text, np.uint8)
        off = np.ascontiguousarray(offsets, np.uint64)
        tk = self.submit_ptr(buf.ctypes.data if buf.size else 0, off.ctypes.data, len(off) - 1)
        self._keep[tk] = (buf, off)
        return tk

    def wait_raw(self, ticket: int) -> "_lib.Result8":
        r = _lib.Result8()
        _lib.check(self._L.kp_queue_wait(self._h, ticket, C.byref(r)))
        return r

    def wait(self, ticket: int) -> BatchResult:
        r = self.wait_raw(ticket)
        buf, off = self._keep.pop(ticket)
        return result8_to_batch(copy_result8(r), off)

    def close(self):
        if getattr(self, "_h", None):
            self._L.kp_queue_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
