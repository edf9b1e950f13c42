"""Multi-GPU sharding of the tokenizer path: one process per GPU, `torch.distributed` for the plumbing.

Sentences are independent and the dictionary is read-only, so the path has no exchange step inside
the algorithm.  Exactly two collectives exist, both outside the kernels:

  * `broadcast_dict_blob` — ONE broadcast of the packed dictionary blob (17.5 MB for IPADIC) from the
    rank that built it; every rank stages its handle from the blob in HBM
    (`kp_dict_create_from_device_blob`: checksum-validated, copied into memory the handle owns).
  * `NcclGather.gather`   — the same gather behind the C ABI (`kp_gather_*`): the library's own grouped
    ncclSend / ncclRecv + a compaction kernel on rank 0; what bench.py times.
  * `TokenGather.gather`  — tokens of all shards to one rank, in global sentence order: ONE `gather`
    collective over preallocated fixed-capacity buffers (counts ride in a 16-byte header; no count
    exchange, no allocation, one host sync on the receiving rank).

`shard_by_bytes` (corpus.py) gives the contiguous, byte-balanced sentence ranges.  Backend `nccl`
moves device tensors over NVLink; backend `gloo` (CPU tensors) runs the same host logic in tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .tokenizer import BatchResult, TOKEN8_DTYPE, expand_tokens8


class DeviceBytes:
    """Zero-copy torch view of `nbytes` of device memory owned by the C library."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


def device_view(ptr: int, nbytes: int, device: int) -> torch.Tensor:
    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device="cuda:%d" % device)
    return torch.as_tensor(DeviceBytes(ptr, nbytes), device="cuda:%d" % device)


def broadcast_dict_blob(blob, src: int = 0, device=None, group=None) -> torch.Tensor:
    """`blob`: uint8 numpy array on rank `src` (ignored elsewhere).  Returns the blob as a uint8 tensor
    on `device` on every rank."""
    device = torch.device("cpu") if device is None else torch.device(device)
    size = torch.zeros(1, dtype=torch.int64, device=device)
    if dist.get_rank(group) == src:
        size[0] = int(blob.size)
    dist.broadcast(size, src, group=group)
    if dist.get_rank(group) == src:
        t = torch.from_numpy(np.ascontiguousarray(blob, np.uint8)).to(device)
    else:
        t = torch.empty(int(size.item()), dtype=torch.uint8, device=device)
    dist.broadcast(t, src, group=group)
    return t


class TokenGather:
    """Gather of the shards' packed results (compact form: kp_token8 records, u32 offsets, i32 costs) on
    rank `dst`, as ONE collective per call over preallocated buffers.

    Every rank copies its result into a fixed-capacity send buffer
        [n_sent, n_tok : int64 x 2][tok_off u32 x (cap_sent+1)][eos i32 x cap_sent][tokens 8 B x cap_tok]
    and `dist.gather` (NCCL: grouped send/recv over NVLink; gloo on CPU) lands the buffers of all ranks
    in rank order on `dst`.  No count exchange precedes the transfer and nothing is allocated per call;
    `dst` then reads the 16-byte headers back (its only host sync) and compacts: offsets rebased by the
    tokens of the preceding shards, token records and costs concatenated in rank order (= global
    sentence order for contiguous shards).  A shard larger than the capacity raises on every rank at
    the next call (the header says so), never silently truncates."""

    HEADER = 16

    def __init__(self, device, cap_sent: int, cap_tok: int, dst: int = 0, group=None):
        self.group, self.dst = group, dst
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device)
        self.cap_sent, self.cap_tok = int(cap_sent), int(cap_tok)
        self.o_off = self.HEADER
        self.o_eos = self.o_off + 4 * (self.cap_sent + 1)
        self.o_tok = (self.o_eos + 4 * self.cap_sent + 7) // 8 * 8
        self.nbytes = self.o_tok + 8 * self.cap_tok
        self.send = torch.zeros(self.nbytes, dtype=torch.uint8, device=self.device)
        self.recv = (torch.zeros(self.world, self.nbytes, dtype=torch.uint8, device=self.device)
                     if self.rank == dst else None)

    def fits(self, n_sent: int, n_tok: int) -> bool:
        return n_sent <= self.cap_sent and n_tok <= self.cap_tok

    def gather(self, tok_off: torch.Tensor, tokens: torch.Tensor, eos_cost: torch.Tensor):
        """tok_off uint8-view of u32 [n_sent+1] (shard-relative), tokens uint8 [n_tok*8], eos_cost uint8-view of
        i32 [n_sent]; all on self.device.  -> (tok_off int64 [S+1], tokens uint8 [T*8], eos int32 [S]) on
        `dst`, None elsewhere."""
        n_sent, n_tok = eos_cost.numel() // 4, tokens.numel() // 8
        if not self.fits(n_sent, n_tok):
            raise ValueError("shard (%d sentences, %d tokens) exceeds the gather capacity (%d, %d)"
                             % (n_sent, n_tok, self.cap_sent, self.cap_tok))
        s = self.send
        s[:16].view(torch.int64).copy_(torch.tensor([n_sent, n_tok], dtype=torch.int64), non_blocking=True)
        s[self.o_off:self.o_off + 4 * (n_sent + 1)].copy_(tok_off)
        s[self.o_eos:self.o_eos + 4 * n_sent].copy_(eos_cost)
        s[self.o_tok:self.o_tok + 8 * n_tok].copy_(tokens)
        if self.world > 1:
            dist.gather(s, list(self.recv.unbind(0)) if self.rank == self.dst else None, dst=self.dst, group=self.group)
        elif self.rank == self.dst:
            self.recv[0].copy_(s)
        if self.rank != self.dst:
            return None
        r = self.recv
        counts = r[:, :16].contiguous().view(torch.int64).view(self.world, 2).cpu()      # the one host sync
        ns, nt = counts[:, 0].tolist(), counts[:, 1].tolist()
        tok_base = [0]
        for k in nt[:-1]:
            tok_base.append(tok_base[-1] + k)
        offs = r[:, self.o_off:self.o_off + 4 * (self.cap_sent + 1)].contiguous().view(torch.int32).view(self.world, -1)
        eoss = r[:, self.o_eos:self.o_eos + 4 * self.cap_sent].contiguous().view(torch.int32).view(self.world, -1)
        g_off = torch.cat([offs[k, :ns[k]].to(torch.int64) + tok_base[k] for k in range(self.world)]
                          + [torch.tensor([tok_base[-1] + nt[-1]], dtype=torch.int64, device=self.device)])
        g_eos = torch.cat([eoss[k, :ns[k]] for k in range(self.world)])
        g_tok = torch.cat([r[k, self.o_tok:self.o_tok + 8 * nt[k]] for k in range(self.world)])
        return g_off, g_tok, g_eos


class NcclGather:
    """kp_gather_*: the token gather behind the C ABI for one-process-per-GPU launchers.  The NCCL unique id
    travels from rank 0 through `torch.distributed` (any channel would do); from then on the transfer is the
    library's own grouped ncclSend / ncclRecv of one fixed-capacity block per rank plus a compaction kernel on
    rank 0 -- no count exchange, no host round trip, nothing allocated per call."""

    def __init__(self, device: int, cap_sent: int, cap_tok: int, group=None):
        import ctypes as C
        from . import _lib
        self._L = _lib.load()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device
        uid = np.zeros(128, np.uint8)
        if self.rank == 0:
            _lib.check(self._L.kp_gather_unique_id(uid.ctypes.data_as(C.c_void_p)))
        backend = dist.get_backend(group)
        t = torch.from_numpy(uid)
        if backend == "nccl":
            t = t.to("cuda:%d" % device)
        dist.broadcast(t, 0, group=group)
        uid = t.cpu().numpy()
        self._h = C.c_void_p()
        _lib.check(self._L.kp_gather_create(device, self.rank, self.world, uid.ctypes.data_as(C.c_void_p), int(cap_sent),
                                            int(cap_tok), C.byref(self._h)))

    def gather(self, mine):
        """mine: _lib.Result8 with DEVICE pointers (Tokenizer.tokenize_batch_device8).  -> Result8 on rank 0 (device
        pointers owned by the gather, valid until its next call), an empty Result8 elsewhere."""
        import ctypes as C
        from . import _lib
        out = _lib.Result8()
        _lib.check(self._L.kp_gather_tokens(self._h, C.byref(mine), C.byref(out)))
        return out

    def last_ms(self) -> float:
        import ctypes as C
        from . import _lib
        ms = C.c_float()
        _lib.check(self._L.kp_gather_last_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def to_host(self, r):
        """Gathered device result -> (tok_off u32, tokens TOKEN8_DTYPE, eos i32) numpy copies."""
        import ctypes as C
        from . import _lib
        n, nt = int(r.n_sent), int(r.n_tokens)
        tok_off = np.empty(n + 1, np.uint32)
        tokens = np.empty(nt, TOKEN8_DTYPE)
        eos = np.empty(n, np.int32)
        for dst, src in ((tok_off, r.tok_off), (tokens, r.tokens), (eos, r.eos_cost)):
            if dst.nbytes:
                _lib.check(self._L.kp_gather_copy_to_host(self._h, dst.ctypes.data_as(C.c_void_p), src, dst.nbytes))
        return tok_off, tokens, eos

    def close(self):
        if getattr(self, "_h", None):
            self._L.kp_gather_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def to_batch_result(tok_off: torch.Tensor, tokens8: torch.Tensor, eos_cost: torch.Tensor, offsets) -> BatchResult:
    """Gathered compact result (TokenGather.gather on dst) + the global sentence offsets -> BatchResult."""
    off = tok_off.cpu().numpy().astype(np.uint64)
    t8 = tokens8.cpu().numpy().view(TOKEN8_DTYPE)
    return BatchResult(off, expand_tokens8(off, t8, offsets), eos_cost.cpu().numpy().copy())


class ShardedTokenizer:
    """One rank's member of a data-parallel tokenizer group (one process per GPU).

    Rank `src` passes the `Dict`; the others pass None.  The packed dictionary travels by ONE
    broadcast; every rank (src included) then stages a private handle from the blob it holds in HBM
    (`kp_dict_create_from_device_blob`: validated by checksum, copied into memory the handle owns) and
    drops the receive buffer.  The caller's Dict is never modified.  `tokenize_global` takes the same
    global batch on every rank, tokenizes this rank's byte-balanced contiguous range and gathers the
    tokens on rank `dst`."""

    def __init__(self, dict=None, device: int = 0, src: int = 0, group=None):
        from . import Dict, Tokenizer
        self.group, self.device = group, device
        dev = torch.device("cuda:%d" % device)
        blob = dict.pack() if dist.get_rank(group) == src else None
        blob_t = broadcast_dict_blob(blob, src, dev, group)
        z8, z16 = np.zeros(0, np.uint8), np.zeros((0, 3), np.int16)
        own = Dict(da=np.zeros((0, 2), np.int32), dup_ids=np.zeros(0, np.int64), dup_counts=np.zeros(0, np.uint64),
                   morphs=z16, conn_row=0, conn_col=0, conn=np.zeros(0, np.int16), char_category=z8,
                   invoke_list=z8, group_list=z8, unk_cat=z8, unk_first_id=np.zeros(0, np.int64),
                   unk_count=np.zeros(0, np.uint64), unk_morphs=z16)
        own.attach_device_blob(blob_t.data_ptr(), blob_t.numel(), device)
        del blob_t                                   # the handle owns its own copy
        self.dict = own
        self.tokenizer = Tokenizer(own, device=device)
        self._gather = None

    def tokenize_shard(self, text: np.ndarray, offsets: np.ndarray):
        """-> (tok_off u32, tokens TOKEN8_DTYPE, eos i32) of this rank's sentences (compact form)."""
        return self.tokenizer.tokenize_batch8_bytes(text, offsets)

    def tokenize_global(self, text: np.ndarray, offsets: np.ndarray, dst: int = 0):
        from .corpus import shard_by_bytes
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        ranges = shard_by_bytes(offsets, world)
        s0, s1 = ranges[rank]
        off = np.ascontiguousarray(offsets[s0:s1 + 1])
        tok_off, tokens8, eos = self.tokenize_shard(text[int(off[0]):int(off[-1])], off - off[0])
        # capacity from the batch itself: a shard has at most bytes + sentences tokens
        cap_sent = max(b - a for a, b in ranges)
        cap_tok = max(int(offsets[b] - offsets[a]) + (b - a) for a, b in ranges)
        if self._gather is None or not (self._gather.cap_sent >= cap_sent and self._gather.cap_tok >= cap_tok):
            self._gather = TokenGather("cuda:%d" % self.device, cap_sent, cap_tok, dst, self.group)
        dev = self._gather.device
        g = self._gather.gather(torch.from_numpy(tok_off.view(np.uint8)).to(dev),
                                torch.from_numpy(tokens8.view(np.uint8)).to(dev),
                                torch.from_numpy(eos.view(np.uint8)).to(dev))
        return to_batch_result(*g, offsets) if g is not None else None


class Shards:
    """kp_shards_*: all visible GPUs (or `devices`) driven from ONE process -- what a Rust
    `Tokenizer::tokenize_batch` binds.  The dictionary is packed once and broadcast with NCCL; a batch is
    split into byte-balanced contiguous sentence ranges, one per GPU."""

    def __init__(self, dict, devices=None):
        import ctypes as C
        from . import _lib
        self._L = _lib.load()
        if devices is None:
            n = C.c_int()
            _lib.check(self._L.kp_device_count(C.byref(n)))
            devices = list(range(n.value))
        self.devices = list(devices)
        arr = (C.c_int * len(self.devices))(*self.devices)
        a, self._keep = dict._arrays()
        self._h = C.c_void_p()
        _lib.check(self._L.kp_shards_create(C.byref(a), arr, len(self.devices), C.byref(self._h)))

    def _call(self, fn, text, offsets):
        import ctypes as C
        from . import _lib
        buf = np.frombuffer(text, np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else np.ascontiguousarray(text, np.uint8)
        off = np.ascontiguousarray(offsets, np.uint64)
        r = _lib.Result8()
        _lib.check(fn(self._h, buf.ctypes.data_as(C.c_void_p) if buf.size else None, off.ctypes.data_as(C.c_void_p),
                      len(off) - 1, C.byref(r)))
        return r, off

    def tokenize(self, text, offsets) -> BatchResult:
        """Host text in, host result out (every GPU copies its tokens to their global offsets)."""
        from .tokenizer import copy_result8, result8_to_batch
        r, off = self._call(self._L.kp_shards_tokenize, text, offsets)
        return result8_to_batch(copy_result8(r), off)

    def tokenize_gather(self, text, offsets) -> BatchResult:
        """Same batch, result gathered in device memory of devices[0] over NVLink (ncclSend / ncclRecv),
        then copied to the host here for inspection."""
        import ctypes as C
        from . import _lib
        from .tokenizer import result8_to_batch
        r, off = self._call(self._L.kp_shards_tokenize_gather, text, offsets)
        n, nt = int(r.n_sent), int(r.n_tokens)
        tok_off = np.empty(n + 1, np.uint32)
        tokens = np.empty(nt, TOKEN8_DTYPE)
        eos = np.empty(n, np.int32)
        for dst, src in ((tok_off, r.tok_off), (tokens, r.tokens), (eos, r.eos_cost)):
            if dst.nbytes:
                _lib.check(self._L.kp_shards_copy_to_host(self._h, dst.ctypes.data_as(C.c_void_p), src, dst.nbytes))
        return result8_to_batch((tok_off, tokens, eos), off)

    def times(self) -> dict:
        import ctypes as C
        from . import _lib
        ms = (C.c_float * 4)()
        _lib.check(self._L.kp_shards_times(self._h, C.byref(ms)))
        return {"call_ms": ms[0], "slowest_pass_ms": ms[1], "nccl_gather_ms": ms[2], "dict_broadcast_ms": ms[3]}

    def close(self):
        if getattr(self, "_h", None):
            self._L.kp_shards_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
