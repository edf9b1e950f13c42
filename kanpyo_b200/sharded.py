"""Multi-GPU sharding of the tokenizer path: one process per GPU, `torch.distributed` for the plumbing.

Sentences are independent and the dictionary is read-only, so the path has no exchange step inside
the algorithm.  Exactly two collectives exist, both outside the kernels:

  * `broadcast_dict_blob` — ONE broadcast of the packed dictionary blob (17.5 MB for IPADIC) from the
    rank that built it; every other rank stages its HBM copy straight from the receive buffer
    (`kp_dict_create_from_device_blob`).
  * `gather_results`      — tokens of all shards to one rank, in global sentence order: an
    all-gather of the per-rank (sentences, tokens) counts followed by grouped point-to-point sends of
    the exact payloads (no padding).

`shard_by_bytes` (corpus.py) gives the contiguous, byte-balanced sentence ranges.  Backend `nccl`
moves device tensors over NVLink; backend `gloo` (CPU tensors) runs the same host logic in tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .tokenizer import BatchResult, TOKEN_DTYPE


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


def gather_results(tok_off: torch.Tensor, tokens: torch.Tensor, eos_cost: torch.Tensor, dst: int = 0, group=None):
    """Per-rank shard results -> on rank `dst` the concatenation in rank order (= global sentence order
    when shards are contiguous ranges), else None.

    tok_off  int64/uint64-as-int64 [n_sent+1], shard-relative (tok_off[0] == 0)
    tokens   uint8 [n_tokens*16] (kp_token records)
    eos_cost int32 [n_sent]
    All three live on the same device (cuda for nccl, cpu for gloo).  Returns (tok_off, tokens, eos_cost)
    tensors on that device."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = tokens.device
    n_sent = eos_cost.numel()
    n_tok = tokens.numel() // TOKEN_DTYPE.itemsize
    mine = torch.tensor([n_sent, n_tok], dtype=torch.int64, device=dev)
    counts = torch.empty(world, 2, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts.view(-1), mine, group=group)
    counts = counts.cpu()
    if world == 1:
        return tok_off.clone(), tokens.clone(), eos_cost.clone()
    sent_base = torch.cumsum(counts[:, 0], 0) - counts[:, 0]
    tok_base = torch.cumsum(counts[:, 1], 0) - counts[:, 1]
    body = tok_off[1:].contiguous().view(torch.int64)     # tok_off[0] is implied by the base
    if rank == dst:
        S, T = int(counts[:, 0].sum()), int(counts[:, 1].sum())
        g_off = torch.zeros(S + 1, dtype=torch.int64, device=dev)
        g_tok = torch.empty(T * TOKEN_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        g_eos = torch.empty(S, dtype=torch.int32, device=dev)
        ops = []
        for r in range(world):
            s0, s1 = int(sent_base[r]), int(sent_base[r] + counts[r, 0])
            t0, t1 = int(tok_base[r]) * 16, int(tok_base[r] + counts[r, 1]) * 16
            if r == rank:
                g_off[1 + s0:1 + s1] = body
                g_tok[t0:t1] = tokens
                g_eos[s0:s1] = eos_cost
                continue
            peer = dist.get_global_rank(group, r) if group is not None else r
            if s1 > s0:
                ops.append(dist.P2POp(dist.irecv, g_off[1 + s0:1 + s1], peer, group))
                ops.append(dist.P2POp(dist.irecv, g_eos[s0:s1], peer, group))
            if t1 > t0:
                ops.append(dist.P2POp(dist.irecv, g_tok[t0:t1], peer, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        # rebase every shard's offsets by the tokens that precede it
        for r in range(world):
            s0, s1 = int(sent_base[r]), int(sent_base[r] + counts[r, 0])
            g_off[1 + s0:1 + s1] += int(tok_base[r])
        return g_off, g_tok, g_eos
    peer = dist.get_global_rank(group, dst) if group is not None else dst
    ops = []
    if n_sent:
        ops.append(dist.P2POp(dist.isend, body, peer, group))
        ops.append(dist.P2POp(dist.isend, eos_cost.contiguous(), peer, group))
    if n_tok:
        ops.append(dist.P2POp(dist.isend, tokens.contiguous(), peer, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return None


def to_batch_result(tok_off: torch.Tensor, tokens: torch.Tensor, eos_cost: torch.Tensor) -> BatchResult:
    return BatchResult(tok_off.cpu().numpy().view(np.uint64).copy(),
                       tokens.cpu().numpy().view(TOKEN_DTYPE).copy(), eos_cost.cpu().numpy().copy())


class ShardedTokenizer:
    """One rank's member of a data-parallel tokenizer group.

    Rank `src` passes the `Dict`; the others pass None and receive the packed blob by broadcast.
    `tokenize_shard` runs the CUDA path on this rank's sentences; `tokenize_global` takes the same
    global batch on every rank, tokenizes this rank's byte-balanced contiguous range and gathers the
    tokens on rank `dst`."""

    def __init__(self, dict=None, device: int = 0, src: int = 0, group=None):
        from . import Dict, Tokenizer
        self.group, self.device = group, device
        dev = torch.device("cuda:%d" % device)
        blob = dict.pack() if dist.get_rank(group) == src else None
        self.blob = broadcast_dict_blob(blob, src, dev, group)
        if dict is None:
            z8, z16 = np.zeros(0, np.uint8), np.zeros((0, 3), np.int16)
            dict = Dict(da=np.zeros((0, 2), np.int32), dup_ids=np.zeros(0, np.int64), dup_counts=np.zeros(0, np.uint64),
                        morphs=z16, conn_row=0, conn_col=0, conn=np.zeros(0, np.int16), char_category=z8,
                        invoke_list=z8, group_list=z8, unk_cat=z8, unk_first_id=np.zeros(0, np.int64),
                        unk_count=np.zeros(0, np.uint64), unk_morphs=z16)
        dict.attach_device_blob(self.blob.data_ptr(), self.blob.numel(), device)
        self.dict = dict
        self.tokenizer = Tokenizer(dict, device=device)

    def tokenize_shard(self, text: np.ndarray, offsets: np.ndarray) -> BatchResult:
        return self.tokenizer.tokenize_batch_bytes(text, offsets)

    def tokenize_global(self, text: np.ndarray, offsets: np.ndarray, dst: int = 0):
        from .corpus import shard_by_bytes
        rank, world = dist.get_rank(self.group), dist.get_world_size(self.group)
        s0, s1 = shard_by_bytes(offsets, world)[rank]
        res = self.tokenize_shard(text, np.ascontiguousarray(offsets[s0:s1 + 1]))
        dev = torch.device("cuda:%d" % self.device)
        g = gather_results(torch.from_numpy(res.tok_off.view(np.int64)).to(dev),
                           torch.from_numpy(res.tokens.view(np.uint8)).to(dev),
                           torch.from_numpy(res.eos_cost).to(dev), dst, self.group)
        return to_batch_result(*g) if g is not None else None
