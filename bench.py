#!/usr/bin/env python
"""bench.py — UTF-8 bytes/s tokenized (IPADIC, synthetic JA corpus) on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W                      # product arm (CUDA path)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --gpus N --steps K --warmup W     # restated reference CPU path

A step = one pass of the hot path (lattice build + Viterbi + back-trace, `Tokenizer::tokenize` per
sentence) over one batch: BASELINE.json configs[1], 65 536 synthetic sentences (mean 80 chars,
Zipf vocabulary over IPADIC) per GPU.  Weak scaling: every rank tokenizes its own 65 536-sentence
shard of an N x 65 536-sentence batch; the dictionary reaches ranks > 0 through one NCCL broadcast;
no collective runs inside the timed data path.

  value     input bytes of all ranks / device time (CUDA events on the library's stream around the
            whole step), text already resident in HBM; max over ranks
  e2e       same metric through the C-ABI `kp_tokenize_batch` with pinned HOST buffers: H2D of the
            text + offsets, kernels, D2H of the token records, wall clock around the call
  roofline  dominant kernel (kp_viterbi): algorithmic bytes 16 N + 8 E / its
            CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference: oracle/ref_tokenize.cpp (the reference's CPU algorithm restated in
            C++, same shape: per-node heap strings, vector-of-vector buckets) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "UTF-8 bytes/sec tokenized (IPADIC, synthetic JA corpus)"
WORKLOADS = {"cfg2": ("cfg2", 65536), "cfg3": ("cfg3", 1000000), "cfg4": ("cfg4", 4096)}
WORKLOAD_DESC = {
    "cfg2": "BASELINE.json configs[1]: 65536 synthetic JA sentences per GPU, mean 80 chars, Zipf vocab, IPADIC",
    "cfg3": "BASELINE.json configs[2]: 1M Wikipedia-shape synthetic sentences per GPU, IPADIC",
    "cfg4": "BASELINE.json configs[3]: 4096 sentences x 4096 chars per GPU (deep lattice), IPADIC",
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region: NVML polled every ~5 ms from a thread (the
    timed region is tens of milliseconds, too short for `nvidia-smi -lms`); same fields as the
    B200_PROFILING.md recipe."""

    def __init__(self, gpus):
        self.gpus = list(gpus)
        self.samples = []          # (sm_mhz, reasons bitmask) per poll per gpu
        self.max_mhz = None
        self._stop = threading.Event()
        self._th = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.handles = [nv.nvmlDeviceGetHandleByIndex(i) for i in self.gpus]
            self.max_mhz = max(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM) for h in self.handles)
        except Exception:
            self.nv = None
            return

        def poll():
            nv = self.nv
            while not self._stop.is_set():
                for h in self.handles:
                    try:
                        self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                                             nv.nvmlDeviceGetCurrentClocksEventReasons(h)))
                    except Exception:
                        pass
                time.sleep(0.004)

        self._th = threading.Thread(target=poll, daemon=True)
        self._th.start()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if self._th is None:
            return out
        self._stop.set()
        self._th.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        if self.samples:
            sm = sorted(x[0] for x in self.samples)
            # "under load": the upper half of the samples (the GPU idles between steps while the host flushes L2)
            out.update(sm_mhz=float(statistics.median(sm[len(sm) // 2:])), samples=len(sm))
            bits = 0
            for _, r in self.samples:
                bits |= int(r)
            out["reasons"] = sorted(k for k, v in names.items() if bits & v)
        return out


def host_info():
    model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return len(os.sched_getaffinity(0)), model


def load_dict_and_corpus(kind, n_sent, seed_offset, rank=0, world=1, barrier=None):
    from kanpyo_b200 import builder, corpus
    if world > 1 and rank != 0:
        barrier()                       # rank 0 builds (or loads) the cache first
    d = builder.ipadic()
    if world > 1 and rank == 0:
        barrier()
    vocab = corpus.Vocabulary(d.keywords, d.morphs)
    text, off = corpus.synth_corpus(vocab, n_sent, kind, seed=corpus.SEED + seed_offset)
    return d, text, off


# ----------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm restated (oracle/ref_tokenize.cpp) on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_sample(otk, text, off, n, threads):
    n = min(n, len(off) - 1)
    t0 = time.perf_counter()
    otk.tokenize_batch(text[:int(off[n])], off[:n + 1], threads=threads, collect=False)
    return time.perf_counter() - t0, int(off[n])


def cpu_pick_sample(otk, text, off, threads, target_s):
    """Sentences for ~target_s of wall time with `threads` threads (calibrated on a small slice)."""
    n = min(2048, len(off) - 1)
    for _ in range(2):      # second pass re-calibrates at the chosen size (thread start-up, caches)
        dt, nbytes = cpu_sample(otk, text, off, n, threads)
        want = nbytes / max(dt, 1e-6) * target_s
        n = int(np.searchsorted(off, np.uint64(min(want, float(off[-1]))), side="right")) - 1
        n = max(256, min(n, len(off) - 1))
    return n


def cpu_baseline(text, off, target_s=12.0):
    from oracle import oracle
    otk = oracle.OracleTokenizer(oracle.load_ipadic())
    cores, model = host_info()
    n = cpu_pick_sample(otk, text, off, cores, target_s / 2)
    best = None
    for _ in range(2):
        dt, nbytes = cpu_sample(otk, text, off, n, cores)
        best = dt if best is None else min(best, dt)
    dt1, nb1 = cpu_sample(otk, text, off, max(256, n // max(cores, 1)), 1)
    return {"value": nbytes / best, "unit": "bytes/s", "cores": cores, "kind": "port",
            "sample": "first %d sentences (%d bytes) of the step's batch, %d std::threads, best of 2; "
                      "oracle/ref_tokenize.cpp (C++ restatement of the reference's CPU path, g++ -O2); "
                      "the reference is Rust and cannot be built here" % (n, nbytes, cores),
            "single_thread_value": nb1 / dt1, "cpu_model": model}


def reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    from oracle import oracle
    from kanpyo_b200 import corpus
    kind, n_sent = WORKLOADS[args.workload]
    n_sent = args.sentences or n_sent
    od = oracle.load_ipadic()
    otk = oracle.OracleTokenizer(od)
    vocab = corpus.Vocabulary(od.keywords, od.morphs)
    text, off = corpus.synth_corpus(vocab, n_sent, kind, seed=corpus.SEED)
    cores, model = host_info()
    # bound the whole run to a few minutes: ~2 s per step
    per_step = min(2.0, 150.0 / max(1, args.steps + args.warmup))
    n = cpu_pick_sample(otk, text, off, cores, per_step)
    for _ in range(args.warmup):
        cpu_sample(otk, text, off, n, cores)
    times = []
    nbytes = 0
    for _ in range(args.steps):
        dt, nbytes = cpu_sample(otk, text, off, n, cores)
        times.append(dt)
    total = sum(times)
    value = nbytes * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "bytes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "sentences_per_gpu": n_sent,
                   "l2": "n/a (CPU arm)"},
        "cpu_baseline": {"value": value, "unit": "bytes/s", "cores": cores, "kind": "port", "cpu_model": model,
                         "sample": "each step = first %d sentences (%d bytes) of the workload, %d std::threads; "
                                   "oracle/ref_tokenize.cpp (C++ restatement; the Rust reference cannot be built "
                                   "in this image)" % (n, nbytes, cores)},
        "e2e": {"value": value, "unit": "bytes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------
# product arm
# ----------------------------------------------------------------------------------------------------
def product_arm(args):
    import torch
    import torch.distributed as dist

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product arm has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda:%d" % local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import kanpyo_b200
    from kanpyo_b200 import sharded

    kind, n_sent = WORKLOADS[args.workload]
    n_sent = args.sentences or n_sent
    d, text, off = load_dict_and_corpus(kind, n_sent, rank, rank, world, barrier)

    # ---- dictionary: staged to HBM once; ranks > 0 receive the packed blob by ONE NCCL broadcast ----
    bcast_ms = None
    if world > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        blob = d.pack() if rank == 0 else None
        warm = torch.zeros(1, device=dev)
        dist.all_reduce(warm)              # NCCL communicator setup is not part of the broadcast
        barrier()
        e0.record()
        blob_t = sharded.broadcast_dict_blob(blob, 0, dev)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        d.attach_device_blob(blob_t.data_ptr(), blob_t.numel(), local)
    tk = kanpyo_b200.Tokenizer(d, device=local)
    if args.chunk_mib:
        tk.set_chunk_bytes(args.chunk_mib << 20)

    n_bytes = int(off[-1])
    S = len(off) - 1
    off_i64 = np.ascontiguousarray(off).view(np.int64)
    d_text = torch.from_numpy(text.copy()).to(dev)
    d_off = torch.from_numpy(off_i64.copy()).to(dev)
    h_text = torch.from_numpy(text.copy()).pin_memory()
    h_off = torch.from_numpy(off_i64.copy()).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def step_device():
        flush.zero_()
        torch.cuda.synchronize()
        r = tk.tokenize_batch_device(d_text.data_ptr(), d_off.data_ptr(), S, 0, n_bytes)
        return r, tk.profile()

    def step_e2e():
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = tk.tokenize_batch_ptr(h_text.data_ptr(), h_off.data_ptr(), S)
        return r, time.perf_counter() - t0

    # exact work counters (outside any timed region)
    tk.set_count_work(True)
    step_device()
    ctr = tk.counters()
    tk.set_count_work(False)

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(range(world)) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    w0 = time.perf_counter()
    dev_ms, stages, launches = [], {}, 0
    for _ in range(args.steps):
        r, p = step_device()
        dev_ms.append(p["total_ms"])
        launches += p["kernel_launches"]
        for k in ("prep_ms", "lattice_ms", "bucket_ms", "viterbi_ms", "backtrace_ms"):
            stages[k] = stages.get(k, 0.0) + p[k]
    barrier()
    wall_s = time.perf_counter() - w0
    n_tokens = int(r.n_tokens)

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    e2e_s = []
    for _ in range(args.steps):
        r2, dt = step_e2e()
        e2e_s.append(dt)
    barrier()
    clocks = sampler.stop() if sampler else None
    assert int(r2.n_tokens) == n_tokens, "host and device entry points disagree on the token count"

    # ---- token gather to rank 0 over NVLink (north_star's second collective; outside the timed path) --
    gather_ms = None
    if world > 1:
        r = tk.tokenize_batch_device(d_text.data_ptr(), d_off.data_ptr(), S, 0, n_bytes)
        t_off = sharded.device_view(r.tok_off, 8 * (S + 1), local).view(torch.int64)
        t_tok = sharded.device_view(r.tokens, 16 * int(r.n_tokens), local)
        t_eos = sharded.device_view(r.eos_cost, 4 * S, local).view(torch.int32)
        sharded.gather_results(t_off, t_tok, t_eos)             # warm the P2P channels
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        g = sharded.gather_results(t_off, t_tok, t_eos)
        e1.record()
        torch.cuda.synchronize()
        gather_ms = e0.elapsed_time(e1)
        if rank == 0:
            assert g[2].numel() == S * world

    # ---- reduce over ranks: bytes summed, times max ------------------------------------------------
    vals = torch.tensor([sum(dev_ms), sum(e2e_s) * 1e3, wall_s * 1e3, stages["viterbi_ms"], stages["lattice_ms"],
                         gather_ms or 0.0], dtype=torch.float64, device=dev)
    h2d_b = int(h_text.numel() + 8 * h_off.numel())
    d2h_b = int(16 * n_tokens + 8 * (S + 1) + 4 * S)
    sums = torch.tensor([n_bytes, n_tokens, launches, h2d_b, d2h_b], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dev_total_ms, e2e_total_ms, wall_ms, vit_ms, lat_ms, gather_max = vals.tolist()
    world_bytes, world_tokens, world_launches, world_h2d, world_d2h = sums.tolist()

    if rank == 0:
        K = args.steps
        peak, peak_src = measured_peak()
        # algorithmic bytes of the dominant kernel (kp_viterbi), this rank's launch: DESIGN.md section 5
        N_nodes, E_pairs = ctr["nodes"], ctr["pairs"]
        a_vit = 16 * N_nodes + 8 * E_pairs
        a_total = (ctr["bytes"] + ctr["chars"] + 8 * (ctr["probes"] + ctr["probes_ok"]) + 40 * N_nodes + 8 * E_pairs
                   + 20 * ctr["tokens"])
        vit_launch_ms = stages["viterbi_ms"] / K
        achieved = a_vit / (vit_launch_ms * 1e-3) / 1e9
        kern_ms = sum(stages.values()) / K
        traffic, traffic_src = None, None
        try:   # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                tj = next(v for k, v in json.load(f).items() if k.startswith("kp_viterbi"))   # kp_viterbi<lanes>
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": world_bytes * K / (dev_total_ms * 1e-3), "unit": "bytes/s",
            "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": dev_total_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "sentences_per_gpu": S, "bytes_per_gpu": n_bytes,
                       "parallelism": "dp%d (independent sentence shards, no data-path collective)" % world,
                       "l2": "flushed between steps (256 MiB memset outside the timed events); per-step scratch "
                             "(%.2f GB of node records / buckets) also exceeds the 126 MB L2"
                             % ((N_nodes * 34 + ctr["chars"] * 36) / 1e9),
                       "timing": "CUDA events on the library stream around each whole step, summed over steps, "
                                 "max over ranks; wall_ms_per_step includes the L2 flushes"},
            "wall_ms_per_step": wall_ms / K,
            "e2e": {"value": world_bytes * K / (e2e_total_ms * 1e-3), "unit": "bytes/s",
                    "h2d_bytes_per_step": int(world_h2d), "d2h_bytes_per_step": int(world_d2h),
                    "ms_per_step": e2e_total_ms / K,
                    "api": "kp_tokenize_batch (C ABI) on pinned host buffers, wall clock around the call"},
            "gpu_launches": int(world_launches),
            "roofline": {"bound": "hbm", "kernel": "kp_viterbi", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": a_vit, "launch_ms": vit_launch_ms,
                         "share_of_kernel_time": vit_launch_ms / kern_ms,
                         "whole_path": {"algorithmic_bytes": a_total, "bytes_per_input_byte": a_total / ctr["bytes"],
                                        "kernel_ms": kern_ms, "achieved": a_total / (kern_ms * 1e-3) / 1e9,
                                        "frac": a_total / (kern_ms * 1e-3) / 1e9 / peak}},
            "stages_ms_per_step": {k: v / K for k, v in stages.items()},
            "counters": ctr,
            "clocks": clocks,
        }
        if world > 1:
            line["dict_broadcast_ms"] = bcast_ms
            line["token_gather_ms"] = gather_max
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(text, off)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--sentences", type=int, default=0, help="override sentences per GPU (debug)")
    ap.add_argument("--chunk-mib", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)
    return product_arm(args)


if __name__ == "__main__":
    sys.exit(main())
