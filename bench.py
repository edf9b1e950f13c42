#!/usr/bin/env python
"""bench.py — UTF-8 bytes/s tokenized (IPADIC, synthetic JA corpus) on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W                      # product arm (CUDA path)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --gpus N --steps K --warmup W     # restated reference CPU path

A step = one pass of the hot path (lattice build + Viterbi + back-trace, `Tokenizer::tokenize` per
sentence) over one batch: BASELINE.json configs[1], 65 536 synthetic sentences (mean 80 chars,
Zipf vocabulary over IPADIC) per GPU.  The main line is weak scaling: every rank tokenizes its own
65 536-sentence batch; the dictionary reaches ranks > 0 through one NCCL broadcast; at N > 1 every
timed step ends with the NCCL gather of the token records on rank 0.  `strong_scaling` (N > 1) is
BASELINE.json configs[4]: ONE 65 536-sentence batch split into byte-balanced shards, gather included.

  value     input bytes of all ranks / device time (CUDA events on the library's stream around the
            whole step, plus the gather's events at N > 1), text already resident in HBM; max over ranks
  e2e       same metric through the C-ABI queue (`kp_queue_submit` / `kp_queue_wait`) with pinned HOST
            buffers: H2D of the text + offsets, kernels, D2H of the compact token records, all inside
            the wall clock, successive steps overlapped; `sync_call` is the blocking `kp_tokenize_batch8`
  parity    the benched batch's result (device path, e2e path, gathered results at N > 1) compared
            bit-for-bit with the oracle on the same bytes; a mismatch fails the run (exit code 3)
  roofline  dominant kernel: algorithmic bytes / its CUDA-event duration, against
            MEASURED_PEAKS.json hbm_gbs
  latency_us  one sentence per call (`kp_tokenize`), p50 / p99, beside the CPU port's time
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


def bind_to_gpu_numa_node(gpu_index):
    """What `numactl --cpunodebind --membind` does for a one-process-per-GPU launch: run this rank (and first-touch its
    pinned buffers) on the CPUs next to its GPU, so the H2D / D2H of eight ranks do not all cross the socket link.
    Topology from NVML (the GPU's ideal CPU set), else sysfs (the PCI device's NUMA node); any failure leaves the
    affinity untouched.  Returns a description or None."""
    allowed = os.sched_getaffinity(0)
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(gpu_index)
        try:
            words = nv.nvmlDeviceGetCpuAffinity(h, (max(allowed) // 64) + 1)
            cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1} & allowed
            if 2 <= len(cpus) < len(allowed):
                os.sched_setaffinity(0, cpus)
                return "NVML cpu affinity (%d of %d cpus)" % (len(cpus), len(allowed))
        except Exception:   # noqa: BLE001
            pass
        bus = nv.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:            # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= allowed
        if len(cpus) < 2:
            return None
        os.sched_setaffinity(0, cpus)
        return "numa node %d (%d cpus)" % (node, len(cpus))
    except Exception:   # noqa: BLE001
        return None


def load_dict_and_corpus(kind, n_sent, seed_offset, rank=0, world=1, barrier=None):
    """The product's dictionary and the synthetic batch of rank `seed_offset` (seed = corpus.SEED + rank)."""
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
def oracle_reference(text, off, threads):
    """The oracle's answer for a batch: (tok_off, tokens int64[n,6], eos_cost).  Checker only."""
    from oracle import oracle
    otk = oracle.OracleTokenizer(oracle.load_ipadic())
    o_off, o_tok, o_cost, _ = otk.tokenize_batch(text, off, threads=threads)
    return o_off, o_tok, o_cost


def compare_with_oracle(ref, tok_off, tokens16, eos):
    """Bit-exact comparison of a (tok_off, kp_token records, dp[EOS]) result with the oracle's."""
    o_off, o_tok, o_cost = ref
    if len(tok_off) != len(o_off) or not np.array_equal(np.asarray(tok_off, np.uint64), o_off):
        return "token offsets differ"
    if not np.array_equal(np.asarray(eos, np.int32), o_cost):
        return "dp[EOS] differs"
    t = tokens16
    for name, col in (("id", 0), ("cls", 1), ("position", 2), ("start", 3)):
        if not np.array_equal(t[name].astype(np.int64), o_tok[:, col]):
            return "token %s differs" % name
    if not np.array_equal(t["start"].astype(np.int64) + t["char_len"], o_tok[:, 4]):
        return "token end differs"
    return None


def latency_probe(tk, otk, sentences, reps=200):
    """Single-sentence latency of kp_tokenize (the reference CLI's call pattern, src/bin/kanpyo.rs:115-122:
    one line per call) beside the CPU port's time for the same sentence."""
    out = {}
    for name, s in sentences.items():
        b = s.encode("utf-8")
        off = np.array([0, len(b)], np.uint64)
        for _ in range(20):
            tk.tokenize_batch_bytes(b, off)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            tk.tokenize_batch_bytes(b, off)
            ts.append((time.perf_counter() - t0) * 1e6)
        ts.sort()
        cpu = None
        if otk is not None:
            buf = np.frombuffer(b, np.uint8)
            t0 = time.perf_counter()
            for _ in range(reps):
                otk.tokenize_batch(buf, off, threads=1, collect=False)
            cpu = (time.perf_counter() - t0) * 1e6 / reps
        out[name] = {"chars": len(s), "p50": ts[len(ts) // 2], "p99": ts[min(len(ts) - 1, int(len(ts) * 0.99))],
                     "cpu_port_us": cpu}
    return out


def product_arm(args):
    import torch
    import torch.distributed as dist

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product arm has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda:%d" % local)
    numa = bind_to_gpu_numa_node(local) if world > 1 and not args.no_numa else None
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(minutes=4))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import kanpyo_b200
    from kanpyo_b200 import corpus, sharded
    from kanpyo_b200.tokenizer import Queue, copy_result8, expand_tokens8

    kind, n_sent = WORKLOADS[args.workload]
    n_sent = args.sentences or n_sent
    d, text, off = load_dict_and_corpus(kind, n_sent, rank, rank, world, barrier)

    # ---- dictionary: staged to HBM once; ranks > 0 receive the packed blob by ONE NCCL broadcast ----
    bcast_ms = None
    if world > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        blob = d.pack() if rank == 0 else None
        warm = torch.zeros(1, device=dev)
        dist.all_reduce(warm)              # NCCL communicator / channel setup is not part of the broadcast
        dist.broadcast(torch.zeros(1 << 20, dtype=torch.uint8, device=dev), 0)
        barrier()
        e0.record()
        blob_t = sharded.broadcast_dict_blob(blob, 0, dev)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        d.attach_device_blob(blob_t.data_ptr(), blob_t.numel(), local)
        del blob_t
    tk = kanpyo_b200.Tokenizer(d, device=local)
    if args.chunk_mib:
        tk.set_chunk_bytes(args.chunk_mib << 20)
    if args.path != "auto":
        tk.set_path(args.path)

    n_bytes = int(off[-1])
    S = len(off) - 1
    off_i64 = np.ascontiguousarray(off).view(np.int64)
    d_text = torch.from_numpy(text.copy()).to(dev)
    d_off = torch.from_numpy(off_i64.copy()).to(dev)
    h_text = torch.from_numpy(text.copy()).pin_memory()
    h_off = torch.from_numpy(off_i64.copy()).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    # ---- token gather to rank 0 (north_star's second collective), INSIDE every timed step at N > 1 ----
    gather = None
    if world > 1:      # capacities are part of the block layout: the same on every rank (max over the ranks' batches)
        caps = torch.tensor([S, int(n_bytes * 0.14) + S], dtype=torch.int64, device=dev)     # 0.12 tokens per input byte here
        dist.all_reduce(caps, op=dist.ReduceOp.MAX)
        gather = sharded.NcclGather(local, int(caps[0]), int(caps[1]))

    def step_device():
        """-> (result, device ms of the step: library stream events + the gather's events on torch's stream)."""
        flush.zero_()
        torch.cuda.synchronize()
        r = tk.tokenize_batch_device8(d_text.data_ptr(), d_off.data_ptr(), S, 0, n_bytes)
        p = tk.profile()
        g, gms = None, 0.0
        if gather is not None:
            g = gather.gather(r)
            gms = gather.last_ms()
        return r, p, g, gms

    # exact work counters (outside any timed region)
    tk.set_count_work(True)
    tk.set_path("pipeline")               # the counting kernels belong to the multi-kernel pipeline
    step_device()
    ctr = tk.counters()
    tk.set_count_work(False)
    tk.set_path(args.path)

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(range(world)) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    w0 = time.perf_counter()
    dev_ms, gather_ms_sum, stages, launches = [], 0.0, {}, 0
    fused_sentences = 0
    g_last = None
    for _ in range(args.steps):
        r, p, g_last, gms = step_device()
        dev_ms.append(p["total_ms"] + gms)
        gather_ms_sum += gms
        launches += p["kernel_launches"]
        fused_sentences = p.get("fused_sentences", 0)
        for k in ("prep_ms", "lattice_ms", "bucket_ms", "viterbi_ms", "backtrace_ms", "fused_ms"):
            stages[k] = stages.get(k, 0.0) + p.get(k, 0.0)
    barrier()
    wall_s = time.perf_counter() - w0
    n_tokens = int(r.n_tokens)
    dev_result = tk.copy_device_result8(r)

    # ---- e2e: the asynchronous queue (kp_queue_*), pinned HOST buffers in, pinned HOST result out; every
    # step's H2D and D2H are inside the timed region, overlapped with the kernels of the neighbouring steps
    q = Queue(d, device=local, depth=args.queue_depth)
    q.set_path(args.path)
    # measured on 8 GPUs / 32 cores (24 queue workers): spinning 50.0 GB/s, sleeping on blocking-sync events 41.7 GB/s;
    # on 1 GPU 8.4 vs 7.5 GB/s -- the workers spin unless told otherwise
    blocking = args.blocking_sync == "on"
    q.set_blocking_sync(blocking)

    def run_queue(qq, k, text_ptr, off_ptr, ns):
        """k steps through the queue, `depth` in flight: submit step i + depth only after step i was waited for
        (its result buffers are the ones step i + depth reuses)."""
        pending, last = [], None
        for _ in range(k):
            if len(pending) == qq.depth:
                last = qq.wait_raw(pending.pop(0))
            pending.append(qq.submit_ptr(text_ptr, off_ptr, ns))
        for t in pending:
            last = qq.wait_raw(t)
        return last

    run_queue(q, max(args.warmup, 2 * args.queue_depth), h_text.data_ptr(), h_off.data_ptr(), S)
    barrier()
    t0 = time.perf_counter()
    r_e2e = run_queue(q, args.steps, h_text.data_ptr(), h_off.data_ptr(), S)
    torch.cuda.synchronize()
    e2e_total_s = time.perf_counter() - t0
    barrier()
    e2e_result = copy_result8(r_e2e)
    # what a caller that wants the 16-byte kp_token records (absolute position / start) pays on top: kp_expand_tokens8
    # over the step's pinned result, up to 8 host threads (the Rust shim does this while it builds its Vec<Token>)
    import ctypes as C
    from kanpyo_b200 import _lib as kp_lib
    from kanpyo_b200.tokenizer import TOKEN_DTYPE
    expanded = np.empty(int(r_e2e.n_tokens), TOKEN_DTYPE)
    expand_ms = None
    for _ in range(3):
        t0 = time.perf_counter()
        kp_lib.check(kp_lib.load().kp_expand_tokens8(C.byref(r_e2e), C.c_void_p(h_off.data_ptr()), C.c_void_p(expanded.ctypes.data)))
        dt = (time.perf_counter() - t0) * 1e3
        expand_ms = dt if expand_ms is None else min(expand_ms, dt)
    # the synchronous call (one batch at a time, nothing overlapped), for comparison
    for _ in range(args.warmup):
        tk.tokenize_batch8_ptr(h_text.data_ptr(), h_off.data_ptr(), S)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tk.tokenize_batch8_ptr(h_text.data_ptr(), h_off.data_ptr(), S)
    sync_total_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None
    q.close()

    # ---- strong scaling (BASELINE.json configs[4]): ONE batch of n_sent sentences split into byte-balanced
    # contiguous shards; shard text resident in HBM (the scatter happens once, outside); every timed step =
    # tokenize the shard + gather the tokens on rank 0
    strong = None
    g_strong = None
    if world > 1:
        _, text0, off0 = load_dict_and_corpus(kind, n_sent, 0)          # every rank generates rank 0's batch
        ranges = corpus.shard_by_bytes(off0, world)
        s0, s1 = ranges[rank]
        sh_off = np.ascontiguousarray(off0[s0:s1 + 1] - off0[s0]).view(np.int64)
        sh_text = text0[int(off0[s0]):int(off0[s1])]
        ds_text = torch.from_numpy(sh_text.copy()).to(dev)
        ds_off = torch.from_numpy(sh_off.copy()).to(dev)
        hs_text = torch.from_numpy(sh_text.copy()).pin_memory()
        hs_off = torch.from_numpy(sh_off.copy()).pin_memory()
        Ss, Bs = s1 - s0, int(sh_off[-1])
        cap_s = max(b - a for a, b in ranges)
        cap_t = max(int((off0[b] - off0[a]) * 0.14) + (b - a) for a, b in ranges)
        sg = sharded.NcclGather(local, cap_s, cap_t)

        def step_strong():
            flush.zero_()
            torch.cuda.synchronize()
            rr = tk.tokenize_batch_device8(ds_text.data_ptr(), ds_off.data_ptr(), Ss, 0, Bs)
            pm = tk.profile()["total_ms"]
            gg = sg.gather(rr)
            return gg, pm, sg.last_ms()

        for _ in range(args.warmup):
            step_strong()
        barrier()
        sm, sgm = 0.0, 0.0
        for _ in range(args.steps):
            g_strong, pm, gm = step_strong()
            sm += pm + gm
            sgm += gm
        barrier()
        # e2e of the strong split: each rank's queue on its shard, host buffers in and out
        q2 = Queue(d, device=local, depth=args.queue_depth)
        q2.set_path(args.path)
        q2.set_blocking_sync(blocking)
        run_queue(q2, 2 * args.queue_depth, hs_text.data_ptr(), hs_off.data_ptr(), Ss)
        barrier()
        t0 = time.perf_counter()
        run_queue(q2, args.steps, hs_text.data_ptr(), hs_off.data_ptr(), Ss)
        strong_e2e_s = time.perf_counter() - t0
        barrier()
        q2.close()
        sv = torch.tensor([sm, sgm, strong_e2e_s * 1e3], dtype=torch.float64, device=dev)
        dist.all_reduce(sv, op=dist.ReduceOp.MAX)
        sm, sgm, strong_e2e_ms = sv.tolist()
        strong = {"sentences": int(len(off0) - 1), "bytes": int(off0[-1]), "value": int(off0[-1]) * args.steps / (sm * 1e-3),
                  "unit": "bytes/s", "ms_per_step": sm / args.steps, "token_gather_ms": sgm / args.steps,
                  "e2e_value": int(off0[-1]) * args.steps / (strong_e2e_ms * 1e-3),
                  "what": "BASELINE.json configs[4]: one batch split into %d byte-balanced contiguous shards, shard text "
                          "resident in HBM; every timed step = tokenize the shard + NCCL gather of the token records on "
                          "rank 0 (kp_gather_tokens over NVLink); device time, max over ranks.  e2e_value: "
                          "every rank's kp_queue on its shard from pinned host memory to pinned host memory" % world}

    # ---- reduce over ranks: bytes summed, times max ------------------------------------------------
    vals = torch.tensor([sum(dev_ms), e2e_total_s * 1e3, wall_s * 1e3, stages["viterbi_ms"], stages["lattice_ms"],
                         gather_ms_sum, sync_total_s * 1e3], dtype=torch.float64, device=dev)
    h2d_b = int(h_text.numel() + 8 * h_off.numel())
    d2h_b = int(8 * n_tokens + 4 * (S + 1) + 4 * S)
    sums = torch.tensor([n_bytes, n_tokens, launches, h2d_b, d2h_b], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dev_total_ms, e2e_total_ms, wall_ms, vit_ms, lat_ms, gather_total_ms, sync_total_ms = vals.tolist()
    world_bytes, world_tokens, world_launches, world_h2d, world_d2h = sums.tolist()

    # ---- parity gate: every number above is reported only with its result checked against the oracle --
    parity = {"checked": False}
    if not args.no_parity:
        cores, _ = host_info()
        ref = oracle_reference(text, off, cores)
        checks = {}

        def guarded(fn):          # a malformed result must fail the comparison, not leave the other ranks waiting
            try:
                return fn()
            except Exception as e:   # noqa: BLE001
                return "comparison raised %s: %s" % (type(e).__name__, e)

        checks["device"] = guarded(lambda: compare_with_oracle(
            ref, dev_result[0], expand_tokens8(dev_result[0], dev_result[1], off), dev_result[2]))
        checks["e2e_queue"] = guarded(lambda: compare_with_oracle(
            ref, e2e_result[0], expand_tokens8(e2e_result[0], e2e_result[1], off), e2e_result[2]))
        sentences = S
        if world > 1:
            # rank 0 holds the gathered results: the weak one over all ranks' batches, the strong one over batch 0
            if rank == 0:
                texts, offs = [text], [off]
                for rk in range(1, world):
                    _, tx, of = load_dict_and_corpus(kind, n_sent, rk)
                    texts.append(tx)
                    offs.append(of)
                g_text = np.concatenate(texts)
                g_offs = [offs[0]]
                for of in offs[1:]:
                    g_offs.append(of[1:] + g_offs[-1][-1])
                g_offv = np.concatenate(g_offs)
                gref = oracle_reference(g_text, g_offv, cores)
                gw = gather.to_host(g_last)
                checks["gathered_weak"] = guarded(lambda: compare_with_oracle(
                    gref, gw[0], expand_tokens8(gw[0], gw[1], g_offv), gw[2]))
                gs = sg.to_host(g_strong)
                checks["gathered_strong"] = guarded(lambda: compare_with_oracle(
                    ref, gs[0], expand_tokens8(gs[0], gs[1], off0), gs[2]))
                sentences = len(g_offv) - 1
        bad = {k: v for k, v in checks.items() if v}
        flag = torch.tensor([1.0 if bad else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        parity = {"checked": True, "sentences": int(sentences), "match": flag.item() == 0.0,
                  "against": "oracle/ref_tokenize.cpp on the same bytes: tok_off, (id, class, position, start, end) of "
                             "every token, dp[EOS] of every sentence, bit-exact",
                  "results_checked": sorted(checks)}
        if bad:
            parity["mismatch"] = bad

    if rank == 0:
        K = args.steps
        peak, peak_src = measured_peak()
        N_nodes, E_pairs = ctr["nodes"], ctr["pairs"]
        a_total = (ctr["bytes"] + ctr["chars"] + 8 * (ctr["probes"] + ctr["probes_ok"]) + 40 * N_nodes + 8 * E_pairs
                   + 20 * ctr["tokens"])
        kern_ms = sum(stages.values()) / K
        fused = stages.get("fused_ms", 0.0) > 0.5 * sum(stages.values())
        if fused:      # the fused kernel does the whole path: its algorithmic bytes are the path's
            dom, a_dom, dom_ms = "kp_fused", a_total, stages["fused_ms"] / K
        else:          # multi-kernel pipeline: the sweep dominates
            dom, a_dom, dom_ms = "kp_viterbi", 16 * N_nodes + 8 * E_pairs, stages["viterbi_ms"] / K
        achieved = a_dom / (dom_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:   # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                tj = next(v for k, v in json.load(f).items() if k.startswith(dom))
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": world_bytes * K / (dev_total_ms * 1e-3), "unit": "bytes/s",
            "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": dev_total_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[args.workload], "sentences_per_gpu": S, "bytes_per_gpu": n_bytes,
                       "parallelism": ("dp%d: independent sentence shards; at N > 1 every timed step ends with the NCCL gather "
                                       "of the token records on rank 0" % world),
                       "path": args.path,
                       "l2": "flushed between steps (256 MiB memset outside the timed events)",
                       "timing": "CUDA events: the library stream around each whole step plus, at N > 1, the gather's stream "
                                 "around kp_gather_tokens; summed over steps, max over ranks"},
            "wall_ms_per_step": wall_ms / K,
            "e2e": {"value": world_bytes * K / (e2e_total_ms * 1e-3), "unit": "bytes/s",
                    "h2d_bytes_per_step": int(world_h2d), "d2h_bytes_per_step": int(world_d2h),
                    "ms_per_step": e2e_total_ms / K,
                    "queue_workers_wait": "blocking" if blocking else "spinning",
                    "expand_tokens8_ms_per_step": expand_ms,
                    "api": "kp_queue_submit / kp_queue_wait (C ABI), depth %d: pinned host text + offsets in, pinned host "
                           "kp_token8 records + offsets + dp[EOS] out, every step's copies inside the wall clock and "
                           "overlapped with the neighbouring steps' kernels; the records are the packed form the Rust "
                           "shim expands into Vec<Token> (that expansion is host work and not in this number: expand_tokens8_ms_per_step is kp_expand_tokens8 over one step's result on up to 8 host threads, best of 3)"
                           % args.queue_depth,
                    "sync_call": {"value": world_bytes * K / (sync_total_ms * 1e-3), "ms_per_step": sync_total_ms / K,
                                  "api": "kp_tokenize_batch8, one blocking call per step, nothing overlapped"}},
            "gpu_launches": int(world_launches),
            "parity": parity,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": a_dom, "launch_ms": dom_ms,
                         "share_of_kernel_time": dom_ms / kern_ms if kern_ms else None,
                         "whole_path": {"algorithmic_bytes": a_total, "bytes_per_input_byte": a_total / ctr["bytes"],
                                        "kernel_ms": kern_ms, "achieved": a_total / (kern_ms * 1e-3) / 1e9,
                                        "frac": a_total / (kern_ms * 1e-3) / 1e9 / peak}},
            "stages_ms_per_step": {k: v / K for k, v in stages.items()},
            "fused_sentences_per_step": fused_sentences,
            "counters": ctr,
            "clocks": clocks,
        }
        if world > 1:
            line["config"]["host_affinity"] = ("each rank bound to its GPU's " + numa) if numa else "not bound"
            line["dict_broadcast_ms"] = bcast_ms
            line["token_gather_ms"] = gather_total_ms / K
            line["collectives"] = ("NCCL: one broadcast of the packed dictionary (torch.distributed), one gather of token records per "
                                   "step (kp_gather_tokens: the library's grouped ncclSend / ncclRecv of one block per rank)")
            line["strong_scaling"] = strong
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(text, off)
            from oracle import oracle
            otk = oracle.OracleTokenizer(oracle.load_ipadic())
            first80 = bytes(text[int(off[0]):int(off[1])]).decode("utf-8")
            line["latency_us"] = latency_probe(tk, otk, {"cfg1": "すもももももももものうち", "cfg2_sentence0": first80})
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity.get("checked") and not parity.get("match"):
        return 3
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
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the benched batch (debug)")
    ap.add_argument("--path", default="auto", choices=["auto", "pipeline", "fused"])
    ap.add_argument("--queue-depth", type=int, default=3)
    ap.add_argument("--blocking-sync", default="off", choices=["on", "off"],
                    help="queue workers sleep on blocking-sync events instead of spinning while the device works")
    ap.add_argument("--no-numa", action="store_true", help="N > 1: do not bind each rank to its GPU's NUMA node")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return reference_arm(args)
    return product_arm(args)


if __name__ == "__main__":
    sys.exit(main())
