// kp_api.cu — tokenizer handle, chunk pipeline and the C ABI entry points of include/kanpyo_b200.h.
//
// Host-side mirror of the reference's `Tokenizer` (src/tokenizer.rs:7-45): `kp_tokenizer_create` is
// Tokenizer::new, `kp_tokenize` is Tokenizer::tokenize; the batch calls run the same function over
// many independent sentences in one device pass.  There is no CPU implementation behind any of them.
#include <limits.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "kp_kernels.cuh"

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return KP_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            kp_set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
            p = nullptr;
            return KP_ERR_NOMEM;
        }
        cap = want;
        return KP_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return (T*)p; }
};

struct PinBuf {   // pinned host memory, preserved on growth
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes, size_t keep) {
        if (bytes <= cap) return KP_OK;
        size_t want = std::max(bytes, cap * 2) + 4096;
        void* q = nullptr;
        cudaError_t e = cudaHostAlloc(&q, want, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            cudaGetLastError();
            kp_set_error("cudaHostAlloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return KP_ERR_NOMEM;
        }
        if (p && keep) memcpy(q, p, keep);
        if (p) cudaFreeHost(p);
        p = q;
        cap = want;
        return KP_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const { return (T*)p; }
};

constexpr uint32_t KP_PERM_REFRESH = 16;
#ifndef KP_FUSED_AUTO_SENTENCES
#define KP_FUSED_AUTO_SENTENCES 3584   // KP_PATH_AUTO: batches up to this many sentences take the fused kernel (the paths cross
                                       // between 3072 and 4096 sentences, profiles/r02_device_sweep.txt)
#endif
#ifndef KP_FUSED_BYTES
#define KP_FUSED_BYTES 192, 256, 320, 448, 768, 1536     // byte limits of the fused kernel's size classes
#endif
enum { EV_START, EV_H2D, EV_PIPE0, EV_PREP, EV_LATTICE, EV_BUCKET, EV_VITERBI, EV_BACKTRACE, EV_PACK, EV_FUSED0, EV_FUSED1, EV_END, EV_COUNT };

}  // namespace

struct kp_tokenizer {
    const kp_dict* dict = nullptr;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[EV_COUNT] = {};
    uint64_t chunk_bytes = 64ull << 20;
    bool count_work = false;
    int path_mode = KP_PATH_AUTO;
    bool fused_ok = false;             // the dictionary fits the fused kernel's packed records
    kp_fused_classes fclasses = {};
    cudaStream_t fstream[KP_FUSED_MAX_CLASSES] = {};
    cudaEvent_t fev_fork = nullptr, fev_join[KP_FUSED_MAX_CLASSES] = {};
    DevBuf flists, fctl;
    DevBuf small_in, small_out;        // small-batch path: one staging block in, one block out
    PinBuf h_small_in, h_small_out;
    uint64_t small_calls = 0;
    bool blocking_sync = false;        // wait on a blocking-sync event instead of spinning (many contexts per core)
    cudaEvent_t ev_block = nullptr;
    uint32_t fused_max_batch = KP_FUSED_AUTO_SENTENCES;
    kp_perm perm = {};
    uint32_t perm_age = 0;     // passes since the column order was last ranked
    DevBuf perm_hist, perm_map, perm_conn;
    // chunk scratch
    DevBuf text, off, nchar, coff, binfo, ncount, noff, bcount, boff, ucount, rbk, nhit, hits, bfill, rec, tgt, red, ndp, bnode, path, pre, lenhist, order,
        tcount, toff32, scan_tmp, totals, err, stage, sel;
    // device outputs
    DevBuf d_tok_off, d_tokens, d_eos;
    // host staging
    PinBuf h_totals, h_tok_off, h_tokens, h_eos, h_misc;
    std::vector<kp_lattice_node> lattice_nodes;
    kp_counters counters = {};
    kp_profile profile = {};
    bool pipeline_timed = false;
    uint64_t last_tokens = 0;
    kp_chunk pass = {};        // chunk of a two-phase pass (kp_pass_begin / kp_pass_pack)
    const void* res_tok_off = nullptr;   // host result of the last kp_tokenize_batch[8] call
    const void* res_tokens = nullptr;
    const int32_t* res_eos = nullptr;
};

namespace {

#define KP_TRY(x)            \
    do {                     \
        int rc__ = (x);      \
        if (rc__ < 0) return rc__; \
    } while (0)
#define KP_LAUNCH(x)                         \
    do {                                     \
        int rc__ = (x);                      \
        if (rc__ < 0) return rc__;           \
        t->profile.kernel_launches += rc__;  \
    } while (0)

// Wait for everything queued on the tokenizer's stream.  Default: cudaStreamSynchronize (the driver spins: lowest
// latency).  With blocking_sync the thread sleeps on a cudaEventBlockingSync event instead: a queue of three contexts
// per GPU on eight GPUs is 24 waiting threads, and spinning ones starve each other on a 32-core host.
static int kp_wait_stream(kp_tokenizer* t) {
    if (!t->blocking_sync) {
        KP_CUDA(cudaStreamSynchronize(t->stream));
        return KP_OK;
    }
    if (!t->ev_block) KP_CUDA(cudaEventCreateWithFlags(&t->ev_block, cudaEventBlockingSync | cudaEventDisableTiming));
    KP_CUDA(cudaEventRecord(t->ev_block, t->stream));
    KP_CUDA(cudaEventSynchronize(t->ev_block));
    return KP_OK;
}

struct StageTimes {
    float prep = 0, lattice = 0, bucket = 0, viterbi = 0, backtrace = 0;
};

// The multi-kernel pipeline over the sentences of a chunk (all of them, or the `c.sel` subset the fused
// kernel left): lattice, buckets, sweep, back-trace; tokens end up staged (c.stage), token counts in
// c.tcount and dp[EOS] in c.eos_cost, all indexed by chunk sentence.
int run_pipeline(kp_tokenizer* t, kp_chunk& c, StageTimes* times) {
    cudaStream_t st = t->stream;
    const kp_ddict& d = t->dict->view;
    const uint32_t S = c.S;
    KP_TRY(t->nchar.ensure(sizeof(uint32_t) * (S + 1)));
    KP_TRY(t->coff.ensure(sizeof(uint32_t) * (S + 2)));
    KP_TRY(t->scan_tmp.ensure(sizeof(uint64_t) * kp_scan_tmp_elems(std::max(S, c.S_all) + 1)));
    KP_TRY(t->lenhist.ensure(sizeof(uint32_t) * kp_len_bins()));
    KP_TRY(t->order.ensure(sizeof(uint32_t) * (S + 1)));
    c.lenhist = t->lenhist.as<uint32_t>();
    c.order = t->order.as<uint32_t>();
    c.nchar = t->nchar.as<uint32_t>();
    c.coff = t->coff.as<uint32_t>();
    c.scan_tmp = t->scan_tmp.as<uint64_t>();
    uint64_t* h_tot = t->h_totals.as<uint64_t>();
    uint32_t* h_err = (uint32_t*)(h_tot + 16);

    KP_CUDA(cudaEventRecord(t->ev[EV_PIPE0], st));
    KP_CUDA(cudaMemsetAsync(c.lenhist, 0, sizeof(uint32_t) * kp_len_bins(), st));
    KP_LAUNCH(kp_launch_prep_count(c, st));
    KP_LAUNCH(kp_launch_scan(c.nchar, c.coff, S, c.scan_tmp, &c.totals[0], st));
    KP_LAUNCH(kp_launch_length_order(c, st));       // scan of the length histogram prep_count filled
    KP_CUDA(cudaMemcpyAsync(h_tot, c.totals, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    KP_CUDA(cudaMemcpyAsync(h_err, c.err, sizeof(uint32_t) * 2, cudaMemcpyDeviceToHost, st));
    KP_TRY(kp_wait_stream(t));
    if (h_err[1]) {
        kp_set_error("sentence offsets are not ascending or exceed the text length");
        return KP_ERR_ARG;
    }
    if (h_err[0]) {
        kp_set_error("input contains invalid UTF-8");
        return KP_ERR_UTF8;
    }
    if (h_tot[0] + S + 1 >= (1ull << 32)) return KP_ERR_TOO_LARGE;
    c.C = (uint32_t)h_tot[0];
    c.NB = c.C + S;
    const size_t NB = c.NB;
    KP_TRY(t->binfo.ensure(sizeof(uint4) * (NB + 1)));
    KP_TRY(t->ncount.ensure(sizeof(uint32_t) * (NB + 1)));
    KP_TRY(t->noff.ensure(sizeof(uint32_t) * (NB + 2)));
    KP_TRY(t->bcount.ensure(sizeof(uint32_t) * (NB + 1)));
    KP_TRY(t->boff.ensure(sizeof(uint32_t) * (NB + 2)));
    KP_TRY(t->bfill.ensure(sizeof(uint2) * (NB + 1)));
    KP_TRY(t->ucount.ensure(sizeof(uint32_t) * (NB + 1)));
    KP_TRY(t->nhit.ensure(NB + 16));
    KP_TRY(t->hits.ensure(sizeof(uint4) * 2 * (NB + 1)));
    KP_TRY(t->rbk.ensure(sizeof(uint2) * (NB + 1)));
    KP_TRY(t->scan_tmp.ensure(sizeof(uint64_t) * kp_scan_tmp_elems((uint32_t)NB + 1)));
    c.scan_tmp = t->scan_tmp.as<uint64_t>();
    c.binfo = t->binfo.as<uint4>();
    c.ncount = t->ncount.as<uint32_t>();
    c.noff = t->noff.as<uint32_t>();
    c.bcount = t->bcount.as<uint32_t>();
    c.boff = t->boff.as<uint32_t>();
    c.bfill = t->bfill.as<uint2>();
    c.ucount = t->ucount.as<uint32_t>();
    c.nhit = t->nhit.as<uint8_t>();
    c.hits = t->hits.as<uint4>();
    c.rbk = t->rbk.as<uint2>();

    KP_LAUNCH(kp_launch_prep_fill(c, d, st));        // boundary table + scatter of the Viterbi work order
    KP_CUDA(cudaEventRecord(t->ev[EV_PREP], st));
    KP_LAUNCH(kp_launch_lattice_count(c, d, t->count_work, st));
    KP_LAUNCH(kp_launch_scan2(c.ncount, c.bcount, c.noff, c.boff, c.NB, c.scan_tmp, &c.totals[1], &c.totals[2], st));
    KP_CUDA(cudaMemcpyAsync(h_tot, c.totals, sizeof(uint64_t) * 8, cudaMemcpyDeviceToHost, st));
    KP_TRY(kp_wait_stream(t));
    if (h_tot[1] >= (1ull << 31) - 1) return KP_ERR_TOO_LARGE;   // node indices carry a flag in bit 31 (KP_SLOT_SHARED)
    c.N = (uint32_t)h_tot[1];
    if (h_tot[2] != h_tot[1]) {
        kp_set_error("internal: bucket entries %llu != nodes %llu", (unsigned long long)h_tot[2],
                     (unsigned long long)h_tot[1]);
        return KP_ERR_CUDA;
    }
    const size_t N = c.N;
    KP_TRY(t->rec.ensure(sizeof(uint4) * (N + 1)));
    KP_TRY(t->tgt.ensure(sizeof(uint2) * (N + 1)));
    KP_TRY(t->red.ensure(sizeof(int2) * (N + 1)));
    KP_TRY(t->ndp.ensure(sizeof(int32_t) * (N + 1)));
    KP_TRY(t->bnode.ensure(sizeof(uint32_t) * (N + 1)));
    KP_TRY(t->path.ensure(sizeof(uint32_t) * (NB + 1)));
    c.rec = t->rec.as<uint4>();
    c.tgt = t->tgt.as<uint2>();
    c.red = t->red.as<int2>();
    c.ndp = t->ndp.as<int32_t>();
    c.bnode = t->bnode.as<uint32_t>();
    c.path = t->path.as<uint32_t>();

    KP_LAUNCH(kp_launch_lattice_fill(c, d, st));
    KP_CUDA(cudaEventRecord(t->ev[EV_LATTICE], st));
    KP_CUDA(cudaMemsetAsync(c.bfill, 0, sizeof(uint2) * (NB + 1), st));
    if (t->perm_age % KP_PERM_REFRESH == 0) KP_LAUNCH(kp_launch_column_order(c, d, t->perm, st));
    t->perm_age++;
    KP_LAUNCH(kp_launch_bucketize(c, d, t->perm, st));
    KP_CUDA(cudaEventRecord(t->ev[EV_BUCKET], st));
    KP_LAUNCH(kp_launch_viterbi(c, d, t->perm, st));
    if (t->count_work) KP_LAUNCH(kp_launch_pair_count(c, st));
    KP_CUDA(cudaEventRecord(t->ev[EV_VITERBI], st));
    KP_LAUNCH(kp_launch_backtrace_count(c, d, st));
    c.direct_emit = c.sel == nullptr;               // the whole chunk is here: no staging needed
    if (!c.direct_emit) KP_LAUNCH(kp_launch_backtrace_stage(c, st));
    KP_CUDA(cudaEventRecord(t->ev[EV_BACKTRACE], st));
    t->counters.chars += c.C;
    t->counters.nodes += (uint64_t)c.N + S;   // + one BOS per sentence (not materialised on the device)
    t->pipeline_timed = true;
    (void)times;
    return KP_OK;
}

// The fused per-sentence kernel over the chunk: classify by sentence bytes, one persistent launch per size
// class (largest first, each on its own stream so that the tail of one class overlaps the start of the
// next), then the flags and the list of sentences left to the pipeline come back in one round trip.
int run_fused(kp_tokenizer* t, kp_chunk& c, uint32_t* left) {
    cudaStream_t st = t->stream;
    const kp_ddict& d = t->dict->view;
    const uint32_t S = c.S_all, nc = t->fclasses.n;
    // fctl: [0..nc) list lengths, [8..8+nc) tickets, [16] sentences left to the pipeline
    KP_TRY(t->sel.ensure(sizeof(uint32_t) * (S + 1)));
    KP_TRY(t->flists.ensure(sizeof(uint32_t) * ((size_t)nc * S + 1)));
    KP_TRY(t->fctl.ensure(sizeof(uint32_t) * 32));
    c.sel_out = t->sel.as<uint32_t>();
    uint32_t* fctl = t->fctl.as<uint32_t>();
    uint64_t* h_tot = t->h_totals.as<uint64_t>();
    uint32_t* h_err = (uint32_t*)(h_tot + 16);
    uint32_t* h_ctl = h_err + 4;
    KP_CUDA(cudaMemsetAsync(fctl, 0, sizeof(uint32_t) * 32, st));
    KP_CUDA(cudaEventRecord(t->ev[EV_FUSED0], st));
    KP_LAUNCH(kp_launch_fused_classify(c, t->fclasses, t->flists.as<uint32_t>(), fctl, fctl + 16, st));
    KP_CUDA(cudaEventRecord(t->fev_fork, st));
    // classes 0 .. nc-2 side by side (larger first: the tail of one overlaps the start of the next); a sentence that
    // overflows its class is appended to the LAST class's list, whose kernel runs after the others; only what
    // overflows that one is left to the pipeline
    uint32_t* const esc_list = t->flists.as<uint32_t>() + (size_t)(nc - 1) * S;
    for (uint32_t i = 1; i < nc; i++) {
        const uint32_t k = nc - 1 - i;
        KP_CUDA(cudaStreamWaitEvent(t->fstream[k], t->fev_fork, 0));
        KP_LAUNCH(kp_launch_fused(c, d, t->fclasses.c[k], t->flists.as<uint32_t>() + (size_t)k * S, fctl + k, fctl + 8 + k,
                                  esc_list, fctl + (nc - 1), S, t->fstream[k]));
        KP_CUDA(cudaEventRecord(t->fev_join[k], t->fstream[k]));
        KP_CUDA(cudaStreamWaitEvent(st, t->fev_join[k], 0));
    }
    KP_LAUNCH(kp_launch_fused(c, d, t->fclasses.c[nc - 1], esc_list, fctl + (nc - 1), fctl + 8 + (nc - 1), c.sel_out, fctl + 16,
                              S, st));
    KP_CUDA(cudaEventRecord(t->ev[EV_FUSED1], st));
    KP_CUDA(cudaMemcpyAsync(h_err, c.err, sizeof(uint32_t) * 4, cudaMemcpyDeviceToHost, st));
    KP_CUDA(cudaMemcpyAsync(h_ctl, fctl, sizeof(uint32_t) * 32, cudaMemcpyDeviceToHost, st));
    KP_CUDA(cudaMemcpyAsync(h_tot + 8, c.totals + 8, sizeof(uint64_t) * 4, cudaMemcpyDeviceToHost, st));
    KP_TRY(kp_wait_stream(t));
    if (h_err[1]) {
        kp_set_error("sentence offsets are not ascending or exceed the text length");
        return KP_ERR_ARG;
    }
    if (h_err[0]) {
        kp_set_error("input contains invalid UTF-8");
        return KP_ERR_UTF8;
    }
    *left = h_ctl[16];
    t->counters.chars += h_tot[8];
    t->counters.nodes += h_tot[9];
    t->profile.fused_sentences += (uint32_t)h_tot[10];
    float ms = 0;
    cudaEventElapsedTime(&ms, t->ev[EV_FUSED0], t->ev[EV_FUSED1]);
    t->profile.fused_ms += ms;
    return KP_OK;
}

// One device pass over a chunk whose text / offsets are already in device memory, in two halves so
// that a multi-GPU caller can learn every shard's token count before any shard packs its result:
//   chunk_compute  lattice, Viterbi, back-trace; tokens staged; token counts scanned; *n_tokens read back
//                  (single-GPU calls hand it the token base and it queues chunk_pack before waiting)
//   chunk_pack     staged tokens -> c.tok_off / c.tokens (offsets rebased by tok_base); c.eos_cost is final
int chunk_pack(kp_tokenizer* t, kp_chunk& c, uint64_t tok_base, bool compact);

// pack_base: the chunk's token base when the caller already knows it (single-GPU calls), nullptr otherwise.  With a
// known base the pack kernel is queued BEHIND the scan before the host waits for the token count, so the round trip
// hides behind it.
int chunk_compute(kp_tokenizer* t, kp_chunk& c, uint64_t* n_tokens, StageTimes* times, const uint64_t* pack_base = nullptr,
                  bool compact = false) {
    cudaStream_t st = t->stream;
    const uint32_t S = c.S;
    c.S_all = S;
    c.sel = nullptr;
    KP_TRY(t->totals.ensure(sizeof(uint64_t) * 16));
    KP_TRY(t->err.ensure(sizeof(uint32_t) * 4));
    KP_TRY(t->h_totals.ensure(sizeof(uint64_t) * 40, 0));
    KP_TRY(t->tcount.ensure(sizeof(uint32_t) * (S + 1)));
    KP_TRY(t->toff32.ensure(sizeof(uint32_t) * (S + 2)));
    KP_TRY(t->stage.ensure(sizeof(kp_token) * ((size_t)c.B + S + 2)));
    KP_TRY(t->scan_tmp.ensure(sizeof(uint64_t) * kp_scan_tmp_elems(S + 1)));
    c.totals = t->totals.as<uint64_t>();
    c.err = t->err.as<uint32_t>();
    c.tcount = t->tcount.as<uint32_t>();
    c.toff32 = t->toff32.as<uint32_t>();
    c.stage = t->stage.as<kp_token>();
    c.scan_tmp = t->scan_tmp.as<uint64_t>();
    uint64_t* h_tot = t->h_totals.as<uint64_t>();
    KP_CUDA(cudaEventRecord(t->ev[EV_H2D], st));          // the inputs are on the device from here on
    KP_CUDA(cudaMemsetAsync(c.totals, 0, sizeof(uint64_t) * 16, st));
    KP_CUDA(cudaMemsetAsync(c.err, 0, sizeof(uint32_t) * 4, st));
    t->pipeline_timed = false;
    // which device path: the fused per-sentence kernel keeps a sentence on chip but holds few sentences per SM, so
    // it wins where the batch is too small to fill the pipeline's kernels (and for single sentences, where the
    // pipeline's dozen launches and three round trips dominate); KP_PATH_FUSED / _PIPELINE force either one
    // (a batch whose MEAN sentence is longer than the largest size class -- long-line input -- is the pipeline's anyway)
    const uint64_t largest = t->fclasses.n ? t->fclasses.c[t->fclasses.n - 1].max_bytes : 0;
    const bool fused = t->fused_ok && !t->count_work && S > 0 &&
                       (t->path_mode == KP_PATH_FUSED ||
                        (t->path_mode == KP_PATH_AUTO && S <= t->fused_max_batch && (uint64_t)c.B <= largest * S));
    if (fused) {
        // fused per-sentence kernel first; the pipeline then takes the sentences it left (c.sel)
        uint32_t left = 0;
        KP_TRY(run_fused(t, c, &left));
        if (left) {
            c.sel = t->sel.as<uint32_t>();
            c.S = left;
            KP_TRY(run_pipeline(t, c, times));
        }
    } else {
        KP_TRY(run_pipeline(t, c, times));
    }
    c.scan_tmp = t->scan_tmp.as<uint64_t>();
    KP_LAUNCH(kp_launch_scan(c.tcount, c.toff32, c.S_all, c.scan_tmp, &c.totals[3], st));
    KP_CUDA(cudaMemcpyAsync(h_tot, c.totals, sizeof(uint64_t) * 8, cudaMemcpyDeviceToHost, st));
    if (pack_base) {
        KP_TRY(chunk_pack(t, c, *pack_base, compact));
        KP_CUDA(cudaEventRecord(t->ev[EV_END], st));      // end of the device work (a host caller records it again after its D2H)
    }
    KP_TRY(kp_wait_stream(t));
    *n_tokens = h_tot[3];
    t->counters.bytes += c.B;
    t->counters.tokens += h_tot[3];
    t->counters.sentences += c.S_all;
    t->counters.probes += h_tot[4];
    t->counters.probes_ok += h_tot[5];
    t->counters.pairs += h_tot[6];
    if (times && t->pipeline_timed) {
        float ms = 0;
        cudaEventElapsedTime(&ms, t->ev[EV_PIPE0], t->ev[EV_PREP]);     times->prep += ms;
        cudaEventElapsedTime(&ms, t->ev[EV_PREP], t->ev[EV_LATTICE]);   times->lattice += ms;
        cudaEventElapsedTime(&ms, t->ev[EV_LATTICE], t->ev[EV_BUCKET]); times->bucket += ms;
        cudaEventElapsedTime(&ms, t->ev[EV_BUCKET], t->ev[EV_VITERBI]); times->viterbi += ms;
        cudaEventElapsedTime(&ms, t->ev[EV_VITERBI], t->ev[EV_BACKTRACE]); times->backtrace += ms;
    }
    t->profile.chunks++;
    return KP_OK;
}

int chunk_pack(kp_tokenizer* t, kp_chunk& c, uint64_t tok_base, bool compact) {
    cudaStream_t st = t->stream;
    if (c.direct_emit) KP_LAUNCH(kp_launch_tokens_emit(c, tok_base, compact, st));
    else KP_LAUNCH(kp_launch_tokens_pack(c, tok_base, compact, st));
    KP_CUDA(cudaEventRecord(t->ev[EV_PACK], st));
    return KP_OK;
}

int run_chunk(kp_tokenizer* t, kp_chunk& c, uint64_t tok_base, bool compact, uint64_t* n_tokens, StageTimes* times) {
    return chunk_compute(t, c, n_tokens, times, &tok_base, compact);
}

void begin_call(kp_tokenizer* t) {
    memset(&t->counters, 0, sizeof(t->counters));
    memset(&t->profile, 0, sizeof(t->profile));
}

void store_times(kp_tokenizer* t, const StageTimes& s) {
    t->profile.prep_ms = s.prep;
    t->profile.lattice_ms = s.lattice;
    t->profile.bucket_ms = s.bucket;
    t->profile.viterbi_ms = s.viterbi;
    t->profile.backtrace_ms = s.backtrace;
}

}  // namespace

extern "C" int kp_tokenizer_create(const kp_dict* d, kp_tokenizer** out) {
    if (!d || !out) return KP_ERR_ARG;
    *out = nullptr;
    KP_CUDA(cudaSetDevice(d->device));
    kp_tokenizer* t = new kp_tokenizer();
    t->dict = d;
    t->device = d->device;
    cudaError_t e = cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking);
    for (int i = 0; i < EV_COUNT && e == cudaSuccess; i++) e = cudaEventCreate(&t->ev[i]);
    if (e != cudaSuccess) {
        kp_set_error("stream/event creation failed: %s", cudaGetErrorString(e));
        delete t;
        return KP_ERR_CUDA;
    }
    {   // column order of the sweep's connection matrix: a private, permuted copy of the transposed matrix
        const kp_ddict& v = d->view;
        int rc = KP_OK;
        if ((rc = t->perm_hist.ensure(sizeof(uint32_t) * (v.conn_col + 1))) ||
            (rc = t->perm_map.ensure(sizeof(uint16_t) * (v.conn_col + 1))) ||
            (rc = t->perm_conn.ensure(sizeof(int16_t) * ((size_t)v.conn_row * v.connT_stride + 8)))) {
            kp_tokenizer_destroy(t);
            return rc;
        }
        t->perm.hist = t->perm_hist.as<uint32_t>();
        t->perm.perm = t->perm_map.as<uint16_t>();
        t->perm.connP = t->perm_conn.as<int16_t>();
        cudaMemset(t->perm.connP, 0, sizeof(int16_t) * (size_t)v.conn_row * v.connT_stride);   // row padding
    }
    {   // fused per-sentence path: size classes by sentence bytes (measured on the synthetic corpora: 0.34-0.43
        // chars, <= 1.6 known nodes and <= 1.8 reduced slots per byte at the 99th percentile; what does not fit
        // its class goes to the pipeline)
        const kp_blob_header* h = (const kp_blob_header*)d->host_blob.data();
        t->fused_ok = kp_fused_dict_ok(d->view, (const kp_catinfo*)(d->host_blob.data() + h->off_catinfo));
        static const uint32_t limits[] = {KP_FUSED_BYTES};
        t->fclasses.n = 0;
        for (uint32_t lim : limits) {
            if (t->fclasses.n >= KP_FUSED_MAX_CLASSES) break;
            kp_fused_class& k = t->fclasses.c[t->fclasses.n++];
            k.max_bytes = lim;
            k.cap_c = std::min<uint32_t>(lim * 9 / 20 + 8, 1020);
            k.cap_k = std::min<uint32_t>((lim * 17 / 10 + 32) & ~1u, 4000);
            k.cap_r = std::min<uint32_t>(lim * 19 / 10 + 48, 4090);
            if (k.cap_k * 3 < 8 * (k.cap_c + 1)) k.cap_k = ((8 * (k.cap_c + 1) + 2) / 3 + 1) & ~1u;   // the path overlays t_slot | hy
            if (k.cap_r < k.cap_k / 2) k.cap_r = k.cap_k / 2;                                           // the hits overlay k_dp
        }
        if (t->fused_ok) {
            int rc = kp_fused_prepare(&t->fclasses, t->device);
            cudaError_t e2 = cudaEventCreateWithFlags(&t->fev_fork, cudaEventDisableTiming);
            for (uint32_t k = 0; k < t->fclasses.n && e2 == cudaSuccess; k++) {
                e2 = cudaStreamCreateWithFlags(&t->fstream[k], cudaStreamNonBlocking);
                if (e2 == cudaSuccess) e2 = cudaEventCreateWithFlags(&t->fev_join[k], cudaEventDisableTiming);
            }
            if (rc || e2 != cudaSuccess) {
                if (!rc) kp_set_error("fused path: stream/event creation failed: %s", cudaGetErrorString(e2));
                kp_tokenizer_destroy(t);
                return rc ? rc : KP_ERR_CUDA;
            }
        }
    }
    *out = t;
    return KP_OK;
}

extern "C" void kp_tokenizer_destroy(kp_tokenizer* t) {
    if (!t) return;
    cudaSetDevice(t->device);
    if (t->stream) cudaStreamSynchronize(t->stream);
    DevBuf* bufs[] = {&t->text, &t->off, &t->nchar, &t->coff, &t->binfo, &t->ncount, &t->noff, &t->bcount, &t->boff,
                      &t->bfill, &t->ucount, &t->rbk, &t->nhit, &t->hits, &t->rec, &t->tgt, &t->red, &t->ndp, &t->bnode, &t->path, &t->pre, &t->lenhist, &t->order, &t->tcount, &t->toff32,
                      &t->scan_tmp, &t->totals, &t->err, &t->stage, &t->sel, &t->flists, &t->fctl, &t->small_in, &t->small_out, &t->d_tok_off, &t->d_tokens, &t->d_eos, &t->perm_hist, &t->perm_map,
                      &t->perm_conn};
    for (DevBuf* b : bufs) b->release();
    PinBuf* pins[] = {&t->h_totals, &t->h_tok_off, &t->h_tokens, &t->h_eos, &t->h_misc, &t->h_small_in, &t->h_small_out};
    for (PinBuf* b : pins) b->release();
    for (int i = 0; i < EV_COUNT; i++)
        if (t->ev[i]) cudaEventDestroy(t->ev[i]);
    if (t->ev_block) cudaEventDestroy(t->ev_block);
    if (t->fev_fork) cudaEventDestroy(t->fev_fork);
    for (uint32_t k = 0; k < KP_FUSED_MAX_CLASSES; k++) {
        if (t->fev_join[k]) cudaEventDestroy(t->fev_join[k]);
        if (t->fstream[k]) cudaStreamDestroy(t->fstream[k]);
    }
    if (t->stream) cudaStreamDestroy(t->stream);
    delete t;
}

extern "C" int kp_tokenizer_set_chunk_bytes(kp_tokenizer* t, uint64_t bytes) {
    if (!t || bytes == 0) return KP_ERR_ARG;
    t->chunk_bytes = std::min<uint64_t>(bytes, (1ull << 31) - 1);
    return KP_OK;
}

extern "C" int kp_tokenizer_set_blocking_sync(kp_tokenizer* t, int on) {
    if (!t) return KP_ERR_ARG;
    t->blocking_sync = on != 0;
    return KP_OK;
}

extern "C" int kp_tokenizer_set_count_work(kp_tokenizer* t, int on) {
    if (!t) return KP_ERR_ARG;
    t->count_work = on != 0;
    return KP_OK;
}

extern "C" int kp_tokenizer_sync(kp_tokenizer* t) {
    if (!t) return KP_ERR_ARG;
    KP_CUDA(cudaSetDevice(t->device));
    KP_CUDA(cudaStreamSynchronize(t->stream));
    return KP_OK;
}

extern "C" int kp_copy_to_host(kp_tokenizer* t, void* dst, const void* device_src, uint64_t bytes) {
    if (!t || (bytes && (!dst || !device_src))) return KP_ERR_ARG;
    KP_CUDA(cudaSetDevice(t->device));
    if (bytes) {
        KP_CUDA(cudaMemcpyAsync(dst, device_src, bytes, cudaMemcpyDeviceToHost, t->stream));
        KP_CUDA(cudaStreamSynchronize(t->stream));
    }
    return KP_OK;
}

extern "C" int kp_last_counters(const kp_tokenizer* t, kp_counters* out) {
    if (!t || !out) return KP_ERR_ARG;
    *out = t->counters;
    return KP_OK;
}

extern "C" int kp_last_profile(const kp_tokenizer* t, kp_profile* out) {
    if (!t || !out) return KP_ERR_ARG;
    *out = t->profile;
    return KP_OK;
}

extern "C" int kp_tokenizer_set_path(kp_tokenizer* t, int path) {
    if (!t || path < KP_PATH_AUTO || path > KP_PATH_FUSED) return KP_ERR_ARG;
    t->path_mode = path;
    return KP_OK;
}

static int kp_tokenize_device(kp_tokenizer* t, const uint8_t* d_utf8, const uint64_t* d_offsets, uint64_t n_sent,
                              uint64_t first_offset, uint64_t n_bytes, bool compact, kp_chunk* used, uint64_t* n_tokens) {
    if (!t || !d_offsets || (n_bytes && !d_utf8)) return KP_ERR_ARG;
    if (n_bytes >= (1ull << 31) || n_sent >= (1ull << 31) - 2) return KP_ERR_TOO_LARGE;
    KP_CUDA(cudaSetDevice(t->device));
    begin_call(t);
    cudaStream_t st = t->stream;
    kp_chunk c;
    memset(&c, 0, sizeof(c));
    c.text = d_utf8 + first_offset;
    c.off = d_offsets;
    c.base = first_offset;
    c.S = (uint32_t)n_sent;
    c.B = (uint32_t)n_bytes;
    // tokens <= chars + sentences <= bytes + sentences: size the outputs without a device round trip
    KP_TRY(t->d_tok_off.ensure(sizeof(uint64_t) * (n_sent + 1)));
    KP_TRY(t->d_eos.ensure(sizeof(int32_t) * (n_sent + 1)));
    KP_TRY(t->d_tokens.ensure(sizeof(kp_token) * (n_bytes + n_sent + 1)));
    c.tok_off = t->d_tok_off.p;
    c.eos_cost = t->d_eos.as<int32_t>();
    c.tokens = t->d_tokens.p;
    KP_CUDA(cudaEventRecord(t->ev[EV_START], st));
    StageTimes times;
    KP_TRY(run_chunk(t, c, 0, compact, n_tokens, &times));     // returns with the stream drained, EV_END recorded behind the pack
    store_times(t, times);
    cudaEventElapsedTime(&t->profile.total_ms, t->ev[EV_START], t->ev[EV_END]);
    *used = c;
    return KP_OK;
}

extern "C" int kp_tokenize_batch_device(kp_tokenizer* t, const uint8_t* d_utf8, const uint64_t* d_offsets,
                                        uint64_t n_sent, uint64_t first_offset, uint64_t n_bytes, kp_result* out) {
    if (!out) return KP_ERR_ARG;
    kp_chunk c;
    uint64_t ntok = 0;
    KP_TRY(kp_tokenize_device(t, d_utf8, d_offsets, n_sent, first_offset, n_bytes, false, &c, &ntok));
    out->n_sent = n_sent;
    out->n_tokens = ntok;
    out->tok_off = (const uint64_t*)c.tok_off;
    out->tokens = (const kp_token*)c.tokens;
    out->eos_cost = c.eos_cost;
    return KP_OK;
}

extern "C" int kp_tokenize_batch_device8(kp_tokenizer* t, const uint8_t* d_utf8, const uint64_t* d_offsets,
                                         uint64_t n_sent, uint64_t first_offset, uint64_t n_bytes, kp_result8* out) {
    if (!out) return KP_ERR_ARG;
    kp_chunk c;
    uint64_t ntok = 0;
    KP_TRY(kp_tokenize_device(t, d_utf8, d_offsets, n_sent, first_offset, n_bytes, true, &c, &ntok));
    if (ntok >= (1ull << 32)) return KP_ERR_TOO_LARGE;
    out->n_sent = n_sent;
    out->n_tokens = ntok;
    out->tok_off = (const uint32_t*)c.tok_off;
    out->tokens = (const kp_token8*)c.tokens;
    out->eos_cost = c.eos_cost;
    return KP_OK;
}

// ---- small batches (the reference's own call pattern is one line per call, src/bin/kanpyo.rs:115-122) ----
// Up to KP_SMALL_SENT sentences / KP_SMALL_BYTES bytes go through the fused kernel with ONE host round trip:
// the host sorts the sentences into the size classes (it knows their byte lengths), one staging block goes
// in (offsets, class lists, counters), the class kernels + scan + pack run back to back, and one block comes
// out (counters, flags, token offsets, dp[EOS], and the token records up to their upper bound bytes +
// sentences).  Returns 1 when the batch was done here, 0 when the general path has to take it (something
// did not fit the fused kernel), negative on error.
constexpr uint64_t KP_SMALL_SENT = 64, KP_SMALL_BYTES = 48 << 10;

static inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

static int tokenize_small(kp_tokenizer* t, const uint8_t* utf8, const uint64_t* offsets, uint64_t n_sent, bool compact,
                   uint64_t* n_tokens, const void** h_tok_off, const void** h_tokens, const int32_t** h_eos) {
    const uint32_t S = (uint32_t)n_sent, nc = t->fclasses.n;
    const uint64_t nbytes = offsets[n_sent] - offsets[0];
    cudaStream_t st = t->stream;
    const kp_ddict& d = t->dict->view;
    const size_t tok_sz = compact ? sizeof(kp_token8) : sizeof(kp_token), off_sz = compact ? sizeof(uint32_t) : sizeof(uint64_t);
    const size_t bound = nbytes + S;                        // tokens <= chars + sentences <= bytes + sentences
    // staging in: [totals u64 x16 | err u32 x4 | fctl u32 x32 | offsets u64 x(S+1) | lists u32 x nc*S]
    const size_t in_tot = 0, in_err = 128, in_ctl = 144, in_off = 272, in_lists = in_off + 8 * (size_t)(S + 1);
    const size_t in_bytes = align16(in_lists + 4 * (size_t)nc * S);
    // block out: [totals | err | fctl | tok_off | eos | tokens]
    const size_t out_off = 272, out_eos = align16(out_off + off_sz * (S + 1)), out_tok = align16(out_eos + 4 * (size_t)S);
    const size_t out_bytes = out_tok + tok_sz * bound;
    KP_TRY(t->small_in.ensure(in_bytes));
    KP_TRY(t->small_out.ensure(out_bytes));
    KP_TRY(t->h_small_in.ensure(in_bytes, 0));
    KP_TRY(t->h_small_out.ensure(out_bytes, 0));
    KP_TRY(t->text.ensure(nbytes + 16));
    KP_TRY(t->tcount.ensure(sizeof(uint32_t) * (S + 1)));
    KP_TRY(t->toff32.ensure(sizeof(uint32_t) * (S + 2)));
    KP_TRY(t->stage.ensure(sizeof(kp_token) * (bound + 2)));
    KP_TRY(t->scan_tmp.ensure(sizeof(uint64_t) * kp_scan_tmp_elems(S + 1)));
    KP_TRY(t->sel.ensure(sizeof(uint32_t) * (S + 1)));
    char* hin = (char*)t->h_small_in.p;
    memset(hin, 0, in_off);
    memcpy(hin + in_off, offsets, 8 * (size_t)(S + 1));
    uint32_t* h_ctl = (uint32_t*)(hin + in_ctl);
    uint32_t* h_lists = (uint32_t*)(hin + in_lists);
    // one launch: every sentence runs in the class the longest one needs (a handful of blocks: occupancy is moot,
    // serialising several class kernels would not be)
    uint32_t kmax = 0;
    for (uint32_t s = 0; s < S; s++) {
        const uint64_t b = offsets[s + 1] - offsets[s];
        while (kmax < nc && b > t->fclasses.c[kmax].max_bytes) kmax++;
        if (kmax == nc) return 0;                           // longer than every class: the general path
    }
    for (uint32_t s = 0; s < S; s++) h_lists[(size_t)kmax * S + h_ctl[kmax]++] = s;
    char* din = (char*)t->small_in.p;
    char* dout = (char*)t->small_out.p;
    KP_CUDA(cudaEventRecord(t->ev[EV_START], st));
    if (nbytes) KP_CUDA(cudaMemcpyAsync(t->text.p, utf8 + offsets[0], nbytes, cudaMemcpyHostToDevice, st));
    KP_CUDA(cudaMemcpyAsync(din, hin, in_bytes, cudaMemcpyHostToDevice, st));
    kp_chunk c;
    memset(&c, 0, sizeof(c));
    c.text = t->text.as<uint8_t>();
    c.off = (const uint64_t*)(din + in_off);
    c.base = offsets[0];
    c.S = c.S_all = S;
    c.B = (uint32_t)nbytes;
    c.totals = (uint64_t*)(din + in_tot);
    c.err = (uint32_t*)(din + in_err);
    c.tcount = t->tcount.as<uint32_t>();
    c.toff32 = t->toff32.as<uint32_t>();
    c.stage = t->stage.as<kp_token>();
    c.scan_tmp = t->scan_tmp.as<uint64_t>();
    c.sel_out = t->sel.as<uint32_t>();
    c.tok_off = dout + out_off;
    c.eos_cost = (int32_t*)(dout + out_eos);
    c.tokens = dout + out_tok;
    uint32_t* fctl = (uint32_t*)(din + in_ctl);
    KP_CUDA(cudaEventRecord(t->ev[EV_H2D], st));
    KP_CUDA(cudaEventRecord(t->ev[EV_FUSED0], st));
    for (uint32_t i = 0; i < nc; i++) {
        const uint32_t k = nc - 1 - i;
        if (h_ctl[k] == 0) continue;
        KP_LAUNCH(kp_launch_fused(c, d, t->fclasses.c[k], (const uint32_t*)(din + in_lists) + (size_t)k * S, fctl + k, fctl + 8 + k,
                                  c.sel_out, fctl + 16, h_ctl[k], st));
    }
    KP_CUDA(cudaEventRecord(t->ev[EV_FUSED1], st));
    KP_LAUNCH(kp_launch_scan(c.tcount, c.toff32, S, c.scan_tmp, &c.totals[3], st));
    KP_LAUNCH(kp_launch_tokens_pack(c, 0, compact, st));
    KP_CUDA(cudaEventRecord(t->ev[EV_PACK], st));
    KP_CUDA(cudaMemcpyAsync(dout, din, 272, cudaMemcpyDeviceToDevice, st));          // counters + flags ride out with the result
    KP_CUDA(cudaMemcpyAsync(t->h_small_out.p, dout, out_bytes, cudaMemcpyDeviceToHost, st));
    KP_CUDA(cudaEventRecord(t->ev[EV_END], st));
    KP_TRY(kp_wait_stream(t));
    const char* hout = (const char*)t->h_small_out.p;
    const uint64_t* r_tot = (const uint64_t*)hout;
    const uint32_t* r_err = (const uint32_t*)(hout + in_err);
    const uint32_t* r_ctl = (const uint32_t*)(hout + in_ctl);
    if (r_err[1]) {
        kp_set_error("sentence offsets are not ascending or exceed the text length");
        return KP_ERR_ARG;
    }
    if (r_err[0]) {
        kp_set_error("input contains invalid UTF-8");
        return KP_ERR_UTF8;
    }
    if (r_ctl[16]) return 0;                                 // a sentence did not fit its class: the general path
    *n_tokens = r_tot[3];
    *h_tok_off = hout + out_off;
    *h_eos = (const int32_t*)(hout + out_eos);
    *h_tokens = hout + out_tok;
    t->counters.bytes = nbytes;
    t->counters.chars = r_tot[8];
    t->counters.nodes = r_tot[9];
    t->counters.tokens = r_tot[3];
    t->counters.sentences = S;
    t->profile.fused_sentences = (uint32_t)r_tot[10];
    t->profile.chunks = 1;
    cudaEventElapsedTime(&t->profile.h2d_ms, t->ev[EV_START], t->ev[EV_H2D]);
    cudaEventElapsedTime(&t->profile.fused_ms, t->ev[EV_FUSED0], t->ev[EV_FUSED1]);
    cudaEventElapsedTime(&t->profile.d2h_ms, t->ev[EV_PACK], t->ev[EV_END]);
    cudaEventElapsedTime(&t->profile.total_ms, t->ev[EV_START], t->ev[EV_END]);
    t->small_calls++;
    return 1;
}

// Host text in, host result out, chunk by chunk.  compact = kp_token8 records + 32-bit offsets.
// The device result of a chunk is sized by its token COUNT of the previous call when that is known
// to be enough, so the D2H moves the tokens that exist, not an upper bound.
static int kp_tokenize_host(kp_tokenizer* t, const uint8_t* utf8, const uint64_t* offsets, uint64_t n_sent, bool compact) {
    if (!t || !offsets) return KP_ERR_ARG;
    t->res_tok_off = nullptr;
    if (n_sent && offsets[n_sent] > offsets[0] && !utf8) return KP_ERR_ARG;
    for (uint64_t s = 0; s < n_sent; s++)
        if (offsets[s + 1] < offsets[s]) {
            kp_set_error("offsets[%llu] > offsets[%llu]", (unsigned long long)s, (unsigned long long)(s + 1));
            return KP_ERR_ARG;
        }
    KP_CUDA(cudaSetDevice(t->device));
    begin_call(t);
    if (n_sent >= 1 && n_sent <= KP_SMALL_SENT && offsets[n_sent] - offsets[0] <= KP_SMALL_BYTES && t->fused_ok &&
        t->path_mode != KP_PATH_PIPELINE && !t->count_work) {
        uint64_t ntok = 0;
        const int rc = tokenize_small(t, utf8, offsets, n_sent, compact, &ntok, &t->res_tok_off, &t->res_tokens, &t->res_eos);
        if (rc < 0) return rc;
        if (rc == 1) {
            t->last_tokens = ntok;
            return KP_OK;
        }
        t->res_tok_off = nullptr;
        begin_call(t);
    }
    cudaStream_t st = t->stream;
    const size_t tok_sz = compact ? sizeof(kp_token8) : sizeof(kp_token);
    const size_t off_sz = compact ? sizeof(uint32_t) : sizeof(uint64_t);
    KP_TRY(t->h_tok_off.ensure(off_sz * (n_sent + 1), 0));
    KP_TRY(t->h_eos.ensure(sizeof(int32_t) * (n_sent + 1), 0));
    memset(t->h_tok_off.p, 0, off_sz);
    uint64_t tok_total = 0;
    StageTimes times;
    float h2d_ms = 0, d2h_ms = 0, total_ms = 0;
    uint64_t s0 = 0;
    while (s0 < n_sent) {
        // greedy chunk: as many whole sentences as fit in chunk_bytes (at least one)
        uint64_t s1 = s0 + 1;
        while (s1 < n_sent && offsets[s1 + 1] - offsets[s0] <= t->chunk_bytes && s1 - s0 < (1u << 30)) s1++;
        const uint64_t nbytes = offsets[s1] - offsets[s0];
        const uint64_t S = s1 - s0;
        if (nbytes >= (1ull << 31)) return KP_ERR_TOO_LARGE;
        KP_TRY(t->text.ensure(nbytes + 16));
        KP_TRY(t->off.ensure(sizeof(uint64_t) * (S + 1)));
        KP_TRY(t->d_tok_off.ensure(sizeof(uint64_t) * (S + 1)));
        KP_TRY(t->d_eos.ensure(sizeof(int32_t) * (S + 1)));
        KP_TRY(t->d_tokens.ensure(sizeof(kp_token) * (nbytes + S + 1)));
        KP_CUDA(cudaEventRecord(t->ev[EV_START], st));
        if (nbytes) KP_CUDA(cudaMemcpyAsync(t->text.p, utf8 + offsets[s0], nbytes, cudaMemcpyHostToDevice, st));
        KP_CUDA(cudaMemcpyAsync(t->off.p, offsets + s0, sizeof(uint64_t) * (S + 1), cudaMemcpyHostToDevice, st));
        kp_chunk c;
        memset(&c, 0, sizeof(c));
        c.text = t->text.as<uint8_t>();
        c.off = t->off.as<uint64_t>();
        c.base = offsets[s0];
        c.S = (uint32_t)S;
        c.B = (uint32_t)nbytes;
        c.tok_off = t->d_tok_off.p;
        c.eos_cost = t->d_eos.as<int32_t>();
        c.tokens = t->d_tokens.p;
        uint64_t ntok = 0;
        KP_TRY(run_chunk(t, c, tok_total, compact, &ntok, &times));
        if (compact && tok_total + ntok >= (1ull << 32)) return KP_ERR_TOO_LARGE;
        KP_TRY(t->h_tokens.ensure(tok_sz * (tok_total + ntok + 1), tok_sz * tok_total));
        if (ntok)
            KP_CUDA(cudaMemcpyAsync((char*)t->h_tokens.p + tok_sz * tok_total, c.tokens, tok_sz * ntok,
                                    cudaMemcpyDeviceToHost, st));
        KP_CUDA(cudaMemcpyAsync((char*)t->h_tok_off.p + off_sz * s0, c.tok_off, off_sz * (S + 1), cudaMemcpyDeviceToHost, st));
        KP_CUDA(cudaMemcpyAsync(t->h_eos.as<int32_t>() + s0, c.eos_cost, sizeof(int32_t) * S, cudaMemcpyDeviceToHost, st));
        KP_CUDA(cudaEventRecord(t->ev[EV_END], st));
        KP_TRY(kp_wait_stream(t));
        float ms = 0;
        cudaEventElapsedTime(&ms, t->ev[EV_START], t->ev[EV_H2D]);    h2d_ms += ms;
        cudaEventElapsedTime(&ms, t->ev[EV_PACK], t->ev[EV_END]);     d2h_ms += ms;
        cudaEventElapsedTime(&ms, t->ev[EV_START], t->ev[EV_END]);    total_ms += ms;
        tok_total += ntok;
        s0 = s1;
    }
    store_times(t, times);
    t->profile.h2d_ms = h2d_ms;
    t->profile.d2h_ms = d2h_ms;
    t->profile.total_ms = total_ms;
    t->last_tokens = tok_total;
    t->res_tok_off = t->h_tok_off.p;
    t->res_tokens = t->h_tokens.p;
    t->res_eos = t->h_eos.as<int32_t>();
    return KP_OK;
}

// ---- two-phase single-chunk pass for the multi-GPU path (kp_queue.cu) --------------------------------
int kp_pass_begin(kp_tokenizer* t, const uint8_t* utf8, const uint64_t* offsets, uint64_t n_sent, uint64_t* n_tokens) {
    if (!t || !offsets || !n_tokens) return KP_ERR_ARG;
    const uint64_t nbytes = offsets[n_sent] - offsets[0];
    if (nbytes > t->chunk_bytes || nbytes >= (1ull << 31) || n_sent >= (1u << 30)) {
        kp_set_error("shard of %llu bytes exceeds the tokenizer's chunk size", (unsigned long long)nbytes);
        return KP_ERR_TOO_LARGE;
    }
    KP_CUDA(cudaSetDevice(t->device));
    begin_call(t);
    cudaStream_t st = t->stream;
    KP_TRY(t->text.ensure(nbytes + 16));
    KP_TRY(t->off.ensure(sizeof(uint64_t) * (n_sent + 1)));
    KP_TRY(t->d_tok_off.ensure(sizeof(uint64_t) * (n_sent + 1)));
    KP_TRY(t->d_eos.ensure(sizeof(int32_t) * (n_sent + 1)));
    KP_TRY(t->d_tokens.ensure(sizeof(kp_token) * (nbytes + n_sent + 1)));
    KP_CUDA(cudaEventRecord(t->ev[EV_START], st));
    if (nbytes) KP_CUDA(cudaMemcpyAsync(t->text.p, utf8 + offsets[0], nbytes, cudaMemcpyHostToDevice, st));
    KP_CUDA(cudaMemcpyAsync(t->off.p, offsets, sizeof(uint64_t) * (n_sent + 1), cudaMemcpyHostToDevice, st));
    kp_chunk& c = t->pass;
    memset(&c, 0, sizeof(c));
    c.text = t->text.as<uint8_t>();
    c.off = t->off.as<uint64_t>();
    c.base = offsets[0];
    c.S = (uint32_t)n_sent;
    c.B = (uint32_t)nbytes;
    c.tok_off = t->d_tok_off.p;
    c.eos_cost = t->d_eos.as<int32_t>();
    c.tokens = t->d_tokens.p;
    StageTimes times;
    KP_TRY(chunk_compute(t, c, n_tokens, &times));
    store_times(t, times);
    t->last_tokens = *n_tokens;
    return KP_OK;
}

// packs the pass's tokens (compact form, offsets rebased by tok_base) and leaves them on the device
int kp_pass_pack(kp_tokenizer* t, uint64_t tok_base, const uint32_t** d_tok_off, const kp_token8** d_tokens,
                 const int32_t** d_eos, cudaStream_t* stream) {
    KP_CUDA(cudaSetDevice(t->device));
    KP_TRY(chunk_pack(t, t->pass, tok_base, true));
    *d_tok_off = (const uint32_t*)t->pass.tok_off;
    *d_tokens = (const kp_token8*)t->pass.tokens;
    *d_eos = t->pass.eos_cost;
    *stream = t->stream;
    return KP_OK;
}

extern "C" int kp_tokenize_batch(kp_tokenizer* t, const uint8_t* utf8, const uint64_t* offsets, uint64_t n_sent,
                                 kp_result* out) {
    if (!out) return KP_ERR_ARG;
    KP_TRY(kp_tokenize_host(t, utf8, offsets, n_sent, false));
    out->n_sent = n_sent;
    out->n_tokens = t->last_tokens;
    out->tok_off = (const uint64_t*)t->res_tok_off;
    out->tokens = (const kp_token*)t->res_tokens;
    out->eos_cost = t->res_eos;
    return KP_OK;
}

extern "C" int kp_tokenize_batch8(kp_tokenizer* t, const uint8_t* utf8, const uint64_t* offsets, uint64_t n_sent,
                                  kp_result8* out) {
    if (!out) return KP_ERR_ARG;
    KP_TRY(kp_tokenize_host(t, utf8, offsets, n_sent, true));
    out->n_sent = n_sent;
    out->n_tokens = t->last_tokens;
    out->tok_off = (const uint32_t*)t->res_tok_off;
    out->tokens = (const kp_token8*)t->res_tokens;
    out->eos_cost = t->res_eos;
    return KP_OK;
}

// kp_token8 -> kp_token on the host: walk every sentence's tokens backwards from its EOS token
// (position = sentence bytes, start = n_chars carried by the EOS record); see the header.
// Sentences [s0, s1); returns the first sentence whose last record is not EOS, or UINT64_MAX.
static uint64_t kp_expand_range(const kp_result8* r, const uint64_t* offsets, kp_token* out, uint64_t s0, uint64_t s1) {
    for (uint64_t s = s0; s < s1; s++) {
        const uint32_t a = r->tok_off[s], b = r->tok_off[s + 1];
        if (a == b) continue;                     // empty path (dp[EOS] = INF): no tokens, no EOS
        uint32_t pos = (uint32_t)(offsets[s + 1] - offsets[s]), start = 0;
        for (uint32_t k = b; k-- > a;) {
            const kp_token8 x = r->tokens[k];
            const uint32_t cls = x.id_cls >> 30;
            kp_token o;
            o.id = (int32_t)(x.id_cls & 0x3FFFFFFFu);
            o.cls = (uint8_t)cls;
            o.reserved = 0;
            if (k == b - 1) {
                if (cls != KP_CLASS_DUMMY) return s;
                start = (uint32_t)x.byte_len | ((uint32_t)x.char_len << 16);
                o.char_len = 3;
            } else {
                pos -= x.byte_len;
                start -= x.char_len;
                o.char_len = x.char_len;
            }
            o.position = pos;
            o.start = start;
            out[k] = o;
        }
    }
    return UINT64_MAX;
}

// Large results are split over a few host threads (contiguous sentence ranges of about equal token
// counts): one thread streams ~12 GB/s of records, a batch of 2M tokens would take it 4 ms.
extern "C" int kp_expand_tokens8(const kp_result8* r, const uint64_t* offsets, kp_token* out) {
    if (!r || !offsets || (r->n_tokens && (!out || !r->tokens)) || (r->n_sent && !r->tok_off)) return KP_ERR_ARG;
    const uint64_t per_thread = 1u << 17;
    uint64_t nthr = std::min<uint64_t>({r->n_tokens / per_thread, (uint64_t)std::thread::hardware_concurrency(), 8});
    uint64_t bad = UINT64_MAX;
    if (nthr <= 1) {
        bad = kp_expand_range(r, offsets, out, 0, r->n_sent);
    } else {
        std::vector<uint64_t> cut(nthr + 1, r->n_sent), first(nthr, UINT64_MAX);
        cut[0] = 0;
        for (uint64_t i = 1; i < nthr; i++)       // first sentence starting at or after the i-th share of the tokens
            cut[i] = (uint64_t)(std::lower_bound(r->tok_off, r->tok_off + r->n_sent, (uint32_t)(r->n_tokens * i / nthr)) - r->tok_off);
        std::vector<std::thread> th;
        for (uint64_t i = 1; i < nthr; i++)
            th.emplace_back([&, i] { first[i] = kp_expand_range(r, offsets, out, cut[i], cut[i + 1]); });
        first[0] = kp_expand_range(r, offsets, out, cut[0], cut[1]);
        for (auto& x : th) x.join();
        for (uint64_t i = 0; i < nthr; i++) bad = std::min(bad, first[i]);
    }
    if (bad != UINT64_MAX) {
        kp_set_error("sentence %llu: last token is not EOS", (unsigned long long)bad);
        return KP_ERR_ARG;
    }
    return KP_OK;
}

extern "C" int kp_tokenize(kp_tokenizer* t, const uint8_t* utf8, uint64_t len, kp_result* out) {
    const uint64_t off[2] = {0, len};
    return kp_tokenize_batch(t, utf8, off, 1, out);
}

// Lattice::build + viterbi internals of one sentence, in the reference's node order (BOS first).
extern "C" int kp_lattice_dump(kp_tokenizer* t, const uint8_t* utf8, uint64_t len, kp_lattice* out) {
    if (!t || !out || (len && !utf8)) return KP_ERR_ARG;
    kp_result r;
    const int mode = t->path_mode;
    t->path_mode = KP_PATH_PIPELINE;          // the node table exists only in the pipeline's scratch
    const int rc_tok = kp_tokenize(t, utf8, len, &r);
    t->path_mode = mode;
    KP_TRY(rc_tok);
    // scratch of the (single-chunk) pass is still intact
    const uint32_t N = (uint32_t)(t->counters.nodes - 1);   // device nodes (no BOS)
    const uint32_t NB = (uint32_t)t->counters.chars + 1;
    // pre_nodes for every node: recomputed on the device from the dp table (the sweep keeps dp only)
    {
        KP_TRY(t->pre.ensure(sizeof(uint32_t) * ((size_t)N + 1)));
        kp_chunk c;
        memset(&c, 0, sizeof(c));
        c.N = N;
        c.rec = t->rec.as<uint4>();
        c.tgt = t->tgt.as<uint2>();
        c.ndp = t->ndp.as<int32_t>();
        c.bnode = t->bnode.as<uint32_t>();
        c.boff = t->boff.as<uint32_t>();
        c.pre = t->pre.as<uint32_t>();
        KP_TRY(kp_launch_fill_pre(c, t->dict->view, t->stream));
        KP_CUDA(cudaStreamSynchronize(t->stream));
    }
    std::vector<uint4> rec(N), binfo(NB);
    std::vector<uint32_t> bnode(N), pre(N);
    std::vector<int32_t> ndp(N);
    KP_CUDA(cudaMemcpy(rec.data(), t->rec.p, sizeof(uint4) * N, cudaMemcpyDeviceToHost));
    KP_CUDA(cudaMemcpy(binfo.data(), t->binfo.p, sizeof(uint4) * NB, cudaMemcpyDeviceToHost));

    KP_CUDA(cudaMemcpy(bnode.data(), t->bnode.p, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost));
    KP_CUDA(cudaMemcpy(pre.data(), t->pre.p, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost));
    KP_CUDA(cudaMemcpy(ndp.data(), t->ndp.p, sizeof(int32_t) * N, cudaMemcpyDeviceToHost));
    t->lattice_nodes.assign((size_t)N + 1, kp_lattice_node{});
    kp_lattice_node& bos = t->lattice_nodes[0];
    bos.id = 0;
    bos.cls = KP_CLASS_DUMMY;
    bos.dp = INT_MIN;
    bos.pre = -1;
    for (uint32_t i = 0; i < N; i++) {
        kp_lattice_node& o = t->lattice_nodes[i + 1];
        const uint4 x = rec[i];
        const uint32_t kind = x.x >> KP_KIND_SHIFT;
        o.id = (int32_t)(x.x & KP_ID_MASK);
        o.cls = (uint8_t)kind;
        o.byte_pos = binfo[x.y].x;
        o.char_pos = x.y;
        o.end_char = kind == KP_CLASS_DUMMY ? x.y + 1 : x.y + (x.w >> 16);
        o.left_id = (int16_t)(x.z & 0xFFFF);
        o.right_id = (int16_t)(x.z >> 16);
        o.cost = (int16_t)(x.w & 0xFFFF);
        o.dp = ndp[i];
        o.pre = pre[i] == KP_NONE ? -1 : (bnode[pre[i]] == KP_NONE ? 0 : (int32_t)bnode[pre[i]] + 1);
    }
    out->n_nodes = (uint64_t)N + 1;
    out->nodes = t->lattice_nodes.data();
    return KP_OK;
}

extern "C" int kp_da_common_prefix(kp_tokenizer* t, const uint8_t* utf8, uint64_t len, int expand_dup, int64_t* ids,
                                   uint64_t* byte_lens, uint64_t cap, uint64_t* n) {
    if (!t || !n || (len && !utf8) || (cap && (!ids || !byte_lens)) || len >= (1ull << 31)) return KP_ERR_ARG;
    KP_CUDA(cudaSetDevice(t->device));
    cudaStream_t st = t->stream;
    const uint32_t dcap = (uint32_t)std::min<uint64_t>(cap, 1u << 20);
    KP_TRY(t->text.ensure(len + 16));
    KP_TRY(t->rec.ensure(sizeof(int64_t) * (dcap + 1)));
    KP_TRY(t->red.ensure(sizeof(uint64_t) * (dcap + 1)));
    KP_TRY(t->err.ensure(sizeof(uint32_t) * 2));
    if (len) KP_CUDA(cudaMemcpyAsync(t->text.p, utf8, len, cudaMemcpyHostToDevice, st));
    KP_TRY(kp_launch_common_prefix(t->dict->view, t->text.as<uint8_t>(), (uint32_t)len, expand_dup, t->rec.as<int64_t>(),
                                   t->red.as<uint64_t>(), dcap, t->err.as<uint32_t>(), st));
    uint32_t hn = 0;
    KP_CUDA(cudaMemcpyAsync(&hn, t->err.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    KP_CUDA(cudaStreamSynchronize(st));
    const uint32_t m = std::min(hn, dcap);
    if (m) {
        KP_CUDA(cudaMemcpy(ids, t->rec.p, sizeof(int64_t) * m, cudaMemcpyDeviceToHost));
        KP_CUDA(cudaMemcpy(byte_lens, t->red.p, sizeof(uint64_t) * m, cudaMemcpyDeviceToHost));
    }
    *n = hn;
    return KP_OK;
}
