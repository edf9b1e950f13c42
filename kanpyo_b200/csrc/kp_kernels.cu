// kp_kernels.cu — hand-written sm_100a kernels of the Kanpyo hot path.
//
// Integer / byte work, HBM- and L2-gather bound: no tensor cores.  The dictionary
// (17.5 MB for IPADIC) is L2-resident on B200 (126 MB L2); node / bucket arrays stream through HBM.
//
// Reference semantics implemented (file:line in the reference tree):
//   trie walk                 DoubleArray::search_common_prefix_of   kanpyo-dict/src/trie/da.rs:155-182
//   duplicate expansion       IndexTable::search_common_prefix_of    kanpyo-dict/src/index.rs:40-53
//   known / unknown nodes     Lattice::process_{known,unknown}_words src/lattice.rs:24-99
//   node order, end buckets   Lattice::add_*_node / `edges`          src/lattice.rs:156-201
//   forward DP + back-trace   Lattice::viterbi                       src/lattice.rs:116-154
//   Node -> Token             Tokenizer::tokenize                    src/tokenizer.rs:16-45
#include "kp_kernels.cuh"

#define KP_FULL 0xFFFFFFFFu

static inline int kp_launch_check(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        kp_set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return KP_ERR_CUDA;
    }
    return 1;
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// =================================================================================================
// Exclusive scans, one launch each (decoupled look-back).  Two arrays ride together.
// =================================================================================================
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

uint32_t kp_scan_tmp_elems(uint32_t n) { return 2 * ((n + SCAN_TILE - 1) / SCAN_TILE + 1); }

// Single pass: tile sums chained across blocks by decoupled look-back.
// `state[t]` = flag << 62 | value: flag 1 = sum of tile t alone, flag 2 = sum of tiles 0..t.  Blocks take
// their tile by ticket, so a block's predecessors are running or done.  One launch (plus a memset of
// the state), the input read once.
__device__ __forceinline__ uint64_t kp_scan_lookback(unsigned long long* state, uint32_t tile, uint64_t agg) {
    // called by the 32 lanes of warp 0; returns the sum of tiles 0..tile-1
    uint64_t excl = 0;
    if (tile == 0) {
        if (lane_id() == 0) atomicExch(&state[0], (2ull << 62) | agg);
        return 0;
    }
    if (lane_id() == 0) atomicExch(&state[tile], (1ull << 62) | agg);
    int32_t j0 = (int32_t)tile - 1;
    while (true) {
        const int32_t j = j0 - (int32_t)lane_id();
        unsigned long long v = 2ull << 62;               // before tile 0: inclusive sum 0
        if (j >= 0) {
            do {
                v = *(volatile unsigned long long*)&state[j];
            } while ((v >> 62) == 0);
        }
        const uint32_t done = __ballot_sync(KP_FULL, (v >> 62) == 2);
        const uint32_t upto = done ? (uint32_t)__ffs(done) - 1 : 31u;   // up to the nearest inclusive sum
        uint64_t part = lane_id() <= upto ? (uint64_t)(v & ((1ull << 62) - 1)) : 0ull;
        for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(KP_FULL, part, o);
        excl += __shfl_sync(KP_FULL, part, 0);
        if (done) break;
        j0 -= 32;
    }
    if (lane_id() == 0) atomicExch(&state[tile], (2ull << 62) | (excl + agg));
    return excl;
}

template <bool TWO>
__global__ void __launch_bounds__(SCAN_THREADS) kp_scan_onepass(const uint32_t* __restrict__ a,
                                                                const uint32_t* __restrict__ b, uint32_t n,
                                                                uint32_t* __restrict__ out_a, uint32_t* __restrict__ out_b,
                                                                uint64_t* __restrict__ tmp, uint32_t ntiles,
                                                                uint64_t* __restrict__ total_a,
                                                                uint64_t* __restrict__ total_b) {
    __shared__ uint32_t wsa[SCAN_THREADS / 32], wsb[SCAN_THREADS / 32];
    __shared__ uint32_t sh_tile;
    __shared__ uint64_t sh_ea, sh_eb;
    unsigned long long* state_a = (unsigned long long*)tmp;
    unsigned long long* state_b = state_a + ntiles;
    if (threadIdx.x == 0) sh_tile = atomicAdd((uint32_t*)(tmp + 2 * (size_t)ntiles), 1u);
    __syncthreads();
    const uint32_t tile = sh_tile;
    const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t va[SCAN_ITEMS], vb[SCAN_ITEMS];
    uint32_t sa = 0, sb = 0;
    const bool full = base + SCAN_ITEMS <= n;     // 32 contiguous, 32-byte aligned bytes per thread
    if (full) {
        *(uint4*)&va[0] = *(const uint4*)(a + base);
        *(uint4*)&va[4] = *(const uint4*)(a + base + 4);
        if (TWO) {
            *(uint4*)&vb[0] = *(const uint4*)(b + base);
            *(uint4*)&vb[4] = *(const uint4*)(b + base + 4);
        }
    }
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint32_t i = base + k;
        if (!full) {
            va[k] = i < n ? a[i] : 0;
            if (TWO) vb[k] = i < n ? b[i] : 0;
        }
        sa += va[k];
        if (TWO) sb += vb[k];
    }
    uint32_t xa = sa, xb = sb;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t ya = __shfl_up_sync(KP_FULL, xa, o);
        const uint32_t yb = __shfl_up_sync(KP_FULL, xb, o);
        if (lane_id() >= (uint32_t)o) {
            xa += ya;
            xb += yb;
        }
    }
    if (lane_id() == 31) {
        wsa[threadIdx.x >> 5] = xa;
        wsb[threadIdx.x >> 5] = xb;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        uint64_t agg_a = 0, agg_b = 0;
        for (int w = 0; w < SCAN_THREADS / 32; w++) {
            agg_a += wsa[w];
            agg_b += wsb[w];
        }
        const uint64_t ea = kp_scan_lookback(state_a, tile, agg_a);
        const uint64_t eb = TWO ? kp_scan_lookback(state_b, tile, agg_b) : 0ull;
        if (threadIdx.x == 0) {
            sh_ea = ea;
            sh_eb = eb;
            if (tile == ntiles - 1) {              // the last tile holds the end of the array: totals out
                out_a[n] = (uint32_t)(ea + agg_a);
                if (total_a) *total_a = ea + agg_a;
                if (TWO) {
                    out_b[n] = (uint32_t)(eb + agg_b);
                    if (total_b) *total_b = eb + agg_b;
                }
            }
        }
    }
    __syncthreads();
    uint32_t oa = (uint32_t)sh_ea, ob = TWO ? (uint32_t)sh_eb : 0;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) {
        oa += wsa[w];
        ob += wsb[w];
    }
    oa += xa - sa;
    ob += xb - sb;
    uint32_t ra[SCAN_ITEMS], rb[SCAN_ITEMS];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        ra[k] = oa;
        oa += va[k];
        if (TWO) {
            rb[k] = ob;
            ob += vb[k];
        }
    }
    if (full) {                                   // out_a / out_b have n + 1 elements: base + 8 <= n is in range
        *(uint4*)(out_a + base) = *(uint4*)&ra[0];
        *(uint4*)(out_a + base + 4) = *(uint4*)&ra[4];
        if (TWO) {
            *(uint4*)(out_b + base) = *(uint4*)&rb[0];
            *(uint4*)(out_b + base + 4) = *(uint4*)&rb[4];
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            const uint32_t i = base + k;
            if (i < n) {
                out_a[i] = ra[k];
                if (TWO) out_b[i] = rb[k];
            }
        }
    }
}

template <bool TWO>
static int kp_scan_impl(const uint32_t* a, const uint32_t* b, uint32_t* oa, uint32_t* ob, uint32_t n, uint64_t* tmp,
                        uint64_t* ta, uint64_t* tb, cudaStream_t st) {
    uint32_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (ntiles == 0) ntiles = 1;   // n == 0: still writes out[0] = 0 and the totals
    cudaMemsetAsync(tmp, 0, sizeof(uint64_t) * (2 * (size_t)ntiles + 1), st);   // tile states + ticket
    kp_scan_onepass<TWO><<<ntiles, SCAN_THREADS, 0, st>>>(a, b, n, oa, ob, tmp, ntiles, ta, tb);
    int rc = kp_launch_check("kp_scan");
    return rc < 0 ? rc : 1;
}

int kp_launch_scan(const uint32_t* in, uint32_t* out, uint32_t n, uint64_t* tmp, uint64_t* total, cudaStream_t st) {
    return kp_scan_impl<false>(in, nullptr, out, nullptr, n, tmp, total, nullptr, st);
}
int kp_launch_scan2(const uint32_t* a, const uint32_t* b, uint32_t* oa, uint32_t* ob, uint32_t n, uint64_t* tmp,
                    uint64_t* ta, uint64_t* tb, cudaStream_t st) {
    return kp_scan_impl<true>(a, b, oa, ob, n, tmp, ta, tb, st);
}

// =================================================================================================
// Prep: UTF-8 validation + chars per sentence; then per-boundary info.  One warp per sentence.
// =================================================================================================
#ifndef KP_PREP_THREADS
#define KP_PREP_THREADS 256
#endif
constexpr int PREP_THREADS = KP_PREP_THREADS;
constexpr uint32_t LEN_BINS = 4096;    // counting-sort bins of the Viterbi work order (sentence length in chars)

__device__ __forceinline__ bool is_cont(uint32_t c) { return (c & 0xC0u) == 0x80u; }
__device__ __forceinline__ uint32_t lead_len(uint32_t c) {  // 0 = not a valid lead byte
    if (c < 0x80u) return 1;
    if (c >= 0xC2u && c <= 0xDFu) return 2;
    if (c >= 0xE0u && c <= 0xEFu) return 3;
    if (c >= 0xF0u && c <= 0xF4u) return 4;
    return 0;
}

// Rust's &str is valid UTF-8 by construction (src/tokenizer.rs:16 takes &str); the C ABI has to check.
__global__ void __launch_bounds__(PREP_THREADS) kp_prep_count(const uint8_t* __restrict__ text,
                                                              const uint64_t* __restrict__ off, uint64_t base,
                                                              const uint32_t* __restrict__ sel,
                                                              uint32_t S, uint32_t B, uint32_t* __restrict__ nchar,
                                                              uint32_t* __restrict__ err, uint32_t* __restrict__ lenhist) {
    uint32_t s = (blockIdx.x * PREP_THREADS + threadIdx.x) >> 5;
    if (s >= S) return;
    const uint32_t so = sel ? sel[s] : s;         // sentence of the chunk this pass's slot s stands for
    uint64_t lo64 = off[so] - base, hi64 = off[so + 1] - base;
    if (off[so] < base || hi64 < lo64 || hi64 > B) {
        if (lane_id() == 0) {
            atomicOr(&err[1], 1u);
            nchar[s] = 0;
            atomicAdd(&lenhist[0], 1u);
        }
        return;
    }
    uint32_t lo = (uint32_t)lo64, hi = (uint32_t)hi64;
    uint32_t cnt = 0, conts = 0, claimed = 0;
    bool bad = false;
    for (uint32_t i = lo + lane_id(); i < hi; i += 32) {
        uint32_t c = text[i];
        if (!is_cont(c)) {
            cnt++;
            uint32_t L = lead_len(c);
            if (L == 0 || i + L > hi) {
                bad = true;
            } else if (L > 1) {
                claimed += L - 1;
                uint32_t c1 = text[i + 1];
                uint32_t lo1 = 0x80u, hi1 = 0xBFu;
                if (c == 0xE0u) lo1 = 0xA0u;
                if (c == 0xEDu) hi1 = 0x9Fu;
                if (c == 0xF0u) lo1 = 0x90u;
                if (c == 0xF4u) hi1 = 0x8Fu;
                if (c1 < lo1 || c1 > hi1) bad = true;
                for (uint32_t k = 2; k < L; k++)
                    if (!is_cont(text[i + k])) bad = true;
            }
        } else {
            conts++;
        }
    }
    // every lead byte verified the continuation bytes it claims, and claims cannot overlap; the
    // sentence has no stray continuation byte iff their number equals the number claimed
    if (__reduce_add_sync(KP_FULL, conts) != __reduce_add_sync(KP_FULL, claimed)) bad = true;
    cnt = __reduce_add_sync(KP_FULL, cnt);
    if (__any_sync(KP_FULL, bad) && lane_id() == 0) atomicOr(&err[0], 1u);
    if (lane_id() == 0) {
        nchar[s] = cnt;
        atomicAdd(&lenhist[min(cnt, LEN_BINS - 1)], 1u);   // length histogram of the Viterbi work order
    }
}

__global__ void __launch_bounds__(PREP_THREADS) kp_prep_fill(const uint8_t* __restrict__ text,
                                                             const uint64_t* __restrict__ off, uint64_t base,
                                                             const uint32_t* __restrict__ sel, uint32_t S,
                                                             const uint32_t* __restrict__ coff, kp_ddict d,
                                                             uint4* __restrict__ binfo, uint32_t* __restrict__ bcount,
                                                             uint32_t* __restrict__ ucount, uint32_t* __restrict__ cursor,
                                                             uint32_t* __restrict__ order) {
    uint32_t s = (blockIdx.x * PREP_THREADS + threadIdx.x) >> 5;
    if (s >= S) return;
    const uint32_t lane = lane_id();
    const uint32_t so = sel ? sel[s] : s;
    const uint32_t lo = (uint32_t)(off[so] - base), hi = (uint32_t)(off[so + 1] - base);
    const uint32_t bb = coff[s] + s;              // first boundary of this sentence
    const uint32_t n = coff[s + 1] - coff[s];     // chars
    if (lane == 0)                                // counting-sort scatter of the Viterbi work order (cursor = scanned histogram)
        order[atomicAdd(&cursor[min(n, LEN_BINS - 1)], 1u)] = s;
    // forward: byte offset + class of every char (Lattice::build's chars().enumerate(), lattice.rs:105)
    uint32_t run = 0;
    for (uint32_t i0 = lo; i0 < hi; i0 += 32) {
        uint32_t i = i0 + lane;
        uint32_t c = i < hi ? text[i] : 0x80u;
        bool st = !is_cont(c);
        uint32_t m = __ballot_sync(KP_FULL, st);
        if (st) {
            uint32_t cp;
            if (c < 0x80u) cp = c;
            else if (c < 0xE0u) cp = ((c & 0x1Fu) << 6) | (text[i + 1] & 0x3Fu);
            else if (c < 0xF0u) cp = ((c & 0x0Fu) << 12) | ((text[i + 1] & 0x3Fu) << 6) | (text[i + 2] & 0x3Fu);
            else cp = ((c & 0x07u) << 18) | ((text[i + 1] & 0x3Fu) << 12) | ((text[i + 2] & 0x3Fu) << 6) | (text[i + 3] & 0x3Fu);
            // CharCategoryDef::char_category: out-of-table code points use entry 0 (char_category_def.rs:33-38)
            uint32_t cat = d.cat[cp < d.n_cat ? cp : 0];
            uint32_t p = run + __popc(m & lanemask_lt());
            binfo[bb + p] = make_uint4(i, hi, 0u, cat);
            bcount[bb + p] = p == 0 ? 1u : 0u;    // BOS sits in edges[0] (lattice.rs:156-164)
            ucount[bb + p] = 0u;
        }
        run += __popc(m);
    }
    if (lane == 0) {
        binfo[bb + n] = make_uint4(hi, hi, bb + n + 1, 0xFFFFu);   // EOS boundary (lattice.rs:165-175)
        bcount[bb + n] = n == 0 ? 1u : 0u;
        ucount[bb + n] = 0u;
    }
    __syncwarp();
    // backward: end of the same-class run each char belongs to (lattice.rs:69-84), capped at 1024 chars
    uint32_t carry = n;
    for (int32_t k = (int32_t)((n + 31) / 32) - 1; k >= 0; k--) {
        uint32_t p = (uint32_t)k * 32 + lane;
        bool valid = p < n;
        uint32_t cat = valid ? binfo[bb + p].w : 0xFFFEu;
        uint32_t catn = (p + 1 < n) ? binfo[bb + p + 1].w : 0xFFFDu;
        bool last = valid && cat != catn;
        uint32_t m = __ballot_sync(KP_FULL, last);
        uint32_t mge = m & ~lanemask_lt();
        uint32_t runend = mge ? (uint32_t)k * 32 + (uint32_t)__ffs(mge) : carry;   // boundary after the run's last char
        carry = __shfl_sync(KP_FULL, runend, 0);
        if (valid) {
            bool group = (d.catinfo[cat].flags & 2u) != 0;
            uint32_t uend = group ? min(runend, p + KP_MAX_UNKNOWN_LEN) : p + 1;
            binfo[bb + p].z = bb + uend;
        }
    }
}

int kp_launch_prep_count(const kp_chunk& c, cudaStream_t st) {
    if (c.S == 0) return 0;
    uint32_t blocks = (uint32_t)(((uint64_t)c.S * 32 + PREP_THREADS - 1) / PREP_THREADS);
    kp_prep_count<<<blocks, PREP_THREADS, 0, st>>>(c.text, c.off, c.base, c.sel, c.S, c.B, c.nchar, c.err, c.lenhist);
    return kp_launch_check("kp_prep_count");
}

int kp_launch_prep_fill(const kp_chunk& c, const kp_ddict& d, cudaStream_t st) {
    if (c.S == 0) return 0;
    uint32_t blocks = (uint32_t)(((uint64_t)c.S * 32 + PREP_THREADS - 1) / PREP_THREADS);
    kp_prep_fill<<<blocks, PREP_THREADS, 0, st>>>(c.text, c.off, c.base, c.sel, c.S, c.coff, d, c.binfo, c.bcount, c.ucount,
                                                  c.lenhist, c.order);
    return kp_launch_check("kp_prep_fill");
}

// =================================================================================================
// Lattice build: one thread per boundary.  Pass 1 (kp_lattice_count below; kp_lattice_walk<false, true>
// is the plain statement of the same walk and serves the work counters) walks the double array from
// that char to the end of the sentence and counts nodes per start boundary and per end boundary;
// pass 2 (kp_lattice_walk<true, false>, after the scans) writes the node records in the reference's
// insertion order, replaying the hits pass 1 remembered.
// =================================================================================================
#ifndef KP_LAT_THREADS
#define KP_LAT_THREADS 64     // small blocks: walk lengths are heavy-tailed, a block lives as long as its longest walk
#endif
constexpr int LAT_THREADS = KP_LAT_THREADS;

constexpr uint32_t LAT_HITS = 4;   // trie hits per start boundary remembered from the counting walk
#ifndef KP_WALK_T1
#define KP_WALK_T1 1
#endif
#ifndef KP_WALK_SKIP_MID
#define KP_WALK_SKIP_MID 1
#endif
#ifndef KP_CNT_MINB
#define KP_CNT_MINB 32
#endif
#ifndef KP_FILL_MINB
#define KP_FILL_MINB 20
#endif

// A remembered hit is {id, chars | (duplicates << 16)}: the fill pass needs neither the trie nor `dup`.
// A morph {left, right, cost, 0} read as two words: {left | right << 16, cost}.
__device__ __forceinline__ uint2 ld_morph(const short4* __restrict__ m, uint32_t i) {
    return __ldg((const uint2*)m + i);
}
__device__ __forceinline__ uint4 kp_node_rec(uint32_t id, uint32_t kind, uint32_t b, uint2 m, uint32_t nch) {
    return make_uint4(id | (kind << KP_KIND_SHIFT), b, m.x, (m.y & 0xFFFFu) | (nch << 16));
}

// Fill pass, boundaries with more than LAT_HITS hits (a few percent): walk the trie again and write
// one Known node per (hit x duplicate) (lattice.rs:177-188, index.rs:46-51).  Kept out of line so its
// registers do not count against the replay path.  Returns the next free record index.
struct kp_trie_view {
    const int2* da;
    uint32_t da_len;
    const uint16_t* dup;
    const short4* morphs;
};
__device__ __noinline__ uint32_t kp_fill_rewalk(const uint8_t* __restrict__ text, uint32_t bp, uint32_t send,
                                                uint32_t b, kp_trie_view d, uint4* __restrict__ rec, uint32_t o) {
    int prev = KP_ROOT_ID;
    int base = d.da[KP_ROOT_ID].x;
    uint32_t nch = 0;
    for (uint32_t i = bp; i < send; i++) {
        const uint32_t c = text[i];
        nch += !is_cont(c);
        const int q = base + (int)c;                                         // da.rs:160
        if ((uint32_t)q >= d.da_len) break;                                  // Vec::get -> None (da.rs:161)
        const int2 nq = d.da[q];
        if (nq.y != prev) break;                                             // da.rs:162-164
        const int ahead = nq.x;                                              // + TERMINATOR (0), da.rs:165
        if ((uint32_t)ahead < d.da_len) {
            const int2 na = d.da[ahead];
            if (na.y == q && na.x < 0) {                                     // da.rs:167-174
                const uint32_t id = (uint32_t)(-na.x);
                const uint32_t k = (uint32_t)d.dup[id] + 1;                  // index.rs:46-51
                for (uint32_t j = 0; j < k; j++)
                    rec[o++] = kp_node_rec(id + j, KP_CLASS_KNOWN, b, ld_morph(d.morphs, id + j - 1), nch);   // lattice.rs:182
            }
        }
        prev = q;
        base = nq.x;
    }
    return o;
}

template <bool FILL, bool WORK>
__global__ void __launch_bounds__(LAT_THREADS, FILL ? KP_FILL_MINB : KP_CNT_MINB) kp_lattice_walk(const uint8_t* __restrict__ text,
                                                          const uint4* __restrict__ binfo, uint32_t NB, kp_ddict d,
                                                          uint32_t* __restrict__ ncount, uint32_t* __restrict__ bcount,
                                                          uint32_t* __restrict__ ucount, uint8_t* __restrict__ nhit,
                                                          uint4* __restrict__ hits,
                                                          const uint32_t* __restrict__ noff, uint4* __restrict__ rec,
                                                          uint64_t* __restrict__ totals) {
    uint32_t b = blockIdx.x * LAT_THREADS + threadIdx.x;
    uint32_t probes = 0, probes_ok = 0;
    if (b < NB) {
        const uint4 bi = binfo[b];
        const uint32_t bp = bi.x, send = bi.y;
        uint32_t o = FILL ? noff[b] : 0;
        uint32_t total = 0;
        // one Known node per (hit x duplicate) (lattice.rs:177-188, index.rs:46-51), with the duplicate
        // count and the first morph already in registers; the remaining morph loads go out two at a
        // time ahead of the stores that need them
        auto expand_known = [&](uint32_t id, uint32_t y, uint2 m0) {
            const uint32_t nch = y & 0xFFFFu, k = (y >> 16) + 1;
            rec[o++] = kp_node_rec(id, KP_CLASS_KNOWN, b, m0, nch);
            for (uint32_t j = 1; j < k; j += 2) {
                const uint2 ma = ld_morph(d.morphs, id + j - 1);
                const uint2 mb = ld_morph(d.morphs, id + min(j + 1, k - 1) - 1);
                rec[o++] = kp_node_rec(id + j, KP_CLASS_KNOWN, b, ma, nch);
                if (j + 1 < k) rec[o++] = kp_node_rec(id + j + 1, KP_CLASS_KNOWN, b, mb, nch);
            }
        };
        if (bp == send) {
            // EOS: Dummy node with morph (0,0,0) (lattice.rs:165-175)
            if (FILL) rec[o] = make_uint4((uint32_t)KP_CLASS_DUMMY << KP_KIND_SHIFT, b, 0u, 0u);
            total = 1;
        } else {
            // The counting walk remembers its first LAT_HITS hits {id, chars | dups << 16}; the fill pass
            // replays them and walks the trie again only for the few boundaries with more hits than that.
            uint4 h01 = make_uint4(0, 0, 0, 0), h23 = make_uint4(0, 0, 0, 0);
            uint32_t nh = FILL ? nhit[b] : 0;
            if (FILL) h01 = hits[2 * (size_t)b];   // not waiting for nh (unused when nh == 0)
            if (FILL && nh <= LAT_HITS) {
                if (nh > 2) h23 = hits[2 * (size_t)b + 1];
                // first morph of every hit: four independent gathers in flight together
                const uint2 z = make_uint2(0u, 0u);
                const uint2 m0 = nh > 0 ? ld_morph(d.morphs, h01.x - 1) : z;
                const uint2 m1 = nh > 1 ? ld_morph(d.morphs, h01.z - 1) : z;
                const uint2 m2 = nh > 2 ? ld_morph(d.morphs, h23.x - 1) : z;
                const uint2 m3 = nh > 3 ? ld_morph(d.morphs, h23.z - 1) : z;
                if (nh > 0) expand_known(h01.x, h01.y, m0);
                if (nh > 1) expand_known(h01.z, h01.w, m1);
                if (nh > 2) expand_known(h23.x, h23.y, m2);
                if (nh > 3) expand_known(h23.z, h23.w, m3);
            } else if (FILL) {
                if (d.da_len > KP_ROOT_ID)
                    o = kp_fill_rewalk(text, bp, send, b, kp_trie_view{d.da, d.da_len, d.dup, d.morphs}, rec, o);
            } else if (d.da_len > KP_ROOT_ID) {
                nh = 0;
                int prev = KP_ROOT_ID;
                int base = d.da[KP_ROOT_ID].x;
                uint32_t nch = 0;
                for (uint32_t i = bp; i < send; i++) {
                    uint32_t c = text[i];
                    nch += !is_cont(c);
                    int q = base + (int)c;                                   // da.rs:160
                    if (WORK) probes++;
                    if ((uint32_t)q >= d.da_len) break;                      // Vec::get -> None (da.rs:161)
                    int2 nq = d.da[q];
                    if (nq.y != prev) break;                                 // da.rs:162-164
                    if (WORK) probes_ok++;
                    int ahead = nq.x;                                        // + TERMINATOR (0), da.rs:165
                    if ((uint32_t)ahead < d.da_len) {
                        int2 na = d.da[ahead];
                        if (na.y == q && na.x < 0) {                         // da.rs:167-174
                            uint32_t id = (uint32_t)(-na.x);
                            const uint32_t k = (uint32_t)d.dup[id] + 1;      // index.rs:46-51
                            total += k;
                            atomicAdd(&bcount[b + nch], k);
                            const uint32_t y = nch | ((k - 1) << 16);
                            if (nh == 0) { h01.x = id; h01.y = y; }
                            else if (nh == 1) { h01.z = id; h01.w = y; }
                            else if (nh == 2) { h23.x = id; h23.y = y; }
                            else if (nh == 3) { h23.z = id; h23.w = y; }
                            nh++;
                        }
                    }
                    prev = q;
                    base = nq.x;
                }
                nhit[b] = (uint8_t)min(nh, 255u);
                if (nh > 0) hits[2 * (size_t)b] = h01;
                if (nh > 2) hits[2 * (size_t)b + 1] = h23;
            } else {
                nhit[b] = 0;
            }
            const bool matched = nh > 0;
            // unknown words (lattice.rs:42-99)
            const kp_catinfo ci = d.catinfo[bi.w & 0xFFu];
            if ((!matched || (ci.flags & 1u)) && ci.unk_count) {
                if (!FILL) {
                    total += ci.unk_count;
                    atomicAdd(&bcount[bi.z], ci.unk_count);
                    atomicAdd(&ucount[bi.z], ci.unk_count);
                } else {
                    uint32_t ulen = bi.z - b;
                    for (uint32_t j = 0; j < ci.unk_count; j++)             // lattice.rs:195
                        rec[o++] = kp_node_rec((uint32_t)(ci.unk_first + j), KP_CLASS_UNKNOWN, b,
                                               ld_morph(d.unk_morphs, ci.unk_first + j - 1), ulen);
                }
            }
        }
        if (!FILL) ncount[b] = total;
    }
    if (WORK) {
        probes = __reduce_add_sync(KP_FULL, probes);
        probes_ok = __reduce_add_sync(KP_FULL, probes_ok);
        if (lane_id() == 0) {
            atomicAdd((unsigned long long*)&totals[4], (unsigned long long)probes);
            atomicAdd((unsigned long long*)&totals[5], (unsigned long long)probes_ok);
        }
    }
}

// =================================================================================================
// Counting walk, the production version (kp_lattice_walk<false, *> above is the plain statement of
// the same walk; it still serves the work counters).  One thread per start boundary:
//   * the first character is one lookup in the first-character table (kp_dict.cu);
//   * per further byte, the terminator probe da[base[q]] and the next transition da[base[q] + byte]
//     are issued together (both depend only on da[q], da.rs:160-174) and the text byte is fetched a
//     step ahead; a failed range test clears `alive` instead of loading; the probe is skipped in the
//     middle of a character when every key of the dictionary ends on a character boundary;
//   * the first LAT_HITS hits {id, chars} go to a per-thread column of shared memory (no register
//     rotation in the loop); their duplicate counts are looked up after the walk, off its chain.
// ~7 of 32 lanes are live in the loop (walk lengths are heavy-tailed) and the kernel waits on its L2
// gathers: what pays is fewer gathers per walk and small blocks (a block lives as long as its
// longest walk).
// =================================================================================================
__global__ void __launch_bounds__(LAT_THREADS, KP_CNT_MINB) kp_lattice_count(
    const uint8_t* __restrict__ text, const uint4* __restrict__ binfo, uint32_t NB, kp_ddict d,
    uint32_t* __restrict__ ncount, uint32_t* __restrict__ bcount, uint32_t* __restrict__ ucount,
    uint8_t* __restrict__ nhit, uint4* __restrict__ hits) {
    __shared__ uint2 sh_hit[LAT_HITS][LAT_THREADS];          // [slot][thread]: conflict-free columns
    const uint32_t b = blockIdx.x * LAT_THREADS + threadIdx.x;
    if (b >= NB) return;
    const uint4 bi = binfo[b];
    const uint32_t bp = bi.x, send = bi.y;
    if (bp == send) {                            // EOS: one Dummy node (lattice.rs:165-175)
        ncount[b] = 1;
        return;
    }
    uint32_t nh = 0, total = 0;
    if (d.da_len > KP_ROOT_ID) {
        uint32_t i = bp;                         // index of the byte whose transition led to state q
        const uint32_t c = text[i];
        int prev = KP_ROOT_ID, q = 0;
        int2 nq = make_int2(0, 0);
        bool alive = false;
        int2 f = make_int2(KP_FIRST_SLOW, 0);
        if (KP_WALK_T1 && c < 0xF0u) {
            // the text is valid UTF-8, so the continuation bytes are inside the sentence
            uint32_t cp = c, L = 1;
            if (c >= 0xE0u) { cp = ((c & 0x0Fu) << 12) | ((text[i + 1] & 0x3Fu) << 6) | (text[i + 2] & 0x3Fu); L = 3; }
            else if (c >= 0x80u) { cp = ((c & 0x1Fu) << 6) | (text[i + 1] & 0x3Fu); L = 2; }
            f = d.first[cp];
            if (f.x != KP_FIRST_SLOW) {          // arrive in state f.x as if by the character's last byte
                i += L - 1;
                q = f.x;
                alive = f.x >= 0;
                nq = make_int2(f.y, 0);
                prev = 0;
            }
        }
        if (f.x == KP_FIRST_SLOW) {
            q = d.da[KP_ROOT_ID].x + (int)c;                                 // da.rs:160
            alive = (uint32_t)q < d.da_len;                                  // Vec::get -> None (da.rs:161)
            if (alive) nq = d.da[q];
        }
        // chars among the bytes consumed so far, the one being tried included: exact whenever the
        // transition held, which is the only time it is read
        uint32_t nch = 1;
        int c1 = i + 1 < send ? (int)(int8_t)text[i + 1] : 0;               // signed: continuation bytes are < -64
        while (alive && nq.y == prev) {                                      // da.rs:162-164
            const int ahead = nq.x;                                          // + TERMINATOR (0), da.rs:165
            const int q2 = nq.x + (c1 & 0xFF);
            const uint32_t i2 = i + 1;
            // a key can only end where a character ends (kp_dict.cu checks the dictionary): no probe while
            // the next byte is a continuation byte
            const bool pa = (uint32_t)ahead < d.da_len && (!KP_WALK_SKIP_MID || c1 >= -64 || d.mid_char_keys);
            const bool p2 = i2 < send && (uint32_t)q2 < d.da_len;
            int2 na = make_int2(0, 0), nq2 = make_int2(0, 0);
            if (pa) na = d.da[ahead];
            if (p2) nq2 = d.da[q2];
            int c2 = 0;
            if (i2 + 1 < send) c2 = (int)(int8_t)text[i2 + 1];
            if (pa && na.y == q && na.x < 0) {                               // da.rs:167-174
                const uint32_t id = (uint32_t)(-na.x);
                if (nh < LAT_HITS) {
                    sh_hit[nh][threadIdx.x] = make_uint2(id, nch);
                } else {
                    const uint32_t k = (uint32_t)d.dup[id] + 1;              // index.rs:46-51
                    total += k;
                    atomicAdd(&bcount[b + nch], k);
                }
                nh++;
            }
            nch += c1 >= -64;                    // the byte tried next starts a character
            prev = q;
            q = q2;
            nq = nq2;
            alive = p2;
            c1 = c2;
            i = i2;
        }
        // duplicate counts of the remembered hits: independent loads, off the walk's chain
        uint2 h[LAT_HITS];
        uint32_t k[LAT_HITS];
#pragma unroll
        for (uint32_t u = 0; u < LAT_HITS; u++) {
            h[u] = make_uint2(0u, 0u);
            if (u < nh) h[u] = sh_hit[u][threadIdx.x];
        }
#pragma unroll
        for (uint32_t u = 0; u < LAT_HITS; u++) k[u] = u < nh ? (uint32_t)d.dup[h[u].x] + 1 : 0u;
#pragma unroll
        for (uint32_t u = 0; u < LAT_HITS; u++) {
            if (u < nh) {
                total += k[u];
                atomicAdd(&bcount[b + h[u].y], k[u]);
                h[u].y |= (k[u] - 1) << 16;
            }
        }
        if (nh > 0) hits[2 * (size_t)b] = make_uint4(h[0].x, h[0].y, h[1].x, h[1].y);
        if (nh > 2) hits[2 * (size_t)b + 1] = make_uint4(h[2].x, h[2].y, h[3].x, h[3].y);
    }
    nhit[b] = (uint8_t)min(nh, 255u);
    // unknown words (lattice.rs:42-99)
    const kp_catinfo ci = d.catinfo[bi.w & 0xFFu];
    if ((nh == 0 || (ci.flags & 1u)) && ci.unk_count) {
        total += ci.unk_count;
        atomicAdd(&bcount[bi.z], ci.unk_count);
        atomicAdd(&ucount[bi.z], ci.unk_count);
    }
    ncount[b] = total;
}

int kp_launch_lattice_count(const kp_chunk& c, const kp_ddict& d, bool count_work, cudaStream_t st) {
    if (c.NB == 0) return 0;
    uint32_t blocks = (c.NB + LAT_THREADS - 1) / LAT_THREADS;
    if (count_work)
        kp_lattice_walk<false, true><<<blocks, LAT_THREADS, 0, st>>>(c.text, c.binfo, c.NB, d, c.ncount, c.bcount, c.ucount,
                                                                c.nhit, c.hits, nullptr, nullptr, c.totals);
    else
        kp_lattice_count<<<blocks, LAT_THREADS, 0, st>>>(c.text, c.binfo, c.NB, d, c.ncount, c.bcount, c.ucount, c.nhit,
                                                         c.hits);
    return kp_launch_check("kp_lattice_walk<count>");
}

int kp_launch_lattice_fill(const kp_chunk& c, const kp_ddict& d, cudaStream_t st) {
    if (c.NB == 0) return 0;
    uint32_t blocks = (c.NB + LAT_THREADS - 1) / LAT_THREADS;
    kp_lattice_walk<true, false><<<blocks, LAT_THREADS, 0, st>>>(c.text, c.binfo, c.NB, d, nullptr, nullptr, nullptr, c.nhit,
                                                            c.hits, c.noff, c.rec, c.totals);
    return kp_launch_check("kp_lattice_walk<fill>");
}

// =================================================================================================
// Column order of the sweep's connection matrix.  kp_viterbi reads connT[right_j][left_i]: the lanes
// of a group share the row and differ in the column, so a gather costs one L1 wavefront per distinct
// 128-byte line among the group's left ids.  Ranking the left ids by how often they occur in this
// batch's lattice and storing the columns in rank order puts the ids that matter into the first line
// or two of every row (1.8 -> 1.2 lines per group on the synthetic corpora).  The ranking only moves
// data around: every lookup returns the same matrix cell.  It is refreshed on a tokenizer's first
// pass and every KP_PERM_REFRESH passes after that (id frequencies drift slowly).
// =================================================================================================
constexpr uint32_t PERM_SAMPLE = 1u << 15;

__global__ void __launch_bounds__(256) kp_left_hist(uint32_t n, uint32_t stride, const uint4* __restrict__ rec,
                                                    uint32_t* __restrict__ hist) {
    uint32_t i = blockIdx.x * 256 + threadIdx.x;
    uint32_t left = i < n ? (rec[(size_t)i * stride].z & 0xFFFFu) : 0xFFFFFFFFu;
    uint32_t m = __match_any_sync(KP_FULL, left);
    if (i < n && lane_id() == (uint32_t)__ffs(m) - 1) atomicAdd(&hist[left], (uint32_t)__popc(m));
}

// perm[left] = rank of the left id by (count desc, id asc)
__global__ void __launch_bounds__(128) kp_left_rank(const uint32_t* __restrict__ hist, uint32_t n_left,
                                                    uint16_t* __restrict__ perm) {
    const uint32_t id = blockIdx.x * 128 + threadIdx.x;
    if (id >= n_left) return;
    const uint32_t c = hist[id];
    uint32_t rank = 0;
    for (uint32_t o = 0; o < n_left; o++) {
        const uint32_t co = hist[o];
        rank += (co > c) || (co == c && o < id);
    }
    perm[id] = (uint16_t)rank;
}

__global__ void __launch_bounds__(256) kp_conn_permute(const int16_t* __restrict__ connT, uint32_t stride,
                                                       uint32_t n_left, const uint16_t* __restrict__ perm,
                                                       int16_t* __restrict__ connP) {
    const size_t row = (size_t)blockIdx.x * stride;
    for (uint32_t l = threadIdx.x; l < n_left; l += 256) connP[row + perm[l]] = connT[row + l];
}

int kp_launch_column_order(const kp_chunk& c, const kp_ddict& d, const kp_perm& pm, cudaStream_t st) {
    const uint32_t sample = c.N < PERM_SAMPLE ? c.N : PERM_SAMPLE, stride = sample ? c.N / sample : 1;
    cudaMemsetAsync(pm.hist, 0, sizeof(uint32_t) * d.conn_col, st);
    if (sample) kp_left_hist<<<(sample + 255) / 256, 256, 0, st>>>(sample, stride, c.rec, pm.hist);
    kp_left_rank<<<(d.conn_col + 127) / 128, 128, 0, st>>>(pm.hist, d.conn_col, pm.perm);
    kp_conn_permute<<<d.conn_row, 256, 0, st>>>(d.connT, d.connT_stride, d.conn_col, pm.perm, pm.connP);
    int rc = kp_launch_check("kp_column_order");
    return rc < 0 ? rc : (sample ? 3 : 2);
}

// =================================================================================================
// Bucketize.  One warp per sentence walks its nodes in order, 32 at a time, and builds
//   * bnode: the reference's edges[end] lists (stable: ascending node index), used by the back-trace;
//   * red / rbk / tgt: the REDUCED buckets the Viterbi sweep scans (layout in kp_kernels.cuh).
// Why reduced: every start position inside a same-class run emits the class's unknown nodes, and
// they all end at the run's end -- buckets of 100+ entries that differ only in dp.  A successor only
// ever needs min(dp) per distinct (right_id); for unknown nodes the distinct ids are known
// statically (the ids of the class of the char before the boundary), so they get one shared slot
// each and the sweep min-merges into it.  The minimum VALUE is unchanged, and the back-trace picks
// the first predecessor attaining it from the full list, so results are bit-identical.
// =================================================================================================
#ifndef KP_SENT_THREADS
#define KP_SENT_THREADS 64
#endif
constexpr int SENT_THREADS = KP_SENT_THREADS;   // one warp per sentence
#ifndef KP_BK_SMEM
#define KP_BK_SMEM 256
#endif
#ifndef KP_BK_MINB
#define KP_BK_MINB 32          // 32 registers, 2 KB of cursors per warp: 64 resident warps per SM
#endif
constexpr uint32_t BK_SMEM = KP_BK_SMEM;   // boundaries per sentence whose fill cursors fit in shared memory

__global__ void __launch_bounds__(SENT_THREADS, KP_BK_MINB) kp_bucketize(uint32_t S, const uint32_t* __restrict__ coff,
                                                             const uint32_t* __restrict__ noff,
                                                             const uint32_t* __restrict__ boff,
                                                             const uint32_t* __restrict__ bcount,
                                                             const uint32_t* __restrict__ ucount,
                                                             const uint4* __restrict__ binfo,
                                                             const uint4* __restrict__ rec, kp_ddict d,
                                                             const uint16_t* __restrict__ perm,
                                                             uint2* __restrict__ bfill, uint2* __restrict__ rbk,
                                                             uint2* __restrict__ tgt, int2* __restrict__ red,
                                                             uint32_t* __restrict__ bnode) {
    uint32_t s = (blockIdx.x * SENT_THREADS + threadIdx.x) >> 5;
    if (s >= S) return;
    const uint32_t lane = lane_id();
    const uint32_t bb = coff[s] + s, n = coff[s + 1] - coff[s];
    const uint32_t n0 = noff[bb], n1 = noff[bb + n];   // nodes before the EOS node (which is node n1)
    // reduced bucket sizes: known nodes (+BOS) + the unknown ids of the preceding char's class
    for (uint32_t p = lane; p <= n; p += 32) {
        const uint32_t b = bb + p;
        const uint32_t uc = ucount[b];
        uint32_t r = bcount[b] - uc;
        if (uc) r += d.catinfo[binfo[b - 1].w & 0xFFu].unk_count;   // uc > 0 implies p > 0
        rbk[b] = make_uint2(boff[b], r);
    }
    if (lane == 0) {
        uint32_t q = boff[bb];                          // BOS: dp None -> unwrap_or(0) (lattice.rs:127)
        red[q] = make_int2(0, 0);
        bnode[q] = KP_NONE;
        tgt[n1] = make_uint2((uint32_t)perm[0], KP_NONE);   // EOS: morph (0,0,0) (column of left id 0), ends nowhere
    }
    // per-bucket fill cursors {all, known}: in shared memory when the sentence is short enough
    __shared__ uint2 sfill_all[SENT_THREADS / 32][BK_SMEM];
    uint2* sfill = sfill_all[threadIdx.x >> 5];
    const bool insm = n + 1 <= BK_SMEM;
    if (insm) {
        for (uint32_t p = lane; p <= n; p += 32) sfill[p] = make_uint2(0u, 0u);
        __syncwarp();
    }
    // Software pipeline, two rounds deep: the node records of round k+2 and the gathers that depend on
    // the records of round k+1 (bucket base, and for unknown nodes the known count and first id of
    // the class) are in flight while round k does its ranking and stores.
    auto gather = [&](const uint4& r, bool valid, uint32_t& base, uint32_t& uslot) {
        base = 0;
        uslot = 0;
        if (valid) {
            const uint32_t e = r.y + (r.w >> 16);
            base = boff[e];
            if ((r.x >> KP_KIND_SHIFT) != KP_CLASS_KNOWN) {
                const uint32_t cat = binfo[r.y].w & 0xFFu;      // class of the node's first char = of all its chars
                uslot = (bcount[e] - ucount[e]) + ((r.x & KP_ID_MASK) - (uint32_t)d.catinfo[cat].unk_first);
            }
        }
    };
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    uint4 rcur = n0 + lane < n1 ? rec[n0 + lane] : zero4;
    uint4 rnext = n0 + 32 + lane < n1 ? rec[n0 + 32 + lane] : zero4;
    uint32_t gbase, guslot;
    gather(rcur, n0 + lane < n1, gbase, guslot);
    for (uint32_t i0 = n0; i0 < n1; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool valid = i < n1;
        const uint4 r = rcur;
        const uint32_t base = gbase, uslot = guslot;
        rcur = rnext;
        gather(rcur, i + 32 < n1, gbase, guslot);
        rnext = i + 64 < n1 ? rec[i + 64] : zero4;
        const uint32_t e = valid ? r.y + (r.w >> 16) : KP_NONE;   // end boundary = start + char_len (lattice.rs:187,200)
        const bool known = valid && (r.x >> KP_KIND_SHIFT) == KP_CLASS_KNOWN;
        const uint32_t m = __match_any_sync(KP_FULL, e);
        const uint32_t km = m & __ballot_sync(KP_FULL, known);
        const uint32_t leader = (uint32_t)__ffs(m) - 1;
        uint2 old = make_uint2(0, 0);
        if (valid && lane == leader) {
            uint2* cur = insm ? &sfill[e - bb] : &bfill[e];
            old = *cur;
            *cur = make_uint2(old.x + (uint32_t)__popc(m), old.y + (uint32_t)__popc(km));
        }
        old.x = __shfl_sync(KP_FULL, old.x, leader);
        old.y = __shfl_sync(KP_FULL, old.y, leader);
        if (valid) {
            const uint32_t first = e == bb ? 1u : 0u;       // BOS occupies slot 0 of the first bucket
            bnode[base + first + old.x + (uint32_t)__popc(m & lanemask_lt())] = i;
            // known: next free known slot; unknown: shared slot of (end boundary, unknown id)
            const uint32_t slot = known ? base + first + old.y + (uint32_t)__popc(km & lanemask_lt()) : base + uslot;
            red[slot] = make_int2(KP_INF, (int)((r.z >> 16) * d.connT_stride));       // {dp, element offset of row right_id in connT}
            tgt[i] = make_uint2((uint32_t)perm[r.z & 0xFFFFu] | (r.w << 16),           // left id as its column in connP
                                known ? slot : slot | KP_SLOT_SHARED);
        }
        __syncwarp();
    }
}

int kp_launch_bucketize(const kp_chunk& c, const kp_ddict& d, const kp_perm& pm, cudaStream_t st) {
    if (c.S == 0) return 0;
    uint32_t blocks = (uint32_t)(((uint64_t)c.S * 32 + SENT_THREADS - 1) / SENT_THREADS);
    kp_bucketize<<<blocks, SENT_THREADS, 0, st>>>(c.S, c.coff, c.noff, c.boff, c.bcount, c.ucount, c.binfo, c.rec, d,
                                                  pm.perm, c.bfill, c.rbk, c.tgt, c.red, c.bnode);
    return kp_launch_check("kp_bucketize");
}

// =================================================================================================
// Viterbi work order
// =================================================================================================
#ifndef KP_VIT_GROUP
#define KP_VIT_GROUP 0         // lanes per sentence in the sweep; 0 = chosen per batch (kp_launch_viterbi)
#endif
#ifndef KP_VIT_MINB
#define KP_VIT_MINB 10
#endif
#ifndef KP_VIT_LATEPF
#define KP_VIT_LATEPF 1         // next boundary's bounds / targets issued just before the pair loop (see kp_viterbi)
#endif
#ifndef KP_VIT_TGEARLY
#define KP_VIT_TGEARLY 1        // next boundary's first targets fetched a whole step ahead
#endif
#ifndef KP_VIT_UNROLL
#define KP_VIT_UNROLL 4
#endif
constexpr int VIT_UNROLL = KP_VIT_UNROLL;     // pairs per batch of the inner loop
#ifndef KP_VIT_THREADS
#define KP_VIT_THREADS 128
#endif
constexpr int VIT_THREADS = KP_VIT_THREADS;

// Sentences sorted by length, longest first (counting sort on min(chars, LEN_BINS-1)): the four
// sentences a warp steps together then have (nearly) the same number of boundaries, and the long
// ones start first.  Order among equal lengths is arbitrary; results do not depend on it.
__global__ void __launch_bounds__(1024) kp_len_scan(uint32_t* __restrict__ hist) {   // in place: bin -> first slot
    __shared__ uint32_t wsum[32];
    // 4 bins per thread, bins visited from the longest to the shortest
    uint32_t v[4], tsum = 0;
    for (int k = 0; k < 4; k++) {
        v[k] = hist[LEN_BINS - 1 - (threadIdx.x * 4 + k)];
        tsum += v[k];
    }
    uint32_t x = tsum;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(KP_FULL, x, o);
        if (lane_id() >= (uint32_t)o) x += y;
    }
    if (lane_id() == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = wsum[threadIdx.x];
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(KP_FULL, w, o);
            if (lane_id() >= (uint32_t)o) w += y;
        }
        wsum[threadIdx.x] = w;
    }
    __syncthreads();
    uint32_t off = x - tsum + ((threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0u);
    for (int k = 0; k < 4; k++) {
        hist[LEN_BINS - 1 - (threadIdx.x * 4 + k)] = off;
        off += v[k];
    }
}

uint32_t kp_len_bins() { return LEN_BINS; }

// Work order of the Viterbi sweep: kp_prep_count has filled the length histogram; this scans it into
// first slots, and kp_prep_fill scatters the sentences through them.
int kp_launch_length_order(const kp_chunk& c, cudaStream_t st) {
    if (c.S == 0) return 0;
    kp_len_scan<<<1, 1024, 0, st>>>(c.lenhist);
    return kp_launch_check("kp_len_scan");
}

// &base[i] of an int16 array, computed once and kept: `volatile` stops the compiler from recomputing
// it from its inputs at every use inside the pair loop
__device__ __forceinline__ uint64_t elem_ptr_pinned(const int16_t* base, uint32_t i) {
    uint64_t r;
    asm volatile("mad.wide.u32 %0, %1, 2, %2;" : "=l"(r) : "r"(i), "l"(base));
    return r;
}
// *(int16*)(p + 2 * off).  ptxas emits LEA + LEA.HI.X for the address (it keeps the two halves of p in unpaired
// registers, which rules out IMAD.WIDE; a multiplier it cannot fold makes it three instructions -- both measured)
__device__ __forceinline__ int ld_conn_at(uint64_t p, uint32_t off) {
    uint64_t a;
    int v;
    asm("mad.wide.u32 %0, %1, 2, %2;" : "=l"(a) : "r"(off), "l"(p));
    asm("ld.global.nc.s16 %0, [%1];" : "=r"(v) : "l"(a));
    return v;
}
__device__ __forceinline__ int ld_conn(const char* p) {
    int v;
    asm("ld.global.nc.s16 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// =================================================================================================
// Viterbi forward sweep (lattice.rs:116-143, dp values only).
//
// A warp carries 32 / GROUP sentences (taken in length order), GROUP lanes each, and steps all
// of them boundary by boundary with warp-uniform loop bounds, so the groups share every issued
// instruction.  Lanes hold the nodes STARTING at the boundary (targets); each lane folds the
// boundary's reduced bucket (predecessors) with one DPX add-min per pair:
//   best = min(best, dp[j] + conn(right_j, left_i))                                 connection.rs:12-14
//   dp[i] = min(best + cost_i, INF), kept only if < INF                             lattice.rs:127-139
// The loop bound is the warp's largest bucket; lanes without a target and groups whose own bucket is
// exhausted have their loads predicated off, so the memory traffic is each group's own.  Per pair
// the loop issues six instructions: predicate, 8-byte bucket entry at an immediate offset, row
// address (LEA + LEA.HI.X), 2-byte matrix cell, VIADDMNMX.
// The argmin (pre_nodes) is NOT tracked here: the back-trace recomputes it for the ~30 nodes per
// sentence that lie on the best path.  dp goes to ndp[i] and into the node's reduced slot: a plain
// store for a known node (the slot is its own), a min-merge for an unknown node (the slot is shared
// by every unknown node with that id ending there); the value to merge with is fetched before the
// pair loop so its latency hides behind it.  The __syncwarp orders those stores before the next
// boundary's loads.
// =================================================================================================
template <int GROUP>
__global__ void __launch_bounds__(VIT_THREADS, KP_VIT_MINB) kp_viterbi(
    uint32_t S, uint32_t N, const uint32_t* __restrict__ order, const uint32_t* __restrict__ sel,
    const uint32_t* __restrict__ coff, const uint32_t* __restrict__ noff,
    const uint2* __restrict__ rbk, const uint2* __restrict__ tgt, int2* red,
    int32_t* __restrict__ ndp, int32_t* __restrict__ eos_cost, const int16_t* __restrict__ connT) {
    const uint32_t slot = (blockIdx.x * VIT_THREADS + threadIdx.x) / GROUP;
    const uint32_t l = threadIdx.x & (GROUP - 1);
    const bool has = slot < S;
    uint32_t s = 0, bb = 0, n = 0;
    if (has) {
        s = order[slot];
        bb = coff[s] + s;
        n = coff[s + 1] - coff[s];
    }
    const uint32_t steps = __reduce_max_sync(KP_FULL, has ? n + 1 : 0u);
    uint32_t t0 = 0, t1n = 0;
    uint2 bkn = make_uint2(0u, 0u);               // {first entry, entries} of the next boundary's reduced bucket
    if (has) {
        t0 = noff[bb];
        t1n = noff[bb + 1];
        bkn = rbk[bb];
    }
    // connT[right_j][left_i] (connection.rs:12-14, transposed): the lanes of a group share the row of
    // predecessor j and differ in the column, and the left ids of the nodes starting at one boundary
    // cluster (noun / unknown-word ids are neighbours), so a group's gather touches ~2 lines, not ~5
    uint2 tgn = make_uint2(0u, KP_NONE);          // first target chunk of the next boundary, prefetched
    if (has && t0 + l < t1n) tgn = tgt[t0 + l];
    for (uint32_t p = 0; p < steps; p++) {
        const bool act = has && p <= n;
        const uint32_t t1 = act ? t1n : t0, R = act ? bkn.y : 0u;
        const int2* const rbase = red + bkn.x;
#if KP_VIT_TGEARLY
        uint2 tnx = make_uint2(0u, KP_NONE);      // next boundary's first targets, fetched a whole step ahead
#endif
        // Bounds (and first targets) of the next boundary.  Issued just before the pair loop rather than
        // here: ptxas puts these loads on the scoreboard of the merge-value load, and the first wait on
        // that scoreboard would otherwise come a few instructions later, exposing their whole latency.
        auto fetch_next = [&]() {
            if (has && p < n) {
                t1n = noff[bb + p + 2];
                bkn = rbk[bb + p + 1];
#if KP_VIT_TGEARLY
                // the address is known now, whether the lane has a target there only once t1n arrives:
                // fetch anyway (clamped to the array), decide at the end of the step
                tnx = tgt[min(t1 + l, N)];
#endif
            }
        };
#if !KP_VIT_LATEPF
        fetch_next();
#endif
        const uint32_t T = t1 - t0;
        const uint32_t Tmax = __reduce_max_sync(KP_FULL, T), Rmax = __reduce_max_sync(KP_FULL, R);
#if KP_VIT_LATEPF
        if (Tmax == 0) fetch_next();              // no node starts here in any of the warp's sentences
#endif
        for (uint32_t tc = 0; tc < Tmax; tc += GROUP) {
            const bool tv = tc + l < T;
            uint2 tg = tgn;
            if (tc) tg = tv ? tgt[t0 + tc + l] : make_uint2(0u, KP_NONE);
            const uint64_t crow = elem_ptr_pinned(connT, tg.x & 0xFFFFu);   // this target's column; entries carry row offsets
            // reduced slot: bit 31 marks a shared one (unknown node); KP_NONE = EOS, which ends nowhere
            const bool eos = tg.y == KP_NONE;
            int* const slotp = &red[tg.y & ~KP_SLOT_SHARED].x;
            int merged = KP_INF;
            if (tv && !eos && (tg.y & KP_SLOT_SHARED)) merged = *slotp;
#if KP_VIT_LATEPF
            if (tc == 0) fetch_next();
#endif
            int best = INT_MAX;
            const int2* rp = rbase;
            int rem = tv ? (int)R : 0;
            for (uint32_t jj = 0; jj < Rmax; jj += VIT_UNROLL) {
                // all bucket entries of the batch first, then all matrix cells: two load latencies per
                // batch, not two per pair
                int2 e[VIT_UNROLL];
                int cell[VIT_UNROLL];
#pragma unroll
                for (int u = 0; u < VIT_UNROLL; u++)
                    if (u < rem) e[u] = rp[u];
#pragma unroll
                for (int u = 0; u < VIT_UNROLL; u++)
                    if (u < rem) cell[u] = ld_conn_at(crow, (uint32_t)e[u].y);   // crow + 2 * offset
#pragma unroll
                for (int u = 0; u < VIT_UNROLL; u++)
                    if (u < rem) best = __viaddmin_s32(e[u].x, cell[u], best);
                rp += VIT_UNROLL;
                rem -= VIT_UNROLL;
            }
            if (tv) {
                int dp = KP_INF;
                if (R) dp = min(best + (int)(int16_t)(tg.x >> 16), KP_INF);
                ndp[t0 + tc + l] = dp;
                if (eos) eos_cost[sel ? sel[s] : s] = dp;
                else *slotp = min(merged, dp);
            }
        }
        // the next boundary's first targets: the address does not depend on this step's results
#if KP_VIT_TGEARLY
        tgn = (has && p < n && t1 + l < t1n) ? tnx : make_uint2(0u, KP_NONE);
#else
        tgn = (has && p < n && t1 + l < t1n) ? tgt[t1 + l] : make_uint2(0u, KP_NONE);
#endif
        __syncwarp();
        t0 = t1;
    }
}

// Lanes per sentence: 8 when the batch alone fills the machine with warps (148 SMs x 40 resident
// warps), more when it does not -- a batch of few, long sentences (BASELINE.json configs[3]) is a few
// thousand sequential chains, and what matters then is the length of each chain, not lane use.
// Measured (cfg2, sweep ms, 8 / 16 / 32 lanes): 65 536 sentences 0.508 / 0.602 / 0.992; 32 768: 0.317 / 0.316 /
// 0.508; 24 576: 0.307 / 0.258 / 0.391; 16 384: 0.287 / 0.199 / 0.264; 8 192: 0.255 / 0.178 / 0.162.
// Results do not depend on the choice.
int kp_launch_viterbi(const kp_chunk& c, const kp_ddict& d, const kp_perm& pm, cudaStream_t st) {
    if (c.S == 0) return 0;
    const int group = KP_VIT_GROUP ? KP_VIT_GROUP : (c.S >= 30000 ? 8 : c.S >= 12000 ? 16 : 32);
    const uint32_t blocks = (uint32_t)(((uint64_t)c.S * group + VIT_THREADS - 1) / VIT_THREADS);
#define KP_VIT_LAUNCH(G)                                                                                          \
    kp_viterbi<G><<<blocks, VIT_THREADS, 0, st>>>(c.S, c.N, c.order, c.sel, c.coff, c.noff, c.rbk, c.tgt, c.red, c.ndp, c.eos_cost, \
                                                  pm.connP)
    if (group == 8) KP_VIT_LAUNCH(8);
    else if (group == 16) KP_VIT_LAUNCH(16);
    else if (group == 4) KP_VIT_LAUNCH(4);
    else KP_VIT_LAUNCH(32);
#undef KP_VIT_LAUNCH
    return kp_launch_check("kp_viterbi");
}

// First predecessor of a node attaining its dp, scanned in the reference's list order
// (edges[start], ascending node index): pre_nodes[i] of lattice.rs:136-139.
__device__ __forceinline__ uint32_t kp_first_argmin(const uint32_t* __restrict__ bnode, const int32_t* __restrict__ ndp,
                                                    const uint4* __restrict__ rec, uint32_t q0, uint32_t q1,
                                                    const char* crow, int want) {
    for (uint32_t j = q0; j < q1; j++) {
        const uint32_t nd = bnode[j];
        int dpj = 0;
        uint32_t right = 0;                      // BOS: dp None -> 0, morph (0,0,0)
        if (nd != KP_NONE) {
            dpj = ndp[nd];
            right = rec[nd].z >> 16;
        }
        if (dpj + ld_conn(crow + right * 2u) == want) return j;
    }
    return KP_NONE;
}

// pre_nodes[i] for EVERY node, recomputed from the dp values.  Only the lattice dump needs it.
__global__ void __launch_bounds__(256) kp_fill_pre(uint32_t N, const uint4* __restrict__ rec,
                                                   const uint2* __restrict__ tgt, const int32_t* __restrict__ ndp,
                                                   const uint32_t* __restrict__ bnode,
                                                   const uint32_t* __restrict__ boff, uint32_t* __restrict__ pre,
                                                   const int16_t* __restrict__ conn, uint32_t conn_row) {
    uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const uint4 r = rec[i];
    const uint32_t b = r.y;
    const int dp = ndp[i];
    uint32_t pr = KP_NONE;
    if (dp < KP_INF)
        pr = kp_first_argmin(bnode, ndp, rec, boff[b], boff[b + 1],
                             (const char*)conn + (size_t)(r.z & 0xFFFFu) * conn_row * 2,
                             dp - (int)(int16_t)(r.w & 0xFFFFu));
    pre[i] = pr;
}

int kp_launch_fill_pre(const kp_chunk& c, const kp_ddict& d, cudaStream_t st) {
    if (c.N == 0) return 0;
    kp_fill_pre<<<(c.N + 255) / 256, 256, 0, st>>>(c.N, c.rec, c.tgt, c.ndp, c.bnode, c.boff, c.pre, d.conn, d.conn_row);
    return kp_launch_check("kp_fill_pre");
}

// E = sum over boundaries of (#targets x #predecessors): the pairs visited by lattice.rs:122-125
__global__ void __launch_bounds__(256) kp_pair_count(uint32_t NB, const uint32_t* __restrict__ noff,
                                                     const uint32_t* __restrict__ boff, uint64_t* __restrict__ totals) {
    uint32_t b = blockIdx.x * 256 + threadIdx.x;
    unsigned long long e = 0;
    if (b < NB) e = (unsigned long long)(noff[b + 1] - noff[b]) * (boff[b + 1] - boff[b]);
    for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(KP_FULL, e, o);
    if (lane_id() == 0 && e) atomicAdd((unsigned long long*)&totals[6], e);
}

int kp_launch_pair_count(const kp_chunk& c, cudaStream_t st) {
    if (c.NB == 0) return 0;
    kp_pair_count<<<(c.NB + 255) / 256, 256, 0, st>>>(c.NB, c.noff, c.boff, c.totals);
    return kp_launch_check("kp_pair_count");
}

// =================================================================================================
// Back-trace (lattice.rs:144-153) + Node -> Token (tokenizer.rs:22-43).  8 / 16 / 32 lanes per sentence.
// Pass 1 walks from the EOS node; the lanes test that many candidate predecessors at a time for
// dp[j] + conn + cost == dp[i] and a ballot picks the first (list order), which is the predecessor the
// reference's strict '<' update keeps.  The path (node indices, back to front) is parked in `path`;
// pass 2 (after the scan of the path lengths) writes the tokens front to back.
// =================================================================================================
#ifndef KP_BT_THREADS
#define KP_BT_THREADS 128
#endif
constexpr int BT_THREADS = KP_BT_THREADS;
#ifndef KP_BT_GROUP
#define KP_BT_GROUP 0          // lanes per sentence; 0 = chosen per batch (kp_launch_backtrace_count)
#endif

#ifndef KP_BT_ORDER
#define KP_BT_ORDER 1
#endif
template <int BT_GROUP>
__global__ void __launch_bounds__(BT_THREADS) kp_backtrace_find(uint32_t S, const uint32_t* __restrict__ order,
                                                                const uint32_t* __restrict__ sel,
                                                                const uint32_t* __restrict__ coff,
                                                                const uint32_t* __restrict__ noff,
                                                                const uint32_t* __restrict__ boff,
                                                                const uint4* __restrict__ rec,
                                                                const int32_t* __restrict__ ndp,
                                                                const uint32_t* __restrict__ bnode,
                                                                const int16_t* __restrict__ conn, uint32_t conn_row,
                                                                uint32_t* __restrict__ path,
                                                                uint32_t* __restrict__ tcount) {
    const uint32_t slot = (blockIdx.x * BT_THREADS + threadIdx.x) / BT_GROUP;
    if (slot >= S) return;                       // whole groups leave together
    // length order, longest first (the Viterbi work order): the groups of a warp walk paths of about
    // the same length, and the long paths start first
    const uint32_t s = KP_BT_ORDER ? order[slot] : slot;
    const uint32_t l = threadIdx.x & (BT_GROUP - 1), gshift = lane_id() & ~(uint32_t)(BT_GROUP - 1);
    constexpr uint32_t GM = BT_GROUP == 32 ? KP_FULL : ((1u << (BT_GROUP & 31)) - 1u);
    const uint32_t gmask = GM << gshift;
    const uint32_t bb = coff[s] + s, n = coff[s + 1] - coff[s];
    uint32_t cur = noff[bb + n];                 // `self.nodes.len() - 1`: the EOS node
    uint32_t cnt = 0;
    int dp = ndp[cur];
    uint4 r = rec[cur];                          // {id|class, start boundary, left|right<<16, cost|len<<16}
    // bucket of the current node's start boundary.  For every candidate the lanes also fetch ITS start bucket's
    // bounds, next to the matrix cell (both depend only on the candidate's record): the winner's bounds arrive by
    // shuffle and the next token starts one dependent load earlier (four levels per token -> three).
    uint32_t q0 = boff[r.y], q1 = boff[r.y + 1];
    while (true) {
        if (dp >= KP_INF) break;                 // pre_nodes[pos] is None: no total < INF was ever seen
        const char* crow = (const char*)conn + (size_t)(r.z & 0xFFFFu) * conn_row * 2;   // connection.rs:12-14
        const int want = dp - (int)(int16_t)(r.w & 0xFFFFu);
        bool found = false;
        uint32_t next = KP_NONE, q0n = 0, q1n = 0;
        for (uint32_t j0 = q0; j0 < q1; j0 += BT_GROUP) {
            const uint32_t j = j0 + l;
            bool ok = false;
            uint32_t nd = KP_NONE;
            int dpj = 0;                         // BOS: dp None -> 0, morph (0,0,0)
            uint4 rn = make_uint4(0, 0, 0, 0);
            uint32_t nq0 = 0, nq1 = 0;
            if (j < q1) {
                nd = bnode[j];
                if (nd != KP_NONE) {
                    dpj = ndp[nd];
                    rn = rec[nd];
                    nq0 = boff[rn.y];
                    nq1 = boff[rn.y + 1];
                }
                ok = dpj + ld_conn(crow + (rn.z >> 16) * 2u) == want;
            }
            const uint32_t m = (__ballot_sync(gmask, ok) >> gshift) & GM;
            if (m) {                             // first match in list order; its lane already holds the
                const uint32_t src = gshift + (uint32_t)__ffs(m) - 1;   // next node's record, dp and bucket bounds
                next = __shfl_sync(gmask, nd, src);
                dp = __shfl_sync(gmask, dpj, src);
                r.y = __shfl_sync(gmask, rn.y, src);
                r.z = __shfl_sync(gmask, rn.z, src);
                r.w = __shfl_sync(gmask, rn.w, src);
                q0n = __shfl_sync(gmask, nq0, src);
                q1n = __shfl_sync(gmask, nq1, src);
                found = true;
                break;
            }
        }
        if (!found) break;                       // unreachable for a consistent dp table
        if (l == 0) path[bb + cnt] = cur;        // path length <= n + 1 = boundaries of the sentence
        cnt++;
        if (next == KP_NONE) break;              // reached BOS, which has no predecessor and is not emitted
        cur = next;
        q0 = q0n;
        q1 = q1n;
    }
    if (l == 0) tcount[sel ? sel[s] : s] = cnt;
}

// Tokens front to back into the staging area: a whole warp per sentence (~30 tokens: one round), each
// lane one token.  Sentence `so` of the chunk stages its tokens at stage[(off[so] - base) + so ...):
// a path has at most chars + 1 <= bytes + 1 tokens, so the regions of different sentences never
// overlap and no scan is needed before the tokens exist (the fused kernel stages the same way).
#ifndef KP_EMIT_GROUP
#define KP_EMIT_GROUP 32
#endif
constexpr int EMIT_GROUP = KP_EMIT_GROUP;
__global__ void __launch_bounds__(BT_THREADS) kp_backtrace_stage(uint32_t S, const uint32_t* __restrict__ sel,
                                                                 const uint64_t* __restrict__ off, uint64_t base,
                                                                 const uint32_t* __restrict__ coff,
                                                                 const uint4* __restrict__ rec,
                                                                 const uint4* __restrict__ binfo,
                                                                 const uint32_t* __restrict__ path,
                                                                 const uint32_t* __restrict__ tcount,
                                                                 kp_token* __restrict__ stage) {
    const uint32_t s = (blockIdx.x * BT_THREADS + threadIdx.x) / EMIT_GROUP;
    const uint32_t l = threadIdx.x & (EMIT_GROUP - 1);
    if (s >= S) return;
    const uint32_t so = sel ? sel[s] : s;
    const uint32_t bb = coff[s] + s;
    const uint32_t sent_byte0 = binfo[bb].x;     // for n == 0 this is the EOS boundary: also the sentence start
    const uint32_t cnt = tcount[so];
    kp_token* const out = stage + ((off[so] - base) + so);
    for (uint32_t k = l; k < cnt; k += EMIT_GROUP) {
        const uint4 r = rec[path[bb + cnt - 1 - k]];
        const uint32_t kind = r.x >> KP_KIND_SHIFT;
        kp_token t;
        t.id = (int32_t)(r.x & KP_ID_MASK);
        t.position = binfo[r.y].x - sent_byte0;
        t.start = r.y - bb;
        t.char_len = kind == KP_CLASS_DUMMY ? 3 : (uint16_t)(r.w >> 16);   // "EOS".chars().count()
        t.cls = (uint8_t)kind;
        t.reserved = 0;
        out[k] = t;
    }
}

// Staged tokens -> the packed result, after the scan of the token counts: a warp per sentence.
// COMPACT = false: kp_token records (16 B) and 64-bit offsets; true: kp_token8 (8 B; see the header for
// how the host rebuilds position / start) and 32-bit offsets.
template <bool COMPACT>
__global__ void __launch_bounds__(BT_THREADS) kp_tokens_pack(uint32_t S, const uint64_t* __restrict__ off, uint64_t base,
                                                             const uint32_t* __restrict__ toff, uint64_t tok_base,
                                                             const kp_token* __restrict__ stage, void* __restrict__ tok_off_out,
                                                             void* __restrict__ tokens_out) {
    const uint32_t s = (blockIdx.x * BT_THREADS + threadIdx.x) >> 5;
    const uint32_t l = lane_id();
    if (s > S) return;
    if (l == 0) {
        if (COMPACT) ((uint32_t*)tok_off_out)[s] = (uint32_t)(tok_base + toff[s]);
        else ((uint64_t*)tok_off_out)[s] = tok_base + toff[s];
    }
    if (s == S) return;
    const uint32_t w0 = toff[s], cnt = toff[s + 1] - w0;
    const uint4* const in = (const uint4*)(stage + ((off[s] - base) + s));
    for (uint32_t k = l; k < cnt; k += 32) {
        const uint4 t = in[k];                   // {id, position, start, char_len | cls << 16}
        if (!COMPACT) {
            ((uint4*)tokens_out)[w0 + k] = t;
        } else {
            const uint32_t cls = (t.w >> 16) & 0xFFu;
            uint32_t lens;
            if (cls == KP_CLASS_DUMMY) lens = t.z;                                   // EOS: n_chars
            else lens = ((in[k + 1].y - t.y) & 0xFFFFu) | (t.w << 16);               // bytes up to the next token | chars
            ((uint2*)tokens_out)[w0 + k] = make_uint2((uint32_t)t.x | (cls << KP_KIND_SHIFT), lens);
        }
    }
}

// Pipeline-only chunks (no sentence went through the fused kernel) skip the staging area: the packed result is
// written straight from the parked paths, a warp per sentence, each lane one token.
template <bool COMPACT>
__global__ void __launch_bounds__(BT_THREADS) kp_tokens_emit(uint32_t S, const uint32_t* __restrict__ coff,
                                                             const uint4* __restrict__ rec, const uint4* __restrict__ binfo,
                                                             const uint32_t* __restrict__ path,
                                                             const uint32_t* __restrict__ toff, uint64_t tok_base,
                                                             void* __restrict__ tok_off_out, void* __restrict__ tokens_out) {
    const uint32_t s = (blockIdx.x * BT_THREADS + threadIdx.x) >> 5;
    const uint32_t l = lane_id();
    if (s > S) return;
    if (l == 0) {
        if (COMPACT) ((uint32_t*)tok_off_out)[s] = (uint32_t)(tok_base + toff[s]);
        else ((uint64_t*)tok_off_out)[s] = tok_base + toff[s];
    }
    if (s == S) return;
    const uint32_t bb = coff[s] + s;
    const uint32_t sent_byte0 = binfo[bb].x;     // for n == 0 this is the EOS boundary: also the sentence start
    const uint32_t w0 = toff[s], cnt = toff[s + 1] - w0;
    for (uint32_t k = l; k < cnt; k += 32) {
        const uint4 r = rec[path[bb + cnt - 1 - k]];
        const uint32_t kind = r.x >> KP_KIND_SHIFT;
        const uint32_t position = binfo[r.y].x - sent_byte0, start = r.y - bb;
        if (!COMPACT) {
            // {id, position, start, char_len | cls << 16}; EOS: char_len = "EOS".chars().count()
            ((uint4*)tokens_out)[w0 + k] = make_uint4(r.x & KP_ID_MASK, position, start,
                                                      (kind == KP_CLASS_DUMMY ? 3u : r.w >> 16) | (kind << 16));
        } else {
            uint32_t lens = start;                                               // EOS: n_chars
            if (kind != KP_CLASS_DUMMY) {
                const uint32_t len = r.w >> 16;                                  // chars; the next token starts len chars on
                lens = ((binfo[r.y + len].x - binfo[r.y].x) & 0xFFFFu) | (len << 16);
            }
            ((uint2*)tokens_out)[w0 + k] = make_uint2((r.x & KP_ID_MASK) | (kind << KP_KIND_SHIFT), lens);
        }
    }
}

int kp_launch_tokens_emit(const kp_chunk& c, uint64_t tok_base, bool compact, cudaStream_t st) {
    const uint32_t blocks = (uint32_t)(((uint64_t)(c.S_all + 1) * 32 + BT_THREADS - 1) / BT_THREADS);
    if (compact)
        kp_tokens_emit<true><<<blocks, BT_THREADS, 0, st>>>(c.S_all, c.coff, c.rec, c.binfo, c.path, c.toff32, tok_base, c.tok_off, c.tokens);
    else
        kp_tokens_emit<false><<<blocks, BT_THREADS, 0, st>>>(c.S_all, c.coff, c.rec, c.binfo, c.path, c.toff32, tok_base, c.tok_off, c.tokens);
    return kp_launch_check("kp_tokens_emit");
}

// Lanes per sentence: the widest group that keeps the batch within ~2.6 waves of resident warps (148 SMs x 64 warps).
// A walk is a chain of dependent loads per token, and a bucket larger than the group takes several rounds of that
// chain: wide groups shorten every chain, narrow groups keep more chains resident.  Measured (cfg2, back-trace ms,
// 8 / 16 / 32 lanes): 65 536 sentences 0.195 / 0.227 / 0.272; 32 768: 0.156 / 0.123 / 0.151; 16 384: 0.141 / 0.111 /
// 0.086; cfg4 (4096 long lines): 3.08 / 2.13 / 1.51.  Results do not depend on the choice.
int kp_launch_backtrace_count(const kp_chunk& c, const kp_ddict& d, cudaStream_t st) {
    if (c.S == 0) return 0;
    const int group = KP_BT_GROUP ? KP_BT_GROUP : (c.S <= 25000 ? 32 : c.S <= 50000 ? 16 : 8);
    const uint32_t blocks = (uint32_t)(((uint64_t)c.S * group + BT_THREADS - 1) / BT_THREADS);
#define KP_BT_LAUNCH(G)                                                                                                    \
    kp_backtrace_find<G><<<blocks, BT_THREADS, 0, st>>>(c.S, c.order, c.sel, c.coff, c.noff, c.boff, c.rec, c.ndp, c.bnode, \
                                                        d.conn, d.conn_row, c.path, c.tcount)
    if (group == 8) KP_BT_LAUNCH(8);
    else if (group == 16) KP_BT_LAUNCH(16);
    else if (group == 4) KP_BT_LAUNCH(4);
    else KP_BT_LAUNCH(32);
#undef KP_BT_LAUNCH
    return kp_launch_check("kp_backtrace_find");
}

int kp_launch_backtrace_stage(const kp_chunk& c, cudaStream_t st) {
    if (c.S == 0) return 0;
    kp_backtrace_stage<<<(uint32_t)(((uint64_t)c.S * EMIT_GROUP + BT_THREADS - 1) / BT_THREADS), BT_THREADS, 0, st>>>(
        c.S, c.sel, c.off, c.base, c.coff, c.rec, c.binfo, c.path, c.tcount, c.stage);
    return kp_launch_check("kp_backtrace_stage");
}

// over ALL sentences of the chunk (S_all), whichever path staged their tokens
int kp_launch_tokens_pack(const kp_chunk& c, uint64_t tok_base, bool compact, cudaStream_t st) {
    const uint32_t blocks = (uint32_t)(((uint64_t)(c.S_all + 1) * 32 + BT_THREADS - 1) / BT_THREADS);
    if (compact)
        kp_tokens_pack<true><<<blocks, BT_THREADS, 0, st>>>(c.S_all, c.off, c.base, c.toff32, tok_base, c.stage, c.tok_off, c.tokens);
    else
        kp_tokens_pack<false><<<blocks, BT_THREADS, 0, st>>>(c.S_all, c.off, c.base, c.toff32, tok_base, c.stage, c.tok_off, c.tokens);
    return kp_launch_check("kp_tokens_pack");
}

// =================================================================================================
// Single-query common-prefix search (parity tests of the reference's da.rs / index.rs vectors).
// =================================================================================================
__global__ void kp_common_prefix(kp_ddict d, const uint8_t* __restrict__ text, uint32_t len, int expand_dup,
                                 int64_t* __restrict__ ids, uint64_t* __restrict__ lens, uint32_t cap,
                                 uint32_t* __restrict__ n_out) {
    if (threadIdx.x || blockIdx.x) return;
    uint32_t n = 0;
    if (d.da_len > KP_ROOT_ID) {
        int prev = KP_ROOT_ID, base = d.da[KP_ROOT_ID].x;
        for (uint32_t i = 0; i < len; i++) {
            int q = base + (int)text[i];
            if ((uint32_t)q >= d.da_len) break;
            int2 nq = d.da[q];
            if (nq.y != prev) break;
            int ahead = nq.x;
            if ((uint32_t)ahead < d.da_len) {
                int2 na = d.da[ahead];
                if (na.y == q && na.x < 0) {
                    uint32_t id = (uint32_t)(-na.x);
                    uint32_t k = expand_dup ? (uint32_t)d.dup[id] + 1 : 1;
                    for (uint32_t j = 0; j < k; j++) {
                        if (n < cap) {
                            ids[n] = (int64_t)id + j;
                            lens[n] = (uint64_t)i + 1;
                        }
                        n++;
                    }
                }
            }
            prev = q;
            base = nq.x;
        }
    }
    *n_out = n;
}

int kp_launch_common_prefix(const kp_ddict& d, const uint8_t* d_text, uint32_t len, int expand_dup, int64_t* d_ids,
                            uint64_t* d_lens, uint32_t cap, uint32_t* d_n, cudaStream_t st) {
    kp_common_prefix<<<1, 32, 0, st>>>(d, d_text, len, expand_dup, d_ids, d_lens, cap, d_n);
    return kp_launch_check("kp_common_prefix");
}

