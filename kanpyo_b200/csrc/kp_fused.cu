// kp_fused.cu — the whole path of one sentence in one warp, on chip.
//
// The multi-kernel pipeline (kp_kernels.cu) moves ~250 bytes of scratch per input byte through HBM and
// every one of its kernels waits on dependent gathers into that scratch.  A sentence of the sizes real
// text has (BASELINE.json configs[1]: 82 chars, ~270 dictionary nodes, ~30 tokens) fits in a few KB, so
// here ONE WARP owns a sentence from its bytes to its tokens and keeps the lattice in shared memory:
//
//   decode    UTF-8 check, char offsets, classes, unknown-run ends            lattice.rs:105-111, 55-84
//   walk      lane = start position: double-array walk, hits {id, start, chars} da.rs:155-182
//   count     duplicates per hit, nodes per start / per end, scans            index.rs:40-53
//   expand    node {left, cost, id} at its start, {right} at its end bucket    lattice.rs:156-201
//   sweep     boundary by boundary: lanes = (target x predecessor slice),      lattice.rs:116-143
//             DPX add-min per pair, xor-shuffle min across the slices          connection.rs:12-14
//   trace     from EOS: first predecessor attaining dp (list order)           lattice.rs:136-153
//   tokens    staged at the sentence's slot of the staging area               tokenizer.rs:22-43
//
// Only the dictionary (L2-resident) and the sentence's own bytes are read from global memory; only the
// tokens, the token count and dp[EOS] are written.  Sentences come in size classes (by bytes), each
// launch with its own capacities; a sentence that does not fit its class (chars, nodes, hits) is appended
// to the largest class's list, which runs last, and what does not fit that one is taken by the pipeline
// afterwards: results are identical either way (tests/test_gpu_parity.py runs both).
//
// Where it is used: one line per call and small batches (one launch, one host round trip), where the
// pipeline's dozen launches dominate.  For tens of thousands of sentences the pipeline is ~2x faster: it
// amortises a sweep step over four sentences per warp and keeps 160 sentences in flight per SM where
// shared memory holds 8-17 here (profiles/r02_fused_kernel.md has the ncu evidence).
//
// Exactness notes
//   * Unknown nodes are not materialised.  Every start inside a same-class run emits the class's unknown
//     ids and they all end at the run's end; a successor only ever needs min(dp) per id, so each (end, id)
//     has ONE slot, min-merged in start order with a strict '<' (the first minimal start is kept).  The
//     minimum VALUE a successor sees is unchanged (the pipeline's reduced buckets do the same).
//   * The reference keeps the FIRST predecessor attaining the minimum in `edges[p]` order, which is
//     insertion order: ascending start, known before unknown, ascending id (two nodes of one bucket with
//     the same start have the same length, hence come from the same trie hit or the same class).  Bucket
//     slots are handed out by atomics here, so the back-trace compares the key (start, slot) instead of
//     the slot alone: a hit's duplicates take consecutive slots in id order, the unknown ids' shared slots
//     sit behind the known ones in id order, and a shared slot remembers its first minimal start --
//     ascending (start, slot) is the reference's order among the entries that can tie.
//   * dp arithmetic is the reference's: dp = min(min_j(dp_j + conn) + cost, INF), BOS = 0, a node without
//     predecessor (or with dead ones only) keeps INF and cuts the path.
#include <limits.h>

#include "kp_kernels.cuh"

#define KP_FULL 0xFFFFFFFFu

namespace {

constexpr uint32_t NONE16 = 0xFFFFu;
constexpr uint32_t ID_BITS = 20, ID_MASK20 = (1u << ID_BITS) - 1;   // ids the packed records can carry
constexpr uint32_t MAX_K = 2047;          // duplicates + 1 per hit
constexpr uint32_t MAX_NCH = 63;          // chars of a dictionary word
constexpr uint32_t EOS_INFO = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ bool is_cont(uint32_t c) { return (c & 0xC0u) == 0x80u; }
__device__ __forceinline__ uint32_t lead_len(uint32_t c) {  // 0 = not a valid lead byte
    if (c < 0x80u) return 1;
    if (c >= 0xC2u && c <= 0xDFu) return 2;
    if (c >= 0xE0u && c <= 0xEFu) return 3;
    if (c >= 0xF0u && c <= 0xF4u) return 4;
    return 0;
}
__device__ __forceinline__ uint2 ld_morph(const short4* __restrict__ m, uint32_t i) {   // {left | right << 16, cost}
    return __ldg((const uint2*)m + i);
}
}  // namespace

// Shared-memory layout of one warp's sentence (byte offsets).  C1 = cap_c + 2 boundary entries.
//   unk_lc  u32[64]   unknown morphs {left | cost << 16}, unk_ro u32[64] their row offsets right * stride
//                     (copied once per block; dictionaries with more unknown morphs are not eligible)
//   nstart  u32[C1]   known nodes starting at p: count -> first index -> (after expand) END index
//   bstart  u32[C1]   reduced-bucket size of boundary e -> first slot (bstart[n+1] = slots in use)
//   bcur    u32[C1]   known nodes ending at e (| bit 31: an unknown node ends here) -> fill cursor ->
//                     (after expand) first SHARED slot of bucket e
//   binfo   u32[C1]   per start: unknown end (10) | emits unknown << 10 | class unk_count (5) << 11 | unk_first (8) << 16
//                     | class invokes unknown words << 24
//   udesc   u32[C1]   per start: first shared slot of the bucket its unknown nodes end in | unk_first << 16
//   desc    uint2[C1] per boundary, what one step of the sweep needs in one load:
//                     {first target | known targets << 12 | targets << 22, first slot | slots << 12 | lane split << 24}
//   t_lc    u32[K]    known node: left | cost << 16                      (K = cap_k, index = node)
//   t_info  u32[K]    known node: id | chars << 20; EOS_INFO for the EOS node
//   t_slot  u16[K]    known node: its slot in the bucket of its end      } the path {id|kind<<30, start|chars<<16}
//   hy      u16[K/2]  hit: start (10) | chars (6) << 10                   } overlays these two after the sweep
//   k_dr    int2[R+1] bucket slot: {dp, row offset of right_id in connT}; slot R is where EOS (which ends nowhere)
//                     parks its dp.  hx u32[K/2] (hit id | k << 20) overlays it until the hits are expanded
//   k_node  u16[R+1]  bucket slot: node index (known), NONE16 (BOS), first minimal start (shared unknown slot)
//   bpos    u16[C1]   byte offset of char p inside the sentence (bpos[n] = bytes)
//   bcls    u8 [C1]   class of char p
//   tbytes  u8 [max_bytes + 8]   the sentence's bytes
struct kp_fused_layout {
    uint32_t unk_lc, unk_ro, nstart, bstart, bcur, binfo, udesc, desc, bpos, bcls, t_lc, t_info, t_slot, hy, k_dr, k_node,
        tbytes, total;
};

static __host__ __device__ inline kp_fused_layout kp_fused_make_layout(uint32_t cap_c, uint32_t cap_k, uint32_t cap_r,
                                                                       uint32_t max_bytes) {
    kp_fused_layout L;
    const uint32_t C1 = cap_c + 2;
    uint32_t o = 0;
    L.unk_lc = o; o += 4 * 64;
    L.unk_ro = o; o += 4 * 64;
    L.desc = o; o += 8 * C1;
    L.k_dr = o; o += 8 * (cap_r + 1);
    L.nstart = o; o += 4 * C1;
    L.bstart = o; o += 4 * C1;
    L.bcur = o; o += 4 * C1;
    L.binfo = o; o += 4 * C1;
    L.udesc = o; o += 4 * C1;
    L.t_lc = o; o += 4 * cap_k;
    L.t_info = o; o += 4 * cap_k;
    o = (o + 7) & ~7u;
    L.t_slot = o; o += 2 * cap_k;
    L.hy = o; o += 2 * (cap_k / 2);
    L.k_node = o; o += 2 * (cap_r + 1);
    L.bpos = o; o += 2 * C1;
    L.bcls = o; o += C1;
    L.tbytes = o; o += max_bytes + 8;
    L.total = (o + 15) & ~15u;
    return L;
}

uint32_t kp_fused_smem_bytes(const kp_fused_class& k) {
    return kp_fused_make_layout(k.cap_c, k.cap_k, k.cap_r, k.max_bytes).total;
}

// ---------------------------------------------------------------------------------------------------------
// Classification: one thread per sentence.  Sentences go to the list of the first class whose byte limit
// holds them, the rest (and everything when no class exists) to the pipeline's list.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kp_fused_classify(const uint64_t* __restrict__ off, uint64_t base, uint32_t S,
                                                         uint32_t B, kp_fused_classes cls, uint32_t* __restrict__ lists,
                                                         uint32_t* __restrict__ counts, uint32_t* __restrict__ sel,
                                                         uint32_t* __restrict__ nsel, uint32_t* __restrict__ err,
                                                         uint32_t* __restrict__ tcount) {
    const uint32_t s = blockIdx.x * 256 + threadIdx.x;
    if (s >= S) return;
    const uint64_t o0 = off[s], o1 = off[s + 1];
    if (o0 < base || o1 < o0 || o1 - base > B) {
        atomicOr(&err[1], 1u);
        tcount[s] = 0;
        return;
    }
    const uint64_t bytes = o1 - o0;
    for (uint32_t k = 0; k < cls.n; k++)
        if (bytes <= cls.c[k].max_bytes) {
            lists[(size_t)k * S + atomicAdd(&counts[k], 1u)] = s;
            return;
        }
    sel[atomicAdd(nsel, 1u)] = s;
}

// ---------------------------------------------------------------------------------------------------------
// The fused kernel: persistent one-warp blocks take sentences from their class's list by ticket.
// ---------------------------------------------------------------------------------------------------------
struct kp_fused_args {
    const uint8_t* text;
    const uint64_t* off;
    uint64_t base;
    const uint32_t* list;      // sentences of this class
    const uint32_t* count;     // how many
    uint32_t* cursor;          // ticket counter
    uint32_t cap_c, cap_k, cap_r, max_bytes;
    kp_token* stage;
    uint32_t* tcount;
    int32_t* eos_cost;
    uint32_t* sel;             // sentences that do not fit this class: appended here (the largest class's list,
    uint32_t* nsel;            // or, from the largest class, the pipeline's) ... and their count
    uint32_t* err;             // [0] invalid UTF-8
    unsigned long long* totals;   // [8] chars, [9] nodes (BOS and EOS included), [10] sentences done here
};

namespace {

constexpr int WALKS = 3;       // start positions a lane walks at the same time (their gathers overlap)

// One double-array walk in flight (da.rs:155-182).  `alive`: the transition into state q was in range and
// nq = da[q] is loaded; the walk goes on while check[q] == prev.
struct kp_walk {
    uint32_t p, i, nch, nh;
    int prev, q, c1;
    int2 nq;
    bool alive;
};

}  // namespace

__global__ void __launch_bounds__(32) kp_fused(const kp_fused_args a, const kp_ddict d) {
    extern __shared__ __align__(16) unsigned char smem[];
    const kp_fused_layout L = kp_fused_make_layout(a.cap_c, a.cap_k, a.cap_r, a.max_bytes);
    uint32_t* const unk_lc = (uint32_t*)(smem + L.unk_lc);
    uint32_t* const unk_ro = (uint32_t*)(smem + L.unk_ro);
    uint32_t* const nstart = (uint32_t*)(smem + L.nstart);
    uint32_t* const bstart = (uint32_t*)(smem + L.bstart);
    uint32_t* const bcur = (uint32_t*)(smem + L.bcur);
    uint32_t* const binfo = (uint32_t*)(smem + L.binfo);
    uint32_t* const udesc = (uint32_t*)(smem + L.udesc);
    uint2* const desc = (uint2*)(smem + L.desc);
    uint16_t* const bpos = (uint16_t*)(smem + L.bpos);
    uint8_t* const bcls = smem + L.bcls;
    uint32_t* const t_lc = (uint32_t*)(smem + L.t_lc);
    uint32_t* const t_info = (uint32_t*)(smem + L.t_info);
    uint16_t* const t_slot = (uint16_t*)(smem + L.t_slot);
    uint16_t* const hy = (uint16_t*)(smem + L.hy);
    int2* const k_dr = (int2*)(smem + L.k_dr);
    uint32_t* const hx = (uint32_t*)(smem + L.k_dr);
    uint16_t* const k_node = (uint16_t*)(smem + L.k_node);
    uint2* const path = (uint2*)(smem + L.t_slot);
    uint8_t* const tb = smem + L.tbytes;
    __shared__ uint32_t sh_hits;
    const uint32_t lane = lane_id();
    const uint32_t n_list = *a.count;
    const uint32_t cap_h = a.cap_k / 2;
    const int16_t* const connT = d.connT;
    const uint32_t stride = d.connT_stride;

    // the unknown morphs, once per block
    for (uint32_t u = lane; u < 64; u += 32) {
        uint2 m = make_uint2(0u, 0u);
        if (u < d.n_unk_morphs) m = ld_morph(d.unk_morphs, u);
        unk_lc[u] = (m.x & 0xFFFFu) | (m.y << 16);
        unk_ro[u] = (m.x >> 16) * stride;
    }
    uint32_t ticket = 0;
    if (lane == 0) ticket = atomicAdd(a.cursor, 1u);
    ticket = __shfl_sync(KP_FULL, ticket, 0);

    while (ticket < n_list) {
        const uint32_t s = a.list[ticket];
        if (lane == 0) ticket = atomicAdd(a.cursor, 1u);          // the next sentence's ticket, a whole sentence ahead
        const uint32_t lo = (uint32_t)(a.off[s] - a.base), hi = (uint32_t)(a.off[s + 1] - a.base);
        const uint32_t nbytes = hi - lo;
        bool give_up = false;                       // does not fit: the pipeline takes the sentence

        // ---- the sentence's bytes into shared memory (loads of a batch in flight together) ----
        {
            const uint8_t* const text = a.text + lo;
            for (uint32_t i0 = 0; i0 < nbytes; i0 += 256) {
                uint32_t c[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint32_t i = i0 + 32 * k + lane;
                    c[k] = i < nbytes ? text[i] : 0u;
                }
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint32_t i = i0 + 32 * k + lane;
                    if (i < nbytes) tb[i] = (uint8_t)c[k];
                }
            }
        }
        __syncwarp();

        // ---- decode: UTF-8 validation (the C ABI must check what &str guarantees), char offsets, classes ----
        uint32_t n = 0;
        {
            uint32_t conts = 0, claimed = 0;
            bool bad = false;
            for (uint32_t i0 = 0; i0 < nbytes; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool inr = i < nbytes;
                const uint32_t c = inr ? tb[i] : 0x80u;
                const bool st = inr && !is_cont(c);
                if (inr && !st) conts++;
                const uint32_t m = __ballot_sync(KP_FULL, st);
                if (st) {
                    const uint32_t len = lead_len(c);
                    if (len == 0 || i + len > nbytes) {
                        bad = true;
                    } else {
                        uint32_t cp = c;
                        if (len > 1) {
                            claimed += len - 1;
                            const uint32_t c1 = tb[i + 1];
                            uint32_t lo1 = 0x80u, hi1 = 0xBFu;
                            if (c == 0xE0u) lo1 = 0xA0u;
                            if (c == 0xEDu) hi1 = 0x9Fu;
                            if (c == 0xF0u) lo1 = 0x90u;
                            if (c == 0xF4u) hi1 = 0x8Fu;
                            if (c1 < lo1 || c1 > hi1) bad = true;
                            if (len == 2) cp = ((c & 0x1Fu) << 6) | (c1 & 0x3Fu);
                            else {
                                const uint32_t c2 = tb[i + 2];
                                if (!is_cont(c2)) bad = true;
                                if (len == 3) cp = ((c & 0x0Fu) << 12) | ((c1 & 0x3Fu) << 6) | (c2 & 0x3Fu);
                                else {
                                    const uint32_t c3 = tb[i + 3];
                                    if (!is_cont(c3)) bad = true;
                                    cp = ((c & 0x07u) << 18) | ((c1 & 0x3Fu) << 12) | ((c2 & 0x3Fu) << 6) | (c3 & 0x3Fu);
                                }
                            }
                        }
                        const uint32_t p = n + __popc(m & lanemask_lt());
                        if (p < a.cap_c) {
                            // CharCategoryDef::char_category: out-of-table code points use entry 0 (char_category_def.rs:33-38)
                            bpos[p] = (uint16_t)i;
                            bcls[p] = d.cat[cp < d.n_cat ? cp : 0];
                        }
                    }
                }
                n += __popc(m);
            }
            if (__reduce_add_sync(KP_FULL, conts) != __reduce_add_sync(KP_FULL, claimed)) bad = true;
            if (__any_sync(KP_FULL, bad)) {
                if (lane == 0) {
                    atomicOr(&a.err[0], 1u);
                    a.tcount[s] = 0;
                }
                ticket = __shfl_sync(KP_FULL, ticket, 0);
                continue;
            }
            if (n > a.cap_c) give_up = true;
        }
        if (!give_up) {
            if (lane == 0) {
                bpos[n] = (uint16_t)nbytes;
                sh_hits = 0;
            }
            for (uint32_t q = lane; q <= n + 1; q += 32) {
                nstart[q] = 0;
                bcur[q] = 0;
            }
            __syncwarp();
            // unknown-word extent of every start: to the end of its same-class run when the class groups,
            // at most 1024 chars, else one char (lattice.rs:55-84); the class's unknown ids ride along
            uint32_t carry = n;
            for (int32_t k = (int32_t)((n + 31) / 32) - 1; k >= 0; k--) {
                const uint32_t p = (uint32_t)k * 32 + lane;
                const bool valid = p < n;
                const uint32_t cat = valid ? bcls[p] : 0xFFFEu;
                const uint32_t catn = (p + 1 < n) ? bcls[p + 1] : 0xFFFDu;
                const uint32_t m = __ballot_sync(KP_FULL, valid && cat != catn);
                const uint32_t mge = m & ~lanemask_lt();
                const uint32_t runend = mge ? (uint32_t)k * 32 + (uint32_t)__ffs(mge) : carry;
                carry = __shfl_sync(KP_FULL, runend, 0);
                if (valid) {
                    const kp_catinfo ci = d.catinfo[cat];
                    const uint32_t uend = (ci.flags & 2u) ? min(runend, p + KP_MAX_UNKNOWN_LEN) : p + 1;
                    binfo[p] = uend | (ci.unk_count << 11) | (((uint32_t)ci.unk_first & 0xFFu) << 16) | ((ci.flags & 1u) << 24);
                }
            }
            if (lane == 0) {
                binfo[n] = 0;
                nstart[n] = 1;                      // the EOS node starts at boundary n (lattice.rs:165-175)
            }
            __syncwarp();

            // ---- walk: WALKS start positions per lane at a time (da.rs:155-182; the walk of kp_lattice_count) ----
            if (d.da_len > KP_ROOT_ID) {
                const int root_base = d.da[KP_ROOT_ID].x;
                for (uint32_t p0 = 0; p0 < n; p0 += 32 * WALKS) {
                    kp_walk w[WALKS];
#pragma unroll
                    for (int k = 0; k < WALKS; k++) {
                        kp_walk& x = w[k];
                        x.p = p0 + 32 * k + lane;
                        x.nh = 0;
                        x.alive = false;
                        x.nch = 1;
                        x.prev = KP_ROOT_ID;
                        x.q = 0;
                        x.nq = make_int2(0, 0);
                        x.c1 = 0;
                        x.i = 0;
                        if (x.p < n) {
                            uint32_t i = bpos[x.p];
                            const uint32_t c = tb[i];
                            int2 f = make_int2(KP_FIRST_SLOW, 0);
                            if (c < 0xF0u) {
                                uint32_t cp = c, len = 1;
                                if (c >= 0xE0u) { cp = ((c & 0x0Fu) << 12) | ((tb[i + 1] & 0x3Fu) << 6) | (tb[i + 2] & 0x3Fu); len = 3; }
                                else if (c >= 0x80u) { cp = ((c & 0x1Fu) << 6) | (tb[i + 1] & 0x3Fu); len = 2; }
                                f = d.first[cp];
                                if (f.x != KP_FIRST_SLOW) {          // arrive in state f.x as if by the character's last byte
                                    i += len - 1;
                                    x.q = f.x;
                                    x.alive = f.x >= 0;
                                    x.nq = make_int2(f.y, 0);
                                    x.prev = 0;
                                }
                            }
                            if (f.x == KP_FIRST_SLOW) {
                                x.q = root_base + (int)c;                                        // da.rs:160
                                x.alive = (uint32_t)x.q < d.da_len;                              // Vec::get -> None (da.rs:161)
                                if (x.alive) x.nq = d.da[x.q];
                            }
                            x.i = i;
                            x.c1 = i + 1 < nbytes ? (int)(int8_t)tb[i + 1] : 0;
                        }
                    }
                    while (true) {
                        bool any = false;
#pragma unroll
                        for (int k = 0; k < WALKS; k++) {
                            kp_walk& x = w[k];
                            x.alive = x.alive && x.nq.y == x.prev;                               // da.rs:162-164
                            any = any || x.alive;
                        }
                        if (!any) break;
                        int2 na[WALKS], nq2[WALKS];
                        bool pa[WALKS], p2[WALKS];
                        int q2[WALKS];
#pragma unroll
                        for (int k = 0; k < WALKS; k++) {                                        // all gathers of the step first
                            kp_walk& x = w[k];
                            const int ahead = x.nq.x;                                            // + TERMINATOR (0), da.rs:165
                            q2[k] = x.nq.x + (x.c1 & 0xFF);
                            pa[k] = x.alive && (uint32_t)ahead < d.da_len && (x.c1 >= -64 || d.mid_char_keys);
                            p2[k] = x.alive && x.i + 1 < nbytes && (uint32_t)q2[k] < d.da_len;
                            na[k] = make_int2(0, 0);
                            nq2[k] = make_int2(0, 0);
                            if (pa[k]) na[k] = d.da[ahead];
                            if (p2[k]) nq2[k] = d.da[q2[k]];
                        }
#pragma unroll
                        for (int k = 0; k < WALKS; k++) {
                            kp_walk& x = w[k];
                            if (x.alive) {
                                if (pa[k] && na[k].y == x.q && na[k].x < 0) {                    // da.rs:167-174
                                    const uint32_t id = (uint32_t)(-na[k].x);
                                    const uint32_t h = atomicAdd(&sh_hits, 1u);
                                    if (h < cap_h && x.nch <= MAX_NCH) {
                                        hx[h] = id;
                                        hy[h] = (uint16_t)(x.p | (x.nch << 10));
                                    } else {
                                        atomicOr(&sh_hits, 0x80000000u);                         // does not fit
                                    }
                                    x.nh++;
                                }
                                x.nch += x.c1 >= -64;        // the byte tried next starts a character
                                x.prev = x.q;
                                x.q = q2[k];
                                x.nq = nq2[k];
                                x.alive = p2[k];
                                x.i++;
                                x.c1 = x.i + 1 < nbytes ? (int)(int8_t)tb[x.i + 1] : 0;
                            }
                        }
                    }
                    // unknown words (lattice.rs:42-99): when nothing matched, or the class always invokes them
#pragma unroll
                    for (int k = 0; k < WALKS; k++) {
                        const kp_walk& x = w[k];
                        if (x.p < n) {
                            const uint32_t bi = binfo[x.p];
                            if ((x.nh == 0 || (bi >> 24)) && ((bi >> 11) & 31u)) {
                                binfo[x.p] = bi | (1u << 10);
                                atomicOr(&bcur[bi & 1023u], 0x80000000u);
                            }
                        }
                    }
                }
            } else {
                for (uint32_t p = lane; p < n; p += 32) {
                    const uint32_t bi = binfo[p];
                    if ((bi >> 11) & 31u) {
                        binfo[p] = bi | (1u << 10);
                        atomicOr(&bcur[bi & 1023u], 0x80000000u);
                    }
                }
            }
            __syncwarp();
            const uint32_t H = sh_hits;
            if (H & 0x80000000u) give_up = true;

            // ---- count: duplicates per hit (index.rs:46-51), known nodes per start and per end ----
            if (!give_up) {
                bool over = false;
                for (uint32_t h = lane; h < H; h += 32) {
                    const uint32_t id = hx[h], y = hy[h];
                    const uint32_t k = (uint32_t)d.dup[id] + 1;
                    if (k > MAX_K) over = true;
                    atomicAdd(&nstart[y & 1023u], k);
                    atomicAdd(&bcur[(y & 1023u) + (y >> 10)], k);
                    hx[h] = id | (k << ID_BITS);
                }
                if (__any_sync(KP_FULL, over)) give_up = true;
            }
        }
        __syncwarp();
        uint32_t n_known = 0, n_slots = 0;
        if (!give_up) {
            // exclusive scans over the boundaries 0..n: first node of every start, first slot of every bucket
            uint32_t ca = 0, cb = 0;
            bool wide = false;
            for (uint32_t q0 = 0; q0 <= n; q0 += 32) {
                const uint32_t q = q0 + lane;
                uint32_t va = 0, vb = 0;
                if (q <= n) {
                    va = nstart[q];
                    const uint32_t bc = bcur[q];
                    vb = (bc & 0x7FFFFFFFu) + (q == 0 ? 1u : 0u);                 // + BOS in edges[0] (lattice.rs:156-164)
                    if (bc >> 31) vb += (binfo[q - 1] >> 11) & 31u;               // shared slots: the ids of the class before q
                    if (va > 900u) wide = true;                                   // desc carries 10-bit target counts
                }
                uint32_t xa = va, xb = vb;
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t ya = __shfl_up_sync(KP_FULL, xa, o), yb = __shfl_up_sync(KP_FULL, xb, o);
                    if (lane >= (uint32_t)o) {
                        xa += ya;
                        xb += yb;
                    }
                }
                if (q <= n) {
                    nstart[q] = ca + xa - va;
                    bstart[q] = cb + xb - vb;
                    bcur[q] = cb + xb - vb + (q == 0 ? 1u : 0u);                  // fill cursor (BOS holds slot 0)
                }
                ca += __shfl_sync(KP_FULL, xa, 31);
                cb += __shfl_sync(KP_FULL, xb, 31);
            }
            n_known = ca;
            n_slots = cb;
            if (n_known > a.cap_k || n_slots > a.cap_r || __any_sync(KP_FULL, wide)) give_up = true;
        }
        if (give_up) {
            if (lane == 0) a.sel[atomicAdd(a.nsel, 1u)] = s;
            ticket = __shfl_sync(KP_FULL, ticket, 0);
            continue;
        }
        if (lane == 0) bstart[n + 1] = n_slots;
        __syncwarp();

        // ---- expand, pass 1: one lane per hit hands out node indices and bucket slots (lattice.rs:177-188) ----
        if (lane == 0) {
            const uint32_t i = nstart[n];
            nstart[n] = i + 1;
            t_info[i] = EOS_INFO;                   // EOS: morph (0,0,0); it ends nowhere: its dp parks in slot n_slots
            t_slot[i] = (uint16_t)n_slots;
        }
        {
            const uint32_t H = sh_hits;
            for (uint32_t h = lane; h < H; h += 32) {
                const uint32_t x = hx[h], y = hy[h];
                const uint32_t id = x & ID_MASK20, k = x >> ID_BITS, p = y & 1023u, nch = y >> 10;
                const uint32_t i0 = atomicAdd(&nstart[p], k);
                const uint32_t s0 = atomicAdd(&bcur[p + nch], k);
                for (uint32_t dd = 0; dd < k; dd++) {
                    t_info[i0 + dd] = (id + dd) | (nch << ID_BITS);
                    t_slot[i0 + dd] = (uint16_t)(s0 + dd);
                }
            }
        }
        __syncwarp();                               // the hits are dead from here on: k_dr takes their place
        // pass 2: one lane per node fetches its morph: {left, cost} at the node, {right} at its slot (two in flight)
        for (uint32_t i0 = 0; i0 < n_known; i0 += 64) {
            const uint32_t ia = i0 + lane, ib = i0 + 32 + lane;
            const uint32_t fa = ia < n_known ? t_info[ia] : EOS_INFO, fb = ib < n_known ? t_info[ib] : EOS_INFO;
            uint2 ma = make_uint2(0u, 0u), mb = make_uint2(0u, 0u);
            if (fa != EOS_INFO) ma = ld_morph(d.morphs, (fa & ID_MASK20) - 1);
            if (fb != EOS_INFO) mb = ld_morph(d.morphs, (fb & ID_MASK20) - 1);
            if (ia < n_known) {
                t_lc[ia] = (ma.x & 0xFFFFu) | (ma.y << 16);
                if (fa != EOS_INFO) {
                    const uint32_t sl = t_slot[ia];
                    k_dr[sl].y = (int)((ma.x >> 16) * stride);
                    k_node[sl] = (uint16_t)ia;
                }
            }
            if (ib < n_known) {
                t_lc[ib] = (mb.x & 0xFFFFu) | (mb.y << 16);
                if (fb != EOS_INFO) {
                    const uint32_t sl = t_slot[ib];
                    k_dr[sl].y = (int)((mb.x >> 16) * stride);
                    k_node[sl] = (uint16_t)ib;
                }
            }
        }
        // shared slots of the unknown ids: [bcur[e], bstart[e + 1]); BOS: dp None -> unwrap_or(0) (lattice.rs:127);
        // and the per-boundary descriptors of the sweep
        uint32_t n_unknown = 0;
        for (uint32_t e = lane; e <= n; e += 32) {
            const uint32_t q0 = bcur[e], q1 = bstart[e + 1];
            if (q1 > q0) {
                const uint32_t first = (binfo[e - 1] >> 16) & 0xFFu;
                for (uint32_t q = q0; q < q1; q++) {
                    k_dr[q] = make_int2(INT_MAX, (int)unk_ro[first + (q - q0) - 1]);
                    k_node[q] = 0;
                }
            }
            if (e == 0) {
                k_dr[0] = make_int2(0, 0);
                k_node[0] = (uint16_t)NONE16;
            }
            const uint32_t t0 = e ? nstart[e - 1] : 0u, Tk = nstart[e] - t0;
            const uint32_t bi = binfo[e];
            const uint32_t Tu = (bi >> 10) & 1u ? (bi >> 11) & 31u : 0u;
            const uint32_t T = Tk + Tu;
            n_unknown += Tu;
            uint32_t sh = T <= 1 ? 0u : 32u - (uint32_t)__clz(T - 1);
            if (sh > 5) sh = 5;
            desc[e] = make_uint2(t0 | (Tk << 12) | (T << 22), bstart[e] | ((q1 - bstart[e]) << 12) | (sh << 24));
            udesc[e] = bcur[bi & 1023u] | (bi & 0x00FF0000u);
        }
        n_unknown = __reduce_add_sync(KP_FULL, n_unknown);
        __syncwarp();

        // ---- sweep (lattice.rs:116-143): per boundary, lanes = (target, slice of the predecessors) ----
        // B0 stands for "no predecessor seen": every real dp_j + conn is below it, and B0 + cost >= INF for every cost
        constexpr int B0 = KP_INF + 32768;
        for (uint32_t p = 0; p <= n; p++) {
            const uint2 ds = desc[p];
            const uint32_t T = ds.x >> 22;
            if (T == 0) continue;
            const uint32_t t0 = ds.x & 0xFFFu, Tk = (ds.x >> 12) & 0x3FFu;
            const uint32_t R = (ds.y >> 12) & 0xFFFu, sh = ds.y >> 24;
            const int2* const bucket = k_dr + (ds.y & 0xFFFu);
            const uint32_t ud = udesc[p];
            const uint32_t il = lane & ((1u << sh) - 1u), jo = lane >> sh, J = 32u >> sh;
            for (uint32_t tc = 0; tc < T; tc += 1u << sh) {
                const uint32_t ti = tc + il;
                const bool tv = ti < T, unk = ti >= Tk;
                const uint32_t* const src = unk ? unk_lc + ((ud >> 16) + (ti - Tk) - 1) : t_lc + (t0 + ti);
                const uint32_t lc = tv ? *src : 0u;
                const int16_t* const col = connT + (lc & 0xFFFFu);
                int best = B0;
                const uint32_t Rl = tv ? R : 0u;
                for (uint32_t j = jo; j < Rl; j += J) {
                    const int2 e = bucket[j];
                    best = __viaddmin_s32(e.x, (int)__ldg(col + e.y), best);      // connection.rs:12-14
                }
                if (sh < 5) best = min(best, __shfl_xor_sync(KP_FULL, best, 16));
                if (sh < 4) best = min(best, __shfl_xor_sync(KP_FULL, best, 8));
                if (sh < 3) best = min(best, __shfl_xor_sync(KP_FULL, best, 4));
                if (sh < 2) best = min(best, __shfl_xor_sync(KP_FULL, best, 2));
                if (sh < 1) best = min(best, __shfl_xor_sync(KP_FULL, best, 1));
                if (tv && jo == 0) {
                    const int dp = min(best + (int)(int16_t)(lc >> 16), KP_INF);          // lattice.rs:127-139
                    const uint32_t dst = unk ? (ud & 0xFFFFu) + (ti - Tk) : t_slot[t0 + ti];
                    const int old = unk ? k_dr[dst].x : INT_MAX;
                    if (dp < old) {                   // unknown: strict, the first minimal start is kept
                        k_dr[dst].x = dp;
                        if (unk) k_node[dst] = (uint16_t)p;
                    }
                }
            }
            __syncwarp();
        }
        const int eos_dp = k_dr[n_slots].x;

        // ---- back-trace (lattice.rs:144-153): the first predecessor attaining dp, in `edges` order: ascending
        // start, known before unknown, ascending id = ascending (start, slot) ----
        uint32_t cnt = 0;
        {
            uint32_t cur_id = 0, cur_kind = KP_CLASS_DUMMY, cur_p = n, cur_len = 3, left = 0;
            int cost = 0, dpc = eos_dp;
            while (dpc < KP_INF) {                   // pre_nodes[pos] is None otherwise
                const int want = dpc - cost;
                const uint2 ds = desc[cur_p];
                const uint32_t r0 = ds.y & 0xFFFu, R = (ds.y >> 12) & 0xFFFu, rs = bcur[cur_p];
                const int16_t* const col = connT + left;
                uint32_t best_key = 0xFFFFFFFFu;
                for (uint32_t j0 = 0; j0 < R; j0 += 32) {
                    const uint32_t j = j0 + lane;
                    uint32_t key = 0xFFFFFFFFu;
                    if (j < R) {
                        const int2 e = k_dr[r0 + j];
                        if (e.x != INT_MAX && e.x + (int)__ldg(col + e.y) == want) {
                            const uint32_t nd = k_node[r0 + j];
                            uint32_t start = nd;                                  // shared slot: its first minimal start
                            if (r0 + j < rs) start = nd == NONE16 ? 0u : cur_p - (t_info[nd] >> ID_BITS);
                            key = (start << 12) | j;
                        }
                    }
                    best_key = min(best_key, __reduce_min_sync(KP_FULL, key));
                }
                if (best_key == 0xFFFFFFFFu) break;  // unreachable for a consistent dp table
                if (lane == 0) path[cnt] = make_uint2(cur_id | (cur_kind << 30), cur_p | (cur_len << 16));
                cnt++;
                const uint32_t bj = r0 + (best_key & 0xFFFu);
                const uint32_t nd = k_node[bj];
                if (bj < rs && nd == NONE16) break;  // BOS: no predecessor, not emitted
                dpc = k_dr[bj].x;
                if (bj >= rs) {                      // the class's unknown id (bj - rs), started at nd
                    cur_id = ((binfo[cur_p - 1] >> 16) & 0xFFu) + (bj - rs);
                    const uint32_t lc = unk_lc[cur_id - 1];
                    cur_kind = KP_CLASS_UNKNOWN;
                    cur_len = cur_p - nd;
                    cur_p = nd;
                    left = lc & 0xFFFFu;
                    cost = (int)(int16_t)(lc >> 16);
                } else {
                    const uint32_t inf = t_info[nd], lc = t_lc[nd];
                    cur_kind = KP_CLASS_KNOWN;
                    cur_id = inf & ID_MASK20;
                    cur_len = inf >> ID_BITS;
                    cur_p -= cur_len;
                    left = lc & 0xFFFFu;
                    cost = (int)(int16_t)(lc >> 16);
                }
            }
        }
        __syncwarp();

        // ---- tokens (tokenizer.rs:22-43), front to back, into the sentence's slot of the staging area ----
        if (lane == 0) {
            a.tcount[s] = cnt;
            a.eos_cost[s] = eos_dp;
            atomicAdd(&a.totals[8], (unsigned long long)n);
            atomicAdd(&a.totals[9], (unsigned long long)n_known + n_unknown + 1);   // + BOS
            atomicAdd(&a.totals[10], 1ull);
        }
        uint4* const out = (uint4*)(a.stage + ((size_t)lo + s));
        for (uint32_t k = lane; k < cnt; k += 32) {
            const uint2 e = path[cnt - 1 - k];
            const uint32_t kind = e.x >> 30, p = e.y & 0xFFFFu;
            // {id, position, start, char_len | cls << 16}; EOS: char_len = "EOS".chars().count()
            out[k] = make_uint4(e.x & KP_ID_MASK, bpos[p], p, (e.y >> 16) | (kind << 16));
        }
        ticket = __shfl_sync(KP_FULL, ticket, 0);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
bool kp_fused_dict_ok(const kp_ddict& d, const kp_catinfo* host_catinfo) {
    if (d.n_morphs >= (1u << ID_BITS) || d.n_unk_morphs > 64u || d.conn_col > 65536u || d.conn_row > 65536u) return false;
    for (int c = 0; c < 256; c++)
        if (host_catinfo[c].unk_count > 31u || (uint32_t)host_catinfo[c].unk_first + host_catinfo[c].unk_count > 65535u)
            return false;
    return true;
}

static int kp_fused_check(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        kp_set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return KP_ERR_CUDA;
    }
    return 1;
}

int kp_fused_prepare(kp_fused_classes* cls, int device) {
    // opt in to large dynamic shared memory once, and learn how many one-warp blocks of each class fit an SM
    int sms = 0, max_optin = 0;
    KP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    KP_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    uint32_t largest = 0;
    for (uint32_t k = 0; k < cls->n; k++) {
        cls->c[k].smem = kp_fused_smem_bytes(cls->c[k]);
        largest = cls->c[k].smem > largest ? cls->c[k].smem : largest;
    }
    if ((int)largest > max_optin) {
        kp_set_error("fused kernel class needs %u bytes of shared memory, the device offers %d", largest, max_optin);
        return KP_ERR_ARG;
    }
    KP_CUDA(cudaFuncSetAttribute(kp_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)largest));
    KP_CUDA(cudaFuncSetAttribute(kp_fused, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    for (uint32_t k = 0; k < cls->n; k++) {
        int per_sm = 0;
        KP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kp_fused, 32, cls->c[k].smem));
        if (per_sm < 1) per_sm = 1;
        cls->c[k].blocks = (uint32_t)(per_sm * sms);
    }
    return KP_OK;
}

int kp_launch_fused_classify(const kp_chunk& c, const kp_fused_classes& cls, uint32_t* lists, uint32_t* counts,
                             uint32_t* nsel, cudaStream_t st) {
    if (c.S_all == 0) return 0;
    kp_fused_classify<<<(c.S_all + 255) / 256, 256, 0, st>>>(c.off, c.base, c.S_all, c.B, cls, lists, counts, (uint32_t*)c.sel_out,
                                                              nsel, c.err, c.tcount);
    return kp_fused_check("kp_fused_classify");
}

int kp_launch_fused(const kp_chunk& c, const kp_ddict& d, const kp_fused_class& k, const uint32_t* list,
                    const uint32_t* count, uint32_t* cursor, uint32_t* over_list, uint32_t* over_count, uint32_t expected,
                    cudaStream_t st) {
    kp_fused_args a;
    a.text = c.text;
    a.off = c.off;
    a.base = c.base;
    a.list = list;
    a.count = count;
    a.cursor = cursor;
    a.cap_c = k.cap_c;
    a.cap_k = k.cap_k;
    a.cap_r = k.cap_r;
    a.max_bytes = k.max_bytes;
    a.stage = c.stage;
    a.tcount = c.tcount;
    a.eos_cost = c.eos_cost;
    a.sel = over_list;
    a.nsel = over_count;
    a.err = c.err;
    a.totals = (unsigned long long*)c.totals;
    uint32_t blocks = k.blocks;
    if (expected < blocks) blocks = expected ? expected : 1;
    kp_fused<<<blocks, 32, k.smem, st>>>(a, d);
    return kp_fused_check("kp_fused");
}
