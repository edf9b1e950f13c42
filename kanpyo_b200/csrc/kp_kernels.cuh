// kp_kernels.cuh — launch wrappers of the sm_100a kernels (definitions in kp_kernels.cu).
//
// Device data layout of one chunk (S sentences, B bytes, C chars, NB = C + S boundaries, N nodes):
//   text    u8 [B]        input bytes (chunk-relative)
//   off     u64[S+1]      caller's sentence offsets (absolute; `base` = offset of the chunk's first byte)
//   coff    u32[S+1]      exclusive scan of chars per sentence; boundary base of sentence s = coff[s] + s
//   binfo   uint4[NB]     per boundary {byte offset, sentence end byte, unknown-word end boundary, class}
//                         boundary p of a sentence = position before char p; p = n_chars is the EOS boundary
//   ncount/noff u32[NB+1] nodes STARTING at each boundary (count / exclusive scan) = reference insertion order
//   bcount/boff u32[NB+1] nodes ENDING at each boundary (+BOS at p=0)  = the reference's `edges` buckets
//   rec     uint4[N]      node {id|class<<30, start boundary, left|right<<16, cost|char_len<<16}
//   ucount  u32[NB]       UNKNOWN nodes ending at each boundary (subset of bcount)
//   bnode   u32[N]        the reference's `edges` lists: per bucket entry (ascending node index) the node
//                         index (KP_NONE for BOS); bucket of boundary b = [boff[b], boff[b+1])
//   nhit    u8[NB]        trie hits the counting walk saw at each start boundary (saturating)
//   hits    uint4[2 NB]   its first four, {id, chars | duplicates << 16} each: the fill pass replays them
//   red     int2[N]       REDUCED buckets, what the Viterbi sweep scans as predecessors: {min dp, element offset of
//                         row right_id in the transposed connection matrix}.
//                         Region of boundary b = rbk[b].x .. + rbk[b].y: one entry per known node
//                         ending at b (BOS first at a sentence's first boundary), then one entry per
//                         unknown-morph id of the class of the char before b, shared by ALL unknown nodes
//                         with that id ending at b (they have the same right_id; only their minimum dp
//                         can matter to a successor)
//   rbk     uint2[NB]     {boff[b], entries in the reduced bucket} of each boundary (one load per step of the sweep)
//   tgt     uint2[N]      per node, what the sweep needs of a TARGET: {column of left_id | cost<<16, reduced slot}
//                         (global index into red, | KP_SLOT_SHARED for an unknown node's shared slot;
//                         KP_NONE for EOS)
//   ndp     i32[N]        dp of every node (the back-trace and the lattice dump read it)
//   path    u32[NB]       best path of each sentence, back to front, at the sentence's boundary base
//   stage   kp_token[B+S+1] tokens of sentence s in path order at (off[s] - base) + s (a path has at most
//                         bytes + 1 tokens); kp_tokens_pack moves them to the packed result after the
//                         scan of the token counts
//   pre     u32[N]        (lattice dump only) predecessor as a bucket slot (KP_NONE = Option::None)
#pragma once
#include "kp_common.cuh"

struct kp_chunk {
    // inputs
    const uint8_t* text;     // device, chunk-relative base
    const uint64_t* off;     // device, [S+1]
    uint64_t base;           // off[0]
    uint32_t S, B;           // S = sentences of THIS pass (all of the chunk, or the `sel` subset)
    uint32_t S_all;          // sentences of the chunk
    bool direct_emit;        // pipeline-only chunk: tokens go from the parked paths to the packed result, no staging
    const uint32_t* sel;     // [S] chunk sentence behind each slot of this pass (nullptr = identity); outputs
                             // (eos_cost, tcount, staged tokens) are indexed by chunk sentence
    uint32_t* sel_out;       // [S_all] list the fused path appends the sentences it leaves to the pipeline to
    // sizes learnt on the way
    uint32_t C, NB, N;
    // scratch (device)
    uint32_t* nchar;  uint32_t* coff;
    uint4* binfo;
    uint32_t* ncount; uint32_t* noff; uint32_t* bcount; uint32_t* boff; uint32_t* ucount; uint2* rbk;
    uint8_t* nhit; uint4* hits;   // per start boundary: trie hits seen by the counting walk, and the first 4 {id, chars}
    uint2* bfill;            // running {all, known} counts while bucketizing
    uint4* rec; uint2* tgt; int2* red; int32_t* ndp; uint32_t* bnode; uint32_t* path; uint32_t* pre;
    int32_t* eos_cost; uint32_t* tcount; uint32_t* toff32;
    kp_token* stage;         // [B + S_all + 1] staged tokens: sentence s at (off[s] - base) + s, in path order
    void* tok_off;           // [S_all+1] output (rebased by tok_base): u64, or u32 for the compact form
    void* tokens;            // output: kp_token[], or kp_token8[] for the compact form
    uint64_t* scan_tmp;      // tile partials for the scans
    uint64_t* totals;        // [8] device scalars: 0 chars, 1 nodes, 2 bucket entries, 3 tokens, 4 P, 5 P_ok, 6 E
    uint32_t* err;           // [2] device flags: 0 utf8, 1 offsets
    uint32_t* lenhist;       // [kp_len_bins()] counting-sort bins of sentence lengths
    uint32_t* order;         // [S] sentences sorted by length, longest first (Viterbi work order)
};
uint32_t kp_len_bins();

// Column order of the Viterbi sweep's connection matrix (see kp_kernels.cu)
struct kp_perm {
    uint32_t* hist;          // [conn_col] left-id histogram of a node sample
    uint16_t* perm;          // [conn_col] left id -> column
    int16_t* connP;          // [conn_row * connT_stride] transposed matrix with permuted columns
};

uint32_t kp_scan_tmp_elems(uint32_t n);   // uint64 elements of scan_tmp needed for an n-element scan

// each returns the number of kernels launched (negative kp_status on launch failure)
int kp_launch_prep_count(const kp_chunk& c, cudaStream_t st);
int kp_launch_prep_fill(const kp_chunk& c, const kp_ddict& d, cudaStream_t st);
int kp_launch_lattice_count(const kp_chunk& c, const kp_ddict& d, bool count_work, cudaStream_t st);
int kp_launch_lattice_fill(const kp_chunk& c, const kp_ddict& d, cudaStream_t st);
int kp_launch_column_order(const kp_chunk& c, const kp_ddict& d, const kp_perm& pm, cudaStream_t st);
int kp_launch_bucketize(const kp_chunk& c, const kp_ddict& d, const kp_perm& pm, cudaStream_t st);
int kp_launch_length_order(const kp_chunk& c, cudaStream_t st);
int kp_launch_viterbi(const kp_chunk& c, const kp_ddict& d, const kp_perm& pm, cudaStream_t st);
int kp_launch_pair_count(const kp_chunk& c, cudaStream_t st);
int kp_launch_backtrace_count(const kp_chunk& c, const kp_ddict& d, cudaStream_t st);
int kp_launch_fill_pre(const kp_chunk& c, const kp_ddict& d, cudaStream_t st);
int kp_launch_backtrace_stage(const kp_chunk& c, cudaStream_t st);
int kp_launch_tokens_pack(const kp_chunk& c, uint64_t tok_base, bool compact, cudaStream_t st);
// pipeline-only chunk: packed result straight from the parked paths (no staging)
int kp_launch_tokens_emit(const kp_chunk& c, uint64_t tok_base, bool compact, cudaStream_t st);
// exclusive scans, n inputs -> n+1 outputs; total (u64) written to *total
int kp_launch_scan(const uint32_t* in, uint32_t* out, uint32_t n, uint64_t* tmp, uint64_t* total, cudaStream_t st);
int kp_launch_scan2(const uint32_t* in_a, const uint32_t* in_b, uint32_t* out_a, uint32_t* out_b, uint32_t n,
                    uint64_t* tmp, uint64_t* total_a, uint64_t* total_b, cudaStream_t st);
int kp_launch_common_prefix(const kp_ddict& d, const uint8_t* d_text, uint32_t len, int expand_dup, int64_t* d_ids,
                            uint64_t* d_lens, uint32_t cap, uint32_t* d_n, cudaStream_t st);

// ---- fused per-sentence path (kp_fused.cu) --------------------------------------------------------------
// A class = the sentences of at most max_bytes bytes, run by one-warp blocks with room for cap_c chars,
// cap_k known nodes (EOS included; cap_k / 2 trie hits) and cap_r reduced bucket slots in shared memory.
constexpr uint32_t KP_FUSED_MAX_CLASSES = 6;
struct kp_fused_class {
    uint32_t max_bytes, cap_c, cap_k, cap_r;
    uint32_t smem, blocks;   // filled by kp_fused_prepare: dynamic shared memory per block, persistent blocks to launch
};
struct kp_fused_classes {
    uint32_t n;
    kp_fused_class c[KP_FUSED_MAX_CLASSES];
};
uint32_t kp_fused_smem_bytes(const kp_fused_class& k);
bool kp_fused_dict_ok(const kp_ddict& d, const kp_catinfo* host_catinfo);
int kp_fused_prepare(kp_fused_classes* cls, int device);
// lists: [cls.n][S_all] sentence lists per class; counts: [cls.n]; nsel: sentences left to the pipeline (c.sel_out)
int kp_launch_fused_classify(const kp_chunk& c, const kp_fused_classes& cls, uint32_t* lists, uint32_t* counts,
                             uint32_t* nsel, cudaStream_t st);
// over_list / over_count: where the sentences that do not fit class k are appended
int kp_launch_fused(const kp_chunk& c, const kp_ddict& d, const kp_fused_class& k, const uint32_t* list,
                    const uint32_t* count, uint32_t* cursor, uint32_t* over_list, uint32_t* over_count, uint32_t expected,
                    cudaStream_t st);
