/* kanpyo_b200.h — C ABI of the B200-native Kanpyo hot path (lattice build + Viterbi).
 *
 * This is the drop-in boundary for the reference's `Tokenizer::tokenize()` path
 * (togatoga/kanpyo @ f1931e2c).  Everything behind it is hand-written CUDA for sm_100a; there is
 * NO CPU fallback: every compute entry point returns KP_ERR_CUDA when no usable GPU is present.
 *
 * Plain C: pointers and sizes only, no C++/torch types.  All functions return 0 (KP_OK) or a
 * negative kp_status; nothing throws across the boundary.  A Rust/ctypes/cgo caller binds exactly
 * these symbols (see INTEGRATION.md for the Rust shim that keeps the reference's signatures).
 *
 * Reference interfaces replaced (file:line relative to the reference tree):
 *   kp_dict_create            <- the fields of `Dict` the hot path reads           kanpyo-dict/src/dict.rs:21-30
 *                                (IndexTable{da,dup} index.rs:10-13, Morphs morph.rs:24,
 *                                 ConnectionTable connection.rs:5-9, CharCategoryDef
 *                                 char_category_def.rs:15-20, UnkDict unk_dict.rs:12-16)
 *   kp_tokenizer_create       <- Tokenizer::new(dict)                               src/tokenizer.rs:12-14
 *   kp_tokenize               <- Tokenizer::tokenize(&self, &str) -> Vec<Token>     src/tokenizer.rs:16-45
 *   kp_tokenize_batch[_device]<- the same call over many independent sentences (new: the reference
 *                                has no batch entry; semantics = tokenize() per sentence)
 *   kp_token                  <- Token / TokenClass                                 src/token.rs:4-18
 *   kp_lattice_dump           <- Lattice::build + Lattice::viterbi (pub nodes/edges) src/lattice.rs:6-10,101,116
 *   kp_da_common_prefix       <- IndexTable::search_common_prefix_of                kanpyo-dict/src/index.rs:40-53
 *                                (DoubleArray::search_common_prefix_of              kanpyo-dict/src/trie/da.rs:155-182)
 * New with ABI 2 (no counterpart in the reference, which is single-threaded and takes one line per call):
 *   kp_tokenize_batch8 / kp_token8   the batch call with 8-byte transfer records (half the device-to-host bytes)
 *   kp_queue_*                       successive batches in flight: copies hidden behind the neighbours' kernels
 *   kp_shards_*                      every GPU of the box from one process (NCCL broadcast + sentence shards)
 *   kp_gather_*                      the token gather for one-process-per-GPU launchers (NCCL send / recv)
 * Threading: a kp_dict is immutable and shareable; a kp_tokenizer, kp_queue, kp_shards or kp_gather handle takes
 * calls from one thread at a time (each owns its streams, scratch and result buffers).
 */
#ifndef KANPYO_B200_H
#define KANPYO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KP_ABI_VERSION 2           /* 2: kp_profile grew fused_ms; compact results, queues and shards are new symbols */

typedef enum kp_status {
    KP_OK = 0,
    KP_ERR_ARG = -1,          /* null pointer / inconsistent sizes / offsets not ascending */
    KP_ERR_CUDA = -2,         /* CUDA runtime error or no sm_100 device (kp_last_error() has the text) */
    KP_ERR_DICT = -3,         /* dictionary arrays fail validation (an index the reference would panic on) */
    KP_ERR_UTF8 = -4,         /* input is not valid UTF-8 (Rust's &str guarantees validity; the ABI checks) */
    KP_ERR_NOMEM = -5,        /* host or device allocation failed */
    KP_ERR_TOO_LARGE = -6,    /* a single chunk exceeds the 2^31-byte / 2^31-node device index range */
    KP_ERR_BLOB = -7          /* packed dictionary blob has a bad magic / version / size */
} kp_status;

/* TokenClass (src/token.rs:4-8) */
enum { KP_CLASS_DUMMY = 0, KP_CLASS_KNOWN = 1, KP_CLASS_UNKNOWN = 2 };

/* Host arrays describing the read-only dictionary, exactly the data the reference's hot path reads.
 * Nothing is retained after kp_dict_create returns. */
typedef struct kp_dict_arrays {
    const int32_t* da;              /* [da_len][2] = {base, check}                      da.rs:14-20 */
    uint64_t da_len;
    const int64_t* dup_ids;         /* BTreeMap<KeywordID,usize> keys                   index.rs:12 */
    const uint64_t* dup_counts;     /*                           values                             */
    uint64_t n_dup;
    const int16_t* morphs;          /* [n_morphs][3] = {left_id, right_id, cost}        morph.rs:7-11,24 */
    uint64_t n_morphs;
    uint64_t conn_row, conn_col;    /* ConnectionTable{row,col}; get(r,c)=data[row*c+r] connection.rs:5-14 */
    const int16_t* conn;            /* [conn_row*conn_col]                                           */
    const uint8_t* char_category;   /* [n_char_category] code point -> class            char_category_def.rs:15-38 */
    uint64_t n_char_category;
    const uint8_t* invoke_list;     /* [n_invoke] bool                                  char_category_def.rs:18 */
    uint64_t n_invoke;
    const uint8_t* group_list;      /* [n_group] bool                                   char_category_def.rs:19 */
    uint64_t n_group;
    const uint8_t* unk_cat;         /* BTreeMap<u8,(KeywordID,usize)> keys              unk_dict.rs:15 */
    const int64_t* unk_first_id;    /*   .0 : first 1-based unknown morph id                          */
    const uint64_t* unk_count;      /*   .1 : number of consecutive ids                               */
    uint64_t n_unk_map;
    const int16_t* unk_morphs;      /* [n_unk_morphs][3]                                unk_dict.rs:13 */
    uint64_t n_unk_morphs;
} kp_dict_arrays;

typedef struct kp_dict kp_dict;             /* immutable, device-resident; shareable across tokenizers/threads */
typedef struct kp_tokenizer kp_tokenizer;   /* one CUDA stream + scratch; NOT thread-safe (one per caller thread) */

/* Token (src/token.rs:11-18) without the heap string.  `end` = start + char_len.  `surface` is "EOS"
 * for KP_CLASS_DUMMY (src/tokenizer.rs:28-31); otherwise it is the sentence's bytes
 * [position, next.position) where `next` is the following token of the same sentence: consecutive
 * path nodes are adjacent in the input (src/lattice.rs:124-125 pairs a node only with nodes ending
 * where it starts) and every non-empty path ends with the EOS token, whose position = sentence length. */
typedef struct kp_token {
    int32_t id;          /* Token.id (KeywordID); 0 for EOS                                      */
    uint32_t position;   /* Token.position: byte offset inside its sentence                      */
    uint32_t start;      /* Token.start: char offset inside its sentence                         */
    uint16_t char_len;   /* Token.end - Token.start (EOS: 3 = "EOS".chars().count())             */
    uint8_t cls;         /* KP_CLASS_*                                                           */
    uint8_t reserved;    /* 0                                                                    */
} kp_token;              /* 16 bytes */

/* Result of a batch call.  Pointers are owned by the tokenizer and stay valid until its next call
 * or kp_tokenizer_destroy.  For kp_tokenize_batch they are (pinned) HOST pointers; for
 * kp_tokenize_batch_device they are DEVICE pointers on the tokenizer's device. */
typedef struct kp_result {
    uint64_t n_sent;
    uint64_t n_tokens;
    const uint64_t* tok_off;    /* [n_sent+1]: tokens of sentence s are tokens[tok_off[s] .. tok_off[s+1]) */
    const kp_token* tokens;     /* [n_tokens]                                                              */
    const int32_t* eos_cost;    /* [n_sent]: dp[EOS] of Lattice::viterbi (src/lattice.rs:116-143)          */
} kp_result;

/* Compact token record for transfers: half the bytes of kp_token.  `position` and `start` are prefix
 * sums the host rebuilds (consecutive path nodes are adjacent in the input, see kp_token): walking a
 * sentence's tokens BACKWARDS from its EOS token, position -= byte_len and start -= char_len.  The EOS
 * token (cls == KP_CLASS_DUMMY, always the last token of a non-empty path) carries the sentence's char
 * count in its two length fields: n_chars = byte_len | char_len << 16 (Token.start of EOS,
 * src/tokenizer.rs:24-34); its position is the sentence's byte length, which the caller knows.
 * kp_expand_tokens8 does exactly this on the host. */
typedef struct kp_token8 {
    uint32_t id_cls;     /* Token.id | KP_CLASS_* << 30                                          */
    uint16_t byte_len;   /* surface bytes  (EOS: low half of n_chars)                            */
    uint16_t char_len;   /* Token.end - Token.start (EOS: high half of n_chars)                  */
} kp_token8;             /* 8 bytes */

typedef struct kp_result8 {
    uint64_t n_sent;
    uint64_t n_tokens;
    const uint32_t* tok_off;    /* [n_sent+1] */
    const kp_token8* tokens;    /* [n_tokens] */
    const int32_t* eos_cost;    /* [n_sent]   */
} kp_result8;

/* Exact work counters of the last batch call (used for the roofline's algorithmic bytes). */
typedef struct kp_counters {
    uint64_t bytes;      /* B: input bytes            */
    uint64_t chars;      /* C: input chars            */
    uint64_t nodes;      /* N: lattice nodes incl. BOS and EOS, as in Lattice.nodes.len() summed */
    uint64_t tokens;     /* T: emitted tokens         */
    uint64_t sentences;
    /* filled only when kp_tokenizer_set_count_work(t, 1) was called (extra counting kernels run): */
    uint64_t probes;     /* P: trie transitions attempted   (da.rs:160-162)                      */
    uint64_t probes_ok;  /* P_ok: transitions whose check matched (then read `ahead`, da.rs:166-167) */
    uint64_t pairs;      /* E: (target, previous) pairs visited by viterbi (lattice.rs:125)      */
} kp_counters;

/* Per-stage device times of the last batch call, CUDA events on the tokenizer's stream (milliseconds). */
typedef struct kp_profile {
    float h2d_ms, prep_ms, lattice_ms, bucket_ms, viterbi_ms, backtrace_ms, d2h_ms, total_ms;
    uint32_t kernel_launches;   /* kernels of this library launched by the call */
    uint32_t chunks;
    float fused_ms;             /* the fused per-sentence kernel (the other stage times then cover only the
                                   sentences that did not fit it) */
    uint32_t fused_sentences;   /* sentences the fused kernel completed */
} kp_profile;

int kp_abi_version(void);
const char* kp_strerror(int status);
const char* kp_last_error(void);            /* thread-local detail for the last failing call */
int kp_device_count(int* n);                /* CUDA devices visible to the library            */

/* ---- dictionary ------------------------------------------------------------------------------- */
/* Validates every index the hot path can form (KP_ERR_DICT instead of the reference's panic), packs
 * the arrays into one blob and stages it to HBM on `device` once. */
int kp_dict_create(const kp_dict_arrays* arrays, int device, kp_dict** out);
/* Validate + pack only (host work, no device needed).  dst == NULL: just report *size. */
int kp_dict_pack(const kp_dict_arrays* arrays, void* dst, uint64_t cap, uint64_t* size);
/* The packed blob (host copy kept by the handle / the staged copy in HBM): what a multi-GPU launcher
 * broadcasts to the other ranks. */
int kp_dict_blob(const kp_dict* d, const void** host_ptr, uint64_t* size);
int kp_dict_device_blob(const kp_dict* d, const void** device_ptr, uint64_t* size);
/* Build a handle from a packed blob in host memory, or from one already in device memory (e.g. the
 * receive buffer of an NCCL broadcast). */
int kp_dict_create_from_blob(const void* host_blob, uint64_t size, int device, kp_dict** out);
int kp_dict_create_from_device_blob(const void* device_blob, uint64_t size, int device, kp_dict** out);
void kp_dict_destroy(kp_dict* d);

/* ---- tokenizer -------------------------------------------------------------------------------- */
int kp_tokenizer_create(const kp_dict* d, kp_tokenizer** out);     /* Tokenizer::new */
void kp_tokenizer_destroy(kp_tokenizer* t);
/* Upper bound on input bytes processed per device pass by kp_tokenize_batch (default 64 MiB). */
int kp_tokenizer_set_chunk_bytes(kp_tokenizer* t, uint64_t bytes);

/* Tokenizer::tokenize for one sentence (host UTF-8, not NUL-terminated). */
int kp_tokenize(kp_tokenizer* t, const uint8_t* utf8, uint64_t len, kp_result* out);
/* Sentences s = utf8[offsets[s] .. offsets[s+1]) for s in [0,n_sent); HOST pointers in, HOST result out. */
int kp_tokenize_batch(kp_tokenizer* t, const uint8_t* utf8, const uint64_t* offsets, uint64_t n_sent,
                      kp_result* out);
/* Same, with utf8/offsets already resident in device memory and the result left in device memory.
 * n_bytes = offsets[n_sent] - offsets[0] (the host must know it).  The batch is one device pass. */
int kp_tokenize_batch_device(kp_tokenizer* t, const uint8_t* d_utf8, const uint64_t* d_offsets, uint64_t n_sent,
                             uint64_t first_offset, uint64_t n_bytes, kp_result* out);
/* kp_tokenize_batch with the compact 8-byte token records and 32-bit token offsets (half the
 * device-to-host bytes).  KP_ERR_TOO_LARGE if the batch has 2^32 tokens or more. */
int kp_tokenize_batch8(kp_tokenizer* t, const uint8_t* utf8, const uint64_t* offsets, uint64_t n_sent,
                       kp_result8* out);
/* kp_tokenize_batch_device with the compact result (device pointers). */
int kp_tokenize_batch_device8(kp_tokenizer* t, const uint8_t* d_utf8, const uint64_t* d_offsets, uint64_t n_sent,
                              uint64_t first_offset, uint64_t n_bytes, kp_result8* out);
/* Host-side expansion kp_token8 -> kp_token (fills position / start / char_len; EOS char_len = 3).
 * offsets = the caller's sentence offsets [n_sent+1]; out has r->n_tokens entries.  Pure host code; results of
 * more than 2^18 tokens are split over up to 8 host threads (contiguous sentence ranges). */
int kp_expand_tokens8(const kp_result8* r, const uint64_t* offsets, kp_token* out);
int kp_last_counters(const kp_tokenizer* t, kp_counters* out);
/* on != 0: also count P, P_ok, E (slower; never enabled inside a timed region). */
int kp_tokenizer_set_count_work(kp_tokenizer* t, int on);
int kp_last_profile(const kp_tokenizer* t, kp_profile* out);
/* Copy `bytes` from device memory of a kp_tokenize_batch_device result to host memory (for callers
 * that do not bind a CUDA runtime themselves). */
int kp_copy_to_host(kp_tokenizer* t, void* dst, const void* device_src, uint64_t bytes);
/* Blocks until all work queued by this tokenizer is complete. */
int kp_tokenizer_sync(kp_tokenizer* t);

/* Which device path kp_tokenize* takes.  Two exist, with identical results (tests run both):
 *   the pipeline  a dozen kernels over the whole batch with the lattice in HBM: the throughput path, it needs
 *                 tens of thousands of sentences in flight to hide its dependent gathers;
 *   the fused kernel  one warp owns one sentence from its bytes to its tokens in shared memory: one launch and,
 *                 for batches of at most 64 sentences / 48 KiB, ONE host round trip per call -- the path for
 *                 the reference's own call pattern (one line per call) and for small batches.
 * KP_PATH_AUTO (default) picks by batch size (fused up to 3584 sentences per chunk, where the two paths cross); sentences that do not fit
 * the fused kernel's shared-memory budget go through the pipeline in the same call.  KP_PATH_PIPELINE and
 * KP_PATH_FUSED force one path for every batch size (measurements, parity tests). */
enum { KP_PATH_AUTO = 0, KP_PATH_PIPELINE = 1, KP_PATH_FUSED = 2 };
int kp_tokenizer_set_path(kp_tokenizer* t, int path);
/* How the calling thread waits for the device.  0 (default): the driver spins (lowest latency).  1: the thread
 * sleeps on a blocking-sync event, leaving the core to others -- for hosts with fewer cores than waiting contexts.
 * (Measured with 24 queue workers on a 32-core host: spinning 50.0 GB/s, sleeping 41.7 GB/s; hence the default.) */
int kp_tokenizer_set_blocking_sync(kp_tokenizer* t, int on);

/* ---- asynchronous batches: H2D(n+1) || kernels(n) || D2H(n-1) across calls ------------------------
 * A queue owns `depth` independent tokenizer contexts (stream + scratch + pinned result buffers),
 * each served by its own host thread.  kp_queue_submit returns at once with a ticket; the caller's
 * utf8 / offsets must stay valid until kp_queue_wait(ticket) returns.  Tickets count 0, 1, 2, ...;
 * ticket k runs on context k % depth, and its result stays valid until ticket k + depth is submitted
 * (submit blocks while ticket k is still running).  Results are the compact form (kp_result8, pinned
 * host memory).  The reference has no such call: this is Tokenizer::tokenize over successive batches
 * with the copies of one batch hidden behind the kernels of its neighbours. */
typedef struct kp_queue kp_queue;
int kp_queue_create(const kp_dict* d, uint32_t depth, kp_queue** out);
int kp_queue_set_path(kp_queue* q, int path);    /* kp_tokenizer_set_path on every context */
int kp_queue_set_blocking_sync(kp_queue* q, int on);   /* kp_tokenizer_set_blocking_sync on every context */
int kp_queue_submit(kp_queue* q, const uint8_t* utf8, const uint64_t* offsets, uint64_t n_sent, uint64_t* ticket);
int kp_queue_wait(kp_queue* q, uint64_t ticket, kp_result8* out);
void kp_queue_destroy(kp_queue* q);

/* ---- multi-GPU in one process: sentence shards, one dictionary broadcast, token gather --
 * kp_shards_create packs the dictionary once, stages it on devices[0] and sends the packed blob to the
 * other devices with ONE ncclBroadcast (communicators from ncclCommInitAll; libnccl.so.2 is loaded
 * at run time, KP_ERR_CUDA if it is missing); every device gets a tokenizer context and a host
 * thread.  kp_shards_tokenize splits the batch into contiguous sentence ranges balanced by bytes, one
 * per device (global sentence order is kept, so gathering is concatenation); every device copies its
 * range in, runs the path, and copies its tokens out straight into ONE pinned result at their global
 * offsets -- for a host consumer the gather is the D2H itself.  kp_shards_tokenize_gather leaves the
 * result in device memory of devices[0] instead: token records travel over NVLink with grouped
 * ncclSend / ncclRecv (exact sizes).  Results are the compact form. */
typedef struct kp_shards kp_shards;
int kp_shards_create(const kp_dict_arrays* arrays, const int* devices, int n_devices, kp_shards** out);
int kp_shards_tokenize(kp_shards* g, const uint8_t* utf8, const uint64_t* offsets, uint64_t n_sent, kp_result8* out);
int kp_shards_tokenize_gather(kp_shards* g, const uint8_t* utf8, const uint64_t* offsets, uint64_t n_sent,
                              kp_result8* out_device);
/* milliseconds of the last kp_shards_* call: [0] whole call (wall), [1] slowest device pass (H2D + kernels),
 * [2] the NCCL gather (0 for kp_shards_tokenize), [3] the dictionary broadcast at create time */
int kp_shards_times(const kp_shards* g, float ms[4]);
int kp_shards_copy_to_host(kp_shards* g, void* dst, const void* device_src, uint64_t bytes);
void kp_shards_destroy(kp_shards* g);

/* ---- token gather across processes (one rank per GPU: torchrun, mpirun, ...) ---------------------------------
 * The second collective of SURVEY.md 8e for launchers that run one PROCESS per GPU.  Rank 0 obtains an NCCL
 * unique id (kp_gather_unique_id, 128 bytes) and hands it to the other ranks by whatever channel the launcher
 * has; every rank then creates its member with the same capacities (sentences / tokens per rank).
 * kp_gather_tokens takes each rank's device-resident compact result (kp_tokenize_batch_device8) and leaves, on
 * rank 0, the results of all ranks in device memory in rank order -- global sentence order for contiguous
 * shards -- with the token offsets rebased; other ranks get an empty *out.  One grouped ncclSend / ncclRecv of a
 * fixed-capacity block per rank, counts in the block's header, compaction by a kernel on rank 0: no count
 * exchange and no host round trip before the transfer.  KP_ERR_TOO_LARGE when a shard exceeds the capacity. */
typedef struct kp_gather kp_gather;
int kp_gather_unique_id(void* id128);
int kp_gather_create(int device, int rank, int world, const void* id128, uint64_t cap_sent, uint64_t cap_tok,
                     kp_gather** out);
int kp_gather_tokens(kp_gather* g, const kp_result8* mine, kp_result8* out);
int kp_gather_last_ms(const kp_gather* g, float* ms);     /* device time of the last kp_gather_tokens on this rank */
int kp_gather_copy_to_host(kp_gather* g, void* dst, const void* device_src, uint64_t bytes);
void kp_gather_destroy(kp_gather* g);

/* ---- lattice inspection (Lattice{nodes,edges} + viterbi internals) ---------------------------- */
typedef struct kp_lattice_node {
    int32_t id;          /* node.id(): 0 for BOS/EOS                    src/lattice/node.rs:27-33 */
    uint8_t cls;         /* KP_CLASS_* (Dummy = BOS/EOS)                                          */
    uint8_t reserved[3];
    uint32_t byte_pos;   /* node.byte_pos()                                                       */
    uint32_t char_pos;   /* node.char_pos()                                                       */
    uint32_t end_char;   /* index of the `edges` bucket holding the node (BOS 0, EOS n+1)         */
    int16_t left_id, right_id, cost;
    int16_t reserved2;
    int32_t dp;          /* dp[i] after viterbi(); INT32_MIN where the reference has None (BOS)   */
    int32_t pre;         /* pre_nodes[i] as a node index; -1 where the reference has None         */
} kp_lattice_node;       /* 36 bytes */

typedef struct kp_lattice {
    uint64_t n_nodes;                /* nodes in the reference's insertion order (BOS first, EOS last) */
    const kp_lattice_node* nodes;    /* HOST pointer, owned by the tokenizer                           */
} kp_lattice;
int kp_lattice_dump(kp_tokenizer* t, const uint8_t* utf8, uint64_t len, kp_lattice* out);

/* IndexTable::search_common_prefix_of on the device copy of the trie (one query; for parity tests
 * of the reference's known-answer vectors).  Writes up to `cap` (id, byte_len) pairs; *n = total hits
 * (0 where the reference returns None). */
int kp_da_common_prefix(kp_tokenizer* t, const uint8_t* utf8, uint64_t len, int expand_dup, int64_t* ids,
                        uint64_t* byte_lens, uint64_t cap, uint64_t* n);

/* ---- dictionary construction support (host only) ------------------------------------------------ */
/* da::build_with_ids (kanpyo-dict/src/trie/da.rs:206-217): double array over n_keys sorted unique keys
 * blob[off[k] .. off[k+1]) with leaf ids[k]; produces the same base/check array as the reference.
 * *out = malloc'ed int32[2 * *out_len] ({base, check} per node); release with kp_da_free. */
int kp_da_build(const uint8_t* blob, const uint64_t* off, uint64_t n_keys, const int64_t* ids, int32_t** out,
                uint64_t* out_len);
void kp_da_free(int32_t* p);

#ifdef __cplusplus
}
#endif
#endif /* KANPYO_B200_H */
