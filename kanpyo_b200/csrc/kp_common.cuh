// kp_common.cuh — shared types of the sm_100a implementation behind include/kanpyo_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/kanpyo_b200.h"

// ---- constants of the reference's hot path -----------------------------------------------------
constexpr int KP_ROOT_ID = 1;                  // kanpyo-dict/src/trie/da.rs:9
constexpr int KP_INF = 1 << 30;                // src/lattice.rs:117
constexpr uint32_t KP_MAX_UNKNOWN_LEN = 1024;  // src/lattice.rs:55
constexpr uint32_t KP_NONE = 0xFFFFFFFFu;      // Option::None for slot / node indices
constexpr uint32_t KP_SLOT_SHARED = 0x80000000u;   // tgt.y: the reduced slot is shared (unknown node); nodes per chunk < 2^31

// first-character table of the trie walk (kp_dict.cu): code points it covers, and its two markers
constexpr uint32_t KP_FIRST_CPS = 0x10000u;
constexpr int KP_FIRST_DEAD = -1;              // a transition fails inside the character: no hits at all
constexpr int KP_FIRST_SLOW = -2;              // not representable: walk this character from the root

// node.x packs id (30 bits) and class (2 bits)
constexpr uint32_t KP_ID_MASK = 0x3FFFFFFFu;
constexpr int KP_KIND_SHIFT = 30;

// ---- device view of the dictionary (passed by value to kernels) --------------------------------
struct kp_catinfo {   // one per char class (256 entries)
    int32_t unk_first;    // first 1-based unknown morph id (unk_dict.rs:15)
    uint32_t unk_count;   // 0 = class absent from the map
    uint32_t flags;       // bit0 invoke_list[c], bit1 group_list.get(c).unwrap_or(false)
    uint32_t pad;
};

struct kp_ddict {
    const int2* da;            // {base, check}
    uint32_t da_len;
    const uint16_t* dup;       // dense dup[id], id in [0, n_morphs]
    const short4* morphs;      // {left, right, cost, 0}, index id-1
    uint32_t n_morphs;
    const int16_t* conn;       // data[row*left_of_target + right_of_previous]
    uint32_t conn_row, conn_col;
    const int16_t* connT;      // transposed copy: connT[right_of_previous * connT_stride + left_of_target]
    uint32_t connT_stride;     // elements per row (conn_col rounded up to 64)
    const int2* first;         // [KP_FIRST_CPS] trie state after one whole character: {state, base[state]}
    uint32_t mid_char_keys;    // some key ends inside a UTF-8 character: probe for terminators after every byte
    const uint8_t* cat;        // code point -> class
    uint32_t n_cat;
    const kp_catinfo* catinfo; // [256]
    const short4* unk_morphs;
    uint32_t n_unk_morphs;
};

// ---- packed blob ---------------------------------------------------------------------------------
constexpr uint64_t KP_BLOB_MAGIC = 0x3144303032424B50ull;  // "PKB200D1" little-endian
struct kp_blob_header {
    uint64_t magic;
    uint32_t version;
    uint32_t header_size;
    uint64_t total_size;
    uint64_t da_len, n_morphs, conn_row, conn_col, n_cat, n_unk_morphs;
    uint64_t off_da, off_dup, off_morphs, off_conn, off_cat, off_catinfo, off_unk_morphs;
    uint64_t reserved[4];      // [0] offset of the transposed matrix, [1] its row stride in elements,
                               // [2] offset of the first-character table, [3] 1 = some key ends inside a character
    uint64_t checksum;         // FNV-1a (64-bit, 8 bytes at a time) of everything after the header, written by kp_pack:
                               // a blob that passes it carries exactly the indices kp_pack validated
};
constexpr uint32_t KP_BLOB_VERSION = 2;   // layout of kp_blob_header + sections (independent of KP_ABI_VERSION)

struct kp_dict {
    int device;
    void* d_blob;
    uint64_t size;
    std::string host_blob;   // host copy of the packed blob
    kp_ddict view;
};

// ---- error plumbing ------------------------------------------------------------------------------
void kp_set_error(const char* fmt, ...);
#define KP_CUDA(call)                                                                            \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            kp_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return KP_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

int kp_view_from_blob(const kp_blob_header* h, const void* d_blob, kp_ddict* v);
// Validates `host_blob` (header, section extents, checksum) and wraps the copy of it that already sits in
// device memory at d_blob (e.g. the receive buffer of the NCCL broadcast): the handle takes ownership
// of d_blob (cudaFree at kp_dict_destroy).  No second upload.
int kp_dict_adopt_device_blob(std::string&& host_blob, void* d_blob, int device, kp_dict** out);
