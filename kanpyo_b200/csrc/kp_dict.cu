// kp_dict.cu — dictionary validation, packing and one-time staging to HBM.
//
// Replaces (by construction from the same data) the read-only members of the reference's `Dict`
// that the hot path touches: IndexTable{da,dup} (kanpyo-dict/src/index.rs:10-13), Morphs
// (morph.rs:24), ConnectionTable (connection.rs:5-9), CharCategoryDef (char_category_def.rs:15-20)
// and UnkDict (unk_dict.rs:12-16).  Every index the reference would form at tokenize time with a
// hard `[]` (and panic on) is validated here once, so the kernels need no bounds checks on them.
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "kp_common.cuh"

static thread_local char g_err[512] = "";

void kp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* kp_last_error(void) { return g_err; }
extern "C" int kp_abi_version(void) { return KP_ABI_VERSION; }

extern "C" const char* kp_strerror(int s) {
    switch (s) {
        case KP_OK: return "ok";
        case KP_ERR_ARG: return "invalid argument";
        case KP_ERR_CUDA: return "CUDA error or no usable device (no CPU fallback exists)";
        case KP_ERR_DICT: return "dictionary arrays failed validation";
        case KP_ERR_UTF8: return "input is not valid UTF-8";
        case KP_ERR_NOMEM: return "out of memory";
        case KP_ERR_TOO_LARGE: return "chunk too large for 32-bit device indices";
        case KP_ERR_BLOB: return "bad dictionary blob";
        default: return "unknown status";
    }
}

extern "C" int kp_device_count(int* n) {
    if (!n) return KP_ERR_ARG;
    *n = 0;
    KP_CUDA(cudaGetDeviceCount(n));
    return KP_OK;
}

static inline uint64_t align256(uint64_t x) { return (x + 255) & ~uint64_t(255); }

// FNV-1a over 64-bit words (sections are 256-byte aligned, so the payload is a whole number of words)
static uint64_t kp_blob_checksum(const char* blob, uint64_t size) {
    uint64_t h = 0xcbf29ce484222325ull;
    const uint64_t* w = (const uint64_t*)(blob + align256(sizeof(kp_blob_header)));
    const uint64_t n = (size - align256(sizeof(kp_blob_header))) / 8;
    for (uint64_t i = 0; i < n; i++) h = (h ^ w[i]) * 0x100000001b3ull;
    return h;
}

// Pack + validate.  Returns KP_OK and fills `blob`.
static int kp_pack(const kp_dict_arrays* a, std::string* blob) {
    if (!a) return KP_ERR_ARG;
    if ((a->da_len && !a->da) || (a->n_dup && (!a->dup_ids || !a->dup_counts)) || (a->n_morphs && !a->morphs) ||
        !a->conn || (a->n_invoke && !a->invoke_list) || (a->n_group && !a->group_list) ||
        (a->n_unk_map && (!a->unk_cat || !a->unk_first_id || !a->unk_count)) || (a->n_unk_morphs && !a->unk_morphs)) {
        kp_set_error("kp_dict_arrays: null array with non-zero length");
        return KP_ERR_ARG;
    }
    if (a->da_len >= (1ull << 31) || a->n_morphs >= KP_ID_MASK || a->n_unk_morphs >= KP_ID_MASK) {
        kp_set_error("dictionary too large for 30-bit ids / 31-bit trie indices");
        return KP_ERR_DICT;
    }
    // ConnectionTable::get(row,col) = data[self.row*col + row] (connection.rs:12-14) is called with
    // (previous.right_id, target.left_id), BOS/EOS being (0,0): need a non-empty matrix.
    if (a->conn_row == 0 || a->conn_col == 0 || a->conn_row > 65536 || a->conn_col > 65536) {
        kp_set_error("connection matrix shape %llu x %llu unsupported", (unsigned long long)a->conn_row,
                     (unsigned long long)a->conn_col);
        return KP_ERR_DICT;
    }
    // CharCategoryDef::char_category falls back to entry 0 (char_category_def.rs:33-38): table must be non-empty.
    if (a->n_char_category == 0 || !a->char_category) {
        kp_set_error("char_category table is empty (the reference indexes entry 0)");
        return KP_ERR_DICT;
    }
    auto morph_ok = [&](const int16_t* m, uint64_t n, const char* what) -> bool {
        for (uint64_t i = 0; i < n; i++) {
            int l = m[3 * i], r = m[3 * i + 1];
            if (l < 0 || r < 0 || (uint64_t)l >= a->conn_col || (uint64_t)r >= a->conn_row) {
                kp_set_error("%s[%llu]: left_id %d / right_id %d outside the %llux%llu connection matrix", what,
                             (unsigned long long)i, l, r, (unsigned long long)a->conn_row,
                             (unsigned long long)a->conn_col);
                return false;
            }
        }
        return true;
    };
    if (!morph_ok(a->morphs, a->n_morphs, "morphs") || !morph_ok(a->unk_morphs, a->n_unk_morphs, "unk_morphs"))
        return KP_ERR_DICT;

    // dense dup[id] (index.rs:46: `self.dup.get(id).unwrap_or(&0)`)
    std::vector<uint16_t> dup(a->n_morphs + 2, 0);
    for (uint64_t i = 0; i < a->n_dup; i++) {
        int64_t id = a->dup_ids[i];
        uint64_t c = a->dup_counts[i];
        if (id < 1 || (uint64_t)id > a->n_morphs || c > 65535 || (uint64_t)id + c > a->n_morphs) {
            kp_set_error("dup entry %lld -> %llu outside morphs[0..%llu)", (long long)id, (unsigned long long)c,
                         (unsigned long long)a->n_morphs);
            return KP_ERR_DICT;
        }
        dup[(size_t)id] = (uint16_t)c;
    }
    // Every id a common-prefix search can return: a terminator child `q` of some node p
    // (q == base[p], check[q] == p, base[q] < 0) yields id = -base[q] (da.rs:166-174), expanded to
    // id ..= id + dup[id] (index.rs:46-51), each used as morphs[id-1] (lattice.rs:182).
    for (uint64_t q = 0; q < a->da_len; q++) {
        int32_t b = a->da[2 * q], p = a->da[2 * q + 1];
        if (b >= 0 || p < 0 || (uint64_t)p >= a->da_len) continue;
        if (a->da[2 * (uint64_t)p] != (int32_t)q) continue;
        int64_t id = -(int64_t)b;
        if (id < 1 || (uint64_t)id > a->n_morphs || (uint64_t)id + dup[(size_t)id] > a->n_morphs) {
            kp_set_error("trie leaf %llu carries id %lld outside morphs[0..%llu)", (unsigned long long)q, (long long)id,
                         (unsigned long long)a->n_morphs);
            return KP_ERR_DICT;
        }
    }
    // class table: invoke_list[cat] is a hard index (lattice.rs:54)
    for (uint64_t i = 0; i < a->n_char_category; i++) {
        if (a->char_category[i] >= a->n_invoke) {
            kp_set_error("char_category[%llu] = %u but invoke_list has %llu entries", (unsigned long long)i,
                         a->char_category[i], (unsigned long long)a->n_invoke);
            return KP_ERR_DICT;
        }
    }
    std::vector<kp_catinfo> ci(256);
    for (int c = 0; c < 256; c++) {
        ci[c] = kp_catinfo{0, 0, 0, 0};
        if ((uint64_t)c < a->n_invoke && a->invoke_list[c]) ci[c].flags |= 1u;
        if ((uint64_t)c < a->n_group && a->group_list[c]) ci[c].flags |= 2u;
    }
    for (uint64_t i = 0; i < a->n_unk_map; i++) {
        int64_t first = a->unk_first_id[i];
        uint64_t cnt = a->unk_count[i];
        if (cnt == 0) continue;
        if (first < 1 || (uint64_t)first + cnt - 1 > a->n_unk_morphs) {   // unk_dict.morphs[id-1], lattice.rs:195
            kp_set_error("unk map class %u -> (%lld,%llu) outside unk_morphs[0..%llu)", a->unk_cat[i], (long long)first,
                         (unsigned long long)cnt, (unsigned long long)a->n_unk_morphs);
            return KP_ERR_DICT;
        }
        ci[a->unk_cat[i]].unk_first = (int32_t)first;
        ci[a->unk_cat[i]].unk_count = (uint32_t)cnt;
    }

    kp_blob_header h;
    memset(&h, 0, sizeof(h));
    h.magic = KP_BLOB_MAGIC;
    h.version = KP_BLOB_VERSION;
    h.header_size = sizeof(h);
    h.da_len = a->da_len;
    h.n_morphs = a->n_morphs;
    h.conn_row = a->conn_row;
    h.conn_col = a->conn_col;
    h.n_cat = a->n_char_category;
    h.n_unk_morphs = a->n_unk_morphs;
    uint64_t o = align256(sizeof(h));
    h.off_da = o;          o = align256(o + (a->da_len ? a->da_len : 1) * 8);
    h.off_dup = o;         o = align256(o + dup.size() * 2);
    h.off_morphs = o;      o = align256(o + (a->n_morphs ? a->n_morphs : 1) * 8);
    h.off_conn = o;        o = align256(o + a->conn_row * a->conn_col * 2);
    h.off_cat = o;         o = align256(o + a->n_char_category);
    h.off_catinfo = o;     o = align256(o + 256 * sizeof(kp_catinfo));
    h.off_unk_morphs = o;  o = align256(o + (a->n_unk_morphs ? a->n_unk_morphs : 1) * 8);
    // transposed copy for the Viterbi sweep: row = right_id of the predecessor, column = left_id of the
    // target, rows padded to whole 128-byte lines (see kp_viterbi)
    const uint64_t strideT = (a->conn_col + 63) & ~uint64_t(63);
    if (a->conn_row * strideT * 2 >= (1ull << 31)) {
        kp_set_error("connection matrix %llu x %llu too large for 31-bit row offsets", (unsigned long long)a->conn_row,
                     (unsigned long long)a->conn_col);
        return KP_ERR_DICT;
    }
    h.reserved[0] = o;     o = align256(o + a->conn_row * strideT * 2);
    h.reserved[1] = strideT;
    h.reserved[2] = o;     o = align256(o + (uint64_t)KP_FIRST_CPS * 8);   // first-character table (below)
    h.total_size = o;
    blob->assign((size_t)o, '\0');
    char* p = &(*blob)[0];
    memcpy(p, &h, sizeof(h));
    if (a->da_len) memcpy(p + h.off_da, a->da, a->da_len * 8);
    memcpy(p + h.off_dup, dup.data(), dup.size() * 2);
    auto pack_morphs = [](char* dst, const int16_t* m, uint64_t n) {
        int16_t* d = (int16_t*)dst;
        for (uint64_t i = 0; i < n; i++) {
            d[4 * i] = m[3 * i];
            d[4 * i + 1] = m[3 * i + 1];
            d[4 * i + 2] = m[3 * i + 2];
            d[4 * i + 3] = 0;
        }
    };
    pack_morphs(p + h.off_morphs, a->morphs, a->n_morphs);
    memcpy(p + h.off_conn, a->conn, a->conn_row * a->conn_col * 2);
    {
        int16_t* t = (int16_t*)(p + h.reserved[0]);
        for (uint64_t left = 0; left < a->conn_col; left++)
            for (uint64_t right = 0; right < a->conn_row; right++)
                t[right * strideT + left] = a->conn[left * a->conn_row + right];   // get(right, left), connection.rs:12-14
    }
    {
        // First-character table: for every code point below 0x10000, the trie state reached from the
        // root after the bytes of that one character -- {state, base[state]}, or {KP_FIRST_DEAD, 0}
        // when a transition fails inside the character (then search_common_prefix_of returns
        // nothing, da.rs:160-164).  The counting walk starts there instead of at the root.  Keys are
        // whole UTF-8 strings, so no key can end inside a character; a table entry whose skipped
        // terminator probes (da.rs:165-174) would hit anyway is marked KP_FIRST_SLOW and that
        // character is walked byte by byte from the root.
        int32_t* t = (int32_t*)(p + h.reserved[2]);
        const int32_t* da = a->da;
        const uint64_t n = a->da_len;
        for (uint32_t cp = 0; cp < KP_FIRST_CPS; cp++) {
            int32_t st = KP_FIRST_SLOW, stbase = 0;
            uint8_t bytes[3];
            int L = 0;
            if (cp < 0x80) { bytes[0] = (uint8_t)cp; L = 1; }
            else if (cp < 0x800) { bytes[0] = (uint8_t)(0xC0 | (cp >> 6)); bytes[1] = (uint8_t)(0x80 | (cp & 0x3F)); L = 2; }
            else if (cp < 0xD800 || cp >= 0xE000) {
                bytes[0] = (uint8_t)(0xE0 | (cp >> 12)); bytes[1] = (uint8_t)(0x80 | ((cp >> 6) & 0x3F));
                bytes[2] = (uint8_t)(0x80 | (cp & 0x3F)); L = 3;
            }
            if (L && n > (uint64_t)KP_ROOT_ID) {
                int64_t prev = KP_ROOT_ID, base = da[2 * KP_ROOT_ID];
                bool dead = false, slow = false;
                int64_t q = 0;
                for (int k = 0; k < L && !dead; k++) {
                    q = base + bytes[k];
                    if (q < 0 || (uint64_t)q >= n || da[2 * q + 1] != prev) { dead = true; break; }
                    const int64_t ahead = da[2 * q];
                    if (k < L - 1 && ahead >= 0 && (uint64_t)ahead < n && da[2 * ahead + 1] == q && da[2 * ahead] < 0)
                        slow = true;
                    prev = q;
                    base = da[2 * q];
                }
                if (slow) { st = KP_FIRST_SLOW; }
                else if (dead) { st = KP_FIRST_DEAD; }
                else { st = (int32_t)q; stbase = (int32_t)base; }
            }
            t[2 * cp] = st;
            t[2 * cp + 1] = stbase;
        }
    }
    {
        // Do all keys end on a character boundary?  The reference builds its trie from `String`s, so yes
        // for any dictionary it can produce; then a terminator probe (da.rs:165-174) in the middle of a
        // character of (valid UTF-8) input can never hit, and the counting walk skips it.  Checked here
        // for the arrays actually given: for every leaf (base < 0, reached as base[p] + 0 from p =
        // check[leaf]) the bytes on the way up from p must end with a whole character -- an ASCII byte,
        // or k continuation bytes under a lead byte announcing k.  Otherwise reserved[3] = 1 and the
        // walk probes after every byte, as the reference does.
        // The same walk up to the root gives the key's length: node records carry a word's length in 16
        // bits (kp_kernels.cuh), so a key longer than 65 535 bytes is refused here rather than truncated.
        const int32_t* da = a->da;
        const uint64_t n = a->da_len;
        bool mid = false;
        for (uint64_t q = 0; q < n; q++) {
            const int64_t p0 = da[2 * q + 1];
            if (da[2 * q] >= 0 || p0 < 0 || (uint64_t)p0 >= n || da[2 * p0] != (int32_t)q) continue;   // not a terminator
            int64_t st = p0;
            int conts = 0;
            bool tail_done = false;
            uint32_t depth = 0;
            while (st != KP_ROOT_ID) {               // byte that led to `st`: st - base[check[st]]
                const int64_t par = da[2 * st + 1];
                if (par < 0 || (uint64_t)par >= n || ++depth > 65535) {
                    if (depth > 65535) {
                        kp_set_error("trie leaf %llu: key longer than 65535 bytes (or a cycle in `check`)",
                                     (unsigned long long)q);
                        return KP_ERR_DICT;
                    }
                    mid = true;                      // not reachable from the root: never matched, but do not trust it
                    break;
                }
                if (!tail_done) {
                    const int64_t c = st - da[2 * par];
                    if (c < 0 || c > 255) { mid = true; tail_done = true; }
                    else if ((c & 0xC0) == 0x80) { if (++conts > 3) { mid = true; tail_done = true; } }
                    else {
                        const int need = c < 0x80 ? 0 : c >= 0xF0 ? 3 : c >= 0xE0 ? 2 : c >= 0xC0 ? 1 : -1;
                        if (need != conts) mid = true;
                        tail_done = true;
                    }
                }
                st = par;
            }
            if (!tail_done && conts) mid = true;     // continuation bytes straight under the root
        }
        ((kp_blob_header*)p)->reserved[3] = mid ? 1 : 0;
    }
    memcpy(p + h.off_cat, a->char_category, a->n_char_category);
    memcpy(p + h.off_catinfo, ci.data(), 256 * sizeof(kp_catinfo));
    pack_morphs(p + h.off_unk_morphs, a->unk_morphs, a->n_unk_morphs);
    ((kp_blob_header*)p)->checksum = kp_blob_checksum(p, o);
    return KP_OK;
}

// A blob is accepted only if it is byte-for-byte what kp_pack wrote: magic / version / size, every
// section inside the blob, and the payload checksum.  kp_pack validated every index the kernels form
// (they have no bounds checks), so a truncated, stale or foreign blob must stop here.
static int kp_check_blob(const void* blob, uint64_t size) {
    const kp_blob_header* h = (const kp_blob_header*)blob;
    if (size < sizeof(kp_blob_header) || h->magic != KP_BLOB_MAGIC || h->version != KP_BLOB_VERSION ||
        h->header_size != sizeof(kp_blob_header) || h->total_size != size || (size & 255)) {
        kp_set_error("dictionary blob: bad magic/version/size");
        return KP_ERR_BLOB;
    }
    const uint64_t strideT = h->reserved[1];
    if (h->da_len >= (1ull << 31) || h->n_morphs >= KP_ID_MASK || h->n_unk_morphs >= KP_ID_MASK || h->conn_row == 0 ||
        h->conn_col == 0 || h->conn_row > 65536 || h->conn_col > 65536 || h->n_cat == 0 || strideT < h->conn_col ||
        strideT > 65536 + 64 || (strideT & 63)) {
        kp_set_error("dictionary blob: section sizes out of range");
        return KP_ERR_BLOB;
    }
    const struct { uint64_t off, len; } sec[] = {
        {h->off_da, (h->da_len ? h->da_len : 1) * 8},         {h->off_dup, (h->n_morphs + 2) * 2},
        {h->off_morphs, (h->n_morphs ? h->n_morphs : 1) * 8}, {h->off_conn, h->conn_row * h->conn_col * 2},
        {h->off_cat, h->n_cat},                               {h->off_catinfo, 256 * sizeof(kp_catinfo)},
        {h->off_unk_morphs, (h->n_unk_morphs ? h->n_unk_morphs : 1) * 8},
        {h->reserved[0], h->conn_row * strideT * 2},          {h->reserved[2], (uint64_t)KP_FIRST_CPS * 8}};
    for (const auto& x : sec)
        if (x.off < sizeof(kp_blob_header) || (x.off & 255) || x.off > size || x.len > size - x.off) {
            kp_set_error("dictionary blob: a section lies outside the blob");
            return KP_ERR_BLOB;
        }
    if (kp_blob_checksum((const char*)blob, size) != h->checksum) {
        kp_set_error("dictionary blob: checksum mismatch (truncated, stale or foreign blob)");
        return KP_ERR_BLOB;
    }
    return KP_OK;
}

int kp_view_from_blob(const kp_blob_header* h, const void* d_blob, kp_ddict* v) {
    const char* p = (const char*)d_blob;
    v->da = (const int2*)(p + h->off_da);
    v->da_len = (uint32_t)h->da_len;
    v->dup = (const uint16_t*)(p + h->off_dup);
    v->morphs = (const short4*)(p + h->off_morphs);
    v->n_morphs = (uint32_t)h->n_morphs;
    v->conn = (const int16_t*)(p + h->off_conn);
    v->conn_row = (uint32_t)h->conn_row;
    v->conn_col = (uint32_t)h->conn_col;
    v->connT = (const int16_t*)(p + h->reserved[0]);
    v->connT_stride = (uint32_t)h->reserved[1];
    v->first = (const int2*)(p + h->reserved[2]);
    v->mid_char_keys = h->reserved[3] != 0;
    v->cat = (const uint8_t*)(p + h->off_cat);
    v->n_cat = (uint32_t)h->n_cat;
    v->catinfo = (const kp_catinfo*)(p + h->off_catinfo);
    v->unk_morphs = (const short4*)(p + h->off_unk_morphs);
    v->n_unk_morphs = (uint32_t)h->n_unk_morphs;
    return KP_OK;
}

static int kp_require_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        kp_set_error("no CUDA device available (%s); kanpyo_b200 has no CPU fallback",
                     e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return KP_ERR_CUDA;
    }
    if (device < 0 || device >= n) {
        kp_set_error("device %d out of range (%d visible)", device, n);
        return KP_ERR_ARG;
    }
    cudaDeviceProp prop;
    KP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        kp_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return KP_ERR_CUDA;
    }
    return KP_OK;
}

static int kp_upload(std::string&& blob, int device, kp_dict** out) {
    int rc = kp_require_device(device);
    if (rc) return rc;
    KP_CUDA(cudaSetDevice(device));
    kp_dict* d = new kp_dict();
    d->device = device;
    d->size = blob.size();
    d->host_blob = std::move(blob);
    cudaError_t e = cudaMalloc(&d->d_blob, d->size);
    if (e != cudaSuccess) {
        kp_set_error("cudaMalloc(%llu) for the dictionary failed: %s", (unsigned long long)d->size, cudaGetErrorString(e));
        delete d;
        return KP_ERR_NOMEM;
    }
    e = cudaMemcpy(d->d_blob, d->host_blob.data(), d->size, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        kp_set_error("dictionary upload failed: %s", cudaGetErrorString(e));
        cudaFree(d->d_blob);
        delete d;
        return KP_ERR_CUDA;
    }
    kp_view_from_blob((const kp_blob_header*)d->host_blob.data(), d->d_blob, &d->view);
    *out = d;
    return KP_OK;
}

extern "C" int kp_dict_create(const kp_dict_arrays* arrays, int device, kp_dict** out) {
    if (!arrays || !out) return KP_ERR_ARG;
    *out = nullptr;
    std::string blob;
    int rc = kp_pack(arrays, &blob);
    if (rc) return rc;
    return kp_upload(std::move(blob), device, out);
}

// Host-only packing (no device needed): lets a launcher build the blob on one rank and ship it.
extern "C" int kp_dict_pack(const kp_dict_arrays* arrays, void* dst, uint64_t cap, uint64_t* size) {
    if (!arrays || !size) return KP_ERR_ARG;
    std::string blob;
    int rc = kp_pack(arrays, &blob);
    if (rc) return rc;
    *size = blob.size();
    if (dst) {
        if (cap < blob.size()) return KP_ERR_ARG;
        memcpy(dst, blob.data(), blob.size());
    }
    return KP_OK;
}

extern "C" int kp_dict_blob(const kp_dict* d, const void** host_ptr, uint64_t* size) {
    if (!d || !host_ptr || !size) return KP_ERR_ARG;
    *host_ptr = d->host_blob.data();
    *size = d->size;
    return KP_OK;
}

extern "C" int kp_dict_device_blob(const kp_dict* d, const void** device_ptr, uint64_t* size) {
    if (!d || !device_ptr || !size) return KP_ERR_ARG;
    *device_ptr = d->d_blob;
    *size = d->size;
    return KP_OK;
}

extern "C" int kp_dict_create_from_blob(const void* host_blob, uint64_t size, int device, kp_dict** out) {
    if (!host_blob || !out) return KP_ERR_ARG;
    *out = nullptr;
    int rc = kp_check_blob(host_blob, size);
    if (rc) return rc;
    std::string blob((const char*)host_blob, (size_t)size);
    return kp_upload(std::move(blob), device, out);
}

int kp_dict_adopt_device_blob(std::string&& host_blob, void* d_blob, int device, kp_dict** out) {
    int rc = kp_check_blob(host_blob.data(), host_blob.size());
    if (rc) return rc;
    kp_dict* d = new kp_dict();
    d->device = device;
    d->size = host_blob.size();
    d->host_blob = std::move(host_blob);
    d->d_blob = d_blob;
    kp_view_from_blob((const kp_blob_header*)d->host_blob.data(), d->d_blob, &d->view);
    *out = d;
    return KP_OK;
}

// From a packed blob already in device memory (e.g. the receive buffer of an NCCL broadcast): one
// device-to-host copy for validation and for the handle's host copy, one device-to-device copy
// into memory the handle owns.  The caller's buffer is not kept.
extern "C" int kp_dict_create_from_device_blob(const void* device_blob, uint64_t size, int device, kp_dict** out) {
    if (!device_blob || !out || size < sizeof(kp_blob_header) || size >= (1ull << 40)) return KP_ERR_ARG;
    *out = nullptr;
    int rc = kp_require_device(device);
    if (rc) return rc;
    KP_CUDA(cudaSetDevice(device));
    std::string blob((size_t)size, '\0');
    KP_CUDA(cudaMemcpy(&blob[0], device_blob, size, cudaMemcpyDeviceToHost));
    rc = kp_check_blob(blob.data(), size);
    if (rc) return rc;
    void* mine = nullptr;
    if (cudaMalloc(&mine, size) != cudaSuccess) {
        cudaGetLastError();
        kp_set_error("cudaMalloc(%llu) for the dictionary failed", (unsigned long long)size);
        return KP_ERR_NOMEM;
    }
    cudaError_t e = cudaMemcpy(mine, device_blob, size, cudaMemcpyDeviceToDevice);
    if (e != cudaSuccess) {
        cudaFree(mine);
        kp_set_error("dictionary copy failed: %s", cudaGetErrorString(e));
        return KP_ERR_CUDA;
    }
    rc = kp_dict_adopt_device_blob(std::move(blob), mine, device, out);
    if (rc) cudaFree(mine);
    return rc;
}

extern "C" void kp_dict_destroy(kp_dict* d) {
    if (!d) return;
    cudaSetDevice(d->device);
    cudaFree(d->d_blob);
    delete d;
}
