// kp_queue.cu — asynchronous batches (kp_queue_*) and multi-GPU sharding in one process (kp_shards_*).
//
// Both are host-side plumbing around the tokenizer contexts of kp_api.cu; neither adds a device code
// path.  Sentences are independent and the dictionary is read-only, so
//   * successive batches can overlap: the queue runs `depth` contexts, each on its own stream and host
//     thread, so the H2D of batch n+1 and the D2H of batch n-1 hide behind the kernels of batch n;
//   * one batch can be split into contiguous, byte-balanced sentence ranges, one per GPU.  The only
//     inter-GPU traffic is ONE ncclBroadcast of the packed dictionary at create time and, when the
//     caller wants the result in device memory, the gather of token records to devices[0]
//     (grouped ncclSend / ncclRecv of exact sizes).  For a host consumer every GPU copies its tokens
//     straight to their global offsets in one pinned result: the gather is the D2H itself.
// NCCL is bound at run time (dlopen of libnccl.so.2): the library itself links nothing but the static
// CUDA runtime, and a single-GPU kp_shards never touches NCCL.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "kp_common.cuh"

// two-phase single-chunk pass (kp_api.cu)
int kp_pass_begin(kp_tokenizer* t, const uint8_t* utf8, const uint64_t* offsets, uint64_t n_sent, uint64_t* n_tokens);
int kp_pass_pack(kp_tokenizer* t, uint64_t tok_base, const uint32_t** d_tok_off, const kp_token8** d_tokens,
                 const int32_t** d_eos, cudaStream_t* stream);

namespace {

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- run-time binding of NCCL ------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
};

int nccl_load(NcclApi** out) {
    static std::mutex mu;
    static NcclApi api;
    std::lock_guard<std::mutex> lk(mu);
    if (!api.handle) {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            kp_set_error("kp_shards needs NCCL for more than one device: %s", dlerror());
            return KP_ERR_CUDA;
        }
#define KP_NCCL_SYM(field, name)                                          \
    api.field = (decltype(api.field))dlsym(h, name);                      \
    if (!api.field) {                                                     \
        kp_set_error("libnccl.so.2 lacks %s", name);                      \
        dlclose(h);                                                       \
        return KP_ERR_CUDA;                                               \
    }
        KP_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        KP_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        KP_NCCL_SYM(Broadcast, "ncclBroadcast")
        KP_NCCL_SYM(Send, "ncclSend")
        KP_NCCL_SYM(Recv, "ncclRecv")
        KP_NCCL_SYM(GroupStart, "ncclGroupStart")
        KP_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        KP_NCCL_SYM(GetErrorString, "ncclGetErrorString")
        KP_NCCL_SYM(GetVersion, "ncclGetVersion")
        KP_NCCL_SYM(AllReduce, "ncclAllReduce")
#undef KP_NCCL_SYM
        api.handle = h;
    }
    *out = &api;
    return KP_OK;
}

#define KP_NCCL(api, call)                                                                    \
    do {                                                                                      \
        ncclResult_t r__ = (call);                                                            \
        if (r__ != ncclSuccess) {                                                             \
            kp_set_error("%s failed: %s", #call, (api)->GetErrorString(r__));                 \
            return KP_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

// ---- a host thread that runs one job at a time ------------------------------------------------------
struct Worker {
    std::mutex mu;
    std::condition_variable cv;
    std::thread th;
    std::function<int()> job;      // set by post(); cleared when done
    bool busy = false, quit = false;
    int status = KP_OK;
    std::string error;

    void start() {
        th = std::thread([this] {
            std::unique_lock<std::mutex> lk(mu);
            while (true) {
                cv.wait(lk, [this] { return quit || (busy && job); });
                if (quit) return;
                std::function<int()> f = std::move(job);
                job = nullptr;
                lk.unlock();
                int rc = f();
                std::string err = rc ? kp_last_error() : "";
                lk.lock();
                status = rc;
                error = std::move(err);
                busy = false;
                cv.notify_all();
            }
        });
    }
    void post(std::function<int()> f) {       // caller guarantees !busy (wait() first)
        std::lock_guard<std::mutex> lk(mu);
        job = std::move(f);
        busy = true;
        cv.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this] { return !busy; });
        if (status) kp_set_error("%s", error.c_str());
        return status;
    }
    void stop() {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
            cv.notify_all();
        }
        if (th.joinable()) th.join();
    }
};

struct Pinned {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return KP_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        if (cudaHostAlloc(&p, want, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            kp_set_error("cudaHostAlloc(%zu) failed", want);
            return KP_ERR_NOMEM;
        }
        cap = want;
        return KP_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct DeviceMem {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return KP_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            kp_set_error("cudaMalloc(%zu) failed", want);
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
};

}  // namespace

// =====================================================================================================
// kp_queue
// =====================================================================================================
struct kp_queue_slot {
    kp_tokenizer* tok = nullptr;
    Worker worker;
    uint64_t ticket = ~0ull;       // ticket whose result the slot holds / is computing
    kp_result8 result = {};
};

struct kp_queue {
    uint32_t depth = 0;
    uint64_t next_ticket = 0;
    std::vector<kp_queue_slot*> slots;
};

extern "C" int kp_queue_create(const kp_dict* d, uint32_t depth, kp_queue** out) {
    if (!d || !out || depth == 0 || depth > 16) return KP_ERR_ARG;
    *out = nullptr;
    kp_queue* q = new kp_queue();
    q->depth = depth;
    for (uint32_t i = 0; i < depth; i++) {
        kp_queue_slot* s = new kp_queue_slot();
        q->slots.push_back(s);
        int rc = kp_tokenizer_create(d, &s->tok);
        if (rc) {
            kp_queue_destroy(q);
            return rc;
        }
        s->worker.start();
    }
    *out = q;
    return KP_OK;
}

extern "C" int kp_queue_set_path(kp_queue* q, int path) {
    if (!q) return KP_ERR_ARG;
    for (kp_queue_slot* s : q->slots) {
        s->worker.wait();
        int rc = kp_tokenizer_set_path(s->tok, path);
        if (rc) return rc;
    }
    return KP_OK;
}

extern "C" int kp_queue_set_blocking_sync(kp_queue* q, int on) {
    if (!q) return KP_ERR_ARG;
    for (kp_queue_slot* s : q->slots) {
        s->worker.wait();
        int rc = kp_tokenizer_set_blocking_sync(s->tok, on);
        if (rc) return rc;
    }
    return KP_OK;
}

extern "C" int kp_queue_submit(kp_queue* q, const uint8_t* utf8, const uint64_t* offsets, uint64_t n_sent,
                               uint64_t* ticket) {
    if (!q || !offsets || !ticket) return KP_ERR_ARG;
    const uint64_t tk = q->next_ticket;
    kp_queue_slot* s = q->slots[tk % q->depth];
    s->worker.wait();                 // ticket tk - depth still running: its result is about to be replaced
    q->next_ticket = tk + 1;
    s->ticket = tk;
    s->worker.post([s, utf8, offsets, n_sent]() -> int { return kp_tokenize_batch8(s->tok, utf8, offsets, n_sent, &s->result); });
    *ticket = tk;
    return KP_OK;
}

extern "C" int kp_queue_wait(kp_queue* q, uint64_t ticket, kp_result8* out) {
    if (!q || !out || ticket >= q->next_ticket) return KP_ERR_ARG;
    kp_queue_slot* s = q->slots[ticket % q->depth];
    if (s->ticket != ticket) {
        kp_set_error("ticket %llu: its result was replaced by ticket %llu", (unsigned long long)ticket,
                     (unsigned long long)s->ticket);
        return KP_ERR_ARG;
    }
    int rc = s->worker.wait();
    if (rc) return rc;
    *out = s->result;
    return KP_OK;
}

extern "C" void kp_queue_destroy(kp_queue* q) {
    if (!q) return;
    for (kp_queue_slot* s : q->slots) {
        s->worker.wait();
        s->worker.stop();
        if (s->tok) kp_tokenizer_destroy(s->tok);
        delete s;
    }
    delete q;
}

// =====================================================================================================
// kp_shards
// =====================================================================================================
struct kp_shard_dev {
    int device = 0;
    kp_dict* dict = nullptr;
    kp_tokenizer* tok = nullptr;
    ncclComm_t comm = nullptr;
    cudaStream_t bstream = nullptr;    // stream of the dictionary broadcast
    Worker worker;
    // per call
    uint64_t s0 = 0, s1 = 0, ntok = 0, tok_base = 0;
    double pass_ms = 0, gather_ms = 0;
};

struct kp_shards {
    int n = 0;
    NcclApi* nccl = nullptr;
    std::vector<kp_shard_dev*> dev;
    Pinned h_tok_off, h_tokens, h_eos;
    DeviceMem g_tok_off, g_tokens, g_eos;   // gathered result on devices[0]
    float times[4] = {0, 0, 0, 0};
};

static void shard_ranges(const uint64_t* off, uint64_t n_sent, int n, std::vector<uint64_t>* cuts) {
    // contiguous sentence ranges balanced by cumulative bytes
    cuts->assign((size_t)n + 1, 0);
    const uint64_t total = off[n_sent] - off[0];
    for (int k = 1; k < n; k++) {
        const uint64_t target = off[0] + total / (uint64_t)n * (uint64_t)k + total % (uint64_t)n * (uint64_t)k / (uint64_t)n;
        uint64_t c = (uint64_t)(std::lower_bound(off, off + n_sent + 1, target) - off);
        c = std::min(c, n_sent);
        (*cuts)[k] = std::max(c, (*cuts)[k - 1]);
    }
    (*cuts)[n] = n_sent;
}

extern "C" int kp_shards_create(const kp_dict_arrays* arrays, const int* devices, int n_devices, kp_shards** out) {
    if (!arrays || !out || n_devices < 1 || n_devices > 64) return KP_ERR_ARG;
    *out = nullptr;
    std::vector<int> devs(n_devices);
    for (int i = 0; i < n_devices; i++) devs[i] = devices ? devices[i] : i;
    // validated, packed blob (host)
    uint64_t size = 0;
    int rc = kp_dict_pack(arrays, nullptr, 0, &size);
    if (rc) return rc;
    std::string blob((size_t)size, '\0');
    rc = kp_dict_pack(arrays, &blob[0], size, &size);
    if (rc) return rc;

    kp_shards* g = new kp_shards();
    g->n = n_devices;
    for (int i = 0; i < n_devices; i++) {
        g->dev.push_back(new kp_shard_dev());
        g->dev[i]->device = devs[i];
    }
    auto fail = [&](int code) {
        kp_shards_destroy(g);
        return code;
    };
    // devices[0] stages the dictionary from the host
    rc = kp_dict_create_from_blob(blob.data(), size, devs[0], &g->dev[0]->dict);
    if (rc) return fail(rc);
    if (n_devices > 1) {
        rc = nccl_load(&g->nccl);
        if (rc) return fail(rc);
        NcclApi* N = g->nccl;
        std::vector<ncclComm_t> comms(n_devices);
        ncclResult_t nr = N->CommInitAll(comms.data(), n_devices, devs.data());
        if (nr != ncclSuccess) {
            kp_set_error("ncclCommInitAll failed: %s", N->GetErrorString(nr));
            return fail(KP_ERR_CUDA);
        }
        for (int i = 0; i < n_devices; i++) g->dev[i]->comm = comms[i];
        const void* blob0 = nullptr;
        uint64_t sz0 = 0;
        kp_dict_device_blob(g->dev[0]->dict, &blob0, &sz0);
        std::vector<void*> recv(n_devices, nullptr);
        for (int i = 0; i < n_devices; i++) {
            if (cudaSetDevice(devs[i]) != cudaSuccess || cudaStreamCreateWithFlags(&g->dev[i]->bstream, cudaStreamNonBlocking) != cudaSuccess) {
                kp_set_error("device %d: stream creation failed", devs[i]);
                return fail(KP_ERR_CUDA);
            }
            if (i > 0 && cudaMalloc(&recv[i], size) != cudaSuccess) {
                cudaGetLastError();
                kp_set_error("device %d: cudaMalloc(%llu) for the dictionary failed", devs[i], (unsigned long long)size);
                return fail(KP_ERR_NOMEM);
            }
        }
        // ONE broadcast of the packed dictionary from devices[0]
        const double t0 = now_ms();
        nr = N->GroupStart();
        for (int i = 0; i < n_devices && nr == ncclSuccess; i++)
            nr = N->Broadcast(i == 0 ? blob0 : recv[i], i == 0 ? (void*)blob0 : recv[i], size, ncclChar, 0, comms[i],
                              g->dev[i]->bstream);
        ncclResult_t ne = N->GroupEnd();
        if (nr == ncclSuccess) nr = ne;
        if (nr != ncclSuccess) {
            kp_set_error("ncclBroadcast of the dictionary failed: %s", N->GetErrorString(nr));
            for (int i = 1; i < n_devices; i++) {
                cudaSetDevice(devs[i]);
                cudaFree(recv[i]);
            }
            return fail(KP_ERR_CUDA);
        }
        for (int i = 0; i < n_devices; i++) {
            cudaSetDevice(devs[i]);
            cudaStreamSynchronize(g->dev[i]->bstream);
        }
        g->times[3] = (float)(now_ms() - t0);
        // every receiver checks what arrived (header, extents, checksum) and wraps its receive buffer
        for (int i = 1; i < n_devices; i++) {
            cudaSetDevice(devs[i]);
            std::string copy((size_t)size, '\0');
            cudaError_t e = cudaMemcpy(&copy[0], recv[i], size, cudaMemcpyDeviceToHost);
            if (e == cudaSuccess) rc = kp_dict_adopt_device_blob(std::move(copy), recv[i], devs[i], &g->dev[i]->dict);
            else {
                kp_set_error("device %d: reading back the broadcast dictionary failed: %s", devs[i], cudaGetErrorString(e));
                rc = KP_ERR_CUDA;
            }
            if (rc) {
                for (int j = i; j < n_devices; j++) {
                    cudaSetDevice(devs[j]);
                    cudaFree(recv[j]);
                }
                return fail(rc);
            }
        }
    }
    for (int i = 0; i < n_devices; i++) {
        rc = kp_tokenizer_create(g->dev[i]->dict, &g->dev[i]->tok);
        if (rc) return fail(rc);
        g->dev[i]->worker.start();
    }
    *out = g;
    return KP_OK;
}

extern "C" void kp_shards_destroy(kp_shards* g) {
    if (!g) return;
    const int dev0 = g->dev.empty() ? -1 : g->dev[0]->device;
    for (kp_shard_dev* d : g->dev) {
        if (d->worker.th.joinable()) {
            d->worker.wait();
            d->worker.stop();
        }
        if (d->tok) kp_tokenizer_destroy(d->tok);
        if (d->comm && g->nccl) g->nccl->CommDestroy(d->comm);
        if (d->bstream) {
            cudaSetDevice(d->device);
            cudaStreamDestroy(d->bstream);
        }
        if (d->dict) kp_dict_destroy(d->dict);
        delete d;
    }
    if (dev0 >= 0) cudaSetDevice(dev0);        // the gathered result lives on devices[0]
    g->g_tok_off.release();
    g->g_tokens.release();
    g->g_eos.release();
    g->h_tok_off.release();
    g->h_tokens.release();
    g->h_eos.release();
    delete g;
}

// phase 1 on every device: copy the shard in, run the path up to the scanned token counts
static int shards_begin(kp_shards* g, const uint8_t* utf8, const uint64_t* offsets, uint64_t n_sent, uint64_t* total) {
    if (n_sent >= (1ull << 32) - 1) return KP_ERR_TOO_LARGE;
    for (uint64_t s = 0; s < n_sent; s++)
        if (offsets[s + 1] < offsets[s]) {
            kp_set_error("offsets[%llu] > offsets[%llu]", (unsigned long long)s, (unsigned long long)(s + 1));
            return KP_ERR_ARG;
        }
    std::vector<uint64_t> cuts;
    shard_ranges(offsets, n_sent, g->n, &cuts);
    for (int i = 0; i < g->n; i++) {
        kp_shard_dev* d = g->dev[i];
        d->s0 = cuts[i];
        d->s1 = cuts[i + 1];
        d->worker.post([d, utf8, offsets]() -> int {
            const double t0 = now_ms();
            int rc = kp_pass_begin(d->tok, utf8, offsets + d->s0, d->s1 - d->s0, &d->ntok);
            d->pass_ms = now_ms() - t0;
            return rc;
        });
    }
    int rc = KP_OK;
    std::string err;
    for (int i = 0; i < g->n; i++) {
        int r = g->dev[i]->worker.wait();
        if (r && !rc) {
            rc = r;
            err = kp_last_error();
        }
    }
    if (rc) {
        kp_set_error("%s", err.c_str());
        return rc;
    }
    uint64_t base = 0;
    double slowest = 0;
    for (int i = 0; i < g->n; i++) {
        g->dev[i]->tok_base = base;
        base += g->dev[i]->ntok;
        slowest = std::max(slowest, g->dev[i]->pass_ms);
    }
    if (base >= (1ull << 32)) return KP_ERR_TOO_LARGE;
    g->times[1] = (float)slowest;
    *total = base;
    return KP_OK;
}

static int shards_finish(kp_shards* g) {
    int rc = KP_OK;
    std::string err;
    for (int i = 0; i < g->n; i++) {
        int r = g->dev[i]->worker.wait();
        if (r && !rc) {
            rc = r;
            err = kp_last_error();
        }
    }
    if (rc) kp_set_error("%s", err.c_str());
    return rc;
}

extern "C" int kp_shards_tokenize(kp_shards* g, const uint8_t* utf8, const uint64_t* offsets, uint64_t n_sent,
                                  kp_result8* out) {
    if (!g || !offsets || !out || (n_sent && offsets[n_sent] > offsets[0] && !utf8)) return KP_ERR_ARG;
    const double t0 = now_ms();
    uint64_t total = 0;
    int rc = shards_begin(g, utf8, offsets, n_sent, &total);
    if (rc) return rc;
    if ((rc = g->h_tok_off.ensure(sizeof(uint32_t) * (n_sent + 1))) || (rc = g->h_tokens.ensure(sizeof(kp_token8) * (total + 1))) ||
        (rc = g->h_eos.ensure(sizeof(int32_t) * (n_sent + 1))))
        return rc;
    uint32_t* h_off = (uint32_t*)g->h_tok_off.p;
    kp_token8* h_tok = (kp_token8*)g->h_tokens.p;
    int32_t* h_eos = (int32_t*)g->h_eos.p;
    // phase 2: every device packs with its global token base and copies straight to the global offsets
    for (int i = 0; i < g->n; i++) {
        kp_shard_dev* d = g->dev[i];
        const bool last = i == g->n - 1;
        d->worker.post([d, h_off, h_tok, h_eos, last]() -> int {
            const uint32_t* d_off;
            const kp_token8* d_tok;
            const int32_t* d_eos;
            cudaStream_t st;
            int rc = kp_pass_pack(d->tok, d->tok_base, &d_off, &d_tok, &d_eos, &st);
            if (rc) return rc;
            const uint64_t S = d->s1 - d->s0;
            // the entry closing this shard opens the next one: only the last shard writes it
            KP_CUDA(cudaMemcpyAsync(h_off + d->s0, d_off, sizeof(uint32_t) * (S + (last ? 1 : 0)), cudaMemcpyDeviceToHost, st));
            if (d->ntok)
                KP_CUDA(cudaMemcpyAsync(h_tok + d->tok_base, d_tok, sizeof(kp_token8) * d->ntok, cudaMemcpyDeviceToHost, st));
            if (S) KP_CUDA(cudaMemcpyAsync(h_eos + d->s0, d_eos, sizeof(int32_t) * S, cudaMemcpyDeviceToHost, st));
            KP_CUDA(cudaStreamSynchronize(st));
            return (int)KP_OK;
        });
    }
    rc = shards_finish(g);
    if (rc) return rc;
    g->times[2] = 0;
    g->times[0] = (float)(now_ms() - t0);
    out->n_sent = n_sent;
    out->n_tokens = total;
    out->tok_off = h_off;
    out->tokens = h_tok;
    out->eos_cost = h_eos;
    return KP_OK;
}

extern "C" int kp_shards_tokenize_gather(kp_shards* g, const uint8_t* utf8, const uint64_t* offsets, uint64_t n_sent,
                                         kp_result8* out) {
    if (!g || !offsets || !out || (n_sent && offsets[n_sent] > offsets[0] && !utf8)) return KP_ERR_ARG;
    const double t0 = now_ms();
    uint64_t total = 0;
    int rc = shards_begin(g, utf8, offsets, n_sent, &total);
    if (rc) return rc;
    KP_CUDA(cudaSetDevice(g->dev[0]->device));
    if ((rc = g->g_tok_off.ensure(sizeof(uint32_t) * (n_sent + 1))) || (rc = g->g_tokens.ensure(sizeof(kp_token8) * (total + 1))) ||
        (rc = g->g_eos.ensure(sizeof(int32_t) * (n_sent + 1))))
        return rc;
    uint32_t* g_off = (uint32_t*)g->g_tok_off.p;
    kp_token8* g_tok = (kp_token8*)g->g_tokens.p;
    int32_t* g_eos = (int32_t*)g->g_eos.p;
    // phase 2: pack, then token records to devices[0] over NVLink: grouped ncclSend / ncclRecv, exact sizes
    for (int i = 0; i < g->n; i++) {
        kp_shard_dev* d = g->dev[i];
        d->worker.post([g, d, i, g_off, g_tok, g_eos]() -> int {
            const uint32_t* d_off;
            const kp_token8* d_tok;
            const int32_t* d_eos;
            cudaStream_t st;
            int rc = kp_pass_pack(d->tok, d->tok_base, &d_off, &d_tok, &d_eos, &st);
            if (rc) return rc;
            KP_CUDA(cudaStreamSynchronize(st));      // the gather is timed on its own
            const double t1 = now_ms();
            NcclApi* N = g->nccl;
            auto n_off = [g](int k) { return (g->dev[k]->s1 - g->dev[k]->s0) + (k == g->n - 1 ? 1 : 0); };
            if (i == 0) {
                const uint64_t S = d->s1 - d->s0;
                KP_CUDA(cudaMemcpyAsync(g_off, d_off, sizeof(uint32_t) * n_off(0), cudaMemcpyDeviceToDevice, st));
                if (d->ntok) KP_CUDA(cudaMemcpyAsync(g_tok, d_tok, sizeof(kp_token8) * d->ntok, cudaMemcpyDeviceToDevice, st));
                if (S) KP_CUDA(cudaMemcpyAsync(g_eos, d_eos, sizeof(int32_t) * S, cudaMemcpyDeviceToDevice, st));
                if (g->n > 1) {
                    KP_NCCL(N, N->GroupStart());
                    for (int k = 1; k < g->n; k++) {
                        kp_shard_dev* p = g->dev[k];
                        const uint64_t Sk = p->s1 - p->s0;
                        if (n_off(k)) KP_NCCL(N, N->Recv(g_off + p->s0, n_off(k), ncclUint32, k, d->comm, st));
                        if (p->ntok) KP_NCCL(N, N->Recv(g_tok + p->tok_base, p->ntok * 8, ncclChar, k, d->comm, st));
                        if (Sk) KP_NCCL(N, N->Recv(g_eos + p->s0, Sk, ncclInt32, k, d->comm, st));
                    }
                    KP_NCCL(N, N->GroupEnd());
                }
            } else {
                const uint64_t S = d->s1 - d->s0;
                KP_NCCL(N, N->GroupStart());
                if (n_off(i)) KP_NCCL(N, N->Send(d_off, n_off(i), ncclUint32, 0, d->comm, st));
                if (d->ntok) KP_NCCL(N, N->Send(d_tok, d->ntok * 8, ncclChar, 0, d->comm, st));
                if (S) KP_NCCL(N, N->Send(d_eos, S, ncclInt32, 0, d->comm, st));
                KP_NCCL(N, N->GroupEnd());
            }
            KP_CUDA(cudaStreamSynchronize(st));
            d->gather_ms = now_ms() - t1;
            return (int)KP_OK;
        });
    }
    rc = shards_finish(g);
    if (rc) return rc;
    double gm = 0;
    for (int i = 0; i < g->n; i++) gm = std::max(gm, g->dev[i]->gather_ms);
    g->times[2] = (float)gm;
    g->times[0] = (float)(now_ms() - t0);
    out->n_sent = n_sent;
    out->n_tokens = total;
    out->tok_off = g_off;
    out->tokens = g_tok;
    out->eos_cost = g_eos;
    return KP_OK;
}

extern "C" int kp_shards_times(const kp_shards* g, float ms[4]) {
    if (!g || !ms) return KP_ERR_ARG;
    for (int i = 0; i < 4; i++) ms[i] = g->times[i];
    return KP_OK;
}

extern "C" int kp_shards_copy_to_host(kp_shards* g, void* dst, const void* device_src, uint64_t bytes) {
    if (!g || (bytes && (!dst || !device_src))) return KP_ERR_ARG;
    KP_CUDA(cudaSetDevice(g->dev[0]->device));
    if (bytes) KP_CUDA(cudaMemcpy(dst, device_src, bytes, cudaMemcpyDeviceToHost));
    return KP_OK;
}

// =====================================================================================================
// kp_gather: token gather across PROCESSES (one rank per GPU, e.g. under torchrun / mpirun)
// =====================================================================================================
// Every rank hands in its device-resident compact result; rank 0 ends up with the results of all ranks in
// device memory, in rank order (= global sentence order for contiguous shards), offsets rebased.  One grouped
// ncclSend / ncclRecv of a fixed-capacity block per rank -- no count exchange before the transfer, no host
// round trip: the counts ride in the block's header and a kernel on rank 0 compacts from them.
//   block = [n_sent, n_tok : u64 x 2][tok_off u32 x (cap_sent + 1)][eos i32 x cap_sent][tokens 8 B x cap_tok]
struct kp_gather {
    int device = 0, rank = 0, world = 1;
    NcclApi* nccl = nullptr;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint64_t cap_sent = 0, cap_tok = 0;
    size_t o_off = 16, o_eos = 0, o_tok = 0, block = 0;
    void* send = nullptr;            // this rank's block
    void* recv = nullptr;            // rank 0: world blocks
    void* g_tok_off = nullptr;       // rank 0: compacted result
    void* g_tokens = nullptr;
    void* g_eos = nullptr;
    uint64_t* d_totals = nullptr;    // rank 0: {sentences, tokens, overflow flag}
    uint64_t* h_totals = nullptr;    // pinned
    float last_ms = 0;
};

extern "C" int kp_gather_unique_id(void* id128) {
    if (!id128) return KP_ERR_ARG;
    NcclApi* N = nullptr;
    int rc = nccl_load(&N);
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    auto get = (ncclResult_t(*)(ncclUniqueId*))dlsym(N->handle, "ncclGetUniqueId");
    if (!get) {
        kp_set_error("libnccl.so.2 lacks ncclGetUniqueId");
        return KP_ERR_CUDA;
    }
    KP_NCCL(N, get((ncclUniqueId*)id128));
    return KP_OK;
}

// rank 0: tokens / offsets / costs of all blocks -> compact arrays.  One block of threads per (rank, slice).
__global__ void __launch_bounds__(256) kp_gather_compact(const unsigned char* __restrict__ recv, size_t block, uint32_t world,
                                                         size_t o_off, size_t o_eos, size_t o_tok, uint64_t cap_sent,
                                                         uint64_t cap_tok, uint32_t* __restrict__ g_off,
                                                         int32_t* __restrict__ g_eos, uint2* __restrict__ g_tok,
                                                         uint64_t* __restrict__ totals) {
    const uint32_t r = blockIdx.y;
    uint64_t sbase = 0, tbase = 0;
    bool over = false;
    for (uint32_t q = 0; q <= r && q < world; q++) {
        const uint64_t* h = (const uint64_t*)(recv + (size_t)q * block);
        if (h[0] > cap_sent || h[1] > cap_tok) over = true;
        if (q < r) {
            sbase += h[0];
            tbase += h[1];
        }
    }
    const unsigned char* b = recv + (size_t)r * block;
    const uint64_t ns = ((const uint64_t*)b)[0], nt = ((const uint64_t*)b)[1];
    if (over) {                                   // a rank's result did not fit its block: nothing is trusted
        if (threadIdx.x == 0 && blockIdx.x == 0) totals[2] = 1;
        return;
    }
    const uint32_t* off = (const uint32_t*)(b + o_off);
    const int32_t* eos = (const int32_t*)(b + o_eos);
    const uint2* tok = (const uint2*)(b + o_tok);
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < nt; i += stride) g_tok[tbase + i] = tok[i];
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < ns; i += stride) {
        g_off[sbase + i] = (uint32_t)(tbase + off[i]);
        g_eos[sbase + i] = eos[i];
    }
    if (r == world - 1 && blockIdx.x == 0 && threadIdx.x == 0) {
        g_off[sbase + ns] = (uint32_t)(tbase + nt);
        totals[0] = sbase + ns;
        totals[1] = tbase + nt;
    }
}

extern "C" void kp_gather_destroy(kp_gather* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    if (g->stream) cudaStreamSynchronize(g->stream);
    if (g->comm && g->nccl) g->nccl->CommDestroy(g->comm);
    void* bufs[] = {g->send, g->recv, g->g_tok_off, g->g_tokens, g->g_eos, g->d_totals};
    for (void* p : bufs)
        if (p) cudaFree(p);
    if (g->h_totals) cudaFreeHost(g->h_totals);
    if (g->ev0) cudaEventDestroy(g->ev0);
    if (g->ev1) cudaEventDestroy(g->ev1);
    if (g->stream) cudaStreamDestroy(g->stream);
    delete g;
}

extern "C" int kp_gather_create(int device, int rank, int world, const void* id128, uint64_t cap_sent, uint64_t cap_tok,
                                kp_gather** out) {
    if (!out || !id128 || world < 1 || rank < 0 || rank >= world || cap_sent >= (1ull << 31) || cap_tok >= (1ull << 32))
        return KP_ERR_ARG;
    *out = nullptr;
    NcclApi* N = nullptr;
    int rc = nccl_load(&N);
    if (rc) return rc;
    KP_CUDA(cudaSetDevice(device));
    kp_gather* g = new kp_gather();
    g->device = device;
    g->rank = rank;
    g->world = world;
    g->nccl = N;
    g->cap_sent = cap_sent;
    g->cap_tok = cap_tok;
    g->o_eos = g->o_off + 4 * (cap_sent + 1);
    g->o_tok = (g->o_eos + 4 * cap_sent + 15) & ~size_t(15);
    g->block = (g->o_tok + 8 * cap_tok + 255) & ~size_t(255);
    auto init = (ncclResult_t(*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(N->handle, "ncclCommInitRank");
    cudaError_t e = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&g->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&g->ev1);
    if (e == cudaSuccess) e = cudaMalloc(&g->send, g->block);
    if (e == cudaSuccess) e = cudaMemset(g->send, 0, g->block);
    if (e == cudaSuccess && rank == 0) {
        e = cudaMalloc(&g->recv, g->block * (size_t)world);
        if (e == cudaSuccess) e = cudaMalloc(&g->g_tok_off, 4 * ((size_t)cap_sent * world + 1));
        if (e == cudaSuccess) e = cudaMalloc(&g->g_eos, 4 * ((size_t)cap_sent * world + 1));
        if (e == cudaSuccess) e = cudaMalloc(&g->g_tokens, 8 * ((size_t)cap_tok * world + 1));
        if (e == cudaSuccess) e = cudaMalloc((void**)&g->d_totals, 64);
        if (e == cudaSuccess) e = cudaHostAlloc((void**)&g->h_totals, 64, cudaHostAllocDefault);
    }
    if (e != cudaSuccess || !init) {
        kp_set_error("kp_gather_create: %s", init ? cudaGetErrorString(e) : "libnccl.so.2 lacks ncclCommInitRank");
        cudaGetLastError();
        kp_gather_destroy(g);
        return init ? KP_ERR_NOMEM : KP_ERR_CUDA;
    }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t nr = init(&g->comm, world, id, rank);
    if (nr != ncclSuccess) {
        kp_set_error("ncclCommInitRank failed: %s", N->GetErrorString(nr));
        g->comm = nullptr;
        kp_gather_destroy(g);
        return KP_ERR_CUDA;
    }
    {   // the block layout follows from the capacities: every rank must have passed the same ones
        int64_t h[4] = {(int64_t)cap_sent, (int64_t)cap_tok, -(int64_t)cap_sent, -(int64_t)cap_tok};
        int64_t* dv = nullptr;
        cudaError_t e2 = cudaMalloc((void**)&dv, sizeof(h));
        if (e2 == cudaSuccess) e2 = cudaMemcpy(dv, h, sizeof(h), cudaMemcpyHostToDevice);
        ncclResult_t r2 = ncclSuccess;
        if (e2 == cudaSuccess) r2 = N->AllReduce(dv, dv, 4, ncclInt64, ncclMax, g->comm, g->stream);
        if (e2 == cudaSuccess && r2 == ncclSuccess) e2 = cudaStreamSynchronize(g->stream);
        if (e2 == cudaSuccess && r2 == ncclSuccess) e2 = cudaMemcpy(h, dv, sizeof(h), cudaMemcpyDeviceToHost);
        if (dv) cudaFree(dv);
        if (e2 != cudaSuccess || r2 != ncclSuccess || h[0] != -h[2] || h[1] != -h[3]) {
            if (e2 != cudaSuccess || r2 != ncclSuccess) kp_set_error("kp_gather_create: capacity check failed");
            else kp_set_error("kp_gather_create: ranks disagree on the capacities (%lld..%lld sentences, %lld..%lld tokens)",
                              (long long)-h[2], (long long)h[0], (long long)-h[3], (long long)h[1]);
            kp_gather_destroy(g);
            return e2 != cudaSuccess || r2 != ncclSuccess ? KP_ERR_CUDA : KP_ERR_ARG;
        }
    }
    *out = g;
    return KP_OK;
}

extern "C" int kp_gather_tokens(kp_gather* g, const kp_result8* mine, kp_result8* out) {
    if (!g || !mine || !out) return KP_ERR_ARG;
    if (mine->n_sent > g->cap_sent || mine->n_tokens > g->cap_tok) {
        kp_set_error("shard (%llu sentences, %llu tokens) exceeds the gather capacity (%llu, %llu)",
                     (unsigned long long)mine->n_sent, (unsigned long long)mine->n_tokens, (unsigned long long)g->cap_sent,
                     (unsigned long long)g->cap_tok);
        return KP_ERR_TOO_LARGE;
    }
    KP_CUDA(cudaSetDevice(g->device));
    cudaStream_t st = g->stream;
    NcclApi* N = g->nccl;
    char* s = (char*)g->send;
    const uint64_t hdr[2] = {mine->n_sent, mine->n_tokens};
    KP_CUDA(cudaEventRecord(g->ev0, st));
    // this rank's block: header, offsets, costs, token records (the caller's result is complete: its call synchronised)
    KP_CUDA(cudaMemcpyAsync(s, hdr, 16, cudaMemcpyHostToDevice, st));
    KP_CUDA(cudaMemcpyAsync(s + g->o_off, mine->tok_off, 4 * (mine->n_sent + 1), cudaMemcpyDeviceToDevice, st));
    if (mine->n_sent) KP_CUDA(cudaMemcpyAsync(s + g->o_eos, mine->eos_cost, 4 * mine->n_sent, cudaMemcpyDeviceToDevice, st));
    if (mine->n_tokens) KP_CUDA(cudaMemcpyAsync(s + g->o_tok, mine->tokens, 8 * mine->n_tokens, cudaMemcpyDeviceToDevice, st));
    if (g->rank == 0) {
        KP_CUDA(cudaMemcpyAsync(g->recv, s, g->block, cudaMemcpyDeviceToDevice, st));
        if (g->world > 1) {
            KP_NCCL(N, N->GroupStart());
            for (int r = 1; r < g->world; r++)
                KP_NCCL(N, N->Recv((char*)g->recv + (size_t)r * g->block, g->block, ncclChar, r, g->comm, st));
            KP_NCCL(N, N->GroupEnd());
        }
        KP_CUDA(cudaMemsetAsync(g->d_totals, 0, 64, st));
        dim3 grid(64, (unsigned)g->world);
        kp_gather_compact<<<grid, 256, 0, st>>>((const unsigned char*)g->recv, g->block, (uint32_t)g->world, g->o_off, g->o_eos,
                                                g->o_tok, g->cap_sent, g->cap_tok, (uint32_t*)g->g_tok_off, (int32_t*)g->g_eos,
                                                (uint2*)g->g_tokens, g->d_totals);
        KP_CUDA(cudaGetLastError());
        KP_CUDA(cudaMemcpyAsync(g->h_totals, g->d_totals, 24, cudaMemcpyDeviceToHost, st));
    } else {
        KP_NCCL(N, N->Send(s, g->block, ncclChar, 0, g->comm, st));
    }
    KP_CUDA(cudaEventRecord(g->ev1, st));
    KP_CUDA(cudaEventSynchronize(g->ev1));
    cudaEventElapsedTime(&g->last_ms, g->ev0, g->ev1);
    memset(out, 0, sizeof(*out));
    if (g->rank == 0) {
        if (g->h_totals[2]) {
            kp_set_error("a rank's result exceeded the gather capacity");
            return KP_ERR_TOO_LARGE;
        }
        out->n_sent = g->h_totals[0];
        out->n_tokens = g->h_totals[1];
        out->tok_off = (const uint32_t*)g->g_tok_off;
        out->tokens = (const kp_token8*)g->g_tokens;
        out->eos_cost = (const int32_t*)g->g_eos;
    }
    return KP_OK;
}

extern "C" int kp_gather_last_ms(const kp_gather* g, float* ms) {
    if (!g || !ms) return KP_ERR_ARG;
    *ms = g->last_ms;
    return KP_OK;
}

extern "C" int kp_gather_copy_to_host(kp_gather* g, void* dst, const void* device_src, uint64_t bytes) {
    if (!g || (bytes && (!dst || !device_src))) return KP_ERR_ARG;
    KP_CUDA(cudaSetDevice(g->device));
    if (bytes) KP_CUDA(cudaMemcpy(dst, device_src, bytes, cudaMemcpyDeviceToHost));
    return KP_OK;
}
