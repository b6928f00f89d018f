// Thin runtime layer under the C ABI: device memory, copies, stream sync.  CUDA in the product; plain host memory
// when the sources are compiled for the test-only emulator (ZK_EMU).
#pragma once
#include <sched.h>
#include "zk_platform.cuh"
#include <ctime>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace zk {
namespace rt {

struct error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#ifdef ZK_EMU
inline void check(int, const char *) {}
inline void set_device(int) {}
inline void bind(int, zk_stream_t, zk_stream_t = nullptr, zk_stream_t = nullptr) {}
inline void unbind() {}
inline int device_count() { return 1; }
inline void *dmalloc(size_t bytes) {
    void *p = malloc(bytes ? bytes : 1);
    if (!p) throw error("emu: out of memory");
    return p;
}
inline void dfree(void *p) { free(p); }
inline void *hmalloc_pinned(size_t bytes) { return malloc(bytes ? bytes : 1); }
inline void hfree_pinned(void *p) { free(p); }
inline void *hmalloc_mapped(size_t bytes) { return calloc(1, bytes ? bytes : 1); }
inline void *mapped_device_ptr(void *h) { return h; }
inline void host_pin(void *, size_t) {}
inline const void *host_device_ptr(const void *p) { return p; }
inline void host_unpin(void *) {}
inline void h2d(void *dst, const void *src, size_t n, zk_stream_t) { memcpy(dst, src, n); }
inline void d2h(void *dst, const void *src, size_t n, zk_stream_t) { memcpy(dst, src, n); }
inline void d2d(void *dst, const void *src, size_t n, zk_stream_t) { memmove(dst, src, n); }
inline void dzero(void *dst, size_t n, zk_stream_t) { memset(dst, 0, n); }
inline zk_stream_t stream_create(bool = true) { return nullptr; }
inline void stream_destroy(zk_stream_t) {}
inline void sync(zk_stream_t) {}
inline void stream_wait_stream(zk_stream_t, zk_stream_t) {}
inline void check_launch(const char *) {}
typedef double *event_t;   // wall-clock stamp taken at "record" time (launches are synchronous in the emulator)
inline double emu_now_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
inline event_t event_create() { return new double(0); }
inline void event_destroy(event_t e) { delete e; }
inline void event_record(event_t e, zk_stream_t) { *e = emu_now_ms(); }
inline void event_sync(event_t) {}
inline float event_elapsed_ms(event_t a, event_t b) { return (float) (*b - *a); }
#else
inline void check(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw error(std::string(what) + ": " + cudaGetErrorString(e));
}
inline void set_device(int d) { check(cudaSetDevice(d), "cudaSetDevice"); }
inline int device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
// Device memory comes from a small caching pool: cudaFree (and often cudaMalloc) synchronises the whole device, which
// would stall the witness copy that runs on a second stream during a proof, and the per-call scratch buffers of the API
// would pay hundreds of microseconds each.  Freed blocks are kept (up to kPoolLimit bytes) and handed out again to
// requests of at least half their size.
// The pool is shared by every context of the process (several proofs may be in flight on one device, each context with its
// own streams and host thread), so a freed block carries one event per stream of the context that freed it (bind()), recorded
// at the time of the free: the block is only handed out again once those events have completed, i.e. once no kernel or copy
// queued by its previous owner can still touch it.
struct pool_block_t {
    void *p;
    std::vector<cudaEvent_t> fence;
};
struct pool_t {
    std::mutex m;
    std::multimap<std::pair<int, size_t>, pool_block_t> free_blocks;   // (device, bytes) -> block
    std::unordered_map<void *, std::pair<int, size_t>> live;           // block -> (device, bytes)
    size_t cached = 0;
};
inline pool_t &pool() { static pool_t *P = new pool_t; return *P; }   // leaked on purpose: no destruction-order problems at exit
constexpr size_t kPoolLimit = 48ull << 30;
// the streams of the context the calling thread is working for (set at every C-ABI entry)
struct bound_t { zk_stream_t s[3] = {nullptr, nullptr, nullptr}; };
inline bound_t &bound() { static thread_local bound_t b; return b; }
inline void bind(int device, zk_stream_t s0, zk_stream_t s1 = nullptr, zk_stream_t s2 = nullptr) {
    set_device(device);
    bound_t &b = bound();
    b.s[0] = s0; b.s[1] = s1; b.s[2] = s2;
}
inline void unbind() { bound() = bound_t(); }
inline bool fence_done(pool_block_t &b) {
    for (cudaEvent_t e : b.fence) {
        const cudaError_t q = cudaEventQuery(e);
        if (q == cudaErrorNotReady) return false;
        if (q != cudaSuccess) cudaGetLastError();   // a failed stream: nothing of it will run any more
    }
    for (cudaEvent_t e : b.fence) cudaEventDestroy(e);
    b.fence.clear();
    return true;
}
inline void *dmalloc(size_t bytes) {
    bytes = ((bytes ? bytes : 1) + 511) & ~(size_t) 511;
    int dev = 0;
    check(cudaGetDevice(&dev), "cudaGetDevice");
    pool_t &P = pool();
    {
        std::lock_guard<std::mutex> g(P.m);
        for (auto it = P.free_blocks.lower_bound({dev, bytes}); it != P.free_blocks.end() && it->first.first == dev && it->first.second <= 2 * bytes; ++it) {
            if (!fence_done(it->second)) continue;   // its previous owner may still be using it: look at the next candidate
            void *p = it->second.p;
            P.cached -= it->first.second;
            P.live[p] = it->first;
            P.free_blocks.erase(it);
            return p;
        }
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) {   // give the cache back and try once more
        cudaGetLastError();
        cudaDeviceSynchronize();
        std::lock_guard<std::mutex> g(P.m);
        for (auto &kv : P.free_blocks) {
            for (cudaEvent_t ev : kv.second.fence) cudaEventDestroy(ev);
            cudaFree(kv.second.p);
        }
        P.free_blocks.clear();
        P.cached = 0;
        e = cudaMalloc(&p, bytes);
    }
    check(e, "cudaMalloc");
    std::lock_guard<std::mutex> g(P.m);
    P.live[p] = {dev, bytes};
    return p;
}
inline void dfree(void *p) {
    if (!p) return;
    pool_t &P = pool();
    std::pair<int, size_t> key;
    {
        std::lock_guard<std::mutex> g(P.m);
        auto it = P.live.find(p);
        if (it == P.live.end()) { cudaFree(p); return; }
        key = it->second;
        P.live.erase(it);
        if (P.cached + key.second > kPoolLimit) { cudaFree(p); return; }
    }
    pool_block_t b;
    b.p = p;
    for (zk_stream_t s : bound().s)
        if (s) {
            cudaEvent_t e;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(e, s) != cudaSuccess) { cudaGetLastError(); continue; }
            b.fence.push_back(e);
        }
    std::lock_guard<std::mutex> g(P.m);
    P.free_blocks.emplace(key, std::move(b));
    P.cached += key.second;
}
inline void *hmalloc_pinned(size_t bytes) {
    void *p = nullptr;
    check(cudaMallocHost(&p, bytes ? bytes : 1), "cudaMallocHost");
    return p;
}
inline void hfree_pinned(void *p) { if (p) cudaFreeHost(p); }
// pinned host memory the device can write directly (zero-copy result mailbox)
inline void *hmalloc_mapped(size_t bytes) {
    void *p = nullptr;
    check(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocMapped | cudaHostAllocPortable), "cudaHostAlloc(mapped)");
    memset(p, 0, bytes);
    return p;
}
inline void *mapped_device_ptr(void *h) {
    void *d = nullptr;
    check(cudaHostGetDevicePointer(&d, h, 0), "cudaHostGetDevicePointer");
    return d;
}
inline void host_pin(void *p, size_t bytes) { check(cudaHostRegister(p, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable), "cudaHostRegister"); }
// device-side address of page-locked, mapped host memory (nullptr if the range is not mapped)
inline const void *host_device_ptr(const void *p) {
    void *d = nullptr;
    if (cudaHostGetDevicePointer(&d, const_cast<void *>(p), 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return d;
}
inline void host_unpin(void *p) { check(cudaHostUnregister(p), "cudaHostUnregister"); }
inline void h2d(void *dst, const void *src, size_t n, zk_stream_t s) { if (n) check(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, s), "h2d"); }
inline void d2h(void *dst, const void *src, size_t n, zk_stream_t s) { if (n) check(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, s), "d2h"); }
inline void d2d(void *dst, const void *src, size_t n, zk_stream_t s) { if (n) check(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, s), "d2d"); }
inline void dzero(void *dst, size_t n, zk_stream_t s) { if (n) check(cudaMemsetAsync(dst, 0, n, s), "memset"); }
// high = true: the proof stream (greatest priority); false: background copies (least priority)
inline zk_stream_t stream_create(bool high = true) {
    cudaStream_t s;
    int lo = 0, hi = 0;
    check(cudaDeviceGetStreamPriorityRange(&lo, &hi), "cudaDeviceGetStreamPriorityRange");
    check(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high ? hi : lo), "cudaStreamCreate");
    return s;
}
inline void stream_destroy(zk_stream_t s) { if (s) cudaStreamDestroy(s); }
// Waiting for a stream.  cudaStreamSynchronize spins on a host core, which is the lowest latency while every prover thread has a core
// of its own.  With several provers per GPU and several GPUs per box the waiting threads outnumber the cores, so the wait can give
// the core away:  ZK_HOST_WAIT=spin (default) | yield (poll an event, sched_yield between polls: as fast as spinning on an idle host,
// lets runnable threads in when there are none to spare) | block (sleep on a blocking-sync event: no core burnt, a few tens of
// microseconds more per wait).  ZK_BLOCKING_SYNC=1 is the older spelling of "block".
enum wait_mode_t { kWaitSpin = 0, kWaitYield = 1, kWaitBlock = 2 };
inline wait_mode_t wait_mode() {
    static const wait_mode_t m = [] {
        const char *e = getenv("ZK_HOST_WAIT");
        if (e && !strcmp(e, "yield")) return kWaitYield;
        if (e && !strcmp(e, "block")) return kWaitBlock;
        if (e && *e && strcmp(e, "spin")) fprintf(stderr, "zkcnn_b200: ZK_HOST_WAIT=%s not understood (spin | yield | block); spinning\n", e);
        const char *b = getenv("ZK_BLOCKING_SYNC");
        return (!e || !*e) && b && *b && *b != '0' ? kWaitBlock : kWaitSpin;
    }();
    return m;
}
inline void sync(zk_stream_t s) {
    const wait_mode_t mode = wait_mode();
    if (mode == kWaitSpin) { check(cudaStreamSynchronize(s), "cudaStreamSynchronize"); return; }
    static thread_local cudaEvent_t ev = nullptr;
    static thread_local int ev_dev = -1;
    int dev = 0;
    check(cudaGetDevice(&dev), "cudaGetDevice");
    if (!ev || ev_dev != dev) {
        check(cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming), "cudaEventCreate");
        ev_dev = dev;
    }
    check(cudaEventRecord(ev, s), "cudaEventRecord");
    if (mode == kWaitYield) {
        cudaError_t q;
        while ((q = cudaEventQuery(ev)) == cudaErrorNotReady) sched_yield();
        check(q, "cudaEventQuery");
        return;
    }
    check(cudaEventSynchronize(ev), "cudaEventSynchronize");
}
// make stream `s` wait (on the device) for everything queued on `other` so far
inline void stream_wait_stream(zk_stream_t s, zk_stream_t other) {
    cudaEvent_t e;
    check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    check(cudaEventRecord(e, other), "cudaEventRecord");
    check(cudaStreamWaitEvent(s, e, 0), "cudaStreamWaitEvent");
    cudaEventDestroy(e);
}
inline void check_launch(const char *what) { check(cudaGetLastError(), what); }
typedef cudaEvent_t event_t;
inline event_t event_create() { cudaEvent_t e; check(cudaEventCreate(&e), "cudaEventCreate"); return e; }
inline void event_destroy(event_t e) { cudaEventDestroy(e); }
inline void event_record(event_t e, zk_stream_t s) { check(cudaEventRecord(e, s), "cudaEventRecord"); }
inline void event_sync(event_t e) { check(cudaEventSynchronize(e), "cudaEventSynchronize"); }
inline float event_elapsed_ms(event_t a, event_t b) { float ms = 0; check(cudaEventElapsedTime(&ms, a, b), "cudaEventElapsedTime"); return ms; }
#endif

// grow-only device buffer
struct dbuf {
    void *p = nullptr;
    size_t cap = 0;
    dbuf() = default;
    dbuf(const dbuf &) = delete;
    dbuf &operator=(const dbuf &) = delete;
    dbuf(dbuf &&o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
    dbuf &operator=(dbuf &&o) noexcept {
        if (this != &o) { dfree(p); p = o.p; cap = o.cap; o.p = nullptr; o.cap = 0; }
        return *this;
    }
    ~dbuf() { dfree(p); }
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        dfree(p);
        p = nullptr;
        cap = 0;
        p = dmalloc(bytes);
        cap = bytes;
    }
    // like ensure() but keeps the old contents (device-to-device copy)
    void ensure_keep(size_t bytes, zk_stream_t s) {
        if (bytes <= cap) return;
        void *q = dmalloc(bytes);
        if (p && cap) { d2d(q, p, cap, s); sync(s); }
        dfree(p);
        p = q;
        cap = bytes;
    }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

}  // namespace rt
}  // namespace zk
