// DEVELOPER / TEST HARNESS ONLY -- not part of the product, never loaded by zkcnn_b200.
//
// A minimal CUDA execution-model emulator that lets the *same kernel sources* under zkcnn_b200/csrc be compiled
// with g++ (-DZK_EMU) and executed on host threads, so that kernel logic and the prover state machine can be
// checked against the oracle in the `-m "not gpu"` suite before a B200 is available.  One OS thread runs one
// CTA at a time; the CTA's CUDA threads are fibers (hand-rolled x86-64 context switch) that yield at
// __syncthreads().  Warp shuffles are emulated with a CTA-wide exchange buffer, which is valid for kernels whose
// threads execute shuffles uniformly (all kernels here do).
//
// The product library (libzkcnn_b200.so, built by nvcc) contains none of this: see zk_platform.cuh.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static thread_local
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __align__(n) alignas(n)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 { uint32_t x, y, z, w; };
struct ulonglong2 { unsigned long long x, y; };
struct ulonglong4 { unsigned long long x, y, z, w; };

namespace zkemu {
extern thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
void sync_threads();
void *dyn_smem();
uint64_t *exchange_buf();  // blockDim.x 64-bit slots
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
}  // namespace zkemu

namespace zkemu {
// evaluates the launch arguments once (like a real launch copies them into parameter space), then runs the grid
template <class... KArgs, class... Args>
inline void launch_k(dim3 grid, dim3 block, size_t smem, void (*kernel)(KArgs...), Args &&...args) {
    std::tuple<std::decay_t<KArgs>...> params(std::forward<Args>(args)...);
    launch(grid, block, smem, [&]() { std::apply(kernel, params); });
}
}  // namespace zkemu

#define threadIdx (zkemu::t_threadIdx)
#define blockIdx (zkemu::t_blockIdx)
#define blockDim (zkemu::t_blockDim)
#define gridDim (zkemu::t_gridDim)

static inline void __syncthreads() { zkemu::sync_threads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { zkemu::sync_threads(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <class T> static inline T zkemu_shfl(T v, unsigned src_tid) {
    static_assert(sizeof(T) <= 8, "shuffle of <= 64-bit values only");
    uint64_t *buf = zkemu::exchange_buf();
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    buf[threadIdx.x] = raw;
    zkemu::sync_threads();
    uint64_t got = buf[src_tid];
    zkemu::sync_threads();
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
    unsigned lane = threadIdx.x % width;
    unsigned src = lane + delta < (unsigned) width ? threadIdx.x + delta : threadIdx.x;
    if (src >= blockDim.x) src = threadIdx.x;
    return zkemu_shfl(v, src);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, unsigned m, int width = 32) {
    unsigned src = (threadIdx.x & ~(unsigned) (width - 1)) | ((threadIdx.x % width) ^ m);
    if (src >= blockDim.x) src = threadIdx.x;
    return zkemu_shfl(v, src);
}
template <class T> static inline T __shfl_sync(unsigned, T v, int srcLane, int width = 32) {
    unsigned src = (threadIdx.x & ~(unsigned) (width - 1)) | ((unsigned) srcLane % width);
    if (src >= blockDim.x) src = threadIdx.x;
    return zkemu_shfl(v, src);
}

// redux.sync: sum over the (full) warp of the calling thread
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    uint64_t *buf = zkemu::exchange_buf();
    buf[threadIdx.x] = v;
    zkemu::sync_threads();
    const unsigned w0 = threadIdx.x & ~31u;
    unsigned s = 0;
    for (unsigned t = w0; t < w0 + 32 && t < blockDim.x; ++t)
        if (mask >> (t - w0) & 1u) s += (unsigned) buf[t];
    zkemu::sync_threads();
    return s;
}
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicMax(unsigned *p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline int atomicMax(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
static inline unsigned __brev(unsigned x) {
    x = (x >> 16) | (x << 16);
    x = ((x & 0xff00ff00u) >> 8) | ((x & 0x00ff00ffu) << 8);
    x = ((x & 0xf0f0f0f0u) >> 4) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x & 0xccccccccu) >> 2) | ((x & 0x33333333u) << 2);
    return ((x & 0xaaaaaaaau) >> 1) | ((x & 0x55555555u) << 1);
}
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned) (((uint64_t) a * b) >> 32); }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
