// Latency of the building blocks of a sumcheck round kernel, measured with clock64() by ONE warp on an otherwise idle SM
// (run on the B200 through gpurun):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I zkcnn_b200/csrc tools/latbench.cu -o gpurun_out/latbench
#include "mont.cuh"
#include <cstdio>
#include <vector>

using namespace zk;

__global__ void k_lat(const fr_t *in, fr_t *out, long long *cyc) {
    const int lane = threadIdx.x & 31;
    fr_t a = in[threadIdx.x], b = in[threadIdx.x + 32], c;
    long long t0, t1;
    int slot = 0;
#define TIC() do { __syncwarp(); t0 = clock64(); } while (0)
#define TOC() do { __syncwarp(); t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = t1 - t0; ++slot; } while (0)
    // 0: one out-of-line Montgomery multiplication (dependent chain of 4 -> / 4)
    TIC();
    c = a * b; c = c * b; c = c * b; c = c * b;
    TOC();
    // 1: field add + sub (4 of each)
    TIC();
    c = c + a; c = c - b; c = c + a; c = c - b; c = c + a; c = c - b; c = c + a; c = c - b;
    TOC();
    // 2: unreduced product + accumulate (4 macs)
    fr_lazy_t acc;
    acc.clear();
    TIC();
    acc.mac(c, a); acc.mac(c, b); acc.mac(a, b); acc.mac(a, a);
    TOC();
    // 3: 34 full-mask redux
    uint32_t s = 0;
    TIC();
#pragma unroll
    for (int k = 0; k < 17; ++k) {
        s += __reduce_add_sync(0xffffffffu, acc.w[k] & 0xffffu);
        s += __reduce_add_sync(0xffffffffu, acc.w[k] >> 16);
    }
    TOC();
    // 4: 34 role-masked redux (4 groups)
    const uint32_t mask = 0x11111111u << (lane & 3);
    TIC();
#pragma unroll
    for (int k = 0; k < 17; ++k) {
        s += __reduce_add_sync(mask, acc.w[k] & 0xffffu);
        s += __reduce_add_sync(mask, acc.w[k] >> 16);
    }
    TOC();
    // 5: 24 shuffles (3 field elements)
    fr_t o;
    TIC();
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = __shfl_xor_sync(0xffffffffu, c.v[j], 1);
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] ^= __shfl_xor_sync(0xffffffffu, a.v[j], 2);
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] ^= __shfl_xor_sync(0xffffffffu, b.v[j], 2);
    TOC();
    // 6: warp sum of one field element by 5 shuffle levels (8 shuffles + 1 add per level)
    fr_t w = c;
    TIC();
    for (int d = 16; d > 0; d >>= 1) {
        fr_t p;
#pragma unroll
        for (int j = 0; j < 8; ++j) p.v[j] = __shfl_xor_sync(0xffffffffu, w.v[j], d);
        w = w + p;
    }
    TOC();
    // 7: global load of 2 entries (cold) -> dependent use
    TIC();
    fr_t g = in[64 + threadIdx.x * 4];
    c = c + g;
    TOC();
    // 8: __threadfence
    TIC();
    __threadfence();
    TOC();
    // 9: atomicAdd with return (ticket)
    TIC();
    if (lane == 0) s += atomicAdd(reinterpret_cast<unsigned *>(cyc + 60), 1u);
    TOC();
    // 10: __threadfence_system + store
    TIC();
    __threadfence_system();
    TOC();
    // 11: __syncthreads (single warp)
    TIC();
    __syncthreads();
    TOC();
    // 12: 4 dependent Fp multiplications; 13: 4 dependent PAIRS of Fp multiplications (8 products)
    fp_t fa, fb;
#pragma unroll
    for (int j = 0; j < 12; ++j) { fa.v[j] = a.v[j & 7] ^ (j * 77u); fb.v[j] = b.v[j & 7] + j; }
    fa.v[11] &= 0x0fffffffu; fb.v[11] &= 0x0fffffffu;
    TIC();
    fa = fa * fb; fa = fa * fb; fa = fa * fb; fa = fa * fb;
    TOC();
    fp_t fc = fb;
    TIC();
    {
        fp_t::pair_t q = fp_t::mul2(fa, fb, fc, fb);
        q = fp_t::mul2(q.a, fb, q.b, fa);
        q = fp_t::mul2(q.a, fb, q.b, fa);
        q = fp_t::mul2(q.a, fb, q.b, fa);
        fa = q.a; fc = q.b;
    }
    TOC();
    // 14: 4 dependent pairs of Fr multiplications
    fr_t ra = a, rb = b;
    TIC();
    {
        fr_t::pair_t q = fr_t::mul2(ra, b, rb, a);
        q = fr_t::mul2(q.a, b, q.b, a);
        q = fr_t::mul2(q.a, b, q.b, a);
        q = fr_t::mul2(q.a, b, q.b, a);
        ra = q.a; rb = q.b;
    }
    TOC();
    if (fa.v[0] == 0x12345u && fc.v[3] == 7u) out[1] = ra + rb;
    out[threadIdx.x] = c + o + w;
    if (s == 0x12345678u) out[0] = a;
}

int main() {
    std::vector<fr_t> h(4096);
    uint64_t seed = 88172645463325252ULL;
    for (auto &e : h) {
        for (int i = 0; i < 8; ++i) {
            seed ^= seed >> 12; seed ^= seed << 25; seed ^= seed >> 27;
            e.v[i] = (uint32_t) ((seed * 0x2545F4914F6CDD1DULL) >> 32);
        }
        e.v[7] &= 0x3fffffffu;
    }
    fr_t *din, *dout;
    long long *dc;
    cudaMalloc(&din, h.size() * sizeof(fr_t));
    cudaMalloc(&dout, 64 * sizeof(fr_t));
    cudaMalloc(&dc, 64 * sizeof(long long));
    cudaMemcpy(din, h.data(), h.size() * sizeof(fr_t), cudaMemcpyHostToDevice);
    const char *names[] = {"4 x mul_call (dependent)", "8 x field add/sub", "4 x lazy mac", "34 x redux full mask", "34 x redux role masks",
                           "24 x shfl", "warp sum by shuffles (1 fr)", "global load (cold) + use", "__threadfence", "atomicAdd ticket",
                           "__threadfence_system", "__syncthreads", "4 x Fp mul_call (dependent)", "4 x Fp mul2 pairs (dependent)",
                           "4 x Fr mul2 pairs (dependent)"};
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(dc, 0, 64 * sizeof(long long));
        k_lat<<<1, 32>>>(din, dout, dc);
        cudaDeviceSynchronize();
        long long c[64];
        cudaMemcpy(c, dc, sizeof c, cudaMemcpyDeviceToHost);
        if (rep == 0) continue;
        printf("rep %d\n", rep);
        for (int i = 0; i < 15; ++i) printf("  %-32s %8lld clk\n", names[i], c[i]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
