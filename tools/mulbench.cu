// Throughput of the Montgomery multiplier variants on the device (run on the B200 through gpurun):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I zkcnn_b200/csrc tools/mulbench.cu -o gpurun_out/mulbench
// Each thread runs CHAINS independent chains of ITERS dependent multiplications; the result is checked against the portable
// multiplier on the host for a few threads.  Prints multiplications per second and the per-SM-cycle rate.
#define ZK_INLINE_FIELD_MUL 1
#include "mont.cuh"
#include <vector>

using namespace zk;

template <class F, int VARIANT> __device__ __forceinline__ F mulv(const F &a, const F &b) {
    F r;
    if (VARIANT == 0) F::mul_ptx(r.v, a.v, b.v);
    else F::mul_wide(r.v, a.v, b.v);
    return r;
}
template <class F, int VARIANT> __device__ __noinline__ F mulv_call(F a, F b) { return mulv<F, VARIANT>(a, b); }

template <class F, int VARIANT, int CHAINS, bool CALL> __global__ void __launch_bounds__(256) k_chain(const F *in, F *out, int iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    F x[CHAINS], y = in[t];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = in[t + c + 1];
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = CALL ? mulv_call<F, VARIANT>(x[c], y) : mulv<F, VARIANT>(x[c], y);
    }
    F s = x[0];
#pragma unroll
    for (int c = 1; c < CHAINS; ++c) s = s + x[c];
    out[t] = s;
}

template <class F> static void fill(std::vector<F> &v, uint64_t seed) {
    for (auto &e : v) {
        for (int i = 0; i < F::N; ++i) {
            seed ^= seed >> 12; seed ^= seed << 25; seed ^= seed >> 27;
            e.v[i] = (uint32_t) ((seed * 0x2545F4914F6CDD1DULL) >> 32);
        }
        e.v[F::N - 1] &= 0x0fffffffu;   // < modulus
    }
}

template <class F, int VARIANT, int CHAINS, bool CALL> static void run(const char *name, int blocks_per_sm, int iters) {
    const int blocks = 148 * blocks_per_sm, threads = 256, n = blocks * threads;
    std::vector<F> h(n + CHAINS + 1), o(n);
    fill(h, 0x9E3779B97F4A7C15ULL);
    F *din, *dout;
    cudaMalloc(&din, h.size() * sizeof(F));
    cudaMalloc(&dout, n * sizeof(F));
    cudaMemcpy(din, h.data(), h.size() * sizeof(F), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_chain<F, VARIANT, CHAINS, CALL><<<blocks, threads>>>(din, dout, 8);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_chain<F, VARIANT, CHAINS, CALL><<<blocks, threads>>>(din, dout, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    cudaMemcpy(o.data(), dout, n * sizeof(F), cudaMemcpyDeviceToHost);
    // host check on 3 threads
    int bad = 0;
    for (int t : {0, 12345 % n, n - 1}) {
        F x[CHAINS], y = h[t];
        for (int c = 0; c < CHAINS; ++c) x[c] = h[t + c + 1];
        for (int i = 0; i < iters; ++i)
            for (int c = 0; c < CHAINS; ++c) { F r; F::mul_portable(r.v, x[c].v, y.v); x[c] = r; }
        F s = x[0];
        for (int c = 1; c < CHAINS; ++c) s = s + x[c];
        if (!(s == o[t])) ++bad;
    }
    const double muls = (double) n * CHAINS * iters;
    printf("%-34s blocks/SM %d  %8.3f ms  %7.2f Gmul/s  %6.3f mul/clk/SM (1.9GHz)  %s %s\n", name, blocks_per_sm, best, muls / best / 1e6,
           muls / (best * 1e-3) / 148 / 1.9e9, bad ? "MISMATCH" : "ok", err ? cudaGetErrorString(err) : "");
    cudaFree(din); cudaFree(dout);
}

int main() {
    const int it = 2000;
    for (int bps : {2, 4, 8}) {
        run<fr_t, 0, 1, false>("fr cios   inline chains=1", bps, it);
        run<fr_t, 1, 1, false>("fr wide   inline chains=1", bps, it);
        run<fr_t, 0, 2, false>("fr cios   inline chains=2", bps, it);
        run<fr_t, 1, 2, false>("fr wide   inline chains=2", bps, it);
        run<fr_t, 0, 2, true>("fr cios   call   chains=2", bps, it);
        run<fr_t, 1, 2, true>("fr wide   call   chains=2", bps, it);
        run<fp_t, 0, 1, false>("fp cios   inline chains=1", bps, it / 2);
        run<fp_t, 1, 1, false>("fp wide   inline chains=1", bps, it / 2);
        run<fp_t, 0, 2, true>("fp cios   call   chains=2", bps, it / 2);
        run<fp_t, 1, 2, true>("fp wide   call   chains=2", bps, it / 2);
    }
    return 0;
}
