// Witness generation on the device: the evaluation of the layered circuit that neuralNetwork::create
// does on the host while it builds the layers (src/neuralNetwork.cpp:918-965 calcNormalLayer / calcDotProdLayer / calcFFTLayer,
// the number-theoretic transform of src/utils.cpp:105-145 and the bit-decomposition helpers :899-916).  The topology and the
// quantised weights stay resident; a new picture costs one small upload, these kernels, and nothing over PCIe.
//
//   k_eval_items      out[g] = sum over the gates of g of  val[.][u] (* val[.][v]) * two_mul[sc]      (calcNormalLayer)
//   k_dotprod_eval    out[g block] += src[u block] .* src[v block]                                    (calcDotProdLayer)
//   k_ntt_blocks      forward / inverse radix-2 NTT of every 2^n block of a layer, in shared memory   (calcFFTLayer + fft)
//   k_aux_bits        sign / magnitude bits of earlier gate values -> auxiliary inputs in val[0]      (prepareSignBit, prepareDecmpBit)
//   k_aux_max*        running maximum of the (ReLU-ed) window elements                                (prepareMax)
//   k_layer_range     largest positive value and largest magnitude of a negative one                  (getNextBit, :967-977)
#pragma once
#include "cubic_kernels.cuh"

namespace zk {

// mcl's getInt64 on the device: values >= (r+1)/2 stand for x - r.  *mag = |x| (low 64 bits), returns the sign.
__device__ __forceinline__ bool fr_sign_magnitude(const fr_t &x, unsigned long long *mag) {
    uint32_t c[8];
    x.to_canonical(c);
    const bool neg = fr_t::ge_raw(c, fr_cfg::half());
    if (neg) {
        const uint32_t *p = fr_cfg::mod();
        long long bw = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { bw += (long long) p[k] - (long long) c[k]; c[k] = (uint32_t) bw; bw >>= 32; }
    }
    *mag = (unsigned long long) c[0] | ((unsigned long long) c[1] << 32);
    return neg;
}

// ---- gate evaluation ---------------------------------------------------------------------------------------------------
// records sorted by OUTPUT gate g and cut into items (the machinery of the sumcheck gate passes, build_schedule):
//   x = index of the u operand, g = index of the v operand (binary gates), meta: bits 0-8 sc, bit 16 binary gate,
//   bit 17 u lives in the previous layer (else in layer 0, absolute index), bit 18 v lives in the previous layer
constexpr uint32_t kEvBin = 1u << 16, kEvUPrev = 1u << 17, kEvVPrev = 1u << 18;

__global__ void __launch_bounds__(kBlock) k_eval_items(gate_args_t A) {
    ZK_PDL_ENTRY();
    for (uint32_t it = blockIdx.x * kBlock + threadIdx.x; it < A.n_items; it += gridDim.x * kBlock) {
        const item_t I = A.items[it];
        const uint32_t cnt = I.count_flags & 0xffffu;
        fr_lazy_t acc;
        acc.clear();
        for (uint32_t k = 0; k < cnt; ++k) {
            const gate_rec_t R = A.recs[I.begin + kItemGroup * k];
            const uint32_t sc = R.meta & 0x1ffu;
            const fr_t x = ld_fr_g(((R.meta & kEvUPrev) ? A.val_prev : A.val0) + R.x);
            if (R.meta & kEvBin) {
                const fr_t y = ld_fr_g(((R.meta & kEvVPrev) ? A.val_prev : A.val0) + R.g);
                if (sc) acc.mac(x * y, ld_fr_g(A.two_mul + sc));
                else acc.mac(x, y);
            } else acc.mac(x, sc ? ld_fr_g(A.two_mul + sc) : fr_t::one());
        }
        store_item(A, I.dest, fr_lazy_reduce_upto16(acc));
    }
}

__global__ void __launch_bounds__(kBlock) k_scale_vec(fr_t *v, uint64_t n, fr_t s) {
    for (uint64_t i = (uint64_t) blockIdx.x * kBlock + threadIdx.x; i < n; i += (uint64_t) gridDim.x * kBlock) st_fr_g(v + i, ld_fr_g(v + i) * s);
}

// DOT_PROD layer: CSR by output block g of (u block, v block) pairs; one thread per (g, t)
struct dp_eval_t { uint32_t u, v; };
__global__ void __launch_bounds__(kBlock) k_dotprod_eval(fr_t *out, const fr_t *src, const uint32_t *row_ptr, const dp_eval_t *gates, uint32_t n_rows,
                                                         uint32_t fft_bl) {
    ZK_PDL_ENTRY();
    const uint32_t fft_len = 1u << fft_bl;
    const size_t total = (size_t) n_rows << fft_bl;
    for (size_t idx = (size_t) blockIdx.x * kBlock + threadIdx.x; idx < total; idx += (size_t) gridDim.x * kBlock) {
        const uint32_t g = (uint32_t) (idx >> fft_bl), t = (uint32_t) idx & (fft_len - 1);
        const uint32_t k0 = row_ptr[g], k1 = row_ptr[g + 1];
        fr_lazy_t acc;
        acc.clear();
        for (uint32_t k = k0; k < k1; ++k) {
            const dp_eval_t G = gates[k];
            acc.mac(ld_fr_g(src + (((size_t) G.u << fft_bl) | t)), ld_fr_g(src + (((size_t) G.v << fft_bl) | t)));
        }
        st_fr_g(out + idx, k1 - k0 <= 16 ? fr_lazy_reduce_upto16(acc) : fr_lazy_reduce_any(acc));
    }
}

// ---- number-theoretic transform of every block of a layer (src/utils.cpp:105-145; src/neuralNetwork.cpp:946-965) -------
// forward: block k of `src` has 2^(n-1) entries, zero-extended to 2^n, transformed, all 2^n outputs kept
// inverse: block k of `src` has 2^n entries, inverse transform (1/2^n included), the first 2^(n-1) outputs kept
// pw[i] = w^i for the 2^n-th root of unity w (its inverse for the inverse transform).  One CTA per block, data in shared memory.
__global__ void __launch_bounds__(kBlock) k_ntt_blocks(fr_t *out, const fr_t *src, const fr_t *pw, uint32_t n_blocks, uint32_t n, uint32_t inverse, fr_t ilen) {
    ZK_DYN_SMEM(fr_t, a);
    const uint32_t len = 1u << n, half = len >> 1;
    for (uint32_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const fr_t *in = src + (size_t) blk * (inverse ? len : half);
        for (uint32_t i = threadIdx.x; i < len; i += kBlock) {   // bit-reversed load
            const uint32_t rev = n ? __brev(i) >> (32 - n) : 0;
            st_fr(a + rev, (inverse || i < half) ? ld_fr_g(in + i) : fr_t::zero());
        }
        __syncthreads();
        for (uint32_t span = 2; span <= len; span <<= 1) {
            const uint32_t hs = span >> 1, step = len / span;
            for (uint32_t b = threadIdx.x; b < half; b += kBlock) {
                const uint32_t k = b & (hs - 1), j = (b / hs) * span;
                const fr_t u = ld_fr(a + j + k), v = ld_fr(a + j + k + hs) * ld_fr_g(pw + (size_t) step * k);
                st_fr(a + j + k, u + v);
                st_fr(a + j + k + hs, u - v);
            }
            __syncthreads();
        }
        fr_t *o = out + (size_t) blk * (inverse ? half : len);
        if (inverse)
            for (uint32_t i = threadIdx.x; i < half; i += kBlock) st_fr_g(o + i, ld_fr(a + i) * ilen);
        else
            for (uint32_t i = threadIdx.x; i < len; i += kBlock) st_fr_g(o + i, ld_fr(a + i));
        __syncthreads();
    }
}

// ---- auxiliary inputs (bit decompositions) -----------------------------------------------------------------------------
struct aux_op_t {
    uint32_t src, dst;   // index into the source layer / into val[0]
    uint32_t meta;       // bits 0-7: bit position; bits 8-9: kAuxSign / kAuxBit / kAuxMax
};
constexpr uint32_t kAuxSign = 0, kAuxBit = 1, kAuxMax = 2;

__global__ void __launch_bounds__(kBlock) k_aux_bits(fr_t *val0, const fr_t *src, const aux_op_t *ops, uint64_t n_ops) {
    ZK_PDL_ENTRY();
    for (uint64_t i = (uint64_t) blockIdx.x * kBlock + threadIdx.x; i < n_ops; i += (uint64_t) gridDim.x * kBlock) {
        const aux_op_t op = ops[i];
        unsigned long long mag;
        const bool neg = fr_sign_magnitude(ld_fr_g(src + op.src), &mag);
        const bool bit = ((op.meta >> 8) & 3u) == kAuxSign ? neg : ((mag >> (op.meta & 0xffu)) & 1ull) != 0;
        st_fr_g(val0 + op.dst, bit ? fr_t::one() : fr_t::zero());
    }
}
// running maximum of max(0, value) over the window elements of a pooling cell: scratch[dst - base] (zero on entry)
__global__ void __launch_bounds__(kBlock) k_aux_max(unsigned long long *scratch, uint32_t base, const fr_t *src, const aux_op_t *ops, uint64_t n_ops) {
    ZK_PDL_ENTRY();
    for (uint64_t i = (uint64_t) blockIdx.x * kBlock + threadIdx.x; i < n_ops; i += (uint64_t) gridDim.x * kBlock) {
        const aux_op_t op = ops[i];
        unsigned long long mag;
        const bool neg = fr_sign_magnitude(ld_fr_g(src + op.src), &mag);
        if (!neg && mag) atomicMax(scratch + (op.dst - base), mag);
    }
}
__global__ void __launch_bounds__(kBlock) k_aux_max_store(fr_t *val0, uint32_t base, const unsigned long long *scratch, uint32_t n) {
    ZK_PDL_ENTRY();
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) st_fr_g(val0 + base + i, fr_t::from_u64(scratch[i]));
}

// out[0] = largest non-negative value, out[1] = largest magnitude of a negative value (both zero on entry)
__global__ void __launch_bounds__(kBlock) k_layer_range(const fr_t *val, uint64_t n, unsigned long long *out) {
    unsigned long long mx = 0, mn = 0;
    for (uint64_t i = (uint64_t) blockIdx.x * kBlock + threadIdx.x; i < n; i += (uint64_t) gridDim.x * kBlock) {
        unsigned long long mag;
        if (fr_sign_magnitude(ld_fr_g(val + i), &mag)) mn = mag > mn ? mag : mn;
        else mx = mag > mx ? mag : mx;
    }
    if (mx) atomicMax(out, mx);
    if (mn) atomicMax(out + 1, mn);
}

}  // namespace zk
