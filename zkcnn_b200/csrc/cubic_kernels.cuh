// K2: the cubic sumcheck round of the FFT-convolution (DOT_PROD) layer, prover::sumcheckDotProdUpdate1
// (src/prover.cpp:103-144), on the machinery of the K1 kernels (sc_kernels.cuh): unreduced (lazy) accumulation of the
// products, exact grid-wide limb sums instead of shared-memory trees, TMA-staged row blocks for tables that stream from HBM.
//
// Tables of the round (evaluation form, one Fr per entry):
//   V1 = V_mult[1]: the previous (FFT) layer's values: activation blocks followed by weight blocks, 2^fft_bl entries each
//   V0 = V_mult[0]: sum over gates of beta_g * weight FFT, one block per ACTIVATION block -- rows without gates are zero
//                   (src/prover.cpp:86-91 only ever writes rows gate.u), so V0 is live only below live0 <= live1
//   m  = mult_array[1]: eq table over the frequency index, 2^fft_bl entries, periodic in the table index
// and the round polynomial is  sum_i m[i mod P](x) V1[i](x) V0[i](x)  (cubic in x).
//
// Two facts shape the kernels:
//   * beyond live0 the product vanishes, so that part of V1 is only FOLDED (2 multiplications per output pair, no
//     products): CTAs [0, nb0) work on the product segment, CTAs [nb0, nb0 + nb1) on the fold-only segment.
//   * m depends on i only through j = i mod P.  A thread whose pair index advances by a multiple of P sees ONE j, so it
//     accumulates the quadratic  Q_j(x) = sum_i V1[i](x) V0[i](x)  (three unreduced products per pair, no full
//     multiplication) and multiplies by m_j(x) once, after its loop ("factored" mode).  Threads with few iterations use
//     the direct form (three full products + six unreduced ones per pair) whose fixed cost is lower.
//   The multiplier table is folded on the fly (each thread needs two entries of it); the product segment also writes the
//   folded table out for the next round, so there is no separate launch for it.
#pragma once
#include "sc_kernels.cuh"

namespace zk {

constexpr int kCubicLimbs = 4 * fr_lazy_t::W;   // four unreduced sums (a, b, c, d) of 17 limbs

struct cubic_args_t {
    const fr_t *v1_in, *v0_in;
    fr_t *v1_out, *v0_out;      // folded tables (n_in / 2); unused when fold == 0
    uint32_t n_in;              // evaluations per table (power of two)
    uint32_t live1, live0;      // entries >= live are zero; live0 <= live1
    uint32_t fold;              // 0: first round of the phase (previous_random = 0): no fold
    const fr_t *m_in;           // multiplier table BEFORE this round's fold, m_n entries
    fr_t *m_out;                // fold != 0 && m_n >= 2: receives the folded multiplier table (m_n / 2 entries)
    uint32_t m_n;
    uint32_t nb0, nb1;          // CTAs of the product segment / of the fold-only segment
    fr_t r;                     // previous_random
    unsigned long long *acc;    // [kCubicLimbs] grid-wide limb sums: zero on entry, zero again on exit
    uint32_t *counter;          // "CTAs done" ticket of the product segment; self-resetting
    fr_t *out;                  // (a, b, c, d)
    uint32_t *flag;             // != nullptr: publish `seq` here once `out` is written (mapped host memory)
    uint32_t seq;
};

// any number of products (< 2^32) of reduced operands: T / R mod r
ZK_HD __forceinline__ fr_t fr_lazy_reduce_any(const fr_lazy_t &a) {
    uint32_t top[8] = {a.w[16], 0, 0, 0, 0, 0, 0, 0};
    return montgomery_of_wide<fr_cfg>(a.w, a.w + 8, top);
}

// this thread's multiplier pair for the round: (m0, m1 - m0) of entry pair j of the CURRENT table (folded on the fly)
struct cubic_mult_t { fr_t m0, dm; };
__device__ __forceinline__ fr_t cubic_fold_entry(const cubic_args_t &A, uint32_t k) {   // entry k of the current multiplier table
    if (A.fold && A.m_n >= 2) {
        const fr_t x0 = ld_fr_g(A.m_in + 2 * k), x1 = ld_fr_g(A.m_in + 2 * k + 1);
        return x0 + A.r * (x1 - x0);
    }
    return ld_fr_g(A.m_in + k);
}
__device__ __forceinline__ uint32_t cubic_cur_n(const cubic_args_t &A) { return (A.fold && A.m_n >= 2) ? A.m_n >> 1 : A.m_n; }
__device__ __forceinline__ cubic_mult_t cubic_mult_of(const cubic_args_t &A, uint32_t first_pair) {
    const uint32_t cur_n = cubic_cur_n(A);
    cubic_mult_t M;
    if (cur_n >= 2) {
        const uint32_t j = first_pair & ((cur_n >> 1) - 1);
        M.m0 = cubic_fold_entry(A, 2 * j);
        M.dm = cubic_fold_entry(A, 2 * j + 1) - M.m0;
    } else {   // constant multiplier (src/prover.cpp:135: total[0] == 0 -> tmp_mult[0])
        M.m0 = cubic_fold_entry(A, 0);
        M.dm = fr_t::zero();
    }
    return M;
}
// (qa x^2 + qb x + qc)(dm x + m0) added, unreduced, to the four coefficient sums
__device__ __forceinline__ void cubic_times_mult(fr_lazy_t (&co)[4], const cubic_mult_t &M, const fr_t &qa, const fr_t &qb, const fr_t &qc) {
    co[0].mac(M.dm, qa);
    co[1].mac(M.dm, qb);
    co[1].mac(M.m0, qa);
    co[2].mac(M.dm, qc);
    co[2].mac(M.m0, qb);
    co[3].mac(M.m0, qc);
}

// Last CTA of the product segment: sh_tot holds the four exact limb sums; 12 threads turn the three 8-limb chunks of each
// into field elements (see round_quad_publish), four threads add up, thread 0 stores / publishes.
__device__ __forceinline__ void round_cubic_publish(const cubic_args_t &A, const unsigned long long *sh_tot, fr_t *sh_fr /* [16] */) {
    if (threadIdx.x < 12) {
        const int k = threadIdx.x / 3, c = threadIdx.x % 3;
        uint32_t t[fr_lazy_t::W + 2 + 5];
        limb_sums_normalise(sh_tot + k * fr_lazy_t::W, fr_lazy_t::W, t);
#pragma unroll
        for (int j = fr_lazy_t::W + 2; j < fr_lazy_t::W + 7; ++j) t[j] = 0;
        fr_t x, y;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            x.v[j] = t[8 * c + j];
            y.v[j] = c == 0 ? (j == 0 ? 1u : 0u) : c == 1 ? fr_cfg::one()[j] : fr_cfg::r2()[j];
        }
        uint32_t pm[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pm[j] = fr_cfg::mod()[j];
        while (fr_t::ge_raw(x.v, pm)) fr_t::raw_sub(x.v, pm);
        st_fr(sh_fr + threadIdx.x, x * y);
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        const fr_t s = sh_fr[3 * threadIdx.x] + sh_fr[3 * threadIdx.x + 1] + sh_fr[3 * threadIdx.x + 2];
        st_fr(A.out + threadIdx.x, s);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0 && A.flag) publish(A.flag, A.seq);
}

// after the loop of a product-segment thread: bring its sums into the four coefficient accumulators
template <bool FACTORED>
__device__ __forceinline__ void cubic_finish_thread(fr_lazy_t (&co)[4], const fr_lazy_t (&q)[3], const cubic_mult_t &M, bool any) {
    if (!FACTORED) return;   // the direct form adds into co[] as it goes
    if (!any) return;
    const fr_t qa = fr_lazy_reduce_any(q[0]), qc = fr_lazy_reduce_any(q[1]), qe = fr_lazy_reduce_any(q[2]);
    cubic_times_mult(co, M, qa, qe - qa - qc, qc);
}

// one output pair: (p0, p1) of V0, (q0, q1) of V1
template <bool FACTORED>
__device__ __forceinline__ void cubic_accumulate(fr_lazy_t (&co)[4], fr_lazy_t (&q)[3], const cubic_mult_t &M, const fr_t &p0, const fr_t &p1, const fr_t &q0,
                                                 const fr_t &q1) {
    if (FACTORED) {
        q[0].mac(fr_t::sub_lazy(p1, p0), fr_t::sub_lazy(q1, q0));
        q[1].mac(p0, q0);
        q[2].mac(p1, q1);
    } else {
        const fr_t qa = (p1 - p0) * fr_t::sub_lazy(q1, q0), qc = p0 * q0, qe = p1 * q1;   // (at most one unreduced operand per full multiplication)
        cubic_times_mult(co, M, qa, qe - qa - qc, qc);
    }
}

// Generic kernel: first round of a phase (no fold), tables below the TMA threshold, and the emulator build.
// Host contract for the product segment: nb0 * kRoundBlock is a multiple of the multiplier period (pairs), or every
// thread has at most one iteration -- either way a thread sees one j.
template <bool FACTORED> __global__ void __launch_bounds__(kRoundBlock) k_round_cubic(cubic_args_t A) {
    __shared__ unsigned long long sh_warp[(kRoundBlock / 32) * kCubicLimbs];
    __shared__ unsigned long long sh_tot[kCubicLimbs];
    __shared__ uint32_t sh_ticket;
    __shared__ fr_t sh_fr[16];
    ZK_PDL_ENTRY();
    const uint32_t n_pairs = A.fold ? A.n_in >> 2 : A.n_in >> 1;
    const uint32_t lp0 = A.fold ? (A.live0 + 3) >> 2 : (A.live0 + 1) >> 1, lp1 = A.fold ? (A.live1 + 3) >> 2 : (A.live1 + 1) >> 1;
    const uint32_t P0 = lp0 < n_pairs ? lp0 : n_pairs, P1 = lp1 < n_pairs ? lp1 : n_pairs;
    const fr_t r = A.r;
    if (blockIdx.x >= A.nb0) {   // fold-only segment: V1 beyond the live part of V0
        const uint32_t stride = A.nb1 * kRoundBlock;
        for (uint32_t i = P0 + (blockIdx.x - A.nb0) * kRoundBlock + threadIdx.x; i < P1; i += stride) {
            const uint32_t base = i << 2;
            const fr_t x0 = ld_fr_live(A.v1_in, base, A.live1), x1 = ld_fr_live(A.v1_in, base + 1, A.live1);
            const fr_t x2 = ld_fr_live(A.v1_in, base + 2, A.live1), x3 = ld_fr_live(A.v1_in, base + 3, A.live1);
            st_fr_g(A.v1_out + 2 * i, x0 + r * fr_t::sub_lazy(x1, x0));
            st_fr_g(A.v1_out + 2 * i + 1, x2 + r * fr_t::sub_lazy(x3, x2));
        }
        return;
    }
    const uint32_t stride = A.nb0 * kRoundBlock;
    const uint32_t first = blockIdx.x * kRoundBlock + threadIdx.x;
    // the folded multiplier table for the next round
    if (A.fold && A.m_n >= 2)
        for (uint32_t k = first; k < (A.m_n >> 1); k += stride) st_fr_g(A.m_out + k, cubic_fold_entry(A, k));
    const cubic_mult_t M = cubic_mult_of(A, first);
    fr_lazy_t co[4], q[3];
#pragma unroll
    for (int k = 0; k < 4; ++k) co[k].clear();
#pragma unroll
    for (int k = 0; k < 3; ++k) q[k].clear();
    for (uint32_t i = first; i < P0; i += stride) {
        fr_t p0, p1, q0, q1;
        if (A.fold) {
            const uint32_t base = i << 2;
            fr_t x0 = ld_fr_live(A.v1_in, base, A.live1), x1 = ld_fr_live(A.v1_in, base + 1, A.live1);
            fr_t x2 = ld_fr_live(A.v1_in, base + 2, A.live1), x3 = ld_fr_live(A.v1_in, base + 3, A.live1);
            q0 = x0 + r * fr_t::sub_lazy(x1, x0);
            q1 = x2 + r * fr_t::sub_lazy(x3, x2);
            st_fr_g(A.v1_out + 2 * i, q0);
            st_fr_g(A.v1_out + 2 * i + 1, q1);
            x0 = ld_fr_live(A.v0_in, base, A.live0); x1 = ld_fr_live(A.v0_in, base + 1, A.live0);
            x2 = ld_fr_live(A.v0_in, base + 2, A.live0); x3 = ld_fr_live(A.v0_in, base + 3, A.live0);
            p0 = x0 + r * fr_t::sub_lazy(x1, x0);
            p1 = x2 + r * fr_t::sub_lazy(x3, x2);
            st_fr_g(A.v0_out + 2 * i, p0);
            st_fr_g(A.v0_out + 2 * i + 1, p1);
        } else {
            q0 = ld_fr_live(A.v1_in, 2 * i, A.live1); q1 = ld_fr_live(A.v1_in, 2 * i + 1, A.live1);
            p0 = ld_fr_live(A.v0_in, 2 * i, A.live0); p1 = ld_fr_live(A.v0_in, 2 * i + 1, A.live0);
        }
        cubic_accumulate<FACTORED>(co, q, M, p0, p1, q0, q1);
    }
    cubic_finish_thread<FACTORED>(co, q, M, first < P0);
    uint32_t limb[kCubicLimbs];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < fr_lazy_t::W; ++j) limb[k * fr_lazy_t::W + j] = co[k].w[j];
    if (!grid_limb_sum<kCubicLimbs, kRoundBlock>(limb, A.acc, A.counter, A.nb0, sh_warp, sh_tot, &sh_ticket)) return;
    round_cubic_publish(A, sh_tot, sh_fr);
}

#if !defined(ZK_EMU)
// --------------------------------------------------------------------------------------------------------------------
// K2, HBM-streaming rounds: the TMA pipeline of k_round_quad_tma (one double-buffered pair of 4 KB boxes per warp, 128-byte
// swizzle, one mbarrier per (warp, stage)).  A warp iteration takes two boxes of 32 rows (one row = the four inputs of one
// output pair) and produces 64 folded output pairs:
//   product segment  : box A = rows g of V1, box B = rows g of V0; products of the 32 pairs, factored by the multiplier
//   fold-only segment: box A = rows g, box B = rows g + 1 of V1 (beyond the live part of V0)
// --------------------------------------------------------------------------------------------------------------------
struct alignas(64) cubic_tma_args_t {
    cubic_args_t C;
    CUtensorMap tm_v1, tm_v0;
};

__global__ void __launch_bounds__(kRoundBlock, kTmaCtasPerSm) k_round_cubic_tma(const __grid_constant__ cubic_tma_args_t T) {
    extern __shared__ __align__(1024) unsigned char dyn_smem[];
    __shared__ unsigned long long sh_warp[(kRoundBlock / 32) * kCubicLimbs];
    __shared__ unsigned long long sh_tot[kCubicLimbs];
    __shared__ uint32_t sh_ticket;
    __shared__ fr_t sh_fr[16];
    __shared__ __align__(8) unsigned long long sh_bar[(kRoundBlock / 32) * 2];
    const cubic_args_t &A = T.C;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n_pairs = A.n_in >> 2;
    const uint32_t lp0 = (A.live0 + 3) >> 2, lp1 = (A.live1 + 3) >> 2;
    const uint32_t P0 = lp0 < n_pairs ? lp0 : n_pairs, P1 = lp1 < n_pairs ? lp1 : n_pairs;
    const uint32_t G0 = (P0 + 31) >> 5, G1 = (P1 + 31) >> 5;                 // row blocks of the product segment / of V1
    const uint32_t full0 = (A.live0 >> 2) >> 5, full1 = (A.live1 >> 2) >> 5;  // row blocks whose 128 inputs are all live
    const bool seg1 = blockIdx.x >= A.nb0;
    const uint32_t bx = seg1 ? blockIdx.x - A.nb0 : blockIdx.x;
    const uint32_t gstride = (seg1 ? A.nb1 : A.nb0) * (kRoundBlock / 32);   // in warp iterations
    const uint32_t n_iter = seg1 ? (G1 - G0 + 1) >> 1 : G0;                  // warp iterations of this segment
    const CUtensorMap *tm_a = &T.tm_v1, *tm_b = seg1 ? &T.tm_v1 : &T.tm_v0;
    const fr_t *in_a = A.v1_in, *in_b = seg1 ? A.v1_in : A.v0_in;
    fr_t *out_a = A.v1_out, *out_b = seg1 ? A.v1_out : A.v0_out;
    const uint32_t live_a = A.live1, live_b = seg1 ? A.live1 : A.live0;
    const uint32_t lim_a = P1, lim_b = seg1 ? P1 : P0;                        // output pairs with any live input, per box

    const uint32_t box0 = ((smem_addr(dyn_smem) + 1023u) & ~1023u) + warp * kTmaWarpBytes;
    const uint32_t bar0 = smem_addr(sh_bar + 2 * warp);
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    // row blocks of warp iteration `it` and whether both can be taken as full TMA boxes
    auto blk_a = [&](uint32_t it) { return seg1 ? G0 + 2 * it : it; };
    auto blk_b = [&](uint32_t it) { return seg1 ? G0 + 2 * it + 1 : it; };
    auto is_full = [&](uint32_t it) { return seg1 ? blk_b(it) < full1 : it < full0; };   // (full0 <= full1: box A of the product segment is full too)

    const uint32_t first_pair = (bx * (kRoundBlock / 32) + warp) * 32 + lane;   // product segment: this thread's first output pair
    cubic_mult_t M;
    if (!seg1) {
        const uint32_t tg = bx * kRoundBlock + threadIdx.x, tstride = A.nb0 * kRoundBlock;
        if (A.m_n >= 2)
            for (uint32_t k = tg; k < (A.m_n >> 1); k += tstride) st_fr_g(A.m_out + k, cubic_fold_entry(A, k));
        M = cubic_mult_of(A, first_pair);
    }
    fr_lazy_t q[3];
    q[0].clear(); q[1].clear(); q[2].clear();
    const fr_t r = A.r;
    uint32_t stage = 0, parity = 0;
    uint32_t it = bx * (kRoundBlock / 32) + warp;
    if (it < n_iter && is_full(it) && lane == 0) {
        mbar_expect_tx(bar0, 2 * kTmaBoxBytes);
        tma_load_rows(box0, tm_a, blk_a(it) * 32, bar0);
        tma_load_rows(box0 + kTmaBoxBytes, tm_b, blk_b(it) * 32, bar0);
    }
    for (; it < n_iter; it += gstride) {
        const uint32_t nx = it + gstride;
        __syncwarp();   // every lane is done reading the other stage
        if (nx < n_iter && is_full(nx) && lane == 0) {
            const uint32_t b = bar0 + 8 * (stage ^ 1u), dst = box0 + (stage ^ 1u) * 2 * kTmaBoxBytes;
            mbar_expect_tx(b, 2 * kTmaBoxBytes);
            tma_load_rows(dst, tm_a, blk_a(nx) * 32, b);
            tma_load_rows(dst + kTmaBoxBytes, tm_b, blk_b(nx) * 32, b);
        }
        const uint32_t ia = blk_a(it) * 32 + lane, ib = blk_b(it) * 32 + lane;   // output pairs of this lane
        fr_t x0, x1, x2, x3, y0, y1, y2, y3;
        if (is_full(it)) {
            mbar_wait(bar0 + 8 * stage, (parity >> stage) & 1u);
            parity ^= 1u << stage;
            const uint32_t arow = box0 + stage * 2 * kTmaBoxBytes + lane * 128u, brow = arow + kTmaBoxBytes;
            x0 = lds_fr_swz(arow, lane, 0); x1 = lds_fr_swz(arow, lane, 1); x2 = lds_fr_swz(arow, lane, 2); x3 = lds_fr_swz(arow, lane, 3);
            y0 = lds_fr_swz(brow, lane, 0); y1 = lds_fr_swz(brow, lane, 1); y2 = lds_fr_swz(brow, lane, 2); y3 = lds_fr_swz(brow, lane, 3);
        } else {
            const bool on_a = ia < lim_a, on_b = ib < lim_b;
            x0 = on_a ? ld_fr_live(in_a, 4 * ia, live_a) : fr_t::zero(); x1 = on_a ? ld_fr_live(in_a, 4 * ia + 1, live_a) : fr_t::zero();
            x2 = on_a ? ld_fr_live(in_a, 4 * ia + 2, live_a) : fr_t::zero(); x3 = on_a ? ld_fr_live(in_a, 4 * ia + 3, live_a) : fr_t::zero();
            y0 = on_b ? ld_fr_live(in_b, 4 * ib, live_b) : fr_t::zero(); y1 = on_b ? ld_fr_live(in_b, 4 * ib + 1, live_b) : fr_t::zero();
            y2 = on_b ? ld_fr_live(in_b, 4 * ib + 2, live_b) : fr_t::zero(); y3 = on_b ? ld_fr_live(in_b, 4 * ib + 3, live_b) : fr_t::zero();
        }
        const fr_t dx0 = fr_t::sub_lazy(x1, x0), dx1 = fr_t::sub_lazy(x3, x2), dy0 = fr_t::sub_lazy(y1, y0), dy1 = fr_t::sub_lazy(y3, y2);
        const fr_t a0 = x0 + ZK_TMA_MUL(r, dx0);
        const fr_t a1 = x2 + ZK_TMA_MUL(r, dx1);
        const fr_t b0 = y0 + ZK_TMA_MUL(r, dy0);
        const fr_t b1 = y2 + ZK_TMA_MUL(r, dy1);
        if (ia < lim_a) {
            st_fr(out_a + 2 * ia, a0);
            st_fr(out_a + 2 * ia + 1, a1);
        }
        if (ib < lim_b) {
            st_fr(out_b + 2 * ib, b0);
            st_fr(out_b + 2 * ib + 1, b1);
        }
        if (!seg1) {   // (a0, a1) = (q0, q1) of V1, (b0, b1) = (p0, p1) of V0
            q[0].mac(fr_t::sub_lazy(b1, b0), fr_t::sub_lazy(a1, a0));
            q[1].mac(b0, a0);
            q[2].mac(b1, a1);
        }
        stage ^= 1u;
    }
    if (seg1) return;
    fr_lazy_t co[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) co[k].clear();
    cubic_finish_thread<true>(co, q, M, first_pair < P0);
    uint32_t limb[kCubicLimbs];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < fr_lazy_t::W; ++j) limb[k * fr_lazy_t::W + j] = co[k].w[j];
    if (!grid_limb_sum<kCubicLimbs, kRoundBlock>(limb, A.acc, A.counter, A.nb0, sh_warp, sh_tot, &sh_ticket)) return;
    round_cubic_publish(A, sh_tot, sh_fr);
}
#endif  // !ZK_EMU

// --------------------------------------------------------------------------------------------------------------------
// Dense passes of the FFT-convolution path with unreduced accumulation (one Montgomery reduction per output instead of
// one per term).
// --------------------------------------------------------------------------------------------------------------------
// K5b, DOT_PROD phase 1:  V0[(u << fft_bl) | t] = sum_{gates with that u} beta_g[g] * val[(v << fft_bl) | t]
// (src/prover.cpp:86-91).  CSR by u built at upload; one thread per (u, t), coalesced over t; rows >= n_rows stay unwritten
// (they are beyond live0).
struct dp_gate_t { uint32_t g, v; };
// `splits` threads share one output (u, t): thread s takes every splits-th gate of the row and leaves a partial sum in
// dst[s * total + idx] (splits == 1: dst is the table itself); k_colsum_finish adds the partials.  With two pictures a layer has as
// few as 6 rows of 512 gates each: without the split the launch would be 128 CTAs walking 512 gates one after the other.
__global__ void __launch_bounds__(kBlock) k_dotprod_axpy(fr_t *dst, const fr_t *val, const fr_t *beta_g, const uint32_t *row_ptr,
                                                         const dp_gate_t *gates, uint32_t n_rows, uint32_t fft_bl, uint32_t splits) {
    ZK_PDL_ENTRY();
    const uint32_t fft_len = 1u << fft_bl;
    const size_t total = (size_t) n_rows << fft_bl, work = total * splits;
    for (size_t w = (size_t) blockIdx.x * kBlock + threadIdx.x; w < work; w += (size_t) gridDim.x * kBlock) {
        const size_t idx = w % total;
        const uint32_t sp = (uint32_t) (w / total);
        const uint32_t u = (uint32_t) (idx >> fft_bl), t = (uint32_t) idx & (fft_len - 1);
        const uint32_t k0 = row_ptr[u], k1 = row_ptr[u + 1];
        fr_lazy_t acc;
        acc.clear();
        uint32_t cnt = 0;
        for (uint32_t k = k0 + sp; k < k1; k += splits, ++cnt) {
            const dp_gate_t G = gates[k];
            acc.mac(ld_fr_g(beta_g + G.g), ld_fr_g(val + (((size_t) G.v << fft_bl) | t)));
        }
        st_fr_g(dst + w, cnt <= 16 ? fr_lazy_reduce_upto16(acc) : fr_lazy_reduce_any(acc));
    }
}

// K4b: FFT/IFFT layers: V[u] = sum_g val[(g << shift) | u] * beta_g[g],  u < 2^shift  (src/prover.cpp:190-197); the same shape
// is the R^T Z pass of the Hyrax opening (polyProver.cpp:66-68).  `val` is a row-major [n_g][2^shift] matrix, i.e. ONE linear
// array: thread t of T (T a multiple of 2^shift, so a thread keeps its column) streams elements t, t + T, ... with unreduced
// accumulation and leaves one partial sum; k_colsum_finish adds the T / 2^shift partials of a column with one warp per column
// (columns can be as few as 16 while the rows number hundreds of thousands: a thread per column would add them one by one).
__global__ void __launch_bounds__(kBlock) k_dense_colsum(const fr_t *val, const fr_t *beta_g, uint32_t shift, uint64_t total, fr_t *partial) {
    ZK_PDL_ENTRY();
    const uint64_t T = (uint64_t) gridDim.x * kBlock, t = (uint64_t) blockIdx.x * kBlock + threadIdx.x;
    fr_lazy_t acc;
    acc.clear();
    uint32_t cnt = 0;
    for (uint64_t i = t; i < total; i += T, ++cnt) acc.mac(ld_fr_g(val + i), ld_fr_g(beta_g + (i >> shift)));
    st_fr_g(partial + t, cnt <= 16 ? fr_lazy_reduce_upto16(acc) : fr_lazy_reduce_any(acc));
}
// out[u] = sum_k partial[u + k * n_u], k < per_u; one warp per column u
__global__ void __launch_bounds__(kBlock) k_colsum_finish(const fr_t *partial, uint32_t n_u, uint32_t per_u, fr_t *out) {
    ZK_PDL_ENTRY();
    const uint32_t lane = threadIdx.x & 31u, wpc = kBlock / 32;
    const uint32_t n_groups = (n_u + wpc - 1) / wpc;
    for (uint32_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {   // CTA-uniform trip count: every lane takes part in the shuffles
        const uint32_t u = grp * wpc + (threadIdx.x >> 5);
        fr_t acc = fr_t::zero();
        if (u < n_u) {
            uint32_t k = lane;
            for (; k + 96 < per_u; k += 128) {   // four independent loads in flight
                const fr_t a = ld_fr_g(partial + u + (size_t) k * n_u), b = ld_fr_g(partial + u + (size_t) (k + 32) * n_u);
                const fr_t c = ld_fr_g(partial + u + (size_t) (k + 64) * n_u), d = ld_fr_g(partial + u + (size_t) (k + 96) * n_u);
                acc = acc + ((a + b) + (c + d));
            }
            for (; k < per_u; k += 32) acc = acc + ld_fr_g(partial + u + (size_t) k * n_u);
        }
        for (uint32_t d = 16; d; d >>= 1) {
            fr_t o;
#pragma unroll
            for (int j = 0; j < 8; ++j) o.v[j] = __shfl_xor_sync(0xffffffffu, acc.v[j], d);
            acc = acc + o;
        }
        if (lane == 0 && u < n_u) st_fr_g(out + u, acc);
    }
}

// K5 (DOT_PROD phase 2): V[v] = sum_t val[(v << fft_bl) | t] * beta_gs[t]  (src/prover.cpp:277-284).
// `lanes` (a power of two <= 32) consecutive lanes share one row; their unreduced sums meet through xor-shuffles (exact
// integer additions), the first lane of the group reduces and stores.
__global__ void __launch_bounds__(kBlock) k_dense_rowdot(fr_t *out, const fr_t *val, const fr_t *beta_gs, uint32_t n_rows, uint32_t fft_bl, uint32_t lanes) {
    ZK_PDL_ENTRY();
    const uint32_t fft_len = 1u << fft_bl;
    const uint32_t rows_per_cta = kBlock / lanes, sub = threadIdx.x % lanes;
    const uint32_t n_groups = (n_rows + rows_per_cta - 1) / rows_per_cta;
    for (uint32_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {   // CTA-uniform trip count: every lane takes part in the shuffles
        const uint32_t v = grp * rows_per_cta + threadIdx.x / lanes;
        fr_lazy_t acc;
        acc.clear();
        if (v < n_rows)
            for (uint32_t t = sub; t < fft_len; t += lanes) acc.mac(ld_fr_g(val + (((size_t) v << fft_bl) | t)), ld_fr_g(beta_gs + t));
        for (uint32_t d = lanes >> 1; d; d >>= 1) {
            uint32_t o[fr_lazy_t::W];
#pragma unroll
            for (int j = 0; j < fr_lazy_t::W; ++j) o[j] = __shfl_xor_sync(0xffffffffu, acc.w[j], d);
            uint64_t c = 0;
#pragma unroll
            for (int j = 0; j < fr_lazy_t::W; ++j) {
                c += (uint64_t) acc.w[j] + o[j];
                acc.w[j] = (uint32_t) c;
                c >>= 32;
            }
        }
        if (sub == 0 && v < n_rows) st_fr_g(out + v, fr_lazy_reduce_any(acc));
    }
}

// micro-benchmark (zk_bench_field_mul): two independent dependent-multiplication chains per thread: the sustained multiplication rate with every SM busy
template <class F> __global__ void __launch_bounds__(kBlock) k_mul_chain(F *out, uint32_t iters, uint64_t seed) {
    F x = F::from_u64(seed + blockIdx.x * kBlock + threadIdx.x + 2), y = F::from_u64(seed * 3 + threadIdx.x + 5), a = x, b = y;
    for (uint32_t i = 0; i < iters; ++i) { a = a * x; b = b * y; }
    a = a + b;
    if (a.v[0] == 0x12345678u && a.v[1] == 0x9abcdef0u) out[blockIdx.x * kBlock + threadIdx.x] = a;   // (keeps the chain alive; practically never taken)
}

}  // namespace zk
