// Sumcheck kernels (GKR prover hot loops (1) and (2) of BASELINE.json north_star).
//
// All bookkeeping tables are kept in EVALUATION form (one Fr per entry) instead of the reference's linear_poly
// (a,b) = (v1 - v0, v0) pairs (src/prover.cpp:13-15), which halves the HBM traffic; the round polynomials that leave
// the device are the same field elements (field arithmetic is exact, so summation order does not matter).
//
//   K1  k_round_quad      prover::sumcheckUpdateEach        src/prover.cpp:396-426
//   K2  k_round_cubic / k_round_cubic_tma   prover::sumcheckDotProdUpdate1   src/prover.cpp:103-144   (cubic_kernels.cuh)
//   K3  k_half_tables / k_beta_expand   initBetaTable + initHalfTable   src/utils.cpp:32-51,147-180
//   K3b k_phi_table       phiGInit                          src/utils.cpp:61-103
//   K4/K5 k_gate_items_p1 / k_gate_items_p2 / k_sum_partials   gate loops of sumcheckInitPhase1/2
//                                                           src/prover.cpp:224-233,286-288,297-305
//   K4b k_dense_colsum    FFT/IFFT dense contraction        src/prover.cpp:190-197   (cubic_kernels.cuh)
//   K5b k_dotprod_axpy    sumcheckDotProdInitPhase1         src/prover.cpp:86-91     (cubic_kernels.cuh)
//   K5  k_dense_rowdot    DOT_PROD phase-2 V table          src/prover.cpp:277-284   (cubic_kernels.cuh)
//   K6  k_liu_scatter     sumcheckLiuInit                   src/prover.cpp:334-353
#pragma once
#include "mont.cuh"
#ifndef ZK_EMU
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#endif

namespace zk {

constexpr int kBlock = 256;          // threads per CTA for all streaming kernels

// Programmatic dependent launch.  A kernel that starts with ZK_PDL_ENTRY() may be launched with the
// programmaticStreamSerialization attribute (ZK_KLAUNCH_PDL): its CTAs are then placed, and its parameters fetched, while the
// previous kernel of the stream drains; nothing the predecessor wrote is touched before the wait.  Letting the next launch go
// right after the wait means at most ONE future kernel is ever parked behind a running one.  Launched normally, both
// instructions are no-ops.  One proof is ~1800 mostly tiny, strictly dependent launches: this hides ~2 us of each.
#if ZK_ON_DEVICE
#define ZK_PDL_ENTRY()                                              \
    do {                                                            \
        asm volatile("griddepcontrol.wait;" ::: "memory");          \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
    } while (0)
#else
#define ZK_PDL_ENTRY() do { } while (0)
#endif
constexpr int kMaxGridX = ZK_SM_COUNT * 8;  // persistent-style grids: a multiple of the 148 SMs

// --------------------------------------------------------------------------------------------------------------------
// vectorised global access: an Fr is two 16-byte words
// --------------------------------------------------------------------------------------------------------------------
ZK_HD __forceinline__ fr_t ld_fr(const fr_t *p) {
#if ZK_ON_DEVICE
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 lo = q[0], hi = q[1];
    fr_t r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
#else
    return *p;
#endif
}
ZK_HD __forceinline__ void st_fr(fr_t *p, const fr_t &x) {
#if ZK_ON_DEVICE
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
#else
    *p = x;
#endif
}
// The same for pointers known to be in GLOBAL memory and 32-byte aligned (every table, witness layer and schedule-addressed element is):
// ONE 256-bit access (LDG.E.256 / STG.E.256, sm_100) instead of two 128-bit ones -- a random 32-byte gather is one request, one sector.
ZK_HD __forceinline__ fr_t ld_fr_g(const fr_t *p) {
#if ZK_ON_DEVICE
    fr_t r;
    asm volatile("ld.global.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
#else
    return *p;
#endif
}
ZK_HD __forceinline__ void st_fr_g(fr_t *p, const fr_t &x) {
#if ZK_ON_DEVICE
    asm volatile("st.global.v8.u32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(x.v[0]), "r"(x.v[1]), "r"(x.v[2]), "r"(x.v[3]), "r"(x.v[4]), "r"(x.v[5]),
                 "r"(x.v[6]), "r"(x.v[7]) : "memory");
#else
    *p = x;
#endif
}
// L2-coherent load for data produced by other CTAs of the same launch ("last CTA finishes" reductions)
ZK_HD __forceinline__ fr_t ld_fr_cg(const fr_t *p) {
#if ZK_ON_DEVICE
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 lo = __ldcg(q), hi = __ldcg(q + 1);
    fr_t r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
#else
    return *p;
#endif
}
// guarded load: entries at or beyond `live` are zero by construction (src/prover.cpp:409-417 clears them)
ZK_HD __forceinline__ fr_t ld_fr_live(const fr_t *p, uint32_t idx, uint32_t live) {
    return idx < live ? ld_fr_g(p + idx) : fr_t::zero();
}

// --------------------------------------------------------------------------------------------------------------------
// CTA-wide sum of K field elements per thread; result valid in thread 0.  sh must hold K * kBlock elements.
// --------------------------------------------------------------------------------------------------------------------
template <int K> __device__ __forceinline__ void block_sum(fr_t (&acc)[K], fr_t *sh) {
    const int t = threadIdx.x;
#pragma unroll
    for (int k = 0; k < K; ++k) sh[k * kBlock + t] = acc[k];
    __syncthreads();
    for (int s = kBlock / 2; s > 0; s >>= 1) {
        if (t < s) {
#pragma unroll
            for (int k = 0; k < K; ++k) sh[k * kBlock + t] = sh[k * kBlock + t] + sh[k * kBlock + t + s];
        }
        __syncthreads();
    }
    if (t == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = sh[k * kBlock];
    }
}

// --------------------------------------------------------------------------------------------------------------------
// Exact grid-wide sum of K 32-bit limbs per thread (K <= 64): the integers  sum_threads limb[k]  for every k.
//   warp:  redux.sync on the 16-bit halves of each limb (32 x 65535 < 2^21, no overflow) -> one 64-bit sum per limb
//   CTA :  the warps' sums meet in shared memory, thread k adds them and issues ONE 64-bit red.global.add
//   grid:  "last CTA finishes": the CTA that draws the last ticket reads the K totals back (and clears them for the
//          next launch); they are left in sh_tot[0..K-1] for it.  Returns true in that CTA only (uniformly).
// Used for sums of field elements (K = 8 per element) and of unreduced products (K = 17): integer addition is exact,
// so the order of the atomics does not matter and the result is bit-reproducible.
// --------------------------------------------------------------------------------------------------------------------
// CTA + grid stages: sh_warp[w * K + k] holds warp w's sum of limb k (written by the caller, no barrier yet)
template <int K, int BLOCK>
__device__ __forceinline__ bool grid_limb_sum_finish(unsigned long long *acc, uint32_t *counter, uint32_t n_ctas, unsigned long long *sh_warp /* [BLOCK/32][K] */,
                                                     unsigned long long *sh_tot /* [K] */, uint32_t *sh_ticket) {
    static_assert(K <= BLOCK, "one thread per limb in the CTA stage");
    constexpr int NW = BLOCK / 32;
    __syncthreads();
    if (threadIdx.x < K) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) t += sh_warp[w * K + threadIdx.x];
        if (n_ctas == 1) sh_tot[threadIdx.x] = t;
        else if (t) atomicAdd(acc + threadIdx.x, t);
        __threadfence();
    }
    __syncthreads();
    if (n_ctas == 1) return true;
    if (threadIdx.x == 0) *sh_ticket = atomicAdd(counter, 1u);
    __syncthreads();
    if (*sh_ticket != n_ctas - 1) return false;
    __threadfence();
    if (threadIdx.x < K) {
#if ZK_ON_DEVICE
        sh_tot[threadIdx.x] = __ldcg(acc + threadIdx.x);
#else
        sh_tot[threadIdx.x] = acc[threadIdx.x];
#endif
        acc[threadIdx.x] = 0;
    }
    if (threadIdx.x == 0) *counter = 0;
    __syncthreads();
    return true;
}
// The same for CTAs that each own one SLICE of a wider set of sums: this CTA adds its K limbs at `offset`, the last CTA gets
// all K_TOTAL totals in sh_tot.  (k_round_quad / k_round_quad_tma: one slice of 51 limbs per table pair.)
template <int K, int K_TOTAL, int BLOCK>
__device__ __forceinline__ bool grid_limb_sum_slice(const uint32_t (&limb)[K], uint32_t offset, unsigned long long *acc, uint32_t *counter, uint32_t n_ctas,
                                                    unsigned long long *sh_warp /* [BLOCK/32][K] */, unsigned long long *sh_tot /* [K_TOTAL] */,
                                                    uint32_t *sh_ticket) {
    static_assert(K <= BLOCK && K_TOTAL <= BLOCK, "one thread per limb in the CTA stage");
    constexpr int NW = BLOCK / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t lo = __reduce_add_sync(0xffffffffu, limb[k] & 0xffffu);
        const uint32_t hi = __reduce_add_sync(0xffffffffu, limb[k] >> 16);
        if (lane == (k & 31)) sh_warp[warp * K + k] = (unsigned long long) lo + ((unsigned long long) hi << 16);
    }
    if (n_ctas == 1 && threadIdx.x < K_TOTAL) sh_tot[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x < K) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) t += sh_warp[w * K + threadIdx.x];
        if (n_ctas == 1) sh_tot[offset + threadIdx.x] = t;
        else if (t) atomicAdd(acc + offset + threadIdx.x, t);
        __threadfence();
    }
    __syncthreads();
    if (n_ctas == 1) return true;
    if (threadIdx.x == 0) *sh_ticket = atomicAdd(counter, 1u);
    __syncthreads();
    if (*sh_ticket != n_ctas - 1) return false;
    __threadfence();
    if (threadIdx.x < K_TOTAL) {
#if ZK_ON_DEVICE
        sh_tot[threadIdx.x] = __ldcg(acc + threadIdx.x);
#else
        sh_tot[threadIdx.x] = acc[threadIdx.x];
#endif
        acc[threadIdx.x] = 0;
    }
    if (threadIdx.x == 0) *counter = 0;
    __syncthreads();
    return true;
}
template <int K, int BLOCK>
__device__ __forceinline__ bool grid_limb_sum(const uint32_t (&limb)[K], unsigned long long *acc, uint32_t *counter, uint32_t n_ctas,
                                              unsigned long long *sh_warp /* [BLOCK/32][K] */, unsigned long long *sh_tot /* [K] */,
                                              uint32_t *sh_ticket) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t lo = __reduce_add_sync(0xffffffffu, limb[k] & 0xffffu);
        const uint32_t hi = __reduce_add_sync(0xffffffffu, limb[k] >> 16);
        if (lane == (k & 31)) sh_warp[warp * K + k] = (unsigned long long) lo + ((unsigned long long) hi << 16);
    }
    return grid_limb_sum_finish<K, BLOCK>(acc, counter, n_ctas, sh_warp, sh_tot, sh_ticket);
}
// carry-propagate `n` 64-bit limb sums (weight 2^(32 k)) into n + 2 32-bit limbs
ZK_HD __forceinline__ void limb_sums_normalise(const unsigned long long *tot, int n, uint32_t *out) {
    unsigned long long c = 0;
    for (int k = 0; k < n; ++k) {
        const unsigned long long lo = (tot[k] & 0xffffffffull) + (c & 0xffffffffull);
        out[k] = (uint32_t) lo;
        c = (tot[k] >> 32) + (c >> 32) + (lo >> 32);
    }
    out[n] = (uint32_t) c;
    out[n + 1] = (uint32_t) (c >> 32);
}

// --------------------------------------------------------------------------------------------------------------------
// K1: one sumcheck round on up to two (V, mult) table pairs in one launch.  CTAs [0, pair[0].n_blocks) work on pair 0,
// the rest on pair 1; the round polynomial that leaves the device is the SUM over both pairs (the caller adds them
// anyway, src/prover.cpp:368-383).
// --------------------------------------------------------------------------------------------------------------------
constexpr int kRoundBlock = 128;                 // 4 warps: one per SM sub-partition
#ifndef ZK_ROUND_CTAS_PER_SM
#define ZK_ROUND_CTAS_PER_SM 4
#endif
constexpr int kRoundMaxGrid = ZK_SM_COUNT * ZK_ROUND_CTAS_PER_SM;   // resident CTAs per SM (4: <= 128 registers per thread)
constexpr int kRoundLimbs = 3 * fr_lazy_t::W;    // three unreduced sums of 17 limbs

struct round_pair_t {
    const fr_t *v_in, *m_in;  // current tables, n_in evaluations each, entries >= live are zero
    fr_t *v_out, *m_out;      // folded tables (n_in / 2); unused when fold == 0
    uint32_t n_in, live;
    uint32_t fold;            // 0: first round of a phase (previous_random = 0, src/verifier.cpp:168) -> no fold
    uint32_t n_blocks;        // CTAs assigned to this pair (0 = pair inactive this round)
};
struct round_args_t {
    round_pair_t pair[2];
    fr_t r;                   // previous_random
    unsigned long long *acc;  // [kRoundLimbs] grid-wide limb sums: zero on entry, zero again on exit
    uint32_t *counter;        // "CTAs done" ticket; self-resetting
    fr_t *out;                // (a, b, c), summed over both pairs (may be mapped host memory)
    uint32_t *flag;           // != nullptr: after `out` is written, publish `seq` here (mapped host memory)
    uint32_t seq;
    uint4 *tagged;            // != nullptr: publish (a, b, c) as 8 self-validating 16-byte words instead (see publish_tagged)
    // streaming kernels only (k_round_quad, k_round_quad_tma): the round polynomial of each pair is kept on the device, and
    // derive_b[p] != 0 says that state[p] holds pair p's polynomial of the PREVIOUS round.  Then the third product is skipped:
    // the folded tables satisfy  sum_i m_i v_i = P_prev(r)  (a table identity: every entry is the old pair evaluated at r), i.e.
    // P(0) + P(1) = P_prev(r), so  b = P_prev(r) - 2c - a.
    fr_t *state;              // [2][4]: (a, b, c, -) per pair
    uint32_t derive_b[2];
};
// Fence-free result mailbox: the 24 result words leave as eight 16-byte stores {w0, w1, w2, seq}.  Each store reaches
// host memory as one write, so a word whose tag equals the awaited sequence number carries valid data: no
// __threadfence_system() (1.4 us on B200, tools/latbench.cu) between data and flag.
__device__ __forceinline__ void publish_tagged(uint4 *box, const fr_t &a, const fr_t &b, const fr_t &c, uint32_t seq) {
    uint32_t w[24];
#pragma unroll
    for (int j = 0; j < 8; ++j) { w[j] = a.v[j]; w[8 + j] = b.v[j]; w[16 + j] = c.v[j]; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#if ZK_ON_DEVICE
        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(box + j), "r"(w[3 * j]), "r"(w[3 * j + 1]), "r"(w[3 * j + 2]), "r"(seq) : "memory");
#else
        uint4 q;
        q.x = w[3 * j]; q.y = w[3 * j + 1]; q.z = w[3 * j + 2]; q.w = seq;
        box[j] = q;
#endif
    }
}
// make the results visible to the host, then raise the sequence number it is spinning on
__device__ __forceinline__ void publish(uint32_t *flag, uint32_t seq) {
#if ZK_ON_DEVICE
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(flag) = seq;
#else
    *flag = seq;
#endif
}

// Last CTA of a round kernel: sh_tot holds the three exact limb sums (A, C, E; 17 limbs each, weight 2^(32 j)).
// Thread (k, c) turns chunk c of sum k into its share of the field element, three threads add up, thread 0 publishes
// (a, b, c) = (A, E - A - C, C).
__device__ __forceinline__ void round_quad_publish(const round_args_t &A, const unsigned long long *sh_tot, fr_t *sh_fr /* [12] */) {
    if (threadIdx.x < 9) {
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
        while (fr_t::ge_raw(x.v, pm)) fr_t::raw_sub(x.v, pm);   // the multiplier wants operands below r (at most two steps)
        st_fr(sh_fr + threadIdx.x, x * y);
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const fr_t s = sh_fr[3 * threadIdx.x] + sh_fr[3 * threadIdx.x + 1] + sh_fr[3 * threadIdx.x + 2];
        st_fr(sh_fr + 9 + threadIdx.x, s);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const fr_t a = sh_fr[9], c = sh_fr[10], e = sh_fr[11];
        if (A.tagged) publish_tagged(A.tagged, a, e - a - c, c, A.seq);
        else {
            st_fr(A.out + 0, a);
            st_fr(A.out + 1, e - a - c);
            st_fr(A.out + 2, c);
            if (A.flag) publish(A.flag, A.seq);
        }
    }
}

// The per-pair version for the streaming kernels: sh_tot holds [pair][A, C, E][17 limbs]; sh_fr needs 26 elements.
__device__ __forceinline__ void round_quad_publish_pairs(const round_args_t &A, const unsigned long long *sh_tot, fr_t *sh_fr /* [26] */) {
    if (threadIdx.x < 18) {
        const int k = threadIdx.x / 3, c = threadIdx.x % 3;   // k = pair * 3 + sum
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
    if (threadIdx.x < 6) {   // sum k of pair p
        const fr_t s = sh_fr[3 * threadIdx.x] + sh_fr[3 * threadIdx.x + 1] + sh_fr[3 * threadIdx.x + 2];
        st_fr(sh_fr + 18 + threadIdx.x, s);
    }
    __syncthreads();
    if (threadIdx.x < 2) {   // one thread per pair: (a, b, c) of the pair, kept for the next round
        const int p = threadIdx.x;
        const fr_t a = sh_fr[18 + 3 * p], c = sh_fr[19 + 3 * p], e = sh_fr[20 + 3 * p];
        fr_t b;
        if (A.derive_b[p]) {
            const fr_t pa = ld_fr(A.state + 4 * p), pb = ld_fr(A.state + 4 * p + 1), pc = ld_fr(A.state + 4 * p + 2);
            const fr_t at_r = (pa * A.r + pb) * A.r + pc;     // P_prev(r) = P(0) + P(1)
            b = at_r - c - c - a;
        } else b = e - a - c;
        if (A.pair[p].n_blocks) {
            st_fr(A.state + 4 * p, a);
            st_fr(A.state + 4 * p + 1, b);
            st_fr(A.state + 4 * p + 2, c);
        }
        st_fr(sh_fr + 24 + p, b);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const fr_t a = sh_fr[18] + sh_fr[21], b = sh_fr[24] + sh_fr[25], c = sh_fr[19] + sh_fr[22];
        if (A.tagged) publish_tagged(A.tagged, a, b, c, A.seq);
        else {
            st_fr(A.out + 0, a);
            st_fr(A.out + 1, b);
            st_fr(A.out + 2, c);
            if (A.flag) publish(A.flag, A.seq);
        }
    }
}

// Per output pair (v0,v1),(m0,m1) of the (folded) tables the round polynomial contributes
//   a += (m1-m0)(v1-v0),  c += m0 v0,  b += (m1-m0) v0 + m0 (v1-v0) = m1 v1 - a - c
// (linear_poly * linear_poly, src/polynomial.cpp:116-118, evaluated Karatsuba-style with 3 products).  The three sums
// are accumulated unreduced (fr_lazy_t): the four fold multiplications per output pair are full Montgomery
// multiplications (their results are stored), the three products only pay the multiplication half.
__global__ void __launch_bounds__(kRoundBlock, ZK_ROUND_CTAS_PER_SM) k_round_quad(round_args_t A) {
    ZK_PDL_ENTRY();
    __shared__ unsigned long long sh_warp[(kRoundBlock / 32) * kRoundLimbs];
    __shared__ unsigned long long sh_tot[2 * kRoundLimbs];
    __shared__ uint32_t sh_ticket;
    __shared__ fr_t sh_fr[26];
    const uint32_t nb0 = A.pair[0].n_blocks, nb = nb0 + A.pair[1].n_blocks;
    const bool second = blockIdx.x >= nb0;
    const fr_t *v_in = second ? A.pair[1].v_in : A.pair[0].v_in, *m_in = second ? A.pair[1].m_in : A.pair[0].m_in;
    fr_t *v_out = second ? A.pair[1].v_out : A.pair[0].v_out, *m_out = second ? A.pair[1].m_out : A.pair[0].m_out;
    const uint32_t n_in = second ? A.pair[1].n_in : A.pair[0].n_in, live = second ? A.pair[1].live : A.pair[0].live;
    const uint32_t fold = second ? A.pair[1].fold : A.pair[0].fold;
    const uint32_t bx = second ? blockIdx.x - nb0 : blockIdx.x;
    const uint32_t stride = (second ? A.pair[1].n_blocks : nb0) * kRoundBlock;
    const bool derive = (second ? A.derive_b[1] : A.derive_b[0]) != 0;   // b comes from the previous round: no E
    fr_lazy_t acc[3];  // A, C, E
    acc[0].clear(); acc[1].clear(); acc[2].clear();
    if (fold) {
        const uint32_t n_pairs = n_in >> 2;
        const uint32_t live_pairs = (live + 3) >> 2;
        const fr_t r = A.r;
        for (uint32_t i = bx * kRoundBlock + threadIdx.x; i < n_pairs && i < live_pairs; i += stride) {
            const uint32_t base = i << 2;
            fr_t x0 = ld_fr_live(v_in, base, live), x1 = ld_fr_live(v_in, base + 1, live);
            fr_t x2 = ld_fr_live(v_in, base + 2, live), x3 = ld_fr_live(v_in, base + 3, live);
            fr_t y0 = ld_fr_live(m_in, base, live), y1 = ld_fr_live(m_in, base + 1, live);
            fr_t y2 = ld_fr_live(m_in, base + 2, live), y3 = ld_fr_live(m_in, base + 3, live);
            const fr_t v0 = x0 + r * fr_t::sub_lazy(x1, x0);
            const fr_t v1 = x2 + r * fr_t::sub_lazy(x3, x2);
            st_fr_g(v_out + 2 * i, v0);
            st_fr_g(v_out + 2 * i + 1, v1);
            const fr_t m0 = y0 + r * fr_t::sub_lazy(y1, y0);
            const fr_t m1 = y2 + r * fr_t::sub_lazy(y3, y2);
            st_fr_g(m_out + 2 * i, m0);
            st_fr_g(m_out + 2 * i + 1, m1);
            acc[0].mac(fr_t::sub_lazy(m1, m0), fr_t::sub_lazy(v1, v0));
            acc[1].mac(m0, v0);
            if (!derive) acc[2].mac(m1, v1);
        }
    } else {
        const uint32_t n_pairs = n_in >> 1;
        const uint32_t live_pairs = (live + 1) >> 1;
        for (uint32_t i = bx * kRoundBlock + threadIdx.x; i < n_pairs && i < live_pairs; i += stride) {
            const fr_t v0 = ld_fr_live(v_in, 2 * i, live), v1 = ld_fr_live(v_in, 2 * i + 1, live);
            const fr_t m0 = ld_fr_live(m_in, 2 * i, live), m1 = ld_fr_live(m_in, 2 * i + 1, live);
            acc[0].mac(fr_t::sub_lazy(m1, m0), fr_t::sub_lazy(v1, v0));
            acc[1].mac(m0, v0);
            acc[2].mac(m1, v1);
        }
    }
    uint32_t limb[kRoundLimbs];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < fr_lazy_t::W; ++j) limb[k * fr_lazy_t::W + j] = acc[k].w[j];
    if (!grid_limb_sum_slice<kRoundLimbs, 2 * kRoundLimbs, kRoundBlock>(limb, second ? kRoundLimbs : 0, A.acc, A.counter, nb, sh_warp, sh_tot, &sh_ticket)) return;
    round_quad_publish_pairs(A, sh_tot, sh_fr);
}

// K1, latency-bound rounds (tables up to 2^16 entries): the seven multiplications of an output pair are spread over
// FOUR lanes, so the critical path of a round is one fold multiplication + one unreduced product instead of seven
// multiplications in a row.  Lane roles within a quad (lanes 4q .. 4q+3 work on output pair q):
//   fold round:   role 0: v0 = fold(V[4q], V[4q+1])   role 1: v1 = fold(V[4q+2], V[4q+3])   role 2: m0   role 3: m1
//                 then three shuffle exchanges hand role 0 (v0, m0) -> C, role 1 (v1, m1) -> E, role 2 (dm, dv) -> A
//   first round:  role 0: (v0, m0) -> C   role 1: (v1, m1) -> E   role 2: loads all four -> A
// Each lane's unreduced product goes through role-masked redux into the same exact limb sums as k_round_quad.
__global__ void __launch_bounds__(kRoundBlock) k_round_quad_thin(round_args_t A) {
    __shared__ unsigned long long sh_warp[(kRoundBlock / 32) * kRoundLimbs];
    __shared__ unsigned long long sh_tot[kRoundLimbs];
    __shared__ uint32_t sh_ticket;
    __shared__ fr_t sh_fr[12];
    ZK_PDL_ENTRY();
    const uint32_t nb0 = A.pair[0].n_blocks, nb = nb0 + A.pair[1].n_blocks;
    const bool second = blockIdx.x >= nb0;
    const fr_t *v_in = second ? A.pair[1].v_in : A.pair[0].v_in, *m_in = second ? A.pair[1].m_in : A.pair[0].m_in;
    fr_t *v_out = second ? A.pair[1].v_out : A.pair[0].v_out, *m_out = second ? A.pair[1].m_out : A.pair[0].m_out;
    const uint32_t n_in = second ? A.pair[1].n_in : A.pair[0].n_in, live = second ? A.pair[1].live : A.pair[0].live;
    const uint32_t fold = second ? A.pair[1].fold : A.pair[0].fold;
    const uint32_t bx = second ? blockIdx.x - nb0 : blockIdx.x;
    const uint32_t stride = (second ? A.pair[1].n_blocks : nb0) * (kRoundBlock / 4);   // output pairs per grid pass
    const uint32_t role = threadIdx.x & 3u, lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lim_n = fold ? n_in >> 2 : n_in >> 1, lim_l = fold ? (live + 3) >> 2 : (live + 1) >> 1;
    const uint32_t limit = lim_n < lim_l ? lim_n : lim_l;
    fr_lazy_t acc;
    acc.clear();
    const fr_t r = A.r;
    // the trip count is CTA-uniform (the CTA's first pair decides): every lane takes part in the shuffles
    for (uint32_t q0 = bx * (kRoundBlock / 4); q0 < limit; q0 += stride) {
        const uint32_t q = q0 + (threadIdx.x >> 2);
        const bool on = q < limit;
        fr_t a_op = fr_t::zero(), b_op = fr_t::zero();
        if (fold) {
            const fr_t *src = role < 2 ? v_in : m_in;
            const uint32_t idx = 4 * q + 2 * (role & 1u);
            const fr_t x0 = on ? ld_fr_live(src, idx, live) : fr_t::zero(), x1 = on ? ld_fr_live(src, idx + 1, live) : fr_t::zero();
            const fr_t y = x0 + r * fr_t::sub_lazy(x1, x0);
            if (on) st_fr_g((role < 2 ? v_out : m_out) + 2 * q + (role & 1u), y);
            fr_t o1, o2, d, d2;
#pragma unroll
            for (int j = 0; j < 8; ++j) o1.v[j] = __shfl_xor_sync(0xffffffffu, y.v[j], 1);
            d = (role & 1u) ? fr_t::sub_lazy(y, o1) : fr_t::sub_lazy(o1, y);   // (entry 1) - (entry 0) of this lane's table
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                o2.v[j] = __shfl_xor_sync(0xffffffffu, y.v[j], 2);
                d2.v[j] = __shfl_xor_sync(0xffffffffu, d.v[j], 2);
            }
            if (role < 2) { a_op = y; b_op = o2; }          // v0 m0 (role 0), v1 m1 (role 1)
            else if (role == 2) { a_op = d; b_op = d2; }    // (m1 - m0)(v1 - v0)
        } else if (on) {
            if (role < 2) {
                a_op = ld_fr_live(v_in, 2 * q + role, live);
                b_op = ld_fr_live(m_in, 2 * q + role, live);
            } else if (role == 2) {
                a_op = fr_t::sub_lazy(ld_fr_live(v_in, 2 * q + 1, live), ld_fr_live(v_in, 2 * q, live));
                b_op = fr_t::sub_lazy(ld_fr_live(m_in, 2 * q + 1, live), ld_fr_live(m_in, 2 * q, live));
            }
        }
        acc.mac(a_op, b_op);
    }
    // warp sums per role: role 2 -> A (slot 0), role 0 -> C (slot 1), role 1 -> E (slot 2); role 3 carries zeros.  Full-mask
    // redux with the other roles' lanes contributing zero: member masks narrower than the warp are emulated in software
    // (measured 420 clk per redux against 12.5, tools/latbench.cu).
#pragma unroll
    for (int sl = 0; sl < 3; ++sl) {
        const bool mine = role == (sl == 0 ? 2u : sl == 1 ? 0u : 1u);
#pragma unroll
        for (int k = 0; k < fr_lazy_t::W; ++k) {
            const uint32_t x = mine ? acc.w[k] : 0u;
            const uint32_t lo = __reduce_add_sync(0xffffffffu, x & 0xffffu);
            const uint32_t hi = __reduce_add_sync(0xffffffffu, x >> 16);
            if (lane == (uint32_t) (k & 31)) sh_warp[warp * kRoundLimbs + sl * fr_lazy_t::W + k] = (unsigned long long) lo + ((unsigned long long) hi << 16);
        }
    }
    if (!grid_limb_sum_finish<kRoundLimbs, kRoundBlock>(A.acc, A.counter, nb, sh_warp, sh_tot, &sh_ticket)) return;
    round_quad_publish(A, sh_tot, sh_fr);
}

// --------------------------------------------------------------------------------------------------------------------
// K1, the TAIL of a phase in ONE launch (batched phases only: every challenge of the phase is known up front).  Once both
// table pairs are down to kTailMaxEntries entries, a single CTA keeps them in shared memory and runs all remaining rounds:
// per round the four-lane split of k_round_quad_thin, the exact limb sums inside the CTA, the collapse of an exhausted pair
// (src/prover.cpp:400-404) and, after the last round, the (at most two) surviving entries of each table written back for the
// Finalize call.  Replaces ~10 dependent launches per phase (about a thousand per vgg11 proof) by one.
// --------------------------------------------------------------------------------------------------------------------
constexpr int kTailBlock = 512;
constexpr uint32_t kTailMaxEntries = 1024;
constexpr int kTailMaxRounds = 12;
constexpr uint32_t kTailSmemBytes = 4 * (kTailMaxEntries + kTailMaxEntries / 2) * sizeof(fr_t);   // per table: buffer A (1024) + buffer B (512)

struct tail_args_t {
    const fr_t *v_in[2], *m_in[2];   // tables at the start of the tail: n_in entries, the first `live` non-zero
    uint32_t n_in[2], live[2];       // n_in == 0: the pair takes no part
    uint32_t first;                  // the first tail round is the first round of the phase (previous_random = 0): no fold
    uint32_t n_rounds;
    fr_t r[kTailMaxRounds];          // previous_random of each tail round
    fr_t *slots;                     // [n_rounds][16] result blocks: (a, b, c) at 0, collapse values (v, m) of pair p at 8 + 2p
    fr_t *v_out[2], *m_out[2];       // the entries left after the last round (<= 2 per table)
};

__global__ void __launch_bounds__(kTailBlock, 1) k_round_tail(tail_args_t A) {
    ZK_DYN_SMEM(fr_t, sm);
    __shared__ unsigned long long sh_warp[(kTailBlock / 32) * kRoundLimbs];
    __shared__ unsigned long long sh_tot[kRoundLimbs];
    __shared__ fr_t sh_fr[12];
    ZK_PDL_ENTRY();
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, role = tid & 3u, grp = tid >> 2;
    constexpr uint32_t kGroups = kTailBlock / 4;
    // table t = 2 * pair + (0: V, 1: mult); buffer A of table t at sm + t * 1536, buffer B behind it
    fr_t *cur[4], *nxt[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) { cur[t] = sm + t * (kTailMaxEntries + kTailMaxEntries / 2); nxt[t] = cur[t] + kTailMaxEntries; }
    uint32_t n[2] = {A.n_in[0], A.n_in[1]};
    for (int p = 0; p < 2; ++p)
        for (uint32_t i = tid; i < n[p]; i += kTailBlock) {
            st_fr(cur[2 * p] + i, ld_fr_live(A.v_in[p], i, A.live[p]));
            st_fr(cur[2 * p + 1] + i, ld_fr_live(A.m_in[p], i, A.live[p]));
        }
    __syncthreads();
    for (uint32_t j = 0; j < A.n_rounds; ++j) {
        const bool fold = !(A.first && j == 0);
        const fr_t r = A.r[j];
        fr_t *slot = A.slots + (size_t) j * 16;
        uint32_t Q[2], fin[2];
        for (int p = 0; p < 2; ++p) {
            const uint32_t n_after = n[p] ? (fold ? n[p] >> 1 : n[p]) : 0;
            fin[p] = n_after == 1;
            Q[p] = fin[p] ? 0 : n_after >> 1;
        }
        // collapse of an exhausted pair: its two tables evaluated at r (four lanes of warp 15 -- idle in every round that has one)
        if (tid >= kTailBlock - 4) {
            const int t = tid - (kTailBlock - 4), p = t >> 1;
            if (fin[p]) {
                fr_t x0 = ld_fr(cur[t]);
                if (fold) { const fr_t x1 = ld_fr(cur[t] + 1); x0 = x0 + r * (x1 - x0); }
                st_fr(slot + 8 + 2 * p + (t & 1), x0);
            }
        }
        const uint32_t tasks = Q[0] + Q[1];
        fr_lazy_t acc;
        acc.clear();
        for (uint32_t t0 = 0; t0 < tasks; t0 += kGroups) {   // CTA-uniform trip count: every lane takes part in the shuffles
            const uint32_t t = t0 + grp;
            const bool on = t < tasks;
            const int p = (on && t >= Q[0]) ? 1 : 0;
            const uint32_t q = on ? (p ? t - Q[0] : t) : 0;
            fr_t a_op = fr_t::zero(), b_op = fr_t::zero();
            if (fold) {
                const fr_t *src = cur[2 * p + (role >> 1)];          // roles 0, 1: the V table, roles 2, 3: the mult table
                const uint32_t idx = 4 * q + 2 * (role & 1u);
                const fr_t x0 = on ? ld_fr(src + idx) : fr_t::zero(), x1 = on ? ld_fr(src + idx + 1) : fr_t::zero();
                const fr_t y = x0 + r * fr_t::sub_lazy(x1, x0);
                if (on) st_fr(nxt[2 * p + (role >> 1)] + 2 * q + (role & 1u), y);
                fr_t o1, o2, d, d2;
#pragma unroll
                for (int k = 0; k < 8; ++k) o1.v[k] = __shfl_xor_sync(0xffffffffu, y.v[k], 1);
                d = (role & 1u) ? fr_t::sub_lazy(y, o1) : fr_t::sub_lazy(o1, y);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    o2.v[k] = __shfl_xor_sync(0xffffffffu, y.v[k], 2);
                    d2.v[k] = __shfl_xor_sync(0xffffffffu, d.v[k], 2);
                }
                if (role < 2) { a_op = y; b_op = o2; }          // v0 m0 (role 0), v1 m1 (role 1)
                else if (role == 2) { a_op = d; b_op = d2; }    // (m1 - m0)(v1 - v0)
            } else if (on) {
                const fr_t *v = cur[2 * p], *m = cur[2 * p + 1];
                if (role < 2) { a_op = ld_fr(v + 2 * q + role); b_op = ld_fr(m + 2 * q + role); }
                else if (role == 2) {
                    a_op = fr_t::sub_lazy(ld_fr(v + 2 * q + 1), ld_fr(v + 2 * q));
                    b_op = fr_t::sub_lazy(ld_fr(m + 2 * q + 1), ld_fr(m + 2 * q));
                }
            }
            acc.mac(a_op, b_op);
        }
        // exact limb sums: role 2 -> A, role 0 -> C, role 1 -> E; only the warps that had a task take part
        const uint32_t active_groups = tasks < kGroups ? tasks : kGroups;
        const uint32_t active_warps = (active_groups * 4 + 31) >> 5;
#ifdef ZK_EMU
        const bool take_part = true;    // (the test emulator implements warp collectives with a CTA-wide exchange: every thread must call them)
#else
        const bool take_part = warp < active_warps;
#endif
        if (take_part) {
#pragma unroll
            for (int sl = 0; sl < 3; ++sl) {
                const bool mine = role == (sl == 0 ? 2u : sl == 1 ? 0u : 1u);
#pragma unroll
                for (int k = 0; k < fr_lazy_t::W; ++k) {
                    const uint32_t x = mine ? acc.w[k] : 0u;
                    const uint32_t lo = __reduce_add_sync(0xffffffffu, x & 0xffffu);
                    const uint32_t hi = __reduce_add_sync(0xffffffffu, x >> 16);
                    if (lane == (uint32_t) (k & 31)) sh_warp[warp * kRoundLimbs + sl * fr_lazy_t::W + k] = (unsigned long long) lo + ((unsigned long long) hi << 16);
                }
            }
        }
        __syncthreads();
        if (tid < kRoundLimbs) {
            unsigned long long t = 0;
            for (uint32_t w = 0; w < active_warps; ++w) t += sh_warp[w * kRoundLimbs + tid];
            sh_tot[tid] = t;
        }
        __syncthreads();
        round_args_t R;
        R.out = slot;
        R.flag = nullptr;
        R.tagged = nullptr;
        R.seq = 0;
        round_quad_publish(R, sh_tot, sh_fr);    // (a, b, c) = (A, E - A - C, C) -> slot[0..2]
        for (int p = 0; p < 2; ++p) {
            if (fin[p]) n[p] = 0;
            else if (fold && n[p]) {
                n[p] >>= 1;
                fr_t *t0 = cur[2 * p], *t1 = cur[2 * p + 1];
                cur[2 * p] = nxt[2 * p]; cur[2 * p + 1] = nxt[2 * p + 1];
                nxt[2 * p] = t0; nxt[2 * p + 1] = t1;
            }
        }
        __syncthreads();   // the folded tables and sh_* are settled before the next round touches them
    }
    for (int p = 0; p < 2; ++p)
        if (tid < n[p]) {
            st_fr_g(A.v_out[p] + tid, ld_fr(cur[2 * p] + tid));
            st_fr_g(A.m_out[p] + tid, ld_fr(cur[2 * p + 1] + tid));
        }
}

#if !defined(ZK_EMU)
// --------------------------------------------------------------------------------------------------------------------
// K1, HBM-streaming rounds (tables of 2^17 entries and more): the same arithmetic as k_round_quad, fed by TMA.
// A warp owns a double-buffered pair of 4 KB boxes per table: 32 rows of 128 bytes, one row = the four input entries of
// one output pair.  One elected lane issues cp.async.bulk.tensor for the NEXT row block of V and of mult while the warp
// multiplies the current one; the boxes land in shared memory with the 128-byte swizzle (16-byte chunk c of row q sits at
// chunk c ^ (q & 7)), so each lane reads its own row with conflict-free LDS.128 and no LSU/L1 traffic or registers are
// spent on data in flight.  Completion is tracked by one mbarrier per (warp, stage) with a transaction count of 8 KB.
// Row blocks that are not completely live (the last one of a table) take the guarded global-load path.
// --------------------------------------------------------------------------------------------------------------------
// (compile-time variants measured in round 1, all within 2 % of each other at 2^24 entries: one stage refilled as soon as the
//  block is in registers with 4 or 5 CTAs per SM, the out-of-line multiplier; the arithmetic, not the feeding, sets the pace)
#ifndef ZK_TMA_STAGES
#define ZK_TMA_STAGES 2
#endif
#ifndef ZK_TMA_CTAS
#define ZK_TMA_CTAS 3
#endif
constexpr int kTmaStages = ZK_TMA_STAGES;   // 2: next block lands while this one is multiplied; 1: refill as soon as the block is in registers
constexpr int kTmaCtasPerSm = ZK_TMA_CTAS;
constexpr int kTmaMaxGrid = ZK_SM_COUNT * kTmaCtasPerSm;
constexpr uint32_t kTmaBoxBytes = 4096;                                     // 32 rows x 128 bytes
constexpr uint32_t kTmaWarpBytes = kTmaStages * 2 /* tables */ * kTmaBoxBytes;
constexpr uint32_t kTmaSmemBytes = (kRoundBlock / 32) * kTmaWarpBytes + 1024;   // + slack to align the boxes to 1 KB

struct alignas(64) round_tma_args_t {
    round_args_t R;
    CUtensorMap tm[2][2];    // [pair][V, mult]: the input tables as (n_in / 4) rows of 32 x u32, box 32 x 32, SWIZZLE_128B
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_rows(uint32_t dst, const CUtensorMap *tm, uint32_t row0, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(0), "r"(row0), "r"(bar) : "memory");
}
__device__ __forceinline__ fr_t lds_fr_swz(uint32_t row_base, uint32_t row, int entry) {   // entry 0..3 of the row
    fr_t x;
    const uint32_t a0 = row_base + (((2 * entry) ^ (row & 7u)) << 4), a1 = row_base + (((2 * entry + 1) ^ (row & 7u)) << 4);
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.v[0]), "=r"(x.v[1]), "=r"(x.v[2]), "=r"(x.v[3]) : "r"(a0));
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x.v[4]), "=r"(x.v[5]), "=r"(x.v[6]), "=r"(x.v[7]) : "r"(a1));
    return x;
}
// the multiplier inlined (the four fold multiplications of an output pair are independent: their carry chains interleave)
__device__ __forceinline__ fr_t fr_mul_inline(const fr_t &a, const fr_t &b) {
    fr_t r;
    fr_t::mul_wide(r.v, a.v, b.v);
    return r;
}
#ifdef ZK_TMA_MUL_CALL
#define ZK_TMA_MUL(a, b) ((a) * (b))
#else
#define ZK_TMA_MUL(a, b) fr_mul_inline(a, b)
#endif

__global__ void __launch_bounds__(kRoundBlock, kTmaCtasPerSm) k_round_quad_tma(const __grid_constant__ round_tma_args_t T) {
    extern __shared__ __align__(1024) unsigned char dyn_smem[];
    __shared__ unsigned long long sh_warp[(kRoundBlock / 32) * kRoundLimbs];
    __shared__ unsigned long long sh_tot[2 * kRoundLimbs];
    __shared__ uint32_t sh_ticket;
    __shared__ fr_t sh_fr[26];
    __shared__ __align__(8) unsigned long long sh_bar[(kRoundBlock / 32) * 2];
    const round_args_t &A = T.R;
    const uint32_t nb0 = A.pair[0].n_blocks, nb = nb0 + A.pair[1].n_blocks;
    const bool second = blockIdx.x >= nb0;
    const fr_t *v_in = second ? A.pair[1].v_in : A.pair[0].v_in, *m_in = second ? A.pair[1].m_in : A.pair[0].m_in;
    fr_t *v_out = second ? A.pair[1].v_out : A.pair[0].v_out, *m_out = second ? A.pair[1].m_out : A.pair[0].m_out;
    const uint32_t n_in = second ? A.pair[1].n_in : A.pair[0].n_in, live = second ? A.pair[1].live : A.pair[0].live;
    const CUtensorMap *tm_v = second ? &T.tm[1][0] : &T.tm[0][0], *tm_m = second ? &T.tm[1][1] : &T.tm[0][1];
    const uint32_t bx = second ? blockIdx.x - nb0 : blockIdx.x;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t n_pairs = n_in >> 2, live_pairs = (live + 3) >> 2;
    const uint32_t limit = n_pairs < live_pairs ? n_pairs : live_pairs;        // output pairs with any live input
    const uint32_t n_groups = (limit + 31) >> 5;                               // row blocks of 32 output pairs
    const uint32_t full_groups = (live >> 2) >> 5 < n_groups ? (live >> 2) >> 5 : n_groups;   // blocks whose 128 inputs are all live
    const uint32_t gstride = (second ? A.pair[1].n_blocks : nb0) * (kRoundBlock / 32);
    const bool derive = (second ? A.derive_b[1] : A.derive_b[0]) != 0;   // b comes from the previous round: no E

    const uint32_t box0 = ((smem_addr(dyn_smem) + 1023u) & ~1023u) + warp * kTmaWarpBytes;   // [stage][table] boxes of this warp
    const uint32_t bar0 = smem_addr(sh_bar + 2 * warp);
    if (lane == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    fr_lazy_t acc[3];  // A, C, E
    acc[0].clear(); acc[1].clear(); acc[2].clear();
    const fr_t r = A.r;
    uint32_t stage = 0, parity = 0;   // bit s of `parity`: phase to wait for on stage s
    uint32_t g = bx * (kRoundBlock / 32) + warp;
    if (g < full_groups && lane == 0) {
        mbar_expect_tx(bar0, 2 * kTmaBoxBytes);
        tma_load_rows(box0, tm_v, g * 32, bar0);
        tma_load_rows(box0 + kTmaBoxBytes, tm_m, g * 32, bar0);
    }
    for (; g < n_groups; g += gstride) {
        const uint32_t gn = g + gstride;
        if (kTmaStages == 2) {
            __syncwarp();   // every lane is done reading the other stage (previous iteration)
            if (gn < full_groups && lane == 0) {
                const uint32_t b = bar0 + 8 * (stage ^ 1u), dst = box0 + (stage ^ 1u) * 2 * kTmaBoxBytes;
                mbar_expect_tx(b, 2 * kTmaBoxBytes);
                tma_load_rows(dst, tm_v, gn * 32, b);
                tma_load_rows(dst + kTmaBoxBytes, tm_m, gn * 32, b);
            }
        }
        const uint32_t i = g * 32 + lane;   // output pair of this lane
        fr_t x0, x1, x2, x3, y0, y1, y2, y3;
        if (g < full_groups) {
            const uint32_t st = kTmaStages == 2 ? stage : 0u;
            mbar_wait(bar0 + 8 * st, (parity >> st) & 1u);
            parity ^= 1u << st;
            const uint32_t vrow = box0 + st * 2 * kTmaBoxBytes + lane * 128u, mrow = vrow + kTmaBoxBytes;
            x0 = lds_fr_swz(vrow, lane, 0); x1 = lds_fr_swz(vrow, lane, 1); x2 = lds_fr_swz(vrow, lane, 2); x3 = lds_fr_swz(vrow, lane, 3);
            y0 = lds_fr_swz(mrow, lane, 0); y1 = lds_fr_swz(mrow, lane, 1); y2 = lds_fr_swz(mrow, lane, 2); y3 = lds_fr_swz(mrow, lane, 3);
        } else {
            const uint32_t base = i << 2;
            const bool on = i < limit;
            x0 = on ? ld_fr_live(v_in, base, live) : fr_t::zero(); x1 = on ? ld_fr_live(v_in, base + 1, live) : fr_t::zero();
            x2 = on ? ld_fr_live(v_in, base + 2, live) : fr_t::zero(); x3 = on ? ld_fr_live(v_in, base + 3, live) : fr_t::zero();
            y0 = on ? ld_fr_live(m_in, base, live) : fr_t::zero(); y1 = on ? ld_fr_live(m_in, base + 1, live) : fr_t::zero();
            y2 = on ? ld_fr_live(m_in, base + 2, live) : fr_t::zero(); y3 = on ? ld_fr_live(m_in, base + 3, live) : fr_t::zero();
        }
        const fr_t dx0 = fr_t::sub_lazy(x1, x0), dx1 = fr_t::sub_lazy(x3, x2), dy0 = fr_t::sub_lazy(y1, y0), dy1 = fr_t::sub_lazy(y3, y2);
        if (kTmaStages == 1) {
            // single buffer: the subtractions above consumed all sixteen loaded registers, so every lane's LDS has
            // completed (in-order issue); `dep` makes the refill depend on them so that it cannot be scheduled earlier
            uint32_t dep = dx0.v[0] ^ dx1.v[0] ^ dy0.v[0] ^ dy1.v[0] ^ dx0.v[7] ^ dx1.v[7] ^ dy0.v[7] ^ dy1.v[7];
            dep = (uint32_t) __popc(dep) >> 6;   // always 0, opaque to the compiler
            __syncwarp();
            if (gn < full_groups && lane == 0) {
                mbar_expect_tx(bar0, 2 * kTmaBoxBytes);
                tma_load_rows(box0, tm_v, gn * 32 + dep, bar0);
                tma_load_rows(box0 + kTmaBoxBytes, tm_m, gn * 32 + dep, bar0);
            }
        }
        const fr_t v0 = x0 + ZK_TMA_MUL(r, dx0);
        const fr_t v1 = x2 + ZK_TMA_MUL(r, dx1);
        const fr_t m0 = y0 + ZK_TMA_MUL(r, dy0);
        const fr_t m1 = y2 + ZK_TMA_MUL(r, dy1);
        if (i < limit) {
            st_fr(v_out + 2 * i, v0);
            st_fr(v_out + 2 * i + 1, v1);
            st_fr(m_out + 2 * i, m0);
            st_fr(m_out + 2 * i + 1, m1);
        }
        acc[0].mac(fr_t::sub_lazy(m1, m0), fr_t::sub_lazy(v1, v0));
        acc[1].mac(m0, v0);
        if (!derive) acc[2].mac(m1, v1);
        stage ^= 1u;
    }
    uint32_t limb[kRoundLimbs];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < fr_lazy_t::W; ++j) limb[k * fr_lazy_t::W + j] = acc[k].w[j];
    if (!grid_limb_sum_slice<kRoundLimbs, 2 * kRoundLimbs, kRoundBlock>(limb, second ? kRoundLimbs : 0, A.acc, A.counter, nb, sh_warp, sh_tot, &sh_ticket)) return;
    round_quad_publish_pairs(A, sh_tot, sh_fr);
}
#endif  // !ZK_EMU

// Copy from MAPPED, page-locked host memory by the SMs (zero-copy loads over PCIe) instead of the copy engine: the DMA queue
// serves copies in submission order whatever their stream, so a 10 ms witness prefetch on the copy engine delays every small
// host->device copy of the proof that runs meanwhile; loads issued by a few low-priority CTAs do not.  Four independent
// 16-byte loads per thread keep ~64 KB per CTA in flight.
__global__ void __launch_bounds__(kBlock) k_copy_from_host(uint4 *dst, const uint4 *src, uint64_t n16) {
    const uint64_t stride = (uint64_t) gridDim.x * kBlock;
    uint64_t i = (uint64_t) blockIdx.x * kBlock + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        const uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n16; i += stride) dst[i] = src[i];
}

// Compact witness: most witness values are small signed integers (quantised activations, weights, bit decompositions), so
// the host keeps them as int64 (8 bytes instead of 32 over PCIe) and the device rebuilds the Montgomery form: one
// multiplication by R^2 per element.  `in` may be mapped host memory (zero-copy loads, see k_copy_from_host); four
// independent loads per thread keep the link busy.  Values that do not fit go through k_scatter_fr afterwards.
__global__ void __launch_bounds__(kBlock) k_expand_i64(fr_t *out, const long long *in, uint64_t n) {
    const uint64_t stride = (uint64_t) gridDim.x * kBlock;
    uint64_t i = (uint64_t) blockIdx.x * kBlock + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const long long a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
        st_fr_g(out + i, fr_t::from_i64(a));
        st_fr_g(out + i + stride, fr_t::from_i64(b));
        st_fr_g(out + i + 2 * stride, fr_t::from_i64(c));
        st_fr_g(out + i + 3 * stride, fr_t::from_i64(d));
    }
    for (; i < n; i += stride) st_fr_g(out + i, fr_t::from_i64(in[i]));
}
__global__ void __launch_bounds__(kBlock) k_scatter_fr(fr_t *out, const uint32_t *idx, const fr_t *val, uint32_t n) {
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) st_fr_g(out + idx[i], ld_fr_g(val + i));
}

// fold a 2-entry table pair down to single values (the "total == 1" collapse, src/prover.cpp:400-404, and the
// eval(previous_random) of the Finalize calls, src/prover.cpp:146-153,459-497).  out[2*i], out[2*i+1] = v, m of pair i.
struct final_fold_args_t {
    const fr_t *v_in[3], *m_in[3];
    uint32_t live[3];
    uint32_t fold[3];  // 1: n_eval == 2 -> v0 + r (v1 - v0);  0: n_eval == 1 -> copy entry 0
    uint32_t active[3];
    fr_t r;
    fr_t *out;  // [3][2]
    uint32_t *flag;   // see round_args_t
    uint32_t seq;
};
__global__ void k_final_fold(final_fold_args_t A) {
    ZK_PDL_ENTRY();
    const int i = threadIdx.x >> 1, which = threadIdx.x & 1;
    const fr_t *p = (i < 3 && A.active[i]) ? (which ? A.m_in[i] : A.v_in[i]) : nullptr;
    if (p) {
        fr_t x0 = ld_fr_live(p, 0, A.live[i]);
        if (A.fold[i]) {
            fr_t x1 = ld_fr_live(p, 1, A.live[i]);
            x0 = x0 + A.r * (x1 - x0);
        }
        st_fr(A.out + 2 * i + which, x0);
    }
    if (A.flag) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) publish(A.flag, A.seq);
    }
}

// --------------------------------------------------------------------------------------------------------------------
// K3: eq / beta tables.  initHalfTable (src/utils.cpp:32-51): f[0] = init, then for every variable i the table doubles:
// f[j | 2^i] = f[j] r_i, f[j] -= f[j] r_i.  One CTA per half table; both halves (and both points of the 6-argument
// overload) are built by one launch.
// --------------------------------------------------------------------------------------------------------------------
struct half_job_t {
    fr_t *out;           // 2^bits entries
    const fr_t *r;       // bits challenges
    uint32_t bits;
    fr_t init;
};
struct half_args_t { half_job_t job[4]; };
__global__ void __launch_bounds__(kBlock) k_half_tables(half_args_t A) {
    ZK_PDL_ENTRY();
    const half_job_t J = A.job[blockIdx.x];
    if (!J.out) return;
    if (threadIdx.x == 0) st_fr(J.out, J.init);
    __syncthreads();
    for (uint32_t i = 0; i < J.bits; ++i) {
        const uint32_t n = 1u << i;
        const fr_t ri = ld_fr(J.r + i);
        for (uint32_t j = threadIdx.x; j < n; j += kBlock) {
            fr_t x = ld_fr(J.out + j);
            fr_t t = x * ri;
            st_fr(J.out + (j | n), t);
            st_fr(J.out + j, x - t);
        }
        __threadfence();
        __syncthreads();
    }
}

// out[i] = f0[i & mask] s0[i >> fh]  (+ f1[i & mask] s1[i >> fh]),   then  *= tail_scale for i >= tail_start
// (the relu_rou scaling of the bit-decomposition rows, src/prover.cpp:221-222)
struct beta_args_t {
    fr_t *out;
    const fr_t *f0, *s0, *f1, *s1;  // f1 == nullptr: single point
    uint32_t bits, first_half;
    uint32_t tail_start;            // >= 2^bits: no tail scaling
    fr_t tail_scale;
};
__global__ void __launch_bounds__(kBlock) k_beta_expand(beta_args_t A) {
    ZK_PDL_ENTRY();
    const uint32_t n = 1u << A.bits, mask = (1u << A.first_half) - 1;
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
        fr_t x = A.f0 ? ld_fr_g(A.f0 + (i & mask)) * ld_fr_g(A.s0 + (i >> A.first_half)) : fr_t::zero();
        if (A.f1) x = x + ld_fr_g(A.f1 + (i & mask)) * ld_fr_g(A.s1 + (i >> A.first_half));
        if (i >= A.tail_start) x = x * A.tail_scale;
        st_fr_g(A.out + i, x);
    }
}

// PADDING layer: beta_g[g] = beta_g_prev[g >> blh] * beta_gs[g & (2^blh - 1)]   (src/prover.cpp:214-219)
__global__ void __launch_bounds__(kBlock) k_beta_outer(fr_t *out, const fr_t *hi, const fr_t *lo, uint32_t bits, uint32_t blh,
                                                       uint32_t tail_start, fr_t tail_scale) {
    ZK_PDL_ENTRY();
    const uint32_t n = 1u << bits, mask = (1u << blh) - 1;
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
        fr_t x = ld_fr_g(hi + (i >> blh)) * ld_fr_g(lo + (i & mask));
        if (i >= tail_start) x = x * tail_scale;
        st_fr_g(out + i, x);
    }
}

// --------------------------------------------------------------------------------------------------------------------
// K3b: phiGInit (src/utils.cpp:61-103): closed-form MLE of the (I)FFT butterfly network at the point rx.
// One CTA; `pw` holds the 2^n powers of the 2^n-th root of unity (or of its inverse for the IFFT).
// --------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_phi_table(fr_t *phi, const fr_t *rx, const fr_t *pw, fr_t scale, int n, int is_ifft) {
    const fr_t one = fr_t::one();
    if (threadIdx.x == 0) {
        st_fr(phi, scale);
        if (is_ifft) st_fr(phi + 1, scale);
    }
    __threadfence();
    __syncthreads();
    const int i0 = is_ifft ? 2 : 1, i1 = is_ifft ? n : n - 1;
    for (int i = i0; i <= i1; ++i) {
        const uint32_t half = 1u << (i - 1);
        const int m = n - i;
        const fr_t rm = ld_fr(rx + m);
        const fr_t t1 = one - rm;
        for (uint32_t b = threadIdx.x; b < half; b += kBlock) {
            fr_t t2 = rm * ld_fr(pw + ((size_t) b << m));
            fr_t x = ld_fr(phi + b);
            st_fr(phi + (b ^ half), x * (t1 - t2));
            st_fr(phi + b, x * (t1 + t2));
        }
        __threadfence();
        __syncthreads();
    }
    if (!is_ifft) {
        const uint32_t half = 1u << (n - 1);
        const fr_t r0 = ld_fr(rx);
        const fr_t t1 = one - r0;
        for (uint32_t b = threadIdx.x; b < half; b += kBlock) {
            fr_t t2 = r0 * ld_fr(pw + b);
            st_fr(phi + b, ld_fr(phi + b) * (t1 + t2));
        }
    }
}

// --------------------------------------------------------------------------------------------------------------------
// K4/K5: gate gather-reduce.  The circuit topology is static, so at upload time every phase of every layer gets a
// schedule: gates sorted by destination row and cut into work items of <= kItemLen consecutive records; rows that span
// several items are finished by further levels that add up the per-item partial sums.  No atomics, no ordering
// dependence (Fr addition is exact), one thread per item.
// --------------------------------------------------------------------------------------------------------------------
constexpr int kItemLen = 16;   // <= 16: the bound fr_lazy_reduce_upto16 relies on
constexpr int kItemGroup = 32;  // items per record group: record k of item i at item.begin + kItemGroup * k (see build_schedule)

struct gate_rec_t {   // 12 bytes
    uint32_t g;       // output gate -> index into beta_g
    uint32_t x;       // phase 1: index of the v operand (absolute index into val[0], or index into val[l-1]);  phase 2: u
    uint32_t meta;    // bits 0-8: sc (index into two_mul, 0 = multiply by one);  bits 16-17: kind
};
// phase-1 kinds: 0 = uni gate (no value factor), 1 = v operand in layer 0, 2 = v operand in layer l-1
// phase-2 kinds: 0 = scale by V_u0, 1 = scale by V_u1
struct item_t {       // 12 bytes
    uint32_t begin;   // level 0: position of the item's first record (its records are kItemGroup apart); level >= 1: first partial
    uint32_t dest;    // bit 31: 1 = final (index into out table selected by bit 30), 0 = partial slot
    uint32_t count_flags;  // bits 0-15: count;  bits 16-17: kind of the records (phase 2 only)
};
constexpr uint32_t kDestFinal = 0x80000000u, kDestTable1 = 0x40000000u, kDestScalar = 0x20000000u;

struct gate_args_t {
    const gate_rec_t *recs;
    const item_t *items;
    uint32_t n_items;
    const fr_t *beta_g;
    const fr_t *val0, *val_prev;   // phase 1: value sources
    const fr_t *beta_u;            // phase 2
    const fr_t *two_mul;
    fr_t vu[2];                    // phase 2: V_u0, V_u1
    fr_t *out0, *out1;             // mult tables
    fr_t *out_scalar;              // phase 2: add_term accumulator slot
    fr_t *partial;                 // partial sums written by this level
};

__device__ __forceinline__ void store_item(const gate_args_t &A, uint32_t dest, const fr_t &acc) {
    if (dest & kDestFinal) {
        if (dest & kDestScalar) st_fr_g(A.out_scalar, acc);
        else st_fr_g(((dest & kDestTable1) ? A.out1 : A.out0) + (dest & 0x1fffffffu), acc);
    } else st_fr_g(A.partial + dest, acc);
}

// An item's <= 16 products are accumulated unreduced (lazy_acc_t) and reduced once (fr_lazy_reduce_upto16): these kernels
// are bound by the field multiplications (116 M binary gates x 2 phases for vgg11), not by the gathers.
__global__ void __launch_bounds__(kBlock) k_gate_items_p1(gate_args_t A) {
    ZK_PDL_ENTRY();
    for (uint32_t it = blockIdx.x * kBlock + threadIdx.x; it < A.n_items; it += gridDim.x * kBlock) {
        const item_t I = A.items[it];
        const uint32_t cnt = I.count_flags & 0xffffu;
        fr_lazy_t acc;
        acc.clear();
        for (uint32_t k = 0; k < cnt; ++k) {
            const gate_rec_t R = A.recs[I.begin + kItemGroup * k];
            const fr_t bg = ld_fr_g(A.beta_g + R.g);
            const uint32_t kind = (R.meta >> 16) & 3u, sc = R.meta & 0x1ffu;
            if (kind == 0) acc.mac(bg, sc ? ld_fr_g(A.two_mul + sc) : fr_t::one());
            else {
                const fr_t v = ld_fr_g((kind == 1 ? A.val0 : A.val_prev) + R.x);
                if (sc) acc.mac(bg * v, ld_fr_g(A.two_mul + sc));
                else acc.mac(bg, v);
            }
        }
        store_item(A, I.dest, fr_lazy_reduce_upto16(acc));
    }
}

__global__ void __launch_bounds__(kBlock) k_gate_items_p2(gate_args_t A) {
    ZK_PDL_ENTRY();
    for (uint32_t it = blockIdx.x * kBlock + threadIdx.x; it < A.n_items; it += gridDim.x * kBlock) {
        const item_t I = A.items[it];
        const uint32_t cnt = I.count_flags & 0xffffu;
        fr_lazy_t acc;
        acc.clear();
        for (uint32_t k = 0; k < cnt; ++k) {
            const gate_rec_t R = A.recs[I.begin + kItemGroup * k];
            const fr_t bg = ld_fr_g(A.beta_g + R.g), bu = ld_fr_g(A.beta_u + R.x);
            const uint32_t sc = R.meta & 0x1ffu;
            if (sc) acc.mac(bg * bu, ld_fr_g(A.two_mul + sc));
            else acc.mac(bg, bu);
        }
        store_item(A, I.dest, fr_lazy_reduce_upto16(acc) * A.vu[(I.count_flags >> 16) & 1u]);
    }
}

// level >= 1: add up `count` consecutive partial sums of the previous level
__global__ void __launch_bounds__(kBlock) k_sum_partials(gate_args_t A, const fr_t *src) {
    ZK_PDL_ENTRY();
    for (uint32_t it = blockIdx.x * kBlock + threadIdx.x; it < A.n_items; it += gridDim.x * kBlock) {
        const item_t I = A.items[it];
        const uint32_t cnt = I.count_flags & 0xffffu;
        fr_t acc = fr_t::zero();
        for (uint32_t k = 0; k < cnt; ++k) acc = acc + ld_fr_g(src + I.begin + k);
        store_item(A, I.dest, acc);
    }
}

// V table of operands that live in layer 0: V[u] = val[0][ori_id[u]]  (getCirValue, src/prover.cpp:499-501)
__global__ void __launch_bounds__(kBlock) k_gather(fr_t *out, const fr_t *val0, const uint32_t *ori, uint32_t n) {
    ZK_PDL_ENTRY();
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) st_fr_g(out + i, ld_fr_g(val0 + ori[i]));
}

// --------------------------------------------------------------------------------------------------------------------
// K6: Liu init: mult[ori[h]] += beta(r, sigma)[h] for one (layer, side); beta computed on the fly from half tables.
// Within one launch the ori[] entries are distinct (initSubset de-duplicates, src/circuit.cpp:17-23), and launches
// are stream-ordered, so plain read-modify-write is race free.
// --------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_liu_scatter(fr_t *mult, const uint32_t *ori, uint32_t n, const fr_t *f, const fr_t *s,
                                                        uint32_t first_half) {
    ZK_PDL_ENTRY();
    const uint32_t mask = (1u << first_half) - 1;
    for (uint32_t h = blockIdx.x * kBlock + threadIdx.x; h < n; h += gridDim.x * kBlock) {
        fr_t b = ld_fr_g(f + (h & mask)) * ld_fr_g(s + (h >> first_half));
        const uint32_t x = ori[h];
        st_fr_g(mult + x, ld_fr_g(mult + x) + b);
    }
}

// --------------------------------------------------------------------------------------------------------------------
// Vres (src/prover.cpp:434-457): MLE of the (tiny) output layer at r.  One CTA, in-place halving in shared memory.
// --------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_vres(const fr_t *val, uint32_t output_size, const fr_t *r, uint32_t r_size, fr_t *scratch,
                                                 fr_t *out) {
    // scratch holds 2^r_size entries
    const uint32_t whole = 1u << r_size;
    for (uint32_t i = threadIdx.x; i < whole; i += kBlock) st_fr(scratch + i, i < output_size ? ld_fr(val + i) : fr_t::zero());
    __threadfence();
    __syncthreads();
    uint32_t n = whole;
    for (uint32_t i = 0; i < r_size; ++i) {
        const fr_t ri = ld_fr(r + i);
        const uint32_t half = n >> 1;
        // two passes so that in-place writes do not race with reads of later pairs
        for (uint32_t base = 0; base < half; base += kBlock) {
            const uint32_t j = base + threadIdx.x;
            fr_t x;
            if (j < half) {
                fr_t x0 = ld_fr(scratch + 2 * j), x1 = ld_fr(scratch + 2 * j + 1);
                x = x0 + ri * (x1 - x0);
            }
            __syncthreads();
            if (j < half) st_fr(scratch + j, x);
            __threadfence();
            __syncthreads();
        }
        n = half;
    }
    if (threadIdx.x == 0) st_fr(out, ld_fr(scratch));
}

// --------------------------------------------------------------------------------------------------------------------
// element-wise helpers (unit parity tests, Hyrax scalar bookkeeping)
// --------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_fr_binop(const fr_t *a, const fr_t *b, fr_t *out, uint32_t n, int op) {
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
        fr_t x = ld_fr(a + i), y = ld_fr(b + i);
        st_fr(out + i, op == 0 ? x + y : op == 1 ? x - y : x * y);
    }
}

}  // namespace zk
