// Hyrax commitment kernels (hot loop (3) of BASELINE.json north_star).
//
//   K8  k_msm_rowinfo / k_msm_window / k_msm_finish   polyProver::commit -> G1::mulVec
//                                                     3rd/hyrax-bls12-381/src/polyProver.cpp:19-34, mcl ec.hpp:1570-1597
//   K9  k_bullet_scalars / k_dot2 / k_bullet_fold     polyProver::bulletProve / bulletUpdate   polyProver.cpp:76-109
//       (RZ = R^T Z of initBulletProve, polyProver.cpp:66-68, reuses k_dense_colsum of sc_kernels.cuh)
//
// MSM design.  All MSMs of a proof share one generator set (the sqrt(n) Pedersen generators), so a fixed-base window
// table T[w][j] = 2^(8w) * G_j (affine) is built once per generator set.  A scalar is split into sign and magnitude
// (mcl's isNegative convention: x >= (r+1)/2 is handled as -(r-x) with the negated point), and the magnitude into
// unsigned 8-bit digits; digit d of window w adds +-T[w][j] into bucket d.  Because the table already carries the
// 2^(8w) factor, every window's bucket sum is simply added up at the end: no doubling chain.  One CTA handles one
// (row, chunk, window): counting sort of the digits in shared memory, balanced bucket accumulation with mixed adds,
// parallel  sum_b b * B_b.  The zkCNN witness is tiny-valued (hard part 3): windows above the
// row's widest magnitude exit immediately, so a typical row costs one window.
//
// mcl's mulVec is interleaved wNAF (not Pippenger); results agree as GROUP ELEMENTS, which is why every point that
// leaves the device is normalised to affine.
#pragma once
#include "g1.cuh"
#include "sc_kernels.cuh"

namespace zk {

constexpr int kMsmWindows = 32;      // 8-bit digits of a < 2^255 magnitude (|x| <= (r-1)/2 < 2^254)
constexpr int kMsmChunk = 4096;      // entries per CTA
constexpr int kMsmBuckets = 256;

ZK_HD __forceinline__ fp_t ld_fp(const fp_t *p) {
#if ZK_ON_DEVICE
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1], c = q[2];
    fp_t r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    r.v[8] = c.x; r.v[9] = c.y; r.v[10] = c.z; r.v[11] = c.w;
    return r;
#else
    return *p;
#endif
}

// sign / magnitude of a Montgomery-form scalar.  Returns the number of significant bytes of |x| (0 for x == 0).
ZK_HD __forceinline__ uint32_t scalar_sign_mag(const fr_t &s, uint32_t mag[8], uint32_t &neg) {
    uint32_t c[8];
    s.to_canonical(c);
    neg = fr_t::ge_raw(c, fr_cfg::half()) ? 1u : 0u;
    if (neg) {
        const uint32_t *p = fr_cfg::mod();
        int64_t bw = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            bw += (int64_t) p[i] - (int64_t) c[i];
            mag[i] = (uint32_t) bw;
            bw >>= 32;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) mag[i] = c[i];
    }
    uint32_t nbytes = 0;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        if (nbytes == 0 && mag[i]) {
            uint32_t x = mag[i];
            nbytes = 4 * i + (x >> 24 ? 4 : x >> 16 ? 3 : x >> 8 ? 2 : 1);
        }
    }
    return nbytes;
}

// ---- generator preparation --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_g1_to_affine(const g1_jac_t *in, g1_aff_t *out, uint32_t n) {
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) out[i] = g1_to_affine(in[i]);
}

// T[w][j] = 2^(8w) * G_j, affine.  One thread per generator (once per generator set): a chain of 248 doublings in Jacobian
// form, then ONE inversion for all 31 points (Montgomery's trick; zs / pre are [31][n] scratch for the z's and their prefix
// products).  32-thread CTAs spread the n chains over as many SMs as possible: the chain is latency bound.
constexpr int kTableBuildBlock = 32;
__global__ void __launch_bounds__(kTableBuildBlock) k_msm_table_build(const g1_aff_t *gens, g1_aff_t *table, fp_t *zs, fp_t *pre, uint32_t n) {
    const uint32_t j = blockIdx.x * kTableBuildBlock + threadIdx.x;
    if (j >= n) return;
    const g1_aff_t a = gens[j];
    table[j] = a;
    if (a.is_inf()) {
        for (int w = 1; w < kMsmWindows; ++w) table[(size_t) w * n + j] = g1_aff_t::inf();
        return;
    }
    g1_jac_t p = g1_jac_t::from_affine(a);
    fp_t run = fp_t::one();
    for (int w = 1; w < kMsmWindows; ++w) {
        for (int k = 0; k < 8; ++k) p = g1_dbl(p);          // never infinity: G_j has prime order
        g1_aff_t *e = table + (size_t) w * n + j;
        e->x = p.x;
        e->y = p.y;
        run = run * p.z;
        zs[(size_t) (w - 1) * n + j] = p.z;
        pre[(size_t) (w - 1) * n + j] = run;
    }
    fp_t inv = run.inverse();
    for (int w = kMsmWindows - 1; w >= 1; --w) {
        const fp_t zi = w > 1 ? inv * pre[(size_t) (w - 2) * n + j] : inv;   // 1 / z_w
        inv = inv * zs[(size_t) (w - 1) * n + j];
        const fp_t zi2 = zi.sqr();
        g1_aff_t *e = table + (size_t) w * n + j;
        e->x = e->x * zi2;
        e->y = e->y * zi2 * zi;
    }
}

// widest magnitude (in bytes) per row
__global__ void __launch_bounds__(kBlock) k_msm_rowinfo(const fr_t *scalars, uint64_t n, uint32_t n_rows, uint32_t *rowinfo) {
    ZK_PDL_ENTRY();
    const uint64_t total = n * n_rows;
    for (uint64_t i = (uint64_t) blockIdx.x * kBlock + threadIdx.x; i < total; i += (uint64_t) gridDim.x * kBlock) {
        fr_t s = ld_fr(scalars + i);
        if (s.is_zero()) continue;
        uint32_t mag[8], neg;
        uint32_t nb = scalar_sign_mag(s, mag, neg);
        atomicMax(rowinfo + (uint32_t) (i / n), nb);
    }
}

struct msm_smem_t {
    g1_jac_t bucket[kMsmBuckets];
    g1_jac_t first[kBlock];
    g1_jac_t last[kBlock];
    uint32_t count[kMsmBuckets];
    uint32_t off[kMsmBuckets + 1];
    uint32_t cursor[kMsmBuckets];
    uint16_t sorted[kMsmChunk];
    uint8_t dig[kMsmChunk];
    uint8_t sgn[kMsmChunk];
};

struct msm_args_t {
    const fr_t *scalars;     // [n_rows][n]
    const g1_aff_t *table;   // [kMsmWindows][n_table]
    const uint32_t *rowinfo; // widest magnitude per row, bytes
    uint64_t n;              // row length (== number of generators used)
    uint32_t n_table;        // generators in the table (row stride of the table)
    uint32_t n_chunks;       // CTAs per row and window
    uint32_t chunk;          // entries per CTA (<= kMsmChunk)
    uint32_t wide_only;      // 1: scalars that fit one byte are skipped (the small-multiples path took them, msm_kernels.cuh)
    g1_jac_t *partial;       // [n_rows][n_chunks][kMsmWindows] window sums
    unsigned long long *ops; // != nullptr (profiling): += bucket additions (mixed) of this CTA; the reduction's 2 x 255 full additions are added by the host
};

// grid = (n_rows * n_chunks, kMsmWindows)
__global__ void __launch_bounds__(kBlock) k_msm_window(msm_args_t A) {
    ZK_PDL_ENTRY();
    ZK_DYN_SMEM(msm_smem_t, S);
    const uint32_t t = threadIdx.x;
    const uint32_t row = blockIdx.x / A.n_chunks, chunk = blockIdx.x % A.n_chunks, w = blockIdx.y;
    g1_jac_t *dst = A.partial + ((size_t) blockIdx.x * kMsmWindows + w);
    if (w >= A.rowinfo[row]) {   // no scalar of this row reaches this window
        if (t == 0) *dst = g1_jac_t::inf();
        return;
    }
    const uint64_t base = (uint64_t) chunk * A.chunk;
    const uint32_t nc = (uint32_t) (A.n - base < (uint64_t) A.chunk ? A.n - base : (uint64_t) A.chunk);
    const fr_t *sc = A.scalars + (uint64_t) row * A.n + base;
    const g1_aff_t *T = A.table + (size_t) w * A.n_table + base;

    S->count[t] = 0;
    S->bucket[t] = g1_jac_t::inf();
    S->first[t] = g1_jac_t::inf();
    S->last[t] = g1_jac_t::inf();
    __syncthreads();
    // digits of this window + histogram
    for (uint32_t j = t; j < nc; j += kBlock) {
        fr_t s = ld_fr(sc + j);
        uint32_t d = 0, neg = 0;
        if (!s.is_zero()) {
            uint32_t mag[8];
            const uint32_t nb = scalar_sign_mag(s, mag, neg);
            d = (mag[w >> 2] >> ((w & 3) * 8)) & 0xffu;
            if (A.wide_only && nb <= 1) d = 0;
        }
        S->dig[j] = (uint8_t) d;
        S->sgn[j] = (uint8_t) neg;
        if (d) atomicAdd(&S->count[d], 1u);
    }
    __syncthreads();
    if (t == 0) {
        uint32_t o = 0;
        S->off[0] = 0;
        S->off[1] = 0;
        for (int b = 1; b < kMsmBuckets; ++b) {
            S->cursor[b] = o;
            o += S->count[b];
            S->off[b + 1] = o;
        }
    }
    __syncthreads();
    const uint32_t E = S->off[kMsmBuckets];
    if (A.ops && t == 0 && E) atomicAdd(A.ops, (unsigned long long) E);
    if (E == 0) {
        if (t == 0) *dst = g1_jac_t::inf();
        return;
    }
    for (uint32_t j = t; j < nc; j += kBlock) {
        const uint32_t d = S->dig[j];
        if (d) S->sorted[atomicAdd(&S->cursor[d], 1u)] = (uint16_t) (j | ((uint32_t) S->sgn[j] << 15));
    }
    __syncthreads();
    // balanced accumulation: thread t owns sorted[t*q, (t+1)*q)
    const uint32_t q = (E + kBlock - 1) / kBlock;
    const uint32_t sb = t * q, se = sb + q < E ? sb + q : E;
    if (sb < E) {
        g1_jac_t acc = g1_jac_t::inf();
        uint32_t cur_b = S->dig[S->sorted[sb] & 0x7fffu], run_start = sb;
        for (uint32_t p = sb; p <= se; ++p) {
            uint32_t b = 0, j = 0, neg = 0;
            if (p < se) {
                const uint32_t e = S->sorted[p];
                j = e & 0x7fffu;
                neg = e >> 15;
                b = S->dig[j];
            }
            if (p == se || b != cur_b) {   // flush the finished run
                if (S->off[cur_b] >= sb && S->off[cur_b + 1] <= se) S->bucket[cur_b] = acc;   // bucket lies inside this span
                else if (run_start == sb) S->first[t] = acc;
                else S->last[t] = acc;
                if (p == se) break;
                acc = g1_jac_t::inf();
                cur_b = b;
                run_start = p;
            }
            g1_aff_t pt;
            pt.x = ld_fp(&T[j].x);
            pt.y = ld_fp(&T[j].y);
            if (neg) pt.y = -pt.y;   // (0,0) stays (0,0): infinity is its own negative
            acc = g1_add_mixed(acc, pt);
        }
    }
    __syncthreads();
    // buckets shared by several threads: add up their partial runs
    if (t >= 1 && S->count[t]) {
        const uint32_t lo = S->off[t] / q, hi = (S->off[t + 1] - 1) / q;
        if (lo != hi) {
            g1_jac_t s = g1_jac_t::inf();
            for (uint32_t k = lo; k <= hi; ++k) s = g1_add(s, S->off[t] <= k * q ? S->first[k] : S->last[k]);
            S->bucket[t] = s;
        }
    }
    __syncthreads();
    // sum_b b * B_b = sum_{k = 1..255} S_k with the suffix sums S_k = sum_{b >= k} B_b: a parallel suffix scan (8 steps of one
    // addition per thread, ping-pong between `bucket` and `first`, which is free by now) and a tree sum of S_1..S_255 -- 16
    // dependent point additions instead of the 8 doublings + up to 8 additions + 8 tree levels of scaling every bucket by b.
    {
        g1_jac_t *in = S->bucket, *out = S->first;
        for (uint32_t d = 1; d < (uint32_t) kMsmBuckets; d <<= 1) {
            out[t] = t + d < (uint32_t) kMsmBuckets ? g1_add(in[t], in[t + d]) : in[t];
            __syncthreads();
            g1_jac_t *tmp = in; in = out; out = tmp;
        }
        // 8 steps: the result is back in S->bucket; bucket 0 is empty by construction and its suffix sum is not a term
        if (t == 0) S->bucket[0] = g1_jac_t::inf();
    }
    __syncthreads();
    for (uint32_t s = kBlock / 2; s > 0; s >>= 1) {
        if (t < s) S->bucket[t] = g1_add(S->bucket[t], S->bucket[t + s]);
        __syncthreads();
    }
    if (t == 0) *dst = S->bucket[0];
}

// out[row] = normalised sum of the row's (chunk, window) partial sums
__global__ void __launch_bounds__(64) k_msm_finish(const g1_jac_t *partial, uint32_t n_rows, uint32_t per_row, g1_jac_t *out) {
    const uint32_t row = blockIdx.x * 64 + threadIdx.x;
    if (row >= n_rows) return;
    g1_jac_t s = g1_jac_t::inf();
    for (uint32_t k = 0; k < per_row; ++k) {
        g1_jac_t p = partial[(size_t) row * per_row + k];
        if (!p.is_inf()) s = g1_add(s, p);
    }
    out[row] = g1_normalize(s);
}

// ---- bullet (inner-product argument) rounds ---------------------------------------------------------------------------
// After k folds the reference's generators are g_k[i] = sum_{j = i mod m} c_k(j) G_j (m = n / 2^k) with
// c_{k+1}(j) = c_k(j) * (bit (log n - 1 - k) of j clear ? 1/rho_k : 1)   (polyProver.cpp:104).  Instead of folding points
// (h full scalar multiplications per round) the coefficients are folded and each round's two MSMs run over the
// ORIGINAL generators with scalars a_k[j mod m] * c_k(j):  rows[0] takes the j with (j mod m) < h, rows[1] the others.
__global__ void __launch_bounds__(kBlock) k_bullet_scalars(const fr_t *a, const fr_t *coef, uint32_t n, uint32_t m, fr_t *rows) {
    ZK_PDL_ENTRY();
    const uint32_t h = m >> 1;
    for (uint32_t j = blockIdx.x * kBlock + threadIdx.x; j < n; j += gridDim.x * kBlock) {
        const uint32_t i = j & (m - 1);
        fr_t s = ld_fr(a + i) * ld_fr(coef + j);
        const bool left = i < h;
        st_fr(rows + j, left ? s : fr_t::zero());
        st_fr(rows + n + j, left ? fr_t::zero() : s);
    }
}
// coef[j] *= rinv where bit `bit` of j is clear
__global__ void __launch_bounds__(kBlock) k_bullet_coef(fr_t *coef, uint32_t n, uint32_t bit, fr_t rinv) {
    ZK_PDL_ENTRY();
    for (uint32_t j = blockIdx.x * kBlock + threadIdx.x; j < n; j += gridDim.x * kBlock)
        if (!((j >> bit) & 1u)) st_fr(coef + j, ld_fr(coef + j) * rinv);
}
// a'[i] = a[i] * r + a[i + h]   (polyProver.cpp:103)
__global__ void __launch_bounds__(kBlock) k_bullet_fold(const fr_t *a, fr_t *out, uint32_t h, fr_t r) {
    ZK_PDL_ENTRY();
    for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < h; i += gridDim.x * kBlock)
        st_fr(out + i, ld_fr(a + i) * r + ld_fr(a + i + h));
}
// out[0] = sum_{i<h} a[i] L[i],  out[1] = sum_{i<h} a[i+h] L[i]   (polyProver.cpp:88-91); one CTA
__global__ void __launch_bounds__(kBlock) k_dot2(const fr_t *a, const fr_t *L, uint32_t h, fr_t *out) {
    ZK_PDL_ENTRY();
    __shared__ fr_t sh[2 * kBlock];
    fr_t acc[2] = {fr_t::zero(), fr_t::zero()};
    for (uint32_t i = threadIdx.x; i < h; i += kBlock) {
        fr_t l = ld_fr(L + i);
        acc[0] = acc[0] + ld_fr(a + i) * l;
        acc[1] = acc[1] + ld_fr(a + i + h) * l;
    }
    block_sum<2>(acc, sh);
    if (threadIdx.x == 0) { st_fr(out, acc[0]); st_fr(out + 1, acc[1]); }
}
// dot product of two long vectors with the "last CTA finishes" pattern (polyProver::evaluate, polyProver.cpp:36-42)
__global__ void __launch_bounds__(kBlock) k_dot_long(const fr_t *a, const fr_t *b, uint64_t n, fr_t *partials, uint32_t *counter, fr_t *out) {
    __shared__ fr_t sh[kBlock];
    __shared__ uint32_t ticket;
    fr_t acc[1] = {fr_t::zero()};
    for (uint64_t i = (uint64_t) blockIdx.x * kBlock + threadIdx.x; i < n; i += (uint64_t) gridDim.x * kBlock)
        acc[0] = acc[0] + ld_fr(a + i) * ld_fr(b + i);
    block_sum<1>(acc, sh);
    if (threadIdx.x == 0) {
        st_fr(partials + blockIdx.x, acc[0]);
        __threadfence();
        ticket = atomicAdd(counter, 1u);
    }
    __syncthreads();
    if (ticket != gridDim.x - 1) return;
    __threadfence();
    fr_t tot[1] = {fr_t::zero()};
    for (uint32_t i = threadIdx.x; i < gridDim.x; i += kBlock) tot[0] = tot[0] + ld_fr_cg(partials + i);
    __syncthreads();
    block_sum<1>(tot, sh);
    if (threadIdx.x == 0) { st_fr(out, tot[0]); *counter = 0; }
}

// ---- element-wise G1 (parity tests; generator set-up of the stand-alone verifier) ------------------------------------------
__global__ void __launch_bounds__(64) k_g1_vec_op(const g1_jac_t *a, const g1_jac_t *b, const fr_t *k, g1_jac_t *out, uint32_t n, int op) {
    const uint32_t i = blockIdx.x * 64 + threadIdx.x;
    if (i >= n) return;
    g1_jac_t r;
    if (op == 0) r = g1_add(a[i], b[i]);
    else if (op == 1) r = g1_dbl(a[i]);
    else {
        uint32_t c[8];
        ld_fr(k + i).to_canonical(c);
        r = g1_mul_canonical(a[i], c);
    }
    out[i] = g1_normalize(r);
}

// synthetic benchmark inputs: mode 0 = uniform field elements, mode 2 = "witness-like"
// (40 % zero, 8 % one, rest uniform in [-255, 255], negatives stored as r - |x|).  Values are in Montgomery form.
__global__ void __launch_bounds__(kBlock) k_fill_synthetic(fr_t *out, uint64_t n, uint64_t seed, int mode) {
    for (uint64_t i = (uint64_t) blockIdx.x * kBlock + threadIdx.x; i < n; i += (uint64_t) gridDim.x * kBlock) {
        uint64_t st = seed + 0x9E3779B97F4A7C15ULL * (i + 1);
        auto next = [&]() {
            uint64_t z = (st += 0x9E3779B97F4A7C15ULL);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
            return z ^ (z >> 31);
        };
        fr_t x;
        if (mode == 0) {
            uint32_t c[8];
            for (int k = 0; k < 8; k += 2) { uint64_t w = next(); c[k] = (uint32_t) w; c[k + 1] = (uint32_t) (w >> 32); }
            c[7] &= 0x3fffffffu;   // < 2^254 < r
            x = fr_t::from_canonical(c);
        } else {
            const uint32_t u = (uint32_t) (next() % 100u);
            if (u < 40) x = fr_t::zero();
            else if (u < 48) x = fr_t::one();
            else x = fr_t::from_i64((int64_t) (next() % 511u) - 255);
        }
        st_fr(out + i, x);
    }
}

// device self-test: PTX multiplier vs portable multiplier, Fr and Fp.  mismatches += 1 per differing result.
__global__ void __launch_bounds__(kBlock) k_selftest(uint64_t seed, uint32_t n, uint32_t *mismatches) {
    const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
    if (i >= n) return;
    uint64_t st = seed + 0x9E3779B97F4A7C15ULL * (i + 1);
    auto next = [&]() {
        uint64_t z = (st += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    };
    fr_t a, b;
    fp_t c, d;
    for (int k = 0; k < 8; k += 2) { uint64_t x = next(), y = next(); a.v[k] = (uint32_t) x; a.v[k + 1] = (uint32_t) (x >> 32); b.v[k] = (uint32_t) y; b.v[k + 1] = (uint32_t) (y >> 32); }
    for (int k = 0; k < 12; k += 2) { uint64_t x = next(), y = next(); c.v[k] = (uint32_t) x; c.v[k + 1] = (uint32_t) (x >> 32); d.v[k] = (uint32_t) y; d.v[k + 1] = (uint32_t) (y >> 32); }
    a.v[7] &= 0x3fffffffu; b.v[7] &= 0x3fffffffu;   // < r
    c.v[11] &= 0x0fffffffu; d.v[11] &= 0x0fffffffu; // < p
    uint32_t bad = 0;
    {
        fr_t x = a * b, y;
        fr_t::mul_portable(y.v, a.v, b.v);
        bad += x != y;
        // add/sub round trip and distributivity exercise the carry chains of + and -
        bad += ((a + b) - b) != a;
        bad += ((a - b) + b) != a;
        bad += (a * (b + a)) != (x + a * a);
        // unreduced difference as the second operand of a multiplication
        bad += (a * fr_t::sub_lazy(b, a)) != (a * (b - a));
        bad += (b * fr_t::sub_lazy(a, a)) != fr_t::zero();
        // lazy reduction: three unreduced products summed, reduced once == sum of the reduced products
        fr_lazy_t acc;
        acc.clear();
        acc.mac(a, b); acc.mac(fr_t::sub_lazy(a, fr_t::zero()), a); acc.mac(b, fr_t::sub_lazy(b, fr_t::zero()));
        const uint32_t zero8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint32_t top[8] = {acc.w[16], 0, 0, 0, 0, 0, 0, 0};
        bad += montgomery_of_wide<fr_cfg>(acc.w, acc.w + 8, top) != (x + a * a + b * b);
        // the short reduction (<= 16 products of reduced operands) and the single-integer Montgomery reduction under it
        {
            fr_lazy_t l16;
            l16.clear();
            fr_t want = fr_t::zero();
            for (int k = 0; k < 16; ++k) {
                const fr_t u = (k & 1) ? a : b, w = (k & 2) ? x : ((k & 4) ? a : b);
                l16.mac(u, w);
                want = want + u * w;
            }
            bad += fr_lazy_reduce_upto16(l16) != want;
            fr_lazy_t one_prod;
            one_prod.clear();
            one_prod.mac(a, b);
            bad += fr_lazy_reduce_upto16(one_prod) != x;
        }
        // ... and of a sum of field elements seen as a plain integer (chunk 1 of the same formula)
        uint32_t s9[9];
        uint64_t cy = 0;
        for (int k = 0; k < 8; ++k) { cy += (uint64_t) a.v[k] + b.v[k] + x.v[k]; s9[k] = (uint32_t) cy; cy >>= 32; }
        s9[8] = (uint32_t) cy;
        top[0] = s9[8];
        bad += montgomery_of_wide<fr_cfg>(zero8, s9, top) != (a + b + x);
    }
    {
        fp_t x = c * d, y;
        fp_t::mul_portable(y.v, c.v, d.v);
        bad += x != y;
        bad += ((c + d) - d) != c;
        bad += ((c - d) + d) != c;
        bad += (c * (d + c)) != (x + c * c);
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace zk
