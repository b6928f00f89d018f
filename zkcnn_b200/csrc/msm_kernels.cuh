// Small-scalar fast path of the Hyrax commitment MSM (K8) and fixed-base scalar multiplication.
//
// The zkCNN witness is tiny-valued (hard part 3): of the 2^24 scalars of vgg11 a fifth is zero and 99.8 % of the
// rest satisfy |x| < 64 once mcl's sign convention is applied; a few ten thousand are 2-3 bytes wide, none wider.  All sqrt(n)
// commitment rows share one generator set, so a table of small multiples  M[j][d-1] = d * G_j  (d = 1 .. 2^w - 1, affine, 96 B)  turns
// a row commitment into ONE mixed addition per non-zero w-bit digit of a scalar -- no buckets, no window reduction:
//     comm[i] = sum_k 2^(wk) * sum_j sign(z_ij) * M[j][digit_k(|z_ij|)]          (k_msm_small, k_msm_finish_rows)
// The table has 2^w - 1 entries per generator (w = 6: 25 MB for the 4096 vgg11 generators, L2 resident; w = 8: 100 MB) and is rebuilt
// whenever the generator set changes (k_msm_multiples_build, ~39 field multiplications per entry with one shared inversion per 32
// entries).  Scalars of more than 24 bits are left to the bucket kernel (k_msm_window of hyrax_kernels.cuh in "wide only" mode) and
// both partial results meet in k_msm_finish_rows.
#pragma once
#include "hyrax_kernels.cuh"

namespace zk {

constexpr int kMultiples = 255;          // d = 1 .. 255
constexpr int kMulSeg = 32;              // table entries produced by one thread (one shared inversion)
constexpr int kMulThreadsPerGen = 8;     // 8 * 32 = 256 >= 255

// ---- M[j][d-1] = d * G_j ------------------------------------------------------------------------------------------------
// thread (j, s) produces d = 32 s + 1 .. 32 s + 32: a chain of mixed additions in Jacobian form, then ONE inversion of the
// product of all z (Montgomery's trick) to normalise the 32 points.
// n_mult entries per generator (d = 1 .. n_mult; 255 for byte digits, 63 for 6-bit digits), (n_mult + 31) / 32 threads per generator
__global__ void __launch_bounds__(128) k_msm_multiples_build(const g1_aff_t *gens, g1_aff_t *table, uint32_t n, uint32_t n_mult) {
    const uint32_t idx = blockIdx.x * 128 + threadIdx.x;
    const uint32_t tpg = (n_mult + kMulSeg - 1) / kMulSeg;
    const uint32_t j = idx / tpg, s = idx % tpg;
    if (j >= n) return;
    const g1_aff_t G = gens[j];
    g1_aff_t *out = table + (size_t) j * n_mult + s * kMulSeg;
    const int cnt = (s + 1) * kMulSeg > n_mult ? (int) (n_mult - s * kMulSeg) : kMulSeg;
    if (G.is_inf()) {
        for (int k = 0; k < cnt; ++k) out[k] = g1_aff_t::inf();
        return;
    }
    // B = (32 s) * G
    g1_jac_t B = g1_jac_t::inf();
    if (s) {
        g1_jac_t p32 = g1_jac_t::from_affine(G);
        for (int k = 0; k < 5; ++k) p32 = g1_dbl(p32);
        B = p32;
        for (uint32_t t = 1; t < s; ++t) B = g1_add(B, p32);
    }
    fp_t zs[kMulSeg], pre[kMulSeg];
    fp_t run = fp_t::one();
    for (int k = 0; k < cnt; ++k) {
        B = g1_add_mixed(B, G);          // (32 s + k + 1) * G; never infinity: G has prime order r > 255
        out[k].x = B.x;
        out[k].y = B.y;
        zs[k] = B.z;
        run = run * B.z;
        pre[k] = run;
    }
    fp_t inv = run.inverse();
    for (int k = cnt - 1; k >= 0; --k) {
        const fp_t zi = k ? inv * pre[k - 1] : inv;      // 1 / z_k
        inv = inv * zs[k];
        const fp_t zi2 = zi.sqr();
        out[k].x = out[k].x * zi2;
        out[k].y = out[k].y * zi2 * zi;
    }
}

// ---- one warp per (row, segment): comm partial = sum of table points selected by the small scalars ---------------------------
// Scalars of up to kSmallBytes bytes stay on this path: byte k of the magnitude selects a table point that is added into the warp's
// accumulator of LEVEL k, the warp writes one partial sum per level and k_msm_finish_rows combines the levels as
// L_0 + 256 * (L_1 + 256 * L_2)  -- 8 doublings per level and row instead of a bucket reduction per (row, chunk).  The zkCNN witness
// has ~25 000 scalars of 2-3 bytes among 2^24 (biases scaled to the accumulator's fixed point).  Level 0 lives in registers, the
// rarely used higher levels in shared memory; all digits go through the same queue, so the lanes stay balanced.
constexpr int kSmallBytes = 3;    // scalars of up to 3 bytes (24 bits) stay on the small-multiples path
constexpr int kSmallLevels = 4;   // digit levels a warp can hold: 24 bits = 3 levels of 8, 4 levels of 6 or 7 bits
struct msm_small_args_t {
    const fr_t *scalars;       // [n_rows][n]
    const g1_aff_t *table;     // [n][2^digit_bits - 1]
    uint64_t n;
    uint32_t n_rows, n_seg, seg_len;   // seg_len <= 65536
    uint32_t digit_bits;       // 6, 7 or 8: width of a digit = log2(table entries per generator + 1)
    g1_jac_t *partial;         // level 0: [n_rows][n_seg][32], the 32 lane accumulators of every warp as they are (k_msm_finish_rows adds them
                               // up: a 5-level tree per warp here would be a sixth of this kernel's time)
    g1_jac_t *partial_hi;      // levels 1 ..: [n_rows][n_seg][kSmallLevels - 1] warp sums (levels a warp did not use are written as infinity)
    uint32_t *rowinfo;         // [n_rows] widest magnitude (bytes) among the scalars wider than kSmallBytes (atomicMax); [n_rows] = length of
                               // wide_rows; [n_rows + 1 + row] = byte levels present in the row (<= kSmallBytes)
    uint32_t *wide_rows;       // the rows with a scalar wider than kSmallBytes, in the order they were found (work list of k_msm_window)
    unsigned long long *ops;   // != nullptr (profiling): += mixed additions performed (one per non-zero byte of a small scalar)
};
#ifndef ZK_SMALL_WARPS
#define ZK_SMALL_WARPS 4
#endif
constexpr int kSmallWarps = ZK_SMALL_WARPS;   // 4 warps per CTA: three CTAs (12 warps) fit the register file of an SM, one 8-warp CTA would be alone
struct msm_small_smem_t {
    g1_jac_t lvl[kSmallLevels - 1][kSmallWarps * 32];   // per thread: slot k holds the accumulator of level k + 1 -- or the one of level 0 while
                                                        // level k + 1 is the one in registers; at the end the tree sums of levels 1 ..
    uint32_t queue[kSmallWarps][64];
};

__global__ void __launch_bounds__(kSmallWarps * 32, kSmallWarps == 4 ? 3 : 1) k_msm_small(msm_small_args_t A) {
    ZK_DYN_SMEM(msm_small_smem_t, S);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t gw = (uint64_t) blockIdx.x * kSmallWarps + warp;
    const bool active = gw < (uint64_t) A.n_rows * A.n_seg;   // idle warps of the last CTA keep walking through the barriers
    const uint32_t row = active ? (uint32_t) (gw / A.n_seg) : 0u, seg = active ? (uint32_t) (gw % A.n_seg) : 0u;
    const uint64_t base = (uint64_t) seg * A.seg_len;
    const uint64_t end = !active ? base : base + A.seg_len < A.n ? base + A.seg_len : A.n;
    const fr_t *sc = A.scalars + (uint64_t) row * A.n;
    uint32_t *q = S->queue[warp];
    (void) q;   // unused by the (uncompacted) emulator path
    // ONE accumulator in registers, of level acc_level (0 nearly always); a digit of another level swaps it with that level's slot
    g1_jac_t acc = g1_jac_t::inf();
    uint32_t acc_level = 0;
#pragma unroll
    for (int l = 0; l < kSmallLevels - 1; ++l) S->lvl[l][threadIdx.x] = g1_jac_t::inf();
    auto swap_with = [&](uint32_t level) {   // registers: level 0 <-> `level`
        g1_jac_t *slot = &S->lvl[level - 1][threadIdx.x];
        const g1_jac_t t = *slot;
        *slot = acc;
        acc = t;
    };
    const uint32_t w = A.digit_bits, dmask = (1u << w) - 1u, n_mult = dmask;
    uint32_t wide = 0, used = 0, n_adds = 0, head = 0, pending = 0;
    (void) head; (void) pending;
    // code: generator within the segment (16 bits) | digit << 16 (<= 8 bits) | negative << 24 | level << 25
    auto consume = [&](uint32_t code) {
        ++n_adds;
        const uint32_t j = code & 0xffffu, d = (code >> 16) & 0xffu, neg = (code >> 24) & 1u, level = code >> 25;
        const g1_aff_t *e = A.table + ((base + j) * n_mult + (d - 1));
        g1_aff_t pt;
        pt.x = ld_fp(&e->x);
        pt.y = ld_fp(&e->y);
        if (neg) pt.y = -pt.y;
        if (level != acc_level) {
            if (acc_level) swap_with(acc_level);
            if (level) swap_with(level);
            acc_level = level;
            used |= 1u << level;
        }
        acc = g1_add_mixed(acc, pt);
    };
    for (uint64_t b = base; b < end; b += 32) {
        const uint64_t j = b + lane;
        uint32_t digits = 0, neg = 0;   // up to kSmallBytes digits of this lane's scalar
        if (j < end) {
            const fr_t s = ld_fr(sc + j);
            if (!s.is_zero()) {
                uint32_t mag[8];
                const uint32_t nb = scalar_sign_mag(s, mag, neg);
                if (nb > (uint32_t) kSmallBytes) wide = wide > nb ? wide : nb;
                else digits = mag[0];
            }
        }
#pragma unroll
        for (int k = 0; k < kSmallLevels; ++k) {
            const uint32_t d = (digits >> (w * k)) & dmask;
            const uint32_t code = d ? ((uint32_t) (j - base) | (d << 16) | (neg << 24) | ((uint32_t) k << 25)) : 0u;
            // compact the non-zero digits of these 32 entries into the warp's queue, so that every lane of the warp
            // has a point addition to do whenever the queue is drained (the witness is 40 % zeros)
#if ZK_ON_DEVICE
            const uint32_t mask = __ballot_sync(0xffffffffu, code != 0);
            if (mask == 0) continue;   // (warp-uniform; the usual case for k >= 1)
            const uint32_t rank = __popc(mask & ((1u << lane) - 1u));
            if (code) q[(head + pending + rank) & 63u] = code;
            pending += __popc(mask);
            __syncwarp();
            if (pending >= 32) {
                const uint32_t c = q[(head + lane) & 63u];
                head = (head + 32) & 63u;
                pending -= 32;
                __syncwarp();
                consume(c);
            }
#else
            if (code) consume(code);
#endif
        }
    }
#if ZK_ON_DEVICE
    if (lane < pending) consume(q[(head + lane) & 63u]);
    __syncwarp();
    used = __reduce_or_sync(0xffffffffu, used);
#else
    used = (1u << kSmallLevels) - 2u;   // the emulator's barriers want the same trip count in every thread: sum every level
#endif
    if (acc_level) swap_with(acc_level);
    if (active) A.partial[gw * 32 + lane] = acc;
    // warp-level sums of the 32 accumulators of the higher levels in use
    for (uint32_t l = 1; l < (uint32_t) kSmallLevels; ++l) {
        g1_jac_t *my = S->lvl[l - 1] + threadIdx.x;
        if (!((used >> l) & 1u)) {   // (warp-uniform)
            if (active && lane == 0) A.partial_hi[gw * (kSmallLevels - 1) + (l - 1)] = g1_jac_t::inf();
            continue;
        }
        __syncwarp();
        for (uint32_t st = 16; st > 0; st >>= 1) {
            if (lane < st) *my = g1_add(*my, my[st]);
            __syncwarp();
        }
        if (active && lane == 0) A.partial_hi[gw * (kSmallLevels - 1) + (l - 1)] = *my;
    }
    if (wide && atomicMax(A.rowinfo + row, wide) == 0) A.wide_rows[atomicAdd(A.rowinfo + A.n_rows, 1u)] = row;   // first to mark the row lists it
    if (used > 1u && (!ZK_ON_DEVICE || lane == 0)) atomicMax(A.rowinfo + A.n_rows + 1 + row, 32u - (uint32_t) __clz(used));
    if (A.ops) {
#if ZK_ON_DEVICE
        const uint32_t tot = __reduce_add_sync(0xffffffffu, n_adds);
        if (lane == 0 && tot) atomicAdd(A.ops, (unsigned long long) tot);
#else
        if (n_adds) atomicAdd(A.ops, (unsigned long long) n_adds);
#endif
    }
}

// ---- out[row] = (byte levels of the row's small-path partials + sum of its bucket-path partials), Jacobian; eight threads per row -------
// small: the lane accumulators of the row's n_small segments (level 0), small_hi: their warp sums of the higher levels.
// (bucket partials of a row exist when the bucket kernel visited it: every row, or with wide_only the rows whose rowinfo is non-zero)
constexpr int kFinishRows = 16;   // rows per CTA
__global__ void __launch_bounds__(kFinishRows * kGroup, 3) k_msm_finish_rows(const g1_jac_t *small, const g1_jac_t *small_hi, uint32_t n_small,
                                                                              const g1_jac_t *bucket, uint32_t n_bucket, const uint32_t *rowinfo,
                                                                              uint32_t wide_only, uint32_t n_rows, uint32_t digit_bits, g1_jac_t *out) {
    ZK_PDL_ENTRY();
    __shared__ g1_jac_t sh[kFinishRows * kGroup];
    const uint32_t sub = threadIdx.x & (kGroup - 1);
    const uint32_t row = blockIdx.x * kFinishRows + threadIdx.x / kGroup;
    const bool active = row < n_rows;
    if (!active) n_small = n_bucket = 0;
    else if (wide_only && rowinfo[row] == 0) n_bucket = 0;
    g1_jac_t acc = g1_jac_t::inf();
    for (uint32_t k = sub; k < n_small * 32; k += kGroup) {
        const g1_jac_t p = small[(size_t) row * n_small * 32 + k];
        if (!p.is_inf()) acc = g1_add(acc, p);
    }
    for (uint32_t k = sub; k < n_bucket; k += kGroup) {
        const g1_jac_t p = bucket[(size_t) row * n_bucket + k];
        if (!p.is_inf()) acc = g1_add(acc, p);
    }
    g1_jac_t tot = group8_sum(acc, sh);
    if (active && sub == 0) {
        const uint32_t levels = n_small ? rowinfo[n_rows + 1 + row] : 0u;   // > 1 only in the rows with a scalar of more than one digit
        if (levels > 1) {
            g1_jac_t hi = g1_jac_t::inf();
            for (uint32_t p = levels - 1; p >= 1; --p) {
                for (uint32_t k = 0; k < digit_bits; ++k) hi = g1_dbl(hi);
                for (uint32_t k = 0; k < n_small; ++k) {
                    const g1_jac_t x = small_hi[((size_t) row * n_small + k) * (kSmallLevels - 1) + (p - 1)];
                    if (!x.is_inf()) hi = g1_add(hi, x);
                }
            }
            for (uint32_t k = 0; k < digit_bits; ++k) hi = g1_dbl(hi);
            tot = g1_add(tot, hi);
        }
        out[row] = tot;   // Jacobian; k_g1_normalize_rows follows
    }
}
// in place: every point to z = 1 (infinity -> all zero), one thread per point.  Separate from k_msm_finish_rows, where one lane per warp
// would walk through the inversion with the 31 others idle: here 4096 commitment rows are 128 warps instead of 4096.
__global__ void __launch_bounds__(128) k_g1_normalize_rows(g1_jac_t *pts, uint32_t n) {
    ZK_PDL_ENTRY();
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (i < n) pts[i] = g1_normalize(pts[i]);
}

// ---- fixed-base scalar multiplication: out[i] = k_i * B for ONE base point ------------------------------------------------------
// comb[w][d-1] = d * 2^(8w) * B (32 x 255 affine entries, built by k_msm_table_build + k_msm_multiples_build);
// every product is 32 table look-ups and mixed additions.
__global__ void __launch_bounds__(128) k_fixed_base_mul(const g1_aff_t *comb, const fr_t *k, uint32_t n, g1_jac_t *out) {
    const uint32_t i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    uint32_t c[8];
    ld_fr(k + i).to_canonical(c);
    g1_jac_t acc = g1_jac_t::inf();
    for (int w = 0; w < kMsmWindows; ++w) {
        const uint32_t d = (c[w >> 2] >> ((w & 3) * 8)) & 0xffu;
        if (!d) continue;
        const g1_aff_t *e = comb + ((size_t) w * kMultiples + (d - 1));
        g1_aff_t pt;
        pt.x = ld_fp(&e->x);
        pt.y = ld_fp(&e->y);
        acc = g1_add_mixed(acc, pt);
    }
    out[i] = g1_normalize(acc);
}

}  // namespace zk
