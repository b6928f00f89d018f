// Hyrax commitment kernels (hot loop (3) of BASELINE.json north_star).
//
//   K8  k_msm_rowinfo / k_msm_window                  polyProver::commit -> G1::mulVec  (scalars beyond the small-multiples path of
//       (+ msm_kernels.cuh: k_msm_small, row sums)    msm_kernels.cuh)   3rd/hyrax-bls12-381/src/polyProver.cpp:19-34, mcl ec.hpp:1570-1597
//   K9  k_msm_bucket_fill / _merge / _reduce          the MSMs of bulletProve, all rounds in one pass (few rows, full-width scalars)
//   K9  k_bullet_scalars / k_dot2 / k_bullet_fold     polyProver::bulletProve / bulletUpdate   polyProver.cpp:76-109
//       (RZ = R^T Z of initBulletProve, polyProver.cpp:66-68, reuses k_dense_colsum of sc_kernels.cuh)
//
// MSM design.  All MSMs of a proof share one generator set (the sqrt(n) Pedersen generators), so a fixed-base window
// table T[w][j] = 2^(8w) * G_j (affine) is built once per generator set.  A scalar is split into sign and magnitude
// (mcl's isNegative convention: x >= (r+1)/2 is handled as -(r-x) with the negated point), and the magnitude into
// unsigned 8-bit digits; digit d of window w adds +-T[w][j] into bucket d.  Because the table already carries the
// 2^(8w) factor, the digits of ALL windows share one set of 255 buckets: a work item is (row, chunk of generators) with
// up to 32 entries per generator -- counting sort of the entries by digit in shared memory, balanced accumulation with
// mixed adds, one  sum_b b * B_b  per item.  Windows above the row's widest magnitude contribute no entries.
//
// mcl's mulVec is interleaved wNAF (not Pippenger); results agree as GROUP ELEMENTS, which is why every point that
// leaves the device is normalised to affine.
#pragma once
#include "g1.cuh"
#include "sc_kernels.cuh"

namespace zk {

constexpr int kMsmWindows = 32;      // 8-bit digits of a < 2^255 magnitude (|x| <= (r-1)/2 < 2^254)
constexpr int kMsmChunk = 4096;      // upper bound of the "msm_few_rows_chunk" tunable ((generator, window) entries per work item)
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
// products).  The chain is latency bound and runs next to the commitment's kernels: 128-thread CTAs (one warp per scheduler of 32 SMs)
// instead of 32-thread CTAs on 128 SMs -- each of those took a CTA slot away from k_msm_small for the whole 2.7 ms (commit 11.0 -> 10.1 ms).
constexpr int kTableBuildBlock = 128;
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

// Balanced bucket accumulation over a digit-sorted entry list, shared by the bucket kernels.  Thread t of NT adds up the entries
// sorted[t*q, (t+1)*q) run by run (a run = consecutive entries of one bucket).  A run that neither continues from the previous thread
// nor into the next one is a whole bucket: emit(bucket, sum).  A bucket that spans threads is the chain
//     last[lo] + last[lo+1] + ... + last[hi-1] + first[hi]
// (runs that extend to the right go to `last`, the terminal piece to `first`), summed by a segmented inclusive scan over `last` --
// log2(NT) dependent additions however skewed the digits are (one hot bucket used to cost one addition per thread it spanned).
// entry: generator (7 bits) | window << 7 (5 bits) | negative << 12 | digit << 13.  Every thread of the CTA must call this.
template <int NT, class Emit>
__device__ __forceinline__ void msm_accumulate_runs(const uint32_t *sorted, uint32_t E, const g1_aff_t *T, uint32_t n_table, g1_jac_t *first,
                                                     g1_jac_t *last, uint32_t *chain_start, uint32_t *term_bucket, Emit emit) {
    const uint32_t t = threadIdx.x;
    const uint32_t q = (E + NT - 1) / NT;
    const uint32_t sb = t * q, se = sb + q < E ? sb + q : E;
    last[t] = g1_jac_t::inf();
    chain_start[t] = 1u;
    term_bucket[t] = 0u;   // bucket 0 is never used: "no terminal piece"
    if (sb < E) {
        g1_jac_t acc = g1_jac_t::inf();
        uint32_t cur_b = sorted[sb] >> 13, run_start = sb;
        for (uint32_t p = sb; p <= se; ++p) {
            uint32_t b = 0, e = 0;
            if (p < se) {
                e = sorted[p];
                b = e >> 13;
            }
            if (p == se || b != cur_b) {   // flush the finished run
                const bool from_left = run_start == sb && sb > 0 && (sorted[sb - 1] >> 13) == cur_b;
                const bool to_right = p == se && se < E && (sorted[se] >> 13) == cur_b;
                if (to_right) {
                    last[t] = acc;
                    chain_start[t] = from_left ? 0u : 1u;
                } else if (from_left) {
                    first[t] = acc;
                    term_bucket[t] = cur_b;
                } else emit(cur_b, acc);
                if (p == se) break;
                acc = g1_jac_t::inf();
                cur_b = b;
                run_start = p;
            }
            const g1_aff_t *src = T + (size_t) ((e >> 7) & 31u) * n_table + (e & 127u);
            g1_aff_t pt;
            pt.x = ld_fp(&src->x);
            pt.y = ld_fp(&src->y);
            if ((e >> 12) & 1u) pt.y = -pt.y;   // (0,0) stays (0,0): infinity is its own negative
            acc = g1_add_mixed(acc, pt);
        }
    }
    __syncthreads();
    for (uint32_t d = 1; d < (uint32_t) NT; d <<= 1) {   // segmented inclusive scan, in place (read, barrier, write, barrier)
        const bool take = t >= d && !chain_start[t];
        g1_jac_t left = g1_jac_t::inf();
        uint32_t left_start = 0;
        if (take) { left = last[t - d]; left_start = chain_start[t - d]; }
        __syncthreads();
        if (take) { last[t] = g1_add(left, last[t]); chain_start[t] = left_start; }
        __syncthreads();
    }
    if (term_bucket[t]) emit(term_bucket[t], g1_add(last[t - 1], first[t]));
}

constexpr int kMsmEntries = 4096;             // (generator, window) entries one work item sorts in shared memory
constexpr int kMsmGensPerItem = kMsmEntries / kMsmWindows;   // 128 generators x 32 windows

struct msm_smem_t {
    g1_jac_t bucket[kMsmBuckets];
    g1_jac_t first[kBlock];      // terminal piece of a bucket that spans threads; later the ping-pong partner of `bucket` in the suffix scan
    g1_jac_t last[kBlock];       // a thread's run that extends into the next thread's span (msm_accumulate_runs)
    uint32_t count[kMsmBuckets];
    uint32_t off[kMsmBuckets + 1];
    uint32_t cursor[kMsmBuckets];
    uint32_t scan[2][kMsmBuckets];
    uint32_t sorted[kMsmEntries];   // generator (7 bits) | window << 7 (5 bits) | negative << 12 | digit << 13
};

struct msm_args_t {
    const fr_t *scalars;     // [n_rows][n]
    const g1_aff_t *table;   // [kMsmWindows][n_table]
    const uint32_t *rowinfo; // widest magnitude per row, bytes; rowinfo[n_rows] = number of rows listed in wide_rows
    const uint32_t *wide_rows;   // wide_only: the rows that hold a scalar wider than one byte (appended by k_msm_small)
    uint64_t n;              // row length (== number of generators used)
    uint32_t n_rows;
    uint32_t n_table;        // generators in the table (row stride of the table)
    uint32_t n_chunks;       // work items per row
    uint32_t chunk;          // generators per work item (<= kMsmGensPerItem)
    uint32_t wide_only;      // != 0: scalars of at most this many bytes are skipped (the small-multiples path took them, msm_kernels.cuh)
    g1_jac_t *partial;       // [n_rows][n_chunks] bucket-method sums
    unsigned long long *ops; // != nullptr (profiling): += bucket additions (mixed) of this CTA; the reductions are added by the host
};

// Bucket method over ALL windows at once.  The table already carries the 2^(8w) factor of window w, so the digits of every window of
// a scalar go into ONE set of 255 buckets: a work item is (row, chunk of <= 128 generators) with up to 32 entries per generator, and
// one bucket reduction serves all windows (32 of them before).  Persistent CTAs walk the work items; with wide_only the items are
// (listed wide row, chunk), the list being written by k_msm_small, so a witness with a handful of wide scalars costs a handful of
// items.
__global__ void __launch_bounds__(kBlock) k_msm_window(msm_args_t A) {
    ZK_PDL_ENTRY();
    ZK_DYN_SMEM(msm_smem_t, S);
    const uint32_t t = threadIdx.x;
    const uint32_t n_items = (A.wide_only ? A.rowinfo[A.n_rows] : A.n_rows) * A.n_chunks;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t ri = item / A.n_chunks, chunk = item % A.n_chunks;
        const uint32_t row = A.wide_only ? A.wide_rows[ri] : ri;
        g1_jac_t *dst = A.partial + ((size_t) row * A.n_chunks + chunk);
        const uint32_t nw = A.rowinfo[row] < (uint32_t) kMsmWindows ? A.rowinfo[row] : (uint32_t) kMsmWindows;   // live windows of this row
        const uint64_t base = (uint64_t) chunk * A.chunk;
        const uint32_t nc = (uint32_t) (A.n - base < (uint64_t) A.chunk ? A.n - base : (uint64_t) A.chunk);
        const fr_t *sc = A.scalars + (uint64_t) row * A.n + base;
        const g1_aff_t *T = A.table + base;

        S->count[t] = 0;
        S->bucket[t] = g1_jac_t::inf();
        __syncthreads();
        // thread t takes half of the windows of generator t / 2: digits + histogram
        const uint32_t g = t >> 1, w0 = (t & 1u) * (kMsmWindows / 2);
        uint32_t dg[4] = {0, 0, 0, 0}, neg = 0;   // the 16 digits of this thread's windows
        if (g < nc && w0 < nw) {
            const fr_t s = ld_fr(sc + g);
            if (!s.is_zero()) {
                uint32_t mag[8];
                const uint32_t nb = scalar_sign_mag(s, mag, neg);
                if (!(A.wide_only && nb <= A.wide_only)) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) dg[k] = mag[(w0 >> 2) + k];
                }
            }
#pragma unroll
            for (int k = 0; k < kMsmWindows / 2; ++k) {
                const uint32_t d = (dg[k >> 2] >> ((k & 3) * 8)) & 0xffu;
                if (d && w0 + k < nw) atomicAdd(&S->count[d], 1u);
            }
        }
        __syncthreads();
        // exclusive prefix sum of the histogram (bucket 0 stays empty)
        {
            uint32_t (*sa)[kMsmBuckets] = S->scan;
            sa[0][t] = t ? S->count[t] : 0u;
            __syncthreads();
            int cur = 0;
            for (uint32_t d = 1; d < (uint32_t) kMsmBuckets; d <<= 1) {
                sa[cur ^ 1][t] = sa[cur][t] + (t >= d ? sa[cur][t - d] : 0u);
                __syncthreads();
                cur ^= 1;
            }
            const uint32_t incl = sa[cur][t], excl = incl - (t ? S->count[t] : 0u);
            S->off[t] = excl;
            S->cursor[t] = excl;
            if (t == kBlock - 1) S->off[kMsmBuckets] = incl;
        }
        __syncthreads();
        const uint32_t E = S->off[kMsmBuckets];
        if (E == 0) {   // (uniform over the CTA)
            if (t == 0) *dst = g1_jac_t::inf();
            continue;
        }
        if (A.ops && t == 0) atomicAdd(A.ops, (unsigned long long) E);
        if (g < nc && w0 < nw) {
#pragma unroll
            for (int k = 0; k < kMsmWindows / 2; ++k) {
                const uint32_t d = (dg[k >> 2] >> ((k & 3) * 8)) & 0xffu;
                if (d && w0 + k < nw) S->sorted[atomicAdd(&S->cursor[d], 1u)] = g | ((w0 + k) << 7) | (neg << 12) | (d << 13);
            }
        }
        __syncthreads();
        msm_accumulate_runs<kBlock>(S->sorted, E, T, A.n_table, S->first, S->last, S->cursor, S->scan[0], [&](uint32_t b, const g1_jac_t &sum) { S->bucket[b] = sum; });
        __syncthreads();
        // sum_b b * B_b = sum_{k = 1..255} S_k with the suffix sums S_k = sum_{b >= k} B_b: a parallel suffix scan (8 steps of one addition
        // per thread, ping-pong between `bucket` and `first`, which is free by now) and a tree sum of S_1..S_255 -- 16 dependent point
        // additions.  (Per segment of 8 buckets with running sums it is ~1000 additions instead of 4096 but 30 dependent ones, and this
        // kernel holds its SM alone either way: the uniform-scalar 4096 x 4096 MSM took 669 ms that way, 571 ms this way.)
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
        for (uint32_t st = kBlock / 2; st > 0; st >>= 1) {
            if (t < st) S->bucket[t] = g1_add(S->bucket[t], S->bucket[t + st]);
            __syncthreads();
        }
        if (t == 0) *dst = S->bucket[0];
        __syncthreads();   // the next item reuses the shared arrays
    }
}

// ---- few rows (the opening's two MSMs per round): accumulate / merge / reduce as three lean launches -------------------------------
// k_msm_window keeps a whole SM (register file, 130 KB of shared memory) for as long as its slowest phase -- a reduction 32 threads wide --
// takes.  With several proofs in flight that idles the machine, so MSMs of a few rows split the work by the parallelism it has:
//   k_msm_bucket_fill    (row, chunk) items as above, 128 threads, ~57 KB: sort + balanced accumulation only; the 255 bucket sums of
//                        the item go to global memory (three CTAs per SM, every thread adding points for most of the CTA's life)
//   k_msm_bucket_merge   eight threads per (row, bucket): sum over the row's items (the lanes stride the items, then a 3-level tree)
//   k_msm_bucket_reduce  one CTA per row:  sum_b b * B_b = sum_{k<8} 2^k * S_k,  S_k = sum of the buckets whose index has bit k set
//                        (eight 128-leaf trees side by side, then 7 doublings + 7 additions); the result is normalised in place
constexpr int kFillThreads = 128;
struct msm_fill_smem_t {
    g1_jac_t first[kFillThreads];   // partial run at the start / end of a thread's span (buckets shared with the neighbours)
    g1_jac_t last[kFillThreads];
    uint32_t count[kMsmBuckets];
    uint32_t off[kMsmBuckets + 1];
    uint32_t cursor[kMsmBuckets];
    uint32_t scan[2][kFillThreads];
    uint32_t sorted[kMsmEntries];   // generator (7 bits) | window << 7 | negative << 12 | digit << 13
};
struct msm_fill_args_t {
    const fr_t *scalars;     // [n_rows][n]
    const g1_aff_t *table;   // [kMsmWindows][n_table]
    const uint32_t *rowinfo; // widest magnitude per row, bytes
    uint64_t n;
    uint32_t n_rows, n_table, n_chunks, chunk;   // chunk: generators per item (<= kMsmGensPerItem)
    g1_jac_t *buckets;       // [n_rows * n_chunks][kMsmBuckets]; written for items with entries only
    uint32_t *item_entries;  // [n_rows * n_chunks] entries of the item (0: its buckets were not written)
    unsigned long long *ops; // != nullptr (profiling): += bucket additions (mixed)
};

__global__ void __launch_bounds__(kFillThreads, 3) k_msm_bucket_fill(msm_fill_args_t A) {
    ZK_PDL_ENTRY();
    ZK_DYN_SMEM(msm_fill_smem_t, S);
    const uint32_t t = threadIdx.x;
    const uint32_t n_items = A.n_rows * A.n_chunks;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t row = item / A.n_chunks, chunk = item % A.n_chunks;
        const uint32_t nw = A.rowinfo[row] < (uint32_t) kMsmWindows ? A.rowinfo[row] : (uint32_t) kMsmWindows;
        const uint64_t base = (uint64_t) chunk * A.chunk;
        const uint32_t nc = (uint32_t) (A.n - base < (uint64_t) A.chunk ? A.n - base : (uint64_t) A.chunk);
        const fr_t *sc = A.scalars + (uint64_t) row * A.n + base;
        const g1_aff_t *T = A.table + base;
        g1_jac_t *gout = A.buckets + (size_t) item * kMsmBuckets;

        S->count[t] = 0;
        S->count[t + kFillThreads] = 0;
        __syncthreads();
        // thread t takes generator t: the 32 digits of its magnitude + histogram
        uint32_t dg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, neg = 0;
        if (t < nc && nw) {
            const fr_t s = ld_fr(sc + t);
            if (!s.is_zero()) scalar_sign_mag(s, dg, neg);
#pragma unroll
            for (int k = 0; k < kMsmWindows; ++k) {
                const uint32_t d = (dg[k >> 2] >> ((k & 3) * 8)) & 0xffu;
                if (d && (uint32_t) k < nw) atomicAdd(&S->count[d], 1u);
            }
        }
        __syncthreads();
        // exclusive prefix sum of the 256 counts: pairs of buckets per thread, 7-step scan over the pair sums
        {
            const uint32_t c0 = S->count[2 * t], c1 = S->count[2 * t + 1];
            S->scan[0][t] = c0 + c1;
            __syncthreads();
            int cur = 0;
            for (uint32_t d = 1; d < (uint32_t) kFillThreads; d <<= 1) {
                S->scan[cur ^ 1][t] = S->scan[cur][t] + (t >= d ? S->scan[cur][t - d] : 0u);
                __syncthreads();
                cur ^= 1;
            }
            const uint32_t incl = S->scan[cur][t], excl = incl - (c0 + c1);
            S->off[2 * t] = excl;
            S->off[2 * t + 1] = excl + c0;
            S->cursor[2 * t] = excl;
            S->cursor[2 * t + 1] = excl + c0;
            if (t == kFillThreads - 1) S->off[kMsmBuckets] = incl;
        }
        __syncthreads();
        const uint32_t E = S->off[kMsmBuckets];
        if (t == 0) {
            A.item_entries[item] = E;
            if (A.ops && E) atomicAdd(A.ops, (unsigned long long) E);
        }
        if (E == 0) continue;   // (uniform over the CTA)
        if (t < nc) {
#pragma unroll
            for (int k = 0; k < kMsmWindows; ++k) {
                const uint32_t d = (dg[k >> 2] >> ((k & 3) * 8)) & 0xffu;
                if (d && (uint32_t) k < nw) S->sorted[atomicAdd(&S->cursor[d], 1u)] = t | ((uint32_t) k << 7) | (neg << 12) | (d << 13);
            }
        }
        __syncthreads();
        // empty buckets first (count is not touched below), then the sums: every bucket of a non-empty item is written exactly once
        for (uint32_t b = t; b < (uint32_t) kMsmBuckets; b += kFillThreads)
            if (b && S->count[b] == 0) gout[b] = g1_jac_t::inf();
        msm_accumulate_runs<kFillThreads>(S->sorted, E, T, A.n_table, S->first, S->last, S->cursor, S->scan[0], [&](uint32_t b, const g1_jac_t &sum) { gout[b] = sum; });
        __syncthreads();   // the next item reuses the shared arrays
    }
}

// sum of the eight accumulators of a group of 8 consecutive threads (3 tree levels through shared memory); the result is in the group's
// first thread.  Eight lanes per sum instead of a warp per sum: a 5-level tree over 32 lanes spends 5 additions of warp time on 31 useful
// ones, this spends 3 on 28 (four sums per warp), and the sequential part before it keeps every lane busy.
constexpr int kGroup = 8;
__device__ __forceinline__ g1_jac_t group8_sum(const g1_jac_t &acc, g1_jac_t *sh) {
    g1_jac_t *my = sh + threadIdx.x;
    const uint32_t sub = threadIdx.x & (kGroup - 1);
    *my = acc;
    __syncwarp();
    for (uint32_t st = kGroup / 2; st > 0; st >>= 1) {
        if (sub < st) *my = g1_add(*my, my[st]);
        __syncwarp();
    }
    return *my;
}

// merged[row][b] = sum over the row's items of buckets[item][b]; eight threads per (row, b), each striding the items
constexpr int kMergeThreads = 128;
__global__ void __launch_bounds__(kMergeThreads, 3) k_msm_bucket_merge(const g1_jac_t *buckets, const uint32_t *item_entries, uint32_t n_rows, uint32_t n_chunks,
                                                                        g1_jac_t *merged) {
    ZK_PDL_ENTRY();
    __shared__ g1_jac_t sh[kMergeThreads];
    const uint32_t sub = threadIdx.x & (kGroup - 1);
    const uint32_t idx = blockIdx.x * (kMergeThreads / kGroup) + threadIdx.x / kGroup;   // row * 256 + b
    const uint32_t row = idx / kMsmBuckets, b = idx % kMsmBuckets;
    g1_jac_t acc = g1_jac_t::inf();
    if (row < n_rows && b != 0)
        for (uint32_t k = sub; k < n_chunks; k += kGroup) {
            const size_t item = (size_t) row * n_chunks + k;
            if (item_entries[item] == 0) continue;
            const g1_jac_t p = buckets[item * kMsmBuckets + b];
            if (!p.is_inf()) acc = g1_add(acc, p);
        }
    const g1_jac_t tot = group8_sum(acc, sh);
    if (row < n_rows && sub == 0) merged[idx] = tot;
}

// out[row] = normalised sum_b b * merged[row][b]; one CTA per row, thread (k, u) = (threadIdx.x / 32, threadIdx.x % 32) works on S_k.
// With S_out the eight S_k of every row are handed out instead (S_out[row * 8 + k]) and the caller finishes on the host
// (msm_finish_host: 7 doublings + 7 additions + one inversion take ~50 us there, ~400 us on one GPU thread).
__global__ void __launch_bounds__(kBlock) k_msm_bucket_reduce(const g1_jac_t *merged, uint32_t n_rows, g1_jac_t *out, g1_jac_t *S_out) {
    ZK_PDL_ENTRY();
    __shared__ g1_jac_t sh[kBlock];
    const uint32_t row = blockIdx.x, k = threadIdx.x >> 5, u = threadIdx.x & 31u;
    const g1_jac_t *B = merged + (size_t) row * kMsmBuckets;
    // the v-th bucket index with bit k set: bit k inserted into the 7-bit number v
    g1_jac_t acc = g1_jac_t::inf();
    for (uint32_t v = u; v < (uint32_t) kMsmBuckets / 2; v += 32) {
        const uint32_t i = ((v >> k) << (k + 1)) | (1u << k) | (v & ((1u << k) - 1u));
        const g1_jac_t p = B[i];
        if (!p.is_inf()) acc = g1_add(acc, p);
    }
    g1_jac_t *my = sh + threadIdx.x;
    *my = acc;
    __syncthreads();
    for (uint32_t st = 16; st > 0; st >>= 1) {
        if (u < st) *my = g1_add(*my, my[st]);
        __syncthreads();
    }
    if (S_out) {
        if (u == 0) S_out[row * 8 + k] = *my;
        return;
    }
    if (threadIdx.x == 0) {
        g1_jac_t r = sh[7 * 32];
        for (int kk = 6; kk >= 0; --kk) r = g1_add(g1_dbl(r), sh[kk * 32]);
        out[row] = g1_normalize(r);
    }
}
// the host's half of k_msm_bucket_reduce: sum_k 2^k * S[k], normalised
inline g1_jac_t msm_finish_host(const g1_jac_t *S) {
    g1_jac_t r = S[7];
    for (int k = 6; k >= 0; --k) r = g1_add(g1_dbl(r), S[k]);
    return g1_normalize(r);
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
    if (i < 512) {   // inversion: the binary algorithm (host), plain Fermat and windowed Fermat (device) agree, and x * x^-1 == 1
        bad += a.inverse_fermat_w4() != a.inverse_fermat();
        bad += c.inverse_fermat_w4() != c.inverse_fermat();
        bad += a.inverse() != a.inverse_fermat();
        bad += c.inverse() != c.inverse_fermat();
        bad += !a.is_zero() && (a * a.inverse()) != fr_t::one();
        bad += !c.is_zero() && (c * c.inverse()) != fp_t::one();
        bad += !fr_t::zero().inverse().is_zero();
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace zk
