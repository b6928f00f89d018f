// C ABI, part 2: Hyrax polynomial commitment (mirror of 3rd/hyrax-bls12-381/src/polyProver.cpp) and the stateless
// primitives.  Included by capi.cu only.
#pragma once
#include "capi_sumcheck.cuh"
#include "hyrax_kernels.cuh"
#include "msm_kernels.cuh"

namespace zk {

static uint64_t fnv1a64(const void *p, size_t n) {
    const uint8_t *b = static_cast<const uint8_t *>(p);
    uint64_t h = 0xcbf29ce484222325ULL;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ULL; }
    return h;
}

static void msm_configure() {
#if !defined(ZK_EMU)
    static const bool done = [] {   // (thread-safe: several contexts may start at once)
        rt::check(cudaFuncSetAttribute(k_msm_window, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(msm_smem_t)),
                  "cudaFuncSetAttribute(k_msm_window)");
        rt::check(cudaFuncSetAttribute(k_msm_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(msm_small_smem_t)), "cudaFuncSetAttribute(k_msm_small)");
        rt::check(cudaFuncSetAttribute(k_msm_bucket_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(msm_fill_smem_t)),
                  "cudaFuncSetAttribute(k_msm_bucket_fill)");
        return true;
    }();
    (void) done;
#endif
}

// (re)build the fixed-base window table for a generator set given as host Jacobian points
static void msm_prepare_table(zk_ctx *ctx, hyrax_t &H, const uint64_t *gens, uint32_t n_gens) {
    // same public generators as last time?  The 64-bit hash only short-cuts the comparison: the generators are caller-supplied input, so a
    // hit is confirmed byte for byte against the host copy kept from the last build (a collision must not reuse a stale table)
    const size_t bytes = (size_t) n_gens * sizeof(g1_jac_t);
    const uint64_t h = fnv1a64(gens, bytes) ^ n_gens;
    if (H.table_ready && H.gens_hash == h && H.n_gens == n_gens && H.gens_host.size() == bytes && memcmp(H.gens_host.data(), gens, bytes) == 0) return;
    H.gens_host.assign(reinterpret_cast<const uint8_t *>(gens), reinterpret_cast<const uint8_t *>(gens) + bytes);
    H.n_gens = n_gens;
    rt::dbuf &tmp = H.gens_jac;
    tmp.ensure((size_t) n_gens * sizeof(g1_jac_t));
    h2d_staged(ctx, tmp.p, gens, (size_t) n_gens * sizeof(g1_jac_t));
    H.gens_aff.ensure((size_t) n_gens * sizeof(g1_aff_t));
    ZK_KLAUNCH(ctx, k_g1_to_affine, dim3(grid_for(n_gens)), dim3(kBlock), 0, tmp.as<g1_jac_t>(), H.gens_aff.as<g1_aff_t>(), n_gens);
    H.table.ensure((size_t) kMsmWindows * n_gens * sizeof(g1_aff_t));
    H.table_scratch.ensure((size_t) 2 * kMsmWindows * n_gens * sizeof(fp_t));
    // The window table is a chain of 248 doublings per generator: latency bound, a few hundred resident threads.  It runs on a
    // side stream next to the small-multiples build and the commitment's k_msm_small (which do not need it); the main stream
    // waits for it right before the first k_msm_window (msm_wait_table).
#ifndef ZK_EMU
    if (!ctx->aux_stream) ctx->aux_stream = rt::stream_create(true);
    rt::stream_wait_stream(ctx->aux_stream, ctx->stream);
    ZK_KLAUNCH_S(ctx, ctx->aux_stream, ZK_PROF_MSM, (uint64_t) kMsmWindows * n_gens * 96, k_msm_table_build, dim3((n_gens + kTableBuildBlock - 1) / kTableBuildBlock),
                 dim3(kTableBuildBlock), 0, H.gens_aff.as<g1_aff_t>(), H.table.as<g1_aff_t>(), H.table_scratch.as<fp_t>(),
                 H.table_scratch.as<fp_t>() + (size_t) kMsmWindows * n_gens, n_gens);
    H.table_pending = true;
#else
    ZK_KLAUNCH_C(ctx, ZK_PROF_MSM, (uint64_t) kMsmWindows * n_gens * 96, k_msm_table_build, dim3((n_gens + kTableBuildBlock - 1) / kTableBuildBlock),
                 dim3(kTableBuildBlock), 0, H.gens_aff.as<g1_aff_t>(), H.table.as<g1_aff_t>(), H.table_scratch.as<fp_t>(),
                 H.table_scratch.as<fp_t>() + (size_t) kMsmWindows * n_gens, n_gens);
#endif
    H.gens_hash = h;
    H.table_ready = true;
    H.mult_ready = false;
}

static void msm_wait_table(zk_ctx *ctx, hyrax_t &H) {
    if (!H.table_pending) return;
    rt::stream_wait_stream(ctx->stream, ctx->aux_stream);
    H.table_pending = false;
}

// small-multiples table M[j][d-1] = d * G_j for the current generator set (msm_kernels.cuh); built on first use
constexpr uint32_t kMultiplesMaxGens = 1u << 13;   // 8192 generators -> 200 MB
static void msm_prepare_multiples(zk_ctx *ctx, hyrax_t &H) {
    const uint32_t bits = ctx->msm_digit_bits;
    if (H.mult_ready && H.mult_bits == bits) return;
    const uint32_t n_mult = (1u << bits) - 1u;
    H.mult.ensure((size_t) H.n_gens * n_mult * sizeof(g1_aff_t));
    const uint32_t threads = H.n_gens * ((n_mult + kMulSeg - 1) / kMulSeg);
    ZK_KLAUNCH_C(ctx, ZK_PROF_MSM, (uint64_t) H.n_gens * n_mult * 96, k_msm_multiples_build, dim3((threads + 127) / 128), dim3(128), 0,
                 H.gens_aff.as<g1_aff_t>(), H.mult.as<g1_aff_t>(), H.n_gens, n_mult);
    H.mult_ready = true;
    H.mult_bits = bits;
}

// out_dev[k] (normalised) = sum_j scalars[k*n + j] * G_j  for k < n_rows, generators taken from H.table.
// host_S != nullptr: the caller is about to wait for the stream anyway and wants the points on the host.  When the MSM takes the few-row
// path, *host_S is set to a pinned array of 8 partial sums per row (copy queued on the stream) that msm_finish_host turns into the
// normalised points after the wait, and out_dev is NOT written; otherwise *host_S = nullptr and out_dev holds the result as usual.
// wide_scalars: the caller knows the scalars are full-width field elements (the opening's rows): the bucket kernels take all of them, whatever
// the number of rows.
static void msm_run(zk_ctx *ctx, hyrax_t &H, const fr_t *scalars_dev, uint64_t n, uint32_t n_rows, g1_jac_t *out_dev, g1_jac_t **host_S = nullptr,
                    bool wide_scalars = false) {
    if (host_S) *host_S = nullptr;
    ZK_REQUIRE(H.table_ready && n <= H.n_gens, "MSM: generator table missing or too small");
    msm_configure();
    H.msm_rowinfo.ensure((size_t) (2 * n_rows + 1) * 4);   // [n_rows]: length of the wide-row list; [n_rows + 1 + row]: byte levels of the row (small path)
    rt::dzero(H.msm_rowinfo.p, (size_t) (2 * n_rows + 1) * 4, ctx->stream);
    const uint64_t alg_bytes = n * n_rows * 32 + n * 96 + (uint64_t) n_rows * 144;
    // many rows over one generator set: scalars of up to kSmallBytes bytes go through the small-multiples table, the bucket kernel
    // below only sees what is left (the rows k_msm_small lists as holding wider scalars)
    const bool small_path = !wide_scalars && n_rows >= 16 && H.n_gens <= kMultiplesMaxGens;
    const bool few_rows = n_rows <= 8 || wide_scalars;
    uint32_t n_seg = 0;
    if (small_path) {
        msm_prepare_multiples(ctx, H);
        const uint32_t seg_len = (uint32_t) std::min<uint64_t>(n, ctx->msm_small_seg);
        n_seg = (uint32_t) ((n + seg_len - 1) / seg_len);
        H.msm_small.ensure((size_t) n_rows * n_seg * 32 * sizeof(g1_jac_t));
        H.msm_small_hi.ensure((size_t) n_rows * n_seg * (kSmallLevels - 1) * sizeof(g1_jac_t));
        H.msm_wide_rows.ensure((size_t) n_rows * 4);
        msm_small_args_t S;
        S.scalars = scalars_dev;
        S.table = H.mult.as<g1_aff_t>();
        S.n = n;
        S.n_rows = n_rows; S.n_seg = n_seg; S.seg_len = seg_len;
        S.digit_bits = H.mult_bits;
        S.partial = H.msm_small.as<g1_jac_t>();
        S.partial_hi = H.msm_small_hi.as<g1_jac_t>();
        S.rowinfo = H.msm_rowinfo.as<uint32_t>();
        S.wide_rows = H.msm_wide_rows.as<uint32_t>();
        S.ops = ctx->prof_on ? prof_ops_counter(ctx) : nullptr;
        const uint64_t warps = (uint64_t) n_rows * n_seg;
        ZK_KLAUNCH_C(ctx, ZK_PROF_MSM, alg_bytes, k_msm_small, dim3((uint32_t) ((warps + kSmallWarps - 1) / kSmallWarps)), dim3(kSmallWarps * 32), sizeof(msm_small_smem_t), S);
    } else {
        ZK_KLAUNCH_PDL(ctx, ZK_PROF_MSM, 0, k_msm_rowinfo, dim3(grid_for(n * n_rows)), dim3(kBlock), 0, scalars_dev, n, n_rows, H.msm_rowinfo.as<uint32_t>());
    }
    // work items of the bucket kernel: (row, chunk of generators), all windows of a chunk in one item.  Few rows: smaller chunks, more CTAs
    const uint32_t chunk = few_rows ? std::max<uint32_t>(1, std::min<uint32_t>((n_rows <= 8 ? ctx->msm_few_rows_chunk : ctx->msm_batch_chunk) / kMsmWindows, kMsmGensPerItem))
                                    : kMsmGensPerItem;
    const uint32_t n_chunks = (uint32_t) ((n + chunk - 1) / chunk);
    const bool waited = H.table_pending;   // a cross-stream wait sits between the next launch and its predecessor: plain launch then
    msm_wait_table(ctx, H);
    if (!small_path && ctx->msm_split && few_rows && (uint64_t) n_rows * n_chunks <= 4096) {
        // few rows (the opening): accumulate / merge / reduce as three lean launches (hyrax_kernels.cuh)
        const uint32_t n_items = n_rows * n_chunks;
        H.msm_buckets.ensure((size_t) n_items * kMsmBuckets * sizeof(g1_jac_t));
        H.msm_item_entries.ensure((size_t) n_items * 4);
        H.msm_merged.ensure((size_t) n_rows * kMsmBuckets * sizeof(g1_jac_t));
        msm_fill_args_t F;
        F.scalars = scalars_dev;
        F.table = H.table.as<g1_aff_t>();
        F.rowinfo = H.msm_rowinfo.as<uint32_t>();
        F.n = n;
        F.n_rows = n_rows; F.n_table = H.n_gens; F.n_chunks = n_chunks; F.chunk = chunk;
        F.buckets = H.msm_buckets.as<g1_jac_t>();
        F.item_entries = H.msm_item_entries.as<uint32_t>();
        F.ops = ctx->prof_on ? prof_ops_counter(ctx) + 1 : nullptr;
        // three CTAs per SM, all of them: an even split over fewer CTAs (1536 items as 384 x 4 instead of 444 x 3.46) measured slower,
        // the SMs left with two CTAs lose more than the last partial pass costs
        const uint32_t fgrid = std::min<uint32_t>(n_items, 3 * ZK_SM_COUNT);
        if (waited) ZK_KLAUNCH_C(ctx, ZK_PROF_MSM, alg_bytes, k_msm_bucket_fill, dim3(fgrid), dim3(kFillThreads), sizeof(msm_fill_smem_t), F);
        else ZK_KLAUNCH_PDL(ctx, ZK_PROF_MSM, alg_bytes, k_msm_bucket_fill, dim3(fgrid), dim3(kFillThreads), sizeof(msm_fill_smem_t), F);
        ZK_KLAUNCH_PDL(ctx, ZK_PROF_MSM, 0, k_msm_bucket_merge, dim3(n_rows * kMsmBuckets / (kMergeThreads / kGroup)), dim3(kMergeThreads), 0, H.msm_buckets.as<g1_jac_t>(),
                       H.msm_item_entries.as<uint32_t>(), n_rows, n_chunks, H.msm_merged.as<g1_jac_t>());
        if (host_S && ctx->msm_host_finish && n_rows <= 8) {
            H.msm_S.ensure((size_t) n_rows * 8 * sizeof(g1_jac_t));
            if (!ctx->msm_S_h) ctx->msm_S_h = static_cast<g1_jac_t *>(rt::hmalloc_pinned((size_t) 8 * 8 * sizeof(g1_jac_t)));
            ZK_KLAUNCH_PDL(ctx, ZK_PROF_MSM, 0, k_msm_bucket_reduce, dim3(n_rows), dim3(kBlock), 0, H.msm_merged.as<g1_jac_t>(), n_rows, out_dev, H.msm_S.as<g1_jac_t>());
            rt::d2h(ctx->msm_S_h, H.msm_S.p, (size_t) n_rows * 8 * sizeof(g1_jac_t), ctx->stream);
            *host_S = ctx->msm_S_h;
        } else
            ZK_KLAUNCH_PDL(ctx, ZK_PROF_MSM, 0, k_msm_bucket_reduce, dim3(n_rows), dim3(kBlock), 0, H.msm_merged.as<g1_jac_t>(), n_rows, out_dev, (g1_jac_t *) nullptr);
        return;
    }
    H.msm_out.ensure((size_t) n_rows * n_chunks * sizeof(g1_jac_t));
    msm_args_t A;
    A.scalars = scalars_dev;
    A.table = H.table.as<g1_aff_t>();
    A.rowinfo = H.msm_rowinfo.as<uint32_t>();
    A.wide_rows = small_path ? H.msm_wide_rows.as<uint32_t>() : nullptr;
    A.n = n;
    A.n_rows = n_rows;
    A.n_table = H.n_gens;
    A.n_chunks = n_chunks;
    A.chunk = chunk;
    A.wide_only = small_path ? (uint32_t) kSmallBytes : 0u;
    A.partial = H.msm_out.as<g1_jac_t>();
    A.ops = ctx->prof_on ? prof_ops_counter(ctx) + 1 : nullptr;
    // persistent CTAs (one per SM: the kernel's shared memory) walk the items; with wide_only their number is only known on the device
    const uint32_t grid = (uint32_t) std::min<uint64_t>((uint64_t) n_rows * n_chunks, 2 * ZK_SM_COUNT);
    if (waited) ZK_KLAUNCH_C(ctx, ZK_PROF_MSM, small_path ? 0 : alg_bytes, k_msm_window, dim3(grid), dim3(kBlock), sizeof(msm_smem_t), A);
    else ZK_KLAUNCH_PDL(ctx, ZK_PROF_MSM, small_path ? 0 : alg_bytes, k_msm_window, dim3(grid), dim3(kBlock), sizeof(msm_smem_t), A);
    ZK_KLAUNCH_PDL(ctx, ZK_PROF_MSM, 0, k_msm_finish_rows, dim3((n_rows + kFinishRows - 1) / kFinishRows), dim3(kFinishRows * kGroup), 0,
                 small_path ? H.msm_small.as<g1_jac_t>() : (const g1_jac_t *) nullptr, small_path ? H.msm_small_hi.as<g1_jac_t>() : (const g1_jac_t *) nullptr, n_seg,
                 H.msm_out.as<g1_jac_t>(), n_chunks,
                 H.msm_rowinfo.as<uint32_t>(), small_path ? 1u : 0u, n_rows, H.mult_bits, out_dev);
    ZK_KLAUNCH_PDL(ctx, ZK_PROF_MSM, 0, k_g1_normalize_rows, dim3((n_rows + 127) / 128), dim3(128), 0, out_dev, n_rows);
}

static void hyrax_bind(zk_ctx *ctx, const fr_t *Z, uint32_t bit_length, const uint64_t *gens, uint32_t n_gens) {
    hyrax_t &H = ctx->hy;
    H.Z = Z;
    H.bit_length = bit_length;
    H.r_bits = bit_length >> 1;
    H.l_bits = bit_length - H.r_bits;
    ZK_REQUIRE(gens && n_gens == (1u << H.l_bits), "Hyrax: need exactly 2^ceil(bl/2) generators (polyProver.cpp:26)");
    msm_prepare_table(ctx, H, gens, n_gens);
    H.bound = true;
    H.cur = 0;
    H.round = 0;
}

}  // namespace zk

extern "C" {

int zk_poly_bind_input(zk_ctx *ctx, const uint64_t *gens, uint32_t n_gens) {   // src/prover.cpp:503-511
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && ctx->circuit_ready, "circuit not uploaded");
    zk::rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    zk::layer_t &L0 = ctx->layers[0];
    ZK_REQUIRE(L0.n_val >= L0.d.size && L0.val.p, "input layer witness missing");
    zk::hyrax_bind(ctx, L0.val.as<zk::fr_t>(), (uint32_t) L0.d.bit_length, gens, n_gens);
    ZK_API_END
}

int zk_poly_create(zk_ctx *ctx, const uint64_t *Z, uint64_t n, const uint64_t *gens, uint32_t n_gens) {   // polyProver.cpp:12-17
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && Z && n >= 1, "bad arguments");
    zk::rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    uint32_t bl = 0;
    while ((1ull << bl) < n) ++bl;
    ZK_REQUIRE(bl <= 30, "polynomial too large");
    zk::hyrax_t &H = ctx->hy;
    H.z_own.ensure(sizeof(zk::fr_t) << bl);
    zk::rt::h2d(H.z_own.p, Z, n * sizeof(zk::fr_t), ctx->stream);
    if ((1ull << bl) > n) zk::rt::dzero(H.z_own.as<zk::fr_t>() + n, ((1ull << bl) - n) * sizeof(zk::fr_t), ctx->stream);
    zk::rt::sync(ctx->stream);
    zk::hyrax_bind(ctx, H.z_own.as<zk::fr_t>(), bl, gens, n_gens);
    ZK_API_END
}

int zk_poly_commit(zk_ctx *ctx, uint64_t *comm_out, uint32_t n_out) {   // polyProver.cpp:19-34
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && ctx->hy.bound && comm_out, "no polynomial bound");
    zk::rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    zk::hyrax_t &H = ctx->hy;
    const uint32_t rsize = 1u << H.r_bits, lsize = 1u << H.l_bits;
    ZK_REQUIRE(n_out == rsize, "commit: output must hold 2^(bl/2) points");
    zk::rt::dbuf &out = H.pts_out;
    out.ensure((size_t) rsize * sizeof(zk::g1_jac_t));
    zk::msm_run(ctx, H, H.Z, lsize, rsize, out.as<zk::g1_jac_t>());
    zk::d2h_staged(ctx, comm_out, out.p, (size_t) rsize * sizeof(zk::g1_jac_t));
    ZK_API_END
}

int zk_poly_evaluate(zk_ctx *ctx, const uint64_t *x, uint32_t n, uint64_t *out) {   // polyProver.cpp:36-42
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && ctx->hy.bound && out && n == ctx->hy.bit_length, "evaluate: wrong number of variables");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    hyrax_t &H = ctx->hy;
    ensure_round_scratch(ctx);
    std::vector<fr_t> xs(n);
    memcpy(xs.data(), x, (size_t) n * 32);
    rt::dbuf X;
    X.ensure(sizeof(fr_t) << n);
    beta_point_t pts[1] = {{xs.data(), fr_t::one()}};
    build_beta(ctx, X.as<fr_t>(), n, pts, 1);
    const uint32_t g = grid_for(1ull << n);
    ZK_KLAUNCH(ctx, k_dot_long, dim3(g), dim3(kBlock), 0, H.Z, X.as<fr_t>(), 1ull << n, ctx->partials.as<fr_t>(), ctx->counters.as<uint32_t>() + 3,
               ctx->round_out.as<fr_t>());
    rt::d2h(ctx->h_out, ctx->round_out.p, sizeof(fr_t), ctx->stream);
    rt::sync(ctx->stream);
    fr_store(out, ctx->h_out[0]);
    ZK_API_END
}

int zk_poly_init_bullet_prove(zk_ctx *ctx, const uint64_t *lx, uint32_t n_lx, const uint64_t *rx, uint32_t n_rx) {   // polyProver.cpp:52-74
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && ctx->hy.bound, "no polynomial bound");
    hyrax_t &H = ctx->hy;
    ZK_REQUIRE(n_lx == H.l_bits && n_rx == H.r_bits && (lx || !n_lx) && (rx || !n_rx), "initBulletProve: split of the point does not match");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const uint32_t lsize = 1u << H.l_bits, rsize = 1u << H.r_bits;
    H.t.resize(n_lx);
    memcpy(H.t.data(), lx, (size_t) n_lx * 32);
    std::vector<fr_t> rxs(n_rx);
    memcpy(rxs.data(), rx, (size_t) n_rx * 32);
    H.L.ensure((size_t) lsize * sizeof(fr_t));
    H.R.ensure((size_t) rsize * sizeof(fr_t));
    beta_point_t pl[1] = {{H.t.data(), fr_t::one()}};
    build_beta(ctx, H.L.as<fr_t>(), n_lx, pl, 1);      // L = expand(lx)
    beta_point_t pr[1] = {{rxs.data(), fr_t::one()}};
    build_beta(ctx, H.R.as<fr_t>(), n_rx, pr, 1);      // R = expand(rx)
    // RZ[i] = sum_j R[j] * Z[j * lsize + i]
    H.a.ensure((size_t) lsize * sizeof(fr_t));
    H.a_next.ensure((size_t) lsize * sizeof(fr_t));
    dense_colsum(ctx, H.Z, H.R.as<fr_t>(), H.l_bits, rsize, H.a.as<fr_t>());
    // bullet_g = gens  ->  coefficient vector of ones;  bullet_a = RZ;  scale = 1
    std::vector<fr_t> ones(lsize, fr_t::one());
    H.coef.ensure((size_t) lsize * sizeof(fr_t));
    rt::h2d(H.coef.p, ones.data(), (size_t) lsize * sizeof(fr_t), ctx->stream);
    rt::sync(ctx->stream);
    H.cur = lsize;
    H.round = 0;
    H.scale = fr_t::one();
    ZK_API_END
}

int zk_poly_bullet_prove(zk_ctx *ctx, uint64_t *lcomm, uint64_t *rcomm, uint64_t *ly, uint64_t *ry) {   // polyProver.cpp:76-96
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && ctx->hy.bound && ctx->hy.cur >= 2 && !ctx->hy.t.empty(), "bulletProve: nothing left to prove");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    hyrax_t &H = ctx->hy;
    ensure_round_scratch(ctx);
    const uint32_t lsize = 1u << H.l_bits, m = H.cur, h = m >> 1;
    // two MSMs over the original generators (see k_bullet_scalars)
    H.scal.ensure((size_t) 2 * lsize * sizeof(fr_t));
    ZK_KLAUNCH_PDL(ctx, ZK_PROF_OTHER, 0, k_bullet_scalars, dim3(grid_for(lsize)), dim3(kBlock), 0, H.a.as<fr_t>(), H.coef.as<fr_t>(), lsize, m, H.scal.as<fr_t>());
    rt::dbuf &pts = H.pts_out;
    pts.ensure(2 * sizeof(g1_jac_t));
    g1_jac_t *S_h = nullptr;
    msm_run(ctx, H, H.scal.as<fr_t>(), lsize, 2, pts.as<g1_jac_t>(), &S_h, true);
    // ly, ry
    ZK_KLAUNCH_PDL(ctx, ZK_PROF_OTHER, 0, k_dot2, dim3(1), dim3(kBlock), 0, H.a.as<fr_t>(), H.L.as<fr_t>(), h, ctx->round_out.as<fr_t>());
    rt::d2h(ctx->h_out, ctx->round_out.p, 2 * sizeof(fr_t), ctx->stream);
    g1_jac_t hp[2];
    if (!S_h) rt::d2h(hp, pts.p, sizeof hp, ctx->stream);
    rt::sync(ctx->stream);
    if (S_h) { hp[0] = msm_finish_host(S_h); hp[1] = msm_finish_host(S_h + 8); }   // the last 14 point operations + the inversion of each point
    H.scale = H.scale * (fr_t::one() - H.t.back()).inverse();
    fr_store(ly, ctx->h_out[0] * H.scale);
    fr_store(ry, ctx->h_out[1] * H.scale);
    memcpy(lcomm, &hp[0], sizeof(g1_jac_t));
    memcpy(rcomm, &hp[1], sizeof(g1_jac_t));
    ZK_API_END
}

int zk_poly_bullet_update(zk_ctx *ctx, const uint64_t *randomness) {   // polyProver.cpp:98-109
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && ctx->hy.bound && ctx->hy.cur >= 2 && !ctx->hy.t.empty(), "bulletUpdate: nothing left to fold");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    hyrax_t &H = ctx->hy;
    const fr_t r = fr_load(randomness);
    const fr_t rinv = r.inverse();
    const uint32_t lsize = 1u << H.l_bits, h = H.cur >> 1;
    ZK_KLAUNCH_PDL(ctx, ZK_PROF_OTHER, 0, k_bullet_fold, dim3(grid_for(h)), dim3(kBlock), 0, H.a.as<fr_t>(), H.a_next.as<fr_t>(), h, r);
    std::swap(H.a, H.a_next);
    // generators fold as g[i] * r^-1 + g[i + h]: the coefficient of G_j picks up r^-1 when j lies in a lower half
    uint32_t bit = 0;
    while ((1u << bit) < h) ++bit;   // h == 2^bit
    ZK_KLAUNCH_PDL(ctx, ZK_PROF_OTHER, 0, k_bullet_coef, dim3(grid_for(lsize)), dim3(kBlock), 0, H.coef.as<fr_t>(), lsize, bit, rinv);
    H.cur = h;
    ++H.round;
    H.t.pop_back();
    ZK_API_END
}

// Every round of the inner-product argument in one device pass (bulletProve + bulletUpdate of polyProver.cpp:76-109, n_rounds times).
// The reference's verifier draws a round's randomness from its RNG after the round's message (polyVerifier.cpp:48-50) but independently of
// it, so a caller that draws the randomness of all rounds first gets the same messages: the folds of `a` and of the generator coefficients
// run back to back, the 2 * n_rounds MSMs over the original generators become ONE bucket MSM of 2 * n_rounds rows, and the host waits once.
int zk_poly_bullet_prove_all(zk_ctx *ctx, const uint64_t *randomness, uint32_t n_rounds, uint64_t *lcomm, uint64_t *rcomm, uint64_t *ly, uint64_t *ry) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && ctx->hy.bound && randomness && lcomm && rcomm && ly && ry && n_rounds >= 1 && n_rounds <= 30, "bulletProveAll: bad arguments");
    hyrax_t &H = ctx->hy;
    ZK_REQUIRE(H.cur == (1u << n_rounds) && H.t.size() == n_rounds, "bulletProveAll: n_rounds must be the number of rounds left");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    ensure_round_scratch(ctx);
    const uint32_t lsize = 1u << H.l_bits;
    H.scal.ensure((size_t) 2 * n_rounds * lsize * sizeof(fr_t));
    H.bullet_dots.ensure((size_t) 2 * n_rounds * sizeof(fr_t));
    uint32_t m = H.cur;
    for (uint32_t k = 0; k < n_rounds; ++k) {
        const uint32_t h = m >> 1;
        const fr_t r = fr_load(randomness + 4 * (size_t) k);
        const fr_t rinv = r.inverse();
        ZK_KLAUNCH_PDL(ctx, ZK_PROF_OTHER, 0, k_bullet_scalars, dim3(grid_for(lsize)), dim3(kBlock), 0, H.a.as<fr_t>(), H.coef.as<fr_t>(), lsize, m,
                       H.scal.as<fr_t>() + (size_t) 2 * k * lsize);
        ZK_KLAUNCH_PDL(ctx, ZK_PROF_OTHER, 0, k_dot2, dim3(1), dim3(kBlock), 0, H.a.as<fr_t>(), H.L.as<fr_t>(), h, H.bullet_dots.as<fr_t>() + 2 * k);
        ZK_KLAUNCH_PDL(ctx, ZK_PROF_OTHER, 0, k_bullet_fold, dim3(grid_for(h)), dim3(kBlock), 0, H.a.as<fr_t>(), H.a_next.as<fr_t>(), h, r);
        std::swap(H.a, H.a_next);
        uint32_t bit = 0;
        while ((1u << bit) < h) ++bit;
        ZK_KLAUNCH_PDL(ctx, ZK_PROF_OTHER, 0, k_bullet_coef, dim3(grid_for(lsize)), dim3(kBlock), 0, H.coef.as<fr_t>(), lsize, bit, rinv);
        m = h;
    }
    rt::dbuf &pts = H.pts_out;
    pts.ensure((size_t) 2 * n_rounds * sizeof(g1_jac_t));
    msm_run(ctx, H, H.scal.as<fr_t>(), lsize, 2 * n_rounds, pts.as<g1_jac_t>(), nullptr, true);
    std::vector<fr_t> dots(2 * n_rounds);
    std::vector<g1_jac_t> hp(2 * n_rounds);
    d2h_staged(ctx, dots.data(), H.bullet_dots.p, dots.size() * sizeof(fr_t));
    d2h_staged(ctx, hp.data(), pts.p, hp.size() * sizeof(g1_jac_t));
    for (uint32_t k = 0; k < n_rounds; ++k) {
        H.scale = H.scale * (fr_t::one() - H.t.back()).inverse();
        H.t.pop_back();
        fr_store(ly + 4 * (size_t) k, dots[2 * k] * H.scale);
        fr_store(ry + 4 * (size_t) k, dots[2 * k + 1] * H.scale);
        memcpy(lcomm + 18 * (size_t) k, &hp[2 * k], sizeof(g1_jac_t));
        memcpy(rcomm + 18 * (size_t) k, &hp[2 * k + 1], sizeof(g1_jac_t));
    }
    H.cur = 1;
    H.round += n_rounds;
    ZK_API_END
}

int zk_poly_bullet_open(zk_ctx *ctx, uint64_t *out) {   // polyProver.cpp:111-116
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && ctx->hy.bound && ctx->hy.cur == 1 && out, "bulletOpen: folding not finished");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    ensure_round_scratch(ctx);
    rt::d2h(ctx->h_out, ctx->hy.a.p, sizeof(fr_t), ctx->stream);
    rt::sync(ctx->stream);
    fr_store(out, ctx->h_out[0]);
    ZK_API_END
}

// ---- stateless primitives ------------------------------------------------------------------------------------------------
int zk_fr_vec_op(zk_ctx *ctx, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, uint64_t n) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && a && b && out && op >= 0 && op <= 2 && n < (1ull << 31), "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    rt::dbuf da, db, dc;
    da.ensure(n * 32); db.ensure(n * 32); dc.ensure(n * 32);
    rt::h2d(da.p, a, n * 32, ctx->stream);
    rt::h2d(db.p, b, n * 32, ctx->stream);
    ZK_KLAUNCH(ctx, k_fr_binop, dim3(grid_for(n)), dim3(kBlock), 0, da.as<fr_t>(), db.as<fr_t>(), dc.as<fr_t>(), (uint32_t) n, op);
    rt::d2h(out, dc.p, n * 32, ctx->stream);
    rt::sync(ctx->stream);
    ZK_API_END
}

int zk_beta_table(zk_ctx *ctx, const uint64_t *r, uint32_t bits, const uint64_t *init, uint64_t *out) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && (r || !bits) && init && out && bits <= 26, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    std::vector<fr_t> rs(bits);
    memcpy(rs.data(), r, (size_t) bits * 32);
    rt::dbuf d;
    d.ensure(sizeof(fr_t) << bits);
    beta_point_t pts[1] = {{rs.data(), fr_load(init)}};
    build_beta(ctx, d.as<fr_t>(), bits, pts, 1);
    rt::d2h(out, d.p, sizeof(fr_t) << bits, ctx->stream);
    rt::sync(ctx->stream);
    ZK_API_END
}

int zk_phi_table(zk_ctx *ctx, const uint64_t *rx, const uint64_t *scale, uint32_t n, int is_ifft, uint64_t *out) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && rx && scale && out && n >= 1 && n <= 20, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const size_t entries = is_ifft ? (size_t) 1 << n : (size_t) 1 << (n - 1);
    rt::dbuf d;
    d.ensure(std::max<size_t>(2, entries) * sizeof(fr_t));
    ctx->d_r.ensure(2 * 64 * sizeof(fr_t));
    rt::h2d(ctx->d_r.p, rx, (size_t) (is_ifft ? n - 1 : n) * 32, ctx->stream);
    const fr_t *pw = phi_powers(ctx, n, is_ifft != 0);
    ZK_KLAUNCH(ctx, k_phi_table, dim3(1), dim3(kBlock), 0, d.as<fr_t>(), ctx->d_r.as<fr_t>(), pw, fr_load(scale), (int) n, is_ifft);
    rt::d2h(out, d.p, entries * sizeof(fr_t), ctx->stream);
    rt::sync(ctx->stream);
    ZK_API_END
}

int zk_fold_rounds(zk_ctx *ctx, const uint64_t *V, const uint64_t *M, uint32_t bits, uint64_t live, const uint64_t *r, uint32_t n_rounds,
                   uint64_t *polys) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && V && M && polys && bits >= 1 && bits <= 28 && n_rounds >= 1 && n_rounds <= bits && live <= (1ull << bits), "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    pair_t &P = ctx->pair[1];
    pair_reset(ctx->pair[0], -1, 0);
    pair_reset(P, (int8_t) bits, (uint32_t) live);
    fr_t *dv = table_init_buf(P.v, 1ull << bits), *dm = table_init_buf(P.m, 1ull << bits);
    rt::h2d(dv, V, live * 32, ctx->stream);
    rt::h2d(dm, M, live * 32, ctx->stream);
    ctx->add_term = fr_t::zero();
    ctx->round = 0;
    for (uint32_t j = 0; j < n_rounds; ++j) {
        const fr_t prev = j == 0 ? fr_t::zero() : fr_load(r + 4 * (j - 1));
        ++ctx->round;
        ctx->add_term = ctx->add_term * (fr_t::one() - prev);
        fr_t abc[3];
        round_quadratic(ctx, prev, 2u, abc);
        abc[1] = abc[1] - ctx->add_term;
        abc[2] = abc[2] + ctx->add_term;
        for (int k = 0; k < 3; ++k) fr_store(polys + (size_t) (3 * j + k) * 4, abc[k]);
    }
    P.n_eval = 0;
    ZK_API_END
}

int zk_msm(zk_ctx *ctx, const uint64_t *bases, const uint64_t *scalars, uint64_t n, uint32_t n_rows, uint64_t *out) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && bases && scalars && out && n >= 1 && n <= (1u << 20) && n_rows >= 1, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    hyrax_t H;   // private table: does not disturb a bound polynomial
    msm_prepare_table(ctx, H, bases, (uint32_t) n);
    rt::dbuf ds, dout;
    ds.ensure(n * n_rows * 32);
    dout.ensure((size_t) n_rows * sizeof(g1_jac_t));
    rt::h2d(ds.p, scalars, n * n_rows * 32, ctx->stream);
    msm_run(ctx, H, ds.as<fr_t>(), n, n_rows, dout.as<g1_jac_t>());
    rt::d2h(out, dout.p, (size_t) n_rows * sizeof(g1_jac_t), ctx->stream);
    rt::sync(ctx->stream);
    ZK_API_END
}

int zk_g1_vec_op(zk_ctx *ctx, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, uint64_t n) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && a && out && op >= 0 && op <= 2 && (b || op == 1) && n < (1ull << 24), "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    rt::dbuf da, db, dc;
    da.ensure(n * sizeof(g1_jac_t));
    dc.ensure(n * sizeof(g1_jac_t));
    rt::h2d(da.p, a, n * sizeof(g1_jac_t), ctx->stream);
    const size_t bsz = op == 0 ? sizeof(g1_jac_t) : sizeof(fr_t);
    if (op != 1) { db.ensure(n * bsz); rt::h2d(db.p, b, n * bsz, ctx->stream); }
    ZK_KLAUNCH(ctx, k_g1_vec_op, dim3((uint32_t) ((n + 63) / 64)), dim3(64), 0, da.as<g1_jac_t>(), op == 0 ? db.as<g1_jac_t>() : nullptr,
               op == 2 ? db.as<fr_t>() : nullptr, dc.as<g1_jac_t>(), (uint32_t) n, op);
    rt::d2h(out, dc.p, n * sizeof(g1_jac_t), ctx->stream);
    rt::sync(ctx->stream);
    ZK_API_END
}

int zk_g1_fixed_base_mul(zk_ctx *ctx, const uint64_t *base, const uint64_t *scalars, uint64_t n, uint64_t *out) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && base && scalars && out && n >= 1 && n < (1ull << 24), "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    if (!ctx->fb_ready || memcmp(ctx->fb_base, base, sizeof(g1_jac_t)) != 0) {   // comb[w][d-1] = d * 2^(8w) * base  (cached per base point, compared byte for byte)
        memcpy(ctx->fb_base, base, sizeof(g1_jac_t));
        rt::dbuf jb, ab, win, scr;
        jb.ensure(sizeof(g1_jac_t)); ab.ensure(sizeof(g1_aff_t)); win.ensure((size_t) kMsmWindows * sizeof(g1_aff_t));
        scr.ensure((size_t) 2 * kMsmWindows * sizeof(fp_t));
        rt::h2d(jb.p, base, sizeof(g1_jac_t), ctx->stream);
        ZK_KLAUNCH(ctx, k_g1_to_affine, dim3(1), dim3(kBlock), 0, jb.as<g1_jac_t>(), ab.as<g1_aff_t>(), 1u);
        ZK_KLAUNCH(ctx, k_msm_table_build, dim3(1), dim3(kTableBuildBlock), 0, ab.as<g1_aff_t>(), win.as<g1_aff_t>(), scr.as<fp_t>(),
                   scr.as<fp_t>() + kMsmWindows, 1u);
        ctx->fb_comb.ensure((size_t) kMsmWindows * kMultiples * sizeof(g1_aff_t));
        ZK_KLAUNCH(ctx, k_msm_multiples_build, dim3((kMsmWindows * kMulThreadsPerGen + 127) / 128), dim3(128), 0, win.as<g1_aff_t>(),
                   ctx->fb_comb.as<g1_aff_t>(), (uint32_t) kMsmWindows, (uint32_t) kMultiples);
        rt::sync(ctx->stream);
        ctx->fb_ready = true;
    }
    rt::dbuf &dk = ctx->fb_k, &dout = ctx->fb_out;
    dk.ensure(n * 32);
    dout.ensure(n * sizeof(g1_jac_t));
    h2d_staged(ctx, dk.p, scalars, n * 32);
    ZK_KLAUNCH(ctx, k_fixed_base_mul, dim3((uint32_t) ((n + 127) / 128)), dim3(128), 0, ctx->fb_comb.as<g1_aff_t>(), dk.as<fr_t>(), (uint32_t) n,
               dout.as<g1_jac_t>());
    d2h_staged(ctx, out, dout.p, n * sizeof(g1_jac_t));
    ZK_API_END
}

int zk_selftest(zk_ctx *ctx, uint64_t seed, uint32_t n) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && n >= 1, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    rt::dbuf d;
    d.ensure(4);
    rt::dzero(d.p, 4, ctx->stream);
    ZK_KLAUNCH(ctx, k_selftest, dim3(grid_for(n)), dim3(kBlock), 0, seed, n, d.as<uint32_t>());
    uint32_t bad = 0;
    rt::d2h(&bad, d.p, 4, ctx->stream);
    rt::sync(ctx->stream);
    ZK_REQUIRE(bad == 0, "device self-test: inline-PTX field arithmetic disagrees with the portable implementation");
    ZK_API_END
}

// ---- device-timed micro-benchmarks ---------------------------------------------------------------------------------------------
int zk_bench_fold(zk_ctx *ctx, uint32_t bits, uint32_t iters, int fold, float *ms) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && ms && bits >= 2 && bits <= 28 && iters >= 1, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    ensure_round_scratch(ctx);
    const uint64_t n = 1ull << bits;
    rt::dbuf v, m, vo, mo;
    v.ensure(n * 32); m.ensure(n * 32); vo.ensure(n * 16); mo.ensure(n * 16);
    ZK_KLAUNCH(ctx, k_fill_synthetic, dim3(grid_for(n)), dim3(kBlock), 0, v.as<fr_t>(), n, 0x9E3779B97F4A7C15ULL, 0);
    ZK_KLAUNCH(ctx, k_fill_synthetic, dim3(grid_for(n)), dim3(kBlock), 0, m.as<fr_t>(), n, 0x243F6A8885A308D3ULL, 0);
    round_args_t A;
    memset(&A, 0, sizeof A);
    A.r = fr_t::from_u64(0x1234567887654321ULL);
    A.acc = ctx->round_acc.as<unsigned long long>();
    A.counter = ctx->counters.as<uint32_t>();
    A.out = ctx->round_out.as<fr_t>();
    round_pair_t &R = A.pair[1];
    R.v_in = v.as<fr_t>(); R.m_in = m.as<fr_t>(); R.v_out = vo.as<fr_t>(); R.m_out = mo.as<fr_t>();
    R.n_in = (uint32_t) n; R.live = (uint32_t) n; R.fold = fold ? 1 : 0;
    const uint64_t out_pairs = fold ? n >> 2 : n >> 1;
    const bool thin = out_pairs <= ctx->thin_max_pairs;   // same choices as round_quadratic()
    A.state = ctx->round_state.as<fr_t>();
    A.derive_b[1] = fold && ctx->derive_b_enabled ? 1u : 0u;   // as in every fold round after the first of a phase
    R.n_blocks = thin ? (uint32_t) ((out_pairs + kRoundBlock / 4 - 1) / (kRoundBlock / 4)) : round_grid_for(out_pairs);
    const uint32_t limit_pairs[2] = {0, (uint32_t) out_pairs};
    (void) limit_pairs;
    auto launch = [&]() {
        if (thin) ZK_KLAUNCH(ctx, k_round_quad_thin, dim3(R.n_blocks), dim3(kRoundBlock), 0, A);
#ifndef ZK_EMU
        else if (fold && n >= ctx->tma_min_entries) launch_round_tma(ctx, ZK_PROF_OTHER, 0, A, limit_pairs);
#endif
        else ZK_KLAUNCH(ctx, k_round_quad, dim3(R.n_blocks), dim3(kRoundBlock), 0, A);
    };
    launch();   // warm-up
    rt::event_t e0 = rt::event_create(), e1 = rt::event_create();
    rt::event_record(e0, ctx->stream);
    for (uint32_t i = 0; i < iters; ++i) launch();
    rt::event_record(e1, ctx->stream);
    rt::event_sync(e1);
    *ms = rt::event_elapsed_ms(e0, e1) / iters;
    rt::event_destroy(e0);
    rt::event_destroy(e1);
    ZK_API_END
}

int zk_bench_msm(zk_ctx *ctx, uint32_t log_rows, uint32_t log_cols, int scalar_mix, uint32_t iters, float *ms) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && ms && log_rows <= 14 && log_cols >= 1 && log_cols <= 14 && iters >= 1 && (scalar_mix == 0 || scalar_mix == 2), "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const uint32_t rows = 1u << log_rows, cols = 1u << log_cols;
    // generators: (j + 1) * G
    std::vector<g1_jac_t> base(cols);
    std::vector<fr_t> k(cols);
    g1_jac_t gen;
    memcpy(gen.x.v, ZK_C(g1_gen_x_mont), 48);
    memcpy(gen.y.v, ZK_C(g1_gen_y_mont), 48);
    gen.z = fp_t::one();
    for (uint32_t j = 0; j < cols; ++j) { base[j] = gen; k[j] = fr_t::from_u64(j + 1); }
    rt::dbuf db, dk, dg;
    db.ensure((size_t) cols * sizeof(g1_jac_t)); dk.ensure((size_t) cols * 32); dg.ensure((size_t) cols * sizeof(g1_jac_t));
    rt::h2d(db.p, base.data(), (size_t) cols * sizeof(g1_jac_t), ctx->stream);
    rt::h2d(dk.p, k.data(), (size_t) cols * 32, ctx->stream);
    ZK_KLAUNCH(ctx, k_g1_vec_op, dim3((cols + 63) / 64), dim3(64), 0, db.as<g1_jac_t>(), (const g1_jac_t *) nullptr, dk.as<fr_t>(), dg.as<g1_jac_t>(), cols, 2);
    std::vector<g1_jac_t> gens(cols);
    rt::d2h(gens.data(), dg.p, (size_t) cols * sizeof(g1_jac_t), ctx->stream);
    rt::sync(ctx->stream);
    hyrax_t H;
    msm_prepare_table(ctx, H, reinterpret_cast<const uint64_t *>(gens.data()), cols);
    rt::dbuf ds, dout;
    const uint64_t n = (uint64_t) rows * cols;
    ds.ensure(n * 32);
    dout.ensure((size_t) rows * sizeof(g1_jac_t));
    ZK_KLAUNCH(ctx, k_fill_synthetic, dim3(grid_for(n)), dim3(kBlock), 0, ds.as<fr_t>(), n, 0x9E3779B97F4A7C15ULL, scalar_mix);
    msm_run(ctx, H, ds.as<fr_t>(), cols, rows, dout.as<g1_jac_t>());   // warm-up
    rt::event_t e0 = rt::event_create(), e1 = rt::event_create();
    rt::event_record(e0, ctx->stream);
    for (uint32_t i = 0; i < iters; ++i) msm_run(ctx, H, ds.as<fr_t>(), cols, rows, dout.as<g1_jac_t>());
    rt::event_record(e1, ctx->stream);
    rt::event_sync(e1);
    *ms = rt::event_elapsed_ms(e0, e1) / iters;
    rt::event_destroy(e0);
    rt::event_destroy(e1);
    ZK_API_END
}

}  // extern "C"
