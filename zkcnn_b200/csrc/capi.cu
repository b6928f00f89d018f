// C ABI implementation (include/zkcnn_b200.h).  Host-side orchestration of the sm_100a kernels; the control flow of
// every entry point mirrors the reference member function cited in the header, the arithmetic runs on the device.
#include "capi_sumcheck.cuh"
#include "capi_witness.cuh"
#include "capi_hyrax.cuh"
#include "capi_verifier.cuh"

using namespace zk;

extern "C" {

const char *zk_last_error(void) { return g_last_error.c_str(); }

const char *zk_version(void) {
#ifdef ZK_EMU
    return "zkcnn_b200 0.1 emu (test-only host build)";
#else
    return "zkcnn_b200 0.1 sm_100a";
#endif
}

int zk_device_count(void) { return rt::device_count(); }

zk_ctx *zk_ctx_create(int device) {
    try {
        ZK_REQUIRE(rt::device_count() > 0, "no CUDA device: zkcnn_b200 has no CPU fallback");
        ZK_REQUIRE(device >= 0 && device < rt::device_count(), "bad device index");
        rt::set_device(device);
        std::unique_ptr<zk_ctx> c(new zk_ctx);
        c->device = device;
        c->stream = rt::stream_create();
        zk_ctx *ctx = c.release();
        if (const char *env = getenv("ZK_TUNABLES")) {   // "name=value,name=value": kernel-selection tunables for contexts the caller does not hold (tests)
            std::string e(env);
            size_t pos = 0;
            while (pos < e.size()) {
                const size_t end = e.find(',', pos), eq = e.find('=', pos);
                if (eq != std::string::npos && (end == std::string::npos || eq < end))
                    zk_set_tunable(ctx, e.substr(pos, eq - pos).c_str(), strtoull(e.c_str() + eq + 1, nullptr, 0));
                if (end == std::string::npos) break;
                pos = end + 1;
            }
        }
        return ctx;
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return nullptr;
    }
}

void zk_ctx_destroy(zk_ctx *ctx) {
    if (!ctx) return;
    zk_stream_t s0 = ctx->stream, s1 = ctx->aux_stream, s2 = ctx->copy_stream;
    try { rt::set_device(ctx->device); } catch (...) {}
    for (zk_stream_t s : {s2, s1, s0})
        if (s) { try { rt::sync(s); } catch (...) {} }
    rt::unbind();   // everything is synchronised: the buffers freed below need no fence (and the streams are about to go)
    rt::hfree_pinned(ctx->h_out);
    if (ctx->msm_S_h) rt::hfree_pinned(ctx->msm_S_h);
    rt::hfree_pinned(ctx->res_h);
    if (ctx->batch_h) rt::hfree_pinned(ctx->batch_h);
    if (ctx->stage_h) rt::hfree_pinned(ctx->stage_h);
    for (auto &r : ctx->prof_pending) { rt::event_destroy(r.a); rt::event_destroy(r.b); }
    for (auto e : ctx->prof_pool) rt::event_destroy(e);
    delete ctx;
    for (zk_stream_t s : {s2, s1, s0})
        if (s) rt::stream_destroy(s);
}

uint64_t zk_ctx_launch_count(const zk_ctx *ctx) { return ctx ? ctx->launches : 0; }

int zk_set_tunable(zk_ctx *ctx, const char *name, uint64_t value) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && name, "bad arguments");
    const std::string n(name);
    if (n == "thin_max_pairs") ctx->thin_max_pairs = (uint32_t) value;
    else if (n == "tma_min_entries") ctx->tma_min_entries = value;
    else if (n == "derive_b") ctx->derive_b_enabled = value ? 1u : 0u;
    else if (n == "pdl") ctx->pdl_enabled = value ? 1u : 0u;
    else if (n == "eval_schedules") ctx->eval_schedules = value != 0;
    else if (n == "axpy_splits") ctx->axpy_splits = (uint32_t) std::min<uint64_t>(value, 64);
    else if (n == "unit_batch") ctx->unit_batch = value ? 1u : 0u;
    else if (n == "tail") ctx->tail_enabled = value ? 1u : 0u;
    else if (n == "tail_max_entries") ctx->tail_max_entries = (uint32_t) std::min<uint64_t>(value, kTailMaxEntries);
    else if (n == "cubic_tma") ctx->cubic_tma_enabled = value ? 1u : 0u;
    else if (n == "cubic_max_grid") ctx->cubic_max_grid = (uint32_t) std::max<uint64_t>(1, value);
    else if (n == "cubic_factored_min_iters") ctx->cubic_factored_min_iters = (uint32_t) std::max<uint64_t>(1, value);
    else if (n == "msm_split") ctx->msm_split = value ? 1u : 0u;
    else if (n == "msm_digit_bits") { ZK_REQUIRE(value >= 6 && value <= 8, "msm_digit_bits: 6, 7 or 8"); ctx->msm_digit_bits = (uint32_t) value; }
    else if (n == "msm_batch_chunk") ctx->msm_batch_chunk = (uint32_t) std::max<uint64_t>(256, std::min<uint64_t>(value, kMsmChunk));
    else if (n == "msm_small_seg") ctx->msm_small_seg = (uint32_t) std::max<uint64_t>(32, std::min<uint64_t>(value, 65536));
    else if (n == "msm_host_finish") ctx->msm_host_finish = value ? 1u : 0u;
    else if (n == "msm_few_rows_chunk") ctx->msm_few_rows_chunk = (uint32_t) std::max<uint64_t>(256, std::min<uint64_t>(value, kMsmChunk));
    else ZK_REQUIRE(false, "unknown tunable");
    ZK_API_END
}

int zk_profile_enable(zk_ctx *ctx, int on) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx, "null ctx");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    rt::sync(ctx->stream);
    prof_resolve(ctx);
    if (!on) prof_print_kernels(ctx);   // (ZK_PROF_KERNELS=1)
    for (int c = 0; c < ZK_PROF_CLASSES; ++c) { ctx->prof_ms[c] = 0; ctx->prof_launches[c] = 0; ctx->prof_bytes[c] = 0; }
    if (ctx->prof_ops.p) { rt::dzero(ctx->prof_ops.p, 64, ctx->stream); rt::sync(ctx->stream); }
    ctx->prof_on = on != 0;
    ZK_API_END
}

// point additions the MSM kernels performed since zk_profile_enable(ctx, 1): out[0] = mixed additions of the small-multiples kernel (one per
// non-zero one-byte scalar), out[1] = mixed additions into buckets of the window kernel (its bucket reductions are not counted)
int zk_profile_msm_ops(zk_ctx *ctx, uint64_t *out) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && out, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    out[0] = out[1] = 0;
    if (ctx->prof_ops.p) {
        unsigned long long h[2];
        d2h_staged(ctx, h, ctx->prof_ops.p, sizeof h);
        out[0] = h[0];
        out[1] = h[1];
    }
    ZK_API_END
}

int zk_profile_get(zk_ctx *ctx, int cls, double *ms, uint64_t *launches, uint64_t *bytes) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && cls >= 0 && cls < ZK_PROF_CLASSES, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    rt::sync(ctx->stream);
    prof_resolve(ctx);
    if (ms) *ms = ctx->prof_ms[cls];
    if (launches) *launches = ctx->prof_launches[cls];
    if (bytes) *bytes = ctx->prof_bytes[cls];
    ZK_API_END
}

int zk_host_pin(const void *p, size_t bytes) {
    ZK_API_BEGIN
    ZK_REQUIRE(p && bytes, "bad arguments");
    rt::host_pin(const_cast<void *>(p), bytes);
    ZK_API_END
}

int zk_host_unpin(const void *p) {
    ZK_API_BEGIN
    ZK_REQUIRE(p, "bad arguments");
    rt::host_unpin(const_cast<void *>(p));
    ZK_API_END
}

// ---- circuit ----------------------------------------------------------------------------------------------------------
int zk_circuit_begin(zk_ctx *ctx, uint32_t n_layers, const uint64_t *two_mul, uint32_t n_two_mul) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && n_layers >= 2 && n_layers <= 255, "bad layer count");
    ZK_REQUIRE(two_mul && n_two_mul >= 1 && n_two_mul <= 512, "bad two_mul table");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    ctx->layers.clear();
    ctx->layers.resize(n_layers);
    ctx->n_layers = n_layers;
    ctx->circuit_ready = false;
    ctx->two_mul_h.resize(n_two_mul);
    memcpy(ctx->two_mul_h.data(), two_mul, (size_t) n_two_mul * 32);
    ctx->two_mul.ensure(512 * sizeof(fr_t));
    rt::h2d(ctx->two_mul.p, two_mul, (size_t) n_two_mul * 32, ctx->stream);
    rt::sync(ctx->stream);
    ZK_API_END
}

int zk_circuit_layer(zk_ctx *ctx, uint32_t id, const zk_layer_desc *D) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && D && id < ctx->n_layers, "bad layer id");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    layer_t &L = ctx->layers[id];
    L.d = *D;
    L.d.uni_gates = nullptr; L.d.bin_gates = nullptr; L.d.ori_id_u = nullptr; L.d.ori_id_v = nullptr;
    L.scale = fr_load(D->scale);
    ZK_REQUIRE(D->bit_length >= 0 && D->bit_length <= 29, "layer too large");
    if (D->size_u[0]) {
        ZK_REQUIRE(D->ori_id_u, "ori_id_u missing");
        L.ori_u.ensure((size_t) D->size_u[0] * 4);
        rt::h2d(L.ori_u.p, D->ori_id_u, (size_t) D->size_u[0] * 4, ctx->stream);
    }
    if (D->size_v[0]) {
        ZK_REQUIRE(D->ori_id_v, "ori_id_v missing");
        L.ori_v.ensure((size_t) D->size_v[0] * 4);
        rt::h2d(L.ori_v.p, D->ori_id_v, (size_t) D->size_v[0] * 4, ctx->stream);
    }
    rt::sync(ctx->stream);
    build_layer_schedules(ctx, id, D);
    L.have_desc = true;
    L.aux_loaded = false;
    ZK_API_END
}

int zk_circuit_end(zk_ctx *ctx) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx, "null ctx");
    for (auto &L : ctx->layers) ZK_REQUIRE(L.have_desc, "a layer was not uploaded");
    ctx->circuit_ready = true;
    ZK_API_END
}

int zk_witness_layer(zk_ctx *ctx, uint32_t id, const uint64_t *val, uint64_t n) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && id < ctx->n_layers && (val || n == 0), "bad witness layer");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    layer_t &L = ctx->layers[id];
    // layer 0 is kept zero-padded to a power of two so that Hyrax can alias it (src/prover.cpp:504-508)
    uint64_t cap = n;
    if (id == 0) { cap = 1; while (cap < n) cap <<= 1; }
    L.val.ensure(std::max<uint64_t>(1, cap) * sizeof(fr_t));
    rt::h2d(L.val.p, val, n * sizeof(fr_t), ctx->stream);
    if (cap > n) rt::dzero(L.val.as<fr_t>() + n, (cap - n) * sizeof(fr_t), ctx->stream);
    rt::sync(ctx->stream);
    L.n_val = n;
    ZK_API_END
}

// The NEXT proof's witness, copied on a second stream while the current proof is running (double buffering of the
// host->device transfer).  The host buffer must stay valid until zk_witness_commit_prefetch.
int zk_witness_layer_prefetch(zk_ctx *ctx, uint32_t id, const uint64_t *val, uint64_t n) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && id < ctx->n_layers && (val || n == 0), "bad witness layer");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    if (!ctx->copy_stream) ctx->copy_stream = rt::stream_create(false);
    layer_t &L = ctx->layers[id];
    uint64_t cap = n;
    if (id == 0) { cap = 1; while (cap < n) cap <<= 1; }
    L.val_next.ensure(std::max<uint64_t>(1, cap) * sizeof(fr_t));
    // In pieces, with at most two of them queued: copies of different streams are served in submission order, so a small
    // host->device copy of the running proof (challenges, generators) would otherwise wait for this whole layer (the
    // 2^24-entry input layer alone is 10 ms of PCIe time).  The pacing blocks the caller: call this from a helper thread.
    const size_t piece = 4u << 20, bytes = n * sizeof(fr_t);
    // preferred: the SMs pull the layer out of mapped host memory (k_copy_from_host), leaving the copy engine to the proof
#ifndef ZK_EMU   // (the test emulator runs one grid at a time: no launches from the helper thread there)
    const void *mapped = bytes ? rt::host_device_ptr(val) : nullptr;
    if (mapped && bytes % 16 == 0) {
        ZK_LAUNCH(k_copy_from_host, dim3(bytes >= (64u << 20) ? 32 : 8), dim3(kBlock), 0, ctx->copy_stream, static_cast<uint4 *>(L.val_next.p),
                  static_cast<const uint4 *>(mapped), (uint64_t) (bytes / 16));
        rt::check_launch("k_copy_from_host");
        if (cap > n) rt::dzero(L.val_next.as<fr_t>() + n, (cap - n) * sizeof(fr_t), ctx->copy_stream);
        L.n_val_next = n;
        L.next_ready = true;
        return 0;
    }
    cudaEvent_t ev[2];
    for (auto &e : ev) rt::check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
#endif
    size_t k = 0;
    for (size_t off = 0; off < bytes; off += piece, ++k) {
#ifndef ZK_EMU
        if (k >= 2) rt::check(cudaEventSynchronize(ev[k & 1]), "cudaEventSynchronize");
#endif
        rt::h2d(static_cast<char *>(L.val_next.p) + off, reinterpret_cast<const char *>(val) + off, std::min(piece, bytes - off), ctx->copy_stream);
#ifndef ZK_EMU
        rt::check(cudaEventRecord(ev[k & 1], ctx->copy_stream), "cudaEventRecord");
#endif
    }
#ifndef ZK_EMU
    for (auto &e : ev) cudaEventDestroy(e);
#endif
    if (cap > n) rt::dzero(L.val_next.as<fr_t>() + n, (cap - n) * sizeof(fr_t), ctx->copy_stream);
    L.n_val_next = n;
    L.next_ready = true;
    ZK_API_END
}
// prover::val[layer_id] in the compact encoding: small[i] is the value as a signed 64-bit integer (mcl's sign convention: a
// field element >= (r+1)/2 is negative), except at the n_wide positions wide_idx[], whose full values are wide_val[] (and
// whose small[] entry is ignored).  prefetch != 0: into the shadow buffer, on the copy stream (see zk_witness_layer_prefetch).
int zk_witness_layer_compact(zk_ctx *ctx, uint32_t id, const int64_t *small, uint64_t n, const uint32_t *wide_idx, const uint64_t *wide_val,
                             uint32_t n_wide, int prefetch) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && id < ctx->n_layers && (small || n == 0) && (n_wide == 0 || (wide_idx && wide_val)), "bad witness layer");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    if (prefetch && !ctx->copy_stream) ctx->copy_stream = rt::stream_create(false);
    zk_stream_t strm = prefetch ? ctx->copy_stream : ctx->stream;
    layer_t &L = ctx->layers[id];
    uint64_t cap = n;
    if (id == 0) { cap = 1; while (cap < n) cap <<= 1; }
    rt::dbuf &dst = prefetch ? L.val_next : L.val;
    dst.ensure(std::max<uint64_t>(1, cap) * sizeof(fr_t));
    fr_t *out = dst.as<fr_t>();
#ifdef ZK_EMU
    for (uint64_t i = 0; i < n; ++i) out[i] = fr_t::from_i64(small[i]);   // (no kernel launches from a helper thread on the emulator)
    for (uint32_t i = 0; i < n_wide; ++i) memcpy(out + wide_idx[i], wide_val + 4 * (size_t) i, sizeof(fr_t));
#else
    if (n) {
        const void *src = rt::host_device_ptr(small);   // page-locked + mapped: the SMs read it in place
        if (!src) {
            L.compact_stage.ensure(n * 8);
            rt::h2d(L.compact_stage.p, small, n * 8, strm);
            src = L.compact_stage.p;
        }
        ZK_LAUNCH(k_expand_i64, dim3(n >= (1u << 22) ? 64 : 8), dim3(kBlock), 0, strm, out, static_cast<const long long *>(src), n);
        rt::check_launch("k_expand_i64");
        ++ctx->launches;
    }
    if (n_wide) {
        const size_t val_off = ((size_t) n_wide * 4 + 31) & ~(size_t) 31;   // indices first, then the 32-byte values
        L.wide_stage.ensure(val_off + (size_t) n_wide * 32);
        uint32_t *di = L.wide_stage.as<uint32_t>();
        fr_t *dv = reinterpret_cast<fr_t *>(static_cast<char *>(L.wide_stage.p) + val_off);
        rt::h2d(di, wide_idx, (size_t) n_wide * 4, strm);
        rt::h2d(dv, wide_val, (size_t) n_wide * 32, strm);
        ZK_LAUNCH(k_scatter_fr, dim3(grid_for(n_wide)), dim3(kBlock), 0, strm, out, di, dv, n_wide);
        rt::check_launch("k_scatter_fr");
        ++ctx->launches;
    }
#endif
    if (cap > n) rt::dzero(out + n, (cap - n) * sizeof(fr_t), strm);
    if (prefetch) {
        L.n_val_next = n;
        L.next_ready = true;
    } else {
        rt::sync(strm);
        L.n_val = n;
    }
    ZK_API_END
}
// wait for the prefetched layers and make them the current witness (the previous buffers become the next shadow copies)
int zk_witness_commit_prefetch(zk_ctx *ctx) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx, "null ctx");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    if (ctx->copy_stream) rt::sync(ctx->copy_stream);
    rt::sync(ctx->stream);
    for (auto &L : ctx->layers)
        if (L.next_ready) {
            std::swap(L.val, L.val_next);
            L.n_val = L.n_val_next;
            L.next_ready = false;
        }
    ZK_API_END
}

// ---- GKR prover -------------------------------------------------------------------------------------------------------
int zk_prover_init(zk_ctx *ctx) {   // prover::init, src/prover.cpp:17-21
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && ctx->circuit_ready, "circuit not uploaded");
    ctx->r_u.assign(ctx->n_layers + 1, {});
    ctx->r_v.assign(ctx->n_layers + 1, {});
    ZK_API_END
}

int zk_vres(zk_ctx *ctx, const uint64_t *r, uint32_t output_size, uint32_t r_size, uint64_t *out) {   // src/prover.cpp:434-457
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && ctx->circuit_ready && out, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    layer_t &L = ctx->layers[ctx->n_layers - 1];
    ZK_REQUIRE(output_size <= L.n_val && r_size <= 24 && (r || r_size == 0), "bad output size");
    ensure_round_scratch(ctx);
    ctx->d_r.ensure(2 * 64 * sizeof(fr_t));
    rt::h2d(ctx->d_r.p, r, (size_t) r_size * 32, ctx->stream);
    ctx->vres_scratch.ensure(sizeof(fr_t) << r_size);
    ZK_KLAUNCH(ctx, k_vres, dim3(1), dim3(kBlock), 0, L.val.as<fr_t>(), output_size, ctx->d_r.as<fr_t>(), r_size,
               ctx->vres_scratch.as<fr_t>(), ctx->round_out.as<fr_t>());
    rt::d2h(ctx->h_out, ctx->round_out.p, sizeof(fr_t), ctx->stream);
    rt::sync(ctx->stream);
    fr_store(out, ctx->h_out[0]);
    ZK_API_END
}

int zk_sumcheck_init_all(zk_ctx *ctx, const uint64_t *r_0, uint32_t n) {   // src/prover.cpp:28-36
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && ctx->circuit_ready && !ctx->r_u.empty(), "prover not initialised");
    ctx->sumcheck_id = ctx->n_layers;
    const int last_bl = ctx->layers[ctx->n_layers - 1].d.bit_length;
    ZK_REQUIRE((int) n >= last_bl, "r_0 too short");
    ctx->r_u[ctx->sumcheck_id].resize(last_bl);
    for (int i = 0; i < last_bl; ++i) ctx->r_u[ctx->sumcheck_id][i] = fr_load(r_0 + 4 * i);
    ZK_API_END
}

int zk_sumcheck_init(zk_ctx *ctx, const uint64_t *alpha, const uint64_t *beta) {   // src/prover.cpp:43-52
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && ctx->circuit_ready && ctx->sumcheck_id >= 1, "bad state");
    ctx->alpha = fr_load(alpha);
    ctx->beta = fr_load(beta);
    --ctx->sumcheck_id;   // r_0 / r_1 are r_u / r_v of level sumcheck_id + 1 from here on
    ZK_API_END
}

int zk_sumcheck_init_phase1(zk_ctx *ctx, const uint64_t *relu_rou_p) {   // src/prover.cpp:155-239
    ZK_API_BEGIN
    layer_t &L = cur_layer(ctx);
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const zk_layer_desc &d = L.d;
    const uint32_t id = ctx->sumcheck_id;
    ZK_REQUIRE(id >= 1 && d.ty != ZK_LAYER_DOT_PROD, "wrong init for this layer");
    for (int b = 0; b < 2; ++b) pair_reset(ctx->pair[b], d.bit_length_u[b], d.size_u[b]);
    ctx->in_dotprod_p1 = false;
    ctx->r_u[id].resize(d.max_bl_u);
    ctx->relu_rou = fr_load(relu_rou_p);
    ctx->add_term = fr_t::zero();
    const std::vector<fr_t> &r0 = ctx->r_u[id + 1];
    const std::vector<fr_t> &r1 = ctx->r_v[id + 1];
    layer_t &prev = ctx->layers[id - 1];
    const uint32_t tail_start = d.zero_start_id < d.size ? d.zero_start_id : 0xffffffffu;

    if (d.ty == ZK_LAYER_FFT || d.ty == ZK_LAYER_IFFT) {
        const bool is_fft = d.ty == ZK_LAYER_FFT;
        const uint32_t fft_bl = d.fft_bit_length, fft_blh = fft_bl - 1;
        const uint32_t cnt_bl = is_fft ? d.bit_length - fft_bl : d.bit_length - fft_blh;
        const uint32_t cnt_len = d.size >> (is_fft ? fft_bl : fft_blh);
        ctx->beta_g.ensure(sizeof(fr_t) << cnt_bl);
        ctx->beta_g_entries = 1u << cnt_bl;
        if (is_fft) {
            ZK_REQUIRE(r0.size() >= fft_bl + cnt_bl, "r_0 too short");
            ZK_REQUIRE(ctx->beta.is_zero() || r1.size() >= cnt_bl, "r_1 too short");
            beta_point_t pts[2] = {{r0.data() + fft_bl, ctx->alpha}, {r1.data(), ctx->beta}};
            build_beta(ctx, ctx->beta_g.as<fr_t>(), cnt_bl, pts, 2);
        } else {
            ZK_REQUIRE(r0.size() >= fft_blh + cnt_bl, "r_0 too short");
            beta_point_t pts[1] = {{r0.data() + fft_blh, ctx->alpha}};
            build_beta(ctx, ctx->beta_g.as<fr_t>(), cnt_bl, pts, 1);
        }
        // V_mult[1][u] = sum_g val[l][g << max_bl_u | u] * beta_g[g]
        pair_t &P = ctx->pair[1];
        ZK_REQUIRE(P.exists && d.size_u[1] == P.n_eval, "unexpected FFT layer shape");
        ZK_REQUIRE(((uint64_t) cnt_len << d.max_bl_u) <= prev.n_val, "FFT source layer too small");
        fr_t *V = table_init_buf(P.v, P.n_eval);
        ZK_REQUIRE(P.n_eval == (1u << d.max_bl_u), "unexpected FFT layer shape");
        dense_colsum(ctx, prev.val.as<fr_t>(), ctx->beta_g.as<fr_t>(), (uint32_t) d.max_bl_u, cnt_len, V);
        // mult_array[1] = phiGInit(r_0, scale)
        fr_t *M = table_init_buf(P.m, P.n_eval);
        ZK_REQUIRE(r0.size() >= fft_bl - (is_fft ? 0 : 1), "r_0 too short for phi table");
        ctx->d_r.ensure(2 * 64 * sizeof(fr_t));
        rt::h2d(ctx->d_r.p, r0.data(), std::min<size_t>(r0.size(), fft_bl) * sizeof(fr_t), ctx->stream);
        const fr_t *pw = phi_powers(ctx, fft_bl, !is_fft);
        ZK_KLAUNCH_C(ctx, ZK_PROF_TABLES, (uint64_t) P.n_eval * 32, k_phi_table, dim3(1), dim3(kBlock), 0, M, ctx->d_r.as<fr_t>(), pw, L.scale, (int) fft_bl, (int) !is_fft);
    } else {
        // V tables
        if (ctx->pair[0].exists) {
            fr_t *V = table_init_buf(ctx->pair[0].v, ctx->pair[0].n_eval);
            if (d.size_u[0])
                ZK_KLAUNCH_PDL(ctx, ZK_PROF_DENSE, (uint64_t) d.size_u[0] * 68, k_gather, dim3(grid_for(d.size_u[0])), dim3(kBlock), 0, V, ctx->layers[0].val.as<fr_t>(),
                           L.ori_u.as<uint32_t>(), d.size_u[0]);
        }
        if (ctx->pair[1].exists) {
            ZK_REQUIRE(d.size_u[1] <= prev.n_val, "previous layer witness missing");
            ctx->pair[1].v.cur = prev.val.as<fr_t>();
        }
        // beta_g
        if (d.ty == ZK_LAYER_PADDING) {
            const uint32_t fft_blh = d.fft_bit_length - 1;
            ZK_REQUIRE(ctx->beta_g_entries >= (1u << (d.bit_length - fft_blh)), "PADDING layer needs the FFT layer's beta_g");
            ctx->beta_gs.ensure(sizeof(fr_t) << fft_blh);
            beta_point_t pts[1] = {{r0.data(), fr_t::one()}};
            ZK_REQUIRE(r0.size() >= fft_blh, "r_0 too short");
            build_beta(ctx, ctx->beta_gs.as<fr_t>(), fft_blh, pts, 1);
            ctx->beta_g_alt.ensure(sizeof(fr_t) << d.bit_length);
            ZK_KLAUNCH_PDL(ctx, ZK_PROF_TABLES, 32ull << d.bit_length, k_beta_outer, dim3(grid_for(1ull << d.bit_length)), dim3(kBlock), 0, ctx->beta_g_alt.as<fr_t>(),
                       ctx->beta_g.as<fr_t>(), ctx->beta_gs.as<fr_t>(), (uint32_t) d.bit_length, fft_blh, tail_start, ctx->relu_rou);
            std::swap(ctx->beta_g, ctx->beta_g_alt);
        } else {
            ctx->beta_g.ensure(sizeof(fr_t) << d.bit_length);
            ZK_REQUIRE(r0.size() >= (size_t) d.bit_length, "r_0 too short");
            ZK_REQUIRE(ctx->beta.is_zero() || r1.size() >= (size_t) d.bit_length, "r_1 too short");
            beta_point_t pts[2] = {{r0.data(), ctx->alpha * L.scale}, {r1.data(), ctx->beta * L.scale}};
            build_beta(ctx, ctx->beta_g.as<fr_t>(), d.bit_length, pts, 2, tail_start, ctx->relu_rou);
        }
        ctx->beta_g_entries = 1u << d.bit_length;
        // mult tables: gate gather-reduce
        gate_args_t A;
        memset(&A, 0, sizeof A);
        for (int b = 0; b < 2; ++b)
            if (ctx->pair[b].exists) {
                fr_t *M = table_init_buf(ctx->pair[b].m, ctx->pair[b].n_eval);
                rt::dzero(M, (size_t) ctx->pair[b].n_eval * sizeof(fr_t), ctx->stream);
                (b ? A.out1 : A.out0) = M;
            }
        A.beta_g = ctx->beta_g.as<fr_t>();
        A.val0 = ctx->layers[0].val.as<fr_t>();
        A.val_prev = prev.val.as<fr_t>();
        A.two_mul = ctx->two_mul.as<fr_t>();
        run_schedule(ctx, L.p1, 1, A);
    }
    ctx->round = 0;
    ZK_API_END
}

int zk_sumcheck_init_phase2(zk_ctx *ctx) {   // src/prover.cpp:241-310
    ZK_API_BEGIN
    layer_t &L = cur_layer(ctx);
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const zk_layer_desc &d = L.d;
    const uint32_t id = ctx->sumcheck_id;
    ZK_REQUIRE(id >= 1 && d.need_phase2, "layer has no phase 2");
    for (int b = 0; b < 2; ++b) pair_reset(ctx->pair[b], d.bit_length_v[b], d.size_v[b]);
    ctx->in_dotprod_p1 = false;
    ctx->r_v[id].resize(d.max_bl_v);
    ctx->add_term = fr_t::zero();
    layer_t &prev = ctx->layers[id - 1];
    const std::vector<fr_t> &ru = ctx->r_u[id];
    ensure_round_scratch(ctx);

    gate_args_t A;
    memset(&A, 0, sizeof A);
    A.beta_g = ctx->beta_g.as<fr_t>();
    A.two_mul = ctx->two_mul.as<fr_t>();
    A.vu[0] = ctx->V_u0;
    A.vu[1] = ctx->V_u1;
    ctx->scalar_slot.ensure(sizeof(fr_t));
    A.out_scalar = ctx->scalar_slot.as<fr_t>();

    if (d.ty == ZK_LAYER_DOT_PROD) {
        const uint32_t fft_bl = d.fft_bit_length, cnt_bl = d.max_bl_v;
        ZK_REQUIRE(ru.size() >= fft_bl + cnt_bl, "r_u too short");
        ctx->beta_u.ensure(sizeof(fr_t) << cnt_bl);
        ctx->beta_gs.ensure(sizeof(fr_t) << fft_bl);
        beta_point_t p1[1] = {{ru.data() + fft_bl, fr_t::one()}};
        build_beta(ctx, ctx->beta_u.as<fr_t>(), cnt_bl, p1, 1);
        beta_point_t p2[1] = {{ru.data(), fr_t::one()}};
        build_beta(ctx, ctx->beta_gs.as<fr_t>(), fft_bl, p2, 1);
        pair_t &P = ctx->pair[1];
        ZK_REQUIRE(P.exists && !ctx->pair[0].exists, "unexpected DOT_PROD shape");
        ZK_REQUIRE(((uint64_t) d.size_v[1] << fft_bl) <= prev.n_val, "DOT_PROD source layer too small");
        fr_t *V = table_init_buf(P.v, P.n_eval);
        if (d.size_v[1]) {
            const uint32_t lanes = std::max(1u, std::min(32u, (1u << fft_bl) / 8));   // lanes per row: eight terms each, up to a warp
            const uint32_t groups = (d.size_v[1] + kBlock / lanes - 1) / (kBlock / lanes);
            ZK_KLAUNCH_PDL(ctx, ZK_PROF_DENSE, ((uint64_t) d.size_v[1] << fft_bl) * 32 + (uint64_t) d.size_v[1] * 32, k_dense_rowdot, dim3(std::min<uint32_t>(groups, kMaxGridX)),
                           dim3(kBlock), 0, V, prev.val.as<fr_t>(), ctx->beta_gs.as<fr_t>(), d.size_v[1], fft_bl, lanes);
        }
        fr_t *M = table_init_buf(P.m, P.n_eval);
        rt::dzero(M, (size_t) P.n_eval * sizeof(fr_t), ctx->stream);
        A.out1 = M;
        A.beta_u = ctx->beta_u.as<fr_t>();
        run_schedule(ctx, L.p2, 2, A);
    } else {
        ZK_REQUIRE(ru.size() >= (size_t) d.max_bl_u, "r_u too short");
        ctx->beta_u.ensure(sizeof(fr_t) << d.max_bl_u);
        beta_point_t p1[1] = {{ru.data(), fr_t::one()}};
        build_beta(ctx, ctx->beta_u.as<fr_t>(), d.max_bl_u, p1, 1);
        if (ctx->pair[0].exists) {
            fr_t *V = table_init_buf(ctx->pair[0].v, ctx->pair[0].n_eval);
            if (d.size_v[0])
                ZK_KLAUNCH_PDL(ctx, ZK_PROF_DENSE, (uint64_t) d.size_v[0] * 68, k_gather, dim3(grid_for(d.size_v[0])), dim3(kBlock), 0, V, ctx->layers[0].val.as<fr_t>(),
                           L.ori_v.as<uint32_t>(), d.size_v[0]);
        }
        if (ctx->pair[1].exists) {
            ZK_REQUIRE(d.size_v[1] <= prev.n_val, "previous layer witness missing");
            ctx->pair[1].v.cur = prev.val.as<fr_t>();
        }
        for (int b = 0; b < 2; ++b)
            if (ctx->pair[b].exists) {
                fr_t *M = table_init_buf(ctx->pair[b].m, ctx->pair[b].n_eval);
                rt::dzero(M, (size_t) ctx->pair[b].n_eval * sizeof(fr_t), ctx->stream);
                (b ? A.out1 : A.out0) = M;
            }
        A.beta_u = ctx->beta_u.as<fr_t>();
        run_schedule(ctx, L.p2, 2, A);
        if (L.p2.has_scalar) {   // add_term of the uni gates (src/prover.cpp:297-300)
            rt::d2h(ctx->h_out, ctx->scalar_slot.p, sizeof(fr_t), ctx->stream);
            rt::sync(ctx->stream);
            ctx->add_term = ctx->h_out[0];
        }
    }
    ctx->round = 0;
    ZK_API_END
}

static int sumcheck_update(zk_ctx *ctx, const uint64_t *prev_p, uint64_t *abc, std::vector<fr_t> &r_arr) {   // src/prover.cpp:368-383
    ZK_API_BEGIN
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const fr_t prev = fr_load(prev_p);
    if (ctx->round) {
        ZK_REQUIRE(ctx->round - 1 < r_arr.size(), "too many rounds");
        r_arr[ctx->round - 1] = prev;
    }
    ++ctx->round;
    ctx->add_term = ctx->add_term * (fr_t::one() - prev);
    fr_t ret[3];
    round_quadratic(ctx, prev, 3u, ret);
    ret[1] = ret[1] - ctx->add_term;
    ret[2] = ret[2] + ctx->add_term;
    for (int k = 0; k < 3; ++k) fr_store(abc + 4 * k, ret[k]);
    ZK_API_END
}

int zk_sumcheck_update1(zk_ctx *ctx, const uint64_t *prev, uint64_t *abc) {
    if (!ctx || !ctx->circuit_ready || ctx->sumcheck_id >= ctx->r_u.size()) { g_last_error = "bad state"; return -1; }
    return sumcheck_update(ctx, prev, abc, ctx->r_u[ctx->sumcheck_id]);
}
int zk_sumcheck_update2(zk_ctx *ctx, const uint64_t *prev, uint64_t *abc) {
    if (!ctx || !ctx->circuit_ready || ctx->sumcheck_id >= ctx->r_v.size()) { g_last_error = "bad state"; return -1; }
    return sumcheck_update(ctx, prev, abc, ctx->r_v[ctx->sumcheck_id]);
}

// All rounds of one phase in one call (see round_quadratic_batch): which = 1 / 2: sumcheckUpdate1 / sumcheckUpdate2 of the
// current layer, 0: sumcheckLiuUpdate.  prevs[j] is the previous_random of round j (prevs[0] = 0), abc receives n x 3 Fr.
int zk_sumcheck_update_batch(zk_ctx *ctx, int which, const uint64_t *prevs_p, uint32_t n_rounds, uint64_t *abc) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && ctx->circuit_ready && prevs_p && abc && n_rounds >= 1 && which >= 0 && which <= 2, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    std::vector<fr_t> prevs(n_rounds);
    for (uint32_t j = 0; j < n_rounds; ++j) prevs[j] = fr_load(prevs_p + 4 * j);
    if (which == 0) {
        ZK_REQUIRE(ctx->sumcheck_id == 0 && ctx->round == 0, "bad state");
        round_quadratic_batch(ctx, prevs.data(), n_rounds, 2u, [&](uint32_t j, const round_rec_t &rec, const fr_t *h_res) {
            fr_t ret[3];
            round_quadratic_book(ctx, rec, h_res, ret);
            for (int k = 0; k < 3; ++k) fr_store(abc + 12 * j + 4 * k, ret[k]);
        });
    } else {
        ZK_REQUIRE(ctx->sumcheck_id < ctx->r_u.size() && ctx->round == 0, "bad state");
        std::vector<fr_t> &r_arr = which == 1 ? ctx->r_u[ctx->sumcheck_id] : ctx->r_v[ctx->sumcheck_id];
        ZK_REQUIRE(n_rounds <= r_arr.size() + 1, "too many rounds");
        for (uint32_t j = 1; j < n_rounds; ++j) r_arr[j - 1] = prevs[j];
        round_quadratic_batch(ctx, prevs.data(), n_rounds, 3u, [&](uint32_t j, const round_rec_t &rec, const fr_t *h_res) {
            ctx->add_term = ctx->add_term * (fr_t::one() - prevs[j]);   // src/prover.cpp:375-378
            fr_t ret[3];
            round_quadratic_book(ctx, rec, h_res, ret);
            ret[1] = ret[1] - ctx->add_term;
            ret[2] = ret[2] + ctx->add_term;
            for (int k = 0; k < 3; ++k) fr_store(abc + 12 * j + 4 * k, ret[k]);
        });
    }
    ZK_API_END
}

int zk_sumcheck_finalize1(zk_ctx *ctx, const uint64_t *prev_p, uint64_t *claim_0, uint64_t *claim_1) {   // src/prover.cpp:459-471
    ZK_API_BEGIN
    cur_layer(ctx);
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const fr_t prev = fr_load(prev_p);
    std::vector<fr_t> &ru = ctx->r_u[ctx->sumcheck_id];
    ZK_REQUIRE(ctx->round >= 1 && ctx->round - 1 < ru.size(), "finalize without rounds");
    ru[ctx->round - 1] = prev;
    fr_t c[2];
    final_values(ctx, prev, c);
    ctx->V_u0 = c[0];
    ctx->V_u1 = c[1];
    fr_store(claim_0, c[0]);
    fr_store(claim_1, c[1]);
    ctx->pair[0].n_eval = ctx->pair[1].n_eval = 0;
    ZK_API_END
}

int zk_sumcheck_finalize2(zk_ctx *ctx, const uint64_t *prev_p, uint64_t *claim_0, uint64_t *claim_1) {   // src/prover.cpp:473-485
    ZK_API_BEGIN
    cur_layer(ctx);
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const fr_t prev = fr_load(prev_p);
    std::vector<fr_t> &rv = ctx->r_v[ctx->sumcheck_id];
    ZK_REQUIRE(ctx->round >= 1 && ctx->round - 1 < rv.size(), "finalize without rounds");
    rv[ctx->round - 1] = prev;
    fr_t c[2];
    final_values(ctx, prev, c);
    fr_store(claim_0, c[0]);
    fr_store(claim_1, c[1]);
    ctx->pair[0].n_eval = ctx->pair[1].n_eval = 0;
    ZK_API_END
}

// ---- FFT-convolution (DOT_PROD) layer ------------------------------------------------------------------------------------
int zk_sumcheck_dotprod_init_phase1(zk_ctx *ctx) {   // src/prover.cpp:57-95
    ZK_API_BEGIN
    layer_t &L = cur_layer(ctx);
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const zk_layer_desc &d = L.d;
    const uint32_t id = ctx->sumcheck_id;
    ZK_REQUIRE(id >= 1 && d.ty == ZK_LAYER_DOT_PROD, "not a DOT_PROD layer");
    const uint32_t fft_bl = d.fft_bit_length;
    layer_t &prev = ctx->layers[id - 1];
    pair_reset(ctx->pair[0], -1, 0);
    pair_reset(ctx->pair[1], d.bit_length_u[1], d.size_u[1]);
    pair_t &P = ctx->pair[1];
    ZK_REQUIRE(P.exists && d.bit_length_u[1] >= (int) fft_bl && d.size_u[1] <= prev.n_val, "unexpected DOT_PROD shape");
    ctx->r_u[id].resize(d.max_bl_u);
    const std::vector<fr_t> &r0 = ctx->r_u[id + 1];
    ZK_REQUIRE(r0.size() >= fft_bl, "r_0 too short");
    // mult_array[1] = beta table over the frequency index
    ctx->mdp_n = 1u << fft_bl;
    fr_t *M = table_init_buf(ctx->mdp, ctx->mdp_n);
    ctx->mdp.next = 0;
    beta_point_t pts[1] = {{r0.data(), fr_t::one()}};
    build_beta(ctx, M, fft_bl, pts, 1);
    // here pair[1].v plays V_mult[1] (activation FFT | weight FFT) and pair[1].m plays V_mult[0] (sum of beta_g * weight FFT).
    // Only rows u that carry a gate are ever written by the reference's loop (src/prover.cpp:86-91): V_mult[0] is zero from
    // row dp_rows_live on, which the round kernels use (cubic_args_t::live0).
    P.v.cur = prev.val.as<fr_t>();
    fr_t *V0 = table_init_buf(P.m, P.n_eval);
    ZK_REQUIRE(L.dp_rows == (P.n_eval >> fft_bl) && L.dp_rows_live <= L.dp_rows, "DOT_PROD schedule missing");
    ctx->dp_live0 = std::min<uint32_t>(L.dp_rows_live << fft_bl, P.live);
    ctx->in_dotprod_p1 = true;
    const uint64_t n_out = (uint64_t) L.dp_rows_live << fft_bl;
    if (n_out) {
        // enough threads to fill the machine: the gates of a row are dealt to `splits` threads whose partial sums k_colsum_finish adds up
        const uint64_t gates_per_row = std::max<uint64_t>(1, d.n_bin / std::max(1u, L.dp_rows_live));
        const uint32_t splits = ctx->axpy_splits ? ctx->axpy_splits
                                                 : (uint32_t) std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(32, gates_per_row / 8), ((uint64_t) ZK_SM_COUNT * 8 * kBlock) / n_out));
        fr_t *dst = V0;
        if (splits > 1) {
            ctx->dense_partial.ensure(n_out * splits * sizeof(fr_t));
            dst = ctx->dense_partial.as<fr_t>();
        }
        ZK_KLAUNCH_PDL(ctx, ZK_PROF_DENSE, (uint64_t) d.n_bin * (32ull << fft_bl) + n_out * 32, k_dotprod_axpy, dim3(grid_for(n_out * splits)), dim3(kBlock), 0, dst,
                       prev.val.as<fr_t>(), ctx->beta_g.as<fr_t>(), L.dp_rowptr.as<uint32_t>(), L.dp_gates.as<dp_gate_t>(), L.dp_rows_live, fft_bl, splits);
        if (splits > 1) {
            const uint32_t groups = (uint32_t) ((n_out + kBlock / 32 - 1) / (kBlock / 32));
            ZK_KLAUNCH_PDL(ctx, ZK_PROF_DENSE, n_out * (splits + 1) * 32, k_colsum_finish, dim3(std::min<uint32_t>(groups, kMaxGridX)), dim3(kBlock), 0, (const fr_t *) dst, (uint32_t) n_out,
                           splits, V0);
        }
    }
    ctx->round = 0;
    ZK_API_END
}

namespace zk {
// One call of sumcheckDotProdUpdate1 (src/prover.cpp:103-144) queued on the stream.  slot_d == nullptr: interactive round, the
// four coefficients go to the host mailbox (the caller waits); otherwise they are left at slot_d.  The host-side table state
// follows from the sizes alone.
static void cubic_round_launch(zk_ctx *ctx, const fr_t &prev, fr_t *slot_d) {
    ensure_round_scratch(ctx);
    if (!ctx->cubic_acc.p) {
        ctx->cubic_acc.ensure(kCubicLimbs * sizeof(unsigned long long));
        rt::dzero(ctx->cubic_acc.p, kCubicLimbs * sizeof(unsigned long long), ctx->stream);
    }
    const bool first = ctx->round == 1;
    pair_t &P = ctx->pair[1];
    ZK_REQUIRE(P.n_eval >= (first ? 2u : 4u), "DOT_PROD tables exhausted");
    cubic_args_t A;
    memset(&A, 0, sizeof A);
    A.v1_in = P.v.cur; A.v0_in = P.m.cur;
    A.n_in = P.n_eval; A.live1 = std::min(P.live, P.n_eval); A.live0 = std::min(ctx->dp_live0, A.live1); A.fold = first ? 0 : 1;
    const uint32_t n_after = first ? P.n_eval : P.n_eval >> 1;
    if (!first) {
        A.v1_out = table_fold_buf(P.v, n_after);
        A.v0_out = table_fold_buf(P.m, n_after);
    }
    A.m_in = ctx->mdp.cur;
    A.m_n = ctx->mdp_n;
    const bool m_fold = !first && ctx->mdp_n >= 2;   // the multiplier table is folded inside the round kernel (src/prover.cpp:112-118)
    const uint32_t cur_n = m_fold ? ctx->mdp_n >> 1 : ctx->mdp_n;
    if (m_fold) A.m_out = table_fold_buf(ctx->mdp, cur_n);
    A.r = prev;
    A.acc = ctx->cubic_acc.as<unsigned long long>();
    A.counter = ctx->counters.as<uint32_t>() + 2;
    if (slot_d) A.out = slot_d;
    else { A.out = ctx->res_d; A.flag = ctx->flag_d; A.seq = ++ctx->seq; }

    const uint32_t n_pairs = first ? P.n_eval >> 1 : P.n_eval >> 2;
    const uint32_t P0 = std::min(n_pairs, first ? (A.live0 + 1) >> 1 : (A.live0 + 3) >> 2);
    const uint32_t P1 = std::min(n_pairs, first ? (A.live1 + 1) >> 1 : (A.live1 + 3) >> 2);
    const uint32_t period = cur_n >= 2 ? cur_n >> 1 : 1;              // output pairs per period of the multiplier
    const uint32_t unit = std::max(1u, period / kRoundBlock);          // CTAs per period: a multi-iteration grid is a multiple of it
    const uint64_t bytes = first ? (uint64_t) A.live0 * 64 : (uint64_t) A.live1 * 48 + (uint64_t) A.live0 * 48;
    const int cls = bytes >= (32u << 20) ? ZK_PROF_FOLD : ZK_PROF_FOLD_SMALL;
    auto pick_nb0 = [&](uint64_t want, uint64_t share) -> uint32_t {
        if (want <= share) return (uint32_t) std::max<uint64_t>(1, want);   // at most one iteration per thread: any grid will do
        return (uint32_t) std::max<uint64_t>(unit, share / unit * unit);
    };
#ifndef ZK_EMU
    if (!first && P.n_eval >= ctx->tma_min_entries && ctx->cubic_tma_enabled) {
        static const bool attr_set = [] {
            rt::check(cudaFuncSetAttribute(k_round_cubic_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kTmaSmemBytes), "cudaFuncSetAttribute");
            return true;
        }();
        (void) attr_set;
        cubic_tma_args_t T;
        memset(&T, 0, sizeof T);
        const uint32_t max_grid = std::min<uint32_t>(kTmaMaxGrid, ctx->cubic_max_grid);
        const uint64_t G0 = (P0 + 31) >> 5, G1 = (P1 + 31) >> 5, it1 = (G1 - G0 + 1) >> 1;
        const uint64_t w0 = G0 * 704, w1 = it1 * 512;                 // wide multiply-adds per warp iteration (4 folds + 3 products | 4 folds)
        const uint64_t want0 = (G0 + kRoundBlock / 32 - 1) / (kRoundBlock / 32), want1 = (it1 + kRoundBlock / 32 - 1) / (kRoundBlock / 32);
        A.nb0 = pick_nb0(want0, std::max<uint64_t>(1, (uint64_t) max_grid * w0 / std::max<uint64_t>(1, w0 + w1)));
        A.nb1 = it1 ? (uint32_t) std::max<uint64_t>(1, std::min<uint64_t>(want1, max_grid > A.nb0 ? max_grid - A.nb0 : 1)) : 0;
        encode_rows_map(&T.tm_v1, A.v1_in, A.n_in);
        encode_rows_map(&T.tm_v0, A.v0_in, A.n_in);
        T.C = A;
        ZK_KLAUNCH_C(ctx, cls, bytes, k_round_cubic_tma, dim3(A.nb0 + A.nb1), dim3(kRoundBlock), kTmaSmemBytes, T);
    } else
#endif
    {
        const uint32_t max_grid = std::min<uint32_t>(ZK_SM_COUNT * 3, ctx->cubic_max_grid);   // three CTAs per SM resident
        const uint64_t want0 = (P0 + kRoundBlock - 1) / kRoundBlock, want1 = first ? 0 : (P1 - P0 + kRoundBlock - 1) / kRoundBlock;
        const uint64_t w0 = (uint64_t) P0 * (first ? 192 : 704), w1 = (uint64_t) (first ? 0 : P1 - P0) * 256;
        A.nb0 = pick_nb0(want0, std::max<uint64_t>(1, (uint64_t) max_grid * w0 / std::max<uint64_t>(1, w0 + w1)));
        A.nb1 = want1 ? (uint32_t) std::max<uint64_t>(1, std::min<uint64_t>(want1, max_grid > A.nb0 ? max_grid - A.nb0 : 1)) : 0;
        const uint64_t iters = ((uint64_t) P0 + (uint64_t) A.nb0 * kRoundBlock - 1) / ((uint64_t) A.nb0 * kRoundBlock);
        if (iters >= ctx->cubic_factored_min_iters) ZK_KLAUNCH_PDL(ctx, cls, bytes, k_round_cubic<true>, dim3(A.nb0 + A.nb1), dim3(kRoundBlock), 0, A);
        else ZK_KLAUNCH_PDL(ctx, cls, bytes, k_round_cubic<false>, dim3(A.nb0 + A.nb1), dim3(kRoundBlock), 0, A);
    }
    if (m_fold) {
        table_advance(ctx->mdp);
        ctx->mdp_n = cur_n;
    }
    if (!first) {
        table_advance(P.m);
        table_advance(P.v);
        P.n_eval >>= 1;
        P.live = (P.live + 1) >> 1;
        ctx->dp_live0 = (ctx->dp_live0 + 1) >> 1;
    }
}
}  // namespace zk

int zk_sumcheck_dotprod_update1(zk_ctx *ctx, const uint64_t *prev_p, uint64_t *abcd) {   // src/prover.cpp:103-144
    ZK_API_BEGIN
    cur_layer(ctx);
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const fr_t prev = fr_load(prev_p);
    std::vector<fr_t> &ru = ctx->r_u[ctx->sumcheck_id];
    if (ctx->round) {
        ZK_REQUIRE(ctx->round - 1 < ru.size(), "too many rounds");
        ru[ctx->round - 1] = prev;
    }
    ++ctx->round;
    cubic_round_launch(ctx, prev, nullptr);
    wait_mailbox(ctx);
    for (int k = 0; k < 4; ++k) fr_store(abcd + 4 * k, ctx->res_h[k]);
    ZK_API_END
}

// every round of the DOT_PROD phase in one call (see zk_sumcheck_update_batch): abcd receives n_rounds x 4 Fr
int zk_sumcheck_dotprod_update_batch(zk_ctx *ctx, const uint64_t *prevs_p, uint32_t n_rounds, uint64_t *abcd) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && prevs_p && abcd && n_rounds >= 1, "bad arguments");
    cur_layer(ctx);
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    std::vector<fr_t> &ru = ctx->r_u[ctx->sumcheck_id];
    ZK_REQUIRE(ctx->round == 0 && n_rounds <= ru.size() + 1, "bad state");
    ensure_round_scratch(ctx);
    ctx->batch_res.ensure((size_t) n_rounds * 16 * sizeof(fr_t));
    if (ctx->batch_cap < n_rounds) {
        if (ctx->batch_h) rt::hfree_pinned(ctx->batch_h);
        ctx->batch_h = static_cast<fr_t *>(rt::hmalloc_pinned((size_t) n_rounds * 16 * sizeof(fr_t)));
        ctx->batch_cap = n_rounds;
    }
    for (uint32_t j = 0; j < n_rounds; ++j) {
        const fr_t prev = fr_load(prevs_p + 4 * j);
        if (j) ru[j - 1] = prev;
        ++ctx->round;
        cubic_round_launch(ctx, prev, ctx->batch_res.as<fr_t>() + (size_t) j * 16);
    }
    rt::d2h(ctx->batch_h, ctx->batch_res.p, (size_t) n_rounds * 16 * sizeof(fr_t), ctx->stream);
    rt::sync(ctx->stream);
    for (uint32_t j = 0; j < n_rounds; ++j)
        for (int k = 0; k < 4; ++k) fr_store(abcd + 16 * (size_t) j + 4 * k, ctx->batch_h[(size_t) j * 16 + k]);
    ZK_API_END
}

int zk_cubic_rounds(zk_ctx *ctx, const uint64_t *mult, uint32_t m_bits, const uint64_t *V0, uint64_t live0, const uint64_t *V1, uint64_t live1, uint32_t bits,
                    const uint64_t *r, uint32_t n_rounds, uint64_t *polys) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && mult && V0 && V1 && polys && bits >= 1 && bits <= 28 && m_bits <= 12 && m_bits <= bits && n_rounds >= 1 && n_rounds <= bits &&
                   live0 <= live1 && live1 <= (1ull << bits) && live1 >= 1 && (r || n_rounds == 1), "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    pair_t &P = ctx->pair[1];
    pair_reset(ctx->pair[0], -1, 0);
    pair_reset(P, (int8_t) bits, (uint32_t) live1);
    fr_t *dv1 = table_init_buf(P.v, 1ull << bits), *dv0 = table_init_buf(P.m, 1ull << bits);
    rt::h2d(dv1, V1, live1 * 32, ctx->stream);
    rt::h2d(dv0, V0, live0 * 32, ctx->stream);
    ctx->mdp_n = 1u << m_bits;
    fr_t *dm = table_init_buf(ctx->mdp, ctx->mdp_n);
    ctx->mdp.next = 0;
    rt::h2d(dm, mult, (size_t) ctx->mdp_n * 32, ctx->stream);
    ctx->dp_live0 = (uint32_t) live0;
    ctx->round = 0;
    for (uint32_t j = 0; j < n_rounds; ++j) {
        const fr_t prev = j == 0 ? fr_t::zero() : fr_load(r + 4 * (j - 1));
        ++ctx->round;
        cubic_round_launch(ctx, prev, nullptr);
        wait_mailbox(ctx);
        for (int k = 0; k < 4; ++k) fr_store(polys + (size_t) (4 * j + k) * 4, ctx->res_h[k]);
    }
    P.n_eval = 0;
    ZK_API_END
}

int zk_fold_rounds2(zk_ctx *ctx, const uint64_t *V0, const uint64_t *M0, int32_t bits0, uint64_t live0, const uint64_t *V1, const uint64_t *M1, int32_t bits1,
                    uint64_t live1, const uint64_t *r, uint32_t n_rounds, uint64_t *polys) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && polys && bits0 <= 28 && bits1 <= 28 && (bits0 >= 0 || bits1 >= 0) && n_rounds >= 1 && (int) n_rounds <= std::max(bits0, bits1) + (std::max(bits0, bits1) == 0), "bad arguments");
    ZK_REQUIRE((bits0 < 0 || (V0 && M0 && live0 <= (1ull << bits0))) && (bits1 < 0 || (V1 && M1 && live1 <= (1ull << bits1))), "bad tables");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const uint64_t *Vs[2] = {V0, V1}, *Ms[2] = {M0, M1};
    const int32_t bits[2] = {bits0, bits1};
    const uint64_t lives[2] = {live0, live1};
    for (int b = 0; b < 2; ++b) {
        pair_t &P = ctx->pair[b];
        pair_reset(P, (int8_t) bits[b], bits[b] >= 0 ? (uint32_t) lives[b] : 0);
        if (bits[b] < 0) continue;
        fr_t *dv = table_init_buf(P.v, 1ull << bits[b]), *dm = table_init_buf(P.m, 1ull << bits[b]);
        rt::h2d(dv, Vs[b], lives[b] * 32, ctx->stream);
        rt::h2d(dm, Ms[b], lives[b] * 32, ctx->stream);
    }
    ctx->in_dotprod_p1 = false;
    ctx->add_term = fr_t::zero();
    ctx->round = 0;
    if (ctx->unit_batch) {   // the phase-batched path (zk_sumcheck_update_batch), fused tail included
        std::vector<fr_t> prevs(n_rounds, fr_t::zero());
        for (uint32_t j = 1; j < n_rounds; ++j) prevs[j] = fr_load(r + 4 * (j - 1));
        round_quadratic_batch(ctx, prevs.data(), n_rounds, 3u, [&](uint32_t j, const round_rec_t &rec, const fr_t *h_res) {
            ctx->add_term = ctx->add_term * (fr_t::one() - prevs[j]);
            fr_t abc[3];
            round_quadratic_book(ctx, rec, h_res, abc);
            abc[1] = abc[1] - ctx->add_term;
            abc[2] = abc[2] + ctx->add_term;
            for (int k = 0; k < 3; ++k) fr_store(polys + (size_t) (3 * j + k) * 4, abc[k]);
        });
    } else
    for (uint32_t j = 0; j < n_rounds; ++j) {
        const fr_t prev = j == 0 ? fr_t::zero() : fr_load(r + 4 * (j - 1));
        ++ctx->round;
        ctx->add_term = ctx->add_term * (fr_t::one() - prev);
        fr_t abc[3];
        round_quadratic(ctx, prev, 3u, abc);
        abc[1] = abc[1] - ctx->add_term;
        abc[2] = abc[2] + ctx->add_term;
        for (int k = 0; k < 3; ++k) fr_store(polys + (size_t) (3 * j + k) * 4, abc[k]);
    }
    ctx->pair[0].n_eval = ctx->pair[1].n_eval = 0;
    ZK_API_END
}

int zk_mle_eval(zk_ctx *ctx, const uint64_t *values, uint32_t n, const uint64_t *r, uint32_t r_size, uint64_t *out) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && values && out && n >= 1 && r_size <= 24 && n <= (1u << r_size) && (r || r_size == 0), "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    ensure_round_scratch(ctx);
    rt::dbuf dv;
    dv.ensure((size_t) n * 32);
    rt::h2d(dv.p, values, (size_t) n * 32, ctx->stream);
    ctx->d_r.ensure(2 * 64 * sizeof(fr_t));
    rt::h2d(ctx->d_r.p, r, (size_t) r_size * 32, ctx->stream);
    ctx->vres_scratch.ensure(sizeof(fr_t) << r_size);
    ZK_KLAUNCH(ctx, k_vres, dim3(1), dim3(kBlock), 0, dv.as<fr_t>(), n, ctx->d_r.as<fr_t>(), r_size, ctx->vres_scratch.as<fr_t>(), ctx->round_out.as<fr_t>());
    rt::d2h(ctx->h_out, ctx->round_out.p, sizeof(fr_t), ctx->stream);
    rt::sync(ctx->stream);
    fr_store(out, ctx->h_out[0]);
    ZK_API_END
}

int zk_debug_table_hash(zk_ctx *ctx, int sel, uint64_t *fnv1a, uint64_t *n_entries) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && fnv1a && n_entries && sel >= 0 && sel <= 4, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const fr_t *src = nullptr;
    uint64_t n = 0, live = 0;
    if (sel == 4) { src = ctx->mdp.cur; n = live = ctx->mdp_n; }
    else {
        const pair_t &P = ctx->pair[sel >> 1];
        if (P.exists && P.n_eval) {
            src = (sel & 1) ? P.m.cur : P.v.cur;
            n = P.n_eval;
            live = std::min<uint64_t>(n, (sel == 3 && ctx->in_dotprod_p1) ? ctx->dp_live0 : P.live);
        }
    }
    std::vector<fr_t> h(live);
    if (live) d2h_staged(ctx, h.data(), src, live * sizeof(fr_t));
    uint64_t hash = 0xcbf29ce484222325ULL;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (i < live) h[i].to_canonical(c);
        const uint8_t *b = reinterpret_cast<const uint8_t *>(c);
        for (int k = 0; k < 32; ++k) { hash ^= b[k]; hash *= 0x100000001b3ULL; }
    }
    *fnv1a = hash;
    *n_entries = n;
    ZK_API_END
}

// ---- micro-benchmarks of the K2 kernels and of the field multiplier (CUDA events on the launching stream) -------------------------
// sustained Fp (is_fp != 0) or Fr multiplications per second, in units of 10^9
int zk_bench_field_mul(zk_ctx *ctx, int is_fp, float *gmul_per_s) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && gmul_per_s, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const uint32_t grid = ZK_SM_COUNT * 8, iters = 2000;
    rt::dbuf out;
    out.ensure((size_t) grid * kBlock * sizeof(fp_t));
    rt::event_t e0 = rt::event_create(), e1 = rt::event_create();
    for (int pass = 0; pass < 2; ++pass) {   // pass 0 warms up
        rt::event_record(e0, ctx->stream);
        if (is_fp) ZK_KLAUNCH(ctx, k_mul_chain<fp_t>, dim3(grid), dim3(kBlock), 0, out.as<fp_t>(), iters, 7ull);
        else ZK_KLAUNCH(ctx, k_mul_chain<fr_t>, dim3(grid), dim3(kBlock), 0, out.as<fr_t>(), iters, 7ull);
        rt::event_record(e1, ctx->stream);
        rt::event_sync(e1);
    }
    *gmul_per_s = (float) ((double) grid * kBlock * iters * 2 / (rt::event_elapsed_ms(e0, e1) * 1e-3) / 1e9);
    rt::event_destroy(e0);
    rt::event_destroy(e1);
    ZK_API_END
}

// one K2 fold round on tables of 2^bits entries (V_mult[0] live over the first 2^live0_bits), multiplier period 2^m_bits
int zk_bench_cubic(zk_ctx *ctx, uint32_t bits, uint32_t live0_bits, uint32_t m_bits, uint32_t iters, float *ms) {
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && ms && bits >= 3 && bits <= 27 && live0_bits <= bits && m_bits >= 1 && m_bits <= 12 && m_bits < bits && iters >= 1, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const uint64_t n = 1ull << bits;
    pair_t &P = ctx->pair[1];
    pair_reset(ctx->pair[0], -1, 0);
    rt::event_t e0 = rt::event_create(), e1 = rt::event_create();
    float total = 0;
    for (uint32_t i = 0; i <= iters; ++i) {   // iteration 0 warms up; every iteration starts from fresh full-size tables
        pair_reset(P, (int8_t) bits, (uint32_t) n);
        fr_t *dv1 = table_init_buf(P.v, n), *dv0 = table_init_buf(P.m, n);
        if (i == 0) {
            ZK_KLAUNCH(ctx, k_fill_synthetic, dim3(grid_for(n)), dim3(kBlock), 0, dv1, n, 0x9E3779B97F4A7C15ULL, 0);
            ZK_KLAUNCH(ctx, k_fill_synthetic, dim3(grid_for(n)), dim3(kBlock), 0, dv0, n, 0x243F6A8885A308D3ULL, 0);
        }
        ctx->mdp_n = 1u << m_bits;
        fr_t *dm = table_init_buf(ctx->mdp, ctx->mdp_n);
        ctx->mdp.next = 0;
        if (i == 0) ZK_KLAUNCH(ctx, k_fill_synthetic, dim3(grid_for(ctx->mdp_n)), dim3(kBlock), 0, dm, (uint64_t) ctx->mdp_n, 0x13198A2E03707344ULL, 0);
        ctx->dp_live0 = 1u << live0_bits;
        ctx->round = 2;   // a fold round (not the first of its phase)
        ensure_round_scratch(ctx);
        ctx->batch_res.ensure(16 * sizeof(fr_t));
        rt::event_record(e0, ctx->stream);
        cubic_round_launch(ctx, fr_t::from_u64(0x1234567887654321ULL), ctx->batch_res.as<fr_t>());
        rt::event_record(e1, ctx->stream);
        rt::event_sync(e1);
        if (i) total += rt::event_elapsed_ms(e0, e1);
    }
    P.n_eval = 0;
    *ms = total / iters;
    rt::event_destroy(e0);
    rt::event_destroy(e1);
    ZK_API_END
}

int zk_sumcheck_dotprod_finalize1(zk_ctx *ctx, const uint64_t *prev_p, uint64_t *claim_1) {   // src/prover.cpp:146-153
    ZK_API_BEGIN
    cur_layer(ctx);
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    ensure_round_scratch(ctx);
    const fr_t prev = fr_load(prev_p);
    std::vector<fr_t> &ru = ctx->r_u[ctx->sumcheck_id];
    ZK_REQUIRE(ctx->round >= 1 && ctx->round - 1 < ru.size(), "finalize without rounds");
    ru[ctx->round - 1] = prev;
    pair_t &P = ctx->pair[1];
    ZK_REQUIRE(P.n_eval == 2 && ctx->mdp_n >= 1 && ctx->mdp_n <= 2, "finalize called before the last round");
    final_fold_args_t F;
    memset(&F, 0, sizeof F);
    F.r = prev;
    F.out = ctx->res_d + 8;
    F.v_in[0] = P.v.cur; F.live[0] = P.live; F.fold[0] = 1; F.active[0] = 1;
    F.v_in[1] = ctx->mdp.cur; F.live[1] = ctx->mdp_n; F.fold[1] = ctx->mdp_n == 2; F.active[1] = 1;
    F.flag = ctx->flag_d;
    F.seq = ++ctx->seq;
    ZK_KLAUNCH_PDL(ctx, ZK_PROF_OTHER, 0, k_final_fold, dim3(1), dim3(32), 0, F);
    wait_mailbox(ctx);
    const fr_t c1 = ctx->res_h[8], m = ctx->res_h[10];
    ctx->V_u1 = c1 * m;
    fr_store(claim_1, c1);
    P.n_eval = 0;
    ZK_API_END
}

// ---- input layer ("Liu") sumcheck -----------------------------------------------------------------------------------------
int zk_sumcheck_liu_init(zk_ctx *ctx, const uint64_t *s_u, const uint64_t *s_v, uint32_t n) {   // src/prover.cpp:312-358
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && ctx->circuit_ready && !ctx->r_u.empty() && n + 1 >= ctx->n_layers, "bad state");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    ctx->sumcheck_id = 0;
    layer_t &L0 = ctx->layers[0];
    const zk_layer_desc &d0 = L0.d;
    ZK_REQUIRE(L0.n_val >= d0.size, "input layer witness missing");
    pair_reset(ctx->pair[0], -1, 0);
    pair_reset(ctx->pair[1], d0.bit_length, d0.size);
    ctx->in_dotprod_p1 = false;
    pair_t &P = ctx->pair[1];
    ctx->r_u[0].resize(d0.bit_length);
    ctx->add_term = fr_t::zero();
    P.v.cur = L0.val.as<fr_t>();
    fr_t *M = table_init_buf(P.m, P.n_eval);
    rt::dzero(M, (size_t) P.n_eval * sizeof(fr_t), ctx->stream);
    for (uint32_t i = 1; i < ctx->n_layers; ++i) {
        layer_t &L = ctx->layers[i];
        for (int side = 0; side < 2; ++side) {
            const int bl = side ? L.d.bit_length_v[0] : L.d.bit_length_u[0];
            const uint32_t sz = side ? L.d.size_v[0] : L.d.size_u[0];
            if (bl < 0) continue;
            const std::vector<fr_t> &r = side ? ctx->r_v[i] : ctx->r_u[i];
            ZK_REQUIRE(r.size() >= (size_t) bl, "challenge vector of a finished layer is too short");
            const fr_t sigma = fr_load((side ? s_v : s_u) + 4 * (i - 1));
            if (sigma.is_zero() || sz == 0) continue;
            beta_point_t pts[1] = {{r.data(), sigma}};
            halves_t H = build_halves(ctx, bl, pts, 1);
            ZK_KLAUNCH_PDL(ctx, ZK_PROF_DENSE, (uint64_t) sz * 68, k_liu_scatter, dim3(grid_for(sz)), dim3(kBlock), 0, M, (side ? L.ori_v : L.ori_u).as<uint32_t>(), sz, H.f[0],
                       H.s[0], H.first_half);
        }
    }
    ctx->round = 0;
    ZK_API_END
}

int zk_sumcheck_liu_update(zk_ctx *ctx, const uint64_t *prev_p, uint64_t *abc) {   // src/prover.cpp:385-394
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && ctx->circuit_ready && ctx->sumcheck_id == 0, "bad state");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const fr_t prev = fr_load(prev_p);
    ++ctx->round;
    fr_t ret[3];
    round_quadratic(ctx, prev, 2u, ret);
    for (int k = 0; k < 3; ++k) fr_store(abc + 4 * k, ret[k]);
    ZK_API_END
}

int zk_sumcheck_liu_finalize(zk_ctx *ctx, const uint64_t *prev_p, uint64_t *claim_1) {   // src/prover.cpp:487-497
    ZK_API_BEGIN
    ZK_REQUIRE(ctx && ctx->circuit_ready && ctx->sumcheck_id == 0, "bad state");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const fr_t prev = fr_load(prev_p);
    ZK_REQUIRE(ctx->round >= 1 && ctx->round - 1 < ctx->r_u[0].size(), "finalize without rounds");
    ctx->r_u[0][ctx->round - 1] = prev;
    fr_t c[2];
    final_values(ctx, prev, c);
    fr_store(claim_1, c[1]);
    ctx->pair[1].n_eval = 0;
    ZK_API_END
}

}  // extern "C"
