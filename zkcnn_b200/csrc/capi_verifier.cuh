// C ABI, part 4: the verifier's heavy work on the device: the wiring predicates of every layer
// (verifier::betaInitPhase1/2, predicatePhase1/2, src/verifier.cpp:36-116) and the input-layer term gr (:307-325).  The same kernels
// as the prover's Init* passes (eq tables, gate gather-reduce over the resident topology, dot products), driven by the VERIFIER's own
// challenges: nothing of the prover's table state is reused.  Included by capi.cu (after capi_hyrax.cuh: k_dot_long).
#pragma once
#include "capi_sumcheck.cuh"

namespace zk {

constexpr int kVerifierSlots = 8;

static fr_t *vslot(zk_ctx *ctx, int slot, uint64_t entries) {
    ZK_REQUIRE(slot >= 0 && slot < kVerifierSlots, "bad table slot");
    ctx->vt[slot].ensure(std::max<uint64_t>(1, entries) * sizeof(fr_t));
    ctx->vt_n[slot] = entries;
    return ctx->vt[slot].as<fr_t>();
}
// <a, b> over n entries -> host
static fr_t device_dot(zk_ctx *ctx, const fr_t *a, const fr_t *b, uint64_t n) {
    if (n == 0) return fr_t::zero();
    ensure_round_scratch(ctx);
    ZK_KLAUNCH_C(ctx, ZK_PROF_DENSE, n * 64, k_dot_long, dim3(grid_for(n)), dim3(kBlock), 0, a, b, n, ctx->partials.as<fr_t>(), ctx->counters.as<uint32_t>() + 3, ctx->round_out.as<fr_t>());
    rt::d2h(ctx->h_out, ctx->round_out.p, sizeof(fr_t), ctx->stream);
    rt::sync(ctx->stream);
    return ctx->h_out[0];
}

}  // namespace zk

extern "C" {

// slot <- init0 * eq(r0) (+ init1 * eq(r1) when r1 != NULL) over `bits` variables, entries from tail_start on scaled by tail_scale
// (initBetaTable, src/utils.cpp:147-180; the relu_rou scaling of src/verifier.cpp:81-82)
int zk_vtab_eq(zk_ctx *ctx, int slot, uint32_t bits, const uint64_t *r0, const uint64_t *init0, const uint64_t *r1, const uint64_t *init1, uint32_t tail_start,
               const uint64_t *tail_scale) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && bits <= 30 && (r0 || bits == 0) && init0, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    fr_t *out = vslot(ctx, slot, 1ull << bits);
    std::vector<fr_t> a(bits), b(bits);
    for (uint32_t i = 0; i < bits; ++i) { a[i] = fr_load(r0 + 4 * i); if (r1) b[i] = fr_load(r1 + 4 * i); }
    beta_point_t pts[2] = {{a.data(), fr_load(init0)}, {b.data(), r1 && init1 ? fr_load(init1) : fr_t::zero()}};
    const bool any = !pts[0].init.is_zero() || !pts[1].init.is_zero();
    if (!any) rt::dzero(out, sizeof(fr_t) << bits, ctx->stream);
    else build_beta(ctx, out, bits, pts, r1 ? 2 : 1, tail_scale ? tail_start : 0xffffffffu, tail_scale ? fr_load(tail_scale) : fr_t::one());
    rt::sync(ctx->stream);   // (the challenge vectors above are read by the copies queued in build_beta)
    ZK_API_END
}

// slot_out[g] = hi[g >> lo_bits] * lo[g & (2^lo_bits - 1)]   (the PADDING layer's beta_g, src/verifier.cpp:57-66)
int zk_vtab_outer(zk_ctx *ctx, int slot_out, int slot_hi, int slot_lo, uint32_t bits, uint32_t lo_bits) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && bits <= 30 && lo_bits <= bits && slot_hi >= 0 && slot_hi < kVerifierSlots && slot_lo >= 0 && slot_lo < kVerifierSlots, "bad arguments");
    ZK_REQUIRE(ctx->vt_n[slot_hi] >= (1ull << (bits - lo_bits)) && ctx->vt_n[slot_lo] >= (1ull << lo_bits) && slot_out != slot_hi && slot_out != slot_lo, "table slots too small");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    fr_t *out = vslot(ctx, slot_out, 1ull << bits);
    ZK_KLAUNCH_PDL(ctx, ZK_PROF_TABLES, 32ull << bits, k_beta_outer, dim3(grid_for(1ull << bits)), dim3(kBlock), 0, out, (const fr_t *) ctx->vt[slot_hi].as<fr_t>(),
                   (const fr_t *) ctx->vt[slot_lo].as<fr_t>(), bits, lo_bits, 0xffffffffu, fr_t::one());
    ZK_API_END
}

// slot <- phiGInit(rx, scale, n, is_ifft) (src/utils.cpp:61-103), 2^n entries (the FFT form fills the first 2^(n-1); the rest is zero)
int zk_vtab_phi(zk_ctx *ctx, int slot, const uint64_t *rx, const uint64_t *scale, uint32_t n, int is_ifft) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && rx && scale && n >= 1 && n <= 24, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    fr_t *out = vslot(ctx, slot, 1ull << n);
    rt::dzero(out, sizeof(fr_t) << n, ctx->stream);
    ctx->d_r.ensure(2 * 64 * sizeof(fr_t));
    rt::h2d(ctx->d_r.p, rx, (size_t) n * 32, ctx->stream);
    const fr_t *pw = phi_powers(ctx, n, is_ifft != 0);
    ZK_KLAUNCH_C(ctx, ZK_PROF_TABLES, 32ull << n, k_phi_table, dim3(1), dim3(kBlock), 0, out, (const fr_t *) ctx->d_r.as<fr_t>(), pw, fr_load(scale), (int) n, is_ifft ? 1 : 0);
    rt::sync(ctx->stream);
    ZK_API_END
}

// out = sum_{i < n} a[i] b[i]
int zk_vtab_dot(zk_ctx *ctx, int slot_a, int slot_b, uint64_t n, uint64_t *out) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && out && slot_a >= 0 && slot_a < kVerifierSlots && slot_b >= 0 && slot_b < kVerifierSlots && ctx->vt_n[slot_a] >= n && ctx->vt_n[slot_b] >= n, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    fr_store(out, device_dot(ctx, ctx->vt[slot_a].as<fr_t>(), ctx->vt[slot_b].as<fr_t>(), n));
    ZK_API_END
}

// The five wiring-predicate values of one layer (verifier::predicatePhase1/2, src/verifier.cpp:96-123) from the verifier's own eq tables:
//   uni[k]  = sum over unary gates with u in layer (k ? l-1 : 0) of  beta_g[g] beta_u[u] two_mul[sc]        (NOT yet multiplied by beta_v[0])
//   bin[l]  = sum over binary gates of kind l of  beta_g[g] beta_u[u] beta_v[v] (two_mul[sc])               l = 0: (u0, v0), 1: (u1, v1), 2: (u1, v0)
// slot_beta_v < 0 for layers without a second phase (unary gates only).  out: uni[0], uni[1], bin[0], bin[1], bin[2].
int zk_verifier_layer_predicates(zk_ctx *ctx, uint32_t layer_id, int slot_beta_g, int slot_beta_u, int slot_beta_v, uint64_t *out) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && ctx->circuit_ready && layer_id >= 1 && layer_id < ctx->n_layers && out, "bad arguments");
    ZK_REQUIRE(slot_beta_g >= 0 && slot_beta_g < kVerifierSlots && slot_beta_u >= 0 && slot_beta_u < kVerifierSlots && slot_beta_v < kVerifierSlots, "bad table slot");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    layer_t &L = ctx->layers[layer_id];
    const zk_layer_desc &d = L.d;
    ZK_REQUIRE(d.ty != ZK_LAYER_FFT && d.ty != ZK_LAYER_IFFT, "FFT layers have no gate lists (phi table dot eq table: zk_vtab_dot)");
    ZK_REQUIRE(ctx->vt_n[slot_beta_g] >= (1ull << (d.ty == ZK_LAYER_DOT_PROD ? d.bit_length - d.fft_bit_length : d.bit_length)), "beta_g table too small");
    fr_t res[5] = {fr_t::zero(), fr_t::zero(), fr_t::zero(), fr_t::zero(), fr_t::zero()};
    ensure_round_scratch(ctx);
    gate_args_t A;
    memset(&A, 0, sizeof A);
    A.beta_g = ctx->vt[slot_beta_g].as<fr_t>();
    A.two_mul = ctx->two_mul.as<fr_t>();
    ctx->scalar_slot.ensure(sizeof(fr_t));
    A.out_scalar = ctx->scalar_slot.as<fr_t>();
    if (d.need_phase2) {
        ZK_REQUIRE(slot_beta_v >= 0, "this layer has a second phase: beta_v needed");
        const uint64_t n0 = d.bit_length_v[0] >= 0 ? 1ull << d.bit_length_v[0] : 0, n1 = d.bit_length_v[1] >= 0 ? 1ull << d.bit_length_v[1] : 0;
        ZK_REQUIRE(ctx->vt_n[slot_beta_v] >= std::max(n0, n1), "beta_v table too small");
        ctx->vt_m[0].ensure(std::max<uint64_t>(1, n0) * sizeof(fr_t));
        ctx->vt_m[1].ensure(std::max<uint64_t>(1, n1) * sizeof(fr_t));
        A.out0 = ctx->vt_m[0].as<fr_t>();
        A.out1 = ctx->vt_m[1].as<fr_t>();
        A.beta_u = ctx->vt[slot_beta_u].as<fr_t>();
        const fr_t *bv = ctx->vt[slot_beta_v].as<fr_t>();
        for (int kind = 0; kind < 2; ++kind) {   // the gates whose u operand lies in layer 0 (kind 0) / in layer l-1 (kind 1)
            if (n0) rt::dzero(A.out0, n0 * sizeof(fr_t), ctx->stream);
            if (n1) rt::dzero(A.out1, n1 * sizeof(fr_t), ctx->stream);
            rt::dzero(A.out_scalar, sizeof(fr_t), ctx->stream);
            A.vu[kind] = fr_t::one();
            A.vu[kind ^ 1] = fr_t::zero();
            run_schedule(ctx, L.p2, 2, A);
            if (L.p2.has_scalar) {
                rt::d2h(ctx->h_out, ctx->scalar_slot.p, sizeof(fr_t), ctx->stream);
                rt::sync(ctx->stream);
                res[kind] = ctx->h_out[0];
            }
            if (kind == 0) res[2] = device_dot(ctx, A.out0, bv, n0);                     // (u0, v0)
            else { res[4] = device_dot(ctx, A.out0, bv, n0); res[3] = device_dot(ctx, A.out1, bv, n1); }   // (u1, v0), (u1, v1)
        }
    } else {   // unary gates only: phase-1 schedule with the verifier's beta_g, then <beta_u, mult[k]>
        ZK_REQUIRE(d.n_bin == 0, "a layer with binary gates needs its second phase");
        const uint64_t n0 = d.bit_length_u[0] >= 0 ? 1ull << d.bit_length_u[0] : 0, n1 = d.bit_length_u[1] >= 0 ? 1ull << d.bit_length_u[1] : 0;
        ZK_REQUIRE(ctx->vt_n[slot_beta_u] >= std::max(n0, n1), "beta_u table too small");
        ctx->vt_m[0].ensure(std::max<uint64_t>(1, n0) * sizeof(fr_t));
        ctx->vt_m[1].ensure(std::max<uint64_t>(1, n1) * sizeof(fr_t));
        A.out0 = ctx->vt_m[0].as<fr_t>();
        A.out1 = ctx->vt_m[1].as<fr_t>();
        if (n0) rt::dzero(A.out0, n0 * sizeof(fr_t), ctx->stream);
        if (n1) rt::dzero(A.out1, n1 * sizeof(fr_t), ctx->stream);
        A.val0 = ctx->layers[0].val.as<fr_t>();                 // (never read: no binary gates)
        A.val_prev = ctx->layers[layer_id - 1].val.as<fr_t>();
        run_schedule(ctx, L.p1, 1, A);
        const fr_t *bu = ctx->vt[slot_beta_u].as<fr_t>();
        res[0] = device_dot(ctx, A.out0, bu, n0);
        res[1] = device_dot(ctx, A.out1, bu, n1);
    }
    for (int k = 0; k < 5; ++k) fr_store(out + 4 * k, res[k]);
    ZK_API_END
}

// gr of the input-layer check (src/verifier.cpp:307-325): sum over every layer-0 operand slot of eq(r_u[0])[ori_id] times the sigma-weighted
// eq table of the layer it feeds.  r_u / r_v: the challenge vectors of layers 1 .. n_layers-1 back to back (bit_length_u[0] / bit_length_v[0]
// entries per layer that has layer-0 operands, nothing for the others); slot_beta0 = eq(r_u[0]) over the input layer.
int zk_verifier_input_predicate(zk_ctx *ctx, int slot_beta0, const uint64_t *s_u, const uint64_t *s_v, const uint64_t *r_u, const uint64_t *r_v, uint64_t *out) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && ctx->circuit_ready && s_u && s_v && out && slot_beta0 >= 0 && slot_beta0 < kVerifierSlots, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    const zk_layer_desc &d0 = ctx->layers[0].d;
    const uint64_t n0 = 1ull << d0.bit_length;
    ZK_REQUIRE(ctx->vt_n[slot_beta0] >= n0, "beta table of the input layer too small");
    ctx->vt_m[0].ensure(n0 * sizeof(fr_t));
    fr_t *M = ctx->vt_m[0].as<fr_t>();
    rt::dzero(M, n0 * sizeof(fr_t), ctx->stream);
    size_t off_u = 0, off_v = 0;
    for (uint32_t i = 1; i < ctx->n_layers; ++i) {
        layer_t &L = ctx->layers[i];
        for (int side = 0; side < 2; ++side) {
            const int bl = side ? L.d.bit_length_v[0] : L.d.bit_length_u[0];
            const uint32_t sz = side ? L.d.size_v[0] : L.d.size_u[0];
            if (bl < 0) continue;
            size_t &off = side ? off_v : off_u;
            std::vector<fr_t> r(bl);
            for (int k = 0; k < bl; ++k) r[k] = fr_load((side ? r_v : r_u) + 4 * (off + k));
            off += bl;
            const fr_t sigma = fr_load((side ? s_v : s_u) + 4 * (i - 1));
            if (sigma.is_zero() || sz == 0) continue;
            beta_point_t pts[1] = {{r.data(), sigma}};
            halves_t H = build_halves(ctx, bl, pts, 1);
            ZK_KLAUNCH_PDL(ctx, ZK_PROF_DENSE, (uint64_t) sz * 68, k_liu_scatter, dim3(grid_for(sz)), dim3(kBlock), 0, M, (const uint32_t *) (side ? L.ori_v : L.ori_u).as<uint32_t>(), sz,
                           H.f[0], H.s[0], H.first_half);
            rt::sync(ctx->stream);   // `r` is read by the copy queued in build_halves
        }
    }
    fr_store(out, device_dot(ctx, M, ctx->vt[slot_beta0].as<fr_t>(), n0));
    ZK_API_END
}

}  // extern "C"
