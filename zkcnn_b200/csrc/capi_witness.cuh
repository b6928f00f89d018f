// C ABI, part 3: witness generation on the device (witness_kernels.cuh).  Included by capi.cu only.
#pragma once
#include "capi_sumcheck.cuh"

namespace zk {

static void eval_normal_layer(zk_ctx *ctx, uint32_t id) {
    layer_t &L = ctx->layers[id];
    layer_t &prev = ctx->layers[id - 1];
    fr_t *out = L.val.as<fr_t>();
    const schedule_t &S = L.ev;
    // gates without any source stay zero (src/neuralNetwork.cpp:920); a layer whose every gate has a source is written completely by the items
    if (L.ev_rows_covered < L.d.size || S.levels.empty()) rt::dzero(out, (size_t) L.d.size * sizeof(fr_t), ctx->stream);
    if (S.levels.empty()) return;
    ctx->gate_partial[0].ensure((size_t) std::max(1u, S.max_partials) * sizeof(fr_t));
    ctx->gate_partial[1].ensure((size_t) std::max(1u, S.max_partials) * sizeof(fr_t));
    gate_args_t A;
    memset(&A, 0, sizeof A);
    A.recs = S.recs.as<gate_rec_t>();
    A.val0 = ctx->layers[0].val.as<fr_t>();
    A.val_prev = prev.val.as<fr_t>();
    A.two_mul = ctx->two_mul.as<fr_t>();
    A.out0 = out;
    for (size_t k = 0; k < S.levels.size(); ++k) {
        const level_t &V = S.levels[k];
        A.items = V.items.as<item_t>();
        A.n_items = V.n_items;
        A.partial = ctx->gate_partial[k & 1].as<fr_t>();
        if (k == 0) ZK_KLAUNCH_PDL(ctx, ZK_PROF_GATES, S.n_recs * 76 + (uint64_t) V.n_items * 44, k_eval_items, dim3(grid_for(V.n_items)), dim3(kBlock), 0, A);
        else ZK_KLAUNCH_PDL(ctx, ZK_PROF_GATES, (uint64_t) V.n_items * 44 + (uint64_t) S.levels[k - 1].n_partials * 32, k_sum_partials, dim3(grid_for(V.n_items)), dim3(kBlock), 0, A,
                            (const fr_t *) ctx->gate_partial[(k - 1) & 1].as<fr_t>());
    }
    if (!(L.scale == fr_t::one())) ZK_KLAUNCH(ctx, k_scale_vec, dim3(grid_for(L.d.size)), dim3(kBlock), 0, out, (uint64_t) L.d.size, L.scale);
}

static void eval_fft_layer(zk_ctx *ctx, uint32_t id) {
    layer_t &L = ctx->layers[id];
    layer_t &prev = ctx->layers[id - 1];
    const bool inverse = L.d.ty == ZK_LAYER_IFFT;
    const uint32_t n = L.d.fft_bit_length, len = 1u << n, half = len >> 1;
    const uint32_t n_blocks = inverse ? L.d.size / half : L.d.size / len;
    ZK_REQUIRE(n >= 1 && n <= 12, "unsupported FFT size for device witness generation");
    ZK_REQUIRE((uint64_t) n_blocks * (inverse ? half : len) == L.d.size && (uint64_t) n_blocks * (inverse ? len : half) <= prev.n_val, "FFT layer sizes are not whole blocks");
    const fr_t *pw = phi_powers(ctx, n, inverse);
    const fr_t ilen = fr_t::from_u64(len).inverse();
    const size_t smem = (size_t) len * sizeof(fr_t);
#ifndef ZK_EMU
    static const bool attr_set = [] {
        rt::check(cudaFuncSetAttribute(k_ntt_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (sizeof(fr_t) << 12)), "cudaFuncSetAttribute");
        return true;
    }();
    (void) attr_set;
#endif
    ZK_KLAUNCH_C(ctx, ZK_PROF_DENSE, (uint64_t) n_blocks * (len + half) * 32, k_ntt_blocks, dim3(std::min<uint32_t>(n_blocks, kMaxGridX * 4)), dim3(kBlock), smem, L.val.as<fr_t>(),
                 (const fr_t *) prev.val.as<fr_t>(), pw, n_blocks, n, inverse ? 1u : 0u, ilen);
}

static void eval_dotprod_layer(zk_ctx *ctx, uint32_t id) {
    layer_t &L = ctx->layers[id];
    layer_t &prev = ctx->layers[id - 1];
    const uint64_t total = (uint64_t) L.dpe_rows << L.d.fft_bit_length;
    ZK_REQUIRE(total == L.d.size, "DOT_PROD evaluation schedule missing");
    ZK_KLAUNCH_PDL(ctx, ZK_PROF_DENSE, (uint64_t) L.d.n_bin * (64ull << L.d.fft_bit_length) + total * 32, k_dotprod_eval, dim3(grid_for(total)), dim3(kBlock), 0, L.val.as<fr_t>(),
                   (const fr_t *) prev.val.as<fr_t>(), (const uint32_t *) L.dpe_rowptr.as<uint32_t>(), (const dp_eval_t *) L.dpe_gates.as<dp_eval_t>(), L.dpe_rows, (uint32_t) L.d.fft_bit_length);
}

static void run_aux_ops(zk_ctx *ctx, uint32_t id) {
    layer_t &L = ctx->layers[id];
    fr_t *val0 = ctx->layers[0].val.as<fr_t>();
    const fr_t *prev = ctx->layers[id - 1].val.as<fr_t>();
    if (L.n_aux_bits_prev)
        ZK_KLAUNCH_PDL(ctx, ZK_PROF_OTHER, 0, k_aux_bits, dim3(grid_for(L.n_aux_bits_prev)), dim3(kBlock), 0, val0, prev, (const aux_op_t *) L.aux_bits_prev.as<aux_op_t>(), L.n_aux_bits_prev);
    if (L.n_aux_max) {
        ctx->wit_scratch.ensure(std::max<size_t>(64, (size_t) L.aux_max_count * 8));
        unsigned long long *scr = ctx->wit_scratch.as<unsigned long long>();
        rt::dzero(scr, (size_t) L.aux_max_count * 8, ctx->stream);
        ZK_KLAUNCH(ctx, k_aux_max, dim3(grid_for(L.n_aux_max)), dim3(kBlock), 0, scr, L.aux_max_base, prev, (const aux_op_t *) L.aux_max.as<aux_op_t>(), L.n_aux_max);
        ZK_KLAUNCH_PDL(ctx, ZK_PROF_OTHER, 0, k_aux_max_store, dim3(grid_for(L.aux_max_count)), dim3(kBlock), 0, val0, L.aux_max_base, (const unsigned long long *) scr, L.aux_max_count);
    }
    if (L.n_aux_bits_l0)   // (after the maxima: they are what gets decomposed)
        ZK_KLAUNCH_PDL(ctx, ZK_PROF_OTHER, 0, k_aux_bits, dim3(grid_for(L.n_aux_bits_l0)), dim3(kBlock), 0, val0, (const fr_t *) val0, (const aux_op_t *) L.aux_bits_l0.as<aux_op_t>(),
                       L.n_aux_bits_l0);
}

}  // namespace zk

extern "C" {

// The auxiliary inputs that the construction of layer `layer_id` derives from earlier gate values (prepareSignBit / prepareDecmpBit /
// prepareMax of the reference, src/neuralNetwork.cpp:899-916), in the order they were issued.  Each op is {src, dst, meta}: dst indexes
// val[0]; meta bits 0-7 = bit position, bits 8-9 = 0 sign bit, 1 magnitude bit, 2 running maximum of max(0, value); bit 10 = the source is
// val[0][src] (the maxima themselves), else val[layer_id - 1][src].
int zk_circuit_aux_ops(zk_ctx *ctx, uint32_t layer_id, const uint32_t *ops /* n x 3 */, uint64_t n) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && layer_id >= 1 && layer_id < ctx->n_layers && (ops || n == 0), "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    layer_t &L = ctx->layers[layer_id];
    std::vector<aux_op_t> a, m, b;
    uint32_t lo = 0xffffffffu, hi = 0;
    for (uint64_t i = 0; i < n; ++i) {
        aux_op_t op = {ops[3 * i], ops[3 * i + 1], ops[3 * i + 2] & 0x3ffu};
        const uint32_t kind = (op.meta >> 8) & 3u;
        const bool from_l0 = (ops[3 * i + 2] >> 10) & 1u;
        ZK_REQUIRE(kind <= kAuxMax && !(from_l0 && kind != kAuxBit), "bad auxiliary op");
        if (kind == kAuxMax) { m.push_back(op); lo = std::min(lo, op.dst); hi = std::max(hi, op.dst); }
        else (from_l0 ? b : a).push_back(op);
    }
    auto up = [&](rt::dbuf &d, const std::vector<aux_op_t> &v) {
        if (v.empty()) return;
        d.ensure(v.size() * sizeof(aux_op_t));
        rt::h2d(d.p, v.data(), v.size() * sizeof(aux_op_t), ctx->stream);
    };
    up(L.aux_bits_prev, a); up(L.aux_max, m); up(L.aux_bits_l0, b);
    rt::sync(ctx->stream);
    L.n_aux_bits_prev = a.size(); L.n_aux_max = m.size(); L.n_aux_bits_l0 = b.size();
    L.aux_max_base = m.empty() ? 0 : lo;
    L.aux_max_count = m.empty() ? 0 : hi - lo + 1;
    L.aux_loaded = true;
    ZK_API_END
}

// Re-evaluates the whole circuit on the device for a new picture: val[0][0, n_image) <- image (field elements, as the host quantises
// them, src/neuralNetwork.cpp:805-833), then layer by layer the auxiliary inputs and the gate values.  The weights in val[0] are the
// resident ones (a complete witness must have been uploaded once).  ranges (may be NULL) receives, per layer, the largest non-negative
// value and the largest magnitude of a negative one (2 x n_layers values): what getNextBit (:967-977) derives the next quantisation scale
// from -- the caller checks that the circuit's structure (bit widths of the decompositions) still fits the new picture.
int zk_witness_generate_layers(zk_ctx *ctx, const uint64_t *image, uint64_t n_image, uint64_t *ranges, const uint8_t *want_range);
int zk_witness_generate(zk_ctx *ctx, const uint64_t *image, uint64_t n_image, uint64_t *ranges) {
    return zk_witness_generate_layers(ctx, image, n_image, ranges, nullptr);
}
// want_range (NULL: every layer): n_layers flags, ranges are only computed for the layers whose flag is set (the others read as zero)
int zk_witness_generate_layers(zk_ctx *ctx, const uint64_t *image, uint64_t n_image, uint64_t *ranges, const uint8_t *want_range) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && ctx->circuit_ready && image && n_image >= 1, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    layer_t &L0 = ctx->layers[0];
    ZK_REQUIRE(L0.n_val >= L0.d.size && n_image <= L0.d.size, "upload a complete witness once before generating on the device (the weights stay resident)");
    for (uint32_t i = 1; i < ctx->n_layers; ++i) {
        layer_t &L = ctx->layers[i];
        const int ty = L.d.ty;
        ZK_REQUIRE(ty == ZK_LAYER_FFT || ty == ZK_LAYER_IFFT || ty == ZK_LAYER_DOT_PROD ? true : (L.ev.n_recs == L.d.n_uni + L.d.n_bin), "evaluation schedules were not built");
        ZK_REQUIRE(L.aux_loaded, "zk_circuit_aux_ops was not called for every layer");
    }
    h2d_staged(ctx, L0.val.p, image, n_image * sizeof(fr_t));
    ctx->wit_scratch.ensure(64);
    if (ranges) {
        ctx->vres_scratch.ensure((size_t) ctx->n_layers * 16);
        rt::dzero(ctx->vres_scratch.p, (size_t) ctx->n_layers * 16, ctx->stream);
    }
    for (uint32_t i = 1; i < ctx->n_layers; ++i) {
        layer_t &L = ctx->layers[i];
        L.val.ensure(std::max<uint64_t>(1, L.d.size) * sizeof(fr_t));
        L.n_val = L.d.size;
        run_aux_ops(ctx, i);
        if (L.d.ty == ZK_LAYER_FFT || L.d.ty == ZK_LAYER_IFFT) eval_fft_layer(ctx, i);
        else if (L.d.ty == ZK_LAYER_DOT_PROD) eval_dotprod_layer(ctx, i);
        else eval_normal_layer(ctx, i);
        if (ranges && (!want_range || want_range[i]))
            ZK_KLAUNCH(ctx, k_layer_range, dim3(std::min<uint32_t>(grid_for(L.d.size), ZK_SM_COUNT * 4)), dim3(kBlock), 0, (const fr_t *) L.val.as<fr_t>(), (uint64_t) L.d.size,
                       ctx->vres_scratch.as<unsigned long long>() + 2 * i);
    }
    if (ranges) d2h_staged(ctx, ranges, ctx->vres_scratch.p, (size_t) ctx->n_layers * 16);
    else rt::sync(ctx->stream);
    for (auto &L : ctx->layers) L.next_ready = false;   // a prefetched copy of the previous witness is stale now
    ZK_API_END
}

// FNV-1a-64 over the canonical 32-byte little-endian encodings of val[layer_id][0, n) as it stands on the device (parity of the
// device-generated witness with the reference's values: the h_val column of tests/golden/*.circuit.txt)
int zk_debug_layer_hash(zk_ctx *ctx, uint32_t layer_id, uint64_t n, uint64_t *fnv1a) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && layer_id < ctx->n_layers && fnv1a && n <= ctx->layers[layer_id].n_val, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    std::vector<fr_t> h(n);
    if (n) d2h_staged(ctx, h.data(), ctx->layers[layer_id].val.p, n * sizeof(fr_t));
    uint64_t hash = 0xcbf29ce484222325ULL;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t c[8];
        h[i].to_canonical(c);
        const uint8_t *b = reinterpret_cast<const uint8_t *>(c);
        for (int k = 0; k < 32; ++k) { hash ^= b[k]; hash *= 0x100000001b3ULL; }
    }
    *fnv1a = hash;
    ZK_API_END
}

// the values of a (short) layer, e.g. the output layer for the inferred class
int zk_witness_read(zk_ctx *ctx, uint32_t layer_id, uint64_t first, uint64_t n, uint64_t *out) {
    ZK_API_BEGIN
    using namespace zk;
    ZK_REQUIRE(ctx && layer_id < ctx->n_layers && out && first + n <= ctx->layers[layer_id].n_val, "bad arguments");
    rt::bind(ctx->device, ctx->stream, ctx->aux_stream, ctx->copy_stream);
    if (n) d2h_staged(ctx, out, ctx->layers[layer_id].val.as<fr_t>() + first, n * sizeof(fr_t));
    ZK_API_END
}

}  // extern "C"
