// C ABI, part 1: circuit upload and the GKR prover state machine (mirror of src/prover.cpp of the reference).
// Included by capi.cu only.
#pragma once
#include "ctx.hpp"
#include <algorithm>
#include <atomic>
#include <cstring>
#include <ctime>

namespace zk {

static thread_local std::string g_last_error;

// ZK_TRACE=1: wall time per C-ABI entry point, printed at exit (where does the host-side latency go?)
struct api_trace_t {
    struct row { const char *name; uint64_t calls; double us; };
    std::vector<row> rows;
    bool on = getenv("ZK_TRACE") != nullptr;
    ~api_trace_t() {
        if (!on) return;
        std::sort(rows.begin(), rows.end(), [](const row &a, const row &b) { return a.us > b.us; });
        for (auto &r : rows) fprintf(stderr, "[zk_trace] %-36s calls %8lu total %10.3f ms  avg %9.2f us\n", r.name, (unsigned long) r.calls, r.us / 1e3, r.us / r.calls);
    }
    void add(const char *name, double us) {
        for (auto &r : rows) if (r.name == name) { ++r.calls; r.us += us; return; }
        rows.push_back({name, 1, us});
    }
};
static api_trace_t g_api_trace;
struct api_scope_t {
    const char *name;
    timespec t0;
    explicit api_scope_t(const char *n) : name(n) { if (g_api_trace.on) clock_gettime(CLOCK_MONOTONIC, &t0); }
    ~api_scope_t() {
        if (!g_api_trace.on) return;
        timespec t1;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        g_api_trace.add(name, (t1.tv_sec - t0.tv_sec) * 1e6 + (t1.tv_nsec - t0.tv_nsec) * 1e-3);
    }
};

#define ZK_API_BEGIN zk::api_scope_t zk_api_scope_(__func__); try {
#define ZK_API_END                                            \
    return 0;                                                 \
    } catch (const std::exception &e) {                       \
        zk::g_last_error = e.what();                          \
        return -1;                                            \
    }
#define ZK_REQUIRE(cond, msg) do { if (!(cond)) throw zk::rt::error(msg); } while (0)

static inline fr_t fr_load(const uint64_t *p) { fr_t x; memcpy(x.v, p, 32); return x; }
static inline void fr_store(uint64_t *p, const fr_t &x) { memcpy(p, x.v, 32); }
static inline uint32_t grid_for(uint64_t work_items) {
    uint64_t b = (work_items + kBlock - 1) / kBlock;
    if (b < 1) b = 1;
    if (b > (uint64_t) kMaxGridX) b = kMaxGridX;
    return (uint32_t) b;
}
static inline zk::rt::event_t prof_event(zk_ctx *ctx) {
    if (!ctx->prof_pool.empty()) { auto e = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); return e; }
    return zk::rt::event_create();
}
// launch of class `cls` (ZK_PROF_*) moving `bytes` algorithmic bytes
#define ZK_KLAUNCH_C(ctx, cls, bytes, kernel, grid, block, smem, ...) \
    ZK_KLAUNCH_S(ctx, (ctx)->stream, cls, bytes, kernel, grid, block, smem, __VA_ARGS__)
// the same on an explicit stream (side work that overlaps the main stream, e.g. the window-table build)
#define ZK_KLAUNCH_S(ctx, strm, cls, bytes, kernel, grid, block, smem, ...)               \
    do {                                                                                  \
        zk_ctx::prof_rec zk_pr_{(cls), nullptr, nullptr, #kernel};                        \
        if ((ctx)->prof_on) {                                                             \
            zk_pr_.a = zk::prof_event(ctx);                                               \
            zk_pr_.b = zk::prof_event(ctx);                                               \
            zk::rt::event_record(zk_pr_.a, (strm));                                       \
        }                                                                                 \
        ZK_LAUNCH(kernel, grid, block, smem, (strm), __VA_ARGS__);                        \
        ++(ctx)->launches;                                                                \
        zk::rt::check_launch(#kernel);                                                    \
        if ((ctx)->prof_on) {                                                             \
            zk::rt::event_record(zk_pr_.b, (strm));                                       \
            (ctx)->prof_pending.push_back(zk_pr_);                                        \
            ++(ctx)->prof_launches[(cls)];                                                \
            (ctx)->prof_bytes[(cls)] += (uint64_t) (bytes);                               \
        }                                                                                 \
    } while (0)
#define ZK_KLAUNCH(ctx, kernel, grid, block, smem, ...) ZK_KLAUNCH_C(ctx, ZK_PROF_OTHER, 0, kernel, grid, block, smem, __VA_ARGS__)
// for kernels that start with ZK_PDL_ENTRY() (sc_kernels.cuh): programmatic dependent launch on the main stream; with
// per-launch profiling events in the stream the overlap cannot happen, so the plain launch is used there
#ifndef ZK_EMU
#define ZK_KLAUNCH_PDL(ctx, cls, bytes, kernel, grid, block, smem, ...)                                       \
    do {                                                                                                      \
        if (!(ctx)->prof_on && (ctx)->pdl_enabled) {                                                          \
            cudaLaunchConfig_t zk_cfg_ = {};                                                                  \
            zk_cfg_.gridDim = (grid);                                                                         \
            zk_cfg_.blockDim = (block);                                                                       \
            zk_cfg_.dynamicSmemBytes = (smem);                                                                \
            zk_cfg_.stream = (ctx)->stream;                                                                   \
            cudaLaunchAttribute zk_attr_[1];                                                                  \
            zk_attr_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                              \
            zk_attr_[0].val.programmaticStreamSerializationAllowed = 1;                                       \
            zk_cfg_.attrs = zk_attr_;                                                                         \
            zk_cfg_.numAttrs = 1;                                                                             \
            zk::rt::check(cudaLaunchKernelEx(&zk_cfg_, kernel, __VA_ARGS__), #kernel);                        \
            ++(ctx)->launches;                                                                                \
        } else ZK_KLAUNCH_C(ctx, cls, bytes, kernel, grid, block, smem, __VA_ARGS__);                         \
    } while (0)
#else
#define ZK_KLAUNCH_PDL(ctx, cls, bytes, kernel, grid, block, smem, ...) ZK_KLAUNCH_C(ctx, cls, bytes, kernel, grid, block, smem, __VA_ARGS__)
#endif

// device counters of the MSM kernels under profiling: [0] mixed additions of k_msm_small, [1] bucket (mixed) additions of k_msm_window
static unsigned long long *prof_ops_counter(zk_ctx *ctx) {
    if (!ctx->prof_ops.p) {
        ctx->prof_ops.ensure(64);
        zk::rt::dzero(ctx->prof_ops.p, 64, ctx->stream);
    }
    return ctx->prof_ops.as<unsigned long long>();
}

static bool prof_by_kernel_enabled() { static const bool on = [] { const char *e = getenv("ZK_PROF_KERNELS"); return e && *e && *e != '0'; }(); return on; }
static void prof_print_kernels(zk_ctx *ctx) {
    if (ctx->prof_by_kernel.empty()) return;
    std::vector<std::pair<std::string, std::pair<uint64_t, double>>> v(ctx->prof_by_kernel.begin(), ctx->prof_by_kernel.end());
    std::sort(v.begin(), v.end(), [](const auto &a, const auto &b) { return a.second.second > b.second.second; });
    double tot = 0;
    for (auto &e : v) tot += e.second.second;
    fprintf(stderr, "[zk_prof] %-28s %8s %10s %6s %9s\n", "kernel", "launches", "ms", "share", "avg us");
    for (auto &e : v) fprintf(stderr, "[zk_prof] %-28s %8lu %10.3f %5.1f%% %9.2f\n", e.first.c_str(), (unsigned long) e.second.first, e.second.second,
                              100.0 * e.second.second / tot, 1e3 * e.second.second / e.second.first);
    fprintf(stderr, "[zk_prof] total %.3f ms (CUDA events around every launch on the launching stream)\n", tot);
    ctx->prof_by_kernel.clear();
}
static void prof_resolve(zk_ctx *ctx) {
    for (auto &r : ctx->prof_pending) {
        zk::rt::event_sync(r.b);
        const float ms = zk::rt::event_elapsed_ms(r.a, r.b);
        ctx->prof_ms[r.cls] += ms;
        if (prof_by_kernel_enabled() && r.name) { auto &e = ctx->prof_by_kernel[r.name]; ++e.first; e.second += ms; }
        ctx->prof_pool.push_back(r.a);
        ctx->prof_pool.push_back(r.b);
    }
    ctx->prof_pending.clear();
}

// --------------------------------------------------------------------------------------------------------------------
// schedule construction (host, once per circuit)
// --------------------------------------------------------------------------------------------------------------------
struct src_t { uint32_t rowkey; gate_rec_t rec; };  // rowkey: bits 30-31 = table (0, 1, 2 = scalar), bits 0-29 = row

static void build_schedule(zk_ctx *ctx, schedule_t &S, std::vector<src_t> &src, uint32_t rows0, uint32_t rows1, bool split_kind) {
    S.n_recs = src.size();
    S.n_val_recs = 0;
    if (!split_kind) for (auto &x : src) S.n_val_recs += ((x.rec.meta >> 16) & 3u) != 0;
    S.levels.clear();
    S.max_partials = 0;
    S.has_scalar = false;
    if (src.empty()) return;
    // stable counting sort by (table, row); callers emit kind-0 sources before kind-1 so kinds stay grouped per row
    const uint64_t nkeys = (uint64_t) rows0 + rows1 + 1;
    auto slot = [&](uint32_t rk) -> uint64_t {
        uint32_t t = rk >> 30, r = rk & 0x3fffffffu;
        return t == 0 ? r : t == 1 ? (uint64_t) rows0 + r : (uint64_t) rows0 + rows1;
    };
    std::vector<uint32_t> cnt(nkeys + 1, 0);
    for (auto &s : src) ++cnt[slot(s.rowkey) + 1];
    for (uint64_t i = 0; i < nkeys; ++i) cnt[i + 1] += cnt[i];
    std::vector<src_t> sorted(src.size());
    for (auto &s : src) sorted[cnt[slot(s.rowkey)]++] = s;
    std::vector<src_t>().swap(src);

    // level 0 items
    std::vector<gate_rec_t> recs;
    std::vector<item_t> items;
    std::vector<uint32_t> item_row;
    items.reserve(sorted.size() / kItemLen + 16);
    item_row.reserve(sorted.size() / kItemLen + 16);
    for (size_t i = 0; i < sorted.size();) {
        const uint32_t rk = sorted[i].rowkey;
        const uint32_t kind = (sorted[i].rec.meta >> 16) & 3u;
        size_t j = i;
        while (j < sorted.size() && j - i < (size_t) kItemLen && sorted[j].rowkey == rk &&
               (!split_kind || ((sorted[j].rec.meta >> 16) & 3u) == kind))
            ++j;
        item_t it;
        it.begin = (uint32_t) i;
        it.dest = 0;
        it.count_flags = (uint32_t) (j - i) | (split_kind ? (kind & 1u) << 16 : 0u);
        items.push_back(it);
        item_row.push_back(rk);
        i = j;
    }
    // Record layout: items are handed to threads in order (item i -> lane i mod 32 of a warp), so the records of every group of 32
    // consecutive items are stored TRANSPOSED: record k of item i sits at group_base + 32 k + (i mod 32).  The k-th record loads of a
    // warp then fall into one 384-byte run (12 wavefronts) instead of 32 runs 192 bytes apart (96): the gate kernels are bound by
    // load/store wavefronts, and the record loads were more than half of them.  Groups are padded to their longest item.
    {
        size_t total = 0;
        for (size_t g0 = 0; g0 < items.size(); g0 += kItemGroup) {
            uint32_t longest = 0;
            for (size_t i = g0; i < std::min(items.size(), g0 + kItemGroup); ++i) longest = std::max(longest, items[i].count_flags & 0xffffu);
            total += (size_t) longest * kItemGroup;
        }
        ZK_REQUIRE(total < (1ull << 32), "schedule too large for 32-bit record offsets");
        recs.assign(total, gate_rec_t{0u, 0u, 0u});
        size_t base = 0;
        for (size_t g0 = 0; g0 < items.size(); g0 += kItemGroup) {
            uint32_t longest = 0;
            for (size_t i = g0; i < std::min(items.size(), g0 + kItemGroup); ++i) {
                const uint32_t cnt = items[i].count_flags & 0xffffu, first = items[i].begin;
                longest = std::max(longest, cnt);
                for (uint32_t k = 0; k < cnt; ++k) recs[base + (size_t) k * kItemGroup + (i - g0)] = sorted[first + k].rec;
                items[i].begin = (uint32_t) (base + (i - g0));
            }
            base += (size_t) longest * kItemGroup;
        }
    }
    std::vector<src_t>().swap(sorted);
    S.recs.ensure(std::max<size_t>(1, recs.size()) * sizeof(gate_rec_t));
    rt::h2d(S.recs.p, recs.data(), recs.size() * sizeof(gate_rec_t), ctx->stream);
    rt::sync(ctx->stream);
    std::vector<gate_rec_t>().swap(recs);

    auto final_dest = [&](uint32_t rk) -> uint32_t {
        uint32_t t = rk >> 30, r = rk & 0x3fffffffu;
        if (t == 2) { S.has_scalar = true; return kDestFinal | kDestScalar; }
        ZK_REQUIRE(r < 0x20000000u, "table too large for the schedule encoding");
        return kDestFinal | (t == 1 ? kDestTable1 : 0u) | r;
    };
    // finish a level: rows with a single item go to the table, others get consecutive partial slots and a next level
    for (;;) {
        std::vector<item_t> next_items;
        std::vector<uint32_t> next_row;
        uint32_t n_partials = 0;
        for (size_t i = 0; i < items.size();) {
            size_t j = i;
            while (j < items.size() && item_row[j] == item_row[i]) ++j;
            if (j - i == 1) items[i].dest = final_dest(item_row[i]);
            else {
                const uint32_t first = n_partials;
                for (size_t k = i; k < j; ++k) items[k].dest = n_partials++;
                for (uint32_t b = first; b < n_partials; b += kItemLen) {
                    item_t it;
                    it.begin = b;
                    it.dest = 0;
                    it.count_flags = std::min<uint32_t>(kItemLen, n_partials - b);
                    next_items.push_back(it);
                    next_row.push_back(item_row[i]);
                }
            }
            i = j;
        }
        level_t L;
        L.n_items = (uint32_t) items.size();
        L.n_partials = n_partials;
        L.items.ensure(items.size() * sizeof(item_t));
        rt::h2d(L.items.p, items.data(), items.size() * sizeof(item_t), ctx->stream);
        rt::sync(ctx->stream);
        S.max_partials = std::max(S.max_partials, n_partials);
        S.levels.push_back(std::move(L));
        if (next_items.empty()) break;
        items.swap(next_items);
        item_row.swap(next_row);
    }
}

// evaluation order for the device witness generator: the gates of a layer sorted by OUTPUT gate (calcNormalLayer,
// src/neuralNetwork.cpp:918-932); layer-0 operands as absolute indices into val[0] (the gate arrays hold positions in ori_id_*)
static void build_eval_schedule(zk_ctx *ctx, uint32_t id, const zk_layer_desc *D) {
    layer_t &L = ctx->layers[id];
    if (D->ty == ZK_LAYER_DOT_PROD) {   // CSR by output block (calcDotProdLayer, :934-944)
        const uint32_t fft_bl = D->fft_bit_length;
        const uint32_t n_rows = D->size >> fft_bl;
        std::vector<uint32_t> ptr(n_rows + 1, 0);
        for (uint64_t i = 0; i < D->n_bin; ++i) {
            ZK_REQUIRE(D->bin_gates[i].g < n_rows, "DOT_PROD gate.g out of range");
            ++ptr[D->bin_gates[i].g + 1];
        }
        for (uint32_t i = 0; i < n_rows; ++i) ptr[i + 1] += ptr[i];
        std::vector<dp_eval_t> g(D->n_bin);
        std::vector<uint32_t> fill(ptr.begin(), ptr.end() - 1);
        for (uint64_t i = 0; i < D->n_bin; ++i) g[fill[D->bin_gates[i].g]++] = {D->bin_gates[i].u, D->bin_gates[i].v};
        L.dpe_rows = n_rows;
        L.dpe_rowptr.ensure(ptr.size() * 4);
        L.dpe_gates.ensure(std::max<size_t>(1, g.size()) * sizeof(dp_eval_t));
        rt::h2d(L.dpe_rowptr.p, ptr.data(), ptr.size() * 4, ctx->stream);
        rt::h2d(L.dpe_gates.p, g.data(), g.size() * sizeof(dp_eval_t), ctx->stream);
        rt::sync(ctx->stream);
        return;
    }
    std::vector<src_t> src;
    src.reserve(D->n_uni + D->n_bin);
    for (uint64_t i = 0; i < D->n_uni; ++i) {
        const zk_uni_gate &G = D->uni_gates[i];
        src_t s;
        s.rowkey = G.g;
        s.rec = {0u, G.lu != 0 ? G.u : D->ori_id_u[G.u], (uint32_t) G.sc | (G.lu != 0 ? kEvUPrev : 0u)};
        src.push_back(s);
    }
    for (uint64_t i = 0; i < D->n_bin; ++i) {
        const zk_bin_gate &G = D->bin_gates[i];
        const bool u_prev = G.l != 0, v_prev = (G.l & 1) != 0;
        src_t s;
        s.rowkey = G.g;
        s.rec = {v_prev ? G.v : D->ori_id_v[G.v], u_prev ? G.u : D->ori_id_u[G.u], (uint32_t) G.sc | kEvBin | (u_prev ? kEvUPrev : 0u) | (v_prev ? kEvVPrev : 0u)};
        src.push_back(s);
    }
    {   // how many output gates have a source at all
        std::vector<bool> seen(D->size, false);
        uint32_t covered = 0;
        for (auto &x : src)
            if (x.rowkey < D->size && !seen[x.rowkey]) { seen[x.rowkey] = true; ++covered; }
        L.ev_rows_covered = covered;
    }
    build_schedule(ctx, L.ev, src, 1u << D->bit_length, 0, false);
}

static void build_layer_schedules(zk_ctx *ctx, uint32_t id, const zk_layer_desc *D) {
    layer_t &L = ctx->layers[id];
    const int ty = D->ty;
    if (id == 0 || ty == ZK_LAYER_FFT || ty == ZK_LAYER_IFFT) return;
    if (ctx->eval_schedules) build_eval_schedule(ctx, id, D);
    auto v_abs = [&](uint32_t v) { return D->ori_id_v[v]; };
    const uint32_t rows_u0 = D->bit_length_u[0] >= 0 ? 1u << D->bit_length_u[0] : 0;
    const uint32_t rows_u1 = D->bit_length_u[1] >= 0 ? 1u << D->bit_length_u[1] : 0;
    const uint32_t rows_v0 = D->bit_length_v[0] >= 0 ? 1u << D->bit_length_v[0] : 0;
    const uint32_t rows_v1 = D->bit_length_v[1] >= 0 ? 1u << D->bit_length_v[1] : 0;

    if (ty == ZK_LAYER_DOT_PROD) {
        // phase 1: CSR by u of (g, v)  (src/prover.cpp:86-91)
        const uint32_t fft_bl = D->fft_bit_length;
        const uint32_t n_rows = rows_u1 >> fft_bl;
        std::vector<uint32_t> ptr(n_rows + 1, 0);
        uint32_t rows_live = 0;
        for (uint64_t i = 0; i < D->n_bin; ++i) {
            ZK_REQUIRE(D->bin_gates[i].u < n_rows, "DOT_PROD gate.u out of range");
            ++ptr[D->bin_gates[i].u + 1];
            rows_live = std::max(rows_live, D->bin_gates[i].u + 1);
        }
        L.dp_rows_live = rows_live;
        for (uint32_t i = 0; i < n_rows; ++i) ptr[i + 1] += ptr[i];
        std::vector<dp_gate_t> g(D->n_bin);
        std::vector<uint32_t> fill(ptr.begin(), ptr.end() - 1);
        for (uint64_t i = 0; i < D->n_bin; ++i) g[fill[D->bin_gates[i].u]++] = {D->bin_gates[i].g, D->bin_gates[i].v};
        L.dp_rows = n_rows;
        L.dp_rowptr.ensure(ptr.size() * 4);
        L.dp_gates.ensure(std::max<size_t>(1, g.size()) * sizeof(dp_gate_t));
        rt::h2d(L.dp_rowptr.p, ptr.data(), ptr.size() * 4, ctx->stream);
        rt::h2d(L.dp_gates.p, g.data(), g.size() * sizeof(dp_gate_t), ctx->stream);
        rt::sync(ctx->stream);
        // phase 2: mult[1][v] += beta_g[g] beta_u[u] V_u1  (src/prover.cpp:286-288)
        std::vector<src_t> src(D->n_bin);
        for (uint64_t i = 0; i < D->n_bin; ++i) {
            const zk_bin_gate &G = D->bin_gates[i];
            src[i].rowkey = (1u << 30) | G.v;
            src[i].rec = {G.g, G.u, (1u << 16)};
        }
        build_schedule(ctx, L.p2, src, rows_v0, rows_v1, true);
        return;
    }

    {   // phase 1 (src/prover.cpp:224-233)
        std::vector<src_t> src;
        src.reserve(D->n_uni + D->n_bin);
        for (uint64_t i = 0; i < D->n_uni; ++i) {
            const zk_uni_gate &G = D->uni_gates[i];
            src_t s;
            s.rowkey = ((G.lu != 0 ? 1u : 0u) << 30) | G.u;
            s.rec = {G.g, 0u, (uint32_t) G.sc};
            src.push_back(s);
        }
        for (uint64_t i = 0; i < D->n_bin; ++i) {
            const zk_bin_gate &G = D->bin_gates[i];
            const bool u_prev = G.l != 0, v_prev = (G.l & 1) != 0;   // binGate::getLayerIdU/V, src/circuit.h:31-32
            src_t s;
            s.rowkey = ((u_prev ? 1u : 0u) << 30) | G.u;
            s.rec = {G.g, v_prev ? G.v : v_abs(G.v), (uint32_t) G.sc | ((v_prev ? 2u : 1u) << 16)};
            src.push_back(s);
        }
        build_schedule(ctx, L.p1, src, rows_u0, rows_u1, false);
    }
    if (D->need_phase2) {   // phase 2 (src/prover.cpp:297-305)
        std::vector<src_t> src;
        src.reserve(D->n_uni + D->n_bin);
        for (int kind = 0; kind < 2; ++kind) {
            for (uint64_t i = 0; i < D->n_uni; ++i) {
                const zk_uni_gate &G = D->uni_gates[i];
                if ((G.lu != 0) != (kind != 0)) continue;
                src_t s;
                s.rowkey = 2u << 30;
                s.rec = {G.g, G.u, (uint32_t) G.sc | ((uint32_t) kind << 16)};
                src.push_back(s);
            }
            for (uint64_t i = 0; i < D->n_bin; ++i) {
                const zk_bin_gate &G = D->bin_gates[i];
                if ((G.l != 0) != (kind != 0)) continue;
                src_t s;
                s.rowkey = (((G.l & 1) ? 1u : 0u) << 30) | G.v;
                s.rec = {G.g, G.u, (uint32_t) G.sc | ((uint32_t) kind << 16)};
                src.push_back(s);
            }
        }
        build_schedule(ctx, L.p2, src, rows_v0, rows_v1, true);
    }
}

// --------------------------------------------------------------------------------------------------------------------
// helpers used by the Init* calls
// --------------------------------------------------------------------------------------------------------------------
struct beta_point_t { const fr_t *r; fr_t init; };   // host challenges, multiplier

// uploads up to two challenge vectors and builds their half tables; returns pointers for k_beta_expand / k_liu_scatter
struct halves_t { const fr_t *f[2] = {nullptr, nullptr}, *s[2] = {nullptr, nullptr}; uint32_t first_half = 0; };
static halves_t build_halves(zk_ctx *ctx, uint32_t bits, const beta_point_t *pts, int n_pts) {
    ZK_REQUIRE(bits <= 30, "beta table too large");
    halves_t H;
    const uint32_t fh = bits >> 1, sh = bits - fh;
    H.first_half = fh;
    ctx->d_r.ensure(2 * 64 * sizeof(fr_t));
    half_args_t A;
    memset(&A, 0, sizeof A);
    for (int k = 0; k < n_pts; ++k) {
        if (pts[k].init.is_zero()) continue;   // initBetaTable leaves zeros for a zero multiplier (src/utils.cpp:154-159)
        ctx->half[2 * k].ensure(sizeof(fr_t) << fh);
        ctx->half[2 * k + 1].ensure(sizeof(fr_t) << sh);
        fr_t *dr = ctx->d_r.as<fr_t>() + 64 * k;
        rt::h2d(dr, pts[k].r, bits * sizeof(fr_t), ctx->stream);
        A.job[2 * k] = {ctx->half[2 * k].as<fr_t>(), dr, fh, pts[k].init};
        A.job[2 * k + 1] = {ctx->half[2 * k + 1].as<fr_t>(), dr + fh, sh, fr_t::one()};
        H.f[k] = ctx->half[2 * k].as<fr_t>();
        H.s[k] = ctx->half[2 * k + 1].as<fr_t>();
    }
    ZK_KLAUNCH_PDL(ctx, ZK_PROF_TABLES, 0, k_half_tables, dim3(4), dim3(kBlock), 0, A);
    return H;
}

static void build_beta(zk_ctx *ctx, fr_t *out, uint32_t bits, const beta_point_t *pts, int n_pts, uint32_t tail_start = 0xffffffffu,
                       const fr_t &tail_scale = fr_t::one()) {
    halves_t H = build_halves(ctx, bits, pts, n_pts);
    beta_args_t B;
    B.out = out;
    // compact: the first present point goes to slot 0
    int k0 = H.f[0] ? 0 : 1;
    B.f0 = H.f[k0]; B.s0 = H.s[k0];
    B.f1 = (k0 == 0) ? H.f[1] : nullptr; B.s1 = (k0 == 0) ? H.s[1] : nullptr;
    B.bits = bits; B.first_half = H.first_half;
    B.tail_start = tail_start; B.tail_scale = tail_scale;
    ZK_KLAUNCH_PDL(ctx, ZK_PROF_TABLES, 32ull << bits, k_beta_expand, dim3(grid_for(1ull << bits)), dim3(kBlock), 0, B);
}

static void run_schedule(zk_ctx *ctx, const schedule_t &S, int phase, gate_args_t A) {
    if (S.levels.empty()) return;
    ctx->gate_partial[0].ensure((size_t) std::max(1u, S.max_partials) * sizeof(fr_t));
    ctx->gate_partial[1].ensure((size_t) std::max(1u, S.max_partials) * sizeof(fr_t));
    A.recs = S.recs.as<gate_rec_t>();
    for (size_t k = 0; k < S.levels.size(); ++k) {
        const level_t &L = S.levels[k];
        A.items = L.items.as<item_t>();
        A.n_items = L.n_items;
        A.partial = ctx->gate_partial[k & 1].as<fr_t>();
        if (k == 0) {
            if (phase == 1) ZK_KLAUNCH_PDL(ctx, ZK_PROF_GATES, S.n_recs * 44 + S.n_val_recs * 32 + (uint64_t) L.n_items * 44, k_gate_items_p1, dim3(grid_for(L.n_items)), dim3(kBlock), 0, A);
            else ZK_KLAUNCH_PDL(ctx, ZK_PROF_GATES, S.n_recs * 76 + (uint64_t) L.n_items * 44, k_gate_items_p2, dim3(grid_for(L.n_items)), dim3(kBlock), 0, A);
        } else {
            const fr_t *src = ctx->gate_partial[(k - 1) & 1].as<fr_t>();
            ZK_KLAUNCH_PDL(ctx, ZK_PROF_GATES, (uint64_t) L.n_items * 44 + (uint64_t) S.levels[k - 1].n_partials * 32, k_sum_partials, dim3(grid_for(L.n_items)), dim3(kBlock), 0, A, src);
        }
    }
}

// out[u] = sum_{g < n_g} val[(g << shift) | u] * weight[g] for u < 2^shift  (k_dense_colsum + k_colsum_finish)
static void dense_colsum(zk_ctx *ctx, const fr_t *val, const fr_t *weight, uint32_t shift, uint64_t n_g, fr_t *out) {
    const uint32_t n_u = 1u << shift;
    const uint64_t total = n_g << shift;
    // threads: a multiple of the column count (a thread keeps its column), at most 8 CTAs per SM, at least ~8 elements each
    uint64_t T = std::max<uint64_t>(std::max<uint32_t>(n_u, kBlock), std::min<uint64_t>((uint64_t) ZK_SM_COUNT * 8 * kBlock, (total / 8 + n_u - 1) / n_u * n_u));
    T = (T + std::max<uint32_t>(n_u, kBlock) - 1) / std::max<uint32_t>(n_u, kBlock) * std::max<uint32_t>(n_u, kBlock);
    const uint32_t per_u = (uint32_t) (T / n_u);
    ctx->dense_partial.ensure(T * sizeof(fr_t));
    ZK_KLAUNCH_PDL(ctx, ZK_PROF_DENSE, total * 32 + n_g * 32, k_dense_colsum, dim3((uint32_t) (T / kBlock)), dim3(kBlock), 0, val, weight, shift, total, ctx->dense_partial.as<fr_t>());
    const uint32_t groups = (n_u + kBlock / 32 - 1) / (kBlock / 32);
    ZK_KLAUNCH_PDL(ctx, ZK_PROF_DENSE, T * 32 + (uint64_t) n_u * 32, k_colsum_finish, dim3(std::min<uint32_t>(groups, kMaxGridX)), dim3(kBlock), 0, ctx->dense_partial.as<fr_t>(), n_u, per_u, out);
}

static void pair_reset(pair_t &P, int8_t bit_length, uint32_t size) {
    P.exists = bit_length >= 0;
    P.poly_round = 0;
    P.n_eval = P.exists ? 1u << bit_length : 0;
    P.live = size;
    P.collapsed = false;
    P.cv = P.cm = fr_t::zero();
    P.v.cur = P.m.cur = nullptr;
    P.v.next = P.m.next = 0;
}
static fr_t *table_init_buf(table_t &T, uint64_t entries) {
    T.init.ensure(std::max<uint64_t>(1, entries) * sizeof(fr_t));
    T.cur = T.init.as<fr_t>();
    return T.init.as<fr_t>();
}
static fr_t *table_fold_buf(table_t &T, uint64_t entries) {
    rt::dbuf &b = T.fold[T.next];
    b.ensure(std::max<uint64_t>(1, entries) * sizeof(fr_t));
    return b.as<fr_t>();
}
static void table_advance(table_t &T) {
    T.cur = T.fold[T.next].as<fr_t>();
    T.next ^= 1;
}

// Transfers between PAGEABLE caller memory and the device, staged through a page-locked buffer of the context: a pageable
// cudaMemcpyAsync is served in order with every other copy in flight, including the witness prefetch on the copy stream
// (10 ms), a copy from/to pinned memory is not.  Both helpers leave the stream synchronised.
constexpr size_t kStageBytes = 1u << 20;
static void *stage_buf(zk_ctx *ctx) {
    if (!ctx->stage_h) ctx->stage_h = rt::hmalloc_pinned(kStageBytes);
    return ctx->stage_h;
}
static void h2d_staged(zk_ctx *ctx, void *dst, const void *src, size_t n) {
    char *st = static_cast<char *>(stage_buf(ctx));
    for (size_t off = 0; off < n; off += kStageBytes) {
        const size_t len = std::min(kStageBytes, n - off);
        memcpy(st, static_cast<const char *>(src) + off, len);
        rt::h2d(static_cast<char *>(dst) + off, st, len, ctx->stream);
        rt::sync(ctx->stream);
    }
}
static void d2h_staged(zk_ctx *ctx, void *dst, const void *src, size_t n) {
    char *st = static_cast<char *>(stage_buf(ctx));
    for (size_t off = 0; off < n; off += kStageBytes) {
        const size_t len = std::min(kStageBytes, n - off);
        rt::d2h(st, static_cast<const char *>(src) + off, len, ctx->stream);
        rt::sync(ctx->stream);
        memcpy(static_cast<char *>(dst) + off, st, len);
    }
}

static void ensure_round_scratch(zk_ctx *ctx) {
    ctx->partials.ensure((size_t) 2 * kMaxGridX * 4 * sizeof(fr_t));
    if (!ctx->counters.p) {
        ctx->counters.ensure(8 * sizeof(uint32_t));
        rt::dzero(ctx->counters.p, 8 * sizeof(uint32_t), ctx->stream);
    }
    if (!ctx->round_acc.p) {   // grid-wide limb sums of k_round_quad: zero between launches (the kernel clears them itself)
        ctx->round_acc.ensure(128 * sizeof(unsigned long long));
        rt::dzero(ctx->round_acc.p, 128 * sizeof(unsigned long long), ctx->stream);
        ctx->round_state.ensure(8 * sizeof(fr_t));
        rt::dzero(ctx->round_state.p, 8 * sizeof(fr_t), ctx->stream);
    }
    ctx->round_out.ensure(16 * sizeof(fr_t));
    if (!ctx->h_out) ctx->h_out = static_cast<fr_t *>(rt::hmalloc_pinned(16 * sizeof(fr_t)));
    if (!ctx->res_h) {
        ctx->res_h = static_cast<fr_t *>(rt::hmalloc_mapped(32 * sizeof(fr_t) + 256));
        memset(ctx->res_h, 0, 32 * sizeof(fr_t) + 256);
        ctx->res_d = static_cast<fr_t *>(rt::mapped_device_ptr(ctx->res_h));
        ctx->flag_h = reinterpret_cast<uint32_t *>(ctx->res_h + 32);
        ctx->flag_d = reinterpret_cast<uint32_t *>(ctx->res_d + 32);
        ctx->tag_h = reinterpret_cast<uint32_t *>(ctx->res_h + 36);   // 128 bytes, 128-byte aligned: the tagged mailbox of k_round_quad
        ctx->tag_d = reinterpret_cast<uint32_t *>(ctx->res_d + 36);
    }
}
#ifndef ZK_EMU
// the input tables of a fold round as TMA tensor maps: (n_in / 4) rows of 32 x u32 (= four entries), box 32 x 32 rows,
// 128-byte swizzle.  cuTensorMapEncodeTiled comes from the driver through the runtime (no link-time dependency on libcuda).
static void encode_rows_map(CUtensorMap *tm, const fr_t *table, uint32_t n_in) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        rt::check(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q), "cudaGetDriverEntryPoint");
        if (!p || q != cudaDriverEntryPointSuccess) throw rt::error("cuTensorMapEncodeTiled is not available");
        return reinterpret_cast<encode_fn>(p);
    }();
    const cuuint64_t gdim[2] = {32, n_in / 4};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {32, 32}, estride[2] = {1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<fr_t *>(table), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw rt::error("cuTensorMapEncodeTiled failed (" + std::to_string((int) r) + ")");
}
static void launch_round_tma(zk_ctx *ctx, int cls, uint64_t bytes, round_args_t &A, const uint32_t limit_pairs[2]) {
    static const bool attr_set = [] {
        rt::check(cudaFuncSetAttribute(k_round_quad_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kTmaSmemBytes), "cudaFuncSetAttribute");
        return true;
    }();
    (void) attr_set;
    round_tma_args_t T;
    memset(&T, 0, sizeof T);
    // one wave: the resident CTAs are split between the pairs in proportion to their row blocks
    uint64_t groups[2], total = 0;
    for (int b = 0; b < 2; ++b) { groups[b] = A.pair[b].n_in ? (limit_pairs[b] + 31) / 32 : 0; total += groups[b]; }
    uint32_t gx = 0;
    for (int b = 0; b < 2; ++b) {
        if (!groups[b]) { A.pair[b].n_blocks = 0; continue; }
        const uint64_t want = (groups[b] + kRoundBlock / 32 - 1) / (kRoundBlock / 32);
        const uint64_t share = std::max<uint64_t>(1, (uint64_t) kTmaMaxGrid * groups[b] / total);
        A.pair[b].n_blocks = (uint32_t) std::min(want, share);
        gx += A.pair[b].n_blocks;
        if (A.pair[b].n_in >= 128) {
            encode_rows_map(&T.tm[b][0], A.pair[b].v_in, A.pair[b].n_in);
            encode_rows_map(&T.tm[b][1], A.pair[b].m_in, A.pair[b].n_in);
        }
    }
    T.R = A;
    ZK_KLAUNCH_C(ctx, cls, bytes, k_round_quad_tma, dim3(gx), dim3(kRoundBlock), kTmaSmemBytes, T);
}
#endif
// CTAs for one table pair of a sumcheck round: one output pair per thread while the machine has room, then grid-stride
static inline uint32_t round_grid_for(uint64_t live_pairs) {
    const uint64_t g = (live_pairs + kRoundBlock - 1) / kRoundBlock;
    return (uint32_t) std::max<uint64_t>(1, std::min<uint64_t>(g, kRoundMaxGrid));
}

// wait until the kernel that was given (flag_d, seq) has published its results into res_h
static void wait_mailbox(zk_ctx *ctx) {
#ifndef ZK_EMU
    volatile uint32_t *f = ctx->flag_h;
    uint64_t spins = 0;
    timespec t0{};
    while (*f != ctx->seq) {
        if (++spins == 200000) clock_gettime(CLOCK_MONOTONIC, &t0);   // ~ms without an answer: start watching the stream and the clock
        else if (spins > 200000 && (spins & 0xffff) == 0) {
            const cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) rt::check(q, "sumcheck round kernel");
            if (q == cudaSuccess && *f != ctx->seq) throw rt::error("sumcheck round kernel finished without publishing its result");
            timespec t1;
            clock_gettime(CLOCK_MONOTONIC, &t1);
            if (t1.tv_sec - t0.tv_sec > 20) throw rt::error("sumcheck round kernel did not publish its result within 20 s");
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
#endif
    if (*ctx->flag_h != ctx->seq) throw rt::error("result mailbox out of sequence");
}

// wait for the eight tagged words of publish_tagged() and unpack (a, b, c)
static void wait_tagged(zk_ctx *ctx, fr_t abc[3]) {
    volatile uint32_t *box = ctx->tag_h;
#ifndef ZK_EMU
    uint64_t spins = 0;
    timespec t0{};
    for (;;) {
        bool all = true;
        for (int j = 0; j < 8; ++j) all = all && box[4 * j + 3] == ctx->seq;
        if (all) break;
        if (++spins == 200000) clock_gettime(CLOCK_MONOTONIC, &t0);
        else if (spins > 200000 && (spins & 0xffff) == 0) {
            const cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) rt::check(q, "sumcheck round kernel");
            timespec t1;
            clock_gettime(CLOCK_MONOTONIC, &t1);
            if (t1.tv_sec - t0.tv_sec > 20) throw rt::error("sumcheck round kernel did not publish its result within 20 s");
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
#endif
    uint32_t w[24];
    for (int j = 0; j < 8; ++j) {
        if (box[4 * j + 3] != ctx->seq) throw rt::error("result mailbox out of sequence");
        w[3 * j] = box[4 * j]; w[3 * j + 1] = box[4 * j + 1]; w[3 * j + 2] = box[4 * j + 2];
    }
    for (int k = 0; k < 3; ++k) memcpy(abc[k].v, w + 8 * k, 32);
}

// k_round_quad_thin with programmatic dependent launch: its CTAs are placed (and its parameters fetched) while the previous
// kernel of the stream drains; the kernel itself waits (griddepcontrol.wait) before it touches memory.  With per-launch
// profiling events in the stream the overlap cannot happen, so the plain launch is used there.
static void launch_round_thin(zk_ctx *ctx, int cls, uint64_t bytes, uint32_t gx, const round_args_t &A) {
    ZK_KLAUNCH_PDL(ctx, cls, bytes, k_round_quad_thin, dim3(gx), dim3(kRoundBlock), 0, A);
}

// Which table pairs took part in a round and how: what the host needs to book the round's results.
struct round_rec_t {
    bool any_quad = false, any_final = false, first = false;
    bool quad[2] = {false, false}, fin[2] = {false, false};
};
// Launch the kernels of one round.  slot_d == nullptr: interactive round, the results go to the host mailbox and are
// awaited by the caller (round_quadratic).  Otherwise the results are left at slot_d (16 Fr: abc at 0, collapse values at
// 8) and nothing is awaited: the table sizes, grids and collapses of a round do not depend on its data, so a whole phase
// can be queued back to back (round_quadratic_batch).  The host-side table state is advanced either way.
static round_rec_t round_quadratic_launch(zk_ctx *ctx, const fr_t &prev, unsigned mask, fr_t *slot_d) {
    ensure_round_scratch(ctx);
    const bool first = ctx->round == 1;
    round_rec_t rec;
    rec.first = first;
    round_args_t A;
    memset(&A, 0, sizeof A);
    A.r = prev;
    A.acc = ctx->round_acc.as<unsigned long long>();
    A.counter = ctx->counters.as<uint32_t>();
    A.out = slot_d ? slot_d : ctx->res_d;
    final_fold_args_t F;
    memset(&F, 0, sizeof F);
    F.r = prev;
    F.out = (slot_d ? slot_d : ctx->res_d) + 8;
    bool &any_quad = rec.any_quad, &any_final = rec.any_final;
    bool (&quad)[2] = rec.quad, (&fin)[2] = rec.fin;
    uint32_t gx = 0, max_live_pairs = 0, pairs_of[2] = {0, 0};
    uint64_t fold_bytes = 0;   // algorithmic: read V and mult (live entries), write both halves
    for (int b = 0; b < 2; ++b) {
        pair_t &P = ctx->pair[b];
        if (!(mask & (1u << b)) || P.n_eval == 0) continue;
        const uint32_t n_after = first ? P.n_eval : P.n_eval >> 1;
        if (n_after == 1) {   // total[idx] == 1: evaluate and move the product into add_term (src/prover.cpp:400-404)
            F.v_in[b] = P.v.cur; F.m_in[b] = P.m.cur;
            F.live[b] = P.live; F.fold[b] = first ? 0 : 1; F.active[b] = 1;
            fin[b] = any_final = true;
        } else {
            round_pair_t &R = A.pair[b];
            R.v_in = P.v.cur; R.m_in = P.m.cur;
            R.n_in = P.n_eval; R.live = P.live; R.fold = first ? 0 : 1;
            if (!first) {
                R.v_out = table_fold_buf(P.v, n_after);
                R.m_out = table_fold_buf(P.m, n_after);
            }
            const uint32_t live_pairs = first ? (P.live + 1) >> 1 : (P.live + 3) >> 2;
            max_live_pairs = std::max(max_live_pairs, live_pairs);
            pairs_of[b] = live_pairs;
            quad[b] = any_quad = true;
            fold_bytes += (uint64_t) std::min(P.live, P.n_eval) * (first ? 64 : 96);
        }
    }
    // the last kernel of the round publishes the sequence number the host waits on
    if (!slot_d) {
        ++ctx->seq;
        if (any_final) { F.flag = ctx->flag_d; F.seq = ctx->seq; }
        else { A.tagged = reinterpret_cast<uint4 *>(ctx->tag_d); A.seq = ctx->seq; }
    }
    // tables up to 2^16 entries: four lanes per output pair (latency); beyond: one thread per output pair, grid-stride
    // (throughput), fed by TMA once the tables are large enough to stream from HBM
    const bool thin = max_live_pairs <= ctx->thin_max_pairs;
    uint32_t limit_pairs[2] = {0, 0};
    uint64_t max_n_in = 0;
    for (int b = 0; b < 2; ++b)
        if (quad[b]) {
            limit_pairs[b] = std::max(1u, std::min(pairs_of[b], first ? A.pair[b].n_in >> 1 : A.pair[b].n_in >> 2));
            A.pair[b].n_blocks = thin ? (limit_pairs[b] + kRoundBlock / 4 - 1) / (kRoundBlock / 4) : round_grid_for(limit_pairs[b]);
            gx += A.pair[b].n_blocks;
            max_n_in = std::max<uint64_t>(max_n_in, A.pair[b].n_in);
        }
    if (!thin) {
        // streaming kernels keep each pair's round polynomial on the device; a pair whose previous round went through one of
        // them gets its b coefficient from that polynomial instead of a third product (see round_args_t::derive_b)
        A.state = ctx->round_state.as<fr_t>();
        for (int b = 0; b < 2; ++b)
            if (quad[b]) {
                pair_t &P = ctx->pair[b];
                A.derive_b[b] = (!first && ctx->derive_b_enabled && P.poly_round + 1 == ctx->round) ? 1u : 0u;
                P.poly_round = ctx->round;
            }
    }
    // rounds that stream less than 32 MiB are bound by launch + reduction latency, not by HBM: they are accounted separately
    const int cls = fold_bytes >= (32u << 20) ? ZK_PROF_FOLD : ZK_PROF_FOLD_SMALL;
    if (any_quad && thin) launch_round_thin(ctx, cls, fold_bytes, gx, A);
#ifndef ZK_EMU
    else if (any_quad && !first && max_n_in >= ctx->tma_min_entries) launch_round_tma(ctx, cls, fold_bytes, A, limit_pairs);
#endif
    else if (any_quad) ZK_KLAUNCH_PDL(ctx, cls, fold_bytes, k_round_quad, dim3(gx), dim3(kRoundBlock), 0, A);
    if (any_final) ZK_KLAUNCH_PDL(ctx, ZK_PROF_OTHER, 0, k_final_fold, dim3(1), dim3(32), 0, F);
    // the table state follows from the sizes alone
    for (int b = 0; b < 2; ++b) {
        pair_t &P = ctx->pair[b];
        if (fin[b]) {
            P.collapsed = true;
            P.n_eval = 0;
        } else if (quad[b] && !first) {
            table_advance(P.v);
            table_advance(P.m);
            P.n_eval >>= 1;
            P.live = (P.live + 1) >> 1;
        }
    }
    return rec;
}
// book the results of a round (h_res: host copy of its 16-Fr result block): collapse values into add_term, (a, b, c) out
static void round_quadratic_book(zk_ctx *ctx, const round_rec_t &rec, const fr_t *h_res, fr_t abc[3]) {
    abc[0] = abc[1] = abc[2] = fr_t::zero();
    if (rec.any_quad)
        for (int k = 0; k < 3; ++k) abc[k] = h_res[k];   // already summed over both pairs by the kernel
    for (int b = 0; b < 2; ++b)
        if (rec.fin[b]) {
            pair_t &P = ctx->pair[b];
            P.cv = h_res[8 + 2 * b];
            P.cm = h_res[8 + 2 * b + 1];
            ctx->add_term = ctx->add_term + P.cv * P.cm;
        }
}
// One call of sumcheckUpdateEach for both table pairs (src/prover.cpp:396-426).  `mask` selects the pairs that take part
// (Liu: only pair 1).  Returns the sum of the pairs' round polynomials in abc[3]; add_term is updated for collapses.
static void round_quadratic(zk_ctx *ctx, const fr_t &prev, unsigned mask, fr_t abc[3]) {
    const round_rec_t rec = round_quadratic_launch(ctx, prev, mask, nullptr);
    if (rec.any_final) {
        wait_mailbox(ctx);
        round_quadratic_book(ctx, rec, ctx->res_h, abc);
    } else if (rec.any_quad) {
        fr_t got[16];
        wait_tagged(ctx, got);
        round_quadratic_book(ctx, rec, got, abc);
    } else round_quadratic_book(ctx, rec, ctx->res_h, abc);
}
// The remaining rounds of a phase in ONE launch (k_round_tail): same per-round results in the same slots as round_quadratic_launch
// would leave, the host-side table state advanced to what the round-by-round path would have reached.
static void round_tail_launch(zk_ctx *ctx, const fr_t *prevs, uint32_t n_rounds, unsigned mask, fr_t *slots_d, round_rec_t *recs) {
#ifndef ZK_EMU
    static const bool attr_set = [] {
        rt::check(cudaFuncSetAttribute(k_round_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kTailSmemBytes), "cudaFuncSetAttribute");
        return true;
    }();
    (void) attr_set;
#endif
    tail_args_t A;
    memset(&A, 0, sizeof A);
    A.first = ctx->round == 0 ? 1u : 0u;
    A.n_rounds = n_rounds;
    for (uint32_t j = 0; j < n_rounds; ++j) A.r[j] = prevs[j];
    A.slots = slots_d;
    uint64_t bytes = 0;
    for (int b = 0; b < 2; ++b) {
        pair_t &P = ctx->pair[b];
        if (!(mask & (1u << b)) || P.n_eval == 0) continue;
        A.v_in[b] = P.v.cur; A.m_in[b] = P.m.cur;
        A.n_in[b] = P.n_eval; A.live[b] = P.live;
        bytes += (uint64_t) std::min(P.live, P.n_eval) * 64;
    }
    // host-side state, round by round (sizes only: which pair folds, which collapses)
    for (uint32_t j = 0; j < n_rounds; ++j) {
        ++ctx->round;
        const bool first = ctx->round == 1;
        round_rec_t &rec = recs[j];
        rec = round_rec_t();
        rec.first = first;
        for (int b = 0; b < 2; ++b) {
            pair_t &P = ctx->pair[b];
            if (!(mask & (1u << b)) || P.n_eval == 0) continue;
            const uint32_t n_after = first ? P.n_eval : P.n_eval >> 1;
            if (n_after == 1) {
                rec.fin[b] = rec.any_final = true;
                P.collapsed = true;
                P.n_eval = 0;
            } else {
                rec.quad[b] = rec.any_quad = true;
                if (!first) { P.n_eval >>= 1; P.live = (P.live + 1) >> 1; }
            }
        }
    }
    for (int b = 0; b < 2; ++b) {
        pair_t &P = ctx->pair[b];
        if (!A.n_in[b]) continue;
        A.v_out[b] = table_fold_buf(P.v, 2);
        A.m_out[b] = table_fold_buf(P.m, 2);
        if (P.n_eval) {   // what is left for the Finalize call
            table_advance(P.v);
            table_advance(P.m);
            P.live = std::min(P.live, P.n_eval);
        }
    }
    ZK_KLAUNCH_PDL(ctx, ZK_PROF_FOLD_SMALL, bytes, k_round_tail, dim3(1), dim3(kTailBlock), kTailSmemBytes, A);
}

// A whole phase queued without waiting for the host in between: round j folds with prevs[j] (prevs[0] = 0).  Legitimate
// because the verifier draws every challenge of a phase BEFORE its first round (src/verifier.cpp:156-160, 207, 275-279):
// the prover messages are the same field elements as in the round-by-round protocol, in the same order.  `hook` books
// round j on the host (add_term scaling, collapse values) once all results are back.
template <class Hook> static void round_quadratic_batch(zk_ctx *ctx, const fr_t *prevs, uint32_t n_rounds, unsigned mask, Hook hook) {
    ensure_round_scratch(ctx);
    ctx->batch_res.ensure((size_t) n_rounds * 16 * sizeof(fr_t));
    if (ctx->batch_cap < n_rounds) {
        if (ctx->batch_h) rt::hfree_pinned(ctx->batch_h);
        ctx->batch_h = static_cast<fr_t *>(rt::hmalloc_pinned((size_t) n_rounds * 16 * sizeof(fr_t)));
        ctx->batch_cap = n_rounds;
    }
    std::vector<round_rec_t> recs(n_rounds);
    for (uint32_t j = 0; j < n_rounds; ++j) {
        // the tail of the phase in one launch once every table that takes part fits the CTA-resident kernel
        uint32_t biggest = 0;
        for (int b = 0; b < 2; ++b)
            if (mask & (1u << b)) biggest = std::max(biggest, ctx->pair[b].n_eval);
        if (ctx->tail_enabled && biggest <= ctx->tail_max_entries && n_rounds - j <= (uint32_t) kTailMaxRounds && n_rounds - j >= 2) {
            round_tail_launch(ctx, prevs + j, n_rounds - j, mask, ctx->batch_res.as<fr_t>() + (size_t) j * 16, recs.data() + j);
            break;
        }
        ++ctx->round;
        recs[j] = round_quadratic_launch(ctx, prevs[j], mask, ctx->batch_res.as<fr_t>() + (size_t) j * 16);
    }
    rt::d2h(ctx->batch_h, ctx->batch_res.p, (size_t) n_rounds * 16 * sizeof(fr_t), ctx->stream);
    rt::sync(ctx->stream);
    for (uint32_t j = 0; j < n_rounds; ++j) hook(j, recs[j], ctx->batch_h + (size_t) j * 16);
}

// value of V_mult[b][0] at the end of a phase (src/prover.cpp:462-463,476-477,490)
static void final_values(zk_ctx *ctx, const fr_t &prev, fr_t out[2]) {
    ensure_round_scratch(ctx);
    final_fold_args_t F;
    memset(&F, 0, sizeof F);
    F.r = prev;
    F.out = ctx->res_d + 8;
    bool any = false;
    for (int b = 0; b < 2; ++b) {
        pair_t &P = ctx->pair[b];
        out[b] = fr_t::zero();
        if (P.n_eval == 0) {
            if (P.exists && P.collapsed) out[b] = P.cv;
            continue;
        }
        ZK_REQUIRE(P.n_eval <= 2, "finalize called before the last round");
        F.v_in[b] = P.v.cur; F.m_in[b] = nullptr;
        F.live[b] = P.live; F.fold[b] = P.n_eval == 2; F.active[b] = 1;
        any = true;
    }
    if (any) {
        F.flag = ctx->flag_d;
        F.seq = ++ctx->seq;
        ZK_KLAUNCH_PDL(ctx, ZK_PROF_OTHER, 0, k_final_fold, dim3(1), dim3(32), 0, F);
        wait_mailbox(ctx);
        for (int b = 0; b < 2; ++b)
            if (F.active[b]) out[b] = ctx->res_h[8 + 2 * b];
    }
}

static const fr_t *phi_powers(zk_ctx *ctx, uint32_t n, bool is_ifft) {
    // phiPowInit (src/utils.cpp:53-59): powers of getRootOfUnit(n) (or of its inverse)
    const uint32_t key = n * 2 + (is_ifft ? 1 : 0);
    for (auto &e : ctx->phi_pw)
        if (e.first == key) return e.second.as<fr_t>();
    ZK_REQUIRE(n >= 1 && n <= 24, "unsupported FFT size");
    fr_t w;
    memcpy(w.v, ZK_C(fr_ROOT32), 32);
    for (uint32_t i = n; i < 32; ++i) w = w.sqr();
    if (is_ifft) w = w.inverse();
    std::vector<fr_t> pw((size_t) 1 << n);
    pw[0] = fr_t::one();
    for (size_t i = 1; i < pw.size(); ++i) pw[i] = pw[i - 1] * w;
    rt::dbuf d;
    d.ensure(pw.size() * sizeof(fr_t));
    rt::h2d(d.p, pw.data(), pw.size() * sizeof(fr_t), ctx->stream);
    rt::sync(ctx->stream);
    ctx->phi_pw.emplace_back(key, std::move(d));
    return ctx->phi_pw.back().second.as<fr_t>();
}

static layer_t &cur_layer(zk_ctx *ctx) {
    ZK_REQUIRE(ctx->circuit_ready, "circuit not uploaded");
    ZK_REQUIRE(ctx->sumcheck_id < ctx->n_layers, "bad sumcheck level");
    return ctx->layers[ctx->sumcheck_id];
}

}  // namespace zk
