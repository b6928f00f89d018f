// See prover.hpp.  Control flow and accounting mirror src/prover.cpp of the reference; the arithmetic is on the device.
#include "prover.hpp"
#include <algorithm>
#include <cstdlib>
#include <stdexcept>
#include <string>

static_assert(sizeof(F) == 32, "F must be 4 x 64-bit Montgomery limbs (mcl Fr layout)");
static_assert(sizeof(uniGate) == sizeof(zk_uni_gate), "uniGate layout (src/circuit.h:15-22)");
static_assert(sizeof(binGate) == sizeof(zk_bin_gate), "binGate layout (src/circuit.h:24-33)");

static inline const uint64_t *w(const F &x) { return reinterpret_cast<const uint64_t *>(&x); }
static inline uint64_t *w(F &x) { return reinterpret_cast<uint64_t *>(&x); }

#define ZK_F_BYTES 32   /* F_BYTE_SIZE == Fr::getByteSize() */

prover::prover() {}

prover::~prover() {
    joinPrefetch();
    unpinWitness();
    poly_p.reset();
    if (ctx_) zk_ctx_destroy(ctx_);
}

void prover::check(int rc, const char *what) const {
    if (rc != 0) throw std::runtime_error(std::string("zkcnn_b200: ") + what + ": " + zk_last_error());
}

void prover::uploadCircuit() {
    check(zk_circuit_begin(ctx_, C.size, w(C.two_mul[0]), (uint32_t) C.two_mul.size()), "zk_circuit_begin");
    for (u32 i = 0; i < C.size; ++i) {
        const layer &cur = C.circuit[i];
        zk_layer_desc d;
        memset(&d, 0, sizeof d);
        d.ty = (int32_t) cur.ty;
        d.size = cur.size;
        for (int b = 0; b < 2; ++b) {
            d.size_u[b] = cur.size_u[b]; d.size_v[b] = cur.size_v[b];
            d.bit_length_u[b] = cur.bit_length_u[b]; d.bit_length_v[b] = cur.bit_length_v[b];
        }
        d.bit_length = cur.bit_length;
        d.max_bl_u = cur.max_bl_u;
        d.max_bl_v = cur.max_bl_v;
        d.need_phase2 = cur.need_phase2;
        d.zero_start_id = cur.zero_start_id;
        d.fft_bit_length = cur.fft_bit_length;
        memcpy(d.scale, &cur.scale, 32);
        d.uni_gates = reinterpret_cast<const zk_uni_gate *>(cur.uni_gates.data());
        d.n_uni = cur.uni_gates.size();
        d.bin_gates = reinterpret_cast<const zk_bin_gate *>(cur.bin_gates.data());
        d.n_bin = cur.bin_gates.size();
        d.ori_id_u = cur.ori_id_u.data();
        d.ori_id_v = cur.ori_id_v.data();
        check(zk_circuit_layer(ctx_, i, &d), "zk_circuit_layer");
        if (i >= 1) {
            const bool have = aux_ops.size() == (size_t) C.size;
            check(zk_circuit_aux_ops(ctx_, i, have && !aux_ops[i].empty() ? aux_ops[i].data() : nullptr, have ? aux_ops[i].size() / 3 : 0), "zk_circuit_aux_ops");
        }
    }
    check(zk_circuit_end(ctx_), "zk_circuit_end");
    circuit_uploaded_ = true;
}

// entries of val[i] that carry data: commitInput() pads val[0] with zeros up to a power of two on the host
// (src/prover.cpp:504-508); the device pads its copy itself, so the padding (124 MB for vgg11) never crosses PCIe
size_t prover::witnessLength(u32 i) const {
    return i == 0 ? std::min<size_t>(val[0].size(), C.circuit[0].size) : val[i].size();
}

void prover::uploadWitness() {
    last_upload_bytes_ = 0;
    for (u32 i = 0; i < C.size; ++i) {
        const size_t n = witnessLength(i);
        if (compact_ready_) {
            check(zk_witness_layer_compact(ctx_, i, compact_[i].data(), n, wide_idx_[i].data(), wide_idx_[i].empty() ? nullptr : w(wide_val_[i][0]),
                                           (uint32_t) wide_idx_[i].size(), 0), "zk_witness_layer_compact");
            last_upload_bytes_ += compactBytes(i);
        } else {
            check(zk_witness_layer(ctx_, i, n ? w(val[i][0]) : nullptr, n), "zk_witness_layer");
            last_upload_bytes_ += n * sizeof(F);
        }
    }
    witness_uploaded_ = true;
}

void prover::pinWitness() {
    unpinWitness();
    if (getenv("ZKCNN_NO_PIN")) return;
    // commitInput() later pads val[0] to 2^bit_length in place (src/prover.cpp:504-508): make room now so that the
    // vector is not reallocated after it has been page-locked
    if (!val.empty() && C.size) val[0].reserve((size_t) 1 << C.circuit[0].bit_length);
#ifndef ZKCNN_DROPIN   // (with the reference's own mcl types the shim has no view of the limbs: plain 32-byte upload there)
    if (!getenv("ZKCNN_NO_COMPACT")) {
        // with the compact encoding only the int64 arrays cross PCIe: those are the buffers to page-lock
        buildCompactWitness();
        for (auto &c : compact_)
            if (!c.empty() && zk_host_pin(c.data(), c.capacity() * sizeof(int64_t)) == 0) pinned_.push_back(c.data());
        return;
    }
#endif
    for (auto &v : val)
        if (!v.empty() && zk_host_pin(v.data(), v.capacity() * sizeof(F)) == 0) pinned_.push_back(v.data());
}

// int64 image of the witness: |x| < 2^62 under mcl's sign convention (a value >= (r+1)/2 stands for x - r) -> compact_, anything
// else -> (index, value) in the wide lists.  Done once per witness (part of zkh_build), on a few host threads.
void prover::buildCompactWitness() {
#ifndef ZKCNN_DROPIN
    compact_.assign(C.size, {});
    wide_idx_.assign(C.size, {});
    wide_val_.assign(C.size, {});
    const unsigned hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    for (u32 l = 0; l < C.size; ++l) {
        const size_t n = witnessLength(l);
        compact_[l].assign(n, 0);
        std::vector<std::vector<uint32_t>> wide(hw);
        auto work = [&](unsigned t) {
            const size_t b = n * t / hw, e = n * (t + 1) / hw;
            for (size_t i = b; i < e; ++i) {
                uint32_t c[8];
                val[l][i].v.to_canonical(c);
                bool small = !(c[2] | c[3] | c[4] | c[5] | c[6] | c[7]) && c[1] < (1u << 30);
                int64_t x = (int64_t) ((uint64_t) c[0] | ((uint64_t) c[1] << 32));
                if (!small) {   // maybe negative: r - c
                    const uint32_t *p = zk::fr_cfg::mod();
                    uint32_t d[8];
                    int64_t bw = 0;
                    for (int k = 0; k < 8; ++k) { bw += (int64_t) p[k] - (int64_t) c[k]; d[k] = (uint32_t) bw; bw >>= 32; }
                    if (!(d[2] | d[3] | d[4] | d[5] | d[6] | d[7]) && d[1] < (1u << 30)) {
                        small = true;
                        x = -(int64_t) ((uint64_t) d[0] | ((uint64_t) d[1] << 32));
                    }
                }
                if (small) compact_[l][i] = x;
                else wide[t].push_back((uint32_t) i);
            }
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < hw; ++t) th.emplace_back(work, t);
        work(0);
        for (auto &x : th) x.join();
        for (auto &wv : wide)
            for (uint32_t i : wv) { wide_idx_[l].push_back(i); wide_val_[l].push_back(val[l][i]); }
    }
    compact_ready_ = true;
    // Digit width of the commitment's small-multiples path (zk_set_tunable "msm_digit_bits").  A scalar costs one point addition per
    // non-zero digit of its magnitude and the table 2^w - 1 entries per generator, rebuilt for every proof's generators: count both
    // for w = 6, 7, 8 in field multiplications (11 per addition, ~39 per table entry) over the input layer.
    if (C.size && !getenv("ZKH_MSM_DIGIT_BITS")) {
        const size_t n = witnessLength(0);
        uint64_t adds[3] = {0, 0, 0};
        for (size_t i = 0; i < n; ++i) {
            const int64_t x = compact_[0][i];
            const uint64_t m = (uint64_t) (x < 0 ? -x : x);
            if (m == 0 || m >> 24) continue;   // zero, or beyond the small path whatever the width
            for (int k = 0; k < 3; ++k) {
                const unsigned w = 6 + k;
                for (uint64_t t = m; t; t >>= w) adds[k] += (t & ((1u << w) - 1)) != 0;
            }
        }
        const uint64_t n_gens = 1ULL << ((C.circuit[0].bit_length + 1) / 2);
        uint64_t best = ~0ULL;
        for (int k = 0; k < 3; ++k) {
            const uint64_t cost = adds[k] * 11 + n_gens * ((1u << (6 + k)) - 1) * 39;
            if (cost < best) { best = cost; msm_digit_bits_ = 6 + k; }
        }
    } else if (const char *e = getenv("ZKH_MSM_DIGIT_BITS")) msm_digit_bits_ = (unsigned) atoi(e);
    if (getenv("ZKH_WITNESS_STATS")) {
        fprintf(stderr, "[witness] digit width of the commitment's small-multiples path: %u bits\n", msm_digit_bits_);   // diagnostic: how the input layer's scalars fall on the commitment MSM's paths (rows of 2^ceil(bl/2) generators)
        const size_t n = witnessLength(0);
        u32 bl = 0;
        while (((size_t) 1 << bl) < n) ++bl;
        const size_t row_len = (size_t) 1 << (bl - bl / 2);
        size_t zero = 0, byte1 = 0, by_bytes[9] = {}, full = wide_idx_[0].size();
        std::vector<uint8_t> row_wide((n + row_len - 1) / row_len, 0), item_wide((n + 127) / 128, 0);
        for (size_t i = 0; i < n; ++i) {
            const int64_t x = compact_[0][i];
            const uint64_t m = (uint64_t) (x < 0 ? -x : x);
            if (m == 0) { ++zero; continue; }
            if (m < 256) { ++byte1; continue; }
            int nb = 0;
            for (uint64_t t = m; t; t >>= 8) ++nb;
            ++by_bytes[nb];
            row_wide[i / row_len] = 1;
            item_wide[i / 128] = 1;
        }
        zero -= full;   // the wide-list entries are zero in the int64 image
        for (uint32_t i : wide_idx_[0]) { row_wide[i / row_len] = 1; item_wide[i / 128] = 1; }
        size_t rows = 0, items = 0;
        fprintf(stderr, "[witness] rows with a wide scalar:");
        for (size_t r = 0; r < row_wide.size(); ++r) if (row_wide[r]) fprintf(stderr, " %zu", r);
        fprintf(stderr, "\n");
        for (auto r : row_wide) rows += r;
        for (auto r : item_wide) items += r;
        size_t hist[9] = {};   // one-byte magnitudes by bit length
        for (size_t i = 0; i < n; ++i) {
            const int64_t x = compact_[0][i];
            const uint64_t m = (uint64_t) (x < 0 ? -x : x);
            if (m == 0 || m > 255) continue;
            int bl2 = 0;
            for (uint64_t t = m; t; t >>= 1) ++bl2;
            ++hist[bl2];
        }
        fprintf(stderr, "[witness] one-byte magnitudes by bit length 1..8:");
        for (int b = 1; b <= 8; ++b) fprintf(stderr, " %zu", hist[b]);
        fprintf(stderr, "\n");
        size_t ones = 0, bit_blocks = 0, bit_block_ones = 0, nonempty_bit_blocks = 0;
        for (size_t i = 0; i + 8 <= n; i += 8) {
            bool pure = true;
            size_t c = 0;
            for (size_t k = 0; k < 8; ++k) { const int64_t x = compact_[0][i + k]; pure = pure && (x == 0 || x == 1); c += x == 1; }
            if (pure) { ++bit_blocks; bit_block_ones += c; nonempty_bit_blocks += c != 0; }
        }
        for (size_t i = 0; i < n; ++i) ones += compact_[0][i] == 1;
        fprintf(stderr, "[witness] ones %zu; aligned blocks of 8 scalars that hold only 0/1: %zu of %zu (%zu non-empty, %zu ones inside)\n", ones, bit_blocks, n / 8,
                nonempty_bit_blocks, bit_block_ones);
        fprintf(stderr, "[witness] input layer: %zu scalars, rows of %zu; zero %zu, one byte %zu, wider:", n, row_len, zero, byte1);
        for (int b = 2; b <= 8; ++b) fprintf(stderr, " %dB %zu", b, by_bytes[b]);
        fprintf(stderr, ", beyond int64 %zu; rows with a wide scalar %zu of %zu, 128-generator chunks with one %zu\n", full, rows, row_wide.size(), items);
    }
#endif
}

void prover::unpinWitness() {
    joinPrefetch();   // a copy in flight reads these buffers
    // the host witness is about to change: neither the device copy nor a prefetched shadow copy describes it any more
    prefetch_pending_ = false;
    witness_uploaded_ = false;
    compact_ready_ = false;
    for (const void *p : pinned_) zk_host_unpin(p);
    pinned_.clear();
}

void prover::joinPrefetch() {
    if (prefetch_thread_.joinable()) prefetch_thread_.join();
}

void prover::prefetchWitness() {
    if (!ctx_ || !circuit_uploaded_) return;   // nothing to overlap with before the first proof
    joinPrefetch();
    prefetch_bytes_ = 0;
    prefetch_error_.clear();
    for (int i = 0; i < C.size; ++i) prefetch_bytes_ += compact_ready_ ? compactBytes(i) : witnessLength(i) * sizeof(F);
    prefetch_thread_ = std::thread([this] {
        for (int i = 0; i < C.size; ++i)
            if ((compact_ready_ ? zk_witness_layer_compact(ctx_, i, compact_[i].data(), witnessLength(i), wide_idx_[i].data(),
                                                           wide_idx_[i].empty() ? nullptr : w(wide_val_[i][0]), (uint32_t) wide_idx_[i].size(), 1)
                                : zk_witness_layer_prefetch(ctx_, i, witnessLength(i) ? w(val[i][0]) : nullptr, witnessLength(i))) != 0) {
                prefetch_error_ = zk_last_error();
                return;
            }
    });
    prefetch_pending_ = true;
}

void prover::init() {   // src/prover.cpp:17-21 (+ upload of what the reference reads in place)
    proof_size = 0;
    upload_timer.start();
    if (!ctx_) {
        int dev = device_;
        if (dev < 0) { const char *e = getenv("ZKCNN_DEVICE"); dev = e ? atoi(e) : 0; }
        ctx_ = zk_ctx_create(dev);
        if (!ctx_) throw std::runtime_error(std::string("zkcnn_b200: cannot create a device context: ") + zk_last_error());
    }
    if (!circuit_uploaded_) { uploadCircuit(); witness_uploaded_ = false; }
    joinPrefetch();
    if (prefetch_pending_ && !prefetch_error_.empty()) throw std::runtime_error("zkcnn_b200: witness prefetch: " + prefetch_error_);
    if (prefetch_pending_ && circuit_uploaded_) {
        check(zk_witness_commit_prefetch(ctx_), "zk_witness_commit_prefetch");
        last_upload_bytes_ = prefetch_bytes_;
        witness_uploaded_ = true;
    } else if (!(witness_resident_ && witness_uploaded_)) uploadWitness();
    else last_upload_bytes_ = 0;
    prefetch_pending_ = false;
    if (prefetch_next_) prefetchWitness();
    check(zk_prover_init(ctx_), "zk_prover_init");
    upload_timer.stop();
}

void prover::generateWitnessOnDevice(const vector<F> &image, std::vector<uint64_t> &ranges, const std::vector<uint8_t> *want_range) {
    upload_timer.start();
    joinPrefetch();
    prefetch_pending_ = false;
    if (!ctx_) {
        int dev = device_;
        if (dev < 0) { const char *e = getenv("ZKCNN_DEVICE"); dev = e ? atoi(e) : 0; }
        ctx_ = zk_ctx_create(dev);
        if (!ctx_) throw std::runtime_error(std::string("zkcnn_b200: cannot create a device context: ") + zk_last_error());
    }
    if (!circuit_uploaded_) { uploadCircuit(); witness_uploaded_ = false; }
    if (!witness_uploaded_) uploadWitness();   // once: the quantised weights stay resident in val[0]
    ranges.assign((size_t) 2 * C.size, 0);
    check(zk_witness_generate_layers(ctx_, w(image[0]), image.size(), ranges.data(), want_range && want_range->size() == (size_t) C.size ? want_range->data() : nullptr),
          "zk_witness_generate");
    last_upload_bytes_ = image.size() * sizeof(F);
    upload_timer.stop();
}

vector<F> prover::readLayer(u32 layer, size_t n) {
    vector<F> out(n);
    if (n) check(zk_witness_read(ctx_, layer, 0, n, w(out[0])), "zk_witness_read");
    return out;
}

void prover::sumcheckInitAll(const vector<F>::const_iterator &r_0_from_v) {   // src/prover.cpp:28-36
    u32 last_bl = C.circuit[C.size - 1].bit_length;
    prove_timer.start();
    check(zk_sumcheck_init_all(ctx_, last_bl ? w(r_0_from_v[0]) : nullptr, last_bl), "zk_sumcheck_init_all");
    prove_timer.stop();
    level_ = (int) C.size;
}

void prover::sumcheckInit(const F &alpha_0, const F &beta_0) {   // src/prover.cpp:43-52
    prove_timer.start();
    check(zk_sumcheck_init(ctx_, w(alpha_0), w(beta_0)), "zk_sumcheck_init");
    prove_timer.stop();
    --level_;   // sumcheck_id of the reference (src/prover.cpp:50)
}

void prover::dumpTables(int layer, const char *tag, bool dot, int first_b) {
    if (!table_dump_) return;
    for (int b = first_b; b < 2; ++b) {
        uint64_t hv = 0, hm = 0, n = 0, nm = 0;
        // DOT_PROD phase 1: the device keeps V_mult[1] as the V table and V_mult[0] as the mult table of pair 1, the 2^fft_bl multiplier apart
        const int sel_v = dot ? (b == 0 ? 3 : 2) : 2 * b, sel_m = dot ? 4 : 2 * b + 1;
        check(zk_debug_table_hash(ctx_, sel_v, &hv, &n), "zk_debug_table_hash");
        if (dot && b == 1) hm = 0xcbf29ce484222325ULL;
        else check(zk_debug_table_hash(ctx_, sel_m, &hm, &nm), "zk_debug_table_hash");
        fprintf(table_dump_, "T %d %s %d n %lu v %016lx m %016lx\n", layer, tag, b, (unsigned long) n, (unsigned long) hv, (unsigned long) hm);
    }
}

void prover::sumcheckDotProdInitPhase1() {   // src/prover.cpp:57-95
    prove_timer.start();
    check(zk_sumcheck_dotprod_init_phase1(ctx_), "zk_sumcheck_dotprod_init_phase1");
    prove_timer.stop();
    dumpTables(level_, "dp1", true, 0);
}

void prover::sumcheckInitPhase1(const F &relu_rou_0) {   // src/prover.cpp:155-239
    prove_timer.start();
    check(zk_sumcheck_init_phase1(ctx_, w(relu_rou_0)), "zk_sumcheck_init_phase1");
    prove_timer.stop();
    dumpTables(level_, "p1", false, 0);
}

void prover::sumcheckInitPhase2() {   // src/prover.cpp:241-310
    prove_timer.start();
    check(zk_sumcheck_init_phase2(ctx_), "zk_sumcheck_init_phase2");
    prove_timer.stop();
    dumpTables(level_, "p2", false, 0);
}

cubic_poly prover::sumcheckDotProdUpdate1(const F &previous_random) {   // src/prover.cpp:103-144
    prove_timer.start();
    F c[4];
    check(zk_sumcheck_dotprod_update1(ctx_, w(previous_random), w(c[0])), "zk_sumcheck_dotprod_update1");
    cubic_poly ret(c[0], c[1], c[2], c[3]);
    proof_size += ZK_F_BYTES * (3 + (!ret.a.isZero()));
    prove_timer.stop();
    if (transcript_) for (int k = 0; k < 4; ++k) transcript_->put_fr(w(c[k]));
    return ret;
}

quadratic_poly prover::sumcheckUpdate1(const F &previous_random) {   // src/prover.cpp:360-362
    prove_timer.start();
    F c[3];
    check(zk_sumcheck_update1(ctx_, w(previous_random), w(c[0])), "zk_sumcheck_update1");
    prove_timer.stop();
    proof_size += ZK_F_BYTES * 3;
    if (transcript_) for (int k = 0; k < 3; ++k) transcript_->put_fr(w(c[k]));
    return quadratic_poly(c[0], c[1], c[2]);
}

quadratic_poly prover::sumcheckUpdate2(const F &previous_random) {   // src/prover.cpp:364-366
    prove_timer.start();
    F c[3];
    check(zk_sumcheck_update2(ctx_, w(previous_random), w(c[0])), "zk_sumcheck_update2");
    prove_timer.stop();
    proof_size += ZK_F_BYTES * 3;
    if (transcript_) for (int k = 0; k < 3; ++k) transcript_->put_fr(w(c[k]));
    return quadratic_poly(c[0], c[1], c[2]);
}

F prover::Vres(const vector<F>::const_iterator &r, u32 output_size, u8 r_size) {   // src/prover.cpp:434-457
    prove_timer.start();
    F res;
    check(zk_vres(ctx_, r_size ? w(r[0]) : nullptr, output_size, r_size, w(res)), "zk_vres");
    prove_timer.stop();
    proof_size += ZK_F_BYTES;
    if (transcript_) transcript_->put_fr(w(res));
    return res;
}

void prover::sumcheckDotProdFinalize1(const F &previous_random, F &claim_1) {   // src/prover.cpp:146-153
    prove_timer.start();
    check(zk_sumcheck_dotprod_finalize1(ctx_, w(previous_random), w(claim_1)), "zk_sumcheck_dotprod_finalize1");
    prove_timer.stop();
    proof_size += ZK_F_BYTES * 1;
    if (transcript_) transcript_->put_fr(w(claim_1));
}

void prover::sumcheckFinalize1(const F &previous_random, F &claim_0, F &claim_1) {   // src/prover.cpp:459-471
    prove_timer.start();
    check(zk_sumcheck_finalize1(ctx_, w(previous_random), w(claim_0), w(claim_1)), "zk_sumcheck_finalize1");
    prove_timer.stop();
    proof_size += ZK_F_BYTES * 2;
    if (transcript_) { transcript_->put_fr(w(claim_0)); transcript_->put_fr(w(claim_1)); }
}

void prover::sumcheckFinalize2(const F &previous_random, F &claim_0, F &claim_1) {   // src/prover.cpp:473-485
    prove_timer.start();
    check(zk_sumcheck_finalize2(ctx_, w(previous_random), w(claim_0), w(claim_1)), "zk_sumcheck_finalize2");
    prove_timer.stop();
    proof_size += ZK_F_BYTES * 2;
    if (transcript_) { transcript_->put_fr(w(claim_0)); transcript_->put_fr(w(claim_1)); }
}

void prover::sumcheckLiuFinalize(const F &previous_random, F &claim_1) {   // src/prover.cpp:487-497
    prove_timer.start();
    check(zk_sumcheck_liu_finalize(ctx_, w(previous_random), w(claim_1)), "zk_sumcheck_liu_finalize");
    prove_timer.stop();
    proof_size += ZK_F_BYTES;
    if (transcript_) transcript_->put_fr(w(claim_1));
}

void prover::sumcheckLiuInit(const vector<F> &s_u, const vector<F> &s_v) {   // src/prover.cpp:312-358
    if (s_u.size() + 1 < C.size || s_v.size() + 1 < C.size) throw std::out_of_range("sumcheckLiuInit: sigma vectors too short");
    prove_timer.start();
    check(zk_sumcheck_liu_init(ctx_, w(s_u[0]), w(s_v[0]), (uint32_t) s_u.size()), "zk_sumcheck_liu_init");
    prove_timer.stop();
    dumpTables(0, "liu", false, 1);
}

quadratic_poly prover::sumcheckLiuUpdate(const F &previous_random) {   // src/prover.cpp:385-394
    prove_timer.start();
    F c[3];
    check(zk_sumcheck_liu_update(ctx_, w(previous_random), w(c[0])), "zk_sumcheck_liu_update");
    prove_timer.stop();
    proof_size += ZK_F_BYTES * 3;
    if (transcript_) for (int k = 0; k < 3; ++k) transcript_->put_fr(w(c[k]));
    return quadratic_poly(c[0], c[1], c[2]);
}

vector<quadratic_poly> prover::sumcheckUpdateAll(int which, const vector<F> &r, int n_rounds) {
    vector<quadratic_poly> polys;
    if (n_rounds <= 0) return polys;
    prove_timer.start();
    vector<F> prevs(n_rounds), c((size_t) 3 * n_rounds);
    prevs[0].clear();
    for (int j = 1; j < n_rounds; ++j) prevs[j] = r[j - 1];
    check(zk_sumcheck_update_batch(ctx_, which, w(prevs[0]), (uint32_t) n_rounds, w(c[0])), "zk_sumcheck_update_batch");
    prove_timer.stop();
    proof_size += (u64) ZK_F_BYTES * 3 * n_rounds;
    polys.reserve(n_rounds);
    for (int j = 0; j < n_rounds; ++j) {
        if (transcript_) for (int k = 0; k < 3; ++k) transcript_->put_fr(w(c[3 * j + k]));
        polys.emplace_back(c[3 * j], c[3 * j + 1], c[3 * j + 2]);
    }
    return polys;
}

vector<cubic_poly> prover::sumcheckDotProdUpdateAll(const vector<F> &r, int n_rounds) {
    vector<cubic_poly> polys;
    if (n_rounds <= 0) return polys;
    prove_timer.start();
    vector<F> prevs(n_rounds), c((size_t) 4 * n_rounds);
    prevs[0].clear();
    for (int j = 1; j < n_rounds; ++j) prevs[j] = r[j - 1];
    check(zk_sumcheck_dotprod_update_batch(ctx_, w(prevs[0]), (uint32_t) n_rounds, w(c[0])), "zk_sumcheck_dotprod_update_batch");
    prove_timer.stop();
    polys.reserve(n_rounds);
    for (int j = 0; j < n_rounds; ++j) {
        proof_size += ZK_F_BYTES * (3 + (!c[4 * j].isZero()));   // src/prover.cpp:137
        if (transcript_) for (int k = 0; k < 4; ++k) transcript_->put_fr(w(c[4 * j + k]));
        polys.emplace_back(c[4 * j], c[4 * j + 1], c[4 * j + 2], c[4 * j + 3]);
    }
    return polys;
}

hyrax_bls12_381::polyProver &prover::commitInput(const vector<G> &gens) {   // src/prover.cpp:503-511
    // pad val[0] with zeros to 2^bit_length (src/prover.cpp:504-508); once: the padding stays zero between proofs, and clearing
    // 3.9 M elements of page-locked memory again would cost ~10 ms per vgg11 proof
    if (val[0].size() != (1ULL << C.circuit[0].bit_length)) {
        const size_t old = val[0].size();
        val[0].resize(1ULL << C.circuit[0].bit_length);
        for (size_t i = old; i < val[0].size(); ++i) val[0][i].clear();
    }
#ifdef ZKCNN_DROPIN_CPU_HYRAX
    poly_p = std::make_unique<hyrax_bls12_381::polyProver>(val[0], gens, transcript_);
#else
    // the device copy of val[0] is already zero-padded (zk_witness_layer); nothing is uploaded again
    if (msm_digit_bits_ >= 6 && msm_digit_bits_ <= 8) check(zk_set_tunable(ctx_, "msm_digit_bits", msm_digit_bits_), "zk_set_tunable");
    poly_p = std::make_unique<hyrax_bls12_381::polyProver>(ctx_, gens, (unsigned char) C.circuit[0].bit_length, transcript_);
#endif
    return *poly_p;
}
