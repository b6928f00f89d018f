// Device-resident prover state behind the C ABI (include/zkcnn_b200.h).
#pragma once
#include <map>
#include "../../include/zkcnn_b200.h"
#include "g1.cuh"
#include "rt.hpp"
#include "sc_kernels.cuh"
#include "cubic_kernels.cuh"
#include "witness_kernels.cuh"
#include <memory>
#include <vector>

namespace zk {

// ---- static gate schedules (built once per circuit at upload) -------------------------------------------------------
struct level_t {
    rt::dbuf items;
    uint32_t n_items = 0;
    uint32_t n_partials = 0;   // partial sums this level writes
};
struct schedule_t {
    rt::dbuf recs;
    uint64_t n_recs = 0, n_val_recs = 0;   // n_val_recs: phase-1 records that also read a value operand
    std::vector<level_t> levels;
    uint32_t max_partials = 0;
    bool has_scalar = false;   // phase 2: uni gates feed add_term
};

struct layer_t {
    zk_layer_desc d{};         // scalar members only (pointers nulled)
    fr_t scale;
    rt::dbuf ori_u, ori_v;     // device copies of ori_id_u / ori_id_v
    schedule_t p1, p2;
    // witness generation on the device (capi_witness.cuh): gates sorted by OUTPUT gate; DOT_PROD: CSR by output block; the auxiliary
    // inputs this layer's construction derives from earlier layers (bit decompositions, window maxima), in execution order
    schedule_t ev;
    uint32_t ev_rows_covered = 0;   // output gates that have at least one source gate (== size: no zero-fill needed)
    rt::dbuf dpe_rowptr, dpe_gates;
    uint32_t dpe_rows = 0;
    rt::dbuf aux_bits_prev, aux_max, aux_bits_l0;   // aux_op_t arrays: bits of val[l-1] entries, maxima over val[l-1] entries, bits of val[0] entries
    uint64_t n_aux_bits_prev = 0, n_aux_max = 0, n_aux_bits_l0 = 0;
    uint32_t aux_max_base = 0, aux_max_count = 0;   // the maxima land in val[0][base, base + count)
    bool aux_loaded = false;
    rt::dbuf dp_rowptr, dp_gates;  // DOT_PROD phase 1: CSR by u
    uint32_t dp_rows = 0;
    uint32_t dp_rows_live = 0;     // rows u < dp_rows_live carry gates: V_mult[0] of the DOT_PROD phase is zero beyond them
    rt::dbuf val;              // prover::val[layer]
    uint64_t n_val = 0;
    rt::dbuf val_next;         // shadow copy filled by zk_witness_layer_prefetch (the next proof's witness)
    uint64_t n_val_next = 0;
    bool next_ready = false;
    rt::dbuf compact_stage, wide_stage;   // zk_witness_layer_compact: staging of the int64 values (when not mapped) and of the wide ones
    bool have_desc = false;
};

// one bookkeeping table in evaluation form
struct table_t {
    const fr_t *cur = nullptr;
    rt::dbuf init, fold[2];
    int next = 0;
};
// a (V, mult) pair that is folded in lock step (src/prover.cpp:396-426)
struct pair_t {
    table_t v, m;
    uint32_t n_eval = 0;       // evaluations currently held (0: pair absent or already collapsed)
    uint32_t live = 0;         // entries >= live are zero
    uint32_t poly_round = 0;   // round whose polynomial of this pair is in zk_ctx::round_state (0: none); see round_args_t::derive_b
    bool exists = false;       // bit_length != -1
    bool collapsed = false;
    fr_t cv, cm;               // values after collapse (V_mult[b][0].b, mult_array[b][0].b)
};

// Hyrax state (3rd/hyrax-bls12-381/src/polyProver.hpp:38-54)
struct hyrax_t {
    bool bound = false;
    const fr_t *Z = nullptr;   // 2^bit_length scalars on the device (aliases val[0] or z_own)
    rt::dbuf z_own;
    uint32_t bit_length = 0, l_bits = 0, r_bits = 0;
    uint32_t n_gens = 0;
    rt::dbuf gens_aff;         // n_gens affine points
    rt::dbuf gens_jac;         // upload staging
    rt::dbuf table;            // fixed-base window table: [kWindows][n_gens] affine, entry = 2^(8w) * gens[j]
    bool table_ready = false;
    bool table_pending = false;   // the build is queued on the side stream; the main stream has not waited for it yet
    rt::dbuf table_scratch;    // z's and prefix products of the table build
    rt::dbuf mult;             // small-multiples table [n_gens][255] affine, entry = d * gens[j]  (msm_kernels.cuh)
    bool mult_ready = false;
    uint32_t mult_bits = 8;    // digit width the table was built for: 2^mult_bits - 1 entries per generator
    uint64_t gens_hash = 0;
    std::vector<uint8_t> gens_host;   // the generator set the tables were built for (confirms a hash hit)
    rt::dbuf L, R, RZ, a, a_next, coef, scal, bullet_dots;
    std::vector<fr_t> t;       // remaining opening point (lx)
    fr_t scale;
    uint32_t cur = 0;          // current length of bullet_a
    uint32_t round = 0;
    std::vector<fr_t> rinv;    // 1 / randomness of the finished rounds
    rt::dbuf msm_out, msm_small, msm_small_hi, msm_rowinfo, msm_wide_rows, msm_buckets, msm_item_entries, msm_merged, msm_S, pts_out;
};

}  // namespace zk

struct zk_ctx {
    int device = 0;
    zk_stream_t stream{};
    zk_stream_t aux_stream{};    // side work of the main stream (window-table build next to the commitment)
    zk_stream_t copy_stream{};   // witness prefetch (zk_witness_layer_prefetch): overlaps the running proof
    uint64_t launches = 0;

    // circuit
    uint32_t n_layers = 0;
    std::vector<zk::layer_t> layers;
    std::vector<zk::fr_t> two_mul_h;
    zk::rt::dbuf two_mul;
    bool circuit_ready = false;
    bool eval_schedules = true;   // zk_circuit_layer also builds the evaluation schedules of the device witness generator (+12 bytes per gate)

    // sumcheck state (members of class prover, src/prover.hpp:55-72)
    uint32_t sumcheck_id = 0, round = 0;
    zk::fr_t alpha, beta, relu_rou, add_term, V_u0, V_u1;
    std::vector<std::vector<zk::fr_t>> r_u, r_v;
    zk::pair_t pair[2];
    zk::table_t mdp;            // DOT_PROD multiplier table (mult_array[1] of size 2^fft_bl)
    uint32_t mdp_n = 0;
    bool in_dotprod_p1 = false; // pair[1].m plays V_mult[0] of a DOT_PROD phase (live below dp_live0)
    uint32_t dp_live0 = 0;      // live entries of V_mult[0] in the DOT_PROD phase (cubic_args_t::live0)
    zk::rt::dbuf beta_g, beta_g_alt, beta_gs, beta_u;
    uint32_t beta_g_entries = 0;

    // scratch
    zk::rt::dbuf half[4], d_r, partials, counters, round_acc, cubic_acc, round_state, round_out, gate_partial[2], dense_partial, vres_scratch, scalar_slot;
    zk::fr_t *h_out = nullptr;  // pinned mirror of round_out
    zk::g1_jac_t *msm_S_h = nullptr;   // pinned: partial sums of a few-row MSM finished on the host (msm_finish_host)
    // result mailbox of the per-round kernels: mapped pinned host memory the last CTA writes directly, followed by a
    // sequence number; the host spins on the sequence number instead of copying + synchronising the stream
    zk::fr_t *res_h = nullptr, *res_d = nullptr;          // [32] host / device view
    uint32_t *flag_h = nullptr, *flag_d = nullptr;
    zk::rt::dbuf batch_res;                               // [rounds][16] result blocks of a batched phase
    zk::fr_t *batch_h = nullptr;
    void *stage_h = nullptr;                              // page-locked staging buffer (h2d_staged / d2h_staged)
    uint32_t batch_cap = 0;
    uint32_t *tag_h = nullptr, *tag_d = nullptr;          // [32] tagged mailbox (publish_tagged)
    uint32_t seq = 0;
    uint32_t thin_max_pairs = 1u << 14;      // see zk_set_tunable
    uint64_t tma_min_entries = 1ull << 17;
    uint32_t pdl_enabled = 1;                // k_round_quad_thin launched with programmatic stream serialization
    uint32_t derive_b_enabled = 1;           // streaming rounds: b from the previous round's polynomial (0: always three products)
    uint32_t axpy_splits = 0;                // K5b: threads per output of k_dotprod_axpy (0: chosen from the shape)
    uint32_t unit_batch = 0;                 // zk_fold_rounds2 goes through the phase-batched path (tests)
    uint32_t tail_enabled = 1;               // batched phases: all rounds on tables of at most tail_max_entries in one launch (k_round_tail)
    uint32_t tail_max_entries = 1024;
    uint32_t cubic_tma_enabled = 1;          // DOT_PROD fold rounds on large tables use k_round_cubic_tma
    uint32_t cubic_max_grid = 1u << 20;       // cap on the CTAs of a K2 launch (tests: forces several iterations per thread on small tables)
    uint32_t cubic_factored_min_iters = 4;   // k_round_cubic: factored form from this many output pairs per thread
    uint32_t msm_digit_bits = 8;             // digit width of the small-multiples path (6, 7 or 8): the table has 2^bits - 1 entries per generator
    uint32_t msm_batch_chunk = 4096;         // the same for the batched opening (2 x rounds rows in one MSM)
    uint32_t msm_small_seg = 1024;           // scalars per warp of k_msm_small (one row segment)
    uint32_t msm_host_finish = 1;            // opening rounds: the last 14 point operations + the normalisation of the two points on the host
    uint32_t msm_split = 1;                  // MSMs of at most 8 rows: accumulate / merge / reduce launches (k_msm_bucket_*) instead of k_msm_window
    uint32_t msm_few_rows_chunk = 2048;      // (generator, window) entries per work item of k_msm_window when an MSM has at most 8 rows (32 per generator)
    std::vector<std::pair<uint32_t, zk::rt::dbuf>> phi_pw;  // cached powers of roots of unity, key = n * 2 + is_ifft

    zk::hyrax_t hy;
    zk::rt::dbuf vt[8], vt_m[2];    // verifier-side tables (capi_verifier.cuh): eq / phi tables by slot, mult-table scratch
    uint64_t vt_n[8] = {};
    zk::rt::dbuf wit_scratch;   // device witness generation: window maxima / layer ranges
    zk::rt::dbuf fb_k, fb_out;
    zk::rt::dbuf fb_comb;       // zk_g1_fixed_base_mul: comb table of the last base point
    uint64_t fb_base[ZK_G1_WORDS] = {};   // the base point the comb table was built for
    bool fb_ready = false;

    // optional per-kernel-class timing (zk_profile_*): CUDA events around every launch of the stream
    bool prof_on = false;
    struct prof_rec { int cls; zk::rt::event_t a, b; const char *name = nullptr; };
    std::map<std::string, std::pair<uint64_t, double>> prof_by_kernel;   // ZK_PROF_KERNELS=1: launches and device ms per kernel name (printed by zk_profile_enable(ctx, 0))
    std::vector<prof_rec> prof_pending;
    std::vector<zk::rt::event_t> prof_pool;
    double prof_ms[ZK_PROF_CLASSES] = {};
    uint64_t prof_launches[ZK_PROF_CLASSES] = {}, prof_bytes[ZK_PROF_CLASSES] = {};
    zk::rt::dbuf prof_ops;      // device counters of point additions (MSM kernels, profiling only)
};
