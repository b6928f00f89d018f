/*
 * zkcnn_b200 -- C ABI of the B200-native zkCNN prover hot path.
 *
 * The reference (TAMUCrypto/zkCNN) has no FFI layer: its hot path is two C++ classes, `prover` (src/prover.hpp:16-78)
 * and `hyrax_bls12_381::polyProver` (3rd/hyrax-bls12-381/src/polyProver.hpp:18-57), called directly by the verifier.
 * This header is the boundary a drop-in puts underneath those classes: every entry point below replaces one public
 * member of one of them (cited as file:line of /root/reference) and is what a cgo / JNI / ctypes binding would bind.
 * zkcnn_b200/host/prover.hpp and polyProver.hpp are the C++ shims with the reference's signatures on top of it.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++ or torch types.
 *   - A field element Fr crosses as uint64_t[4] in mcl's in-memory form: little-endian limbs of a*2^256 mod r
 *     (Montgomery), fully reduced.  `std::vector<Fr>::data()` of the reference can be passed as is.
 *   - A G1 point crosses as uint64_t[18]: Jacobian (x, y, z), each coordinate uint64_t[6] Montgomery mod p
 *     (R = 2^384), the layout of mcl's G1.  Points RETURNED by this library are normalised: z = 1, or
 *     x = y = z = 0 for the point at infinity.
 *   - All pointers are HOST pointers; the library owns every device buffer.  Calls are synchronous: outputs are
 *     written before the call returns.
 *   - Every function returns 0 on success and a negative value on failure; zk_last_error() describes the failure
 *     of the calling thread's last call.  There is no CPU fallback: without a CUDA device zk_ctx_create fails.
 *   - One zk_ctx per (process, GPU); calls on one ctx must be serialised by the caller (the reference prover is
 *     not re-entrant either, src/prover.cpp:9).
 */
#ifndef ZKCNN_B200_H
#define ZKCNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZK_FR_WORDS 4
#define ZK_G1_WORDS 18

typedef struct zk_ctx zk_ctx;

/* ---- library / context ------------------------------------------------------------------------------------------ */
const char *zk_last_error(void);
const char *zk_version(void);      /* "zkcnn_b200 <ver> sm_100a" (or "... emu" for the test-only host build) */
int zk_device_count(void);
zk_ctx *zk_ctx_create(int device); /* NULL on failure */
void zk_ctx_destroy(zk_ctx *ctx);
/* number of kernels launched through this ctx so far (bench.py's gpu_launches) */
uint64_t zk_ctx_launch_count(const zk_ctx *ctx);

/* Per-kernel-class device timing (CUDA events on the launching stream around every launch) with the ALGORITHMIC bytes
 * each launch moves, for roofline reporting.  Off by default (two event records per launch when on). */
enum { ZK_PROF_FOLD = 0,   /* K1/K2 sumcheck round kernels                    */
       ZK_PROF_GATES,      /* K4/K5 gate gather-reduce                        */
       ZK_PROF_MSM,        /* K8/K9 bucket MSM                                */
       ZK_PROF_TABLES,     /* K3 eq / phi tables                              */
       ZK_PROF_DENSE,      /* K4b/K5b/K6 dense contractions, gathers, scatter */
       ZK_PROF_OTHER,
       ZK_PROF_FOLD_SMALL, /* K1/K2 rounds on tables < 32 MiB: latency-bound   */
       ZK_PROF_CLASSES };
int zk_profile_enable(zk_ctx *ctx, int on);     /* also clears the counters */
int zk_profile_get(zk_ctx *ctx, int cls, double *ms, uint64_t *launches, uint64_t *bytes);
/* out[2]: mixed point additions performed by the MSM kernels since profiling was enabled: [0] small-multiples kernel (one per non-zero
 * one-byte scalar), [1] bucket accumulation of the window kernel.  11 Fp multiplications each: the ALU-side work of the MSM class. */
int zk_profile_msm_ops(zk_ctx *ctx, uint64_t *out);

/* Kernel-selection thresholds of the sumcheck rounds (tests and experiments; defaults in parentheses):
 *   "thin_max_pairs"   (16384)  rounds with at most this many output pairs per table use k_round_quad_thin
 *   "tma_min_entries"  (131072) fold rounds on tables of at least this many entries use the TMA-staged k_round_quad_tma
 *   "derive_b"         (1)      streaming rounds take b from the previous round's polynomial (0: always three products)
 *   "pdl"              (1)      k_round_quad_thin is launched with programmatic stream serialization
 *   "eval_schedules"   (1)      zk_circuit_layer also builds the evaluation schedules of zk_witness_generate (+12 bytes per gate)
 *   "axpy_splits"      (0)      K5b: threads that share one output of k_dotprod_axpy (0: chosen from the shape)
 *   "unit_batch"       (0)      zk_fold_rounds2 runs its rounds through the phase-batched path (as zk_sumcheck_update_batch does)
 *   "tail"             (1)      batched phases run all rounds on tables of at most tail_max_entries (1024) in one launch (k_round_tail)
 *   "cubic_tma"        (1)      DOT_PROD fold rounds on tables of at least tma_min_entries use k_round_cubic_tma
 *   "cubic_max_grid"   (none)   cap on the CTAs of a K2 launch (tests: several iterations per thread on small tables)
 *   "cubic_factored_min_iters" (4) k_round_cubic switches to the factored form from this many output pairs per thread
 *   "msm_few_rows_chunk" (2048) (generator, window) entries per work item of the bucket kernels for MSMs of at most 8 rows
 *   "msm_split" (1)      MSMs of at most 8 rows as accumulate / merge / reduce launches; 0: the self-contained bucket kernel
 *   "msm_digit_bits" (8)  digit width of the small-multiples commitment path: 2^bits - 1 multiples per generator are tabulated and a scalar
 *                         costs one addition per non-zero digit; 6 suits witnesses whose magnitudes are mostly below 64 (zkh_build picks it)
 *   "msm_batch_chunk" (4096) the same for the MSM of the batched opening (zk_poly_bullet_prove_all)
 *   "msm_small_seg" (1024) scalars of a row that one warp of the small-multiples kernel takes
 *   "msm_host_finish" (1) opening rounds: the last 14 point operations and the normalisation of the two points on the host */
int zk_set_tunable(zk_ctx *ctx, const char *name, uint64_t value);

/* page-lock / unlock a caller-owned host buffer so that uploads from it are direct DMA (cudaHostRegister) */
int zk_host_pin(const void *p, size_t bytes);
int zk_host_unpin(const void *p);

/* ---- circuit + witness upload: replaces neuralNetwork's friend access to prover::C / prover::val ------------------
 * (src/prover.hpp:48-49,76-77; src/neuralNetwork.cpp:64-68).  Gate structs are bit-identical to src/circuit.h:15-33. */
typedef struct { uint32_t g, u; uint8_t lu, sc; } zk_uni_gate;        /* 12 bytes, == uniGate */
typedef struct { uint32_t g, u, v; uint8_t sc, l; } zk_bin_gate;      /* 16 bytes, == binGate */

/* values of enum class layerType, src/circuit.h:35-37 */
enum {
    ZK_LAYER_INPUT = 0, ZK_LAYER_FFT, ZK_LAYER_IFFT, ZK_LAYER_ADD_BIAS, ZK_LAYER_RELU, ZK_LAYER_SQR, ZK_LAYER_OPT_AVG_POOL,
    ZK_LAYER_MAX_POOL, ZK_LAYER_AVG_POOL, ZK_LAYER_DOT_PROD, ZK_LAYER_PADDING, ZK_LAYER_FCONN, ZK_LAYER_NCONV,
    ZK_LAYER_NCONV_MUL, ZK_LAYER_NCONV_ADD
};

typedef struct {                     /* the scalar members of class layer, src/circuit.h:39-74 */
    int32_t ty;
    uint32_t size, size_u[2], size_v[2];
    int8_t bit_length_u[2], bit_length_v[2], bit_length, max_bl_u, max_bl_v;
    uint8_t need_phase2;
    uint32_t zero_start_id;
    int8_t fft_bit_length;
    uint64_t scale[ZK_FR_WORDS];
    const zk_uni_gate *uni_gates; uint64_t n_uni;
    const zk_bin_gate *bin_gates; uint64_t n_bin;
    const uint32_t *ori_id_u;        /* size_u[0] entries */
    const uint32_t *ori_id_v;        /* size_v[0] entries */
} zk_layer_desc;

/* layeredCircuit::size and ::two_mul (src/circuit.h:77-82, src/circuit.cpp:90-100) */
int zk_circuit_begin(zk_ctx *ctx, uint32_t n_layers, const uint64_t *two_mul, uint32_t n_two_mul);
int zk_circuit_layer(zk_ctx *ctx, uint32_t layer_id, const zk_layer_desc *desc);
int zk_circuit_end(zk_ctx *ctx);     /* builds the device-resident gate schedules */
/* prover::val[layer_id] (src/prover.hpp:49) */
int zk_witness_layer(zk_ctx *ctx, uint32_t layer_id, const uint64_t *val, uint64_t n);
/* double-buffered upload: copy the NEXT proof's prover::val[layer_id] on a second stream while the current proof runs
 * (val must stay valid and unchanged until the commit), then make all prefetched layers current */
int zk_witness_layer_prefetch(zk_ctx *ctx, uint32_t layer_id, const uint64_t *val, uint64_t n);
int zk_witness_commit_prefetch(zk_ctx *ctx);
/* prover::val[layer_id] in a compact encoding (8 instead of 32 bytes per element over PCIe): small[i] = the value as a signed
 * 64-bit integer (mcl's sign convention), except at wide_idx[0..n_wide), whose full Montgomery values are wide_val[];
 * prefetch != 0: into the shadow buffer on the copy stream, made current by zk_witness_commit_prefetch */
int zk_witness_layer_compact(zk_ctx *ctx, uint32_t layer_id, const int64_t *small, uint64_t n, const uint32_t *wide_idx, const uint64_t *wide_val,
                             uint32_t n_wide, int prefetch);

/* ---- witness generation on the device: what neuralNetwork::create computes while it builds the circuit (src/neuralNetwork.cpp:899-965,
 * src/utils.cpp:105-145): gate values layer by layer, number-theoretic transforms of the FFT layers, bit decompositions -----------------
 * zk_circuit_aux_ops: the auxiliary inputs the construction of layer `layer_id` derives from earlier gate values (prepareSignBit /
 * prepareDecmpBit / prepareMax), n triples {src, dst, meta}: dst indexes val[0]; meta bits 0-7 = bit position, bits 8-9 = 0 sign bit,
 * 1 magnitude bit, 2 running maximum of max(0, value); bit 10 = the source is val[0][src], else val[layer_id - 1][src].  Call it for every
 * layer >= 1 (n = 0 where there is none) after zk_circuit_layer. */
int zk_circuit_aux_ops(zk_ctx *ctx, uint32_t layer_id, const uint32_t *ops, uint64_t n);
/* val[0][0, n_image) <- image, then every auxiliary input and every layer re-evaluated on the device; the weights in val[0] are the
 * resident ones (upload a complete witness once).  ranges (NULL or 2 x n_layers values): per layer the largest non-negative value and
 * the largest magnitude of a negative one (what getNextBit, :967-977, turns into the next quantisation scale). */
int zk_witness_generate(zk_ctx *ctx, const uint64_t *image, uint64_t n_image, uint64_t *ranges);
/* the same with the ranges restricted to the layers whose flag is set in want_range[n_layers] (NULL: all) */
int zk_witness_generate_layers(zk_ctx *ctx, const uint64_t *image, uint64_t n_image, uint64_t *ranges, const uint8_t *want_range);
int zk_witness_read(zk_ctx *ctx, uint32_t layer_id, uint64_t first, uint64_t n, uint64_t *out);
/* FNV-1a-64 of the canonical encodings of val[layer_id][0, n) on the device (parity tests: the h_val column of the golden circuit dumps) */
int zk_debug_layer_hash(zk_ctx *ctx, uint32_t layer_id, uint64_t n, uint64_t *fnv1a);

/* ---- GKR prover: one entry point per public member of class prover ---------------------------------------------- */
int zk_prover_init(zk_ctx *ctx);                                                             /* prover.cpp:17  */
int zk_vres(zk_ctx *ctx, const uint64_t *r, uint32_t output_size, uint32_t r_size, uint64_t *out); /* :434 */
int zk_sumcheck_init_all(zk_ctx *ctx, const uint64_t *r_0, uint32_t n);                      /* :28  */
int zk_sumcheck_init(zk_ctx *ctx, const uint64_t *alpha, const uint64_t *beta);              /* :43  */
int zk_sumcheck_dotprod_init_phase1(zk_ctx *ctx);                                            /* :57  */
int zk_sumcheck_init_phase1(zk_ctx *ctx, const uint64_t *relu_rou);                          /* :155 */
int zk_sumcheck_init_phase2(zk_ctx *ctx);                                                    /* :241 */
int zk_sumcheck_dotprod_update1(zk_ctx *ctx, const uint64_t *previous_random, uint64_t *abcd /*4 Fr*/);  /* :103 */
int zk_sumcheck_update1(zk_ctx *ctx, const uint64_t *previous_random, uint64_t *abc /*3 Fr*/);           /* :360 */
int zk_sumcheck_update2(zk_ctx *ctx, const uint64_t *previous_random, uint64_t *abc /*3 Fr*/);           /* :364 */
/* Every round of one phase in one call, queued on the device without a host round trip in between: which = 1 / 2 for
 * sumcheckUpdate1 / sumcheckUpdate2 of the current layer, 0 for sumcheckLiuUpdate.  previous_randoms[j] is the argument
 * round j would get (previous_randoms[0] = 0); abc receives n_rounds x 3 Fr, the same values the per-round calls return.
 * The reference's verifier draws all challenges of a phase before its first round (src/verifier.cpp:156-160,207,275-279),
 * which is what makes this call possible for its own driver loop; the per-round entry points above stay the drop-in API. */
int zk_sumcheck_update_batch(zk_ctx *ctx, int which, const uint64_t *previous_randoms, uint32_t n_rounds, uint64_t *abc);
/* The same for the DOT_PROD phase: every sumcheckDotProdUpdate1 round of the current layer in one call (the verifier draws
 * r_u[i] before the first round for these layers too, src/verifier.cpp:156-160); abcd receives n_rounds x 4 Fr. */
int zk_sumcheck_dotprod_update_batch(zk_ctx *ctx, const uint64_t *previous_randoms, uint32_t n_rounds, uint64_t *abcd);
int zk_sumcheck_dotprod_finalize1(zk_ctx *ctx, const uint64_t *previous_random, uint64_t *claim_1);      /* :146 */
int zk_sumcheck_finalize1(zk_ctx *ctx, const uint64_t *previous_random, uint64_t *claim_0, uint64_t *claim_1); /* :459 */
int zk_sumcheck_finalize2(zk_ctx *ctx, const uint64_t *previous_random, uint64_t *claim_0, uint64_t *claim_1); /* :473 */
int zk_sumcheck_liu_init(zk_ctx *ctx, const uint64_t *s_u, const uint64_t *s_v, uint32_t n);  /* :312 */
int zk_sumcheck_liu_update(zk_ctx *ctx, const uint64_t *previous_random, uint64_t *abc);      /* :385 */
int zk_sumcheck_liu_finalize(zk_ctx *ctx, const uint64_t *previous_random, uint64_t *claim_1);/* :487 */

/* ---- Hyrax polynomial commitment: one entry point per public member of class polyProver --------------------------
 * (3rd/hyrax-bls12-381/src/polyProver.cpp).  A ctx holds one polynomial at a time. */
/* prover::commitInput (prover.cpp:503): zero-pads val[0] to 2^bit_length and binds it, on the device, as Z */
int zk_poly_bind_input(zk_ctx *ctx, const uint64_t *gens, uint32_t n_gens);
/* polyProver::polyProver(Z, gens) (polyProver.cpp:12-17) for a stand-alone polynomial */
int zk_poly_create(zk_ctx *ctx, const uint64_t *Z, uint64_t n, const uint64_t *gens, uint32_t n_gens);
int zk_poly_commit(zk_ctx *ctx, uint64_t *comm_out /* 2^(bl/2) points */, uint32_t n_out);     /* :19  */
int zk_poly_evaluate(zk_ctx *ctx, const uint64_t *x, uint32_t n, uint64_t *out);              /* :36  */
int zk_poly_init_bullet_prove(zk_ctx *ctx, const uint64_t *lx, uint32_t n_lx, const uint64_t *rx, uint32_t n_rx); /* :52 */
int zk_poly_bullet_prove(zk_ctx *ctx, uint64_t *lcomm, uint64_t *rcomm, uint64_t *ly, uint64_t *ry);  /* :76  */
int zk_poly_bullet_update(zk_ctx *ctx, const uint64_t *randomness);                           /* :98  */
int zk_poly_bullet_open(zk_ctx *ctx, uint64_t *out);                                          /* :111 */
/* Not in the reference: all remaining rounds (bulletProve + bulletUpdate, :76-109) in one device pass, for a caller that has drawn the
 * randomness of every round beforehand -- the reference's verifier draws it from its RNG independently of the prover's messages
 * (polyVerifier.cpp:48-50), so the messages are the same.  n_rounds must be the number of rounds left (log2 of the current length);
 * lcomm / rcomm receive n_rounds points (18 words each), ly / ry n_rounds scalars.  Afterwards only zk_poly_bullet_open remains. */
int zk_poly_bullet_prove_all(zk_ctx *ctx, const uint64_t *randomness, uint32_t n_rounds, uint64_t *lcomm, uint64_t *rcomm, uint64_t *ly,
                             uint64_t *ry);

/* ---- verifier side on the device (the caller of the hot path): the wiring predicates of verifier::betaInitPhase1/2
 * and predicatePhase1/2 (src/verifier.cpp:36-123) and the input-layer term gr (:307-325), computed from the VERIFIER's challenges with the
 * kernels of the prover's Init* passes over the resident circuit topology.  Tables live in slots 0..7 of the context. ------------------- */
/* slot <- init0 * eq(r0) (+ init1 * eq(r1) if r1 != NULL) over `bits` variables; entries >= tail_start times tail_scale if tail_scale != NULL */
int zk_vtab_eq(zk_ctx *ctx, int slot, uint32_t bits, const uint64_t *r0, const uint64_t *init0, const uint64_t *r1, const uint64_t *init1, uint32_t tail_start,
               const uint64_t *tail_scale);
/* slot_out[g] = hi[g >> lo_bits] * lo[g & (2^lo_bits - 1)] */
int zk_vtab_outer(zk_ctx *ctx, int slot_out, int slot_hi, int slot_lo, uint32_t bits, uint32_t lo_bits);
/* slot <- phiGInit(rx, scale, n, is_ifft), 2^n entries */
int zk_vtab_phi(zk_ctx *ctx, int slot, const uint64_t *rx, const uint64_t *scale, uint32_t n, int is_ifft);
/* out = sum_{i < n} slot_a[i] * slot_b[i] */
int zk_vtab_dot(zk_ctx *ctx, int slot_a, int slot_b, uint64_t n, uint64_t *out);
/* out[5 Fr] = uni_value[0], uni_value[1] (before the beta_v[0] factor of src/verifier.cpp:111-112), bin_value[0..2]; slot_beta_v < 0: no phase 2 */
int zk_verifier_layer_predicates(zk_ctx *ctx, uint32_t layer_id, int slot_beta_g, int slot_beta_u, int slot_beta_v, uint64_t *out);
/* gr of verifier::verifyFirstLayer; r_u / r_v = the challenge vectors of the layers that have layer-0 operands, back to back */
int zk_verifier_input_predicate(zk_ctx *ctx, int slot_beta0, const uint64_t *s_u, const uint64_t *s_v, const uint64_t *r_u, const uint64_t *r_v, uint64_t *out);

/* ---- stateless primitives: the kernels behind the calls above, exposed for parity tests and micro-benchmarks ------ */
/* out[i] = a[i] (+,-,*) b[i];  op: 0 add, 1 sub, 2 mul */
int zk_fr_vec_op(zk_ctx *ctx, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, uint64_t n);
/* initBetaTable(beta, bits, r, init)  (src/utils.cpp:168-180; also hyrax expand(), utils.cpp:29-62 with init = 1) */
int zk_beta_table(zk_ctx *ctx, const uint64_t *r, uint32_t bits, const uint64_t *init, uint64_t *out);
/* phiGInit (src/utils.cpp:61-103); out has 2^n (IFFT) or 2^(n-1) (FFT) entries */
int zk_phi_table(zk_ctx *ctx, const uint64_t *rx, const uint64_t *scale, uint32_t n, int is_ifft, uint64_t *out);
/* Runs `n_rounds` rounds of sumcheckUpdateEach on one (V, mult) pair of 2^bits entries (first `live` non-zero):
 * round j uses previous_random = (j == 0 ? 0 : r[j-1]).  polys gets 3 Fr per round.  Returns folded tables if non-NULL. */
int zk_fold_rounds(zk_ctx *ctx, const uint64_t *V, const uint64_t *M, uint32_t bits, uint64_t live, const uint64_t *r,
                   uint32_t n_rounds, uint64_t *polys);
/* The same on TWO table pairs folded in lock step, as prover::sumcheckUpdate does (src/prover.cpp:368-383): pair 0 (2^bits0 entries,
 * bits0 < 0: absent) and pair 1; a pair that is exhausted first collapses into add_term.  n_rounds <= max(bits0, bits1). */
int zk_fold_rounds2(zk_ctx *ctx, const uint64_t *V0, const uint64_t *M0, int32_t bits0, uint64_t live0, const uint64_t *V1, const uint64_t *M1, int32_t bits1,
                    uint64_t live1, const uint64_t *r, uint32_t n_rounds, uint64_t *polys);
/* prover::Vres (src/prover.cpp:434-457) on a stand-alone vector: the multilinear extension of values[0..n) at r[0..r_size) */
int zk_mle_eval(zk_ctx *ctx, const uint64_t *values, uint32_t n, const uint64_t *r, uint32_t r_size, uint64_t *out);
/* Runs `n_rounds` rounds of sumcheckDotProdUpdate1 (src/prover.cpp:103-144) on stand-alone tables: mult has 2^m_bits entries
 * (mult_array[1]), V0 / V1 have 2^bits entries of which the first live0 / live1 are non-zero (V_mult[0] / V_mult[1];
 * live0 <= live1).  Round j uses previous_random = (j == 0 ? 0 : r[j-1]).  polys gets 4 Fr (a, b, c, d) per round. */
int zk_cubic_rounds(zk_ctx *ctx, const uint64_t *mult, uint32_t m_bits, const uint64_t *V0, uint64_t live0, const uint64_t *V1, uint64_t live1, uint32_t bits,
                    const uint64_t *r, uint32_t n_rounds, uint64_t *polys);
/* out[k] = sum_j scalars[k*n + j] * bases[j]   (G1::mulVec semantics, mcl ec.hpp:1570-1597), k < n_rows */
int zk_msm(zk_ctx *ctx, const uint64_t *bases, const uint64_t *scalars, uint64_t n, uint32_t n_rows, uint64_t *out);
/* element-wise G1: op 0: out = a + b; 1: out = 2a; 2: out = k*a (k = b reinterpreted as Fr per element) */
int zk_g1_vec_op(zk_ctx *ctx, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, uint64_t n);
/* out[i] = scalars[i] * base for ONE base point (fixed-base comb; the verifier's generator set-up, src/verifier.cpp:121-126) */
int zk_g1_fixed_base_mul(zk_ctx *ctx, const uint64_t *base, const uint64_t *scalars, uint64_t n, uint64_t *out);
/* FNV-1a-64 over the canonical 32-byte little-endian encodings of a bookkeeping table of the sumcheck in progress (entries beyond
 * the live part count as zeros), and its entry count: sel 0 / 1 = V / mult table of pair 0, 2 / 3 = of pair 1, 4 = the multiplier
 * table of a DOT_PROD phase.  For per-function parity tests against the reference's tables (oracle/harness: --dump-dir). */
int zk_debug_table_hash(zk_ctx *ctx, int sel, uint64_t *fnv1a, uint64_t *n_entries);
/* device self-test: inline-PTX field arithmetic against the portable implementation on n random inputs; 0 = equal */
int zk_selftest(zk_ctx *ctx, uint64_t seed, uint32_t n);

/* ---- device-timed micro-benchmarks (CUDA events on the launching stream; inputs resident in HBM) ------------------ */
/* one K1 round on a table pair of 2^bits entries, `iters` launches; returns average ms per launch in *ms */
int zk_bench_fold(zk_ctx *ctx, uint32_t bits, uint32_t iters, int fold, float *ms);
/* one K2 (cubic) fold round on tables of 2^bits entries with V_mult[0] live over the first 2^live0_bits and a multiplier of 2^m_bits entries */
int zk_bench_cubic(zk_ctx *ctx, uint32_t bits, uint32_t live0_bits, uint32_t m_bits, uint32_t iters, float *ms);
/* sustained Fr (is_fp = 0) / Fp (is_fp = 1) Montgomery multiplications per second with every SM busy, in 10^9 per second: the ALU-side peak */
int zk_bench_field_mul(zk_ctx *ctx, int is_fp, float *gmul_per_s);
int zk_bench_msm(zk_ctx *ctx, uint32_t log_rows, uint32_t log_cols, int scalar_mix, uint32_t iters, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* ZKCNN_B200_H */
