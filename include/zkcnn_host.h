/*
 * zkcnn_host -- C entry points of the stand-alone host side (circuit compiler, witness generator, protocol driver)
 * for callers that cannot use the C++ classes directly (bench.py, the pytest suite, FFI users).
 *
 * This is NOT the drop-in boundary (that is include/zkcnn_b200.h, underneath class prover / polyProver); it wraps the
 * caller side of the hot path -- what the reference's demo mains do (src/main_demo_lenet.cpp:19-40,
 * src/main_demo_vgg.cpp:20-42): build a model, neuralNetwork::create, verifier::verify -- behind plain C.
 * All proving arithmetic still goes through include/zkcnn_b200.h to the GPU; there is no CPU prover here.
 */
#ifndef ZKCNN_HOST_H
#define ZKCNN_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct zkh_session zkh_session;

enum {
    ZKH_REAL_GENERATORS  = 1,  /* Hyrax generators = standard G1 generator * challenge (default: the reference's all-infinity set) */
    ZKH_CHECK_PREDICATES = 2,  /* accepted for compatibility: full verification is the default */
    ZKH_WITNESS_RESIDENT = 4,  /* keep the witness on the device between proofs (skip the host->device copy if present) */
    ZKH_FIXED_GENERATORS = 8,  /* reuse the generators of the previous proof (public parameters), keep the window table */
    ZKH_PREFETCH_NEXT    = 32, /* once this proof has its witness, start copying the witness for the NEXT zkh_prove on a second stream
                                 (double buffering: the copy overlaps this proof; the next proof adopts it instead of uploading) */
    ZKH_NO_HASH          = 64, /* leave zkh_stats.fnv1a at 0 (the FNV-1a of the 0.5 MB transcript costs 0.7 ms per vgg11 proof) */
    ZKH_PROVER_ONLY      = 128,/* prover-only timing: the verifier keeps its per-round sum checks and the field-side checks of the opening but
                                 skips the wiring predicates (getFinalValue), the input-layer "gr" recomputation and every G1 check of the
                                 Hyrax opening (src/verifier.cpp:36-116,307-325; polyVerifier.cpp:25-27,53-58); zkh_stats.checks says what ran */
    ZKH_CSPRNG_CHALLENGES = 256, /* draw the challenges from the operating system's CSPRNG like the reference (Fr::setByCSPRNG,
                                 src/verifier.cpp:124,139,...) instead of the seeded stream: `seed` is ignored, the transcript is not reproducible */
    ZKH_HOST_PREDICATES  = 1024,/* evaluate the verifier's wiring predicates on host threads instead of on the device (zk_verifier_*) */
    ZKH_FIAT_SHAMIR      = 512,/* non-interactive mode: every challenge is derived from the transcript so far (see host/challenge_stream.hpp) */
    ZKH_ROUND_BY_ROUND   = 16  /* one device round trip per sumcheck round (the reference's call pattern) instead of one per phase:
                                 the verifier draws a phase's challenges before its first round either way (src/verifier.cpp:156-160),
                                 so the transcript is the same; default is per phase (zk_sumcheck_update_batch) */
};

typedef struct {
    int32_t ok;                 /* verifier accepted (every check listed in `checks` passed) */
    uint32_t n_layers;
    uint64_t input_size;        /* gates in layer 0 (witness size) */
    uint64_t n_fr, n_g1;        /* field elements / points in the proof */
    uint64_t proof_bytes;       /* canonical encoding (SURVEY.md App. A order; Fr 32 B LE, G1 96 B affine) */
    uint64_t fnv1a;             /* FNV-1a 64 of those bytes */
    uint64_t challenges;        /* verifier challenges drawn */
    uint64_t gpu_launches;      /* kernels launched for this proof */
    double prove_s;             /* prover::proveTime()  (GKR part) */
    double poly_s;              /* prover::polyProverTime()  (Hyrax part) */
    double upload_s;            /* witness (and, first time, circuit) upload */
    double wall_s;              /* prover.init() + verifier.verify(), wall clock */
    double verifier_s;          /* verifier-only work inside wall_s (predicates, point checks) */
    double gkr_kb, poly_kb;     /* proof size as the reference counts it */
    uint64_t h2d_bytes;         /* witness bytes copied host->device for this proof */
    uint32_t checks;            /* what the verifier checked: ZKH_CHECKED_* */
    uint32_t witness_path;      /* 0: the witness zkh_build made (zkh_prove); 1: regenerated on the device; 2: circuit rebuilt on the host (zkh_prove_image) */
} zkh_stats;
enum {
    ZKH_CHECKED_ROUND_SUMS = 1,   /* p(0) + p(1) = claim for every sumcheck round; y and bulletOpen of the opening */
    ZKH_CHECKED_PREDICATES = 2,   /* per-layer getFinalValue (wiring predicates) */
    ZKH_CHECKED_INPUT_GR   = 4,   /* input-layer sumcheck against the recomputed gr */
    ZKH_CHECKED_G1         = 8    /* comm_RZ, generator folds and the final point equation of the Hyrax opening */
};

const char *zkh_last_error(void);
/* model: "lenet" (32x32x1, max pooling), "lenet_cifar", "vgg11", "vgg16" (32x32x3, max pooling) or "vgg" with
 * `network` = channel/pool description, e.g. "64 M 128 M 256 256 M 512 512 M 512 512 M" (src/models.cpp:12-41). */
zkh_session *zkh_create(const char *model, const char *network, int pic_cnt, int device);
void zkh_destroy(zkh_session *s);
int64_t zkh_input_count(zkh_session *s);                                 /* decimals build() consumes */
int zkh_input_file(zkh_session *s, const char *path);                    /* the reference's text format */
int zkh_input_values(zkh_session *s, const double *values, uint64_t n);  /* same numbers, in memory */
/* the reference's text format (whitespace-separated decimals, src/neuralNetwork.cpp:805-897) parsed once into doubles: up to `cap` values
 * into `out`, returns how many the file holds (or -1).  What the input cache / the weight broadcast of zkcnn_b200.load_input build on:
 * the 124 MB vgg11 file costs the reference ~20 s of `ifstream >> double` in every run. */
int64_t zkh_parse_numbers(const char *path, double *out, uint64_t cap);
int zkh_build(zkh_session *s);                                           /* circuit + witness (neuralNetwork::create) */
int zkh_prove(zkh_session *s, uint64_t seed, uint32_t flags, zkh_stats *out);
/* A NEW PICTURE for the model zkh_build prepared (same weights): `pixels` are the picture's decimals in the reference's order (the first
 * zkh_input_count() - weights values of the input stream).  The witness is regenerated ON THE DEVICE (zk_witness_generate: only the picture
 * crosses PCIe; layer values, transforms and bit decompositions are recomputed by CUDA kernels from the resident quantised weights), then
 * proved like zkh_prove.  If the picture leads to different quantisation decisions than the one the circuit was built for (its own scale,
 * or the activation scales that getNextBit derives, src/neuralNetwork.cpp:967-977, which fix the widths of the bit decompositions), or the
 * model uses average pooling, the circuit is rebuilt on the host for the new picture instead; zkh_stats.witness_path says which path ran. */
int zkh_prove_image(zkh_session *s, const double *pixels, uint64_t n_pixels, uint64_t seed, uint32_t flags, zkh_stats *out);
/* only the witness part of zkh_prove_image; returns the path taken (1: regenerated on the device, 2: rebuilt on the host) or -1.  Follow it with
 * zkh_prove(..., ZKH_WITNESS_RESIDENT) when it returned 1. */
int zkh_set_image(zkh_session *s, const double *pixels, uint64_t n_pixels);
/* start the host->device copy of the witness for the next zkh_prove now, on a second stream (what ZKH_PREFETCH_NEXT does
 * from inside a proof) */
int zkh_prefetch_witness(zkh_session *s);
const uint8_t *zkh_proof(zkh_session *s, uint64_t *n_bytes);             /* proof of the last zkh_prove */
int zkh_inferred_class(zkh_session *s, int picture);
void *zkh_context(zkh_session *s);   /* the zk_ctx of this session's prover (NULL before the first proof); for zk_profile_* */
/* per-layer shape/hash dump in the format of oracle/harness/ref_run --circuit-hash (parity tests) */
int zkh_circuit_dump(zkh_session *s, const char *path, int with_hashes);
/* test hook: the following zkh_prove calls append the hashes of the prover's bookkeeping tables after every Init* call to `path` (format of
 * oracle/harness/ref_run --dump-dir: per-function parity against the reference); NULL switches it off */
int zkh_table_dump(zkh_session *s, const char *path);

#ifdef __cplusplus
}
#endif
#endif
