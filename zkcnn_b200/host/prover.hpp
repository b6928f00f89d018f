// Host shim with the public interface of the reference's GKR prover, class prover (src/prover.hpp:16-78).
// Every member forwards through the C ABI (include/zkcnn_b200.h) to the device-resident prover state; the shim keeps
// only what the reference keeps on the host side of the interface: the circuit `C`, the witness `val`, timers and
// the proof-size counter.
//
// Build modes (see polyProver.hpp):
//   ZKCNN_DROPIN           : inside the reference tree; takes the place of src/prover.hpp (same include guard), so the
//                            reference's verifier.cpp / neuralNetwork.cpp / models.cpp compile against it unchanged.
//   ZKCNN_DROPIN_CPU_HYRAX : additionally keep the reference's CPU polyProver (BASELINE.json config 2:
//                            "sumcheck fold + MLE table on GPU, Hyrax MSM still CPU").
#ifndef ZKCNN_PROVER_HPP
#define ZKCNN_PROVER_HPP

#ifdef ZKCNN_DROPIN
#ifndef ZKCNN_DROPIN_CPU_HYRAX
#include "polyProver.hpp"   // ours first: its include guard keeps the reference's hyrax/src/polyProver.hpp out
#else
#include "cpu_hyrax_proxy.hpp"   // the reference's CPU polyProver behind a recording forwarder
#endif
#include "global_var.hpp"   // reference: src/
#include "circuit.h"
#include "polynomial.h"
#else
#include "zk_types.hpp"
#include "polyProver.hpp"
#endif
#include "../../include/zkcnn_b200.h"
#include "transcript.hpp"
#include <memory>
#include <string>
#include <thread>

using std::unique_ptr;

class neuralNetwork;
class singleConv;
class prover {
public:
    prover();
    ~prover();
    prover(const prover &) = delete;
    prover &operator=(const prover &) = delete;

    void init();

    void sumcheckInitAll(const vector<F>::const_iterator &r_0_from_v);
    void sumcheckInit(const F &alpha_0, const F &beta_0);
    void sumcheckDotProdInitPhase1();
    void sumcheckInitPhase1(const F &relu_rou_0);
    void sumcheckInitPhase2();

    cubic_poly sumcheckDotProdUpdate1(const F &previous_random);
    quadratic_poly sumcheckUpdate1(const F &previous_random);
    quadratic_poly sumcheckUpdate2(const F &previous_random);

    F Vres(const vector<F>::const_iterator &r, u32 output_size, u8 r_size);

    void sumcheckDotProdFinalize1(const F &previous_random, F &claim_1);
    void sumcheckFinalize1(const F &previous_random, F &claim_0, F &claim_1);
    void sumcheckFinalize2(const F &previous_random, F &claim_0, F &claim_1);
    void sumcheckLiuFinalize(const F &previous_random, F &claim_1);

    void sumcheckLiuInit(const vector<F> &s_u, const vector<F> &s_v);
    quadratic_poly sumcheckLiuUpdate(const F &previous_random);

    // Not in the reference: every round of a phase in one device call (zk_sumcheck_update_batch).  which = 1 / 2 / 0 stands
    // for sumcheckUpdate1 / sumcheckUpdate2 / sumcheckLiuUpdate; r holds the phase's challenges, drawn by the verifier before
    // the first round (src/verifier.cpp:156-160,207,275-279); round j is folded with r[j - 1].  Returns the n_rounds messages.
    vector<quadratic_poly> sumcheckUpdateAll(int which, const vector<F> &r, int n_rounds);
    // the same for the cubic rounds of a DOT_PROD layer (sumcheckDotProdUpdate1, zk_sumcheck_dotprod_update_batch)
    vector<cubic_poly> sumcheckDotProdUpdateAll(const vector<F> &r, int n_rounds);

    hyrax_bls12_381::polyProver &commitInput(const vector<G> &gens);

    timer prove_timer;
    double proveTime() const { return prove_timer.elapse_sec(); }
    double proofSize() const { return (double) proof_size / 1024.0; }
    double polyProverTime() const { return poly_p->getPT(); }
    double polyProofSize() const { return poly_p->getPS(); }

    layeredCircuit C;
    vector<vector<F>> val;        // the output of each gate

    // ---- additions of the B200 build (not in the reference interface) ----------------------------------------------
    // auxiliary-input recipes per layer, written by neuralNetwork::create next to C and val (triples, see zk_circuit_aux_ops); uploaded with
    // the circuit so that zk_witness_generate can rebuild the whole witness on the device for a new picture
    std::vector<std::vector<uint32_t>> aux_ops;
    // new picture, same circuit: upload `image` (val[0][0, image.size())), regenerate every auxiliary input and layer on the device and make
    // that the current witness (the host copy `val` is NOT updated).  ranges: 2 x C.size values (zk_witness_generate).  The circuit and one
    // complete witness are uploaded first if they are not there yet.
    void generateWitnessOnDevice(const vector<F> &image, std::vector<uint64_t> &ranges, const std::vector<uint8_t> *want_range = nullptr);
    vector<F> readLayer(u32 layer, size_t n);   // values of a (short) layer as they stand on the device
    // CUDA device used by this prover (default: env ZKCNN_DEVICE or 0).  Call before init().
    void setDevice(int device) { device_ = device; }
    // record every prover->verifier message in SURVEY.md App. A order (nullptr = off)
    void setTranscript(zkcnn_b200::Transcript *t) { transcript_ = t; }
    // init() uploads C once; call this if C was rebuilt and must be uploaded again
    // (a rebuilt circuit also invalidates the witness on the device and any prefetched copy of it)
    void invalidateCircuit() { joinPrefetch(); circuit_uploaded_ = false; witness_uploaded_ = false; prefetch_pending_ = false; }
    // seconds spent uploading circuit / witness in init() (outside the prove timer, like the reference's allocations)
    double uploadTime() const { return upload_timer.elapse_sec(); }
    // true: init() keeps the witness that is already on the device (same val as the previous proof) instead of copying it again
    void setWitnessResident(bool on) { witness_resident_ = on; }
    // start copying val[] for the NEXT proof on a second stream (it overlaps the proof that is running); the next init() adopts it
    void prefetchWitness();
    void setPrefetchNext(bool on) { prefetch_next_ = on; }   // init() starts the next proof's copy once it has its own witness
    // page-lock val[] so that the witness upload is a direct DMA from host memory (call after val is final; undone by the destructor)
    void pinWitness();
    void unpinWitness();
    // test hook: after every Init* call append one line per table pair to `f` with the hashes of the bookkeeping tables, in the format of
    // oracle/harness/record_proxy.hpp (ref_run --dump-dir): per-function parity of K4 / K4b / K5 / K5b / K6 against the reference
    void setTableDump(FILE *f) { table_dump_ = f; }
    uint64_t lastUploadBytes() const { return last_upload_bytes_; }
    uint64_t gpuLaunches() const { return ctx_ ? zk_ctx_launch_count(ctx_) : 0; }
    zk_ctx *context() { return ctx_; }
    timer upload_timer;

private:
    void check(int rc, const char *what) const;
    void uploadCircuit();
    void uploadWitness();
    size_t witnessLength(u32 layer) const;
    // compact witness (built by pinWitness): int64 per element + the few elements that do not fit; what actually crosses PCIe
    void buildCompactWitness();
    std::vector<std::vector<int64_t>> compact_;
    std::vector<std::vector<uint32_t>> wide_idx_;
    std::vector<std::vector<F>> wide_val_;
    bool compact_ready_ = false;
    unsigned msm_digit_bits_ = 0;      // digit width of the commitment's small-multiples path chosen from the witness (0: library default)
    uint64_t compactBytes(u32 layer) const { return compact_[layer].size() * 8 + wide_idx_[layer].size() * 36; }

    zk_ctx *ctx_ = nullptr;
    int device_ = -1;
    bool circuit_uploaded_ = false;
    bool witness_resident_ = false, witness_uploaded_ = false, prefetch_pending_ = false, prefetch_next_ = false;
    uint64_t prefetch_bytes_ = 0;
    std::thread prefetch_thread_;      // paces the copy (zk_witness_layer_prefetch blocks while it does)
    std::string prefetch_error_;
    void joinPrefetch();
    uint64_t last_upload_bytes_ = 0;
    std::vector<const void *> pinned_;
    u64 proof_size = 0;
    zkcnn_b200::Transcript *transcript_ = nullptr;
    FILE *table_dump_ = nullptr;
    int level_ = 0;   // the layer whose sumcheck is running
    void dumpTables(int layer, const char *tag, bool dot, int first_b);
    unique_ptr<hyrax_bls12_381::polyProver> poly_p;

    friend neuralNetwork;
    friend singleConv;
};

#endif //ZKCNN_PROVER_HPP
