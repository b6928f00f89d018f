// GKR verifier / protocol driver of the stand-alone build: draws every challenge, calls the prover once per sumcheck
// round and checks the round sums, the layer predicates and the Hyrax opening.  Same public interface and the same
// challenge order as the reference's verifier (src/verifier.hpp:11-47, src/verifier.cpp), so a seeded
// run reproduces the reference's transcript.  It is the CALLER of the hot path, not part of it: its heavy loops
// (gate predicates, generator folding) run on the host or through the C-ABI primitives and can be switched off when
// only the prover is being measured.
#pragma once
#include "prover.hpp"

namespace hyrax_bls12_381 {
class polyVerifier {   // 3rd/hyrax-bls12-381/src/polyVerifier.hpp:14-38
public:
    polyVerifier(polyProver &_p, const vector<G1> &_gens, zk_ctx *ctx, bool check_points);
    bool verify(const vector<Fr> &_x, const Fr &RZL);
    double getVT() { return vt.elapse_sec() - p.getPT(); }
    // draw the randomness of every opening round first and let the prover run all rounds in one device pass (polyProver::bulletProveAll);
    // the draws come from the same stream in the same order, so the transcript is the one of the round-by-round exchange.  Off in
    // Fiat-Shamir mode, where a round's randomness depends on the round's message.
    bool batchRounds = false;
private:
    bool bulletVerify(vector<G1> g, vector<Fr> t, G1 comm, Fr y);
    polyProver &p;
    zk_ctx *ctx;
    bool check_points;
    vector<Fr> x, lx, rx;
    vector<G1> gens;
    vector<G1> comm_Z;
    G1 comm_RZ;
    timer vt;
};
vector<Fr> expand(const vector<Fr> &v);                                     // hyrax/src/utils.cpp:29-62
void split(vector<Fr> &L, vector<Fr> &R, const vector<Fr> &r);              // hyrax/src/utils.cpp:20-27
}  // namespace hyrax_bls12_381

class verifier {
public:
    prover *p;
    const layeredCircuit &C;

    verifier(prover *pr, const layeredCircuit &cir);
    bool verify();

    timer total_timer, total_slow_timer;
    double verifierTime() const { return total_timer.elapse_sec(); }
    double verifierSlowTime() const { return total_slow_timer.elapse_sec(); }

    // ---- additions of the B200 build ----------------------------------------------------------------------------------
    // false: keep the round-sum checks but skip the verifier-side gate predicates, the Liu "gr" recomputation and the
    // G1 checks of the opening (prover-only measurements; the challenge order is unchanged)
    bool checkPredicates = true;
    // false (reference behaviour): generators are getG1basePoint() * random with the base point cleared by initPairing,
    // i.e. all infinity (src/verifier.cpp:125, mcl bn.hpp:924).  true: the standard G1 generator is used as base point.
    bool realGenerators = false;
    // one device call per phase (prover::sumcheckUpdateAll) instead of one per round; the messages and their order are the same
    bool batchRounds = true;
    // true: the wiring predicates (betaInitPhase1/2, predicatePhase1/2) and the input-layer term gr run on the device through the
    // zk_vtab_* / zk_verifier_* entry points (same kernels as the prover's Init* passes, the verifier's own challenges); false: on host threads
    bool devicePredicates = true;
    // Fiat-Shamir mode (the active ChallengeStream derives every challenge from the transcript so far): a round's challenge is
    // drawn AFTER the round's message instead of before the phase (src/verifier.cpp:156-160 draws them up front, which is only
    // sound for an interactive verifier), so the rounds go one by one
    bool fiatShamir = false;
    // Hyrax generators are public parameters; when set, they are reused instead of redrawn (the challenge stream is
    // still advanced as if they had been drawn)
    const vector<G> *fixedGenerators = nullptr;
    vector<G> generators;      // the generators used by the last verify()
    double polyVT = 0;

private:
    vector<vector<F>> r_u, r_v;
    vector<F> final_claim_u0, final_claim_v0;
    bool verifyInnerLayers();
    bool verifyFirstLayer();
    bool verifyInput();

    vector<F> beta_g, beta_u, beta_v, beta_gs;
    void betaInitPhase1(u8 depth, const F &alpha, const F &beta, const vector<F> &r_0, const vector<F> &r_1, const F &relu_rou);
    void betaInitPhase2(u8 depth);
    F uni_value[2];
    F bin_value[3];
    void predicatePhase1(u8 layer_id);
    void predicatePhase2(u8 layer_id);
    void predicatesOnDevice(u8 depth, const F &alpha, const F &beta, const F &relu_rou);   // both phases: fills uni_value / bin_value
    F inputPredicateOnDevice(const vector<F> &sig_u, const vector<F> &sig_v);              // gr of verifyFirstLayer
    F getFinalValue(const F &claim_u0, const F &claim_u1, const F &claim_v0, const F &claim_v1);

    F eval_in;
    unique_ptr<hyrax_bls12_381::polyVerifier> poly_v;
};

// host eq-table / phi-table builders used by the verifier (src/utils.cpp:61-103,147-180)
void initBetaTable(vector<F> &beta_g, u8 gLength, const F *r_0, const F *r_1, const F &alpha, const F &beta);
void initBetaTable(vector<F> &beta_g, u8 gLength, const F *r, const F &init);
void phiGInit(vector<F> &phi_g, const F *rx, const F &scale, int n, bool isIFFT);
F getRootOfUnit(int n);   // neuralNetwork.cpp
