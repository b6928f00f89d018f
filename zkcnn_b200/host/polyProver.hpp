// Host shim with the public interface of the reference's Hyrax prover,
//   class hyrax_bls12_381::polyProver   (3rd/hyrax-bls12-381/src/polyProver.hpp:18-57),
// forwarding every call through the C ABI (include/zkcnn_b200.h) to the CUDA kernels.
//
// Two build modes:
//   ZKCNN_DROPIN  : compiled inside the reference tree; Fr / G1 are mcl's types and this header takes the place of
//                   the reference's polyProver.hpp (same include guard), so polyVerifier.cpp builds against it unchanged.
//   (default)     : stand-alone; Fr / G1 come from zk_types.hpp (layout-identical value types).
#ifndef HYRAX_P224_POLYPROVER_HPP
#define HYRAX_P224_POLYPROVER_HPP

#ifdef ZKCNN_DROPIN
#include <hyrax-bls12-381/src/timer.hpp>     // reference: 3rd/hyrax-bls12-381/src
#include <hyrax-bls12-381/src/typedef.hpp>
#include <hyrax-bls12-381/src/utils.hpp>
#include <mcl/bls12_381.hpp>
using namespace mcl::bn;
#else
#include "zk_types.hpp"
#endif
#include <vector>
#include "../../include/zkcnn_b200.h"
#include "transcript.hpp"

namespace hyrax_bls12_381 {
using std::vector;

class polyProver {
public:
    // polyProver.cpp:12-17 -- stand-alone polynomial: uploads Z and the generators to a private device context
    polyProver(const vector<Fr> &_Z, const vector<G1> &_gens);
    // used by prover::commitInput: Z is prover::val[0], already resident on `ctx`'s device (not copied again)
    polyProver(zk_ctx *ctx, const vector<G1> &_gens, unsigned char bit_length, zkcnn_b200::Transcript *tr);
    ~polyProver();
    polyProver(const polyProver &) = delete;
    polyProver &operator=(const polyProver &) = delete;

    vector<G1> commit();
    Fr evaluate(const vector<Fr> &x);
    double getPT() const;
    double getPS() const;
    void initBulletProve(const vector<Fr> &_lx, const vector<Fr> &_rx);
    void bulletProve(G1 &lcomm, G1 &rcomm, Fr &ly, Fr &ry);
    void bulletUpdate(const Fr &randomness);
    Fr bulletOpen();
    const vector<G1> &getGens() const;
    // Not in the reference: every remaining round in one device pass (zk_poly_bullet_prove_all) for a verifier that has drawn the
    // randomness of all rounds beforehand.  The messages are kept here: the bulletProve calls that follow hand them out in order and
    // bulletUpdate only checks that the verifier uses the randomness it announced.
    void bulletProveAll(const vector<Fr> &randomness);

private:
    void check(int rc, const char *what) const;
    zk_ctx *ctx_;
    bool own_ctx_;
    vector<G1> gens;
    unsigned char bit_length;
    timer pt;
    unsigned long long ps;
    zkcnn_b200::Transcript *tr_;
    struct round_msg { G1 lcomm, rcomm; Fr ly, ry, randomness; };
    vector<round_msg> ahead_;      // rounds computed by bulletProveAll and not yet handed out
    size_t ahead_next_ = 0;
    bool ahead_update_due_ = false;
};
}  // namespace hyrax_bls12_381

#endif  // HYRAX_P224_POLYPROVER_HPP
