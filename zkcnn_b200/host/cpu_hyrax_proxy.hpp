// BASELINE config 2 ("sumcheck fold + MLE tables on the GPU, Hyrax MSM still on the CPU"): the drop-in build with
// ZKCNN_DROPIN_CPU_HYRAX keeps the REFERENCE's own CPU polyProver (3rd/hyrax-bls12-381/src/polyProver.cpp, compiled under the
// moved name ref_polyProver, see oracle/Makefile) and puts this forwarder in front of it, so that its messages (row
// commitments, lcomm / rcomm / ly / ry of every bullet round, the final opening) land in the same canonical transcript as
// the GPU build's and the two configurations can be compared byte for byte.
#pragma once
#define polyProver ref_polyProver
#include <hyrax-bls12-381/src/polyProver.hpp>   // the reference class, under its moved name (its include guard is now set)
#undef polyProver
#include "transcript.hpp"

namespace hyrax_bls12_381 {
class polyProver {   // public interface of 3rd/hyrax-bls12-381/src/polyProver.hpp:20-41
public:
    polyProver(const vector<Fr> &Z, const vector<G1> &gens, zkcnn_b200::Transcript *tr = nullptr) : impl_(Z, gens), tr_(tr) {}
    vector<G1> commit() {
        vector<G1> c = impl_.commit();
        if (tr_) for (auto &p : c) tr_->put_g1(reinterpret_cast<const uint64_t *>(&p));
        return c;
    }
    Fr evaluate(const vector<Fr> &x) { return impl_.evaluate(x); }
    double getPT() const { return impl_.getPT(); }
    double getPS() const { return impl_.getPS(); }
    void initBulletProve(const vector<Fr> &lx, const vector<Fr> &rx) { impl_.initBulletProve(lx, rx); }
    void bulletProve(G1 &lcomm, G1 &rcomm, Fr &ly, Fr &ry) {
        impl_.bulletProve(lcomm, rcomm, ly, ry);
        if (tr_) {
            tr_->put_g1(reinterpret_cast<const uint64_t *>(&lcomm));
            tr_->put_g1(reinterpret_cast<const uint64_t *>(&rcomm));
            tr_->put_fr(reinterpret_cast<const uint64_t *>(&ly));
            tr_->put_fr(reinterpret_cast<const uint64_t *>(&ry));
        }
    }
    void bulletUpdate(const Fr &randomness) { impl_.bulletUpdate(randomness); }
    Fr bulletOpen() {
        Fr y = impl_.bulletOpen();
        if (tr_) tr_->put_fr(reinterpret_cast<const uint64_t *>(&y));
        return y;
    }
    const vector<G1> &getGens() const { return impl_.getGens(); }
private:
    ref_polyProver impl_;
    zkcnn_b200::Transcript *tr_;
};
}  // namespace hyrax_bls12_381
