// See polyProver.hpp.  Mirrors 3rd/hyrax-bls12-381/src/polyProver.cpp; all arithmetic is behind the C ABI.
#include "polyProver.hpp"
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>

#define ZK_G1_BYTES 48   /* G1_SIZE in polyProver.cpp:8 is Fp::getByteSize() */
#define ZK_FR_BYTES 32

namespace hyrax_bls12_381 {

static_assert(sizeof(Fr) == 32, "Fr must be 4 x 64-bit Montgomery limbs");
static_assert(sizeof(G1) == 144, "G1 must be Jacobian (x, y, z) of 6 x 64-bit limbs");

static inline const uint64_t *w(const Fr &x) { return reinterpret_cast<const uint64_t *>(&x); }
static inline uint64_t *w(Fr &x) { return reinterpret_cast<uint64_t *>(&x); }
static inline const uint64_t *w(const G1 &x) { return reinterpret_cast<const uint64_t *>(&x); }
static inline uint64_t *w(G1 &x) { return reinterpret_cast<uint64_t *>(&x); }

void polyProver::check(int rc, const char *what) const {
    if (rc != 0) throw std::runtime_error(std::string("zkcnn_b200: ") + what + ": " + zk_last_error());
}

static unsigned char log2_ceil(unsigned long long x) {   // myLog2, hyrax/src/utils.cpp:11-13
    unsigned char r = 0;
    while ((1ULL << r) < x) ++r;
    return r;
}

polyProver::polyProver(const vector<Fr> &_Z, const vector<G1> &_gens)
    : ctx_(nullptr), own_ctx_(true), gens(_gens), ps(0), tr_(nullptr) {
    bit_length = log2_ceil(_Z.size());
    ctx_ = zk_ctx_create(0);
    if (!ctx_) throw std::runtime_error(std::string("zkcnn_b200: cannot create a device context: ") + zk_last_error());
    check(zk_poly_create(ctx_, w(_Z[0]), _Z.size(), gens.empty() ? nullptr : w(gens[0]), (uint32_t) gens.size()), "zk_poly_create");
}

polyProver::polyProver(zk_ctx *ctx, const vector<G1> &_gens, unsigned char bl, zkcnn_b200::Transcript *tr)
    : ctx_(ctx), own_ctx_(false), gens(_gens), bit_length(bl), ps(0), tr_(tr) {
    check(zk_poly_bind_input(ctx_, gens.empty() ? nullptr : w(gens[0]), (uint32_t) gens.size()), "zk_poly_bind_input");
}

polyProver::~polyProver() {
    if (own_ctx_ && ctx_) zk_ctx_destroy(ctx_);
}

vector<G1> polyProver::commit() {   // polyProver.cpp:19-34
    pt.start();
    unsigned char r_bit_length = bit_length >> 1;
    unsigned long long rsize = 1ULL << r_bit_length;
    vector<G1> comm_Z(rsize);
    check(zk_poly_commit(ctx_, w(comm_Z[0]), (uint32_t) rsize), "zk_poly_commit");
    pt.stop();
    ps += ZK_G1_BYTES * comm_Z.size();
    if (tr_ && !comm_Z.empty()) tr_->put_g1_many(w(comm_Z[0]), comm_Z.size());
    return comm_Z;
}

Fr polyProver::evaluate(const vector<Fr> &x) {   // polyProver.cpp:36-42
    Fr res;
    check(zk_poly_evaluate(ctx_, x.empty() ? nullptr : w(x[0]), (uint32_t) x.size(), w(res)), "zk_poly_evaluate");
    return res;
}

double polyProver::getPT() const { return pt.elapse_sec(); }

double polyProver::getPS() const { return ps / 1024.0; }   // KB

void polyProver::initBulletProve(const vector<Fr> &_lx, const vector<Fr> &_rx) {   // polyProver.cpp:52-74
    pt.start();
    check(zk_poly_init_bullet_prove(ctx_, _lx.empty() ? nullptr : w(_lx[0]), (uint32_t) _lx.size(), _rx.empty() ? nullptr : w(_rx[0]),
                                    (uint32_t) _rx.size()),
          "zk_poly_init_bullet_prove");
    pt.stop();
}

void polyProver::bulletProveAll(const vector<Fr> &randomness) {
    if (randomness.empty()) return;
    pt.start();
    const size_t n = randomness.size();
    vector<G1> lc(n), rc(n);
    vector<Fr> ly(n), ry(n);
    check(zk_poly_bullet_prove_all(ctx_, w(randomness[0]), (uint32_t) n, w(lc[0]), w(rc[0]), w(ly[0]), w(ry[0])), "zk_poly_bullet_prove_all");
    pt.stop();
    ahead_.resize(n);
    for (size_t k = 0; k < n; ++k) ahead_[k] = {lc[k], rc[k], ly[k], ry[k], randomness[k]};
    ahead_next_ = 0;
    ahead_update_due_ = false;
}

void polyProver::bulletProve(G1 &lcomm, G1 &rcomm, Fr &ly, Fr &ry) {   // polyProver.cpp:76-96
    if (ahead_next_ < ahead_.size()) {
        if (ahead_update_due_) throw std::logic_error("polyProver: bulletProve twice without bulletUpdate");
        const round_msg &m = ahead_[ahead_next_];
        lcomm = m.lcomm; rcomm = m.rcomm; ly = m.ly; ry = m.ry;
        ahead_update_due_ = true;
    } else {
        pt.start();
        check(zk_poly_bullet_prove(ctx_, w(lcomm), w(rcomm), w(ly), w(ry)), "zk_poly_bullet_prove");
        pt.stop();
    }
    ps += (ZK_G1_BYTES + ZK_FR_BYTES) * 2;
    if (tr_) { tr_->put_g1(w(lcomm)); tr_->put_g1(w(rcomm)); tr_->put_fr(w(ly)); tr_->put_fr(w(ry)); }
}

void polyProver::bulletUpdate(const Fr &randomness) {   // polyProver.cpp:98-109
    if (ahead_update_due_) {   // folded on the device already, with the randomness announced to bulletProveAll
        if (memcmp(&randomness, &ahead_[ahead_next_].randomness, sizeof(Fr)) != 0)
            throw std::logic_error("polyProver: bulletUpdate with another randomness than the one given to bulletProveAll");
        ++ahead_next_;
        ahead_update_due_ = false;
        return;
    }
    pt.start();
    check(zk_poly_bullet_update(ctx_, w(randomness)), "zk_poly_bullet_update");
    pt.stop();
}

Fr polyProver::bulletOpen() {   // polyProver.cpp:111-116
    Fr y;
    check(zk_poly_bullet_open(ctx_, w(y)), "zk_poly_bullet_open");
    ps += ZK_FR_BYTES;
    if (tr_) tr_->put_fr(w(y));
    return y;
}

const vector<G1> &polyProver::getGens() const { return gens; }

}  // namespace hyrax_bls12_381
