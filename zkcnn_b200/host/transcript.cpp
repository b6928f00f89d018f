#define ZK_HOST_ONLY
#include "transcript.hpp"
#include "../csrc/g1.cuh"

namespace zkcnn_b200 {

void fr_words_to_canonical_le(const uint64_t *w, uint8_t out[32]) {
    zk::fr_t x;
    memcpy(x.v, w, 32);
    uint32_t c[8];
    x.to_canonical(c);
    memcpy(out, c, 32);   // little-endian host
}

void g1_words_to_affine_le(const uint64_t *w, uint8_t out[96]) {
    zk::g1_jac_t p;
    memcpy(&p, w, sizeof p);
    memset(out, 0, 96);
    if (p.is_inf()) return;
    // points returned by the library are already normalised (z = 1): no field inversion on the host for those
    zk::g1_aff_t a = p.z == zk::fp_t::one() ? zk::g1_aff_t{p.x, p.y} : zk::g1_to_affine(p);
    uint32_t c[12];
    a.x.to_canonical(c);
    memcpy(out, c, 48);
    a.y.to_canonical(c);
    memcpy(out + 48, c, 48);
}

}  // namespace zkcnn_b200
