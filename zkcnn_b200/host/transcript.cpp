#define ZK_HOST_ONLY
#include "transcript.hpp"
#include <algorithm>
#include <thread>
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

void Transcript::put_g1_many(const uint64_t *w, size_t n) {
    const size_t base = bytes.size();
    bytes.resize(base + n * 96);
    uint8_t *out = bytes.data() + base;
    auto work = [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) g1_words_to_affine_le(w + 18 * i, out + 96 * i);
    };
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t nt = n >= 1024 ? std::min<size_t>(4, hw ? hw : 1) : 1;
    if (nt <= 1) work(0, n);
    else {
        std::vector<std::thread> th;
        const size_t per = (n + nt - 1) / nt;
        for (size_t t = 1; t < nt; ++t) th.emplace_back(work, std::min(n, t * per), std::min(n, (t + 1) * per));
        work(0, std::min(n, per));
        for (auto &x : th) x.join();
    }
    n_g1 += n;
}

}  // namespace zkcnn_b200
