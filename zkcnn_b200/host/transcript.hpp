// Canonical proof transcript: the ordered prover->verifier messages of SURVEY.md Appendix A.
//   Fr -> 32 bytes little-endian canonical (== mcl Fr::serialize);  G1 -> 96 bytes affine x || y, little-endian
//   canonical, all zero for the point at infinity.
// The reference never serialises a proof (src/prover.cpp:137,152,381 only count bytes); this encoding is what the
// multi-GPU launcher gathers, and what the parity tests compare against the reference's recorded messages.
#pragma once
#define ZK_HOST_ONLY_FIELD 1
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace zkcnn_b200 {

// Field / curve values arrive as raw words in mcl's in-memory form (Montgomery); conversion is done with the
// portable arithmetic of csrc/mont.cuh.
void fr_words_to_canonical_le(const uint64_t *mont_words, uint8_t out[32]);
void g1_words_to_affine_le(const uint64_t *jac_words, uint8_t out[96]);

struct Transcript {
    std::vector<uint8_t> bytes;
    uint64_t n_fr = 0, n_g1 = 0;
    void put_fr(const uint64_t *w) {
        uint8_t b[32];
        fr_words_to_canonical_le(w, b);
        bytes.insert(bytes.end(), b, b + 32);
        ++n_fr;
    }
    void put_g1(const uint64_t *w) {
        uint8_t b[96];
        g1_words_to_affine_le(w, b);
        bytes.insert(bytes.end(), b, b + 96);
        ++n_g1;
    }
    // many points at once (the 2^(bl/2) row commitments): the conversions are independent, a few host threads share them
    void put_g1_many(const uint64_t *w, size_t n);
    uint64_t fnv1a() const {
        uint64_t h = 0xcbf29ce484222325ULL;
        for (uint8_t b : bytes) { h ^= b; h *= 0x100000001b3ULL; }
        return h;
    }
    bool save(const std::string &path) const {
        FILE *f = fopen(path.c_str(), "wb");
        if (!f) return false;
        fwrite(bytes.data(), 1, bytes.size(), f);
        fclose(f);
        return true;
    }
    void clear() { bytes.clear(); n_fr = n_g1 = 0; }
};

}  // namespace zkcnn_b200
