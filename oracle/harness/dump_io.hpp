// TEST INFRASTRUCTURE (oracle harness).  Canonical transcript encoding shared by ref_run and the drop-in harness.
//   Fr  -> 32 bytes, little-endian canonical (non-Montgomery) integer   (== mcl Fr::serialize)
//   G1  -> 96 bytes: affine x || y, each 48 bytes little-endian canonical; all zero for the point at infinity
// The byte-identical encoder of the product lives in zkcnn_b200/host/transcript.hpp.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <mcl/bls12_381.hpp>

struct TranscriptSink {
    std::vector<uint8_t> bytes;
    uint64_t n_fr = 0, n_g1 = 0;
    void put(const mcl::bn::Fr &x) {
        uint8_t buf[32];
        size_t n = x.serialize(buf, sizeof(buf));
        if (n != 32) { fprintf(stderr, "Fr serialize failed\n"); abort(); }
        bytes.insert(bytes.end(), buf, buf + 32);
        ++n_fr;
    }
    void put(const mcl::bn::G1 &p) {
        uint8_t buf[96] = {0};
        if (!p.isZero()) {
            mcl::bn::G1 q = p;
            q.normalize();
            // Fp::serialize gives 48 bytes little endian (mcl default ioMode, non-ETH)
            size_t a = q.x.serialize(buf, 48);
            size_t b = q.y.serialize(buf + 48, 48);
            if (a != 48 || b != 48) { fprintf(stderr, "Fp serialize failed\n"); abort(); }
        }
        bytes.insert(bytes.end(), buf, buf + 96);
        ++n_g1;
    }
    bool save(const std::string &path) const {
        FILE *f = fopen(path.c_str(), "wb");
        if (!f) return false;
        fwrite(bytes.data(), 1, bytes.size(), f);
        fclose(f);
        return true;
    }
    uint64_t fnv1a() const {
        uint64_t h = 0xcbf29ce484222325ULL;
        for (uint8_t b : bytes) { h ^= b; h *= 0x100000001b3ULL; }
        return h;
    }
};

inline TranscriptSink &transcript() { static TranscriptSink t; return t; }
