// TEST INFRASTRUCTURE (oracle harness).  Deterministic challenge stream for the reference verifier.
//
// The reference draws every verifier challenge with Fr::setByCSPRNG() (src/verifier.cpp:124,139,157,159,207,
// 249,252,275,277,279; hyrax/src/polyVerifier.cpp:47).  mcl lets the byte source be replaced through
// mcl::fp::RandGen::setRandFunc (mcl/include/mcl/randgen.hpp:141-149).  We install a SplitMix64 stream:
// byte k of the stream is byte (k mod 8) (little endian) of the (k/8)-th SplitMix64 output for `seed`.
// zkcnn_b200/host/challenge_stream.hpp implements the same stream for the stand-alone verifier.
#pragma once
#include <cstdint>
#include <cstring>
#include <mcl/bls12_381.hpp>

struct SeededStream {
    uint64_t state;
    uint64_t calls = 0;
    explicit SeededStream(uint64_t seed) : state(seed) {}
    uint64_t next() {
        uint64_t z = (state += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    static uint32_t read(void *self, void *buf, uint32_t n) {
        auto *s = static_cast<SeededStream *>(self);
        auto *out = static_cast<uint8_t *>(buf);
        // whole words only: mcl always asks for 32 bytes per Fr
        for (uint32_t i = 0; i < n; i += 8) {
            uint64_t w = s->next();
            uint32_t m = n - i < 8 ? n - i : 8;
            memcpy(out + i, &w, m);
        }
        ++s->calls;
        return n;
    }
    void install() { mcl::fp::RandGen::setRandFunc(this, &SeededStream::read); }
};
