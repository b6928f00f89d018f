// Verifier challenge source of the stand-alone build.  The reference draws challenges from mcl's CSPRNG
// (Fr::setByCSPRNG, src/verifier.cpp:124,139,...); here the source is either /dev/urandom or, for reproducible
// transcripts, the SplitMix64 stream that oracle/harness/seeded_rng.hpp installs into the reference:
// byte k of the stream is byte (k mod 8) (little endian) of the (k/8)-th SplitMix64 output for `seed`.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>

namespace zkcnn_b200 {

struct ChallengeStream {
    uint64_t state;
    uint64_t calls = 0;
    explicit ChallengeStream(uint64_t seed) : state(seed) {}
    uint64_t next() {
        uint64_t z = (state += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    void read(uint8_t *out, size_t n) {
        for (size_t i = 0; i < n; i += 8) {
            uint64_t w = next();
            size_t m = n - i < 8 ? n - i : 8;
            memcpy(out + i, &w, m);
        }
        ++calls;
    }
};

// RAII: route Fr::setByCSPRNG() to a seeded stream for the lifetime of this object
struct ScopedChallengeStream {
    ChallengeStream stream;
    ChallengeStream *saved;
    explicit ScopedChallengeStream(uint64_t seed);
    ~ScopedChallengeStream();
};

}  // namespace zkcnn_b200
