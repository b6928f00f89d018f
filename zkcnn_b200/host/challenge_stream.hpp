// Verifier challenge source of the stand-alone build.  The reference draws challenges from mcl's CSPRNG
// (Fr::setByCSPRNG, src/verifier.cpp:124,139,...).  Three sources here:
//   SEEDED      reproducible transcripts: the SplitMix64 stream that oracle/harness/seeded_rng.hpp installs into the reference:
//               byte k of the stream is byte (k mod 8) (little endian) of the (k/8)-th SplitMix64 output for `seed`
//   OS_CSPRNG   /dev/urandom, the reference's behaviour
//   FIAT_SHAMIR non-interactive: challenge k = SplitMix64 stream seeded with FNV-1a-64(seed || k || transcript bytes so far),
//               i.e. every challenge binds all prover messages that precede it in SURVEY.md App. A order.  (SURVEY section 8 f-4;
//               a deployment would swap the 64-bit FNV for a cryptographic hash -- the plumbing is what this mode provides.)
// The active stream is per THREAD: several proofs may run in one process, one per thread.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include "transcript.hpp"

namespace zkcnn_b200 {

struct ChallengeStream {
    enum Mode { SEEDED = 0, OS_CSPRNG = 1, FIAT_SHAMIR = 2 };
    uint64_t state, seed;
    uint64_t calls = 0;
    Mode mode;
    const Transcript *transcript = nullptr;   // FIAT_SHAMIR: the messages sent so far
    uint64_t fs_hash = 0xcbf29ce484222325ULL;
    size_t fs_pos = 0;                         // transcript bytes already absorbed into fs_hash
    explicit ChallengeStream(uint64_t seed_, Mode m = SEEDED) : state(seed_), seed(seed_), mode(m) {}
    uint64_t next() {
        uint64_t z = (state += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    void read(uint8_t *out, size_t n);
};

// RAII: route Fr::setByCSPRNG() of the calling thread to this stream for the lifetime of the object
struct ScopedChallengeStream {
    ChallengeStream stream;
    ChallengeStream *saved;
    explicit ScopedChallengeStream(uint64_t seed, ChallengeStream::Mode mode = ChallengeStream::SEEDED);
    ~ScopedChallengeStream();
};

}  // namespace zkcnn_b200
