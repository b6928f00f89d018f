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

// The shipped demos multiply getG1basePoint() by a random scalar to make the Hyrax generators (src/verifier.cpp:125),
// but initPairing() clears that base point (mcl/include/mcl/bn.hpp:924), so every generator is the point at infinity
// (SURVEY.md section 0, fact 3).  `--gens real` puts the standard BLS12-381 G1 generator (mcl/test/bls12_test.cpp:53-54)
// into mcl's public static before the verifier runs, which makes the same reference code path produce
// non-degenerate generators.  No reference file is modified.
inline void install_real_base_point() {
    mcl::bn::Fp x, y;
    x.setStr("17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb", 16);
    y.setStr("08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1", 16);
    mcl::bn::G1 g;
    g.set(x, y);
    const_cast<mcl::bn::G1 &>(mcl::bn::getG1basePoint()) = g;
}
