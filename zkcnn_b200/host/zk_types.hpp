// Stand-alone host value types with the names and semantics of the types that cross the reference's prover interface:
//   Fr / G1 (mcl value types as used by zkCNN), timer (hyrax/src/timer.hpp), uniGate / binGate / layer / layeredCircuit
//   (src/circuit.h), linear/quadratic/cubic_poly (src/polynomial.h), F / G aliases (src/global_var.hpp).
// Layouts are identical to the reference's (Fr 32 B Montgomery, G1 144 B Jacobian Montgomery, uniGate 12 B, binGate 16 B),
// so the same bytes cross the C ABI in both build modes.  Host arithmetic uses the portable code of csrc/mont.cuh; it is
// only used for O(1)-per-round verifier work and for circuit construction, never for the prover's tables.
#pragma once
#ifndef ZK_HOST_ONLY
#define ZK_HOST_ONLY
#endif
#include "../csrc/g1.cuh"
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned __int128 u128;
typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned char u8;
typedef __int128 i128;
typedef long long i64;
typedef int i32;
typedef char i8;

using std::vector;
using std::string;
using std::max;
using std::min;

namespace zkcnn_b200 {
// deterministic challenge source (see challenge_stream.hpp); nullptr -> /dev/urandom
struct ChallengeStream;
ChallengeStream *&active_challenge_stream();
void challenge_bytes(uint8_t *out, size_t n);
}  // namespace zkcnn_b200

// ---- Fr ------------------------------------------------------------------------------------------------------------------
class Fr {
public:
    zk::fr_t v;
    Fr() { v = zk::fr_t::zero(); }
    Fr(i64 x) { v = zk::fr_t::from_i64(x); }
    Fr(int x) { v = zk::fr_t::from_i64(x); }
    Fr(u64 x) { v = zk::fr_t::from_u64(x); }
    Fr(u32 x) { v = zk::fr_t::from_u64(x); }
    Fr(long x) { v = zk::fr_t::from_i64(x); }
    Fr(unsigned long x) { v = zk::fr_t::from_u64(x); }
    explicit Fr(const zk::fr_t &x) : v(x) {}
    static const Fr &one() { static const Fr o(zk::fr_t::one()); return o; }
    static void inv(Fr &out, const Fr &in) { out.v = in.v.inverse(); }
    static void neg(Fr &out, const Fr &in) { out.v = -in.v; }
    static size_t getByteSize() { return 32; }
    void clear() { v = zk::fr_t::zero(); }
    bool isZero() const { return v.is_zero(); }
    bool isOne() const { return v == zk::fr_t::one(); }
    bool isNegative() const { return v.is_negative(); }   // mcl/include/mcl/fp.hpp:666-671
    Fr operator+(const Fr &b) const { return Fr(v + b.v); }
    Fr operator-(const Fr &b) const { return Fr(v - b.v); }
    Fr operator*(const Fr &b) const { return Fr(v * b.v); }
    Fr operator-() const { return Fr(-v); }
    Fr &operator+=(const Fr &b) { v = v + b.v; return *this; }
    Fr &operator-=(const Fr &b) { v = v - b.v; return *this; }
    Fr &operator*=(const Fr &b) { v = v * b.v; return *this; }
    bool operator==(const Fr &b) const { return v == b.v; }
    bool operator!=(const Fr &b) const { return v != b.v; }
    // ordering on the canonical value, as mcl's FpT::operator< (used by the quantisation helpers)
    static int cmp(const Fr &a, const Fr &b) {
        uint32_t x[8], y[8];
        a.v.to_canonical(x);
        b.v.to_canonical(y);
        for (int i = 7; i >= 0; --i) {
            if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
        }
        return 0;
    }
    bool operator<(const Fr &b) const { return cmp(*this, b) < 0; }
    bool operator>(const Fr &b) const { return cmp(*this, b) > 0; }
    bool operator<=(const Fr &b) const { return cmp(*this, b) <= 0; }
    bool operator>=(const Fr &b) const { return cmp(*this, b) >= 0; }
    // signed small value (mcl getInt64: values >= (r+1)/2 are negative)
    i64 getInt64() const {
        uint32_t c[8];
        bool n = v.is_negative();
        (n ? -v : v).to_canonical(c);
        for (int i = 2; i < 8; ++i) if (c[i]) throw std::range_error("Fr::getInt64: value does not fit");
        u64 m = (u64) c[0] | ((u64) c[1] << 32);
        return n ? -(i64) m : (i64) m;
    }
    size_t serialize(void *buf, size_t n) const {
        if (n < 32) return 0;
        uint32_t c[8];
        v.to_canonical(c);
        memcpy(buf, c, 32);
        return 32;
    }
    // 32 challenge bytes, little endian, masked to 255 bits and, if still >= r, to 254 bits
    // (mcl setByCSPRNG -> setArrayMask, mcl/include/mcl/fp.hpp:407-421,508-517)
    void setByCSPRNG() {
        uint32_t c[8];
        zkcnn_b200::challenge_bytes(reinterpret_cast<uint8_t *>(c), 32);
        c[7] &= 0x7fffffffu;
        if (zk::fr_t::ge_raw(c, zk::fr_cfg::mod())) c[7] &= 0x3fffffffu;
        v = zk::fr_t::from_canonical(c);
    }
    std::string hex() const {
        uint32_t c[8];
        v.to_canonical(c);
        char b[65];
        for (int i = 0; i < 8; ++i) snprintf(b + 8 * i, 9, "%08x", c[7 - i]);
        return b;
    }
};
inline std::ostream &operator<<(std::ostream &os, const Fr &x) { return os << x.hex(); }
static_assert(sizeof(Fr) == 32, "Fr layout");

// ---- G1 ------------------------------------------------------------------------------------------------------------------
class G1 {
public:
    zk::g1_jac_t v;
    G1() { v = zk::g1_jac_t::inf(); }
    explicit G1(const zk::g1_jac_t &p) : v(p) {}
    void clear() { v = zk::g1_jac_t::inf(); }
    bool isZero() const { return v.is_inf(); }
    void normalize() { v = zk::g1_normalize(v); }
    G1 operator+(const G1 &b) const { return G1(zk::g1_add(v, b.v)); }
    G1 operator*(const Fr &k) const {
        uint32_t c[8];
        k.v.to_canonical(c);
        return G1(zk::g1_mul_canonical(v, c));
    }
    bool operator==(const G1 &b) const { return zk::g1_equal(v, b.v); }
    bool operator!=(const G1 &b) const { return !(*this == b); }
    // the standard BLS12-381 generator (mcl/test/bls12_test.cpp:53-54)
    static G1 generator() {
        zk::g1_jac_t p;
        memcpy(p.x.v, h_g1_gen_x_mont, 48);
        memcpy(p.y.v, h_g1_gen_y_mont, 48);
        p.z = zk::fp_t::one();
        return G1(p);
    }
};
static_assert(sizeof(G1) == 144, "G1 layout");

#define F Fr
#define G G1
#define F_ONE (Fr::one())
#define F_ZERO (Fr(0))
#define F_BYTE_SIZE (Fr::getByteSize())

// ---- timer (hyrax/src/timer.hpp:11-25) --------------------------------------------------------------------------------------
class timer {
public:
    timer() { total_time_sec = 0; status = false; }
    void start() { assert(!status); t0 = std::chrono::high_resolution_clock::now(); status = true; }
    void stop() {
        assert(status);
        total_time_sec += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        status = false;
    }
    void clear() { total_time_sec = 0; status = false; }
    double elapse_sec() const { return total_time_sec; }
private:
    std::chrono::high_resolution_clock::time_point t0;
    double total_time_sec;
    bool status;
};

// ---- round-message polynomials (src/polynomial.h:10-48) ---------------------------------------------------------------------
class linear_poly {
public:
    F a, b;   // a x + b
    linear_poly() {}
    linear_poly(const F &aa, const F &bb) : a(aa), b(bb) {}
    F eval(const F &x) const { return a * x + b; }
};
class quadratic_poly {
public:
    F a, b, c;   // a x^2 + b x + c
    quadratic_poly() {}
    quadratic_poly(const F &aa, const F &bb, const F &cc) : a(aa), b(bb), c(cc) {}
    F eval(const F &x) const { return (a * x + b) * x + c; }
    void clear() { a.clear(); b.clear(); c.clear(); }
};
class cubic_poly {
public:
    F a, b, c, d;   // a x^3 + b x^2 + c x + d
    cubic_poly() {}
    cubic_poly(const F &aa, const F &bb, const F &cc, const F &dd) : a(aa), b(bb), c(cc), d(dd) {}
    F eval(const F &x) const { return ((a * x + b) * x + c) * x + d; }
    void clear() { a.clear(); b.clear(); c.clear(); d.clear(); }
};

// ---- circuit IR (src/circuit.h:15-88) -----------------------------------------------------------------------------------------
struct uniGate {
    u32 g, u;
    u8 lu, sc;
    uniGate(u32 _g, u32 _u, u8 _lu, u8 _sc) : g(_g), u(_u), lu(_lu), sc(_sc) {}
};
struct binGate {
    u32 g, u, v;
    u8 sc, l;
    binGate(u32 _g, u32 _u, u32 _v, u8 _sc, u8 _l) : g(_g), u(_u), v(_v), sc(_sc), l(_l) {}
    u8 getLayerIdU(u8 layer_id) const { return !l ? 0 : layer_id - 1; }
    u8 getLayerIdV(u8 layer_id) const { return !(l & 1) ? 0 : layer_id - 1; }
};
static_assert(sizeof(uniGate) == 12 && sizeof(binGate) == 16, "gate layouts (SURVEY.md section 8 T)");

enum class layerType {
    INPUT, FFT, IFFT, ADD_BIAS, RELU, Sqr, OPT_AVG_POOL, MAX_POOL, AVG_POOL, DOT_PROD, PADDING, FCONN, NCONV, NCONV_MUL, NCONV_ADD
};

class layer {
public:
    layerType ty;
    u32 size{}, size_u[2]{}, size_v[2]{};
    i8 bit_length_u[2]{}, bit_length_v[2]{}, bit_length{};
    i8 max_bl_u{}, max_bl_v{};
    bool need_phase2;
    u32 zero_start_id;            // rows >= zero_start_id must evaluate to zero (bit-decomposition checks)
    std::vector<uniGate> uni_gates;
    std::vector<binGate> bin_gates;
    vector<u32> ori_id_u, ori_id_v;
    i8 fft_bit_length;
    F scale;                      // IFFT or average pooling

    layer() : ty(layerType::INPUT) {
        bit_length_u[0] = bit_length_v[0] = -1;
        bit_length_u[1] = bit_length_v[1] = -1;
        need_phase2 = false;
        zero_start_id = 0;
        fft_bit_length = -1;
        scale = F_ONE;
    }
    void updateSize() {
        max_bl_u = std::max(bit_length_u[0], bit_length_u[1]);
        max_bl_v = 0;
        if (!need_phase2) return;
        max_bl_v = std::max(bit_length_v[0], bit_length_v[1]);
    }
};

class layeredCircuit {
public:
    vector<layer> circuit;
    u8 size;
    vector<F> two_mul;
    void init(u8 q_bit_size, u8 _layer_sz);
    void initSubset();
};

i8 ceilPow2BitLength(u32 n);
