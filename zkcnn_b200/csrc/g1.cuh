// BLS12-381 G1 group law (y^2 = x^3 + 4 over Fp), host + device.
//
// Replaces the curve arithmetic that the reference's Hyrax commitment gets from mcl (G1 Jacobian add/dbl:
// mcl/include/mcl/ec.hpp:279,351; mulVec: ec.hpp:1533-1597).  Only the *group element* is comparable with mcl: mcl's
// mulVec is Straus/wNAF and yields different Jacobian coordinates for the same point,
// so every result that crosses the C ABI is normalised to z = 1 (or x = y = z = 0 for the point at infinity, which
// is what mcl's G1::clear() produces).
#pragma once
#include "mont.cuh"

namespace zk {

struct g1_aff_t {  // affine; (0,0) encodes the point at infinity ((0,0) is not on the curve)
    fp_t x, y;
    ZK_HD __forceinline__ bool is_inf() const { return x.is_zero() && y.is_zero(); }
    static ZK_HD __forceinline__ g1_aff_t inf() { return {fp_t::zero(), fp_t::zero()}; }
    ZK_HD __forceinline__ g1_aff_t neg() const { return {x, -y}; }
};

struct g1_jac_t {  // Jacobian (X/Z^2, Y/Z^3); z == 0 encodes infinity.  Same 144-byte layout as mcl's G1.
    fp_t x, y, z;
    ZK_HD __forceinline__ bool is_inf() const { return z.is_zero(); }
    static ZK_HD __forceinline__ g1_jac_t inf() { return {fp_t::zero(), fp_t::zero(), fp_t::zero()}; }
    static ZK_HD __forceinline__ g1_jac_t from_affine(const g1_aff_t &a) {
        if (a.is_inf()) return inf();
        return {a.x, a.y, fp_t::one()};
    }
};
static_assert(sizeof(g1_jac_t) == 144, "G1 must match mcl's 144-byte layout");
static_assert(sizeof(g1_aff_t) == 96, "affine G1 is 96 bytes");

// dbl-2009-l (a = 0): 2M + 5S
ZK_HD inline g1_jac_t g1_dbl(const g1_jac_t &p) {
    if (p.is_inf()) return p;
    fp_t a = p.x.sqr();
    fp_t b = p.y.sqr();
    fp_t c = b.sqr();
    fp_t d = (p.x + b).sqr() - a - c;
    d = d + d;
    fp_t e = a + a + a;
    fp_t f = e.sqr();
    g1_jac_t r;
    r.x = f - (d + d);
    fp_t c8 = c + c;
    c8 = c8 + c8;
    c8 = c8 + c8;
    r.y = e * (d - r.x) - c8;
    r.z = p.y * p.z;
    r.z = r.z + r.z;
    return r;
}

// madd-2007-bl: Jacobian += affine, 7M + 4S, with the exceptional cases handled
ZK_HD inline g1_jac_t g1_add_mixed(const g1_jac_t &p, const g1_aff_t &q) {
    if (q.is_inf()) return p;
    if (p.is_inf()) return g1_jac_t::from_affine(q);
    fp_t z1z1 = p.z.sqr();
    fp_t u2 = q.x * z1z1;
    fp_t s2 = q.y * p.z * z1z1;
    fp_t h = u2 - p.x;
    fp_t rr = s2 - p.y;
    if (h.is_zero()) {
        if (rr.is_zero()) return g1_dbl(p);
        return g1_jac_t::inf();
    }
    rr = rr + rr;
    fp_t hh = h.sqr();
    fp_t i = hh + hh;
    i = i + i;
    fp_t j = h * i;
    fp_t v = p.x * i;
    g1_jac_t r;
    r.x = rr.sqr() - j - (v + v);
    fp_t yj = p.y * j;
    r.y = rr * (v - r.x) - (yj + yj);
    r.z = (p.z + h).sqr() - z1z1 - hh;
    return r;
}

// add-2007-bl: Jacobian + Jacobian, 11M + 5S
ZK_HD inline g1_jac_t g1_add(const g1_jac_t &p, const g1_jac_t &q) {
    if (p.is_inf()) return q;
    if (q.is_inf()) return p;
    fp_t z1z1 = p.z.sqr();
    fp_t z2z2 = q.z.sqr();
    fp_t u1 = p.x * z2z2;
    fp_t u2 = q.x * z1z1;
    fp_t s1 = p.y * q.z * z2z2;
    fp_t s2 = q.y * p.z * z1z1;
    fp_t h = u2 - u1;
    fp_t rr = s2 - s1;
    if (h.is_zero()) {
        if (rr.is_zero()) return g1_dbl(p);
        return g1_jac_t::inf();
    }
    rr = rr + rr;
    fp_t i = (h + h).sqr();
    fp_t j = h * i;
    fp_t v = u1 * i;
    g1_jac_t r;
    r.x = rr.sqr() - j - (v + v);
    fp_t sj = s1 * j;
    r.y = rr * (v - r.x) - (sj + sj);
    r.z = ((p.z + q.z).sqr() - z1z1 - z2z2) * h;
    return r;
}

ZK_HD inline g1_aff_t g1_to_affine(const g1_jac_t &p) {
    if (p.is_inf()) return g1_aff_t::inf();
    fp_t zi = p.z.inverse();
    fp_t zi2 = zi.sqr();
    return {p.x * zi2, p.y * zi2 * zi};
}

// z = 1 form used on the C ABI (infinity -> all-zero, like mcl's clear())
ZK_HD inline g1_jac_t g1_normalize(const g1_jac_t &p) { return g1_jac_t::from_affine(g1_to_affine(p)); }

// k * P for a canonical (non-Montgomery) little-endian 8-limb scalar; plain double-and-add, MSB first
ZK_HD inline g1_jac_t g1_mul_canonical(const g1_jac_t &p, const uint32_t *k) {
    g1_jac_t acc = g1_jac_t::inf();
    bool started = false;
    for (int i = 7; i >= 0; --i)
        for (int b = 31; b >= 0; --b) {
            if (started) acc = g1_dbl(acc);
            if ((k[i] >> b) & 1) {
                acc = g1_add(acc, p);
                started = true;
            }
        }
    return acc;
}

// same group element?  (cross-multiplied comparison, no inversion)
ZK_HD inline bool g1_equal(const g1_jac_t &p, const g1_jac_t &q) {
    if (p.is_inf() || q.is_inf()) return p.is_inf() && q.is_inf();
    fp_t z1z1 = p.z.sqr(), z2z2 = q.z.sqr();
    if (p.x * z2z2 != q.x * z1z1) return false;
    return p.y * q.z * z2z2 == q.y * p.z * z1z1;
}

}  // namespace zk
