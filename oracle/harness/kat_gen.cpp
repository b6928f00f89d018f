// TEST INFRASTRUCTURE (oracle harness).  Known-answer vectors minted from the UNMODIFIED reference: mcl's Fr / G1
// arithmetic and mulVec, the reference's initBetaTable / phiGInit / getRootOfUnit (src/utils.cpp) and the reference's
// Hyrax polyProver (compiled under its moved name, see oracle/Makefile).  Output: JSON on stdout -> tests/golden/kat.json
// (committed, regenerate with `make -C oracle golden`).  All values are hex strings of canonical integers; "raw" entries
// are the in-memory 64-bit words (Montgomery form) that cross the C ABI.
#define polyProver ref_polyProver
#define prover ref_prover
#include <hyrax-bls12-381/src/polyProver.hpp>
#include <utils.hpp>
#include "seeded_rng.hpp"
#include <cstdio>
#include <string>
#include <vector>

using namespace mcl::bn;
using std::vector;
using std::string;

static SeededStream *g_rng;
static Fr rnd() { Fr x; x.setByCSPRNG(); return x; }
static string hx(const Fr &x) { return "\"" + x.getStr(16) + "\""; }
static string hx(const Fp &x) { return "\"" + x.getStr(16) + "\""; }
static string raw(const void *p, size_t words) {
    string s = "[";
    const uint64_t *w = static_cast<const uint64_t *>(p);
    char b[32];
    for (size_t i = 0; i < words; ++i) { snprintf(b, sizeof b, "%s\"%016lx\"", i ? "," : "", w[i]); s += b; }
    return s + "]";
}
static string pt(G1 p) {
    if (p.isZero()) return "null";
    p.normalize();
    return "[" + hx(p.x) + "," + hx(p.y) + "]";
}
static string arr(const vector<Fr> &v) {
    string s = "[";
    for (size_t i = 0; i < v.size(); ++i) s += (i ? "," : "") + hx(v[i]);
    return s + "]";
}
static string parr(const vector<G1> &v) {
    string s = "[";
    for (size_t i = 0; i < v.size(); ++i) s += (i ? "," : "") + pt(v[i]);
    return s + "]";
}
static Fr small_signed() {   // witness-like scalar: 0, 1, or |x| < 2^9 with random sign (stored as r - |x| when negative)
    uint64_t u = g_rng->next();
    int kind = u % 10;
    if (kind < 3) return Fr(0);
    if (kind < 4) return Fr(1);
    int64_t v = (int64_t) ((u >> 8) % 511) - 255;
    return Fr(v);
}

int main() {
    initPairing(mcl::BLS12_381);
    install_real_base_point();
    SeededStream rng(20211115);
    rng.install();
    g_rng = &rng;
    const G1 gen = getG1basePoint();

    printf("{\n");
    // ---- layout pins
    { Fr one(1), two(2), m1(-1);
      printf("\"raw_fr_1\": %s,\n\"raw_fr_2\": %s,\n\"raw_fr_m1\": %s,\n", raw(&one, 4).c_str(), raw(&two, 4).c_str(), raw(&m1, 4).c_str());
      printf("\"raw_g1_gen\": %s,\n\"g1_gen\": %s,\n", raw(&gen, 18).c_str(), pt(gen).c_str()); }
    // ---- Fr arithmetic
    printf("\"fr\": [\n");
    for (int i = 0; i < 20; ++i) {
        Fr a = rnd(), b = rnd();
        if (i == 16) a = 0;
        if (i == 17) { a = 1; b = -1; }
        if (i == 18) { a = -1; b = -1; }
        if (i == 19) { a = small_signed(); b = small_signed(); }
        Fr ia; if (a.isZero()) ia = 0; else Fr::inv(ia, a);
        printf(" {\"a\": %s, \"b\": %s, \"raw_a\": %s, \"add\": %s, \"sub\": %s, \"mul\": %s, \"raw_mul\": %s, \"neg\": %s, \"inv\": %s, \"is_neg\": %d}%s\n",
               hx(a).c_str(), hx(b).c_str(), raw(&a, 4).c_str(), hx(a + b).c_str(), hx(a - b).c_str(), hx(a * b).c_str(),
               [&] { Fr c = a * b; return raw(&c, 4); }().c_str(), hx(-a).c_str(), hx(ia).c_str(), (int) a.isNegative(), i == 19 ? "" : ",");
    }
    printf("],\n");
    // ---- roots of unity (src/utils.cpp:224-232)
    { vector<Fr> r; for (int n = 1; n <= 14; ++n) r.push_back(getRootOfUnit(n)); printf("\"root_of_unity_1_14\": %s,\n", arr(r).c_str()); }
    // ---- eq tables (src/utils.cpp:147-180)
    printf("\"beta\": [\n");
    for (int bits = 0; bits <= 6; ++bits) {
        vector<Fr> r0(bits), r1(bits);
        for (auto &x : r0) x = rnd();
        for (auto &x : r1) x = rnd();
        Fr init = rnd(), alpha = rnd(), beta = bits == 3 ? Fr(0) : rnd();
        vector<Fr> t4(1ULL << bits), t6(1ULL << bits);
        initBetaTable(t4, bits, r0.begin(), init);
        initBetaTable(t6, bits, r0.begin(), r1.begin(), alpha, beta);
        printf(" {\"bits\": %d, \"r0\": %s, \"r1\": %s, \"init\": %s, \"alpha\": %s, \"beta\": %s, \"table4\": %s, \"table6\": %s}%s\n", bits,
               arr(r0).c_str(), arr(r1).c_str(), hx(init).c_str(), hx(alpha).c_str(), hx(beta).c_str(), arr(t4).c_str(), arr(t6).c_str(),
               bits == 6 ? "" : ",");
    }
    printf("],\n");
    // ---- phi tables (src/utils.cpp:61-103)
    printf("\"phi\": [\n");
    for (int k = 0; k < 6; ++k) {
        const int n = 2 + k / 2, ifft = k & 1;
        vector<Fr> rx(n);
        for (auto &x : rx) x = rnd();
        Fr scale = ifft ? rnd() : Fr(1);
        vector<Fr> phi(1ULL << n);
        phiGInit(phi, rx.begin(), scale, n, ifft);
        if (!ifft) phi.resize(1ULL << (n - 1));
        printf(" {\"n\": %d, \"ifft\": %d, \"rx\": %s, \"scale\": %s, \"table\": %s}%s\n", n, ifft, arr(rx).c_str(), hx(scale).c_str(),
               arr(phi).c_str(), k == 5 ? "" : ",");
    }
    printf("],\n");
    // ---- G1: add / dbl / scalar mul / mulVec (mcl ec.hpp:279,351,1570-1597; sizes of mcl/test/common_test.hpp:17-44)
    { G1 Pp = gen * rnd(), Q = gen * rnd(); Fr k = rnd(); G1 d; G1::dbl(d, Pp); G1 O; O.clear();
      printf("\"g1\": {\"P\": %s, \"Q\": %s, \"k\": %s, \"add\": %s, \"dbl\": %s, \"mul\": %s, \"P_plus_negP\": %s, \"P_plus_O\": %s},\n", pt(Pp).c_str(),
             pt(Q).c_str(), hx(k).c_str(), pt(Pp + Q).c_str(), pt(d).c_str(), pt(Pp * k).c_str(), pt(Pp + (-Pp)).c_str(), pt(Pp + O).c_str()); }
    printf("\"mulvec\": [\n");
    const int sizes[] = {1, 2, 3, 5, 16, 33, 70};
    for (int si = 0; si < 7; ++si) {
        const int n = sizes[si];
        vector<G1> pts(n);
        vector<Fr> ks(n);
        for (int i = 0; i < n; ++i) {
            pts[i] = gen * rnd();
            ks[i] = (si & 1) ? small_signed() : rnd();
        }
        if (n >= 5) { pts[2].clear(); ks[3] = 0; pts[4] = pts[0]; }   // infinity among the bases, zero scalar, repeated base
        G1 out;
        G1::mulVec(out, pts.data(), ks.data(), n);
        printf(" {\"n\": %d, \"points\": %s, \"scalars\": %s, \"out\": %s}%s\n", n, parr(pts).c_str(), arr(ks).c_str(), pt(out).c_str(),
               si == 6 ? "" : ",");
    }
    printf("],\n");
    // ---- Hyrax prover (3rd/hyrax-bls12-381/src/polyProver.cpp) on a 2^6 polynomial, 8 generators
    {
        const int bl = 6, lbl = 3, rbl = 3;
        vector<Fr> Z(1 << bl);
        for (size_t i = 0; i < Z.size(); ++i) Z[i] = i % 7 == 0 ? rnd() : small_signed();
        vector<G1> gens(1 << lbl);
        for (auto &g : gens) g = gen * rnd();
        hyrax_bls12_381::ref_polyProver hp(Z, gens);
        vector<G1> comm = hp.commit();
        vector<Fr> x(bl);
        for (auto &v : x) v = rnd();
        Fr ev = hp.evaluate(x);
        vector<Fr> lx(x.begin(), x.begin() + lbl), rx(x.begin() + lbl, x.end());
        hp.initBulletProve(lx, rx);
        printf("\"hyrax\": {\"Z\": %s, \"gens\": %s, \"commit\": %s, \"x\": %s, \"evaluate\": %s, \"rounds\": [\n", arr(Z).c_str(), parr(gens).c_str(),
               parr(comm).c_str(), arr(x).c_str(), hx(ev).c_str());
        for (int j = 0; j < lbl; ++j) {
            G1 lc, rc; Fr ly, ry;
            hp.bulletProve(lc, rc, ly, ry);
            Fr rho = rnd();
            hp.bulletUpdate(rho);
            printf("  {\"lcomm\": %s, \"rcomm\": %s, \"ly\": %s, \"ry\": %s, \"randomness\": %s}%s\n", pt(lc).c_str(), pt(rc).c_str(), hx(ly).c_str(),
                   hx(ry).c_str(), hx(rho).c_str(), j == lbl - 1 ? "" : ",");
        }
        printf(" ], \"open\": %s, \"rsize\": %d, \"rbl\": %d}\n", hx(hp.bulletOpen()).c_str(), 1 << rbl, rbl);
    }
    printf("}\n");
    return 0;
}
