// TEST INFRASTRUCTURE (oracle harness).  Known-answer vectors minted from the UNMODIFIED reference: mcl's Fr / G1
// arithmetic and mulVec, the reference's initBetaTable / phiGInit / getRootOfUnit (src/utils.cpp) and the reference's
// Hyrax polyProver (compiled under its moved name, see oracle/Makefile).  Output: JSON on stdout -> tests/golden/kat.json
// (committed, regenerate with `make -C oracle golden`).  All values are hex strings of canonical integers; "raw" entries
// are the in-memory 64-bit words (Montgomery form) that cross the C ABI.
#define polyProver ref_polyProver
#define prover ref_prover
#include <hyrax-bls12-381/src/polyProver.hpp>
#define private public   // (access only: the fold / cubic / Vres KATs call the reference prover's private round functions on hand-made tables)
#include <prover.hpp>
#undef private
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
        printf(" ], \"open\": %s, \"rsize\": %d, \"rbl\": %d},\n", hx(hp.bulletOpen()).c_str(), 1 << rbl, rbl);
    }
    // ---- K1: prover::sumcheckUpdate / sumcheckUpdateEach (src/prover.cpp:368-383,396-426) on hand-made table pairs: two pairs of
    //      different sizes (the smaller one collapses into add_term on the way), ragged live sizes
    printf("\"fold\": [\n");
    const int fold_cases[][4] = {{5, 19, 3, 5}, {4, 16, -1, 0}, {6, 33, 6, 64}, {3, 5, 1, 2}};   // {bits1, live1, bits0 (-1: absent), live0}
    for (int ci = 0; ci < 4; ++ci) {
        const int b1 = fold_cases[ci][0], l1 = fold_cases[ci][1], b0 = fold_cases[ci][2], l0 = fold_cases[ci][3];
        ref_prover rp;
        vector<Fr> V[2], M[2];
        const int bl[2] = {b0, b1}, lv[2] = {l0, l1};
        for (int b = 0; b < 2; ++b) {
            rp.total[b] = bl[b] >= 0 ? 1u << bl[b] : 0;
            rp.total_size[b] = lv[b];
            rp.V_mult[b].resize(rp.total[b]);
            rp.mult_array[b].resize(rp.total[b]);
            for (u32 i = 0; i < rp.total[b]; ++i) {
                Fr v = (int) i < lv[b] ? (i % 3 ? small_signed() : rnd()) : Fr(0), m = (int) i < lv[b] ? rnd() : Fr(0);
                if ((int) i < lv[b]) { V[b].push_back(v); M[b].push_back(m); }
                rp.V_mult[b][i] = v;
                rp.mult_array[b][i] = m;
            }
        }
        const int rounds = b1 > b0 ? b1 : b0;
        vector<Fr> ch(rounds);
        for (auto &x : ch) x = rnd();
        rp.r_u.assign(2, vector<Fr>(rounds));
        rp.sumcheck_id = 1;
        rp.round = 0;
        rp.add_term.clear();
        printf(" {\"bits\": [%d, %d], \"V0\": %s, \"M0\": %s, \"V1\": %s, \"M1\": %s, \"r\": %s, \"polys\": [", b0, b1, arr(V[0]).c_str(), arr(M[0]).c_str(),
               arr(V[1]).c_str(), arr(M[1]).c_str(), arr(ch).c_str());
        for (int j = 0; j < rounds; ++j) {
            quadratic_poly q = rp.sumcheckUpdate(j ? ch[j - 1] : Fr(0), rp.r_u[1]);
            printf("%s[%s,%s,%s]", j ? "," : "", hx(q.a).c_str(), hx(q.b).c_str(), hx(q.c).c_str());
        }
        printf("], \"add_term\": %s}%s\n", hx(rp.add_term).c_str(), ci == 3 ? "" : ",");
    }
    printf("],\n");
    // ---- K2: prover::sumcheckDotProdUpdate1 (src/prover.cpp:103-144): multiplier over 2^m_bits frequencies, V_mult[0] zero beyond live0
    printf("\"cubic\": [\n");
    const int cubic_cases[][4] = {{5, 2, 12, 29}, {6, 3, 40, 64}, {4, 4, 16, 16}, {7, 1, 30, 100}};   // {bits, m_bits, live0, live1}
    for (int ci = 0; ci < 4; ++ci) {
        const int bits = cubic_cases[ci][0], mb = cubic_cases[ci][1], l0 = cubic_cases[ci][2], l1 = cubic_cases[ci][3];
        ref_prover rp;
        rp.total[0] = 1u << mb;
        rp.total[1] = 1u << bits;
        rp.total_size[1] = l1;
        rp.mult_array[1].resize(rp.total[0]);
        rp.V_mult[0].resize(rp.total[1]);
        rp.V_mult[1].resize(rp.total[1]);
        vector<Fr> mult(rp.total[0]), V0(l0), V1(l1);
        for (auto &x : mult) x = rnd();
        for (auto &x : V0) x = rnd();
        for (size_t i = 0; i < V1.size(); ++i) V1[i] = i % 2 ? small_signed() : rnd();
        for (u32 i = 0; i < rp.total[0]; ++i) rp.mult_array[1][i] = mult[i];
        for (u32 i = 0; i < rp.total[1]; ++i) {
            rp.V_mult[0][i] = (int) i < l0 ? V0[i] : Fr(0);
            rp.V_mult[1][i] = (int) i < l1 ? V1[i] : Fr(0);
        }
        vector<Fr> ch(bits);
        for (auto &x : ch) x = rnd();
        rp.r_u.assign(2, vector<Fr>(bits));
        rp.sumcheck_id = 1;
        rp.round = 0;
        printf(" {\"bits\": %d, \"m_bits\": %d, \"mult\": %s, \"V0\": %s, \"V1\": %s, \"r\": %s, \"polys\": [", bits, mb, arr(mult).c_str(), arr(V0).c_str(),
               arr(V1).c_str(), arr(ch).c_str());
        for (int j = 0; j < bits; ++j) {
            cubic_poly q = rp.sumcheckDotProdUpdate1(j ? ch[j - 1] : Fr(0));
            printf("%s[%s,%s,%s,%s]", j ? "," : "", hx(q.a).c_str(), hx(q.b).c_str(), hx(q.c).c_str(), hx(q.d).c_str());
        }
        Fr claim;
        rp.sumcheckDotProdFinalize1(ch[bits - 1], claim);
        printf("], \"claim_1\": %s, \"V_u1\": %s}%s\n", hx(claim).c_str(), hx(rp.V_u1).c_str(), ci == 3 ? "" : ",");
    }
    printf("],\n");
    // ---- K6b: prover::Vres (src/prover.cpp:434-457): MLE of a short output layer at r
    printf("\"vres\": [\n");
    const int vres_cases[][2] = {{10, 4}, {16, 4}, {1, 0}, {5, 3}};   // {output_size, r_size}
    for (int ci = 0; ci < 4; ++ci) {
        const int n = vres_cases[ci][0], rs = vres_cases[ci][1];
        ref_prover rp;
        rp.C.size = 2;
        rp.val.assign(2, vector<Fr>());
        rp.val[1].resize(n);
        for (auto &x : rp.val[1]) x = small_signed() + rnd();
        vector<Fr> r(rs + 1);
        for (auto &x : r) x = rnd();
        Fr out = rp.Vres(r.begin(), n, rs);
        r.resize(rs);
        printf(" {\"values\": %s, \"r\": %s, \"out\": %s}%s\n", arr(rp.val[1]).c_str(), arr(r).c_str(), hx(out).c_str(), ci == 3 ? "" : ",");
    }
    printf("]\n");
    printf("}\n");
    return 0;
}
