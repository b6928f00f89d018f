// See verifier.hpp.  Protocol order, challenge order and acceptance conditions are those of the reference
// (src/verifier.cpp, 3rd/hyrax-bls12-381/src/polyVerifier.cpp); the code is organised around
// small helpers (challenge drawing, round checking, wiring-predicate evaluation on host threads) instead of the
// reference's timer-interleaved loops.  Field arithmetic is exact, so the eq/phi tables here may be built in any order.
#include "verifier.hpp"
#include <chrono>
#include "challenge_stream.hpp"
#include <array>
#include <functional>
#include <thread>

static inline const uint64_t *w(const F &x) { return reinterpret_cast<const uint64_t *>(&x); }
static inline const uint64_t *w(const G &x) { return reinterpret_cast<const uint64_t *>(&x); }
static inline uint64_t *w(G &x) { return reinterpret_cast<uint64_t *>(&x); }

static void require(int rc, const char *what) {
    if (rc != 0) throw std::runtime_error(std::string("zkcnn_b200: ") + what + ": " + zk_last_error());
}

// ---- eq / phi tables on the host ---------------------------------------------------------------------------------------------
// out[g] = init * prod_j (g_j ? r_j : 1 - r_j)   (initBetaTable, src/utils.cpp:168-180)
static void eqTable(vector<F> &out, int bits, const F *r, const F &init) {
    if (bits < 0) return;
    out.assign((size_t) 1 << bits, F());
    if (init.isZero()) return;
    out[0] = init;
    for (int j = 0; j < bits; ++j) {
        const size_t n = (size_t) 1 << j;
        for (size_t g = 0; g < n; ++g) {
            F hi = out[g] * r[j];
            out[g | n] = hi;
            out[g] -= hi;
        }
    }
}
void initBetaTable(vector<F> &beta_g, u8 gLength, const F *r, const F &init) { eqTable(beta_g, (i8) gLength, r, init); }
// alpha * eq(r_0) + beta * eq(r_1)   (src/utils.cpp:147-165)
void initBetaTable(vector<F> &beta_g, u8 gLength, const F *r_0, const F *r_1, const F &alpha, const F &beta) {
    eqTable(beta_g, (i8) gLength, r_0, alpha);
    if (beta.isZero()) return;
    vector<F> second;
    eqTable(second, (i8) gLength, r_1, beta);
    for (size_t g = 0; g < beta_g.size(); ++g) beta_g[g] += second[g];
}

// MLE of the (inverse) FFT matrix in its row variable at the point rx (phiGInit, src/utils.cpp:61-103): a butterfly
// product table over the powers of the 2^n-th root of unity.  FFT fills 2^(n-1) entries, IFFT 2^n.
void phiGInit(vector<F> &phi, const F *rx, const F &scale, int n, bool isIFFT) {
    vector<F> pw((size_t) 1 << n);
    F root = getRootOfUnit(n);
    if (isIFFT) F::inv(root, root);
    pw[0] = F_ONE;
    for (size_t i = 1; i < pw.size(); ++i) pw[i] = pw[i - 1] * root;
    if (phi.size() < ((size_t) 1 << n)) phi.resize((size_t) 1 << n);
    phi[0] = scale;
    int first = 1, last = n - 1;
    if (isIFFT) { phi[1] = scale; first = 2; last = n; }
    for (int lvl = first; lvl <= last; ++lvl) {
        const u32 half = 1u << (lvl - 1);
        const int m = n - lvl;
        const F keep = F_ONE - rx[m];
        for (u32 b = 0; b < half; ++b) {
            const F tw = rx[m] * pw[(size_t) b << m];
            const F x = phi[b];
            phi[b ^ half] = x * (keep - tw);
            phi[b] = x * (keep + tw);
        }
    }
    if (!isIFFT) {
        const F keep = F_ONE - rx[0];
        for (u32 b = 0; b < (1u << (n - 1)); ++b) phi[b] = phi[b] * (keep + rx[0] * pw[b]);
    }
}

// sum over [0, n) of a 3-component field vector, split over host threads
typedef std::array<F, 3> F3;
static F3 parallelSum3(size_t n, const std::function<void(size_t, size_t, F3 &)> &body) {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (n < 4096) nt = 1;
    nt = std::min<unsigned>(nt, 64);
    vector<F3> part(nt);
    vector<std::thread> th;
    const size_t per = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
        const size_t b = std::min(n, t * per), e = std::min(n, b + per);
        if (nt == 1) body(b, e, part[t]);
        else th.emplace_back([&, t, b, e] { body(b, e, part[t]); });
    }
    for (auto &x : th) x.join();
    F3 tot;
    for (auto &p : part) for (int k = 0; k < 3; ++k) tot[k] += p[k];
    return tot;
}

// ---- Hyrax verifier ------------------------------------------------------------------------------------------------------------
namespace hyrax_bls12_381 {

void split(vector<Fr> &L, vector<Fr> &R, const vector<Fr> &r) {   // low ceil(n/2) variables select the column
    const size_t rs = r.size() >> 1, ls = r.size() - rs;
    L.assign(r.begin(), r.begin() + ls);
    R.assign(r.begin() + ls, r.end());
}

vector<Fr> expand(const vector<Fr> &v) {
    vector<Fr> out;
    eqTable(out, (int) v.size(), v.data(), Fr::one());
    return out;
}

polyVerifier::polyVerifier(polyProver &_p, const vector<G1> &_gens, zk_ctx *_ctx, bool _check_points)
    : p(_p), ctx(_ctx), check_points(_check_points), gens(_gens) {
    vt.start();
    comm_Z = p.commit();   // polyVerifier.cpp:12
    vt.stop();
}

bool polyVerifier::verify(const vector<Fr> &_x, const Fr &RZL) {   // polyVerifier.cpp:18-32
    vt.start();
    x = _x;
    split(lx, rx, x);
    p.initBulletProve(lx, rx);
    comm_RZ.clear();
    if (check_points) {   // comm_RZ = sum_j R[j] comm_Z[j], on the device
        vector<Fr> R = expand(rx);
        if (R.size() != comm_Z.size()) throw std::logic_error("polyVerifier: commitment count does not match the point");
        require(zk_msm(ctx, w(comm_Z[0]), w(R[0]), R.size(), 1, w(comm_RZ)), "zk_msm");
    }
    bool ok = bulletVerify(gens, lx, comm_RZ, RZL);
    vt.stop();
    return ok;
}

bool polyVerifier::bulletVerify(vector<G1> g, vector<Fr> t, G1 comm, Fr y) {   // polyVerifier.cpp:34-78
    G1 lcomm, rcomm;
    Fr ly, ry;
    const size_t logn = t.size();
    if (logn == 0) throw std::logic_error("polyVerifier: empty opening point");
    vector<Fr> ahead;
    if (batchRounds) {
        ahead.resize(logn);
        for (auto &r : ahead) r.setByCSPRNG();
        p.bulletProveAll(ahead);
    }
    for (size_t round = 0;; ++round) {
        p.bulletProve(lcomm, rcomm, ly, ry);
        Fr rho, irho;
        if (batchRounds) rho = ahead[round];
        else rho.setByCSPRNG();
        Fr::inv(irho, rho);
        p.bulletUpdate(rho);
        if (check_points) {
            // g[i] <- g[i] / rho + g[i + h];  comm <- rho lcomm + comm + rcomm / rho   (device-side point arithmetic)
            const size_t h = g.size() >> 1;
            vector<Fr> k(h, irho);
            vector<G1> scaled(h);
            require(zk_g1_vec_op(ctx, 2, w(g[0]), w(k[0]), w(scaled[0]), h), "zk_g1_vec_op(mul)");
            require(zk_g1_vec_op(ctx, 0, w(scaled[0]), w(g[h]), w(g[0]), h), "zk_g1_vec_op(add)");
            g.resize(h);
            G1 pts[2] = {lcomm, rcomm}, out[2];
            Fr ks[2] = {rho, irho};
            require(zk_g1_vec_op(ctx, 2, w(pts[0]), w(ks[0]), w(out[0]), 2), "zk_g1_vec_op(mul)");
            comm = out[0] + comm + out[1];
        }
        if (y != ly * (Fr::one() - t.back()) + ry * t.back()) {
            fprintf(stderr, "y incorrect at %d.\n", (int) (logn - t.size()));
            return false;
        }
        y = ly * rho + ry;
        if (t.size() == 1) {
            bool ok = p.bulletOpen() == y;
            if (ok && check_points) ok = comm == g.back() * y;
            if (!ok) fprintf(stderr, "last step incorrect.\n");
            return ok;
        }
        t.pop_back();
    }
}

}  // namespace hyrax_bls12_381

// ---- GKR verifier ----------------------------------------------------------------------------------------------------------------
verifier::verifier(prover *pr, const layeredCircuit &cir) : p(pr), C(cir) {
    final_claim_u0.assign(C.size + 2, F());
    final_claim_v0.assign(C.size + 2, F());
    r_u.assign(C.size + 2, {});
    r_v.assign(C.size + 2, {});
    p->init();   // src/verifier.cpp:22
}

static void drawChallenges(vector<F> &v, size_t n) {
    v.resize(n);
    for (auto &x : v) x.setByCSPRNG();
}

bool verifier::verify() {   // src/verifier.cpp:118-130
    const auto t0 = std::chrono::steady_clock::now();
    const u8 logn = C.circuit[0].bit_length;
    const u64 n_gens = 1ULL << (logn - (logn >> 1));
    vector<F> k;
    drawChallenges(k, n_gens);   // always drawn, so that the challenge stream does not depend on the generator mode
    if (fixedGenerators) {
        if (fixedGenerators->size() != n_gens) throw std::invalid_argument("verifier: fixedGenerators has the wrong size");
        generators = *fixedGenerators;
    } else if (realGenerators) {
        const G base = G::generator();
        generators.assign(n_gens, G());
        require(zk_g1_fixed_base_mul(p->context(), w(base), w(k[0]), n_gens, w(generators[0])), "zk_g1_fixed_base_mul");
    } else generators.assign(n_gens, G());   // reference default: base point cleared by initPairing -> all infinity
    static const bool trace = getenv("ZKH_TRACE") != nullptr;   // stderr: wall time of the stages of one proof
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t1 = now();
    poly_v.reset(new hyrax_bls12_381::polyVerifier(p->commitInput(generators), generators, p->context(), checkPredicates));
    poly_v->batchRounds = batchRounds && !fiatShamir;
    const auto t2 = now();
    const bool ok1 = verifyInnerLayers();
    const auto t3 = now();
    const bool ok2 = ok1 && verifyFirstLayer();
    const auto t4 = now();
    const bool ok3 = ok2 && verifyInput();
    const auto t5 = now();
    if (trace) fprintf(stderr, "ZKH_TRACE generators %.2f ms  commit %.2f ms  inner layers %.2f ms  input layer %.2f ms  opening %.2f ms\n", ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t4, t5));
    return ok3;
}

F verifier::getFinalValue(const F &claim_u0, const F &claim_u1, const F &claim_v0, const F &claim_v1) {   // src/verifier.cpp:25-34
    return bin_value[0] * (claim_u0 * claim_v0) + bin_value[1] * (claim_u1 * claim_v1) + bin_value[2] * (claim_u1 * claim_v0) +
           uni_value[0] * claim_u0 + uni_value[1] * claim_u1;
}

// eq tables of the layer's output variable (beta_g) and of the u variable (beta_u)   (src/verifier.cpp:36-89)
void verifier::betaInitPhase1(u8 depth, const F &alpha, const F &beta, const vector<F> &r_0, const vector<F> &r_1, const F &relu_rou) {
    const layer &L = C.circuit[depth];
    const int bl = L.bit_length, fft_bl = L.fft_bit_length, fft_blh = fft_bl - 1;
    switch (L.ty) {
        case layerType::FFT:
        case layerType::IFFT:
            beta_gs.assign((size_t) 1 << fft_bl, F());
            phiGInit(beta_gs, r_0.data(), L.scale, fft_bl, L.ty == layerType::IFFT);
            eqTable(beta_u, L.max_bl_u, r_u[depth].data(), F_ONE);
            break;
        case layerType::PADDING: {
            // output index = (channel block | position inside the FFT block): the block part was fixed two layers up
            vector<F> hi;
            initBetaTable(hi, bl - fft_blh, r_u[depth + 2].data() + fft_bl, r_v[depth + 2].data(), alpha, beta);
            eqTable(beta_gs, fft_blh, r_0.data(), F_ONE);
            beta_g.resize((size_t) 1 << bl);
            const u32 mask = (1u << fft_blh) - 1;
            for (size_t g = 0; g < beta_g.size(); ++g) beta_g[g] = hi[g >> fft_blh] * beta_gs[g & mask];
            eqTable(beta_u, L.max_bl_u, r_u[depth].data(), F_ONE);
            break;
        }
        case layerType::DOT_PROD: {
            const int cnt_bl = bl - fft_bl, cnt_bl2 = L.max_bl_u - fft_bl;
            eqTable(beta_g, cnt_bl, r_u[depth + 2].data() + fft_bl - 1, alpha);
            // the frequency variables of r_0 and r_u[depth] must agree: one common factor
            F same = F_ONE;
            for (int j = 0; j < fft_bl; ++j) same *= r_0[j] * r_u[depth][j] + (F_ONE - r_0[j]) * (F_ONE - r_u[depth][j]);
            eqTable(beta_u, cnt_bl2, r_u[depth].data() + fft_bl, same);
            break;
        }
        default:
            initBetaTable(beta_g, bl, r_0.data(), r_1.data(), alpha * L.scale, beta * L.scale);
            if (L.zero_start_id < L.size)
                for (size_t g = L.zero_start_id; g < beta_g.size(); ++g) beta_g[g] *= relu_rou;
            eqTable(beta_u, L.max_bl_u, r_u[depth].data(), F_ONE);
    }
}

void verifier::betaInitPhase2(u8 depth) { eqTable(beta_v, C.circuit[depth].max_bl_v, r_v[depth].data(), F_ONE); }

void verifier::predicatePhase1(u8 layer_id) {   // src/verifier.cpp:96-108
    const layer &L = C.circuit[layer_id];
    F3 s;
    if (L.ty == layerType::FFT || L.ty == layerType::IFFT) {
        s = parallelSum3((size_t) 1 << L.max_bl_u, [&](size_t b, size_t e, F3 &acc) {
            for (size_t u = b; u < e; ++u) acc[1] += beta_gs[u] * beta_u[u];
        });
    } else {
        s = parallelSum3(L.uni_gates.size(), [&](size_t b, size_t e, F3 &acc) {
            for (size_t i = b; i < e; ++i) {
                const uniGate &g = L.uni_gates[i];
                acc[g.lu ? 1 : 0] += beta_g[g.g] * beta_u[g.u] * C.two_mul[g.sc];
            }
        });
    }
    uni_value[0] = s[0];
    uni_value[1] = s[1];
    bin_value[0] = bin_value[1] = bin_value[2] = F_ZERO;
}

void verifier::predicatePhase2(u8 layer_id) {   // src/verifier.cpp:110-123
    uni_value[0] *= beta_v[0];
    uni_value[1] *= beta_v[0];
    const layer &L = C.circuit[layer_id];
    const bool scaled = L.ty != layerType::DOT_PROD;
    F3 s = parallelSum3(L.bin_gates.size(), [&](size_t b, size_t e, F3 &acc) {
        for (size_t i = b; i < e; ++i) {
            const binGate &g = L.bin_gates[i];
            F t = beta_g[g.g] * beta_u[g.u] * beta_v[g.v];
            if (scaled && g.sc) t *= C.two_mul[g.sc];
            acc[g.l] += t;
        }
    });
    for (int k = 0; k < 3; ++k) bin_value[k] = s[k];
}

// The same five values as betaInitPhase1/2 + predicatePhase1/2 above, computed on the device (include/zkcnn_b200.h, "verifier side")
void verifier::predicatesOnDevice(u8 depth, const F &alpha, const F &beta, const F &relu_rou) {
    zk_ctx *ctx = p->context();
    const layer &L = C.circuit[depth];
    const int bl = L.bit_length, fft_bl = L.fft_bit_length, fft_blh = fft_bl - 1;
    const vector<F> &r_0 = r_u[depth + 1], &r_1 = r_v[depth + 1];
    const F one = F_ONE;
    auto ptr = [](const vector<F> &v, size_t off = 0) -> const uint64_t * { return v.size() > off ? w(v[off]) : nullptr; };
    enum { G = 0, U = 1, V = 2, T0 = 3, T1 = 4 };
    uni_value[0] = uni_value[1] = bin_value[0] = bin_value[1] = bin_value[2] = F_ZERO;
    if (L.ty == layerType::FFT || L.ty == layerType::IFFT) {   // sum_u phi(r_0)[u] eq(r_u)[u]
        require(zk_vtab_phi(ctx, T0, ptr(r_0), w(L.scale), fft_bl, L.ty == layerType::IFFT), "zk_vtab_phi");
        require(zk_vtab_eq(ctx, U, L.max_bl_u, ptr(r_u[depth]), w(one), nullptr, nullptr, 0, nullptr), "zk_vtab_eq");
        require(zk_vtab_dot(ctx, T0, U, (uint64_t) 1 << L.max_bl_u, reinterpret_cast<uint64_t *>(&uni_value[1])), "zk_vtab_dot");
        return;
    }
    switch (L.ty) {
        case layerType::PADDING: {
            const F &a2 = alpha, &b2 = beta;
            require(zk_vtab_eq(ctx, T0, bl - fft_blh, ptr(r_u[depth + 2], fft_bl), w(a2), b2.isZero() ? nullptr : ptr(r_v[depth + 2]), w(b2), 0, nullptr), "zk_vtab_eq");
            require(zk_vtab_eq(ctx, T1, fft_blh, ptr(r_0), w(one), nullptr, nullptr, 0, nullptr), "zk_vtab_eq");
            require(zk_vtab_outer(ctx, G, T0, T1, bl, fft_blh), "zk_vtab_outer");
            require(zk_vtab_eq(ctx, U, L.max_bl_u, ptr(r_u[depth]), w(one), nullptr, nullptr, 0, nullptr), "zk_vtab_eq");
            break;
        }
        case layerType::DOT_PROD: {
            const int cnt_bl = bl - fft_bl, cnt_bl2 = L.max_bl_u - fft_bl;
            require(zk_vtab_eq(ctx, G, cnt_bl, ptr(r_u[depth + 2], fft_bl - 1), w(alpha), nullptr, nullptr, 0, nullptr), "zk_vtab_eq");
            F same = F_ONE;
            for (int j = 0; j < fft_bl; ++j) same *= r_0[j] * r_u[depth][j] + (F_ONE - r_0[j]) * (F_ONE - r_u[depth][j]);
            require(zk_vtab_eq(ctx, U, cnt_bl2, ptr(r_u[depth], fft_bl), w(same), nullptr, nullptr, 0, nullptr), "zk_vtab_eq");
            break;
        }
        default: {
            const F a = alpha * L.scale, b = beta * L.scale;
            const bool tail = L.zero_start_id < L.size;
            require(zk_vtab_eq(ctx, G, bl, ptr(r_0), w(a), b.isZero() ? nullptr : ptr(r_1), w(b), tail ? L.zero_start_id : 0, tail ? w(relu_rou) : nullptr), "zk_vtab_eq");
            require(zk_vtab_eq(ctx, U, L.max_bl_u, ptr(r_u[depth]), w(one), nullptr, nullptr, 0, nullptr), "zk_vtab_eq");
        }
    }
    F bv0 = F_ONE;
    if (L.need_phase2) {
        require(zk_vtab_eq(ctx, V, L.max_bl_v, ptr(r_v[depth]), w(one), nullptr, nullptr, 0, nullptr), "zk_vtab_eq");
        for (int j = 0; j < L.max_bl_v; ++j) bv0 *= F_ONE - r_v[depth][j];   // beta_v[0]
    }
    F out[5];
    require(zk_verifier_layer_predicates(ctx, depth, G, U, L.need_phase2 ? V : -1, reinterpret_cast<uint64_t *>(&out[0])), "zk_verifier_layer_predicates");
    uni_value[0] = out[0] * bv0;
    uni_value[1] = out[1] * bv0;
    for (int k = 0; k < 3; ++k) bin_value[k] = out[2 + k];
}

F verifier::inputPredicateOnDevice(const vector<F> &sig_u, const vector<F> &sig_v) {
    zk_ctx *ctx = p->context();
    const layer &in = C.circuit[0];
    const F one = F_ONE;
    require(zk_vtab_eq(ctx, 0, in.bit_length, r_u[0].empty() ? nullptr : w(r_u[0][0]), w(one), nullptr, nullptr, 0, nullptr), "zk_vtab_eq");
    vector<F> ru, rv;
    for (int i = 1; i < C.size; ++i) {
        const layer &L = C.circuit[i];
        if (L.bit_length_u[0] != -1) ru.insert(ru.end(), r_u[i].begin(), r_u[i].begin() + L.bit_length_u[0]);
        if (L.bit_length_v[0] != -1) rv.insert(rv.end(), r_v[i].begin(), r_v[i].begin() + L.bit_length_v[0]);
    }
    ru.push_back(F_ZERO);   // (never empty: a valid pointer for the call)
    rv.push_back(F_ZERO);
    F gr;
    require(zk_verifier_input_predicate(ctx, 0, w(sig_u[0]), w(sig_v[0]), w(ru[0]), w(rv[0]), reinterpret_cast<uint64_t *>(&gr)), "zk_verifier_input_predicate");
    return gr;
}

bool verifier::verifyInnerLayers() {   // src/verifier.cpp:132-266
    const layer &out = C.circuit[C.size - 1];
    F alpha = F_ONE, beta = F_ZERO, relu_rou, claim_u1, claim_v1;
    total_timer.start();
    drawChallenges(r_u[C.size], out.bit_length);
    total_timer.stop();
    F sum = p->Vres(r_u[C.size].begin(), out.size, out.bit_length);
    p->sumcheckInitAll(r_u[C.size].begin());

    for (u8 i = C.size - 1; i; --i) {
        const layer &cur = C.circuit[i];
        const bool dot = cur.ty == layerType::DOT_PROD;
        p->sumcheckInit(alpha, beta);

        // ---- phase 1: all challenges of the phase are drawn before its rounds (src/verifier.cpp:156-160)
        total_timer.start();
        if (fiatShamir) r_u[i].assign(cur.max_bl_u, F());   // drawn round by round, after the message each one answers
        else drawChallenges(r_u[i], cur.max_bl_u);
        if (cur.zero_start_id < cur.size) relu_rou.setByCSPRNG();
        else relu_rou = F_ONE;
        total_timer.stop();
        if (dot) p->sumcheckDotProdInitPhase1();
        else p->sumcheckInitPhase1(relu_rou);

        F prev = F_ZERO;
        vector<quadratic_poly> all;
        vector<cubic_poly> all_dot;
        const bool batch = batchRounds && !fiatShamir;
        if (batch && !dot) all = p->sumcheckUpdateAll(1, r_u[i], cur.max_bl_u);
        if (batch && dot) all_dot = p->sumcheckDotProdUpdateAll(r_u[i], cur.max_bl_u);
        for (int j = 0; j < cur.max_bl_u; ++j) {
            F at0p1, atr;
            if (dot) {
                cubic_poly poly = batch ? all_dot[j] : p->sumcheckDotProdUpdate1(prev);
                if (fiatShamir) r_u[i][j].setByCSPRNG();
                at0p1 = poly.d + poly.eval(F_ONE);
                atr = poly.eval(r_u[i][j]);
            } else {
                quadratic_poly poly = batch ? all[j] : p->sumcheckUpdate1(prev);
                if (fiatShamir) r_u[i][j].setByCSPRNG();
                at0p1 = poly.c + poly.eval(F_ONE);
                atr = poly.eval(r_u[i][j]);
            }
            if (at0p1 != sum) {
                fprintf(stderr, "Verification fail, phase1, circuit %d, current bit %d\n", (int) i, j);
                return false;
            }
            prev = r_u[i][j];
            sum = atr;
        }
        if (dot) p->sumcheckDotProdFinalize1(prev, claim_u1);
        else p->sumcheckFinalize1(prev, final_claim_u0[i], claim_u1);

        total_slow_timer.start();
        if (checkPredicates && !devicePredicates) {
            betaInitPhase1(i, alpha, beta, r_u[i + 1], r_v[i + 1], relu_rou);
            predicatePhase1(i);
        }
        total_slow_timer.stop();

        // ---- phase 2
        if (cur.need_phase2) {
            total_timer.start();
            if (fiatShamir) r_v[i].assign(cur.max_bl_v, F());
            else drawChallenges(r_v[i], cur.max_bl_v);
            total_timer.stop();
            p->sumcheckInitPhase2();
            prev = F_ZERO;
            if (batch) all = p->sumcheckUpdateAll(2, r_v[i], cur.max_bl_v);
            for (int j = 0; j < cur.max_bl_v; ++j) {
                quadratic_poly poly = batch ? all[j] : p->sumcheckUpdate2(prev);
                if (fiatShamir) r_v[i][j].setByCSPRNG();
                if (poly.c + poly.eval(F_ONE) != sum) {
                    fprintf(stderr, "Verification fail, phase2, circuit level %d, current bit %d, total is %d\n", (int) i, j, (int) cur.max_bl_v);
                    return false;
                }
                prev = r_v[i][j];
                sum = poly.eval(prev);
            }
            p->sumcheckFinalize2(prev, final_claim_v0[i], claim_v1);
            total_slow_timer.start();
            if (checkPredicates && !devicePredicates) {
                betaInitPhase2(i);
                predicatePhase2(i);
            }
            total_slow_timer.stop();
        }
        if (checkPredicates && devicePredicates) {
            total_slow_timer.start();
            predicatesOnDevice(i, alpha, beta, relu_rou);
            total_slow_timer.stop();
        }
        if (checkPredicates && sum != getFinalValue(final_claim_u0[i], claim_u1, final_claim_v0[i], claim_v1)) {
            fprintf(stderr, "Verification fail, semi final, circuit level %d\n", (int) i);
            return false;
        }

        // ---- claim for the next layer (src/verifier.cpp:245-255)
        total_timer.start();
        if (cur.ty == layerType::FFT || cur.ty == layerType::IFFT) sum = claim_u1;
        else {
            if (cur.bit_length_u[1] != -1) alpha.setByCSPRNG();
            else alpha.clear();
            if (cur.bit_length_v[1] != -1) beta.setByCSPRNG();
            else beta.clear();
            sum = alpha * claim_u1 + beta * claim_v1;
        }
        total_timer.stop();
        beta_u.clear();
        beta_v.clear();
    }
    return true;
}

bool verifier::verifyFirstLayer() {   // src/verifier.cpp:268-357
    const layer &in = C.circuit[0];
    vector<F> sig_u, sig_v;
    total_timer.start();
    drawChallenges(sig_u, C.size - 1);
    drawChallenges(sig_v, C.size - 1);
    if (fiatShamir) r_u[0].assign(in.bit_length, F());
    else drawChallenges(r_u[0], in.bit_length);
    F sum = F_ZERO;
    for (int i = 1; i < C.size; ++i) {
        if (C.circuit[i].bit_length_u[0] != -1) sum += sig_u[i - 1] * final_claim_u0[i];
        if (C.circuit[i].bit_length_v[0] != -1) sum += sig_v[i - 1] * final_claim_v0[i];
    }
    total_timer.stop();

    p->sumcheckLiuInit(sig_u, sig_v);
    F prev = F_ZERO;
    vector<quadratic_poly> all;
    const bool batch = batchRounds && !fiatShamir;
    if (batch) all = p->sumcheckUpdateAll(0, r_u[0], in.bit_length);
    for (int j = 0; j < in.bit_length; ++j) {
        quadratic_poly poly = batch ? all[j] : p->sumcheckLiuUpdate(prev);
        if (fiatShamir) r_u[0][j].setByCSPRNG();
        if (poly.c + poly.eval(F_ONE) != sum) {
            fprintf(stderr, "Liu fail, circuit 0, current bit %d\n", j);
            return false;
        }
        prev = r_u[0][j];
        sum = poly.eval(prev);
    }
    p->sumcheckLiuFinalize(prev, eval_in);

    if (checkPredicates && devicePredicates) {
        total_slow_timer.start();
        const F gr = inputPredicateOnDevice(sig_u, sig_v);
        total_slow_timer.stop();
        if (eval_in * gr != sum) {
            fprintf(stderr, "Liu fail, semi final, circuit 0.\n");
            return false;
        }
    } else if (checkPredicates) {   // gr = sum over all layer-0 operand slots of eq(r_u[0])[ori_id] * sigma-weighted eq(r_u[i] / r_v[i])
        total_slow_timer.start();
        eqTable(beta_g, in.bit_length, r_u[0].data(), F_ONE);
        F gr = F_ZERO;
        for (int i = 1; i < C.size; ++i) {
            const layer &L = C.circuit[i];
            for (int side = 0; side < 2; ++side) {
                const int bl = side ? L.bit_length_v[0] : L.bit_length_u[0];
                if (bl == -1) continue;
                const vector<u32> &ori = side ? L.ori_id_v : L.ori_id_u;
                vector<F> &tab = side ? beta_v : beta_u;
                eqTable(tab, bl, (side ? r_v[i] : r_u[i]).data(), side ? sig_v[i - 1] : sig_u[i - 1]);
                F3 s = parallelSum3(side ? L.size_v[0] : L.size_u[0], [&](size_t b, size_t e, F3 &acc) {
                    for (size_t j = b; j < e; ++j) acc[0] += beta_g[ori[j]] * tab[j];
                });
                gr += s[0];
            }
        }
        total_slow_timer.stop();
        if (eval_in * gr != sum) {
            fprintf(stderr, "Liu fail, semi final, circuit 0.\n");
            return false;
        }
    }
    beta_g.clear(); beta_gs.clear(); beta_u.clear(); beta_v.clear();
    return true;
}

bool verifier::verifyInput() {   // src/verifier.cpp:359-373
    if (!poly_v->verify(r_u[0], eval_in)) {
        fprintf(stderr, "Verification fail, final input check fail.\n");
        return false;
    }
    polyVT = poly_v->getVT();
    return true;
}
