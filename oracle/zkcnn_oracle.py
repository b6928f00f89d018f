"""TEST INFRASTRUCTURE -- CPU restatement ("port") of the reference's hot-path algorithms in plain Python integers.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(zkcnn_b200/) never does.  Every function follows the reference loop it cites (paths relative to /root/reference),
including the reference's data layout (linear_poly = (a, b) with a = v1 - v0, b = v0) and its zero-padding quirks,
so it is an independent check of the evaluation-form CUDA kernels.  Field elements are canonical Python ints mod r.

Parity status: PINNED.  tests/test_oracle.py checks this module against known-answer vectors produced by the compiled
reference itself (oracle/harness/kat_gen.cpp -> tests/golden/kat.json) and the BLS12-381 constants of
mcl/test/bls12_test.cpp:40-54.
"""

R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001   # Fr modulus  (mcl/test/bls12_test.cpp:41)
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab   # Fp (:40)
G1_GEN = (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
          0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)   # (:53-54)
CURVE_B = 4


# ---- small deterministic generator shared with the tests ---------------------------------------------------------------
class SplitMix64:
    """the challenge stream of oracle/harness/seeded_rng.hpp"""

    def __init__(self, seed):
        self.state = seed & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.state = (self.state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def fr(self):
        """Fr::setByCSPRNG over this stream: 32 bytes LE, masked to 255 bits and, if still >= r, to 254 bits
        (mcl/include/mcl/fp.hpp:407-421,508-517)"""
        v = 0
        for k in range(4):
            v |= self.next() << (64 * k)
        v &= (1 << 255) - 1
        if v >= R:
            v &= (1 << 254) - 1
        return v


def is_negative(x):
    """mcl isNegative: canonical value >= (r + 1) / 2  (mcl/include/mcl/fp.hpp:666-671)"""
    return x % R >= (R + 1) // 2


# ---- eq ("beta") tables: src/utils.cpp:32-51 (initHalfTable), :147-180 (initBetaTable) ----------------------------------
def init_half_table(r, init, first_half, second_half):
    beta_f = [0] * (1 << first_half)
    beta_s = [0] * (1 << second_half)
    beta_f[0], beta_s[0] = init % R, 1
    for i in range(first_half):
        for j in range(1 << i):
            tmp = beta_f[j] * r[i] % R
            beta_f[j | (1 << i)] = tmp
            beta_f[j] = (beta_f[j] - tmp) % R
    for i in range(second_half):
        for j in range(1 << i):
            tmp = beta_s[j] * r[i + first_half] % R
            beta_s[j | (1 << i)] = tmp
            beta_s[j] = (beta_s[j] - tmp) % R
    return beta_f, beta_s


def init_beta_table(g_length, r, init):
    """4-argument overload, src/utils.cpp:168-180; also hyrax expand() (hyrax/src/utils.cpp:29-62) with init = 1"""
    first, second = g_length >> 1, g_length - (g_length >> 1)
    if init % R == 0:
        return [0] * (1 << g_length)
    f, s = init_half_table(r, init, first, second)
    mask = (1 << first) - 1
    return [f[i & mask] * s[i >> first] % R for i in range(1 << g_length)]


def init_beta_table2(g_length, r_0, r_1, alpha, beta):
    """6-argument overload, src/utils.cpp:147-165: alpha * eq(r_0) + beta * eq(r_1)"""
    out = init_beta_table(g_length, r_1, beta) if beta % R else [0] * (1 << g_length)
    if alpha % R == 0:
        return out
    a = init_beta_table(g_length, r_0, alpha)
    return [(x + y) % R for x, y in zip(out, a)]


# ---- FFT-layer phi table: src/utils.cpp:53-103, root of unity :224-232 ----------------------------------------------------
def root_of_unity(n):
    """getRootOfUnit: n-1 successive square roots of -1 as mcl's Fr::squareRoot picks them.  Pinned through the
    constants of SURVEY.md App. C (checked in tests/test_oracle.py): rou(n) = ROOT32^(2^(32-n))."""
    root32 = 0x3f0ee990743a3b6a0d6db230471dd5051ce1e93dfd4b71e59cab6d5c0c17f47c   # Montgomery form of getRootOfUnit(32)
    w = root32 * pow(1 << 256, -1, R) % R
    for _ in range(n, 32):
        w = w * w % R
    return w if n else 1


def phi_g_init(rx, scale, n, is_ifft):
    phi = root_of_unity(n)
    if is_ifft:
        phi = pow(phi, -1, R)
    phi_mul = [1] * (1 << n)
    for i in range(1, 1 << n):
        phi_mul[i] = phi_mul[i - 1] * phi % R
    phi_g = [0] * (1 << n)
    if is_ifft:
        phi_g[0] = phi_g[1] = scale % R
        levels = range(2, n + 1)
    else:
        phi_g[0] = scale % R
        levels = range(1, n)
    for i in levels:
        for b in range(1 << (i - 1)):
            l, r_ = b, b ^ (1 << (i - 1))
            m = n - i
            tmp1, tmp2 = (1 - rx[m]) % R, rx[m] * phi_mul[b << m] % R
            phi_g[r_] = phi_g[l] * (tmp1 - tmp2) % R
            phi_g[l] = phi_g[l] * (tmp1 + tmp2) % R
    if not is_ifft:
        for b in range(1 << (n - 1)):
            tmp1, tmp2 = (1 - rx[0]) % R, rx[0] * phi_mul[b] % R
            phi_g[b] = phi_g[b] * (tmp1 + tmp2) % R
        return phi_g[:1 << (n - 1)]
    return phi_g


# ---- sumcheck rounds in the reference's linear_poly layout: src/prover.cpp:13-15,368-426 -----------------------------------
def _interp(v0, v1):
    return ((v1 - v0) % R, v0 % R)          # linear_poly(a, b) = (one - zero, zero)


def _lin_eval(p, x):
    return (p[0] * x + p[1]) % R


class FoldState:
    """One (V_mult[idx], mult_array[idx]) pair as sumcheckInitPhase1 leaves it: constant polynomials (0, value),
    total = 2^bit_length, total_size = number of genuine entries (src/prover.cpp:159-172)."""

    def __init__(self, v, m, bit_length):
        n = 1 << bit_length
        self.v = [(0, x % R) for x in v] + [(0, 0)] * (n - len(v))
        self.m = [(0, x % R) for x in m] + [(0, 0)] * (n - len(m))
        self.total = n
        self.total_size = len(v)


def sumcheck_update_each(st, previous_random, add_term):
    """prover::sumcheckUpdateEach, src/prover.cpp:396-426.  Returns ((a, b, c), add_term)."""
    v, m = st.v, st.m
    if st.total == 1:
        v[0] = (0, _lin_eval(v[0], previous_random))
        m[0] = (0, _lin_eval(m[0], previous_random))
        add_term = (add_term + v[0][1] * m[0][1]) % R
    a = b = c = 0
    for i in range(st.total >> 1):
        g0, g1 = 2 * i, 2 * i + 1
        if g0 >= st.total_size:
            v[i] = (0, 0)
            m[i] = (0, 0)
            continue
        if g1 >= st.total_size:
            v[g1] = (0, 0)
            m[g1] = (0, 0)
        v[i] = _interp(_lin_eval(v[g0], previous_random), _lin_eval(v[g1], previous_random))
        m[i] = _interp(_lin_eval(m[g0], previous_random), _lin_eval(m[g1], previous_random))
        # linear * linear, src/polynomial.cpp:116-118
        a = (a + m[i][0] * v[i][0]) % R
        b = (b + m[i][0] * v[i][1] + m[i][1] * v[i][0]) % R
        c = (c + m[i][1] * v[i][1]) % R
    st.total >>= 1
    st.total_size = (st.total_size + 1) >> 1
    return (a, b, c), add_term


def sumcheck_rounds(pairs, challenges, n_rounds):
    """prover::sumcheckUpdate for n_rounds rounds (src/prover.cpp:368-383): round j uses previous_random =
    0 for j = 0 else challenges[j-1].  `pairs` is a list of FoldState.  Returns the list of (a, b, c)."""
    add_term = 0
    out = []
    for j in range(n_rounds):
        prev = 0 if j == 0 else challenges[j - 1]
        add_term = add_term * (1 - prev) % R
        a = b = c = 0
        for st in pairs:
            (pa, pb, pc), add_term = sumcheck_update_each(st, prev, add_term)
            a, b, c = (a + pa) % R, (b + pb) % R, (c + pc) % R
        out.append((a, (b - add_term) % R, (c + add_term) % R))
    return out


class DotProdState:
    """Tables of a DOT_PROD layer after sumcheckDotProdInitPhase1 (src/prover.cpp:57-95): mult_array[1] over the
    frequency index (2^fft_bl), V_mult[0] / V_mult[1] over 2^bit_length_u[1] entries."""

    def __init__(self, mult, v0, v1, bit_length, size):
        n = 1 << bit_length
        self.mult = [(0, x % R) for x in mult]
        self.v0 = [(0, x % R) for x in v0] + [(0, 0)] * (n - len(v0))
        self.v1 = [(0, x % R) for x in v1] + [(0, 0)] * (n - len(v1))
        self.total = [len(mult), n]
        self.total_size = size


def sumcheck_dotprod_update1(st, previous_random):
    """prover::sumcheckDotProdUpdate1, src/prover.cpp:103-144.  Returns the cubic (a, b, c, d)."""
    mult, v0, v1 = st.mult, st.v0, st.v1
    if st.total[0] == 1:
        mult[0] = (0, _lin_eval(mult[0], previous_random))
    else:
        for i in range(st.total[0] >> 1):
            mult[i] = _interp(_lin_eval(mult[2 * i], previous_random), _lin_eval(mult[2 * i + 1], previous_random))
    st.total[0] >>= 1
    ret = [0, 0, 0, 0]
    for i in range(st.total[1] >> 1):
        g0, g1 = 2 * i, 2 * i + 1
        if g0 >= st.total_size:
            v0[i] = (0, 0)
            v1[i] = (0, 0)
            continue
        if g1 >= st.total_size:
            v0[g1] = (0, 0)
            v1[g1] = (0, 0)
        v0[i] = _interp(_lin_eval(v0[g0], previous_random), _lin_eval(v0[g1], previous_random))
        v1[i] = _interp(_lin_eval(v1[g0], previous_random), _lin_eval(v1[g1], previous_random))
        mm = mult[i & (st.total[0] - 1)] if st.total[0] else mult[0]
        # (mult * v1) is a quadratic (linear * linear), times v0 a cubic (quadratic * linear, src/polynomial.cpp:78-80)
        qa, qb, qc = mm[0] * v1[i][0] % R, (mm[0] * v1[i][1] + mm[1] * v1[i][0]) % R, mm[1] * v1[i][1] % R
        xa, xb = v0[i]
        ret[0] = (ret[0] + qa * xa) % R
        ret[1] = (ret[1] + qa * xb + qb * xa) % R
        ret[2] = (ret[2] + qb * xb + qc * xa) % R
        ret[3] = (ret[3] + qc * xb) % R
    st.total[1] >>= 1
    st.total_size = (st.total_size + 1) >> 1
    return tuple(ret)


def vres(values, r):
    """prover::Vres, src/prover.cpp:434-457: multilinear extension of `values` at r (variable 0 = lowest index bit)"""
    n = 1 << len(r)
    cur = [x % R for x in values] + [0] * (n - len(values))
    for ri in r:
        cur = [(cur[2 * j] + ri * (cur[2 * j + 1] - cur[2 * j])) % R for j in range(len(cur) // 2)]
    return cur[0]


# ---- BLS12-381 G1 (affine, None = infinity): semantics of mcl's G1 as zkCNN uses it -----------------------------------------
def g1_on_curve(pt):
    return pt is None or (pt[1] * pt[1] - pt[0] ** 3 - CURVE_B) % P == 0


def g1_neg(pt):
    return None if pt is None else (pt[0], (-pt[1]) % P)


def g1_add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def g1_mul(pt, k):
    k %= R
    if k > R // 2:          # same group element, shorter chain (the witness is full of small negative values)
        return g1_neg(g1_mul(pt, R - k))
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, pt)
        pt = g1_add(pt, pt)
        k >>= 1
    return acc


def g1_mul_vec(points, scalars):
    """G1::mulVec as a group element (mcl/include/mcl/ec.hpp:1570-1597): sum_i scalars[i] * points[i]"""
    acc = None
    for pt, k in zip(points, scalars):
        acc = g1_add(acc, g1_mul(pt, k))
    return acc


# ---- Hyrax prover: 3rd/hyrax-bls12-381/src/polyProver.cpp -------------------------------------------------------------------
class HyraxProver:
    def __init__(self, Z, gens):
        self.bit_length = max(0, (len(Z) - 1).bit_length())
        self.Z = [z % R for z in Z] + [0] * ((1 << self.bit_length) - len(Z))
        self.gens = list(gens)

    def commit(self):                                   # polyProver.cpp:19-34
        r_bl = self.bit_length >> 1
        l_bl = self.bit_length - r_bl
        rsize, lsize = 1 << r_bl, 1 << l_bl
        assert lsize == len(self.gens)
        return [g1_mul_vec(self.gens, self.Z[i * lsize:(i + 1) * lsize]) for i in range(rsize)]

    def evaluate(self, x):                              # polyProver.cpp:36-42
        X = init_beta_table(len(x), x, 1)
        return sum(z * e for z, e in zip(self.Z, X)) % R

    def init_bullet_prove(self, lx, rx):                # polyProver.cpp:52-74
        self.t = list(lx)
        self.L = init_beta_table(len(lx), lx, 1)
        Rv = init_beta_table(len(rx), rx, 1)
        lsize = len(self.L)
        self.a = [sum(Rv[j] * self.Z[j * lsize + i] for j in range(len(Rv))) % R for i in range(lsize)]
        self.g = list(self.gens)
        self.scale = 1

    def bullet_prove(self):                             # polyProver.cpp:76-96
        h = len(self.a) >> 1
        lcomm = g1_mul_vec(self.g[:h], self.a[:h])
        rcomm = g1_mul_vec(self.g[h:], self.a[h:])
        self.scale = self.scale * pow((1 - self.t[-1]) % R, -1, R) % R
        ly = sum(self.a[i] * self.L[i] for i in range(h)) * self.scale % R
        ry = sum(self.a[i + h] * self.L[i] for i in range(h)) * self.scale % R
        return lcomm, rcomm, ly, ry

    def bullet_update(self, randomness):                # polyProver.cpp:98-109
        ir = pow(randomness, -1, R)
        h = len(self.a) >> 1
        self.a = [(self.a[i] * randomness + self.a[i + h]) % R for i in range(h)]
        self.g = [g1_add(g1_mul(self.g[i], ir), self.g[i + h]) for i in range(h)]
        self.t.pop()

    def bullet_open(self):                              # polyProver.cpp:111-116
        assert len(self.a) == 1
        return self.a[0]


def fnv1a(data, h=0xcbf29ce484222325):
    for b in data:
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h
