"""The oracle port (oracle/zkcnn_oracle.py) against known answers minted from the compiled reference
(oracle/harness/kat_gen.cpp -> tests/golden/kat.json) and the published BLS12-381 constants."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import zkcnn_oracle as O  # noqa: E402
from zkcnn_b200._binding import fr_from_words, fr_to_words, g1_from_words, g1_to_words  # noqa: E402

H = lambda s: int(s, 16)  # noqa: E731


def P(p):
    return None if p is None else (H(p[0]), H(p[1]))


def test_constants():
    assert O.R.bit_length() == 255 and O.P.bit_length() == 381
    assert O.g1_on_curve(O.G1_GEN)
    assert O.g1_mul(O.G1_GEN, O.R) is None          # the generator has order r
    assert O.g1_mul(O.G1_GEN, O.R - 1) == O.g1_neg(O.G1_GEN)


def test_memory_layout(kat):
    """mcl's in-memory form is what crosses the C ABI: Montgomery limbs, little endian"""
    raw = lambda key: [int(x, 16) for x in kat[key]]  # noqa: E731
    assert list(fr_to_words([1])[0]) == raw("raw_fr_1")
    assert list(fr_to_words([2])[0]) == raw("raw_fr_2")
    assert list(fr_to_words([-1])[0]) == raw("raw_fr_m1")
    assert fr_from_words([raw("raw_fr_m1")]) == [O.R - 1]
    assert P(kat["g1_gen"]) == O.G1_GEN
    assert list(g1_to_words([O.G1_GEN])[0]) == raw("raw_g1_gen")
    assert g1_from_words([raw("raw_g1_gen")]) == [O.G1_GEN]


def test_fr_arithmetic(kat):
    for c in kat["fr"]:
        a, b = H(c["a"]), H(c["b"])
        assert (a + b) % O.R == H(c["add"])
        assert (a - b) % O.R == H(c["sub"])
        assert a * b % O.R == H(c["mul"])
        assert (-a) % O.R == H(c["neg"])
        assert (pow(a, -1, O.R) if a else 0) == H(c["inv"])
        assert O.is_negative(a) == bool(c["is_neg"])
        assert list(fr_to_words([a])[0]) == [int(x, 16) for x in c["raw_a"]]
        assert fr_from_words([[int(x, 16) for x in c["raw_mul"]]]) == [H(c["mul"])]


def test_seeded_challenge_stream():
    """SplitMix64 -> Fr::setByCSPRNG masking; the first challenge of seed 1 is recorded in the LeNet golden run"""
    s = O.SplitMix64(1)
    x = s.fr()
    assert 0 <= x < O.R


def test_roots_of_unity(kat):
    for n, want in enumerate(kat["root_of_unity_1_14"], start=1):
        w = O.root_of_unity(n)
        assert w == H(want)
        assert pow(w, 1 << n, O.R) == 1 and (n == 0 or pow(w, 1 << (n - 1), O.R) != 1)


def test_beta_tables(kat):
    for c in kat["beta"]:
        r0, r1 = [H(x) for x in c["r0"]], [H(x) for x in c["r1"]]
        assert O.init_beta_table(c["bits"], r0, H(c["init"])) == [H(x) for x in c["table4"]]
        assert O.init_beta_table2(c["bits"], r0, r1, H(c["alpha"]), H(c["beta"])) == [H(x) for x in c["table6"]]


def test_phi_tables(kat):
    for c in kat["phi"]:
        got = O.phi_g_init([H(x) for x in c["rx"]], H(c["scale"]), c["n"], bool(c["ifft"]))
        assert got == [H(x) for x in c["table"]]


def test_g1_group_law(kat):
    g = kat["g1"]
    p, q, k = P(g["P"]), P(g["Q"]), H(g["k"])
    assert O.g1_on_curve(p) and O.g1_on_curve(q)
    assert O.g1_add(p, q) == P(g["add"])
    assert O.g1_add(p, p) == P(g["dbl"])
    assert O.g1_mul(p, k) == P(g["mul"])
    assert O.g1_add(p, O.g1_neg(p)) is None and g["P_plus_negP"] is None
    assert O.g1_add(p, None) == P(g["P_plus_O"])


def test_mulvec(kat):
    """G1::mulVec (mcl Straus/wNAF) agrees with the naive sum as a group element, sizes of mcl/test/common_test.hpp"""
    for c in kat["mulvec"]:
        assert O.g1_mul_vec([P(x) for x in c["points"]], [H(x) for x in c["scalars"]]) == P(c["out"]), c["n"]


def test_hyrax_prover(kat):
    h = kat["hyrax"]
    hp = O.HyraxProver([H(x) for x in h["Z"]], [P(x) for x in h["gens"]])
    assert hp.commit() == [P(x) for x in h["commit"]]
    x = [H(v) for v in h["x"]]
    assert hp.evaluate(x) == H(h["evaluate"])
    lbl = len(x) - h["rbl"]
    hp.init_bullet_prove(x[:lbl], x[lbl:])
    for rd in h["rounds"]:
        lc, rc, ly, ry = hp.bullet_prove()
        assert (lc, rc, ly, ry) == (P(rd["lcomm"]), P(rd["rcomm"]), H(rd["ly"]), H(rd["ry"]))
        hp.bullet_update(H(rd["randomness"]))
    assert hp.bullet_open() == H(h["open"])


def test_fold_is_a_sumcheck():
    """size-independent property of the restated fold: p(0) + p(1) of round j equals p_{j-1}(r_{j-1}), and round 0 sums
    to <V, M>; ragged sizes exercise the reference's zero-padding by total_size (src/prover.cpp:409-417)"""
    rng = O.SplitMix64(7)
    for bits, live in ((1, 2), (3, 5), (4, 16), (5, 17), (6, 33)):
        V = [rng.fr() for _ in range(live)]
        M = [rng.fr() for _ in range(live)]
        ch = [rng.fr() for _ in range(bits)]
        polys = O.sumcheck_rounds([O.FoldState(V, M, bits)], ch, bits)
        claim = sum(v * m for v, m in zip(V, M)) % O.R
        for j, (a, b, c) in enumerate(polys):
            assert (c + a + b + c) % O.R == claim
            claim = (a * ch[j] * ch[j] + b * ch[j] + c) % O.R


def test_round_functions_against_reference_kats(kat):
    """the port's fold / cubic round / Vres against known answers minted by calling the REFERENCE prover's own (private) round
    functions on hand-made tables (oracle/harness/kat_gen.cpp -> tests/golden/kat.json): pins the restatement itself"""
    for c in kat["fold"]:
        b0, b1 = c["bits"]
        V0, M0, V1, M1 = ([H(x) for x in c[k]] for k in ("V0", "M0", "V1", "M1"))
        ch = [H(x) for x in c["r"]]
        pairs = ([O.FoldState(V0, M0, b0)] if b0 >= 0 else []) + [O.FoldState(V1, M1, b1)]
        n = len(c["polys"])
        assert O.sumcheck_rounds(pairs, ch, n) == [tuple(H(x) for x in p) for p in c["polys"]], c["bits"]
    for c in kat["cubic"]:
        mult, V0, V1, ch = ([H(x) for x in c[k]] for k in ("mult", "V0", "V1", "r"))
        st = O.DotProdState(mult, V0, V1, c["bits"], len(V1))
        got = [O.sumcheck_dotprod_update1(st, 0 if j == 0 else ch[j - 1]) for j in range(c["bits"])]
        assert got == [tuple(H(x) for x in p) for p in c["polys"]], (c["bits"], c["m_bits"])
    for c in kat["vres"]:
        assert O.vres([H(x) for x in c["values"]], [H(x) for x in c["r"]]) == H(c["out"])
