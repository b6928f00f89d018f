"""CPU suite: whole proofs through the stand-alone host side (circuit compiler, witness generator, protocol driver) on
the emulator build of the kernels, against golden transcripts / circuit dumps minted from the compiled reference
(oracle/harness/make_golden.sh).  Covers naive and FFT convolution, max pooling, ReLU, FC, pic_cnt 1 and 2."""
import os

import pytest

import _cases as cases
from conftest import GOLDEN
from zkcnn_b200._binding import CHECK_PREDICATES, PREFETCH_NEXT, REAL_GENERATORS, Session


def test_lenet_synthetic(emu_host, synthetic_inputs):
    st = cases.prove_and_compare(emu_host, "lenet", "", 1, synthetic_inputs["lenet_syn"], 3, CHECK_PREDICATES, "lenet_syn_p1_seed3", GOLDEN)
    assert st["ok"] == 1 and st["n_layers"] == 24 and st["gpu_launches"] > 0


def test_lenet_synthetic_real_generators(emu_host, synthetic_inputs):
    cases.prove_and_compare(emu_host, "lenet", "", 1, synthetic_inputs["lenet_syn"], 3, REAL_GENERATORS, "lenet_syn_p1_seed3_realgens", GOLDEN)


def test_lenet_two_pictures_fft_path(emu_host, synthetic_inputs):
    # FFT-convolution path (pic_cnt = 2): whole transcript AND the bookkeeping tables after every Init* call (K4b / K5b / K2 set-up)
    st = cases.prove_and_compare(emu_host, "lenet", "", 2, synthetic_inputs["lenet_syn"], 4, 0, "lenet_syn_p2_seed4", GOLDEN, tables=True)
    assert st["ok"] == 1


def test_fft_path_with_other_kernel_variants(emu_host, synthetic_inputs, monkeypatch):
    """the same proof with the kernel selection pushed the other way (ZK_TUNABLES is read when the prover's context is created):
    K5b rows split over three threads, K2 grids capped (several iterations per thread) and factored everywhere, a short fused tail"""
    monkeypatch.setenv("ZK_TUNABLES", "axpy_splits=3,cubic_max_grid=7,cubic_factored_min_iters=1,tail_max_entries=64")
    from zkcnn_b200._binding import PROVER_ONLY
    st = cases.prove_and_compare(emu_host, "lenet", "", 2, synthetic_inputs["lenet_syn"], 4, PROVER_ONLY, "lenet_syn_p2_seed4", GOLDEN)
    assert st["ok"] == 1


def test_small_vgg_naive_conv(emu_host, synthetic_inputs):
    st = cases.prove_and_compare(emu_host, "vgg", synthetic_inputs["smallvgg_config"], 1, synthetic_inputs["smallvgg"], 7, 0,
                                 "smallvgg_p1_seed7", GOLDEN)
    assert st["ok"] == 1


def test_small_vgg_fft_conv(emu_host, synthetic_inputs):
    st = cases.prove_and_compare(emu_host, "vgg", synthetic_inputs["smallvgg_config"], 2, synthetic_inputs["smallvgg"], 7, 0,
                                 "smallvgg_p2_seed7", GOLDEN, tables=True)
    assert st["ok"] == 1


def test_round_by_round_driver_gives_the_same_transcript(emu_host, synthetic_inputs):
    """the default driver hands the device a whole phase of challenges at once (zk_sumcheck_update_batch, fused tail); with
    ZKH_ROUND_BY_ROUND it makes the reference's one call per round -- same messages, same order.  ZKH_PROVER_ONLY skips the verifier's
    predicates without changing the transcript."""
    from zkcnn_b200._binding import PROVER_ONLY, ROUND_BY_ROUND
    st = cases.prove_and_compare(emu_host, "lenet", "", 2, synthetic_inputs["lenet_syn"], 4, PROVER_ONLY | ROUND_BY_ROUND, "lenet_syn_p2_seed4", GOLDEN)
    assert st["ok"] == 1 and st["checks"] == 1


@pytest.mark.parametrize("bits", [6, 7, 8])
def test_commitment_digit_widths_give_the_same_transcript(emu_host, synthetic_inputs, monkeypatch, bits):
    """the commitment's small-multiples path with 63 / 127 / 255 multiples per generator (zkh_build picks the width from the witness;
    ZKH_MSM_DIGIT_BITS forces it): real generators, so the commitment points are in the transcript"""
    from zkcnn_b200._binding import PROVER_ONLY, REAL_GENERATORS
    monkeypatch.setenv("ZKH_MSM_DIGIT_BITS", str(bits))
    cases.prove_and_compare(emu_host, "vgg", synthetic_inputs["smallvgg_config"], 1, synthetic_inputs["smallvgg"], 8, REAL_GENERATORS | PROVER_ONLY,
                            "smallvgg_p1_seed8_realgens", GOLDEN)


def test_shipped_lenet_image(emu_host, mnist_input):
    """BASELINE config 1: the reference's own MNIST demo input (script/demo_lenet.sh)"""
    from zkcnn_b200._binding import HOST_PREDICATES
    st = cases.prove_and_compare(emu_host, "lenet", "", 1, mnist_input, 1, HOST_PREDICATES, "lenet_p1_seed1", GOLDEN)   # predicates on host threads here,
    assert st["ok"] == 1 and st["checks"] == 15                                                                          # on the device everywhere else


def test_witness_lifecycle(emu_host, synthetic_inputs, mnist_input):
    """upload / prefetch / resident witness / rebuild: a proof that prefetches the NEXT witness (double-buffered upload), a proof that
    adopts the prefetched copy, a proof on the resident witness (no upload), then a NEW input and a rebuild -- the stale shadow copy
    and the stale resident witness must not be used -- and a resident proof of the new witness"""
    from zkcnn_b200._binding import PROVER_ONLY, WITNESS_RESIDENT
    g_syn = open(os.path.join(GOLDEN, "lenet_syn_p1_seed3.transcript.bin"), "rb").read()
    g_mnist = open(os.path.join(GOLDEN, "lenet_p1_seed1.transcript.bin"), "rb").read()
    with Session(emu_host, "lenet", "", 1) as s:
        s.input_file(synthetic_inputs["lenet_syn"])
        s.build()
        a = s.prove(3, PREFETCH_NEXT | PROVER_ONLY)     # uploads, then starts the copy for the next proof
        assert a["ok"] == 1 and a["h2d_bytes"] > 0 and s.proof() == g_syn
        b = s.prove(9, PROVER_ONLY)                     # adopts the prefetched copy
        assert b["ok"] == 1 and b["h2d_bytes"] >= a["h2d_bytes"] and s.proof() != g_syn
        s.prefetch_witness()                            # explicit call; the rebuild below must discard it
        s.input_file(mnist_input)
        s.build()
        c = s.prove(1, WITNESS_RESIDENT | PROVER_ONLY)
        assert c["ok"] == 1 and c["h2d_bytes"] > 0 and s.proof() == g_mnist
        d = s.prove(1, WITNESS_RESIDENT | PROVER_ONLY)
        assert d["ok"] == 1 and d["h2d_bytes"] == 0 and s.proof() == g_mnist


def test_challenge_sources(emu_host, synthetic_inputs):
    """the operating system's CSPRNG (the reference's Fr::setByCSPRNG) and Fiat-Shamir challenges, both under full verification"""
    from zkcnn_b200._binding import CSPRNG_CHALLENGES, FIAT_SHAMIR
    g_syn = open(os.path.join(GOLDEN, "lenet_syn_p1_seed3.transcript.bin"), "rb").read()
    with Session(emu_host, "lenet", "", 1) as s:
        s.input_file(synthetic_inputs["lenet_syn"])
        s.build()
        a = s.prove(3, CSPRNG_CHALLENGES)
        pa = s.proof()
        assert a["ok"] == 1 and a["checks"] == 15 and pa != g_syn and len(pa) == len(g_syn)
        # Fiat-Shamir: accepted, reproducible for a seed, every challenge drawn after the message it answers (rounds one by one)
        c = s.prove(3, FIAT_SHAMIR)
        pc = s.proof()
        d = s.prove(3, FIAT_SHAMIR)
        assert c["ok"] == 1 and c["checks"] == 15 and s.proof() == pc and d["fnv1a"] == c["fnv1a"] and pc != g_syn and pc != pa


def test_device_witness_generation(emu_lib, emu_host, synthetic_inputs):
    """witness generation on the device (gate evaluation, NTT of the FFT layers, DOT_PROD products, ReLU / max-pool decompositions):
    LeNet with two pictures (FFT-convolution path) and the small VGG with naive convolutions"""
    import numpy as np
    sys_path = os.path.join(os.path.dirname(GOLDEN), "..", "tools")
    import sys
    sys.path.insert(0, sys_path)
    import gen_synthetic_input as gen
    lenet = gen.generate("lenet", 11).astype(np.float64)
    other = np.random.default_rng(1).random(1024)
    assert cases.device_witness_case(emu_lib, emu_host, "lenet", "", 2, lenet, 1024, "lenet_syn_p2_seed4", GOLDEN, 4, other) == 1
    cfg = synthetic_inputs["smallvgg_config"]
    vgg = gen.generate("vgg11", 5, cfg).astype(np.float64)
    nudged = vgg[:3072].copy()
    nudged[100:110] *= 0.999      # a slightly different picture: same quantisation decisions
    cases.device_witness_case(emu_lib, emu_host, "vgg", cfg, 1, vgg, 3072, "smallvgg_p1_seed7", GOLDEN, 7, nudged)


def test_tampered_witness_is_rejected(emu_host, synthetic_inputs, tmp_path):
    """soundness smoke: a different image gives a different (still accepted) proof; bad arguments are errors"""
    from zkcnn_b200._binding import ZkError
    with pytest.raises(ZkError):
        Session(emu_host, "resnet")
    with Session(emu_host, "lenet", "", 1) as s:
        with pytest.raises(ZkError):
            s.prove(1)                      # not built yet
        with pytest.raises(ZkError):
            s.input_file(tmp_path / "missing.csv")
