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
    st = cases.prove_and_compare(emu_host, "lenet", "", 2, synthetic_inputs["lenet_syn"], 4, CHECK_PREDICATES, "lenet_syn_p2_seed4", GOLDEN)
    assert st["ok"] == 1


def test_small_vgg_naive_conv(emu_host, synthetic_inputs):
    st = cases.prove_and_compare(emu_host, "vgg", synthetic_inputs["smallvgg_config"], 1, synthetic_inputs["smallvgg"], 7, CHECK_PREDICATES,
                                 "smallvgg_p1_seed7", GOLDEN)
    assert st["ok"] == 1


def test_small_vgg_fft_conv(emu_host, synthetic_inputs):
    st = cases.prove_and_compare(emu_host, "vgg", synthetic_inputs["smallvgg_config"], 2, synthetic_inputs["smallvgg"], 7, 0,
                                 "smallvgg_p2_seed7", GOLDEN)
    assert st["ok"] == 1


@pytest.mark.parametrize("model,net,pics,inp,seed,golden", [
    ("lenet", "", 2, "lenet_syn", 4, "lenet_syn_p2_seed4"),
    ("vgg", "small", 2, "smallvgg", 7, "smallvgg_p2_seed7"),
])
def test_init_tables_against_the_reference(emu_host, synthetic_inputs, model, net, pics, inp, seed, golden):
    net = synthetic_inputs["smallvgg_config"] if net == "small" else net
    cases.tables_and_compare(emu_host, model, net, pics, synthetic_inputs[inp], seed, golden, GOLDEN)


def test_round_by_round_driver_gives_the_same_transcript(emu_host, synthetic_inputs):
    """the default driver hands the device a whole phase of challenges at once (zk_sumcheck_update_batch); with
    ZKH_ROUND_BY_ROUND it makes the reference's one call per round -- same messages, same order"""
    from zkcnn_b200._binding import ROUND_BY_ROUND
    cases.prove_and_compare(emu_host, "lenet", "", 1, synthetic_inputs["lenet_syn"], 3, CHECK_PREDICATES | ROUND_BY_ROUND, "lenet_syn_p1_seed3", GOLDEN)
    cases.prove_and_compare(emu_host, "vgg", synthetic_inputs["smallvgg_config"], 1, synthetic_inputs["smallvgg"], 7, ROUND_BY_ROUND,
                            "smallvgg_p1_seed7", GOLDEN)


def test_shipped_lenet_image(emu_host, mnist_input):
    """BASELINE config 1: the reference's own MNIST demo input (script/demo_lenet.sh), degenerate and real generators"""
    st = cases.prove_and_compare(emu_host, "lenet", "", 1, mnist_input, 1, 0, "lenet_p1_seed1", GOLDEN)
    assert st["ok"] == 1 and st["checks"] == 15
    cases.prove_and_compare(emu_host, "lenet", "", 1, mnist_input, 1, REAL_GENERATORS, "lenet_p1_seed1_realgens", GOLDEN)


def test_repeat_proofs_and_resident_witness(emu_host, synthetic_inputs):
    """second proof on the same session: new seed -> new transcript; same seed with the witness kept on the device ->
    identical transcript, no upload"""
    from zkcnn_b200._binding import WITNESS_RESIDENT
    with Session(emu_host, "lenet", "", 1) as s:
        s.input_file(synthetic_inputs["lenet_syn"])
        s.build()
        a = s.prove(3, 0)
        pa = s.proof()
        b = s.prove(9, WITNESS_RESIDENT)
        pb = s.proof()
        c = s.prove(3, WITNESS_RESIDENT)
        pc = s.proof()
    assert a["ok"] and b["ok"] and c["ok"]
    assert a["h2d_bytes"] > 0 and b["h2d_bytes"] == 0 and c["h2d_bytes"] == 0
    assert pa == pc and pa != pb
    assert pa == open(os.path.join(GOLDEN, "lenet_syn_p1_seed3.transcript.bin"), "rb").read()


def test_prefetched_witness_gives_the_same_proofs(emu_host, synthetic_inputs):
    """double-buffered upload: the witness of the next proof is copied while the current one runs; transcripts unchanged"""
    with Session(emu_host, "lenet", "", 1) as s:
        s.input_file(synthetic_inputs["lenet_syn"])
        s.build()
        a = s.prove(3, PREFETCH_NEXT)     # uploads, then starts the copy for the next proof
        pa = s.proof()
        b = s.prove(9, PREFETCH_NEXT)     # adopts the prefetched copy, starts the next one
        s.prefetch_witness()              # explicit call: replaces the pending copy
        c = s.prove(3, 0)
        pc = s.proof()
    assert a["ok"] and b["ok"] and c["ok"] and pa == pc
    # (the first proof pads val[0] to a power of two on the host, src/prover.cpp:504-508: later copies are that much longer)
    assert 0 < a["h2d_bytes"] <= b["h2d_bytes"] <= c["h2d_bytes"]
    assert pa == open(os.path.join(GOLDEN, "lenet_syn_p1_seed3.transcript.bin"), "rb").read()


def test_rebuild_after_prefetch(emu_host, synthetic_inputs, mnist_input):
    """a proof that prefetched the NEXT witness, then a new input and a rebuild: the stale shadow copy must not be adopted
    (the circuit upload clears the device layers), with and without the resident-witness flag"""
    from zkcnn_b200._binding import WITNESS_RESIDENT
    with Session(emu_host, "lenet", "", 1) as s:
        s.input_file(synthetic_inputs["lenet_syn"])
        s.build()
        a = s.prove(3, PREFETCH_NEXT)
        assert a["ok"] == 1 and s.proof() == open(os.path.join(GOLDEN, "lenet_syn_p1_seed3.transcript.bin"), "rb").read()
        s.input_file(mnist_input)
        s.build()
        b = s.prove(1, WITNESS_RESIDENT)
        assert b["ok"] == 1 and b["h2d_bytes"] > 0
        assert s.proof() == open(os.path.join(GOLDEN, "lenet_p1_seed1.transcript.bin"), "rb").read()
        c = s.prove(1, WITNESS_RESIDENT)
        assert c["ok"] == 1 and c["h2d_bytes"] == 0 and s.proof() == open(os.path.join(GOLDEN, "lenet_p1_seed1.transcript.bin"), "rb").read()


def test_challenge_sources(emu_host, synthetic_inputs):
    """seeded (default), the operating system's CSPRNG (the reference's Fr::setByCSPRNG) and Fiat-Shamir challenges"""
    from zkcnn_b200._binding import CSPRNG_CHALLENGES, FIAT_SHAMIR, PROVER_ONLY
    with Session(emu_host, "lenet", "", 1) as s:
        s.input_file(synthetic_inputs["lenet_syn"])
        s.build()
        a = s.prove(3, CSPRNG_CHALLENGES)
        pa = s.proof()
        b = s.prove(3, CSPRNG_CHALLENGES)
        pb = s.proof()
        assert a["ok"] == 1 and b["ok"] == 1 and a["checks"] == 15 and pa != pb and len(pa) == len(pb)
        # Fiat-Shamir: accepted under full verification, reproducible, seed-separated, and every challenge drawn after the message it answers
        c = s.prove(3, FIAT_SHAMIR)
        pc = s.proof()
        d = s.prove(3, FIAT_SHAMIR)
        e = s.prove(4, FIAT_SHAMIR)
        assert c["ok"] == 1 and c["checks"] == 15 and s.proof() != pc and d["fnv1a"] == c["fnv1a"] and e["fnv1a"] != c["fnv1a"]
        assert pc != open(os.path.join(GOLDEN, "lenet_syn_p1_seed3.transcript.bin"), "rb").read()
        p = s.prove(3, PROVER_ONLY)
        assert p["ok"] == 1 and p["checks"] == 1 and s.proof() == open(os.path.join(GOLDEN, "lenet_syn_p1_seed3.transcript.bin"), "rb").read()


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
