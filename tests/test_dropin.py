"""The drop-in check, automated: the REFERENCE's own verifier, circuit builder and model zoo (compiled from /root/reference by
oracle/Makefile with -include zkcnn_b200/host/dropin.hpp) drive the zkcnn_b200 prover through the reference's public prover
interface (src/verifier.cpp:118-373 calls class prover / class polyProver); the transcript must be byte-identical to the one
the unmodified reference produced (tests/golden).  Two configurations:
  dropin_run           GPU sumcheck + GPU Hyrax                                   (BASELINE config 3's boundary)
  dropin_run_cpuhyrax  GPU sumcheck + the reference's own CPU polyProver           (BASELINE config 2)
each on the CUDA library (-m gpu) and on the test-only emulator build of the same kernels (CPU suite)."""
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

REF = os.path.join(ROOT, "oracle", "_ref")


def run_dropin(binary, model_args, seed, golden, tmp_path, extra=()):
    exe = os.path.join(REF, binary)
    if not os.access(exe, os.X_OK):
        pytest.skip(f"{binary} not built in this snapshot (make -C oracle dropin dropin_emu; needs /root/reference)")
    out = tmp_path / (binary + ".bin")
    r = subprocess.run([exe, *model_args, "1", str(seed), "--transcript", str(out), *extra], capture_output=True, text=True, timeout=600)
    res = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")]
    assert res, r.stderr[-500:]
    f = dict(zip(*[iter(res[0].split()[1:21])] * 2))
    want = open(os.path.join(GOLDEN, golden + ".transcript.bin"), "rb").read()
    ref = dict(zip(*[iter(open(os.path.join(GOLDEN, golden + ".result.txt")).read().split()[1:])] * 2))
    assert out.read_bytes() == want, f"{binary}: transcript differs from the reference's ({f})"
    assert f["ok"] == ref["ok"] and f["n_g1"] == ref["n_g1"] and f["fnv"] == ref["fnv"] and f["challenges"] == ref["challenges"]
    return f


@pytest.mark.parametrize("binary", ["dropin_run_emu", "dropin_run_cpuhyrax_emu"])
def test_dropin_on_the_emulator(binary, synthetic_inputs, tmp_path):
    run_dropin(binary, ["lenet", synthetic_inputs["lenet_syn"], "x"], 3, "lenet_syn_p1_seed3", tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("binary", ["dropin_run", "dropin_run_cpuhyrax"])
def test_dropin_on_the_gpu(binary, synthetic_inputs, mnist_input, tmp_path):
    # BASELINE config 1 / 2 input: the reference's shipped MNIST picture; then the synthetic LeNet input and, for the full GPU
    # configuration, real (non-degenerate) generators
    run_dropin(binary, ["lenet", mnist_input, "x"], 1, "lenet_p1_seed1", tmp_path)
    f = run_dropin(binary, ["lenet", synthetic_inputs["lenet_syn"], "x"], 3, "lenet_syn_p1_seed3", tmp_path)
    assert int(f["n_g1"]) > 0
    if binary == "dropin_run":
        run_dropin(binary, ["lenet", synthetic_inputs["lenet_syn"], "x"], 3, "lenet_syn_p1_seed3_realgens", tmp_path, extra=("--gens", "real"))
        run_dropin(binary, ["vgg", synthetic_inputs["smallvgg"], "x", synthetic_inputs["smallvgg"] + ".config"], 7, "smallvgg_p1_seed7", tmp_path)
