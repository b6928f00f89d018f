import hashlib
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
EMU_DIR = os.path.join(ROOT, "tests", "emu", "_build")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make(*targets):
    subprocess.run(["make", "-C", ROOT, *targets], check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def emu_lib():
    """TEST-ONLY: the kernel sources compiled for the CUDA emulator (host threads).  Never used by the product."""
    from zkcnn_b200._binding import Lib
    _make("emu")
    return Lib(os.path.join(EMU_DIR, "libzkcnn_b200_emu.so"))


@pytest.fixture(scope="session")
def emu_host(emu_lib):
    from zkcnn_b200._binding import HostLib
    _make("host_emu")
    return HostLib(os.path.join(EMU_DIR, "libzkcnn_host_emu.so"))


@pytest.fixture(scope="session")
def gpu_lib():
    import zkcnn_b200
    return zkcnn_b200.load()       # raises without the built library or without a GPU: no fallback


@pytest.fixture(scope="session")
def gpu_host(gpu_lib):
    import zkcnn_b200
    return zkcnn_b200.load_host()


@pytest.fixture(scope="session")
def kat():
    return json.load(open(os.path.join(GOLDEN, "kat.json")))


@pytest.fixture(scope="session")
def synthetic_inputs(tmp_path_factory):
    """seeded synthetic inputs of tools/gen_synthetic_input.py, checked against the checksums recorded when the golden
    transcripts were minted (tests/golden/manifest.json)"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_synthetic_input as gen
    manifest = json.load(open(os.path.join(GOLDEN, "manifest.json")))
    d = tmp_path_factory.mktemp("inputs")
    out = {"smallvgg_config": manifest["smallvgg_config"]}
    for name, model, seed, cfg in (("lenet_syn", "lenet", 11, None), ("smallvgg", "vgg11", 5, manifest["smallvgg_config"])):
        path = os.path.join(d, name + ".csv")
        gen.write_text(gen.generate(model, seed, cfg), path)
        if cfg:
            open(path + ".config", "w").write(cfg + "\n")     # the network description file of the reference's `vgg` model class
        digest = hashlib.sha256(open(path, "rb").read()).hexdigest()
        assert digest == manifest[name], f"synthetic input {name} differs from the one the goldens were made with (numpy RNG changed?)"
        out[name] = path
    return out


@pytest.fixture(scope="session")
def mnist_input(tmp_path_factory):
    """BASELINE config 1: the reference's shipped LeNet5 / MNIST input (62 730 decimals: one 32x32 image, then the weights;
    data/lenet5.mnist.relu.max/lenet5.mnist.relu.max-1-images-weights-qint8.csv of the reference's data.tar.gz,
    script/demo_lenet.sh:14-18), committed gzip-compressed as a test fixture"""
    import gzip
    path = os.path.join(tmp_path_factory.mktemp("mnist"), "lenet5_mnist.csv")
    with gzip.open(os.path.join(GOLDEN, "lenet5_mnist_input.csv.gz"), "rb") as f, open(path, "wb") as g:
        g.write(f.read())
    return path


def golden_bytes(name):
    return open(os.path.join(GOLDEN, name + ".transcript.bin"), "rb").read()


def golden_text(name, kind="circuit"):
    return open(os.path.join(GOLDEN, f"{name}.{kind}.txt")).read()
