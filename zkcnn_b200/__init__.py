"""zkcnn_b200: B200-native zkCNN prover (GKR sumcheck + Hyrax over BLS12-381).

The compute path is libzkcnn_b200.so (hand-written sm_100a CUDA behind the C ABI of include/zkcnn_b200.h); this package
only loads it.  There is no CPU fallback: without the built library or without a CUDA device every entry point raises.
"""
import os

from ._binding import (CHECKED_ALL, CHECK_PREDICATES, CSPRNG_CHALLENGES, FIAT_SHAMIR, FIXED_GENERATORS, HOST_PREDICATES, PROVER_ONLY, NO_HASH, PREFETCH_NEXT, REAL_GENERATORS, ROUND_BY_ROUND, WITNESS_RESIDENT, Context, HostLib, Lib, Session, ZkError,
                       PROF_CLASSES, fr_from_words, fr_to_words, g1_from_words, g1_to_words)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libzkcnn_b200.so")
HOST_LIB_PATH = os.path.join(_HERE, "lib", "libzkcnn_host.so")

_lib = None
_host = None


def load():
    """the CUDA library; raises if it is not built or no GPU is visible"""
    global _lib
    if _lib is None:
        lib = Lib(LIB_PATH)
        if lib.device_count() < 1:
            raise ZkError("no CUDA device visible: zkcnn_b200 has no CPU fallback")
        _lib = lib
    return _lib


def load_host():
    global _host
    if _host is None:
        load()
        _host = HostLib(HOST_LIB_PATH)
    return _host


def context(device=0):
    return Context(load(), device)


def session(model, network="", pic_cnt=1, device=0):
    return Session(load_host(), model, network, pic_cnt, device)
