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


def load_input(path, cache=True, dist=None, device=None):
    """The model's input stream (picture + weights, the reference's text format) as a float64 array -- SURVEY section 8 f-2:
      * parsed ONCE: a binary cache (<path>.f64.npy, keyed by the text file's size and mtime) replaces the 124 MB text parse of every run;
      * with `dist` (an initialised torch.distributed group) only rank 0 touches the file system: the array is broadcast to the other ranks
        (NCCL over NVLink when `device` is a CUDA device), instead of every rank parsing its own copy.
    The quantised weights themselves are cached where they are used: they stay resident in val[0] on the device and a new picture regenerates
    only the picture-dependent part of the witness (Session.prove_image)."""
    import numpy as np
    rank = dist.get_rank() if dist is not None else 0
    values = None
    if rank == 0:
        st = os.stat(path)
        cpath = str(path) + ".f64.npy"
        meta = np.array([st.st_size, int(st.st_mtime)], dtype=np.float64)
        if cache and os.path.exists(cpath):
            try:
                arr = np.load(cpath)
                if arr.dtype == np.float64 and len(arr) > 2 and (arr[:2] == meta).all():
                    values = arr[2:]
            except Exception:   # noqa: BLE001
                values = None
        if values is None:
            values = load_host().parse_numbers(path)
            if cache:
                try:
                    np.save(cpath, np.concatenate([meta, values]))
                except OSError:
                    pass
    if dist is not None and dist.get_world_size() > 1:
        import torch
        dev = device if device is not None else "cpu"
        n = torch.tensor([len(values) if rank == 0 else 0], dtype=torch.int64, device=dev)
        dist.broadcast(n, 0)
        t = torch.from_numpy(np.ascontiguousarray(values)).to(dev) if rank == 0 else torch.empty(int(n.item()), dtype=torch.float64, device=dev)
        dist.broadcast(t, 0)
        values = t.cpu().numpy()
    return values
