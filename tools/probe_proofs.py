#!/usr/bin/env python3
"""Prove vgg11 a few times and print the wall time per proof.  usage: probe_proofs.py resident|prefetch [n]
ZKH_TRACE=1 adds the wall time of the stages of each proof, ZK_TRACE=1 the time per C-ABI entry point (at exit)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import gen_synthetic_input as gen
import zkcnn_b200
from zkcnn_b200 import PROVER_ONLY, REAL_GENERATORS, WITNESS_RESIDENT, PREFETCH_NEXT
mode = sys.argv[1]
values = gen.generate("vgg11")
s = zkcnn_b200.session("vgg", "64 M 128 M 256 256 M 512 512 M 512 512 M", 1, device=0)
s.input_values(values.astype(np.float64)); s.build()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
nvtx = None
if os.environ.get("PROBE_NVTX"):      # ncu --nvtx --nvtx-include "proof/": profile exactly the launches of the last proof
    import torch
    nvtx = torch.cuda.nvtx
for i in range(n):
    fl = REAL_GENERATORS | PROVER_ONLY | (WITNESS_RESIDENT if mode == "resident" else PREFETCH_NEXT)
    if nvtx and i == n - 1:
        nvtx.range_push("proof")
    t0 = time.perf_counter(); st = s.prove(100 + i, fl); print(mode, i, f"{(time.perf_counter()-t0)*1e3:.2f} ms launches {st['gpu_launches']}", file=sys.stderr)
    if nvtx and i == n - 1:
        nvtx.range_pop()
s.close()
