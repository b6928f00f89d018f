#!/usr/bin/env python3
"""Prove a batched (pic_cnt > 1, FFT-convolution path) VGG a few times; print wall time per proof and the per-kernel-class
device times (zk_profile_*).  usage: probe_fft_path.py vgg11|vgg16 PICS [n]     (ZKH_TRACE=1 / ZK_TRACE=1 as in probe_proofs.py)"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import gen_synthetic_input as gen
import zkcnn_b200
from zkcnn_b200 import PROF_CLASSES, PROVER_ONLY, REAL_GENERATORS, WITNESS_RESIDENT
model, pics = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = zkcnn_b200.load()
s = zkcnn_b200.session("vgg", gen.CONFIGS[model], pics, device=0)
s.input_values(gen.generate(model).astype(np.float64))
t0 = time.perf_counter(); s.build(); print(f"build {time.perf_counter() - t0:.1f} s", file=sys.stderr)
t0 = time.perf_counter(); st = s.prove(1, 0); print(f"first proof (upload + schedules) {time.perf_counter() - t0:.2f} s fnv {st['fnv1a']:016x} ok {st['ok']} launches {st['gpu_launches']}", file=sys.stderr)
nvtx = None
if os.environ.get("PROBE_NVTX"):      # ncu --nvtx --nvtx-include "proof/": profile exactly the launches of the last un-instrumented proof
    import torch
    nvtx = torch.cuda.nvtx
for i in range(n):
    if nvtx and i == n - 1:
        nvtx.range_push("proof")
    t0 = time.perf_counter(); st = s.prove(100 + i, REAL_GENERATORS | WITNESS_RESIDENT | PROVER_ONLY)
    if nvtx and i == n - 1:
        nvtx.range_pop()
    print(f"resident {i}: {(time.perf_counter() - t0 - 0) * 1e3:.2f} ms  prove_s {st['prove_s']:.4f} poly_s {st['poly_s']:.4f} launches {st['gpu_launches']}", file=sys.stderr)
ctx = s.context_handle()
lib.dll.zk_profile_enable(ctx, 1)
t0 = time.perf_counter(); s.prove(100, REAL_GENERATORS | WITNESS_RESIDENT | PROVER_ONLY); wall = time.perf_counter() - t0
out = {"model": model, "pics": pics, "profiled_wall_ms": round(wall * 1e3, 2)}
for k, name in enumerate(PROF_CLASSES):
    ms, cnt, b = C.c_double(0), C.c_uint64(0), C.c_uint64(0)
    lib.dll.zk_profile_get(ctx, k, C.byref(ms), C.byref(cnt), C.byref(b))
    out[name] = {"ms": round(ms.value, 3), "launches": cnt.value, "GB": round(b.value / 1e9, 3), "GB/s": round(b.value / 1e6 / ms.value, 1) if ms.value else 0}
lib.dll.zk_profile_enable(ctx, 0)
print(json.dumps(out))
s.close()
