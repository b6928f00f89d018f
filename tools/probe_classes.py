#!/usr/bin/env python3
"""Per-kernel-class device time of ONE vgg11 proof (zk_profile_*), for A/B runs under ZK_TUNABLES.  usage: probe_classes.py [pics]
ZK_PROF_KERNELS=1 adds one line per kernel name (launches, device ms) on stderr."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import gen_synthetic_input as gen
import zkcnn_b200
from zkcnn_b200 import PROF_CLASSES, PROVER_ONLY, REAL_GENERATORS, WITNESS_RESIDENT
pics = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lib = zkcnn_b200.load()
s = zkcnn_b200.session("vgg", gen.CONFIGS["vgg11"], pics, device=0)
s.input_values(gen.generate("vgg11").astype(np.float64)); s.build()
fl = REAL_GENERATORS | WITNESS_RESIDENT | PROVER_ONLY
for i in range(3):
    t0 = time.perf_counter(); s.prove(100 + i, fl); wall = time.perf_counter() - t0
ctx = s.context_handle()
lib.dll.zk_profile_enable(ctx, 1)
s.prove(100, fl)
out = {"tunables": os.environ.get("ZK_TUNABLES", ""), "wall_ms": round(wall * 1e3, 2)}
for k, name in enumerate(PROF_CLASSES):
    ms, cnt, b = C.c_double(0), C.c_uint64(0), C.c_uint64(0)
    lib.dll.zk_profile_get(ctx, k, C.byref(ms), C.byref(cnt), C.byref(b))
    out[name] = [round(ms.value, 3), cnt.value]
lib.dll.zk_profile_enable(ctx, 0)
print(json.dumps(out))
s.close()
