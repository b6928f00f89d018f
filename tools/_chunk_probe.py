import sys, os, time, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import gen_synthetic_input as gen
import zkcnn_b200
from zkcnn_b200 import REAL_GENERATORS, WITNESS_RESIDENT
lib = zkcnn_b200.load()
values = gen.generate("vgg11")
s = zkcnn_b200.session("vgg", "64 M 128 M 256 256 M 512 512 M 512 512 M", 1, device=0)
s.input_values(values.astype(np.float64)); s.build()
for i in range(3): s.prove(100 + i, REAL_GENERATORS | WITNESS_RESIDENT)
ctx = s.context_handle()
ref = None
for chunk in (2048, 1024, 512, 256, 2048):
    lib.dll.zk_set_tunable(ctx, b"msm_few_rows_chunk", chunk)
    best = 1e9
    for rep in range(4):
        t0 = time.perf_counter(); st = s.prove(7, REAL_GENERATORS | WITNESS_RESIDENT); best = min(best, time.perf_counter() - t0)
        assert st["ok"] == 1
        if ref is None: ref = st["fnv1a"]
        assert st["fnv1a"] == ref
    print(f"msm_few_rows_chunk {chunk:5d}: best {best*1e3:.2f} ms per proof", flush=True)
s.close()
