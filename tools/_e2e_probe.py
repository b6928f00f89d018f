import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import gen_synthetic_input as gen
import zkcnn_b200
from zkcnn_b200 import REAL_GENERATORS, WITNESS_RESIDENT, PREFETCH_NEXT
values = gen.generate("vgg11")
s = zkcnn_b200.session("vgg", "64 M 128 M 256 256 M 512 512 M 512 512 M", 1, device=0)
s.input_values(values.astype(np.float64)); s.build()
for i in range(3): s.prove(100 + i, REAL_GENERATORS | WITNESS_RESIDENT)
def show(tag, st, t): print(f"{tag:28s} wall {t*1e3:7.2f} ms  prove_s {st['prove_s']*1e3:7.2f} poly_s {st['poly_s']*1e3:6.2f} upload_s {st['upload_s']*1e3:6.2f} wall_s {st['wall_s']*1e3:7.2f}")
for rep in range(3):
    t0 = time.perf_counter(); st = s.prove(7, REAL_GENERATORS | WITNESS_RESIDENT); show("resident", st, time.perf_counter() - t0)
for rep in range(3):
    t0 = time.perf_counter(); st = s.prove(7, REAL_GENERATORS); show("sync upload", st, time.perf_counter() - t0)
for rep in range(4):
    t0 = time.perf_counter(); st = s.prove(7, REAL_GENERATORS | PREFETCH_NEXT); show("prefetch next", st, time.perf_counter() - t0)
s.close()
