#!/usr/bin/env python3
"""M vgg11 provers in flight on one GPU: proofs/s and the host CPU seconds one proof costs.
usage: probe_inflight.py M [K=4 proofs per prover] [image|resident]
ZK_HOST_WAIT=spin|yield|block selects how the prover threads wait for the GPU (csrc/rt.hpp); run under `taskset -c 0-3` to see a host
with fewer cores than waiting threads (the per-GPU share of an 8-GPU box)."""
import os, sys, threading, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import gen_synthetic_input as gen
import zkcnn_b200
from zkcnn_b200 import PROVER_ONLY, REAL_GENERATORS, WITNESS_RESIDENT, NO_HASH

M = int(sys.argv[1])
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mode = sys.argv[3] if len(sys.argv) > 3 else "resident"
NET = "64 M 128 M 256 256 M 512 512 M 512 512 M"
sessions = []
for m in range(M):
    s = zkcnn_b200.session("vgg", NET, 1, device=0)
    s.input_values(gen.generate("vgg11", 20211115 if m == 0 else 7000 + m).astype(np.float64))
    sessions.append(s)
th = [threading.Thread(target=s.build) for s in sessions]
[t.start() for t in th]; [t.join() for t in th]
flags = REAL_GENERATORS | NO_HASH | PROVER_ONLY
pics = [gen.generate("vgg11", 20211115 if m == 0 else 7000 + m)[:3 * 32 * 32] for m in range(M)]


def run(k0, n):
    def work(m):
        for k in range(n):
            if mode == "image":
                st = sessions[m].prove_image(pics[m], k0 + k, flags)
            else:
                st = sessions[m].prove(k0 + k, flags | WITNESS_RESIDENT)
            assert st["ok"] == 1
    th = [threading.Thread(target=work, args=(m,)) for m in range(M)]
    [t.start() for t in th]; [t.join() for t in th]


run(100, 2)
c0, t0 = os.times(), time.perf_counter()
run(200, K)
c1, t1 = os.times(), time.perf_counter()
cpu = (c1.user - c0.user) + (c1.system - c0.system)
n = M * K
print(f"wait={os.environ.get('ZK_HOST_WAIT', 'spin')} cores={len(os.sched_getaffinity(0))} M={M} mode={mode}: {n / (t1 - t0):.2f} proofs/s, "
      f"{(t1 - t0) / n * 1e3:.2f} ms/proof, host CPU {cpu / n * 1e3:.1f} ms per proof ({cpu / (t1 - t0):.2f} cores busy)")
for s in sessions:
    s.close()
