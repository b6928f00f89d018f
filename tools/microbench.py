#!/usr/bin/env python3
"""Device-timed micro-benchmarks of the two headline kernels (CUDA events inside the library, inputs resident in HBM).
usage: microbench.py fold [bits] [iters] | cubic [bits] [iters] | msm [log_rows] [log_cols] [mix] [iters]      -- also the target of ncu captures"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import zkcnn_b200

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "fold"
    a = [int(x) for x in sys.argv[2:]]
    with zkcnn_b200.context(0) as c:
        if what == "fold":
            bits, iters = (a + [24, 20])[:2] if len(a) < 2 else a[:2]
            ms = c.bench_fold(bits, iters, True)
            gbs = 96 * (1 << bits) / 1e9 / (ms / 1e3)
            print(json.dumps({"kernel": "k_round_quad", "bits": bits, "ms": ms, "algorithmic_GB/s": gbs, "frac_of_measured_hbm": gbs / PEAK}))
        elif what == "cubic":
            bits, iters = (a + [24, 10][len(a):])[:2]
            ms = c.bench_cubic(bits, 7, iters)
            gbs = 96 * (1 << bits) / 1e9 / (ms / 1e3)
            print(json.dumps({"kernel": "k_round_cubic_tma", "bits": bits, "ms": ms, "algorithmic_GB/s": gbs, "frac_of_measured_hbm": gbs / PEAK}))
        else:
            lr, lc, mix, iters = (a + [12, 12, 2, 3])[len(a):] if False else (a + [12, 12, 2, 3][len(a):])[:4]
            ms = c.bench_msm(lr, lc, mix, iters)
            b = 32 * (1 << (lr + lc)) + 96 * (1 << lc) + 144 * (1 << lr)
            print(json.dumps({"kernel": "msm", "rows": 1 << lr, "cols": 1 << lc, "mix": mix, "ms": ms, "algorithmic_GB/s": b / 1e9 / (ms / 1e3),
                              "frac_of_measured_hbm": b / 1e9 / (ms / 1e3) / PEAK, "scalars_per_s": (1 << (lr + lc)) / (ms / 1e3)}))


if __name__ == "__main__":
    main()
