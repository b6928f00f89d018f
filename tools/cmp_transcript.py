#!/usr/bin/env python3
"""Compare a recorded transcript with a golden one.  usage: cmp_transcript.py golden.bin mine.bin [--gkr-only N_COMM]

--gkr-only N: `mine` holds only the GKR messages (CPU-Hyrax drop-in variant); compare against the golden slice that
follows the N commitment points (96 bytes each).
"""
import sys


def main():
    g = open(sys.argv[1], "rb").read()
    m = open(sys.argv[2], "rb").read()
    if len(sys.argv) > 3 and sys.argv[3] == "--gkr-only":
        off = int(sys.argv[4]) * 96
        g = g[off:off + len(m)]
    if g == m:
        print(f"TRANSCRIPT MATCH ({len(m)} bytes)")
        return 0
    n = min(len(g), len(m))
    first = next((i for i in range(n) if g[i] != m[i]), n)
    print(f"TRANSCRIPT MISMATCH: golden {len(g)} bytes, mine {len(m)} bytes, first difference at byte {first}")
    return 1


if __name__ == "__main__":
    sys.exit(main())
