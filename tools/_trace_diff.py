import re, sys
def load(path):
    d = {}
    for ln in open(path):
        m = re.match(r"\[zk_trace\] (\S+)\s+calls\s+(\d+) total\s+([\d.]+) ms", ln)
        if m: d[m.group(1)] = (int(m.group(2)), float(m.group(3)))
    return d
a, b, k = load(sys.argv[1]), load(sys.argv[2]), int(sys.argv[3])
rows = [(n, (b[n][0] - a.get(n, (0, 0))[0]) / k, (b[n][1] - a.get(n, (0, 0))[1]) / k) for n in b]
rows.sort(key=lambda r: -r[2])
tot = 0
for n, c, ms in rows:
    if ms > 0.005: print(f"{n:36s} calls/proof {c:7.1f}  ms/proof {ms:8.3f}")
    tot += ms
print("sum of API time per proof", round(tot, 3), "ms")
