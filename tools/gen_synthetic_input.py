#!/usr/bin/env python3
"""Synthetic image + weights in the reference's input format (whitespace-separated decimals, one stream:
image, then per conv layer weights[co][ci][m][m] + bias[co], then per FC layer weights[co][ci] + bias[co];
order fixed by src/neuralNetwork.cpp:805-897 of the reference).

The GPU box has no network and the 124 MB trained vgg11 file of the reference does not travel, so benchmarks and
full-size parity runs use seeded synthetic weights of the same architecture: weights ~ N(0, 2/fan_in) clipped at
6 sigma (the trained vgg11 weights have max/sigma of 6-10), biases ~ N(0, 0.05), image ~ U[0, 1).

usage: gen_synthetic_input.py vgg11|vgg16|lenet OUT.csv [--seed S] [--config "64 M 128 M ..."]
Also writes the network description next to it for the `vgg` model class (OUT.csv.config).
"""
import argparse
import numpy as np

CONFIGS = {
    "vgg11": "64 M 128 M 256 256 M 512 512 M 512 512 M",
    "vgg16": "64 64 M 128 128 M 256 256 256 M 512 512 512 M 512 512 512 M",
}


def vgg_shapes(config, pic=32, ch=3):
    conv, fc = [], []
    n = pic
    for tok in config.split():
        if tok in ("M", "A"):
            n = ((n - 2) >> 1) + 1
        else:
            conv.append((int(tok), ch, 3))
            ch = int(tok)
    fc = [(512, n * n * ch), (512, 512), (10, 512)]
    return (3, pic, pic), conv, fc


def lenet_shapes():
    return (1, 32, 32), [(6, 1, 5), (16, 6, 5)], [(120, 400), (84, 120), (10, 84)]


def generate(model, seed=20211115, config=None):
    if model == "lenet":
        img, conv, fc = lenet_shapes()
    else:
        img, conv, fc = vgg_shapes(config or CONFIGS[model])
    rng = np.random.default_rng(seed)
    parts = [rng.random(int(np.prod(img)), dtype=np.float32)]
    for co, ci, m in conv:
        s = np.sqrt(2.0 / (ci * m * m))
        parts.append(np.clip(rng.normal(0, s, co * ci * m * m), -6 * s, 6 * s).astype(np.float32))
        parts.append(rng.normal(0, 0.05, co).astype(np.float32))
    for co, ci in fc:
        s = np.sqrt(2.0 / ci)
        parts.append(np.clip(rng.normal(0, s, co * ci), -6 * s, 6 * s).astype(np.float32))
        parts.append(rng.normal(0, 0.05, co).astype(np.float32))
    return np.concatenate(parts)


def write_text(values, path):
    # %.9g round-trips float32, so every reader (`istream >> double`, strtod) sees the same doubles
    with open(path, "w") as f:
        step = 1 << 16
        for i in range(0, len(values), step):
            f.write(" ".join(map("%.9g".__mod__, values[i:i + step].tolist())))
            f.write("\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model", choices=["vgg11", "vgg16", "lenet"])
    ap.add_argument("out")
    ap.add_argument("--seed", type=int, default=20211115)
    ap.add_argument("--config", default=None)
    a = ap.parse_args()
    v = generate(a.model, a.seed, a.config)
    write_text(v, a.out)
    if a.model != "lenet":
        with open(a.out + ".config", "w") as f:
            f.write((a.config or CONFIGS[a.model]) + "\n")
    print(f"{a.out}: {len(v)} values")


if __name__ == "__main__":
    main()
