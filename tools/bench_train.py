"""Timing of the ensemble training step (pb_ensemble_train) at the reference's training sizes.
Usage: python tools/bench_train.py [--epochs 2000] [--cpu]   (one JSON line per config)"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pod_uqnn_b200 import _lib                                   # noqa: E402
from pod_uqnn_b200.varneuralnetwork import EnsembleTrainer, glorot_normal   # noqa: E402

CONFIGS = [
    # name, layers, N, M, adv_eps, lr, lam, soft_0
    ("ackley_train (2d_ackley/hyperparams.py: n_s 500 x 0.8, n_M 5)", [3, 128, 128, 128, 14], 400, 5, 0., 1e-3, 1e-3, 0.01),
    ("burgers_train (1dt_burger/hyperparams.py: 50 x 100 x 0.8, n_M 5)", [2, 128, 128, 128, 37], 4000, 5, 0.01, 0.01, 1e-8, 0.01),
    ("burgers_300 (C3 sizes: 300 x 100 x 0.8, n_M 8)", [2, 128, 128, 128, 37], 24000, 8, 0.01, 0.01, 1e-8, 0.01),
]


def flops_per_epoch(layers, N, M, adv):
    dims = layers[:-1] + [2 * layers[-1]]
    mac = sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    # forward + delta chain + weight gradient, 2 flop per multiply-add, per pass
    return (2 if adv is not None else 1) * 3 * 2 * N * mac * M


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=2000)
    ap.add_argument("--cpu", action="store_true", help="also time the oracle port (CPU) for 20 epochs of one member")
    ap.add_argument("--only", type=int, default=-1)
    args = ap.parse_args()
    ctx = _lib.Context(0)
    for ci, (name, layers, N, M, adv, lr, lam, soft_0) in enumerate(CONFIGS):
        if args.only >= 0 and ci != args.only:
            continue
        rng = np.random.default_rng(ci)
        dims = layers[:-1] + [2 * layers[-1]]
        members = [[x for fi, fo in zip(dims[:-1], dims[1:]) for x in (glorot_normal(rng, fi, fo), np.zeros(fo))]
                   for _ in range(M)]
        X = rng.uniform(-1.7, 1.7, (N, layers[0]))
        v = rng.standard_normal((N, layers[-1]))
        out = {"config": name, "layers": layers, "N": N, "M": M, "adv_eps": adv, "epochs": args.epochs}
        for mode in ("graph", "eager"):
            t = EnsembleTrainer(ctx, members, layers, X, v, lr, lam, adv, soft_0, use_graph=(mode == "graph"))
            t.run(50)
            ctx.sync()
            ctx.timer_start()
            losses = t.run(args.epochs)
            ms = ctx.timer_stop()
            out[f"us_per_epoch_{mode}"] = 1e3 * ms / args.epochs
            out[f"tflops_{mode}"] = flops_per_epoch(layers, N, M, adv) / (ms * 1e-3 / args.epochs) / 1e12
            out[f"loss_first_last_{mode}"] = [float(losses[0].mean()), float(losses[-1].mean())]
        if args.cpu:
            from oracle import train as tr
            t0 = time.perf_counter()
            tr.train(X, v, [a.copy() for a in members[0]], layers[-1], 20, lr, lam, adv, soft_0)
            out["cpu_port_us_per_epoch_member"] = 1e6 * (time.perf_counter() - t0) / 20
            out["cpu_cores"] = os.cpu_count()
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
