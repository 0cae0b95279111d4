"""Snapshot kernels on one B200: Ackley C2 through its three kernels (plain / coordinate-indexed / radius-grouped),
Shekel C1, Burgers C3.  Device-event times (PB_T_SNAPSHOTS), best of 5 after a first call; HBM write rate = 8 n_h n_st / t.
Writes gpurun_out/bench_snapshots.json.  `python tools/bench_snapshots.py once` runs the grouped kernel twice (for ncu)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pod_uqnn_b200 import _lib, workloads as wl
from pod_uqnn_b200.acceleration import snapshots_device

ctx = _lib.Context(0)
OUT = {}


def run(name, field, w, n_t=0, t_min=0., t_max=0., reps=5, **kw):
    X = w["x_mesh"][:, 1:].T
    U = None
    best = 1e9
    for i in range(reps + 1):
        U = snapshots_device(ctx, field, X, w["mu"], n_t, t_min, t_max, out=U, **kw)[0]
        if i:
            best = min(best, ctx.ms(_lib.T_SNAPSHOTS))
    nbytes = 8 * U.shape[0] * U.shape[1]
    OUT[name] = {"ms": best, "gbs": nbytes / (best * 1e-3) / 1e9, "bytes": nbytes}
    print(name, OUT[name], flush=True)
    h = U.download() if nbytes < 2e8 else None
    U.release()
    return h


if "once" in sys.argv:
    run("ackley_c2_grouped", "ackley", wl.ackley(400, 400, 1000), reps=1, ackley_path="grouped")
    sys.exit(0)

w = wl.ackley(400, 400, 1000)
for path in ("plain", "indexed", "grouped"):
    run(f"ackley_c2_{path}", "ackley", w, ackley_path=path)
run("shekel_c1_400x1000", "shekel", wl.shekel(400, 1000))
run("burgers_c3_256x100x300", "burgers", wl.burgers(256, 100, 300), 100, 1., 5.)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(OUT, open("gpurun_out/bench_snapshots.json", "w"), indent=1)
