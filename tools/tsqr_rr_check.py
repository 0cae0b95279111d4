"""Development aid (GPU): the rank-revealing TSQR mode against the full one -- rows of the factor, panels run, time."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pod_uqnn_b200 import _lib, workloads as wl           # noqa: E402
from pod_uqnn_b200.acceleration import snapshots_device   # noqa: E402
from pod_uqnn_b200.pod import pod_device                  # noqa: E402
from oracle import compare as cmp                          # noqa: E402  (development check only)

ctx = _lib.Context(0)
L = _lib.lib()
out = []


def stats():
    p, f, fl = C.c_int64(), C.c_int64(), C.c_double()
    ctx.check(L.pb_tsqr_stats(ctx.h, C.byref(p), C.byref(f), C.byref(fl)))
    return p.value, f.value, fl.value


def run(name, Ud, eps=1e-10, modes=("tsqr", "tsqr_full", "gram2")):
    res = {"case": name, "shape": list(Ud.shape)}
    ref = None
    for mode in modes:
        pod_device(ctx, Ud, eps, 0, mode); ctx.sync()
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter(); V, sig = pod_device(ctx, Ud, eps, 0, mode); ctx.sync()
            best = min(best, time.perf_counter() - t0)
        r = {"ms": best * 1e3, "n_L": V.shape[1], "qr_ms": ctx.ms(_lib.T_QR) if mode.startswith("tsqr") else 0.,
             "eig_ms": ctx.ms(_lib.T_EIG)}
        if mode.startswith("tsqr"):
            r["panels"], r["f_rows"], r["flops"] = stats()
        Vh = V.download() if V.shape[0] * V.shape[1] < 4e7 else None
        if mode == "tsqr_full":
            ref = (Vh, sig)
        res[mode] = r
        res[mode]["_V"] = Vh; res[mode]["_sig"] = sig
    if ref is not None and ref[0] is not None:
        for mode in modes:
            if mode == "tsqr_full" or res[mode]["_V"] is None or res[mode]["n_L"] != res["tsqr_full"]["n_L"]:
                continue
            res[mode]["angle_vs_full"] = cmp.largest_principal_angle(res[mode]["_V"], ref[0])
            res[mode]["sigma_err_vs_full"] = cmp.sigma_error(res[mode]["_sig"], ref[1])
    for mode in modes:
        res[mode].pop("_V"); res[mode].pop("_sig")
    print(json.dumps(res), flush=True)
    out.append(res)


w = wl.ackley(400, 400, 1000)
Ud, _, _ = snapshots_device(ctx, "ackley", w["x_mesh"][:, 1:].T, w["mu"])
run("C2 ackley 160000x1000", Ud)
del Ud
w = wl.shekel(400, 1000) if hasattr(wl, "shekel") else None
if w is not None:
    Ud, _, _ = snapshots_device(ctx, "shekel", w["x_mesh"][:, 1:].T, w["mu"])
    run("C1 shekel 400x1000", Ud)
    del Ud
rng = np.random.default_rng(3)
run("gaussian 50000x1024 (full rank)", ctx.to_device(rng.standard_normal((50000, 1024))), modes=("tsqr", "tsqr_full"))
for n_h, n_s in ((250000, 4096), (1250000, 4096)):
    Ud = ctx.alloc(n_h, n_s)
    ctx.check(L.pb_dct_lowrank(ctx.h, n_h, n_s, 256, 0.25, 0, n_h, Ud.ptr, Ud.ld))
    run(f"DCT {n_h}x{n_s}", Ud, modes=("tsqr", "gram2") if n_h > 300000 else ("tsqr", "tsqr_full", "gram2"))
    Ud.release()
json.dump(out, open("gpurun_out/r02_tsqr_rr_check.json", "w"), indent=1)
