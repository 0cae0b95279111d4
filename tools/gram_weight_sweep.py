"""Sweep of the stream-K schedule weights of the special Gram tile kinds (PB_GRAM_WEIGHTS) on the C2 matrix."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pod_uqnn_b200 import _lib, workloads as wl
from pod_uqnn_b200.acceleration import snapshots_device

L = _lib.lib()
ctx0 = _lib.Context(0)
w = wl.ackley(400, 400, 1000)
Ud, _, _ = snapshots_device(ctx0, w["field"], w["x_mesh"][:, 1:].T, w["mu"])
ref = None
for wts in sys.argv[1:] or ["52,40,37"]:
    os.environ["PB_GRAM_WEIGHTS"] = wts
    ctx = _lib.Context(0)
    G = ctx.alloc(1000, 1000)
    ts = []
    for _ in range(6):
        ctx.check(L.pb_gram(ctx.h, Ud.ptr, Ud.shape[0], 1000, Ud.ld, G.ptr))
        ctx.sync()
        ts.append(ctx.ms(_lib.T_TN_KERNEL))
    g = G.download()
    if ref is None:
        ref = g
    print(wts, "kernel ms min %.4f med %.4f" % (min(ts[1:]), float(np.median(ts[1:]))), "max|dG|/max|G| %.2e" % (np.abs(g - ref).max() / np.abs(ref).max()), flush=True)

# per-CTA busy time against the iterations of each tile kind: least-squares cost per iteration and kind
import ctypes
for wts in (sys.argv[1:] or ["52,40,37"])[:2]:
    os.environ["PB_GRAM_WEIGHTS"] = wts
    ctx = _lib.Context(0)
    G = ctx.alloc(1000, 1000)
    ctx.check(L.pb_gram_profile(ctx.h, 1, 0, None, None, None))
    for _ in range(3):
        ctx.check(L.pb_gram(ctx.h, Ud.ptr, Ud.shape[0], 1000, Ud.ld, G.ptr))
    ctx.sync()
    busy = (ctypes.c_double * 1024)(); it = (ctypes.c_int * 4096)(); n = ctypes.c_int()
    ctx.check(L.pb_gram_profile(ctx.h, 1, 1024, busy, it, ctypes.byref(n)))
    P = n.value
    b = np.array(busy[:P]); A = np.array(it[:4 * P], float).reshape(P, 4)
    sol, *_ = np.linalg.lstsq(np.c_[A, np.ones(P)], b, rcond=None)
    print(wts, "busy us min/mean/max", b.min(), b.mean(), b.max(), "cost/iter by kind (ns)", (sol[:4] * 1e3).round(2), "const us", sol[4].round(1),
          "rel to full x64:", (sol[:4] / sol[0] * 64).round(1), flush=True)
    order = np.argsort(b)
    for c in list(order[:4]) + list(order[-4:]):
        print("   cta", c, "busy", b[c].round(1), "iters", A[c].astype(int))
