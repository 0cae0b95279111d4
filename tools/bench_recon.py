"""Timing of pb_reconstruct (V v^T, pod_sig epilogue) and pb_project at C2 sizes.  One JSON line."""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pod_uqnn_b200 import _lib

ctx = _lib.Context(0); L = _lib.lib()
n_h, n = 160000, 1000
rng = np.random.default_rng(0)
Ud = ctx.to_device(rng.standard_normal((n_h, n)))
out = {}
for nl in (14, 32, 64):
    Vd = ctx.to_device(rng.standard_normal((n_h, nl))); vd = ctx.to_device(rng.standard_normal((n, nl)))
    Up, sig = ctx.alloc(n_h, n), ctx.alloc(n_h)
    r = {}
    for name, up in (("pod_sig_only", None), ("recon_and_sig", Up), ("recon_only", Up)):
        best = 1e9
        for _ in range(4):
            ctx.check(L.pb_reconstruct(ctx.h, Vd.ptr, Vd.ld, nl, vd.ptr, n, n_h, Ud.ptr if name != "recon_only" else None,
                                       Ud.ld, up.ptr if up else None, n if up else 0, sig.ptr if name != "recon_only" else None))
            best = min(best, ctx.ms(_lib.T_RECON))
        bytes_ = 8 * n_h * n * {"pod_sig_only": 1, "recon_and_sig": 2, "recon_only": 1}[name]
        r[name] = {"ms": best, "GBps": bytes_ / best / 1e6}
    vv = ctx.alloc(n, nl)
    best = 1e9
    for _ in range(4):
        ctx.check(L.pb_project(ctx.h, Vd.ptr, Vd.ld, nl, Ud.ptr, Ud.ld, n_h, n, vv.ptr))
        best = min(best, ctx.ms(_lib.T_PROJECT))
    r["project"] = {"ms": best, "GBps": 8 * n_h * n / best / 1e6}
    out[f"n_L={nl}"] = r
print(json.dumps(out))
