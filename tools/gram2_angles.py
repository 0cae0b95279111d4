"""Diagnostic: subspace angle of the gram2 mode (with its automatic refinement step) on the golden cases and on
DCT known-answer matrices; prints the a-priori estimate next to the measured angle.  GPU box only."""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import compare as cmp, pod_oracle as po
from pod_uqnn_b200 import _lib, workloads as wl
from pod_uqnn_b200.pod import perform_pod, pod_device

ctx = _lib.Context(0); L = _lib.lib()
out = []
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name, field in [("shekel_c1", "shekel"), ("shekel_tall", "shekel"), ("ackley_64", "ackley"), ("burgers_1step", "burgers")]:
    g = np.load(os.path.join(G, f"pod_{name}.npz"))
    n_t = int(g["n_t"]) if "n_t" in g.files else 0
    _, U, _, _ = po.create_snapshots(g["x_mesh"], 1, n_t, po.FIELDS[field], g["mu"], float(g["t_min"]) if "t_min" in g.files else 0,
                                     float(g["t_max"]) if "t_max" in g.files else 0)
    for mode in ("gram2", "tsqr"):
        V, sig = perform_pod(U, float(g["eps"]), 0, False, mode=mode, ctx=ctx, return_sigma=True)
        st = C.c_int(); L.pb_pod_refine_steps(ctx.h, C.byref(st)); es = C.c_double(); L.pb_pod_refine_estimate(ctx.h, C.byref(es))
        ok = V.shape == g["V"].shape
        out.append({"case": name, "shape": list(U.shape), "mode": mode, "n_L": V.shape[1], "rank_ok": bool(ok), "refine": st.value, "estimate": es.value,
                    "refine_ms": ctx.ms(_lib.T_REFINE),
                    "angle": float(cmp.largest_principal_angle(V, g["V"])) if ok else None})
        print(out[-1], flush=True)
for n_h, n_s, decay in [(65536, 1024, 0.25), (65536, 1024, 0.5), (65536, 1024, 1.0), (200000, 2048, 0.25)]:
    U = ctx.alloc(n_h, n_s)
    ctx.check(L.pb_dct_lowrank(ctx.h, n_h, n_s, 256, decay, 0, n_h, U.ptr, U.ld))
    s = 2. ** (-decay * np.arange(256))
    for mode in ("gram2", "tsqr"):
        V, sig = pod_device(ctx, U, 1e-10, 0, mode)
        st = C.c_int(); L.pb_pod_refine_steps(ctx.h, C.byref(st)); es = C.c_double(); L.pb_pod_refine_estimate(ctx.h, C.byref(es))
        nl = V.shape[1]
        Phi, Cd, Gd = ctx.alloc(n_h, nl), ctx.alloc(nl, nl), ctx.alloc(nl, nl)
        ctx.check(L.pb_dct_basis(ctx.h, n_h, nl, 0, n_h, Phi.ptr, Phi.ld))
        ctx.check(L.pb_gemm_tn(ctx.h, Phi.ptr, Phi.ld, nl, V.ptr, V.ld, nl, n_h, Cd.ptr))
        ctx.check(L.pb_residual_gram(ctx.h, V.ptr, V.ld, nl, Phi.ptr, Phi.ld, nl, n_h, Cd.ptr, Gd.ptr))
        ang = float(np.sqrt(max(0., np.linalg.eigvalsh(Gd.download())[-1])))
        out.append({"case": f"dct {n_h}x{n_s} decay {decay}", "mode": mode, "n_L": nl, "n_L_closed": wl.dct_rank(s, 1e-10), "refine": st.value, "estimate": es.value,
                    "refine_ms": ctx.ms(_lib.T_REFINE), "angle": ang, "sigma_err": float(np.abs(sig[:nl] - s[:nl]).max())})
        print(out[-1], flush=True)
    U.release()
# C2 at full size: gram2 (as the library decides) against the TSQR result of the same matrix
from pod_uqnn_b200.acceleration import snapshots_device
w = wl.ackley(400, 400, 1000)
Ud, _, _ = snapshots_device(ctx, "ackley", w["x_mesh"][:, 1:].T, w["mu"])
Vt, _ = pod_device(ctx, Ud, 1e-10, 0, "tsqr")
Vg, sg = pod_device(ctx, Ud, 1e-10, 0, "gram2")
st = C.c_int(); L.pb_pod_refine_steps(ctx.h, C.byref(st)); es = C.c_double(); L.pb_pod_refine_estimate(ctx.h, C.byref(es))
nl = Vg.shape[1]
Cd, Gd = ctx.alloc(nl, nl), ctx.alloc(nl, nl)
ctx.check(L.pb_gemm_tn(ctx.h, Vt.ptr, Vt.ld, nl, Vg.ptr, Vg.ld, nl, 160000, Cd.ptr))
ctx.check(L.pb_residual_gram(ctx.h, Vg.ptr, Vg.ld, nl, Vt.ptr, Vt.ld, nl, 160000, Cd.ptr, Gd.ptr))
out.append({"case": "C2 full gram2 vs tsqr", "n_L": nl, "n_L_tsqr": Vt.shape[1], "refine": st.value, "estimate": es.value,
            "angle": float(np.sqrt(max(0., np.linalg.eigvalsh(Gd.download())[-1])))})
print(out[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gram2_angles.json", "w"), indent=1)
