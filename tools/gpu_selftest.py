"""GPU self-test / diagnostics (dev aid): runs every kernel family against numpy / the oracle and
prints numbers instead of asserting, so that one gpurun call localises as many bugs as possible.
Writes gpurun_out/selftest.json."""
import json, os, sys, time, traceback
import ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pod_uqnn_b200 import _lib, workloads as wl
from pod_uqnn_b200.pod import pod_device, fast_pod_device
from pod_uqnn_b200.acceleration import snapshots_device
from pod_uqnn_b200.varneuralnetwork import ensemble_predict_device
from oracle import pod_oracle as po, compare as cmp, ensemble as ens

OUT = {}
ctx = _lib.Context(0)
L = _lib.lib()
print("device", ctx.device_info(), flush=True)
only = set(sys.argv[1:])


def section(name):
    def deco(fn):
        if only and name not in only:
            return fn
        t = time.time()
        try:
            OUT[name] = fn()
        except Exception as e:  # noqa: BLE001
            OUT[name] = {"error": repr(e), "trace": traceback.format_exc()[-1500:]}
        OUT[name + "_s"] = round(time.time() - t, 2)
        print(name, json.dumps(OUT[name], default=float)[:1500], flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(OUT, open("gpurun_out/selftest.json", "w"), indent=1, default=float)
        return fn
    return deco


@section("probes")
def _():
    r = {}
    for kind, name in ((0, "dfma"), (1, "dmma"), (2, "mixed")):
        v = C.c_double(); ctx.check(L.pb_probe_fp64_peak(ctx.h, kind, C.byref(v))); r[name + "_tflops"] = v.value
    v = C.c_double(); ctx.check(L.pb_probe_hbm_copy(ctx.h, 1 << 30, C.byref(v))); r["hbm_copy_gbs"] = v.value
    return r


def gemm_tn(A, B, sym=False):
    Ad = ctx.to_device(A); Bd = Ad if sym else ctx.to_device(B)
    Cd = ctx.alloc(A.shape[1], B.shape[1])
    if sym:
        ctx.check(L.pb_gram(ctx.h, Ad.ptr, A.shape[0], A.shape[1], Ad.ld, Cd.ptr))
    else:
        ctx.check(L.pb_gemm_tn(ctx.h, Ad.ptr, Ad.ld, A.shape[1], Bd.ptr, Bd.ld, B.shape[1], A.shape[0], Cd.ptr))
    return Cd.download()


@section("gemm_tn")
def _():
    rng = np.random.default_rng(0); r = {}
    for (rows, ka, kb, sym) in ((1000, 200, 200, True), (5000, 1000, 1000, True), (777, 71, 71, True), (3001, 14, 300, False),
                                (4096, 71, 333, False), (100000, 256, 256, True), (40, 129, 129, True), (9000, 33, 1001, False)):
        A = rng.standard_normal((rows, ka)); B = A if sym else rng.standard_normal((rows, kb))
        Cg = gemm_tn(A, B, sym); Cr = A.T @ B
        r[f"{rows}x{ka}x{kb}{'s' if sym else ''}"] = float(np.abs(Cg - Cr).max() / np.abs(Cr).max())
        if sym: r[f"{rows}x{ka}_asym"] = float(np.abs(Cg - Cg.T).max())
    return r


@section("gemm_nn")
def _():
    rng = np.random.default_rng(1); r = {}
    for (rows, m, k) in ((1000, 200, 14), (5000, 1000, 71), (777, 333, 333), (130, 14, 1000), (4097, 1000, 1000), (300, 71, 2), (999, 15, 33)):
        A = rng.standard_normal((rows, m)); B = rng.standard_normal((m, k))
        Ad, Bd, Yd = ctx.to_device(A), ctx.to_device(B), ctx.alloc(rows, k)
        ctx.check(L.pb_gemm_nn(ctx.h, Ad.ptr, Ad.ld, rows, m, Bd.ptr, Bd.ld, k, Yd.ptr, Yd.ld))
        Yr = A @ B
        r[f"{rows}x{m}x{k}"] = float(np.abs(Yd.download() - Yr).max() / np.abs(Yr).max())
    return r


def pod_report(U, eps, mode, Vr=None, D=None):
    Ud = ctx.to_device(U)
    t = time.time(); Vd, sig = pod_device(ctx, Ud, eps, 0, mode); dt = time.time() - t
    V = Vd.download()
    if Vr is None:
        Vr = po.perform_pod(U, eps); D, _ = po.pod_spectrum(U)
    r = {"n_L": V.shape[1], "n_L_ref": Vr.shape[1], "sig_err_all": cmp.sigma_error(sig, D),
         "sig_err_kept": cmp.sigma_error(sig[:V.shape[1]], D[:V.shape[1]]),
         "wall_s": round(dt, 4), "ms_gram": ctx.ms(_lib.T_GRAM), "ms_eig": ctx.ms(_lib.T_EIG), "ms_pass2": ctx.ms(_lib.T_PASS2),
         "ms_total": ctx.ms(_lib.T_POD_TOTAL), "ms_basis": ctx.ms(_lib.T_BASIS)}
    if V.shape[1] == Vr.shape[1]:
        r["angle"] = cmp.largest_principal_angle(V, Vr)
        r["orth"] = float(np.abs(V.T @ V - np.eye(V.shape[1])).max())
    return r


@section("pod_random_lowrank")
def _():
    rng = np.random.default_rng(2); r = {}
    for (n, m, k) in ((500, 64, 10), (2000, 200, 40), (300, 500, 30)):
        s = 2.0 ** (-np.arange(k) * 0.7)
        U = (np.linalg.qr(rng.standard_normal((n, k)))[0] * s) @ np.linalg.qr(rng.standard_normal((m, k)))[0].T
        for mode in ("gram", "gram2", "accurate"):
            r[f"{n}x{m}r{k}_{mode}"] = pod_report(U, 1e-8, mode)
    return r


@section("fields")
def _():
    r = {}
    for name, w, u in (("shekel", wl.shekel(100, 48), po.u_shekel), ("ackley", wl.ackley(20, 16, 40), po.u_ackley),
                       ("burgers", wl.burgers(64, 20, 10), po.u_burgers)):
        X = w["x_mesh"][:, 1:].T
        Xv_r, U_r, Us_r, _ = po.create_snapshots(w["x_mesh"], 1, w["n_t"], u, w["mu"], w["t_min"], w["t_max"])
        Ud, Xv, t = snapshots_device(ctx, name, X, w["mu"], w["n_t"], w["t_min"], w["t_max"])
        Ug = Ud.download()
        r[name] = {"rel_err": cmp.rel_err(Ug, U_r), "xv_err": float(np.abs(Xv - Xv_r).max()), "ms": ctx.ms(_lib.T_SNAPSHOTS),
                   "nan": int(np.isnan(Ug).sum())}
    return r


@section("pod_cases")
def _():
    r = {}
    cases = []
    w = wl.shekel(400, 300); cases.append(("shekel400x300", po.create_snapshots(w["x_mesh"], 1, 0, po.u_shekel, w["mu"])[1], 1e-10))
    w = wl.shekel(200, 300); cases.append(("shekel200x300w", po.create_snapshots(w["x_mesh"], 1, 0, po.u_shekel, w["mu"])[1], 1e-10))
    w = wl.ackley(48, 48, 160); cases.append(("ackley2304x160", po.create_snapshots(w["x_mesh"], 1, 0, po.u_ackley, w["mu"])[1], 1e-10))
    w = wl.burgers(256, 100, 3); cases.append(("burgers256x300w", po.create_snapshots(w["x_mesh"], 1, 100, po.u_burgers, w["mu"], 1., 5.)[1], 1e-8))
    for name, U, eps in cases:
        Vr = po.perform_pod(U, eps); D, _ = po.pod_spectrum(U)
        for mode in ("gram", "gram2", "accurate"):
            r[f"{name}_{mode}"] = pod_report(U, eps, mode, Vr, D)
    return r


@section("fast_pod")
def _():
    r = {}
    w = wl.burgers(256, 100, 24)
    _, U, Us, _ = po.create_snapshots(w["x_mesh"], 1, 100, po.u_burgers, w["mu"], 1., 5.)
    for eps in (1e-4, 1e-8):
        Vr, ranks_r = po.perform_fast_pod(Us, eps, eps, return_ranks=True)
        for mode in ("gram", "accurate"):
            Ud = ctx.to_device(U)
            t = time.time(); Vd, ranks = fast_pod_device(ctx, Ud, 100, 24, eps, eps, mode); dt = time.time() - t
            V = Vd.download()
            rr = {"n_L": V.shape[1], "n_L_ref": Vr.shape[1], "rank_mismatch": int((ranks != ranks_r).sum()), "wall_s": round(dt, 3)}
            if V.shape[1] == Vr.shape[1]: rr["angle"] = cmp.largest_principal_angle(V, Vr)
            r[f"eps{eps}_{mode}"] = rr
    return r


@section("project_reconstruct")
def _():
    rng = np.random.default_rng(3); r = {}
    for (n_h, n, nl) in ((2304, 160, 14), (1000, 333, 71)):
        U = rng.standard_normal((n_h, n)); V = np.linalg.qr(rng.standard_normal((n_h, nl)))[0]
        Ud, Vd = ctx.to_device(U), ctx.to_device(V)
        v = ctx.alloc(n, nl)
        ctx.check(L.pb_project(ctx.h, Vd.ptr, Vd.ld, nl, Ud.ptr, Ud.ld, n_h, n, v.ptr))
        v_r = po.project_to_v(V, U)
        Up, sig = ctx.alloc(n_h, n), ctx.alloc(n_h)
        ctx.check(L.pb_reconstruct(ctx.h, Vd.ptr, Vd.ld, nl, v.ptr, n, n_h, Ud.ptr, Ud.ld, Up.ptr, Up.ld, sig.ptr))
        Up_r = po.project_to_U(V, v_r)
        r[f"{n_h}x{n}x{nl}"] = {"v": cmp.rel_err(v.download(), v_r), "U_pod": cmp.rel_err(Up.download(), Up_r),
                                 "pod_sig": cmp.rel_err(sig.download(), po.pod_sig(U, Up_r))}
    return r


@section("ensemble")
def _():
    r = {}
    for (layers, M, N, norm) in (([1, 128, 128, 128, 64], 8, 1000, "meanstd"), ([3, 40, 40, 14], 5, 300, "center"),
                                 ([11, 128, 128, 128, 71], 3, 257, "none"), ([2, 16, 5], 1, 100, "meanstd")):
        members = [ens.init_weights(layers, 4000 + i, bias_scale=0.1) for i in range(M)]
        rng = np.random.default_rng(5)
        X = rng.uniform(800, 1200, (N, layers[0]))
        nb = ens.set_normalize_bounds(X, norm)
        vm_r, vs_r = ens.predict_v(X, members, layers[-1], 0.01, norm, nb)
        vm, vs = ensemble_predict_device(ctx, members, layers, X, 0.01, norm, nb)
        r[f"{layers}_M{M}"] = {"mean": cmp.rel_err(vm.download(), vm_r), "sig": cmp.rel_err(vs.download(), vs_r),
                               "ms": ctx.ms(_lib.T_PREDICT)}
    return r


@section("c2_timing")
def _():
    w = wl.ackley(400, 400, 1000)
    X = w["x_mesh"][:, 1:].T
    Ud, _, _ = snapshots_device(ctx, "ackley", X, w["mu"])
    r = {"snap_ms": ctx.ms(_lib.T_SNAPSHOTS)}
    n_h, n_s = Ud.shape
    for mode in ("gram", "gram2"):
        for rep in range(2):
            t = time.time(); Vd, sig = pod_device(ctx, Ud, 1e-10, 0, mode); dt = time.time() - t
        nl = Vd.shape[1]
        r[mode] = {"n_L": nl, "wall_s": round(dt, 4), "ms_gram": ctx.ms(_lib.T_GRAM), "ms_eig": ctx.ms(_lib.T_EIG),
                   "ms_pass2": ctx.ms(_lib.T_PASS2), "ms_total": ctx.ms(_lib.T_POD_TOTAL), "ms_basis": ctx.ms(_lib.T_BASIS),
                   "gram_tflops": n_h * n_s * (n_s + 1) / (ctx.ms(_lib.T_GRAM) * 1e-3) / 1e12,
                   "sig_head": [float(s) for s in sig[:16]]}
        v = ctx.alloc(n_s, nl)
        ctx.check(L.pb_project(ctx.h, Vd.ptr, Vd.ld, nl, Ud.ptr, Ud.ld, n_h, n_s, v.ptr))
        r[mode]["ms_project"] = ctx.ms(_lib.T_PROJECT)
        sigd = ctx.alloc(n_h)
        ctx.check(L.pb_reconstruct(ctx.h, Vd.ptr, Vd.ld, nl, v.ptr, n_s, n_h, Ud.ptr, Ud.ld, None, 0, sigd.ptr))
        r[mode]["ms_podsig"] = ctx.ms(_lib.T_RECON)
        r[mode]["mean_pod_sig"] = float(sigd.download().mean())
    return r


print("launches", ctx.launches())
json.dump(OUT, open("gpurun_out/selftest.json", "w"), indent=1, default=float)
