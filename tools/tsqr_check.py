"""GPU diagnostic of the TSQR mode, stage by stage (development aid): R against LAPACK, the pivoted factor and
Jacobi against the SVD of the same matrix, timings at C2 size."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import pod_oracle as po, compare as cmp      # noqa: E402
from pod_uqnn_b200 import _lib, workloads as wl          # noqa: E402
from pod_uqnn_b200.pod import pod_device                 # noqa: E402
from pod_uqnn_b200.acceleration import snapshots_device  # noqa: E402

ctx = _lib.Context(0)
L = _lib.lib()
out = {}


def tsqr_r(A):
    Ad = ctx.to_device(A)
    m = A.shape[1]
    Rd = ctx.alloc(m, m)
    ctx.check(L.pb_tsqr_r(ctx.h, Ad.ptr, A.shape[0], m, Ad.ld, Rd.ptr))
    return Rd.download()


def check_r(name, A):
    R = tsqr_r(A)
    G = A.T @ A
    e1 = np.abs(R.T @ R - G).max() / np.abs(G).max()
    low = np.abs(np.tril(R, -1)).max()
    Rl = np.linalg.qr(A, mode="r")
    k = min(Rl.shape[0], R.shape[0])
    s = np.sign(np.diag(R)[:k]) * np.sign(np.diag(Rl)[:k]); s[s == 0] = 1
    e2 = np.abs(R[:k] * s[:, None] - Rl[:k]).max() / np.abs(Rl).max()
    print(f"[R] {name} {A.shape}: |RtR-AtA|/|AtA| = {e1:.2e}, below-diag max {low:.1e}, vs LAPACK R {e2:.2e}, finite {np.isfinite(R).all()}", flush=True)
    out[f"r_{name}"] = dict(gram_err=float(e1), lapack_err=float(e2))
    return R


rng = np.random.default_rng(0)
for shape in ((700, 130), (300, 64), (64, 64), (100, 1), (33, 70), (5000, 333), (2000, 1000)):
    check_r("random", rng.standard_normal(shape))


def check_pod(name, U, eps=1e-10):
    Ud = ctx.to_device(U)
    t0 = time.perf_counter()
    Vd, sig = pod_device(ctx, Ud, eps, 0, "tsqr")
    ctx.sync(); dt = time.perf_counter() - t0
    V = Vd.download()
    Vr = po.perform_pod(U, eps)
    D = np.linalg.svd(U, compute_uv=False)
    ang = cmp.largest_principal_angle(V, Vr) if V.shape == Vr.shape else float("nan")
    se = cmp.sigma_error(sig, D)
    print(f"[POD] {name} {U.shape}: n_L {V.shape[1]} (ref {Vr.shape[1]}), sigma err {se:.2e}, angle {ang:.2e}, "
          f"qr {ctx.ms(_lib.T_QR):.2f} ms, eig(qrcp+jacobi) {ctx.ms(_lib.T_EIG):.2f} ms, total {ctx.ms(_lib.T_POD_TOTAL):.2f} ms, wall {dt*1e3:.1f} ms", flush=True)
    out[f"pod_{name}"] = dict(n_L=int(V.shape[1]), n_L_ref=int(Vr.shape[1]), sigma_err=float(se), angle=float(ang),
                              qr_ms=ctx.ms(_lib.T_QR), eig_ms=ctx.ms(_lib.T_EIG), total_ms=ctx.ms(_lib.T_POD_TOTAL))


w = wl.ackley(48, 48, 160)
check_pod("ackley_small", po.create_snapshots(w["x_mesh"], 1, 0, po.u_ackley, w["mu"])[1])
w = wl.shekel(400, 1000)
U = po.create_snapshots(w["x_mesh"], 1, 0, po.u_shekel, w["mu"])[1]
check_pod("shekel_c1_wide", U)
check_pod("shekel_tall", np.ascontiguousarray(U[:, :300]))
A, s, phi = wl.dct_lowrank_host(3000, 192, r=64)
check_pod("dct", A)
check_pod("random_full", rng.standard_normal((900, 200)), eps=0.)

if "--big" in sys.argv:
    w = wl.ackley(400, 400, 1000)
    Ud, _, _ = snapshots_device(ctx, "ackley", w["x_mesh"][:, 1:].T, w["mu"])
    for mode in ("tsqr", "tsqr", "accurate", "gram2"):
        Vd, sig = pod_device(ctx, Ud, 1e-10, 0, mode)
        print(f"[C2] {mode}: n_L {Vd.shape[1]}, total {ctx.ms(_lib.T_POD_TOTAL):.2f} ms, qr {ctx.ms(_lib.T_QR):.2f}, eig {ctx.ms(_lib.T_EIG):.2f}, "
              f"gram {ctx.ms(_lib.T_GRAM):.2f}, pass2 {ctx.ms(_lib.T_PASS2):.2f}", flush=True)
        out[f"c2_{mode}"] = dict(n_L=int(Vd.shape[1]), total_ms=ctx.ms(_lib.T_POD_TOTAL), qr_ms=ctx.ms(_lib.T_QR), eig_ms=ctx.ms(_lib.T_EIG))
        if mode == "tsqr":
            g = np.load("tests/golden/pod_ackley_c2_full.npz")
            n_L = int(g["n_L"])
            print(f"     sigma err vs reference {cmp.sigma_error(sig[:n_L], g['sigma'][:n_L]):.2e}", flush=True)
            V = Vd.download()
            Vrows, sg = cmp.align_signs(V[g["rows"]], g["V_rows"])
            print(f"     V rows err {np.abs(Vrows - g['V_rows']).max() / np.abs(g['V_rows']).max():.2e}", flush=True)
json.dump(out, open("gpurun_out/tsqr_check.json", "w"), indent=1)
