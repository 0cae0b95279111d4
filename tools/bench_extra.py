"""Secondary measurements on one B200 (the other BASELINE.json configs): C1 Shekel, C3 Burgers two-step,
C4 ensemble predict, C5-shaped slab (n_s = 4096).  Writes gpurun_out/bench_extra.json."""
import ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pod_uqnn_b200 import _lib, workloads as wl
from pod_uqnn_b200.acceleration import snapshots_device
from pod_uqnn_b200.pod import pod_device, fast_pod_device
from pod_uqnn_b200.varneuralnetwork import ensemble_predict_device, pack_members
from oracle import ensemble as ens

ctx = _lib.Context(0); L = _lib.lib(); OUT = {}
only = set(sys.argv[1:])
pk = C.c_double(); ctx.check(L.pb_probe_fp64_peak(ctx.h, 1, C.byref(pk))); OUT["fp64_peak_tflops"] = pk.value


def timed(fn, reps=3):
    fn(); ctx.sync()
    best = 1e9
    for _ in range(reps):
        ctx.sync(); ctx.timer_start(); r = fn(); best = min(best, ctx.timer_stop())
    return best, r


def want(name): return not only or name in only


if want("c1"):
    w = wl.shekel(400, 1000)
    Ud, _, _ = snapshots_device(ctx, "shekel", w["x_mesh"][:, 1:].T, w["mu"])
    r = {"snap_ms": ctx.ms(_lib.T_SNAPSHOTS)}
    for mode in ("gram", "gram2", "accurate", "tsqr"):
        ms, (V, sig) = timed(lambda: pod_device(ctx, Ud, 1e-10, 0, mode))
        r[mode] = {"ms": ms, "n_L": V.shape[1], "eig_ms": ctx.ms(_lib.T_EIG), "pass2_ms": ctx.ms(_lib.T_PASS2),
                   "qr_ms": ctx.ms(_lib.T_QR) if mode == "tsqr" else 0.}
    OUT["c1_shekel_400x1000"] = r; print("c1", r, flush=True)

if want("c3"):
    w = wl.burgers(256, 100, 300)
    Ud, _, _ = snapshots_device(ctx, "burgers", w["x_mesh"][:, 1:].T, w["mu"], 100, 1., 5.)
    r = {"snap_ms": ctx.ms(_lib.T_SNAPSHOTS)}
    for mode, eps in (("gram", 1e-4), ("accurate", 1e-4), ("accurate", 1e-10), ("tsqr", 1e-4)):
        ms, (V, ranks) = timed(lambda: fast_pod_device(ctx, Ud, 100, 300, eps, eps, mode), reps=1)
        r[f"two_step_{mode}_{eps:g}"] = {"ms": ms, "n_L": V.shape[1], "ranks_sum": int(ranks.sum()), "ranks_max": int(ranks.max())}
    ms, (V, sig) = timed(lambda: pod_device(ctx, Ud, 1e-4, 0, "accurate"))
    r["one_step_accurate_1e-4"] = {"ms": ms, "n_L": V.shape[1]}
    OUT["c3_burgers_256x100x300"] = r; print("c3", r, flush=True)

if want("c4"):
    r = {}
    for layers, N in (([1, 128, 128, 128, 64], 1_000_000), ([1, 40, 40, 64], 1_000_000)):
        M = 8
        members = [ens.init_weights(layers, 4000 + i) for i in range(M)]
        X = np.linspace(800, 1200, N)[:, None]
        nb = ens.set_normalize_bounds(X, "meanstd")
        Xd = ctx.to_device(X)
        ms, _ = timed(lambda: ensemble_predict_device(ctx, members, layers, Xd, 0.01, "meanstd", nb))
        kms = ctx.ms(_lib.T_PREDICT)
        dims = layers[:-1] + [2 * layers[-1]]
        flops = 2 * N * M * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
        r[str(layers)] = {"call_ms": ms, "kernel_ms": kms, "samples_per_s": N / (kms * 1e-3), "tflops": flops / (kms * 1e-3) / 1e12,
                          "frac_fp64_peak": flops / (kms * 1e-3) / 1e12 / pk.value}
    OUT["c4_ensemble_M8_N1e6"] = r; print("c4", r, flush=True)

if want("c2"):
    w = wl.ackley(400, 400, 1000)
    Ud, _, _ = snapshots_device(ctx, "ackley", w["x_mesh"][:, 1:].T, w["mu"])
    r = {}
    for mode in ("gram2", "accurate", "tsqr"):
        ms, (V, sig) = timed(lambda: pod_device(ctx, Ud, 1e-10, 0, mode))
        r[mode] = {"ms": ms, "n_L": V.shape[1], "gram_ms": ctx.ms(_lib.T_GRAM) if mode != "tsqr" else 0., "eig_ms": ctx.ms(_lib.T_EIG),
                   "pass2_ms": ctx.ms(_lib.T_PASS2), "qr_ms": ctx.ms(_lib.T_QR) if mode == "tsqr" else 0.}
        if mode == "tsqr":      # Householder flop count 2 n_h n_s^2 - (2/3) n_s^3
            r[mode]["qr_tflops"] = (2 * 160000 * 1000 ** 2 - 2 / 3 * 1000 ** 3) / (ctx.ms(_lib.T_QR) * 1e-3) / 1e12
    OUT["c2_ackley_160000x1000"] = r; print("c2", r, flush=True)
    Ud.release()

if want("c5"):
    r = {}
    for n_h, n_s in ((250_000, 4096), (1_250_000, 4096)):
        Ud = ctx.alloc(n_h, n_s)
        t = time.time(); ctx.check(L.pb_dct_lowrank(ctx.h, n_h, n_s, 256, 0.25, 0, n_h, Ud.ptr, Ud.ld)); ctx.sync()
        gen_s = time.time() - t
        rr = {"gen_s": gen_s}
        for mode in ("gram", "gram2", "tsqr"):
            ms, (V, sig) = timed(lambda: pod_device(ctx, Ud, 1e-10, 0, mode), reps=2 if mode != "tsqr" else 1)
            if mode == "tsqr":
                qr = ctx.ms(_lib.T_QR)
                rr[mode] = {"ms": ms, "n_L": V.shape[1], "qr_ms": qr, "eig_ms": ctx.ms(_lib.T_EIG), "basis_ms": ctx.ms(_lib.T_BASIS),
                            "qr_tflops": (2 * n_h * n_s ** 2 - 2 / 3 * n_s ** 3) / (qr * 1e-3) / 1e12,
                            "sig_err_vs_closed_form": float(np.max(np.abs(sig[:V.shape[1]] / sig[0] - 2. ** (-0.25 * np.arange(V.shape[1])))))}
                del V
                continue
            gk = ctx.ms(_lib.T_TN_KERNEL)
            rr[mode] = {"ms": ms, "n_L": V.shape[1], "gram_kernel_ms": gk, "gram_tflops": n_h * n_s * (n_s + 1) / (gk * 1e-3) / 1e12,
                        "gram_frac": n_h * n_s * (n_s + 1) / (gk * 1e-3) / 1e12 / pk.value, "eig_ms": ctx.ms(_lib.T_EIG),
                        "pass2_ms": ctx.ms(_lib.T_PASS2), "basis_ms": ctx.ms(_lib.T_BASIS),
                        "pod_tflops": (n_h * n_s * (n_s + 1) + 2 * n_h * n_s * V.shape[1]) / (ms * 1e-3) / 1e12,
                        "sig_err_vs_closed_form": float(np.max(np.abs(sig[:V.shape[1]] / sig[0] - 2. ** (-0.25 * np.arange(V.shape[1])))))}
            del V
        r[f"{n_h}x{n_s}"] = rr; print("c5", n_h, rr, flush=True)
        Ud.release()
    OUT["c5_slab"] = r

os.makedirs("gpurun_out", exist_ok=True)
json.dump(OUT, open("gpurun_out/bench_extra.json", "w"), indent=1)
