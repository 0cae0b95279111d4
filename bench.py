#!/usr/bin/env python
"""bench.py -- the POD-UQNN hot path on B200 (BASELINE.json metric: fp64 POD TFLOP/s at n_h x n_s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): C2 of BASELINE.json, "2d_ackley-shaped: 400x400 grid (n_h = 160 000)
x n_s = 1000 samples, fp64 Gram POD + projection V^T U".  One step = one pass of the path over one
batch: Gram U^T U (DMMA) -> Cholesky/Jacobi eigensolve -> rank rule -> second Gram pass on the
rotated block ("gram2") -> basis V -> projection v = (V^T U)^T.  At N GPUs every rank owns a
160 000-row slab of a 400 x (400 N) mesh (weak scaling); the Gram / projection partials are summed
with NCCL all-reduces, the small eigensolve is replicated.

value  : algorithmic TFLOP/s, (n_h n_s (n_s+1) + 4 n_h n_s n_L) / wall, U resident in HBM.
e2e    : the same metric through the reference-facing call with HOST buffers: the step uploads the
         U slab from pinned host memory and downloads V and v.
roofline: the split-K DMMA Gram kernel against the fp64 tensor peak measured live in this process
         (MEASURED_PEAKS.json carries no fp64 figure).
cpu_baseline / --impl reference: the reference's algorithm (numpy.linalg.svd -> LAPACK dgesdd, the
         routine numba calls in pod.py:16, + the V and projection products) from oracle/, on the
         host cores, on a bounded row sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_X, N_Y, N_S, EPS, MODE = 400, 400, 1000, 1e-10, "gram2"
METRIC, UNIT = "pod_fp64_tflops", "TFLOP/s"
CPU_SAMPLE_ROWS = 160000   # the whole C2 matrix: ~7 s per pass on the 16 host cores of the GPU box


def algo_flops(n_h, n_s, n_L):
    """SYRK count of the Gram + basis + projection."""
    return n_h * n_s * (n_s + 1) + 4 * n_h * n_s * n_L


def workload_config(n_gpus):
    return {"workload": f"C2 2d_ackley-shaped: {N_X}x{N_Y * n_gpus} grid (n_h = {N_X * N_Y * n_gpus}) x n_s = {N_S}, "
                        f"fp64 Gram POD (mode {MODE}, eps = {EPS:g}) + projection V^T U",
            "n_h_per_gpu": N_X * N_Y, "n_s": N_S, "eps": EPS, "pod_mode": MODE,
            "parallelism": f"row-sharded x{n_gpus}" if n_gpus > 1 else "single GPU",
            "l2_policy": "inputs larger than L2 (U slab = 1.28 GB vs 126 MB L2); no flush needed"}


# ----------------------------------------------------------------------------- CPU arm
def cpu_reference_steps(steps, warmup, rows=CPU_SAMPLE_ROWS):
    """The oracle port of perform_pod + projection on the first `rows` mesh nodes of C2."""
    from oracle import pod_oracle as po
    from pod_uqnn_b200 import workloads as wl
    w = wl.ackley(N_X, N_Y, N_S)
    x_mesh = w["x_mesh"][:rows]
    _, U, _, _ = po.create_snapshots(x_mesh, 1, 0, po.u_ackley, w["mu"])
    times, n_L = [], 0
    for i in range(warmup + steps):
        t = time.perf_counter()
        V = po.perform_pod(U, EPS)
        v = po.project_to_v(V, U)
        dt = time.perf_counter() - t
        n_L = V.shape[1]
        if i >= warmup:
            times.append(dt)
        del v
    sec = statistics.mean(times)
    return algo_flops(rows, N_S, n_L) / sec / 1e12, sec, n_L


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    tf, sec, n_L = cpu_reference_steps(steps, warmup)
    sample = (f"the whole C2 matrix, {CPU_SAMPLE_ROWS} mesh rows x {N_S} snapshots (n_L = {n_L}), "
              f"{steps} timed step(s) after {warmup} warm-up; TFLOP/s uses the same algorithmic flop count")
    line = {"impl": "reference", "metric": METRIC, "value": tf, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": tf, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
            "e2e": {"value": tf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "oracle port of pod.py:7-47 + podnnmodel.py:247 (numpy.linalg.svd = LAPACK dgesdd, all host threads); "
                    "/root/reference is not present on the GPU box"}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 6]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], 0, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1])); pw.append(float(r[2]))
            except ValueError:
                continue
            for n, v in zip(names, r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        load = [s for s, p in zip(sm, pw) if p >= 0.6 * max(pw)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": mx, "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import ctypes as C
    from pod_uqnn_b200 import _lib, sharded, workloads as wl
    from pod_uqnn_b200.acceleration import snapshots_device

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # rank 0 prints exactly ONE JSON line on stdout: keep library chatter (NCCL banner ...) off it
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = _lib.Context(local)
    L = _lib.lib()
    be = sharded.CudaBackend(ctx) if world > 1 else None

    # ---- synthetic input: this rank's slab of the Ackley snapshot matrix, evaluated in HBM
    w = wl.ackley(N_X, N_Y * world, N_S)
    n_h_loc = N_X * N_Y
    X = w["x_mesh"][:, 1:].T
    U, _, _ = snapshots_device(ctx, "ackley", X, w["mu"], row0=rank * n_h_loc, n_rows=n_h_loc)
    U.release()                              # first call: table / workspace allocation; time the second
    U, _, _ = snapshots_device(ctx, "ackley", X, w["mu"], row0=rank * n_h_loc, n_rows=n_h_loc)
    snap_ms = ctx.ms(_lib.T_SNAPSHOTS)
    n_h = n_h_loc * world

    state = {}

    def step_resident():
        if world == 1:
            nl = C.c_int()
            ctx.check(L.pb_pod(ctx.h, U.ptr, n_h_loc, N_S, U.ld, EPS, 0, _lib.POD_MODES[MODE], C.byref(nl)))
            n_L = nl.value
            if state.get("n_L") != n_L:
                state["V"], state["v"], state["n_L"] = ctx.alloc(n_h_loc, n_L), ctx.alloc(N_S, n_L), n_L
            V, v = state["V"], state["v"]
            ctx.check(L.pb_pod_basis_full(ctx.h, U.ptr, n_h_loc, N_S, U.ld, V.ptr, V.ld))
            ctx.check(L.pb_project(ctx.h, V.ptr, V.ld, n_L, U.ptr, U.ld, n_h_loc, N_S, v.ptr))
        else:
            V, _, n_L = sharded.pod_sharded(be, U, EPS, 0, MODE)
            v = sharded.project_sharded(be, V, U)
            state["V"], state["v"], state["n_L"] = V, v, n_L
        return state["n_L"]

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- fp64 peak of this GPU, measured live (roofline denominator)
    pk = C.c_double()
    ctx.check(L.pb_probe_fp64_peak(ctx.h, 1, C.byref(pk)))
    peak_tf = pk.value

    # ---- device-resident timing
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launches()
    gram_ms, parts = [], {"gram": [], "eig": [], "pass2": [], "basis": [], "project": []}
    ctx.timer_start()
    for _ in range(args.steps):
        n_L = step_resident()
        gram_ms.append(ctx.ms(_lib.T_TN_KERNEL))
        for k, t in (("gram", _lib.T_GRAM), ("eig", _lib.T_EIG), ("pass2", _lib.T_PASS2), ("basis", _lib.T_BASIS),
                     ("project", _lib.T_PROJECT)):
            parts[k].append(ctx.ms(t))
    ms_total = ctx.timer_stop()
    barrier()
    launches = ctx.launches() - launches0
    clocks = sampler.stop() if sampler else None
    ms_step = max_over_ranks(ms_total / args.steps)
    value = algo_flops(n_h, N_S, n_L) / (ms_step * 1e-3) / 1e12

    # ---- end to end: host buffers in, host buffers out (pinned), copies inside the timed region
    hU, hV, hv = C.c_void_p(), C.c_void_p(), C.c_void_p()
    ctx.check(L.pb_host_alloc(U.nbytes, C.byref(hU)))
    ctx.check(L.pb_host_alloc(n_h_loc * n_L * 8, C.byref(hV)))
    ctx.check(L.pb_host_alloc(N_S * n_L * 8, C.byref(hv)))
    ctx.check(L.pb_download(ctx.h, hU, U.ptr, U.nbytes))

    def step_e2e():
        if world == 1:
            # the reference-facing host call: chunked upload overlapped with the Gram (pb_pod_host)
            nl_c = C.c_int()
            ctx.check(L.pb_pod_host(ctx.h, hU, n_h_loc, N_S, N_S, U.ptr, EPS, 0, _lib.POD_MODES[MODE], C.byref(nl_c)))
            nl = nl_c.value
            V, v = state["V"], state["v"]
            ctx.check(L.pb_pod_basis_full(ctx.h, U.ptr, n_h_loc, N_S, U.ld, V.ptr, V.ld))
            ctx.check(L.pb_project(ctx.h, V.ptr, V.ld, nl, U.ptr, U.ld, n_h_loc, N_S, v.ptr))
            vptr = v.ptr
        else:
            # row-sharded host call: each rank's chunked upload overlaps its Gram (pb_gram_host), then the same
            # all-reduces as the resident step
            V, _, nl = sharded.pod_sharded(be, U, EPS, 0, MODE, A_host=hU)
            v = sharded.project_sharded(be, V, U)
            state["V"], state["v"], state["n_L"] = V, v, nl
            vptr = v.data_ptr()
        ctx.check(L.pb_download_async(ctx.h, hV, state["V"].ptr, n_h_loc * nl * 8))
        ctx.check(L.pb_download_async(ctx.h, hv, vptr, N_S * nl * 8))
        ctx.sync()

    step_e2e()
    barrier()
    e2e_steps = max(2, min(args.steps, 10))
    ctx.timer_start()
    for _ in range(e2e_steps):
        step_e2e()
    ms_e2e = max_over_ranks(ctx.timer_stop() / e2e_steps)
    barrier()
    e2e_value = algo_flops(n_h, N_S, n_L) / (ms_e2e * 1e-3) / 1e12
    L.pb_host_free(hU); L.pb_host_free(hV); L.pb_host_free(hv)

    # ---- second headline metric: ensemble predict samples/s (C4: M = 8 MLPs [1,128,128,128,2*64], 1e6 points
    #      split over the ranks, no communication)
    # ---- the TSQR accuracy mode on the same resident slab(s): Householder QR per rank, all-gather of the R factors,
    #      replicated QR of the stack + pivoted QR + Jacobi SVD, basis, projection
    def step_tsqr():
        if world == 1:
            nl = C.c_int()
            ctx.check(L.pb_pod(ctx.h, U.ptr, n_h_loc, N_S, U.ld, EPS, 0, _lib.POD_MODES["tsqr"], C.byref(nl)))
            if state.get("t_nL") != nl.value:
                state["tV"], state["tv"], state["t_nL"] = ctx.alloc(n_h_loc, nl.value), ctx.alloc(N_S, nl.value), nl.value
            V, v = state["tV"], state["tv"]
            ctx.check(L.pb_pod_basis_full(ctx.h, U.ptr, n_h_loc, N_S, U.ld, V.ptr, V.ld))
            ctx.check(L.pb_project(ctx.h, V.ptr, V.ld, nl.value, U.ptr, U.ld, n_h_loc, N_S, v.ptr))
            return nl.value
        V, _, nl = sharded.pod_sharded(be, U, EPS, 0, "tsqr")
        sharded.project_sharded(be, V, U)
        return nl
    step_tsqr(); barrier()
    t_best = 1e30
    for _ in range(3):
        barrier(); ctx.timer_start(); nl_t = step_tsqr(); t_best = min(t_best, ctx.timer_stop())
    barrier()
    t_tsqr = max_over_ranks(t_best)
    qr_flops = world * (2 * n_h_loc * N_S ** 2 - 2 / 3 * N_S ** 3)
    tsqr_line = {"workload": "same slabs, mode tsqr (blocked Householder QR of U, pivoted QR + Jacobi SVD of R), basis, projection",
                 "ms_per_step": t_tsqr, "n_L": nl_t, "local_qr_ms": ctx.ms(_lib.T_QR),
                 "householder_tflops": qr_flops / (t_tsqr * 1e-3) / 1e12,
                 "note": "flop count 2 n_h n_s^2 - (2/3) n_s^3 per slab (SURVEY 8d); at N > 1 T_QR is the replicated QR of the stacked R factors"}

    ens_line = ensemble_bench(ctx, L, world, rank, max_over_ranks, peak_tf, with_cpu=(world == 1 and not args.no_cpu_baseline))
    train_lines = train_bench(ctx, world, rank, max_over_ranks, peak_tf, with_cpu=(world == 1 and not args.no_cpu_baseline))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (the split-K DMMA Gram kernel)
    gram_kernel_ms = statistics.mean(gram_ms)
    gram_flops = n_h_loc * N_S * (N_S + 1)
    achieved = gram_flops / (gram_kernel_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "gram_streamk_kernel (Gram U^T U, fp64 DMMA m8n8k4, TMA + mbarrier ring)", "achieved": achieved,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": load_traffic(),
                "algorithmic_flops_per_launch": gram_flops, "kernel_ms": gram_kernel_ms,
                "kernel_share_of_step": gram_kernel_ms / ms_step,
                "peak_source": "live pb_probe_fp64_peak (DMMA m8n8k4 register loop on all SMs, burst); "
                               "MEASURED_PEAKS.json has no fp64 entry (nominal 37.2 TFLOP/s at 1965 MHz)"}

    # ---- CPU baseline (N = 1 only): the oracle port on a bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        tf, sec, nl_cpu = cpu_reference_steps(1, 0)
        cpu = {"value": tf, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"the whole C2 matrix ({CPU_SAMPLE_ROWS} of {n_h_loc} mesh rows x {N_S} snapshots), 1 pass of the oracle's "
                         f"perform_pod + projection ({sec:.2f} s, n_L = {nl_cpu}); same flop formula"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world), "n_L": n_L,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "steps": e2e_steps,
                    "h2d_bytes_per_step": U.nbytes * world, "d2h_bytes_per_step": (n_h + N_S) * n_L * 8},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "stage_ms": {k: statistics.mean(v) for k, v in parts.items()}, "tsqr_mode": tsqr_line,
            "ensemble_predict": ens_line, "ensemble_train": train_lines,
            "snapshot_kernel": {"ms": snap_ms, "gbs": U.nbytes / (snap_ms * 1e-3) / 1e9,
                                "note": "Ackley field -> U slab in HBM (outside the timed step): table kernels + ackley_indexed_kernel, second call"}}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


def ensemble_bench(ctx, L, world, rank, max_over_ranks, peak_tf, with_cpu):
    """C4: predict_v of M = 8 members on 1e6 parameter points (batch split over the ranks)."""
    from pod_uqnn_b200 import _lib
    from pod_uqnn_b200.varneuralnetwork import ensemble_predict_device, glorot_normal
    layers, M, N_total = [1, 128, 128, 128, 64], 8, 1_000_000
    lo, hi = N_total * rank // world, N_total * (rank + 1) // world
    dims_w = layers[:-1] + [2 * layers[-1]]
    members = []
    for i in range(M):                      # synthetic glorot_normal members (Keras initialiser), zero biases
        rng = np.random.default_rng(4000 + i)
        members.append([a for fi, fo in zip(dims_w[:-1], dims_w[1:]) for a in (glorot_normal(rng, fi, fo), np.zeros(fo))])
    X = np.linspace(800, 1200, N_total)[:, None]
    nb = (X.mean(0), X.std(0))              # NORM_MEANSTD bounds, varneuralnetwork.py:70-73
    Xd = ctx.to_device(X[lo:hi])
    best = 1e30
    for _ in range(3):
        vm, vs = ensemble_predict_device(ctx, members, layers, Xd, 0.01, "meanstd", nb)
        best = min(best, ctx.ms(_lib.T_PREDICT))
        del vm, vs
    ms = max_over_ranks(best)
    dims = layers[:-1] + [2 * layers[-1]]
    flops = 2 * N_total * M * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    out = {"workload": "C4 ensemble predict_v: M = 8 x MLP [1,128,128,128,2*64], 1e6 points, fp64, batch split over ranks",
           "samples_per_s": N_total / (ms * 1e-3), "kernel_ms": ms, "tflops": flops / (ms * 1e-3) / 1e12,
           "frac_fp64_peak": flops / (ms * 1e-3) / 1e12 / (peak_tf * world)}
    if with_cpu and rank == 0:
        from oracle import ensemble as ens      # CPU baseline leg only
        n_cpu = 20000
        t = time.perf_counter()
        ens.predict_v(X[:n_cpu], members, layers[-1], 0.01, "meanstd", nb)
        sec = time.perf_counter() - t
        out["cpu_baseline"] = {"samples_per_s": n_cpu / sec, "cores": os.cpu_count(), "kind": "port",
                               "sample": f"{n_cpu} points through the numpy restatement (TensorFlow is not in the image)"}
    return out


def train_bench(ctx, world, rank, max_over_ranks, peak_tf, with_cpu):
    """full-batch ensemble training epochs (varneuralnetwork.py:94-128) at the reference's own training
    sizes; every rank trains its own n_M members (the reference gives one member to each Horovod rank)."""
    from pod_uqnn_b200.varneuralnetwork import EnsembleTrainer, glorot_normal
    cases = [("2d_ackley: [3,128,128,128,2*14], N = 400, n_M = 5 per rank, adv_eps 0", [3, 128, 128, 128, 14], 400, 5, 0., 1e-3, 1e-3, 0.01, 1000),
             ("1dt_burger: [2,128,128,128,2*37], N = 4000, n_M = 5 per rank, adv_eps 0.01", [2, 128, 128, 128, 37], 4000, 5, 0.01, 0.01, 1e-8, 0.01, 300)]
    out = []
    for ci, (name, layers, N, M, adv, lr, lam, soft_0, epochs) in enumerate(cases):
        rng = np.random.default_rng(7000 + 10 * ci + rank)
        dims = layers[:-1] + [2 * layers[-1]]
        members = [[a for fi, fo in zip(dims[:-1], dims[1:]) for a in (glorot_normal(rng, fi, fo), np.zeros(fo))]
                   for _ in range(M)]
        X = rng.uniform(-1.7, 1.7, (N, layers[0]))
        v = rng.standard_normal((N, layers[-1]))
        tr = EnsembleTrainer(ctx, members, layers, X, v, lr, lam, adv, soft_0)
        tr.run(50)
        ctx.sync()
        ctx.timer_start()
        losses = tr.run(epochs)
        ms = max_over_ranks(ctx.timer_stop())
        mac = sum(a * b for a, b in zip(dims[:-1], dims[1:]))
        flops = 2 * 3 * 2 * N * mac * M * world        # 2 loss terms x (forward + delta chain + weight gradient)
        line = {"workload": name, "us_per_epoch": 1e3 * ms / epochs, "member_epochs_per_s": M * world * epochs / (ms * 1e-3),
                "tflops_equiv": flops / (ms * 1e-3 / epochs) / 1e12, "epochs": epochs,
                "loss_first_last": [float(losses[0].mean()), float(losses[-1].mean())],
                "launches_per_epoch": "CUDA graph replay, see DESIGN.md section 9"}
        if with_cpu and rank == 0:
            from oracle import train as otr           # CPU baseline leg only
            n_ep = 10 if N <= 1000 else 3
            t = time.perf_counter()
            otr.train(X, v, [a.copy() for a in members[0]], layers[-1], n_ep, lr, lam, adv, soft_0)
            sec = time.perf_counter() - t
            line["cpu_baseline"] = {"member_epochs_per_s": n_ep / sec, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{n_ep} epochs of one member through the numpy restatement"}
        out.append(line)
    return out


def load_traffic():
    """dram bytes per launch of the Gram kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "gram_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="podb200", choices=["podb200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "podb200" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus != world and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    run_gpu(args)


if __name__ == "__main__":
    main()
