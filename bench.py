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
         U slab from pinned host memory and downloads V and v.  e2e_api: the Python entry points
         pod.perform_pod(ndarray) + PodnnModel.project_to_v(ndarray) on ordinary (pageable) numpy
         arrays, timed with time.perf_counter around the calls.
roofline: the split-K DMMA Gram kernel against the fp64 tensor peak measured live in this process
         (MEASURED_PEAKS.json carries no fp64 figure).
c5     : BASELINE.json configs[4] (the north-star target): the DCT known-answer slab, 1.25 M rows
         per GPU x 4096 snapshots, modes gram2 AND tsqr, POD + projection, with the closed-form
         rank / singular values / modes ASSERTED (the run fails otherwise).
parity : at N > 1 the sharded C2 result is compared on rank 0 with a single-GPU run over the same
         rows (rank, singular values, rank 0's rows of V); the run fails on a mismatch.
cpu_baseline / --impl reference: the reference's OWN perform_pod (oracle/_ref, byte copies of
         poduqnn/pod.py made by oracle/build_ref.py; numba -> LAPACK dgesdd) + the projection of
         podnnmodel.py:437, on all host cores, on the same matrix.  Falls back to the numpy port in
         oracle/pod_oracle.py (kind "port") only when oracle/_ref was never built.
"""
import os
import sys

# ---- host threads of the CPU legs: decided BEFORE numpy / numba load their thread pools.  torchrun exports
#      OMP_NUM_THREADS=1 to every rank; the reference arm must use all host cores (rank 0 alone works).
try:
    HOST_CORES = len(os.sched_getaffinity(0))
except AttributeError:
    HOST_CORES = os.cpu_count() or 1
_IMPL_REFERENCE = any(a == "reference" or a == "--impl=reference" for a in sys.argv[1:])
if _IMPL_REFERENCE or int(os.environ.get("WORLD_SIZE", "1")) == 1:
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMBA_NUM_THREADS"):
        os.environ[_k] = str(HOST_CORES)

import argparse
import json
import statistics
import subprocess
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_X, N_Y, N_S, EPS, MODE = 400, 400, 1000, 1e-10, "gram2"
METRIC, UNIT = "pod_fp64_tflops", "TFLOP/s"
C5_ROWS_PER_GPU, C5_NS, C5_R, C5_DECAY = 1_250_000, 4096, 256, 0.25


def algo_flops(n_h, n_s, n_L):
    """SYRK count of the Gram + basis + projection."""
    return n_h * n_s * (n_s + 1) + 4 * n_h * n_s * n_L


def workload_config(n_gpus):
    return {"workload": f"C2 2d_ackley-shaped: {N_X}x{N_Y * n_gpus} grid (n_h = {N_X * N_Y * n_gpus}) x n_s = {N_S}, "
                        f"fp64 Gram POD (mode {MODE}, eps = {EPS:g}) + projection V^T U",
            "n_h_per_gpu": N_X * N_Y, "n_s": N_S, "eps": EPS, "pod_mode": MODE,
            "parallelism": f"row-sharded x{n_gpus}" if n_gpus > 1 else "single GPU",
            "l2_policy": "inputs larger than L2 (U slab = 1.28 GB vs 126 MB L2); no flush needed"}


def blas_threads():
    """Threads the BLAS / numba pools really run with (reported next to every CPU number)."""
    out = {"host_cores": HOST_CORES}
    try:
        from threadpoolctl import threadpool_info
        out["blas"] = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:  # noqa: BLE001
        out["blas"] = None
    try:
        import numba
        out["numba"] = numba.get_num_threads()
    except Exception:  # noqa: BLE001
        out["numba"] = None
    return out


# ----------------------------------------------------------------------------- CPU arm
def ackley_matrix_host(n_gpus, rows=None):
    """The C2 snapshot matrix of the weak-scaled workload (400 x 400 N mesh, the same mu), built on the host by
    broadcasting the reference's field formula (2d_ackley/hyperparams.py:51-57) over row chunks.  Input generation
    only: never inside a timed region."""
    from pod_uqnn_b200 import workloads as wl
    w = wl.ackley(N_X, N_Y * n_gpus, N_S)
    xm = w["x_mesh"] if rows is None else w["x_mesh"][:rows]
    mu = w["mu"]
    m0, m1, m2 = mu[:, 0][None, :], mu[:, 1][None, :], mu[:, 2][None, :]
    U = np.empty((xm.shape[0], N_S))
    for lo in range(0, xm.shape[0], 20000):
        x, y = xm[lo:lo + 20000, 1][:, None], xm[lo:lo + 20000, 2][:, None]
        U[lo:lo + 20000] = (-20 * (1 + .1 * m2) * np.exp(-.2 * (1 + .1 * m1) * np.sqrt(.5 * (x ** 2 + y ** 2)))
                            - np.exp(.5 * (np.cos(2 * np.pi * (1 + .1 * m0) * x) + np.cos(2 * np.pi * (1 + .1 * m0) * y)))
                            + 20 + np.exp(1))
    return U


class CpuPod:
    """perform_pod + projection on the host cores: the reference's own code when oracle/_ref is built."""

    def __init__(self):
        from oracle import build_ref
        self.kind = "reference" if build_ref.available() else "port"
        if self.kind == "reference":
            self.pod_mod, _ = build_ref.load()
            # numba compiles perform_pod on first use (~7 s): warm it on a tiny matrix of the same dtype / layout
            self.pod_mod.perform_pod(np.ascontiguousarray(np.random.default_rng(0).random((64, 8))), EPS, 0, False)
        else:
            from oracle import pod_oracle as po
            self.po = po

    def step(self, U):
        t = time.perf_counter()
        if self.kind == "reference":
            V = self.pod_mod.perform_pod(U, EPS, 0, False)        # poduqnn/pod.py:7-47, unmodified
            v = (V.T.dot(U)).T                                    # podnnmodel.py:437
        else:
            V = self.po.perform_pod(U, EPS)
            v = self.po.project_to_v(V, U)
        dt = time.perf_counter() - t
        return dt, V.shape[1], v.shape

    def describe(self):
        if self.kind == "reference":
            return ("the reference's own poduqnn/pod.py:perform_pod (byte copy in oracle/_ref, numba -> LAPACK dgesdd) "
                    "+ (V.T.dot(U)).T of podnnmodel.py:437")
        return ("numpy port of pod.py:7-47 + podnnmodel.py:437 (oracle/pod_oracle.py; numpy.linalg.svd = the same LAPACK "
                "dgesdd); oracle/_ref was not built")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_gpus = max(1, args.gpus)
    rows_full = N_X * N_Y * n_gpus
    rows, note = rows_full, ""
    try:
        import psutil
        avail = psutil.virtual_memory().available
        if 6 * rows_full * N_S * 8 > avail:                       # U, its Fortran-order copy, the thin left factor, V ...
            rows = N_X * N_Y
            note = (f"; host memory ({avail / 2**30:.0f} GiB free) cannot hold the SVD workspace of the {rows_full}-row matrix: "
                    f"ran ONE {rows}-row slab, TFLOP/s is per-slab work over per-slab time")
    except ImportError:
        pass
    cpu = CpuPod()
    U = ackley_matrix_host(n_gpus, rows)
    threads = blas_threads()
    dt, n_L, _ = cpu.step(U)                                      # warm-up (page faults, pools)
    steps = max(1, min(args.steps, 3, int(150. / max(dt, 1e-3))))
    times = [cpu.step(U)[0] for _ in range(steps)]
    sec = statistics.mean(times)
    tf = algo_flops(rows, N_S, n_L) / sec / 1e12
    cfg = workload_config(n_gpus)
    cfg["rows_factored"] = rows
    sample = (f"the whole weak-scaled C2 matrix, {rows} mesh rows x {N_S} snapshots (n_L = {n_L}), {steps} timed step(s) "
              f"after 1 warm-up{note}; TFLOP/s uses the same algorithmic flop count")
    line = {"impl": "reference", "metric": METRIC, "value": tf, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": tf, "unit": UNIT, "cores": threads["blas"] or HOST_CORES, "threads": threads,
                             "kind": cpu.kind, "sample": sample},
            "e2e": {"value": tf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "note": cpu.describe()}
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

    def pod_single(Ud, mode, key):
        """single-GPU POD + basis + projection of a resident slab through the C-ABI; buffers cached under `key`."""
        rows, ns = Ud.shape
        nl = C.c_int()
        ctx.check(L.pb_pod(ctx.h, Ud.ptr, rows, ns, Ud.ld, EPS, 0, _lib.POD_MODES[mode], C.byref(nl)))
        n_L = nl.value
        if state.get(key + "n_L") != n_L or state.get(key + "rows") != rows:
            state[key + "V"], state[key + "v"] = ctx.alloc(rows, n_L), ctx.alloc(ns, n_L)
            state[key + "n_L"], state[key + "rows"] = n_L, rows
        V, v = state[key + "V"], state[key + "v"]
        ctx.check(L.pb_pod_basis_full(ctx.h, Ud.ptr, rows, ns, Ud.ld, V.ptr, V.ld))
        ctx.check(L.pb_project(ctx.h, V.ptr, V.ld, n_L, Ud.ptr, Ud.ld, rows, ns, v.ptr))
        return V, v, n_L

    def step_resident():
        if world == 1:
            V, v, n_L = pod_single(U, MODE, "")
        else:
            V, _, n_L = sharded.pod_sharded(be, U, EPS, 0, MODE)
            v = sharded.project_sharded(be, V, U)
        state["V"], state["v"], state["n_L"] = V, v, n_L
        return n_L

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

    def sum_over_ranks(a):
        """small host array summed over the ranks (checks only)"""
        if dist is None:
            return a
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

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
    ms_step = max_over_ranks(ms_total / args.steps)
    value = algo_flops(n_h, N_S, n_L) / (ms_step * 1e-3) / 1e12
    from pod_uqnn_b200.pod import pod_sigma
    sig_resident = pod_sigma(ctx)

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
            ctx.check(L.pb_download_forked(ctx.h, hV, V.ptr, n_h_loc * nl * 8))      # V crosses PCIe while V^T U runs
            ctx.check(L.pb_project(ctx.h, V.ptr, V.ld, nl, U.ptr, U.ld, n_h_loc, N_S, v.ptr))
            ctx.check(L.pb_download_async(ctx.h, hv, v.ptr, N_S * nl * 8))
            ctx.sync()
            return
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
    # the clock / throttle samples cover BOTH timed regions (device-resident steps and the e2e steps)
    clocks = sampler.stop() if sampler else None

    # ---- e2e through the Python API a reference user calls: ordinary numpy arrays in, numpy arrays out
    e2e_api = None
    if world == 1:
        from pod_uqnn_b200 import pod as pod_mod
        U_np = np.ctypeslib.as_array(C.cast(hU, C.POINTER(C.c_double)), shape=(n_h_loc, N_S)).copy()   # pageable copy
        pod_mod.perform_pod(U_np, EPS, 0, False, mode=MODE, ctx=ctx)        # warm: staging ring, device slab
        t_pod, t_prj, t_reuse = [], [], []
        for _ in range(3):
            t0 = time.perf_counter()
            V_np = pod_mod.perform_pod(U_np, EPS, 0, False, mode=MODE, ctx=ctx)
            t1 = time.perf_counter()
            v_np = pod_mod.project_to_v(V_np, U_np, ctx=ctx)
            t2 = time.perf_counter()
            t_pod.append(t1 - t0); t_prj.append(t2 - t1)
            t0 = time.perf_counter()
            V_np = pod_mod.perform_pod(U_np, EPS, 0, False, mode=MODE, ctx=ctx)
            v2_np = pod_mod.project_to_v(V_np, U_np, ctx=ctx, reuse_slab=True)
            t_reuse.append(time.perf_counter() - t0)
        assert np.allclose(v_np, v2_np, rtol=0, atol=1e-9 * np.abs(v_np).max())
        fl = algo_flops(n_h_loc, N_S, V_np.shape[1])
        sec_api, sec_reuse = min(t_pod) + min(t_prj), min(t_reuse)
        e2e_api = {"value": fl / sec_api / 1e12, "unit": UNIT, "ms_per_step": sec_api * 1e3, "mode": MODE,
                   "perform_pod_ms": min(t_pod) * 1e3, "project_to_v_ms": min(t_prj) * 1e3,
                   "call": "V = pod.perform_pod(U_ndarray, eps, mode=...); v = pod.project_to_v(V, U_ndarray): pageable numpy in, "
                           "numpy out, time.perf_counter, best of 3 (U crosses PCIe twice, V three times)",
                   "reuse_slab": {"value": fl / sec_reuse / 1e12, "ms_per_step": sec_reuse * 1e3, "ratio_to_e2e": ms_e2e / (sec_reuse * 1e3),
                                  "call": "the same with project_to_v(..., reuse_slab=True): U and V stay in HBM between the two calls"},
                   "ratio_to_e2e": ms_e2e / (sec_api * 1e3), "v_shape": list(v_np.shape)}
        # the API default (mode "tsqr"): leaf QRs of the row chunks overlap the upload; PB_TSQR_HOST_OVERLAP=0 = upload, then pb_pod
        t_def = []
        pod_mod.perform_pod(U_np, EPS, 0, False, ctx=ctx)
        for _ in range(3):
            t0 = time.perf_counter(); V_def = pod_mod.perform_pod(U_np, EPS, 0, False, ctx=ctx); t_def.append(time.perf_counter() - t0)
        e2e_api["default_mode_tsqr"] = {"perform_pod_ms": min(t_def) * 1e3, "n_L": int(V_def.shape[1]),
                                        "call": "V = pod.perform_pod(U_ndarray, eps): pageable numpy in, numpy out, best of 3"}
        del U_np
    # ---- the host -> device ceiling of this box with every rank copying at once (explains the e2e scaling)
    ctx.sync(); barrier()
    ctx.timer_start()
    ctx.check(L.pb_upload_async(ctx.h, U.ptr, hU, U.nbytes))
    ms_h2d = max_over_ranks(ctx.timer_stop())
    barrier()
    h2d = {"aggregate_gbs": n_h * N_S * 8 / (ms_h2d * 1e-3) / 1e9, "ms": ms_h2d,
           "e2e_floor_ms": ms_h2d, "note": "plain pinned cudaMemcpyAsync of every rank's 1.28 GB slab at the same time, max over ranks"}
    L.pb_host_free(hU); L.pb_host_free(hV); L.pb_host_free(hv)

    # ---- N > 1: the sharded result against a single-GPU run over the same rows, on rank 0 (fails the run on a mismatch)
    parity = None
    if world > 1:
        parity = parity_vs_single_gpu(ctx, be, sharded, snapshots_device, X, w["mu"], U, n_h, rank, state, pod_single, barrier)

    # ---- the TSQR accuracy mode on the same resident slab(s): rank-revealing Householder QR per rank (block pivoting,
    #      stops at the numerical rank), all-gather of the truncated factors, replicated factor of the stack + pivoted QR
    #      + Jacobi SVD, basis, projection.  "tsqr_full" = the same without the early stop (every 64-column panel).
    def qr_stats():
        if world > 1:
            st = be.qr_stats
            return {"panels_local": st["local"][0], "factor_rows_local": st["local"][1],
                    "flops_executed": world * st["local"][2] + st.get("stack", (0, 0, 0.))[2]}
        p_, f_, fl_ = C.c_int64(), C.c_int64(), C.c_double()
        ctx.check(L.pb_tsqr_stats(ctx.h, C.byref(p_), C.byref(f_), C.byref(fl_)))
        return {"panels_local": p_.value, "factor_rows_local": f_.value, "flops_executed": fl_.value}

    tsqr_line = {"workload": "same slabs, TSQR accuracy mode (never forms U^T U): blocked Householder QR of U, pivoted QR + Jacobi "
                             "SVD of the factor, basis, projection"}
    qr_flops_full = world * (2 * n_h_loc * N_S ** 2 - 2 / 3 * N_S ** 3)
    for tmode in ("tsqr", "tsqr_full"):
        def step_tsqr():
            if world == 1:
                return pod_single(U, tmode, "t")[2]
            V, _, nl = sharded.pod_sharded(be, U, EPS, 0, tmode)
            sharded.project_sharded(be, V, U)
            return nl
        step_tsqr(); barrier()
        t_best = 1e30
        for _ in range(3):
            barrier(); ctx.timer_start(); nl_t = step_tsqr(); t_best = min(t_best, ctx.timer_stop())
        barrier()
        t_tsqr = max_over_ranks(t_best)
        st = qr_stats()
        tsqr_line[tmode] = {"ms_per_step": t_tsqr, "n_L": nl_t, "qr_ms_last_factor": ctx.ms(_lib.T_QR), **st,
                            "executed_tflops": st["flops_executed"] / (t_tsqr * 1e-3) / 1e12,
                            "frac_fp64_peak_executed": st["flops_executed"] / (t_tsqr * 1e-3) / 1e12 / (peak_tf * world),
                            "full_householder_equiv_tflops": qr_flops_full / (t_tsqr * 1e-3) / 1e12}
        for k in [k for k in state if k.startswith("t")]:
            state.pop(k)
    tsqr_line["ms_per_step"] = tsqr_line["tsqr"]["ms_per_step"]
    tsqr_line["n_L"] = tsqr_line["tsqr"]["n_L"]
    tsqr_line["note"] = ("tsqr = the API default (rank-revealing: stops when every residual column is at the QR's rounding level); "
                         "flops_executed = Householder count of the panels actually run (pb_tsqr_stats); full count 2 n_h n_s^2 - "
                         "(2/3) n_s^3 per slab (SURVEY 8d)")

    ens_line = ensemble_bench(ctx, L, world, rank, max_over_ranks, peak_tf, with_cpu=(world == 1 and not args.no_cpu_baseline))
    train_lines = train_bench(ctx, world, rank, max_over_ranks, peak_tf, with_cpu=(world == 1 and not args.no_cpu_baseline))

    # ---- C5 (BASELINE.json configs[4], the north-star target): 1.25 M rows per GPU x 4096, both modes, asserted
    c5 = None
    if not args.skip_c5:
        U.release(); state.clear()
        if be is not None:
            be._cache.clear()
        c5 = c5_leg(ctx, L, be, sharded, wl, world, rank, barrier, max_over_ranks, sum_over_ranks, peak_tf, pod_single, state)

    # ---- N > 1: the same C2 step from ONE process (rank 0) driving all N GPUs through pb_multi_* (the library's own
    #      peer-memory collectives, no torch / NCCL on the path); the other ranks idle at the barrier meanwhile
    lib_multi = None
    if world > 1 and not args.skip_lib_multi:
        # the other ranks must wait on the HOST: an NCCL barrier would keep a spinning kernel on their GPUs, and rank 0's
        # kernels would be time-sliced against it (measured: 38 instead of 7 ms per step)
        import datetime
        barrier()
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            try:
                lib_multi = lib_multi_leg(world, X, w["mu"], n_h_loc, n_L, sig_resident, ms_step, snapshots_device)
            finally:
                store.set("podb200_lib_multi_done", "1")
        else:
            store.wait(["podb200_lib_multi_done"], datetime.timedelta(minutes=15))
        barrier()

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

    # ---- CPU baseline (N = 1 only): the reference's perform_pod + projection on the whole C2 matrix, one pass
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cp = CpuPod()
        Uh = ackley_matrix_host(1)
        threads = blas_threads()
        sec, nl_cpu, _ = cp.step(Uh)
        del Uh
        cpu = {"value": algo_flops(n_h_loc, N_S, nl_cpu) / sec / 1e12, "unit": UNIT, "cores": threads["blas"] or HOST_CORES,
               "threads": threads, "kind": cp.kind,
               "sample": f"the whole C2 matrix ({n_h_loc} mesh rows x {N_S} snapshots), 1 pass of {cp.describe()} "
                         f"({sec:.2f} s, n_L = {nl_cpu}); same flop formula"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world), "n_L": n_L,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "steps": e2e_steps,
                    "h2d_bytes_per_step": n_h * N_S * 8, "d2h_bytes_per_step": (n_h + N_S) * n_L * 8},
            "e2e_api": e2e_api, "h2d_probe": h2d,
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "stage_ms": {k: statistics.mean(v) for k, v in parts.items()}, "parity": parity, "tsqr_mode": tsqr_line, "c5": c5, "lib_multi": lib_multi,
            "ensemble_predict": ens_line, "ensemble_train": train_lines,
            "snapshot_kernel": {"ms": snap_ms, "gbs": n_h_loc * N_S * 8 / (snap_ms * 1e-3) / 1e9,
                                "note": "Ackley field -> U slab in HBM (outside the timed step): table kernel + radius-grouped kernel (or the coordinate-indexed one when a slab repeats few radii), second call"}}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


def parity_vs_single_gpu(ctx, be, sharded, snapshots_device, X, mu, U, n_h, rank, state, pod_single, barrier):
    """N > 1: the row-sharded NCCL result (both modes) against ONE GPU factoring all n_h rows (rank 0 regenerates the
    whole matrix in its HBM: 1.28 GB per rank of the job).  Rank, singular values to 1e-10 sigma_max, rank 0's rows of
    V up to the per-mode sign to 1e-8 (north_star tolerances).  Raises on a mismatch; the other ranks wait."""
    from pod_uqnn_b200.pod import pod_sigma
    out = {}
    n_loc = U.shape[0]
    for mode in (MODE, "tsqr"):
        V_sh, sig_sh, nL_sh = sharded.pod_sharded(be, U, EPS, 0, mode)
        V_sh_host = V_sh.download() if rank == 0 else None
        barrier()
        if rank == 0:
            Ufull, _, _ = snapshots_device(ctx, "ackley", X, mu, row0=0, n_rows=n_h)
            V1, _, nL1 = pod_single(Ufull, mode, "p")
            sig1 = pod_sigma(ctx)
            V1_host = V1.download()[:n_loc]
            Ufull.release()
            for k in [k for k in state if k.startswith("p")]:
                state.pop(k)
            smax = float(sig1[0])
            sig_err = float(np.max(np.abs(sig_sh[:nL1] - sig1[:nL1])) / smax)
            signs = np.sign(np.sum(V_sh_host * V1_host, axis=0))
            v_err = float(np.max(np.abs(V_sh_host * signs - V1_host))) if nL_sh == nL1 else float("inf")
            ok = (nL_sh == nL1) and sig_err <= 1e-10 and v_err <= 1e-8
            out[mode] = {"n_L_sharded": nL_sh, "n_L_single_gpu": nL1, "sigma_err_rel_sigma_max": sig_err,
                         "V_rows_rank0_max_abs_err": v_err, "status": "ok" if ok else "MISMATCH"}
        barrier()
    if rank == 0:
        bad = {m: o for m, o in out.items() if o["status"] != "ok"}
        if bad:
            raise SystemExit(f"bench.py: sharded result differs from the single-GPU run: {json.dumps(bad)}")
        out["status"] = "ok"
        out["how"] = ("rank 0 factors all n_h rows on one GPU (pb_pod) and compares with the NCCL row-sharded run: equal n_L, "
                      "sigma to 1e-10 sigma_max, rank 0's rows of V to 1e-8 up to sign")
    return out or None


def lib_multi_leg(world, X, mu, n_h_loc, n_L_ref, sig_ref, ms_torchrun, snapshots_device, steps=20):
    """The weak-scaled C2 step (POD + basis + projection on resident slabs) driven by ONE process over all N GPUs:
    pb_multi_pod / pb_multi_basis / pb_multi_project.  Checked against the torchrun + NCCL run of the same slabs
    (equal n_L, sigma to 1e-10 sigma_max); device time = max over the GPUs of the events around the region."""
    import ctypes as C
    from pod_uqnn_b200 import _lib
    L = _lib.lib()
    m = _lib.MultiContext(world)
    slabs = [snapshots_device(m.ctxs[p], "ackley", X, mu, row0=p * n_h_loc, n_rows=n_h_loc)[0] for p in range(world)]
    n_h = n_h_loc * world
    U_ptrs, rows, lds = _lib._ptr_array([s.ptr for s in slabs]), _lib._i64_array([n_h_loc] * world), _lib._i64_array([N_S] * world)
    out = {"workload": f"the same {world} resident C2 slabs, one process, pb_multi_* (peer-memory all-reduce over NVLink)"}
    for mode in (MODE, "tsqr"):
        nl = C.c_int()
        m.check(L.pb_multi_pod(m.h, U_ptrs, rows, N_S, lds, EPS, 0, _lib.POD_MODES[mode], C.byref(nl)))
        V = [m.ctxs[p].alloc(n_h_loc, nl.value) for p in range(world)]
        v = [m.ctxs[p].alloc(N_S, nl.value) for p in range(world)]
        V_ptrs, v_ptrs = _lib._ptr_array([a.ptr for a in V]), _lib._ptr_array([a.ptr for a in v])

        def step():
            m.check(L.pb_multi_pod(m.h, U_ptrs, rows, N_S, lds, EPS, 0, _lib.POD_MODES[mode], C.byref(nl)))
            m.check(L.pb_multi_basis(m.h, V_ptrs, None))
            m.check(L.pb_multi_project(m.h, V_ptrs, None, nl.value, U_ptrs, lds, rows, N_S, v_ptrs))
        for _ in range(3):
            step()
        m.sync()
        m.timer_start()
        for _ in range(steps):
            step()
        ms = m.timer_stop() / steps
        sig = m.sigma()
        line = {"ms_per_step": ms, "n_L": nl.value, "tflops": algo_flops(n_h, N_S, nl.value) / (ms * 1e-3) / 1e12}
        if mode == MODE:
            smax = float(sig_ref[0])
            err = float(np.max(np.abs(sig[:n_L_ref] - sig_ref[:n_L_ref])) / smax)
            line.update({"sigma_err_vs_torchrun": err, "n_L_torchrun": n_L_ref, "ms_per_step_torchrun": ms_torchrun,
                         "status": "ok" if (nl.value == n_L_ref and err <= 1e-10) else "MISMATCH"})
            if line["status"] != "ok":
                raise SystemExit(f"bench.py: single-process multi-GPU result differs from the torchrun run: {json.dumps(line)}")
        out[mode] = line
        del V, v
    del slabs
    m.close()
    return out


def c5_leg(ctx, L, be, sharded, wl, world, rank, barrier, max_over_ranks, sum_over_ranks, peak_tf, pod_single, state):
    """BASELINE.json configs[4] at 1.25 M rows per GPU: U = sum_k s_k phi_k psi_k^T (DCT-II vectors) filled in
    HBM, POD (gram2 and tsqr) + projection.  ASSERTS the closed form (north_star tolerances): n_L, sigma_k = s_k to
    1e-10 sigma_max, largest principal angle between span(V) and the closed-form modes <= 1e-8 (all rows, all ranks).
    V_rows_err_rel (first rows of V against phi_k, per mode: angle / relative gap) is reported, not asserted."""
    import ctypes as C
    from pod_uqnn_b200 import _lib
    n_loc, n_s, r = C5_ROWS_PER_GPU, C5_NS, C5_R
    n_h = n_loc * world
    Uc = ctx.alloc(n_loc, n_s)
    ctx.check(L.pb_dct_lowrank(ctx.h, n_h, n_s, r, C5_DECAY, rank * n_loc, n_loc, Uc.ptr, Uc.ld))
    ctx.sync()
    s = 2. ** (-C5_DECAY * np.arange(r))
    nL_closed = wl.dct_rank(s, EPS)
    out = {"workload": f"C5 DCT known-answer slab(s): n_h = {n_h} ({n_loc} rows per GPU) x n_s = {n_s}, fp64, eps = {EPS:g}, "
                       f"POD + projection, r = {r} exact singular values 2^(-k/4)", "n_L_closed_form": nL_closed}
    n_chk = 4096
    for mode in ("gram2", "tsqr", "tsqr_full"):
        def step():
            if world == 1:
                V, v, nl = pod_single(Uc, mode, "c")
            else:
                V, _, nl = sharded.pod_sharded(be, Uc, EPS, 0, mode)
                v = sharded.project_sharded(be, V, Uc)
            return V, nl
        step(); barrier()
        best = 1e30
        for _ in range(2):
            barrier(); ctx.timer_start(); V, nl = step(); best = min(best, ctx.timer_stop())
        barrier()
        ms = max_over_ranks(best)
        from pod_uqnn_b200.pod import pod_sigma
        sig = pod_sigma(ctx)
        full_qr = world * (2 * n_loc * n_s ** 2 - (2 / 3) * n_s ** 3) + 4 * n_h * n_s * nl
        if mode == "gram2":
            flops = algo_flops(n_h, n_s, nl)
        elif mode == "tsqr_full" or be is None:
            p_, f_, fl_ = C.c_int64(), C.c_int64(), C.c_double()
            ctx.check(L.pb_tsqr_stats(ctx.h, C.byref(p_), C.byref(f_), C.byref(fl_)))
            flops = (full_qr if mode == "tsqr_full" else fl_.value + 4 * n_h * n_s * nl)
        else:
            flops = world * be.qr_stats["local"][2] + be.qr_stats.get("stack", (0, 0, 0.))[2] + 4 * n_h * n_s * nl
        # closed-form checks (every rank checks its own rows of V; rank 0 reports)
        sig_err = float(np.max(np.abs(sig[:nL_closed] - s[:nL_closed])) / s[0]) if len(sig) >= nL_closed else float("inf")
        Vh = np.empty((n_chk, nl))
        ctx.check(L.pb_download(ctx.h, Vh.ctypes.data, V.ptr, n_chk * nl * 8))
        i = (rank * n_loc + np.arange(n_chk) + .5)[:, None]
        phi = np.sqrt(2. / n_h) * np.cos(np.pi * i * (np.arange(nl) + 1.)[None, :] / n_h)
        sg = np.sign(np.sum(Vh * phi, axis=0))
        v_err = max_over_ranks(float(np.max(np.abs(Vh * sg - phi)) / np.sqrt(2. / n_h)))
        # largest principal angle between span(V) and the closed-form modes, from the residual V - Phi (Phi^T V) itself
        Phi, Cd, Gd = ctx.alloc(n_loc, nl), ctx.alloc(nl, nl), ctx.alloc(nl, nl)
        ctx.check(L.pb_dct_basis(ctx.h, n_h, nl, rank * n_loc, n_loc, Phi.ptr, Phi.ld))
        ctx.check(L.pb_gemm_tn(ctx.h, Phi.ptr, Phi.ld, nl, V.ptr, V.ld, nl, n_loc, Cd.ptr))
        Cd.upload(sum_over_ranks(Cd.download()))
        ctx.check(L.pb_residual_gram(ctx.h, V.ptr, V.ld, nl, Phi.ptr, Phi.ld, nl, n_loc, Cd.ptr, Gd.ptr))
        angle = float(np.sqrt(max(0., np.linalg.eigvalsh(sum_over_ranks(Gd.download()))[-1])))
        Phi.release(); Cd.release(); Gd.release()
        ok = (nl == nL_closed) and sig_err <= 1e-10 and angle <= float(os.environ.get("PB_BENCH_ANGLE_TOL", "1e-8"))
        out[mode] = {"wall_s": ms * 1e-3, "n_L": nl, "tflops": flops / (ms * 1e-3) / 1e12,
                     "frac": flops / (ms * 1e-3) / 1e12 / (peak_tf * world),
                     "sigma_err_rel_sigma_max": sig_err, "subspace_angle": angle, "V_rows_err_rel": v_err,
                     "flop_count": {"gram2": "n_h n_s (n_s+1) + 4 n_h n_s n_L",
                                    "tsqr": "Householder flops of the panels actually run (rank-revealing early stop, pb_tsqr_stats) + 4 n_h n_s n_L",
                                    "tsqr_full": "Householder N (2 n_loc n_s^2 - 2/3 n_s^3) + 4 n_h n_s n_L"}[mode],
                     "full_householder_equiv_tflops": None if mode == "gram2" else full_qr / (ms * 1e-3) / 1e12,
                     "stage_ms": {"gram_or_qr": ctx.ms(_lib.T_GRAM if mode == "gram2" else _lib.T_QR), "eig": ctx.ms(_lib.T_EIG),
                                  "basis": ctx.ms(_lib.T_BASIS), "project": ctx.ms(_lib.T_PROJECT)},
                     "status": "ok" if ok else "MISMATCH"}
        if not ok:
            raise SystemExit(f"bench.py: C5 {mode} does not match the closed form: {json.dumps(out[mode])}")
        for k in [k for k in state if k.startswith("c")]:
            state.pop(k)
        if be is not None:
            be._cache.clear()
    out["parity"] = "ok"
    out["peak_tflops_per_gpu"] = peak_tf
    Uc.release()
    return out


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
        out["cpu_baseline"] = {"samples_per_s": n_cpu / sec, "cores": HOST_CORES, "kind": "port",
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
            line["cpu_baseline"] = {"member_epochs_per_s": n_ep / sec, "cores": HOST_CORES, "kind": "port",
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
    ap.add_argument("--skip-c5", action="store_true", help="skip the C5 (1.25 M x 4096 per GPU) leg")
    ap.add_argument("--skip-lib-multi", action="store_true", help="skip the single-process multi-GPU leg (N > 1)")
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
