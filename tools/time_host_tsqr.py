"""perform_pod(ndarray) in the default TSQR mode on the C2 matrix: overlapped leaf QRs vs upload-then-factor (run twice with
PB_TSQR_HOST_OVERLAP=1 / 0).  Pageable and pinned sources."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pod_uqnn_b200 import _lib, workloads as wl, pod as pod_mod
from pod_uqnn_b200.acceleration import snapshots_device
ctx = _lib.Context(0); L = _lib.lib()
w = wl.ackley(400, 400, 1000)
U = snapshots_device(ctx, "ackley", w["x_mesh"][:, 1:].T, w["mu"])[0].download()
for mode in ("tsqr", "gram2"):
    pod_mod.perform_pod(U, 1e-10, 0, False, mode=mode, ctx=ctx)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); V = pod_mod.perform_pod(U, 1e-10, 0, False, mode=mode, ctx=ctx); ts.append(time.perf_counter() - t0)
    print(f"overlap={os.environ.get('PB_TSQR_HOST_OVERLAP', '1')} pageable {mode}: perform_pod {min(ts) * 1e3:.2f} ms, n_L {V.shape[1]}, "
          f"pod_total {ctx.ms(_lib.T_POD_TOTAL):.2f} ms", flush=True)
hp = C.c_void_p(); L.pb_host_alloc(U.nbytes, C.byref(hp))
hU = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_double)), shape=U.shape); hU[...] = U
Ud, nl = ctx.alloc(*U.shape), C.c_int()
for mode in ("tsqr", "gram2"):
    ts = []
    for _ in range(4):
        ctx.sync(); t0 = time.perf_counter()
        ctx.check(L.pb_pod_host(ctx.h, hU.ctypes.data, U.shape[0], U.shape[1], U.shape[1], Ud.ptr, 1e-10, 0, _lib.POD_MODES[mode], C.byref(nl)))
        ctx.sync(); ts.append(time.perf_counter() - t0)
    print(f"overlap={os.environ.get('PB_TSQR_HOST_OVERLAP', '1')} pinned {mode}: pb_pod_host {min(ts) * 1e3:.2f} ms, n_L {nl.value}", flush=True)
