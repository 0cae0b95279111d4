"""Development aid (GPU box): host-side ceiling of the staged upload and the Python-API end-to-end time it gives.
  python tools/probe_hostcopy.py            -> sweep + one API timing per thread setting (subprocesses)
  python tools/probe_hostcopy.py api        -> the API timing alone, with the current environment"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pod_uqnn_b200 import _lib, pod as pod_mod    # noqa: E402


def api_timing():
    ctx = _lib.Context(0)
    rng = np.random.default_rng(0)
    n_h, n_s = 160000, 1000
    U = (rng.standard_normal((n_h, 24)) @ rng.standard_normal((24, n_s)))      # pageable, low rank like C2
    out = {"threads": os.environ.get("PB_HOST_COPY_THREADS", "default"), "nt": os.environ.get("PB_HOST_COPY_NT", "default")}
    for mode in ("gram2", "tsqr"):
        pod_mod.perform_pod(U, 1e-10, 0, False, mode=mode, ctx=ctx)
        tp, tj, tr = [], [], []
        for _ in range(3):
            t0 = time.perf_counter(); V = pod_mod.perform_pod(U, 1e-10, 0, False, mode=mode, ctx=ctx)
            t1 = time.perf_counter(); v2 = pod_mod.project_to_v(V, U, ctx=ctx, reuse_slab=True)
            t2 = time.perf_counter(); v = pod_mod.project_to_v(V, U, ctx=ctx)
            t3 = time.perf_counter()
            tp.append(t1 - t0); tr.append(t2 - t1); tj.append(t3 - t2)
        out[mode] = {"perform_pod_ms": min(tp) * 1e3, "project_to_v_ms": min(tj) * 1e3, "project_reuse_ms": min(tr) * 1e3,
                     "n_L": V.shape[1]}
    Ud = ctx.to_device(U)
    t0 = time.perf_counter(); Ud.upload(U); t1 = time.perf_counter(); Uh = Ud.download(); t2 = time.perf_counter()
    out["upload_1p28GB_ms"] = (t1 - t0) * 1e3; out["download_1p28GB_ms"] = (t2 - t1) * 1e3
    assert np.array_equal(Uh, U)
    print(json.dumps(out), flush=True)


if len(sys.argv) > 1 and sys.argv[1] == "api":
    api_timing()
    sys.exit(0)
L = _lib.lib()
print("host cores", len(os.sched_getaffinity(0)), flush=True)
for T in (1, 2, 4, 8, 12, 16):
    for nt in (0, 1):
        g = C.c_double()
        L.pb_probe_host_copy(1 << 30, T, nt, C.byref(g))
        print(f"host copy pageable -> pinned: threads {T:2d} streaming {nt}: {g.value:6.1f} GB/s", flush=True)
for T, nt in (("4", "0"), ("8", "0"), ("16", "0"), ("8", "1")):
    env = dict(os.environ, PB_HOST_COPY_THREADS=T, PB_HOST_COPY_NT=nt)
    subprocess.run([sys.executable, os.path.abspath(__file__), "api"], env=env, check=False)
