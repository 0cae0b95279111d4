import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pod_uqnn_b200 import _lib, workloads as wl
from pod_uqnn_b200.acceleration import snapshots_device
from pod_uqnn_b200.pod import pod_device
case, mode = sys.argv[1], sys.argv[2]
ctx = _lib.Context(0)
w = {"c1": lambda: wl.shekel(400, 1000), "c2": lambda: wl.ackley(400, 400, 1000)}[case]()
Ud, _, _ = snapshots_device(ctx, w["field"], w["x_mesh"][:, 1:].T, w["mu"])
for _ in range(2):
    V, sig = pod_device(ctx, Ud, 1e-10, 0, mode)
print(case, mode, "n_L", V.shape[1], "total ms", ctx.ms(_lib.T_POD_TOTAL), "eig", ctx.ms(_lib.T_EIG), "pass2", ctx.ms(_lib.T_PASS2))
