"""One POD of a DCT known-answer slab (n_h x n_s) in the given mode; for launch lists / ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pod_uqnn_b200 import _lib
from pod_uqnn_b200.pod import pod_device
n_h, n_s, mode = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
ctx = _lib.Context(0); L = _lib.lib()
Ud = ctx.alloc(n_h, n_s)
ctx.check(L.pb_dct_lowrank(ctx.h, n_h, n_s, 256, 0.25, 0, n_h, Ud.ptr, Ud.ld))
for _ in range(reps):
    V, sig = pod_device(ctx, Ud, 1e-10, 0, mode)
print(n_h, n_s, mode, "n_L", V.shape[1], "total ms", ctx.ms(_lib.T_POD_TOTAL), "qr", ctx.ms(_lib.T_QR), "eig", ctx.ms(_lib.T_EIG))
