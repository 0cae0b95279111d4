import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pod_uqnn_b200 import _lib
from pod_uqnn_b200.varneuralnetwork import ensemble_predict_device
from oracle import ensemble as ens
ctx = _lib.Context(0)
layers, N, M = [1, 128, 128, 128, 64], int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 4, 8
members = [ens.init_weights(layers, 4000 + i) for i in range(M)]
X = np.linspace(800, 1200, N)[:, None]
nb = ens.set_normalize_bounds(X, "meanstd")
Xd = ctx.to_device(X)
for _ in range(3):
    vm, vs = ensemble_predict_device(ctx, members, layers, Xd, 0.01, "meanstd", nb)
print("kernel ms", ctx.ms(_lib.T_PREDICT), "samples/s", N / ctx.ms(_lib.T_PREDICT) * 1e3)
