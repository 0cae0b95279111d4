"""Stress check of the ensemble predict kernel: R launches on N points against the oracle (evaluated once)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pod_uqnn_b200 import _lib
from pod_uqnn_b200.varneuralnetwork import ensemble_predict_device
from oracle import ensemble as ens
ctx = _lib.Context(0)
layers, M = [1, 128, 128, 128, 64], 8
N, R = int(sys.argv[1]), int(sys.argv[2])
members = [ens.init_weights(layers, 4000 + i, bias_scale=0.1) for i in range(M)]
X = np.linspace(800, 1200, N)[:, None]
nb = ens.set_normalize_bounds(X, "meanstd")
rm = np.empty((N, layers[-1])); rs = np.empty((N, layers[-1]))
for lo in range(0, N, 50000):
    rm[lo:lo + 50000], rs[lo:lo + 50000] = ens.predict_v(X[lo:lo + 50000], members, layers[-1], 0.01, "meanstd", nb)
Xd = ctx.to_device(X)
tot_bad = 0
for r in range(R):
    vm, vs = ensemble_predict_device(ctx, members, layers, Xd, 0.01, "meanstd", nb)
    vm = vm.download(); vs = vs.download()
    err = np.maximum(np.abs(vm - rm).max(1) / np.abs(rm).max(), np.abs(vs - rs).max(1) / np.abs(rs).max())
    bad = int((err > 1e-6).sum()); tot_bad += bad
    print("run", r, "bad points", bad, "max err", err.max(), "ms", ctx.ms(_lib.T_PREDICT), flush=True)
print("TOTAL bad", tot_bad)
