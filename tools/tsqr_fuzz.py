"""Random-shape sweep of the TSQR building blocks on the GPU (development aid): pb_tsqr_r against LAPACK for shapes
around the tiling boundaries (panel widths 16/32/48/64, row counts around multiples of 8 / 128 / the CTA row ranges,
rows < columns, leading dimensions larger than the column count) and the whole mode against the SVD."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import compare as cmp, pod_oracle as po      # noqa: E402
from pod_uqnn_b200 import _lib                           # noqa: E402
from pod_uqnn_b200.pod import pod_device                 # noqa: E402

ctx = _lib.Context(0); L = _lib.lib()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
shapes = [(r, m) for m in (1, 7, 15, 16, 17, 31, 33, 47, 48, 49, 63, 64, 65, 80, 100, 129, 200) for r in (m, m + 1, 2 * m + 3, 127, 128, 129, 1000, 18945)]
shapes += [(int(rng.integers(1, 30000)), int(rng.integers(1, 300))) for _ in range(40)]
shapes += [(int(rng.integers(150000, 260000)), int(rng.integers(8, 40))) for _ in range(4)]          # staged / unstaged boundary
for rows, m in shapes:
    A = rng.standard_normal((rows, m))
    ld = m + int(rng.integers(0, 5))
    big = np.zeros((rows, ld)); big[:, :m] = A
    Ad, Rd = ctx.to_device(big), ctx.alloc(m, m)
    ctx.check(L.pb_tsqr_r(ctx.h, Ad.ptr, rows, m, ld, Rd.ptr))
    R = Rd.download()
    G = A.T @ A
    e = np.abs(R.T @ R - G).max() / np.abs(G).max()
    low = np.abs(np.tril(R, -1)).max()
    if not (e <= 1e-13 and low == 0. and np.isfinite(R).all()):
        bad += 1; print("FAIL R", rows, m, ld, e, low, flush=True)
    Ad.release(); Rd.release()
for _ in range(25):
    n_h, n_s = int(rng.integers(2, 3000)), int(rng.integers(2, 260))
    k = int(rng.integers(1, min(n_h, n_s) + 1))
    U = rng.standard_normal((n_h, k)) * (10. ** -rng.uniform(0, 8, k)) @ rng.standard_normal((k, n_s))
    eps = float(10. ** -rng.uniform(3, 11))
    Vd, sig = pod_device(ctx, ctx.to_device(U), eps, 0, "tsqr")
    Vr = po.perform_pod(U, eps)
    D = np.linalg.svd(U, compute_uv=False)
    ok = Vd.shape == Vr.shape and cmp.sigma_error(sig, D) <= 1e-10
    if ok and Vr.shape[1] > 0:
        ok = cmp.largest_principal_angle(Vd.download(), Vr) <= 1e-8
    if not ok:
        # a rank that differs by one when sigma sits on the threshold is not an error of the kernels: report the margin
        lam = D ** 2; c = np.cumsum(lam) / lam.sum()
        print("CHECK POD", n_h, n_s, k, eps, Vd.shape, Vr.shape, "margin", np.min(np.abs(c - (1 - eps))), flush=True)
        bad += 1
print("fuzz done:", len(shapes), "R shapes, 25 PODs,", bad, "flagged")
