"""GPU parity tests: the CUDA path (through the C-ABI / the reference-named shims) against the
oracle and the golden vectors of the unmodified reference.  Tolerances are BASELINE.json's:
sigma to 1e-10 * sigma_max, subspace angle <= 1e-8, identical rank, ensemble 1e-6 relative."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden, golden_workload
from oracle import pod_oracle as po, compare as cmp, ensemble as ens
from pod_uqnn_b200 import _lib, workloads as wl
from pod_uqnn_b200.acceleration import snapshots_device, loop_u, loop_u_t
from pod_uqnn_b200.pod import perform_pod, perform_fast_pod, pod_device
from pod_uqnn_b200.varneuralnetwork import VarNeuralNetwork, ensemble_predict_device

pytestmark = pytest.mark.gpu
FIELD = {"shekel": po.u_shekel, "ackley": po.u_ackley, "burgers": po.u_burgers}
TOL_FIELD = 1e-13          # relative to max|U|: exp / cos are <= 2 ulp on both sides


def regen(g, field):
    x_mesh, mu, n_t, t_min, t_max = golden_workload(g)
    return po.create_snapshots(x_mesh, 1, n_t, FIELD[field], mu, t_min, t_max)


# ------------------------------------------------------------------ snapshots
@pytest.mark.parametrize("field", ["shekel", "ackley", "burgers"])
def test_snapshot_kernels_match_reference(ctx, field):
    g = load_golden(f"snap_{field}")
    x_mesh, mu, n_t, t_min, t_max = golden_workload(g)
    Ud, X_v, _ = snapshots_device(ctx, field, x_mesh[:, 1:].T, mu, n_t, t_min, t_max)
    assert np.array_equal(X_v, g["X_v"])
    assert cmp.rel_err(Ud.download(), g["U"]) <= TOL_FIELD


def test_loop_u_reference_signature(ctx):
    g = load_golden("snap_ackley")
    x_mesh, mu, *_ = golden_workload(g)
    n_h, n_s = x_mesh.shape[0], mu.shape[0]
    X_v, U, U_nn = np.zeros((n_s, 3)), np.zeros((n_h, n_s)), np.zeros((n_h, n_s))
    out = loop_u("ackley", n_h, X_v, U, U_nn, x_mesh[:, 1:].T, mu, ctx=ctx)
    assert out[0] is X_v and out[1] is U and out[2] is U and out[3] is U_nn    # fills the caller's arrays
    assert cmp.rel_err(U, g["U"]) <= TOL_FIELD and np.array_equal(U_nn, U)
    with pytest.raises(NotImplementedError):
        loop_u("ackley", n_h, X_v, U, U_nn, x_mesh[:, 1:].T, mu, u_noise=0.1, ctx=ctx)
    with pytest.raises(NotImplementedError):
        loop_u(lambda X, t, m: X, n_h, X_v, U, U_nn, x_mesh[:, 1:].T, mu, ctx=ctx)


def test_loop_u_t_reference_signature(ctx):
    g = load_golden("snap_burgers")
    x_mesh, mu, n_t, t_min, t_max = golden_workload(g)
    n_h, n_s = x_mesh.shape[0], mu.shape[0]
    X_v, U = np.zeros((n_s * n_t, 2)), np.zeros((n_h, n_s * n_t))
    U_nn, Us = np.zeros_like(U), np.zeros((n_h, n_t, n_s))
    loop_u_t("burgers", n_t, 1, n_h, n_h, X_v, U, U_nn, Us, x_mesh[:, 1:].T, mu, t_min, t_max, ctx=ctx)
    assert np.array_equal(X_v, g["X_v"])
    assert cmp.rel_err(U, g["U"]) <= TOL_FIELD
    assert cmp.rel_err(Us, g["U_struct"]) <= TOL_FIELD
    assert U[0].max() == 0. and np.isfinite(U).all()       # x = 0 row; exp overflow -> inf -> 0 handled


def test_snapshot_row_sharding(ctx):
    w = wl.ackley(20, 16, 40)
    X = w["x_mesh"][:, 1:].T
    full, _, _ = snapshots_device(ctx, "ackley", X, w["mu"])
    part, _, _ = snapshots_device(ctx, "ackley", X, w["mu"], row0=100, n_rows=57)
    assert np.array_equal(part.download(), full.download()[100:157])


@pytest.mark.gpu
@pytest.mark.parametrize("n_s", [133, 70])          # odd: one column per thread, even: two (16-byte accesses)
def test_ackley_kernel_variants_agree(ctx, n_s):
    """Radius-grouped (sorted nodes, one exp per run of equal radii) == coordinate-indexed, bit for bit; both within
    a few ulp of the plain per-element kernel; ragged node ranges and a non-symmetric mesh included."""
    w = wl.ackley(50, 36, n_s)
    X = w["x_mesh"][:, 1:].T
    plain = snapshots_device(ctx, "ackley", X, w["mu"], ackley_path="plain")[0].download()
    for row0, n_rows in ((0, None), (70, 1), (123, 1001)):
        a = snapshots_device(ctx, "ackley", X, w["mu"], row0=row0, n_rows=n_rows, ackley_path="indexed")[0].download()
        b = snapshots_device(ctx, "ackley", X, w["mu"], row0=row0, n_rows=n_rows, ackley_path="grouped")[0].download()
        assert np.array_equal(a, b)
        hi = X.shape[1] if n_rows is None else row0 + n_rows
        assert np.abs(a - plain[row0:hi]).max() <= 1e-14 * np.abs(plain).max()
    Xs = X.copy(); Xs[0] += .37                    # shifted: few repeated radii, the default picks the indexed kernel
    c = snapshots_device(ctx, "ackley", Xs, w["mu"])[0].download()
    d = snapshots_device(ctx, "ackley", Xs, w["mu"], ackley_path="grouped")[0].download()
    assert np.array_equal(c, d)


# ------------------------------------------------------------------ products
@pytest.mark.parametrize("rows,ka,kb,sym", [(1000, 200, 200, True), (5000, 1000, 1000, True), (777, 71, 71, True),
                                            (3001, 14, 300, False), (4096, 71, 333, False), (40, 129, 129, True),
                                            (9000, 33, 1001, False), (1, 5, 5, True), (100000, 256, 256, True),
                                            (20000, 1000, 1000, True), (30000, 136, 136, True), (9000, 648, 648, True),
                                            (4000, 1920, 1920, True), (60000, 8, 8, True), (25000, 376, 376, True)])
def test_gemm_tn(ctx, rows, ka, kb, sym):
    rng = np.random.default_rng(rows + ka)
    A = rng.standard_normal((rows, ka))
    B = A if sym else rng.standard_normal((rows, kb))
    L, Ad = _lib.lib(), ctx.to_device(A)
    Cd = ctx.alloc(ka, kb)
    if sym:
        ctx.check(L.pb_gram(ctx.h, Ad.ptr, rows, ka, Ad.ld, Cd.ptr))
    else:
        Bd = ctx.to_device(B)
        ctx.check(L.pb_gemm_tn(ctx.h, Ad.ptr, Ad.ld, ka, Bd.ptr, Bd.ld, kb, rows, Cd.ptr))
    Cg, Cr = Cd.download(), A.T @ B
    assert cmp.rel_err(Cg, Cr) < 1e-13
    if sym:
        assert np.array_equal(Cg, Cg.T)                    # exactly symmetric
        Cd2 = ctx.alloc(ka, kb)
        ctx.check(L.pb_gram(ctx.h, Ad.ptr, rows, ka, Ad.ld, Cd2.ptr))
        assert np.array_equal(Cd2.download(), Cg)          # deterministic (fixed-order split-K)


@pytest.mark.parametrize("rows,n", [(3000, 512), (20011, 576), (70000, 1000), (9000, 4096), (1500, 328)])
def test_gemm_tn_qr_trailing_product(ctx, rows, n):
    """Y^T [Y | C] with Y = the first 64 columns of the same matrix: the TSQR trailing product (TMA kernel qr_tn.cu for
    n >= 512, tile rows split over row ranges; narrower shapes take the cp.async kernels) against numpy."""
    rng = np.random.default_rng(rows + n)
    W = rng.standard_normal((rows, n))
    Wd, Zd = ctx.to_device(W), ctx.alloc(64, n)
    L = _lib.lib()
    ctx.check(L.pb_gemm_tn(ctx.h, Wd.ptr, Wd.ld, 64, Wd.ptr, Wd.ld, n, rows, Zd.ptr))
    Z = Zd.download()
    assert cmp.rel_err(Z, W[:, :64].T @ W) < 1e-13
    ctx.check(L.pb_gemm_tn(ctx.h, Wd.ptr, Wd.ld, 64, Wd.ptr, Wd.ld, n, rows, Zd.ptr))
    assert np.array_equal(Zd.download(), Z)                # deterministic (fixed-order split sum)


@pytest.mark.parametrize("rows,m,k", [(1000, 200, 14), (5000, 1000, 71), (777, 333, 333), (130, 14, 1000),
                                      (300, 71, 2), (999, 15, 33), (1, 3, 1), (4001, 512, 100), (3000, 260, 111),
                                      (2500, 96, 118), (2000, 300, 128)])
def test_gemm_nn(ctx, rows, m, k):
    rng = np.random.default_rng(rows + m + k)
    A, B = rng.standard_normal((rows, m)), rng.standard_normal((m, k))
    Ad, Bd, Yd = ctx.to_device(A), ctx.to_device(B), ctx.alloc(rows, k)
    ctx.check(_lib.lib().pb_gemm_nn(ctx.h, Ad.ptr, Ad.ld, rows, m, Bd.ptr, Bd.ld, k, Yd.ptr, Yd.ld))
    assert cmp.rel_err(Yd.download(), A @ B) < 1e-13


# ------------------------------------------------------------------ POD
POD_CASES = [("shekel_c1", "shekel"), ("shekel_tall", "shekel"), ("ackley_64", "ackley"), ("burgers_1step", "burgers")]


@pytest.mark.parametrize("name,field", POD_CASES)
@pytest.mark.parametrize("mode", ["gram", "gram2", "accurate", "tsqr"])
def test_perform_pod_matches_reference(ctx, name, field, mode):
    g = load_golden(f"pod_{name}")
    _, U, _, _ = regen(g, field)
    V, sig = perform_pod(U, float(g["eps"]), 0, False, mode=mode, ctx=ctx, return_sigma=True)
    n_L = g["V"].shape[1]
    assert V.shape == g["V"].shape, "truncation rank differs from the reference"
    assert V.flags["C_CONTIGUOUS"] and V.dtype == np.float64
    assert cmp.sigma_error(sig[:n_L], g["sigma"][:n_L]) <= cmp.TOL_SIGMA_REL
    angle = cmp.largest_principal_angle(V, g["V"])
    if mode in ("accurate", "tsqr"):
        assert angle <= cmp.TOL_ANGLE
        assert cmp.sigma_error(sig, g["sigma"]) <= cmp.TOL_SIGMA_REL       # every singular value
        v = po.project_to_v(V, U)
        Va, s = cmp.align_signs(V, g["V"])
        assert cmp.rel_err(v * s, g["v"]) <= 1e-8                          # coefficients up to sign
    elif mode == "gram2":
        # tolerance-safe since round 2: the library adds one subspace-refinement step when its a-priori estimate of
        # the Gram-path angle exceeds 1e-8 (csrc/api.cu); unrefined it measured 4.7e-7 on shekel_c1, 1.2e-8 on burgers_1step
        assert angle <= cmp.TOL_ANGLE
    else:
        # one Gram pass: bounded by eps_mach * lambda_1 / gap, reported not required
        assert angle <= 1e-4


def test_perform_pod_fixed_rank_and_full_rank(ctx):
    rng = np.random.default_rng(11)
    U = rng.standard_normal((300, 24)) @ np.diag(2. ** -np.arange(24)) @ rng.standard_normal((24, 40))
    V = perform_pod(U, 0., 7, False, ctx=ctx)                   # n_L given: used verbatim (pod.py:25)
    assert V.shape == (300, 7)
    assert cmp.largest_principal_angle(V, po.perform_pod(U, 0., 7)) <= cmp.TOL_ANGLE
    Uf = rng.standard_normal((64, 12))
    Vf = perform_pod(Uf, 0., 0, False, ctx=ctx)                 # eps = 0 (the default): every mode
    assert Vf.shape == (64, 12)
    assert np.abs(Vf.T @ Vf - np.eye(12)).max() < 1e-10
    with pytest.raises(ValueError):
        perform_pod(U, 0., 39, False, mode="gram", ctx=ctx)     # more modes than the numerical rank
    with pytest.raises(KeyError):
        perform_pod(U, 1e-4, 0, False, mode="householder", ctx=ctx)
    Vt = perform_pod(U, 0., 7, False, mode="tsqr", ctx=ctx)
    assert Vt.shape == (300, 7) and cmp.largest_principal_angle(Vt, po.perform_pod(U, 0., 7)) <= cmp.TOL_ANGLE
    Vft = perform_pod(Uf, 0., 0, False, mode="tsqr", ctx=ctx)   # eps = 0: no truncation of the pivoted QR
    assert Vft.shape == (64, 12) and np.abs(Vft.T @ Vft - np.eye(12)).max() < 1e-10


@pytest.mark.parametrize("mode", ["accurate", "tsqr"])
def test_pod_ragged_shapes(ctx, mode):
    rng = np.random.default_rng(12)
    for (n_h, n_st) in ((257, 33), (33, 257), (100, 1), (1, 9), (129, 129)):
        U = rng.standard_normal((n_h, 3)) @ rng.standard_normal((3, n_st)) if min(n_h, n_st) > 3 \
            else rng.standard_normal((n_h, n_st))
        V, sig = perform_pod(U, 1e-6, 0, False, mode=mode, ctx=ctx, return_sigma=True)
        Vr = po.perform_pod(U, 1e-6)
        assert V.shape == Vr.shape, (n_h, n_st)
        assert cmp.largest_principal_angle(V, Vr) <= cmp.TOL_ANGLE
        D = np.linalg.svd(U, compute_uv=False)
        assert cmp.sigma_error(sig[:V.shape[1]], D[:V.shape[1]]) <= cmp.TOL_SIGMA_REL


# ------------------------------------------------------------------ TSQR mode building blocks
@pytest.mark.parametrize("rows,m", [(700, 130), (64, 64), (100, 1), (33, 70), (5000, 333), (1500, 1000), (4097, 65),
                                    (450000, 24)])
def test_tsqr_r_factor(ctx, rows, m, monkeypatch):
    """pb_tsqr_r: R is upper triangular, R^T R = A^T A to rounding, and equals LAPACK's R up to row signs.
    (450 000 rows with the leaf split forced: three row blocks + the QR of their stacked factors.)"""
    if rows > 400000:
        monkeypatch.setenv("PB_QR_LEAVES", "2")       # force the leaf split (by default only above 16 GiB of working copy)
    rng = np.random.default_rng(rows + m)
    A = rng.standard_normal((rows, m))
    Ad, Rd = ctx.to_device(A), ctx.alloc(m, m)
    ctx.check(_lib.lib().pb_tsqr_r(ctx.h, Ad.ptr, rows, m, Ad.ld, Rd.ptr))
    R = Rd.download()
    assert np.array_equal(Ad.download(), A), "pb_tsqr_r must not modify its input"
    assert np.abs(np.tril(R, -1)).max() == 0.
    G = A.T @ A
    assert np.abs(R.T @ R - G).max() <= 1e-13 * np.abs(G).max()
    Rl = np.linalg.qr(A, mode="r")
    k = Rl.shape[0]
    s = np.sign(np.diag(R)[:k]) * np.sign(np.diag(Rl)); s[s == 0] = 1.
    assert np.abs(R[:k] * s[:, None] - Rl).max() <= 1e-12 * np.abs(Rl).max()
    if k < m:
        assert np.abs(R[k:]).max() == 0.


def test_tsqr_r_ill_conditioned_and_strided(ctx):
    """Graded columns (kappa 1e15), a zero column, a leading dimension larger than the column count."""
    rng = np.random.default_rng(5)
    A = rng.standard_normal((900, 96)) * (10. ** -np.linspace(0, 15, 96))
    A[:, 17] = 0.
    big = np.zeros((900, 128)); big[:, :96] = A
    Bd, Rd = ctx.to_device(big), ctx.alloc(96, 96)
    ctx.check(_lib.lib().pb_tsqr_r(ctx.h, Bd.ptr, 900, 96, 128, Rd.ptr))
    R = Rd.download()
    # column-wise backward error: |A - Q R| <= c eps |A_j|  <=>  the Gram entries agree relative to the column norms
    cn = np.linalg.norm(A, axis=0); cn[cn == 0] = 1.
    E = (R.T @ R - A.T @ A) / np.outer(cn, cn)
    assert np.abs(E).max() <= 1e-13
    assert np.abs(R[:, 17]).max() == 0.


def _tsqr_factor(ctx, A, ld=None):
    rows, m = A.shape
    Ad, Fd = ctx.to_device(A), ctx.alloc(m, m)
    fr = C.c_int64()
    ctx.check(_lib.lib().pb_tsqr_factor(ctx.h, Ad.ptr, rows, m, Ad.ld, Fd.ptr, C.byref(fr)))
    panels, frs, fl = C.c_int64(), C.c_int64(), C.c_double()
    ctx.check(_lib.lib().pb_tsqr_stats(ctx.h, C.byref(panels), C.byref(frs), C.byref(fl)))
    assert frs.value == fr.value
    return Fd.download(), fr.value, panels.value


@pytest.mark.parametrize("rows,m,rank", [(3000, 500, 20), (3000, 500, 100), (5000, 333, 333), (700, 130, 130),
                                         (2000, 1000, 64), (100, 1, 1), (33, 70, 33), (4000, 256, 0)])
def test_tsqr_factor_rank_revealing(ctx, rows, m, rank):
    """pb_tsqr_factor: F^T F = A^T A relative to the largest column norm; a numerically low-rank matrix stops after
    ceil(rank / 64) (+1) panels with that many rows instead of m; a full-rank one runs every panel; the rows below the
    factor are zero; a zero matrix gives zero rows."""
    rng = np.random.default_rng(rows + m + rank)
    if rank == 0:
        A = np.zeros((rows, m))
    elif rank < min(rows, m):
        A = rng.standard_normal((rows, rank)) @ rng.standard_normal((rank, m))
    else:
        A = rng.standard_normal((rows, m))
    F, fr, panels = _tsqr_factor(ctx, A)
    G = A.T @ A
    assert np.abs(F.T @ F - G).max() <= 2e-13 * max(np.abs(G).max(), 1e-300)
    assert np.abs(F[fr:]).max() == 0. if fr < m else True
    if rank == 0:
        assert fr == 0
    elif rank < min(rows, m):
        assert fr <= (rank + 63) // 64 * 64 + 64 and panels <= (rank + 63) // 64 + 1, (fr, panels)
    else:
        assert fr == min(m, (min(rows, m) + 63) // 64 * 64) or fr == m


def test_tsqr_factor_adversarial_column_order(ctx):
    """200 copies of one column first, the informative columns last, graded by 1e-12: block pivoting brings the
    informative columns forward, so the factorisation still stops at the numerical rank, and the POD through the
    factor matches the SVD."""
    rng = np.random.default_rng(77)
    B = rng.standard_normal((6000, 40)) @ np.diag(10. ** -np.linspace(0, 12, 40)) @ rng.standard_normal((40, 300))
    A = np.hstack([np.tile(B[:, :1], (1, 200)), B])
    F, fr, panels = _tsqr_factor(ctx, A)
    assert fr <= 128 and panels <= 2, (fr, panels)
    G = A.T @ A
    assert np.abs(F.T @ F - G).max() <= 2e-13 * np.abs(G).max()
    V, sig = perform_pod(A, 1e-12, 0, False, mode="tsqr", ctx=ctx, return_sigma=True)
    Vr = po.perform_pod(A, 1e-12)
    assert V.shape == Vr.shape
    assert cmp.largest_principal_angle(V, Vr) <= cmp.TOL_ANGLE
    D = np.linalg.svd(A, compute_uv=False)
    assert cmp.sigma_error(sig[:len(D)], D) <= cmp.TOL_SIGMA_REL


@pytest.mark.parametrize("name,field", POD_CASES)
def test_tsqr_full_equals_rank_revealing(ctx, name, field):
    """tsqr (early stop) against tsqr_full (every panel): same rank, same singular values, same subspace."""
    g = load_golden(f"pod_{name}")
    _, U, _, _ = regen(g, field)
    V1, s1 = perform_pod(U, float(g["eps"]), 0, False, mode="tsqr", ctx=ctx, return_sigma=True)
    V2, s2 = perform_pod(U, float(g["eps"]), 0, False, mode="tsqr_full", ctx=ctx, return_sigma=True)
    assert V1.shape == V2.shape == g["V"].shape
    assert cmp.sigma_error(s1, s2) <= 1e-13
    assert cmp.largest_principal_angle(V1, V2) <= 1e-9
    assert cmp.largest_principal_angle(V2, g["V"]) <= cmp.TOL_ANGLE


def test_tsqr_logical_shards_equal_unsharded(ctx):
    """TSQR with one leaf per (logical) rank: R of the stacked slab factors gives the POD of the whole matrix."""
    g = load_golden("pod_ackley_64")
    _, U, _, _ = regen(g, "ackley")
    L = _lib.lib()
    n_h, m = U.shape
    bounds = [0, 1500, 1501, 3000, n_h]          # includes a one-row slab (rows < columns)
    slabs = [ctx.to_device(U[a:b]) for a, b in zip(bounds, bounds[1:])]
    stack = np.zeros((len(slabs) * m, m))
    for i, s in enumerate(slabs):
        Rd = ctx.alloc(m, m)
        ctx.check(L.pb_tsqr_r(ctx.h, s.ptr, s.shape[0], m, s.ld, Rd.ptr))
        stack[i * m:(i + 1) * m] = Rd.download()
    Sd, Rd = ctx.to_device(stack), ctx.alloc(m, m)
    ctx.check(L.pb_tsqr_r(ctx.h, Sd.ptr, stack.shape[0], m, m, Rd.ptr))
    nl = C.c_int()
    ctx.check(L.pb_pod_from_r(ctx.h, Rd.ptr, m, 1e-10, 0, C.byref(nl)))
    V = np.zeros((n_h, nl.value))
    for (a, b), s in zip(zip(bounds, bounds[1:]), slabs):
        Vd = ctx.alloc(b - a, nl.value)
        ctx.check(L.pb_pod_basis(ctx.h, s.ptr, b - a, s.ld, None, Vd.ptr, Vd.ld))
        V[a:b] = Vd.download()
    assert V.shape == g["V"].shape
    assert cmp.largest_principal_angle(V, g["V"]) <= cmp.TOL_ANGLE
    from pod_uqnn_b200.pod import pod_sigma
    assert cmp.sigma_error(pod_sigma(ctx), g["sigma"]) <= cmp.TOL_SIGMA_REL
    with pytest.raises(ValueError):              # a Gram matrix is not an input of this mode
        ctx.check(L.pb_pod_from_gram(ctx.h, Rd.ptr, m, 1e-10, 0, 3, C.byref(nl), C.byref(nl)))


def test_tsqr_cuda_backend(ctx):
    from pod_uqnn_b200 import sharded
    g = load_golden("pod_ackley_64")
    _, U, _, _ = regen(g, "ackley")
    be = sharded.CudaBackend(ctx)
    try:
        Ud = ctx.to_device(U)
        V, sig, n_L = sharded.pod_sharded(be, Ud, 1e-10, 0, "tsqr")
        be.stream.synchronize()
        assert n_L == g["V"].shape[1]
        assert cmp.largest_principal_angle(V.download(), g["V"]) <= cmp.TOL_ANGLE
        import torch
        st = torch.from_numpy(np.vstack([U[:2000], U[2000:]])).to(be.dev)       # a torch stack as pb_tsqr_r input
        R = be.tsqr_r(st)
        be.stream.synchronize()
        assert np.abs(R.cpu().numpy().T @ R.cpu().numpy() - U.T @ U).max() <= 1e-13 * np.abs(U.T @ U).max()
    finally:
        be.close()


@pytest.mark.parametrize("eps", ["0.0001", "1e-08"])
def test_perform_fast_pod_tsqr(ctx, eps):
    g = load_golden(f"fastpod_burgers_{eps}")
    _, _, U_struct, _ = regen(g, "burgers")
    V, ranks = perform_fast_pod(U_struct, float(g["eps"]), float(g["eps"]), mode="tsqr", ctx=ctx, return_ranks=True)
    assert np.array_equal(ranks, g["ranks"]) and V.shape == g["V"].shape
    assert cmp.largest_principal_angle(V, g["V"]) <= cmp.TOL_ANGLE


@pytest.mark.parametrize("eps", ["0.0001", "1e-08"])
def test_perform_fast_pod_matches_reference(ctx, eps):
    g = load_golden(f"fastpod_burgers_{eps}")
    _, _, U_struct, _ = regen(g, "burgers")
    V, ranks = perform_fast_pod(U_struct, float(g["eps"]), float(g["eps"]), ctx=ctx, return_ranks=True)
    assert np.array_equal(ranks, g["ranks"]), "per-sample ranks differ from the reference"
    assert V.shape == g["V"].shape
    assert cmp.largest_principal_angle(V, g["V"]) <= cmp.TOL_ANGLE


def test_logical_shards_equal_unsharded(ctx):
    """P row slabs processed one after the other on one device, partial Grams summed in rank
    order (what the NCCL all-reduce does), replicated eigensolve, per-slab basis."""
    g = load_golden("pod_ackley_64")
    _, U, _, _ = regen(g, "ackley")
    L = _lib.lib()
    n_h, m = U.shape
    bounds = [0, 1500, 1501, 3000, n_h]
    slabs = [ctx.to_device(U[a:b]) for a, b in zip(bounds, bounds[1:])]
    G = np.zeros((m, m))
    for s in slabs:
        Gd = ctx.alloc(m, m)
        ctx.check(L.pb_gram(ctx.h, s.ptr, s.shape[0], m, s.ld, Gd.ptr))
        G += Gd.download()
    nl, keep = C.c_int(), C.c_int()
    Gd = ctx.to_device(G)
    ctx.check(L.pb_pod_from_gram(ctx.h, Gd.ptr, m, 1e-10, 0, 1, C.byref(nl), C.byref(keep)))
    Ys, G2 = [], np.zeros((keep.value, keep.value))
    for s in slabs:
        Y, G2d = ctx.alloc(s.shape[0], keep.value), ctx.alloc(keep.value, keep.value)
        ctx.check(L.pb_pod_pass2_gram(ctx.h, s.ptr, s.shape[0], s.ld, Y.ptr, G2d.ptr))
        Ys.append(Y); G2 += G2d.download()
    G2d = ctx.to_device(G2)
    ctx.check(L.pb_pod_pass2_finish(ctx.h, G2d.ptr, 1e-10, 0, C.byref(nl)))
    V = np.zeros((n_h, nl.value))
    for (a, b), s, Y in zip(zip(bounds, bounds[1:]), slabs, Ys):
        Vd = ctx.alloc(b - a, nl.value)
        ctx.check(L.pb_pod_basis(ctx.h, s.ptr, b - a, s.ld, Y.ptr, Vd.ptr, Vd.ld))
        V[a:b] = Vd.download()
    assert V.shape == g["V"].shape
    assert cmp.largest_principal_angle(V, g["V"]) <= cmp.TOL_ANGLE
    V1 = perform_pod(U, 1e-10, 0, False, mode="gram2", ctx=ctx)
    assert cmp.largest_principal_angle(V, V1) <= 1e-10


def test_gram_host_and_cuda_backend(ctx):
    """pb_gram_host (chunked upload overlapped with the chunk Grams; plain upload when too small to chunk) equals
    pb_gram of the resident slab, and sharded.CudaBackend (library and NCCL on one torch stream) reproduces the
    single-GPU POD when there is no process group."""
    from pod_uqnn_b200 import sharded
    L = _lib.lib()
    rng = np.random.default_rng(12)
    for n_h, m in ((40000, 96), (700, 64)):                 # 4 chunks of 10 000 rows; too small: fallback
        U = rng.standard_normal((n_h, m))
        Ud, Gd, G2d, Uh = ctx.to_device(U), ctx.alloc(m, m), ctx.alloc(m, m), ctx.alloc(n_h, m)
        ctx.check(L.pb_gram(ctx.h, Ud.ptr, n_h, m, m, Gd.ptr))
        ctx.check(L.pb_gram_host(ctx.h, U.ctypes.data, n_h, m, m, Uh.ptr, G2d.ptr))
        assert np.array_equal(Uh.download(), U)
        G, G2 = Gd.download(), G2d.download()
        assert np.abs(G - G2).max() <= 1e-13 * np.abs(G).max()
        assert np.abs(G - U.T @ U).max() <= 1e-12 * np.abs(G).max()
    g = load_golden("pod_ackley_64")
    _, U, _, _ = regen(g, "ackley")
    be = sharded.CudaBackend(ctx)
    try:
        Ud = ctx.to_device(U)
        V, sig, n_L = sharded.pod_sharded(be, Ud, 1e-10, 0, "gram2")
        v = sharded.project_sharded(be, V, Ud)
        be.stream.synchronize()
        assert n_L == g["V"].shape[1]
        assert cmp.largest_principal_angle(V.download(), g["V"]) <= cmp.TOL_ANGLE
        assert cmp.rel_err(v.cpu().numpy(), po.project_to_v(V.download(), U)) <= 1e-12
        Uh = ctx.alloc(*U.shape)
        V2, _, n2 = sharded.pod_sharded(be, Uh, 1e-10, 0, "gram2", A_host=np.ascontiguousarray(U).ctypes.data)
        be.stream.synchronize()
        assert n2 == n_L and cmp.largest_principal_angle(V2.download(), V.download()) <= 1e-10
    finally:
        be.close()


# ------------------------------------------------------------------ projection / reconstruction
def test_projection_reconstruction_pod_sig(ctx):
    g = load_golden("pod_ackley_64")
    _, U, _, _ = regen(g, "ackley")
    L = _lib.lib()
    V = g["V"]; n_h, n = U.shape; nl = V.shape[1]
    Ud, Vd = ctx.to_device(U), ctx.to_device(V)
    v = ctx.alloc(n, nl)
    ctx.check(L.pb_project(ctx.h, Vd.ptr, Vd.ld, nl, Ud.ptr, Ud.ld, n_h, n, v.ptr))
    assert cmp.rel_err(v.download(), g["v"]) <= 1e-12
    Up, sig = ctx.alloc(n_h, n), ctx.alloc(n_h)
    ctx.check(L.pb_reconstruct(ctx.h, Vd.ptr, Vd.ld, nl, v.ptr, n, n_h, Ud.ptr, Ud.ld, Up.ptr, Up.ld, sig.ptr))
    assert cmp.rel_err(Up.download(), po.project_to_U(V, g["v"])) <= 1e-12
    assert cmp.rel_err(sig.download(), g["pod_sig"]) <= 1e-9
    sig2 = ctx.alloc(n_h)                                     # pod_sig alone: U_pod never materialised
    ctx.check(L.pb_reconstruct(ctx.h, Vd.ptr, Vd.ld, nl, v.ptr, n, n_h, Ud.ptr, Ud.ld, None, 0, sig2.ptr))
    assert np.array_equal(sig2.download(), sig.download())


@pytest.mark.parametrize("n_h,n,nl", [(1000, 256, 5), (777, 1000, 14), (333, 1001, 32), (129, 513, 37), (64, 300, 64),
                                      (50, 260, 70), (31, 255, 14)])
def test_reconstruction_strip_kernel_shapes(ctx, n_h, n, nl):
    """V v^T and pod_sig = mean_j |U - V v^T| / 2 (podnnmodel.py:251-253) over ragged shapes: the strip kernel
    (n_L <= 64, n >= 256; odd n takes the unaligned staging path), the tiled kernel otherwise."""
    rng = np.random.default_rng(n_h + n + nl)
    V, v, U = rng.standard_normal((n_h, nl)), rng.standard_normal((n, nl)), rng.standard_normal((n_h, n))
    L = _lib.lib()
    Vd, vd, Ud = ctx.to_device(V), ctx.to_device(v), ctx.to_device(U)
    Up, sig, sig2 = ctx.alloc(n_h, n), ctx.alloc(n_h), ctx.alloc(n_h)
    ctx.check(L.pb_reconstruct(ctx.h, Vd.ptr, Vd.ld, nl, vd.ptr, n, n_h, Ud.ptr, Ud.ld, Up.ptr, Up.ld, sig.ptr))
    ref = V @ v.T
    assert np.abs(Up.download() - ref).max() <= 1e-13 * np.abs(ref).max() * nl
    ref_sig = po.pod_sig(U, ref)
    assert cmp.rel_err(sig.download(), ref_sig) <= 1e-12
    ctx.check(L.pb_reconstruct(ctx.h, Vd.ptr, Vd.ld, nl, vd.ptr, n, n_h, Ud.ptr, Ud.ld, None, 0, sig2.ptr))
    assert np.array_equal(sig2.download(), sig.download())
    Up2 = ctx.alloc(n_h, n)                                   # U_pod alone
    ctx.check(L.pb_reconstruct(ctx.h, Vd.ptr, Vd.ld, nl, vd.ptr, n, n_h, None, 0, Up2.ptr, Up2.ld, None))
    assert np.array_equal(Up2.download(), Up.download())


def test_struct_matrix_roundtrip(ctx):
    rng = np.random.default_rng(4)
    Us = rng.standard_normal((37, 9, 5))
    L = _lib.lib()
    Usd, Um = ctx.to_device(Us), ctx.alloc(37, 45)
    ctx.check(L.pb_struct_to_matrix(ctx.h, Usd.ptr, 37, 9, 5, Um.ptr, Um.ld))
    ref = np.concatenate([Us[:, :, k] for k in range(5)], axis=1)        # pod.py:58 slices
    assert np.array_equal(Um.download(), ref)
    back = ctx.alloc(37, 9, 5)
    ctx.check(L.pb_matrix_to_struct(ctx.h, Um.ptr, Um.ld, 37, 9, 5, back.ptr))
    assert np.array_equal(back.download(), Us)


# ------------------------------------------------------------------ ensemble
@pytest.mark.parametrize("layers,M,N,norm", [([1, 128, 128, 128, 64], 8, 1000, "meanstd"), ([3, 40, 40, 14], 5, 300, "center"),
                                             ([11, 128, 128, 128, 71], 3, 257, "none"), ([2, 16, 5], 1, 100, "meanstd"),
                                             ([1, 40, 40, 130], 2, 129, "meanstd"),
                                             # wider than the fused kernel's 128-column tile: the layer-by-layer path
                                             # (1dt_shallowwater's h_layers = [256, 256, 256]; odd widths; one member)
                                             ([2, 256, 256, 256, 20], 3, 1500, "meanstd"), ([3, 140, 301, 7], 2, 333, "center"),
                                             ([1, 129, 3], 1, 64, "none")])
def test_ensemble_predict_v(ctx, layers, M, N, norm):
    members = [ens.init_weights(layers, 4000 + i, bias_scale=0.1) for i in range(M)]
    X = np.random.default_rng(5).uniform(800, 1200, (N, layers[0]))
    nb = ens.set_normalize_bounds(X, norm)
    vm_r, vs_r = ens.predict_v(X, members, layers[-1], 0.01, norm, nb)
    vm, vs = ensemble_predict_device(ctx, members, layers, X, 0.01, norm, nb)
    assert cmp.rel_err(vm.download(), vm_r) <= cmp.TOL_ENSEMBLE_REL
    assert cmp.rel_err(vs.download(), vs_r) <= cmp.TOL_ENSEMBLE_REL


def test_var_neural_network_predict(ctx):
    layers = [2, 32, 32, 6]
    w = ens.init_weights(layers, 77, bias_scale=0.2)
    X = np.random.default_rng(6).standard_normal((50, 2))
    net = VarNeuralNetwork(layers, 0.01, 0.001, soft_0=0.3, norm="meanstd")
    net.set_weights(w)
    net.set_normalize_bounds(X)
    mean, var = net.predict(X, ctx=ctx)
    mean_r, var_r = ens.member_predict(X, w, 6, 0.3, "meanstd", (X.mean(0), X.std(0)))
    assert cmp.rel_err(mean, mean_r) <= cmp.TOL_ENSEMBLE_REL
    assert np.max(np.abs(var - var_r) / var_r) <= cmp.TOL_ENSEMBLE_REL     # own variance, no cancellation
    deep = [1] + [8] * 9 + [2]
    with pytest.raises(ValueError):                                         # more than 8 Dense layers
        ensemble_predict_device(ctx, [ens.init_weights(deep, 1)], deep, X[:, :1])


def test_predict_U_closed_form(ctx):
    rng = np.random.default_rng(8)
    V = np.linalg.qr(rng.standard_normal((500, 9)))[0]
    vm, vs = rng.standard_normal((33, 9)), rng.uniform(0.01, 0.1, (33, 9))
    L = _lib.lib()
    Vd, vmd, vsd = ctx.to_device(V), ctx.to_device(vm), ctx.to_device(vs)
    Um, Us = ctx.alloc(500, 33), ctx.alloc(500, 33)
    ctx.check(L.pb_predict_U(ctx.h, Vd.ptr, Vd.ld, 500, 9, vmd.ptr, vsd.ptr, 33, 0, 0, Um.ptr, Us.ptr, Um.ld))
    Um_r, Us_r = ens.predict_U_moments(V, vm, vs)
    assert cmp.rel_err(Um.download(), Um_r) <= 1e-12 and cmp.rel_err(Us.download(), Us_r) <= 1e-12
    # and the reference's Monte-Carlo estimator converges to it (statistical parity)
    U_mc, Us_mc = ens.predict_mc(V, vm, vs, 4000, np.random.default_rng(9))
    assert np.abs(U_mc - Um_r).max() < 5 * Us_r.max() / np.sqrt(4000) * 1.5
    assert cmp.rel_err(Us_mc, Us_r) < 0.15


# ------------------------------------------------------------------ errors
def test_error_behaviour(ctx):
    L = _lib.lib()
    w = wl.ackley(4, 4, 4)
    X = w["x_mesh"][:, 1:].T
    with pytest.raises(NotImplementedError):
        Xd, mud, Ud = ctx.to_device(X), ctx.to_device(w["mu"]), ctx.alloc(16, 4)
        ctx.check(L.pb_snapshots(ctx.h, 99, 2, 16, Xd.ptr, 4, 3, mud.ptr, 0, None, 0, 16, Ud.ptr, 4))
    with pytest.raises(ValueError):
        snapshots_device(ctx, "ackley", X, w["mu"], n_t=5, t_min=0., t_max=1.)     # steady field with n_t
    with pytest.raises(ValueError):
        snapshots_device(ctx, "shekel", X, w["mu"])                                # wrong dim / n_mu
    with pytest.raises(ValueError):
        pod_device(ctx, ctx.to_device(np.ones((4, 4))), eps=1.5)
    with pytest.raises(ValueError):
        pod_device(ctx, ctx.to_device(np.zeros((8, 4))), eps=1e-4)                  # zero matrix: no pivot
    with pytest.raises(ValueError):
        pod_device(ctx, ctx.to_device(np.zeros((8, 4))), eps=1e-4, mode="gram2")


def test_pod_host_pipelined_upload(ctx):
    """pb_pod_host (chunked upload overlapped with the Gram) == upload + pb_pod, incl. a strided host matrix."""
    rng = np.random.default_rng(21)
    n_h, n_st = 40000, 96
    U = (rng.standard_normal((n_h, 12)) * 2. ** -np.arange(12)) @ rng.standard_normal((12, n_st))
    V1, s1 = perform_pod(U, 1e-8, 0, False, mode="gram2", ctx=ctx, return_sigma=True)          # host path
    V2d, s2 = pod_device(ctx, ctx.to_device(U), 1e-8, 0, "gram2")
    assert V1.shape == V2d.shape
    assert cmp.largest_principal_angle(V1, V2d.download()) <= 1e-10
    assert cmp.sigma_error(s1, s2) <= 1e-12
    Vr = po.perform_pod(U, 1e-8)
    assert V1.shape == Vr.shape and cmp.largest_principal_angle(V1, Vr) <= cmp.TOL_ANGLE
    # leading dimension of the host matrix larger than n_st
    big = np.zeros((n_h, n_st + 5)); big[:, :n_st] = U
    L = _lib.lib()
    Ud, nl = ctx.alloc(n_h, n_st), C.c_int()
    ctx.check(L.pb_pod_host(ctx.h, big.ctypes.data, n_h, n_st, n_st + 5, Ud.ptr, 1e-8, 0, 1, C.byref(nl)))
    assert nl.value == Vr.shape[1] and np.array_equal(Ud.download(), U)


@pytest.mark.gpu
@pytest.mark.parametrize("pinned", [False, True])
def test_pod_host_tsqr_overlapped_upload(ctx, pinned):
    """pb_pod_host in the default TSQR mode factors each row chunk while the next ones cross PCIe (leaf factors,
    then the QR of the stack): same rank / sigma / modes as the device path and as the reference's SVD; the uploaded
    slab is complete."""
    rng = np.random.default_rng(22)
    n_h, n_st = 70000, 200
    U = (rng.standard_normal((n_h, 20)) * 2. ** -np.arange(20)) @ rng.standard_normal((20, n_st))
    L = _lib.lib()
    host = U
    hp = None
    if pinned:
        hp = C.c_void_p(); assert L.pb_host_alloc(U.nbytes, C.byref(hp)) == 0
        host = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_double)), shape=U.shape); host[...] = U
    try:
        Ud, nl = ctx.alloc(n_h, n_st), C.c_int()
        ctx.check(L.pb_pod_host(ctx.h, host.ctypes.data, n_h, n_st, n_st, Ud.ptr, 1e-9, 0, _lib.POD_MODES["tsqr"], C.byref(nl)))
        V1 = ctx.alloc(n_h, nl.value)
        ctx.check(L.pb_pod_basis_full(ctx.h, Ud.ptr, n_h, n_st, Ud.ld, V1.ptr, V1.ld))
        from pod_uqnn_b200.pod import pod_sigma
        s1 = pod_sigma(ctx)
        assert np.array_equal(Ud.download(), U)
        V2d, s2 = pod_device(ctx, Ud, 1e-9, 0, "tsqr")
        Vr = po.perform_pod(U, 1e-9)
        assert nl.value == Vr.shape[1] == V2d.shape[1]
        assert cmp.largest_principal_angle(V1.download(), Vr) <= cmp.TOL_ANGLE
        assert cmp.largest_principal_angle(V1.download(), V2d.download()) <= 1e-10
        assert cmp.sigma_error(s1[:nl.value], s2[:nl.value]) <= 1e-12
    finally:
        if hp is not None:
            L.pb_host_free(hp)


def test_download_forked_overlaps_and_joins_at_sync(ctx):
    """pb_download_forked: D2H on the copy stream, ordered after the producer, joined by pb_sync."""
    rng = np.random.default_rng(44)
    A = rng.standard_normal((200000, 16)); B = rng.standard_normal((16, 16))
    L = _lib.lib()
    Ad, Bd, Yd, Zd = ctx.to_device(A), ctx.to_device(B), ctx.alloc(200000, 16), ctx.alloc(200000, 16)
    hp = C.c_void_p(); assert L.pb_host_alloc(Yd.nbytes, C.byref(hp)) == 0
    try:
        host = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_double)), shape=(200000, 16)); host[...] = 0.
        ctx.check(L.pb_gemm_nn(ctx.h, Ad.ptr, Ad.ld, 200000, 16, Bd.ptr, Bd.ld, 16, Yd.ptr, Yd.ld))      # producer
        ctx.check(L.pb_download_forked(ctx.h, hp, Yd.ptr, Yd.nbytes))
        ctx.check(L.pb_gemm_nn(ctx.h, Yd.ptr, Yd.ld, 200000, 16, Bd.ptr, Bd.ld, 16, Zd.ptr, Zd.ld))      # runs beside the copy
        ctx.sync()
        assert cmp.rel_err(host, A @ B) < 1e-13 and cmp.rel_err(Zd.download(), A @ B @ B) < 1e-13
    finally:
        L.pb_host_free(hp)


def test_refine_rows_override_is_scoped_to_one_decomposition(ctx):
    """pb_pod_set_refine_rows (row-sharded callers: the largest slab of the job) changes the a-priori estimate of the
    decomposition in progress only: every new Gram clears it, so a later single-GPU pb_pod on the same context is not
    decided with another job's row count."""
    rng = np.random.default_rng(33)
    n_h, n_st = 300000, 64
    U = (rng.standard_normal((n_h, 24)) * 2. ** (-1.3 * np.arange(24))) @ rng.standard_normal((24, n_st))
    L, Ud = _lib.lib(), ctx.to_device(U)
    est = C.c_double()

    def estimate_after_pod():
        pod_device(ctx, Ud, 1e-10, 0, "gram2")
        ctx.check(L.pb_pod_refine_estimate(ctx.h, C.byref(est)))
        return est.value

    e_plain = estimate_after_pod()
    ctx.check(L.pb_pod_set_refine_rows(ctx.h, 1000))            # stale override from "another job"
    assert estimate_after_pod() == e_plain                       # pb_pod -> pb_gram cleared it
    # set between the Gram and the eigensolve (what sharded.pod_sharded does): it is used
    G, nl, keep = ctx.alloc(n_st, n_st), C.c_int(), C.c_int()
    ctx.check(L.pb_gram(ctx.h, Ud.ptr, n_h, n_st, Ud.ld, G.ptr))
    ctx.check(L.pb_pod_set_refine_rows(ctx.h, 1000))
    ctx.check(L.pb_pod_from_gram(ctx.h, G.ptr, n_st, 1e-10, 0, _lib.POD_MODES["gram2"], C.byref(nl), C.byref(keep)))
    Y, G2 = ctx.alloc(n_h, keep.value), ctx.alloc(keep.value, keep.value)
    ctx.check(L.pb_pod_pass2_gram(ctx.h, Ud.ptr, n_h, Ud.ld, Y.ptr, G2.ptr))
    ctx.check(L.pb_pod_pass2_finish(ctx.h, G2.ptr, 1e-10, 0, C.byref(nl)))
    ctx.check(L.pb_pod_refine_estimate(ctx.h, C.byref(est)))
    assert 0. < est.value < e_plain                              # 1000 rows instead of 300 000: a smaller estimate


def test_predict_U_monte_carlo(ctx):
    """The reference's MC estimator on the GPU: reproducible per seed, converges to the closed form."""
    rng = np.random.default_rng(18)
    V = np.linalg.qr(rng.standard_normal((700, 11)))[0]
    vm, vs = rng.standard_normal((45, 11)), rng.uniform(0.05, 0.2, (45, 11))
    L = _lib.lib()
    Vd, vmd, vsd = ctx.to_device(V), ctx.to_device(vm), ctx.to_device(vs)

    def run(samples, seed):
        Um, Us = ctx.alloc(700, 45), ctx.alloc(700, 45)
        ctx.check(L.pb_predict_U(ctx.h, Vd.ptr, Vd.ld, 700, 11, vmd.ptr, vsd.ptr, 45, samples, seed, Um.ptr, Us.ptr, Um.ld))
        return Um.download(), Us.download()

    Um_r, Us_r = ens.predict_U_moments(V, vm, vs)
    S = 2000
    a = run(S, 7)
    b = run(S, 7)
    c = run(S, 8)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])          # counter-based: reproducible
    assert not np.array_equal(a[0], c[0])
    # sample mean within 6 standard errors everywhere, sample sigma within 12 % (S = 2000 -> ~1.6 % rms)
    assert np.all(np.abs(a[0] - Um_r) <= 6. * Us_r / np.sqrt(S) + 1e-12)
    assert np.max(np.abs(a[1] / Us_r - 1.)) < 0.12
    # the draws are standard normal: pooled z-score of the mean error has unit variance
    z = (a[0] - Um_r) / (Us_r / np.sqrt(S))
    assert abs(z.mean()) < 0.1 and abs(z.std() - 1.) < 0.1
    with pytest.raises(ValueError):
        run(1, 0)
