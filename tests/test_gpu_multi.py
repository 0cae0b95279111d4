"""Single-process multi-GPU path (pb_multi_*, csrc/multi.cu) against the single-GPU result and the golden vectors of
the unmodified reference.  Logical ranks share cuda:0 on a one-GPU box (every exchange step still runs: peer-memory
reduce kernel, slice copies, event hand-offs between the per-rank host threads); with >= 2 devices the same tests
also run across real GPUs."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden, golden_workload
from oracle import pod_oracle as po, compare as cmp
from pod_uqnn_b200 import _lib
from pod_uqnn_b200.pod import perform_pod, project_to_v

pytestmark = pytest.mark.gpu
FIELD = {"shekel": po.u_shekel, "ackley": po.u_ackley, "burgers": po.u_burgers}


def n_devices():
    import subprocess
    try:
        return len(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.strip().splitlines())
    except OSError:
        return 1


def layouts():
    out = [("logical-2", [0, 0]), ("logical-3", [0, 0, 0])]
    nd = n_devices()
    if nd >= 2:
        out.append(("real-2", [0, 1]))
    if nd >= 4:
        out.append(("real-4", [0, 1, 2, 3]))
    return out


@pytest.fixture(scope="module", params=layouts(), ids=lambda l: l[0])
def multi(request):
    m = _lib.MultiContext(len(request.param[1]), request.param[1])
    yield m
    m.close()


def tall_case(name, field):
    g = load_golden(f"pod_{name}")
    x_mesh, mu, n_t, t_min, t_max = golden_workload(g)
    _, U, _, _ = po.create_snapshots(x_mesh, 1, n_t, FIELD[field], mu, t_min, t_max)
    return g, U


def test_all_reduce_bitwise_replicated(multi):
    """pb_multi_all_reduce: every rank ends with the same bits, equal to the rank-ordered sum; odd counts and
    unaligned buffers take the scalar path."""
    rng = np.random.default_rng(5)
    P = multi.n_gpus
    for count, off in ((1, 0), (7, 0), (1000 * 1000, 0), (4097, 1), (2 * P * 3 + 2, 1)):
        host = [rng.standard_normal(count + off) for _ in range(P)]
        bufs = [multi.ctxs[p].to_device(host[p]) for p in range(P)]
        ptrs = _lib._ptr_array([b.ptr + 8 * off for b in bufs])
        multi.check(_lib.lib().pb_multi_all_reduce(multi.h, ptrs, count))
        multi.sync()
        want = host[0][off:].copy()
        for q in range(1, P):
            want = want + host[q][off:]
        for p in range(P):
            got = bufs[p].download()
            assert np.array_equal(got[off:], want), (count, off, p)
            assert np.array_equal(got[:off], host[p][:off])


@pytest.mark.parametrize("mode", ["gram", "gram2", "accurate", "tsqr", "tsqr_full"])
@pytest.mark.parametrize("name,field", [("ackley_64", "ackley"), ("shekel_tall", "shekel")])
def test_multi_pod_matches_reference(multi, ctx, name, field, mode):
    """Row-sharded slabs in HBM: rank, singular values and subspace against the golden V of the unmodified reference
    (north_star tolerances; the one-pass gram mode is held to the single-GPU run instead)."""
    g, U = tall_case(name, field)
    eps = float(g["eps"])
    P = multi.n_gpus
    rng_rows = [multi.row_range(U.shape[0], p) for p in range(P)]
    slabs = [multi.ctxs[p].to_device(U[lo:hi]) for p, (lo, hi) in enumerate(rng_rows)]
    Vs, sig, n_L = multi.pod(slabs, eps, 0, mode)
    V = np.vstack([Vs[p].download() for p in range(P)])
    V1, sig1 = perform_pod(U, eps, 0, False, mode=mode, ctx=ctx, return_sigma=True)
    assert V.shape == V1.shape
    # one Gram pass resolves small singular values only to eps sigma_1^2 / sigma_k: a different summation order shows there
    assert cmp.sigma_error(sig[:n_L], sig1[:n_L]) <= (cmp.TOL_SIGMA_REL if mode == "gram" else 1e-12)
    if mode != "gram" or field == "ackley":      # (one Gram pass on the ill-conditioned Shekel set: angle ~ eps lambda_1 / gap = 1e-5)
        assert cmp.largest_principal_angle(V, V1) <= (1e-6 if mode == "gram" else 1e-9)
    if mode != "gram":
        assert V.shape == g["V"].shape
        if mode != "gram2" or field == "ackley":
            assert cmp.largest_principal_angle(V, g["V"]) <= cmp.TOL_ANGLE
        D = np.linalg.svd(U, compute_uv=False)
        assert cmp.sigma_error(sig[:n_L], D[:n_L]) <= cmp.TOL_SIGMA_REL
    # projection: sum of the per-rank partials, replicated bit for bit
    out = multi.project(Vs, slabs)
    v0 = out[0].download()
    for p in range(1, P):
        assert np.array_equal(out[p].download(), v0)
    assert cmp.rel_err(v0, po.project_to_v(V, U)) <= 1e-12


@pytest.mark.parametrize("mode", ["gram2", "tsqr"])
def test_multi_host_api(multi, mode, monkeypatch):
    """pod.perform_pod(ndarray, n_gpus=P) + project_to_v through pb_multi_pod_host / basis_host / project_host."""
    rng = np.random.default_rng(11)
    n_h, n_s, r = 4096 * multi.n_gpus + 37, 160, 20
    U = (rng.standard_normal((n_h, r)) * 2. ** -np.arange(r)) @ rng.standard_normal((r, n_s))
    monkeypatch.setitem(_lib._multi, multi.n_gpus, multi)
    V, sig = perform_pod(U, 1e-8, 0, False, mode=mode, n_gpus=multi.n_gpus, return_sigma=True)
    Vr = po.perform_pod(U, 1e-8)
    assert V.shape == Vr.shape
    assert cmp.largest_principal_angle(V, Vr) <= cmp.TOL_ANGLE
    D = np.linalg.svd(U, compute_uv=False)
    assert cmp.sigma_error(sig[:len(D)], D) <= cmp.TOL_SIGMA_REL
    v_ref = po.project_to_v(V, U)
    v1 = project_to_v(V, U, reuse_slab=True, n_gpus=multi.n_gpus)
    v2 = project_to_v(V, U, n_gpus=multi.n_gpus)
    assert cmp.rel_err(v1, v_ref) <= 1e-12 and cmp.rel_err(v2, v_ref) <= 1e-12
    with pytest.raises(ValueError):
        project_to_v(V, U, reuse_slab=True, n_gpus=multi.n_gpus)          # the held matrix was released by the call above


def test_multi_errors(multi):
    """wide matrices are refused with a message; an empty rank (fewer rows than ranks would need) still works."""
    P = multi.n_gpus
    slabs = [multi.ctxs[p].to_device(np.ones((2, 64))) for p in range(P)]
    with pytest.raises(ValueError, match="n_h >= n_st"):
        multi.pod(slabs, 1e-8, 0, "gram2")
    rng = np.random.default_rng(3)
    U = rng.standard_normal((500, 8)) @ rng.standard_normal((8, 40))
    rows = [500] + [0] * (P - 1)
    slabs = [multi.ctxs[0].to_device(U)] + [multi.ctxs[p].alloc(1, 40) for p in range(1, P)]
    for s, r in zip(slabs, rows):
        s.shape = (r, 40)
    Vs, sig, n_L = multi.pod(slabs, 1e-10, 0, "tsqr")
    assert n_L == 8
    assert cmp.largest_principal_angle(Vs[0].download(), po.perform_pod(U, 1e-10)) <= cmp.TOL_ANGLE
