"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python oracle/gen_golden.py

TEST INFRASTRUCTURE.  Imports /root/reference through oracle/reference_loader.py (numba ->
LAPACK dgesdd; tensorflow / tfp / meshio stubbed), runs its own loop_u / loop_u_t /
perform_pod / perform_fast_pod / PodnnModel.create_snapshots on the seeded workloads of
pod_uqnn_b200/workloads.py and stores inputs + outputs as small fixtures.  The numpy
restatement (oracle/pod_oracle.py) is asserted against the reference here, so a fixture is
only written when restatement and reference agree.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_loader, pod_oracle as po, compare as cmp  # noqa: E402
from pod_uqnn_b200 import workloads as wl  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def ref_snapshots(ref, case, w):
    """The reference's PodnnModel.create_snapshots with its own field function."""
    u, _ = reference_loader.load_field(case)
    with tempfile.TemporaryDirectory() as d:
        model = ref.podnnmodel.PodnnModel(d, 1, w["x_mesh"], w["n_t"])
        n_d = w["mu"].shape[1] + (1 if w["n_t"] > 0 else 0)
        n_h = w["x_mesh"].shape[0]
        return model.create_snapshots(n_d, n_h, u, w["mu"], w["t_min"], w["t_max"])


def main():
    ref = reference_loader.load()
    os.makedirs(GOLD, exist_ok=True)
    perform_pod, perform_fast_pod = ref.pod.perform_pod, ref.pod.perform_fast_pod

    # ---- snapshot loops: reference vs restatement, small cases stored in full
    snap = {"shekel": ("1d_shekel", wl.shekel(100, 48), po.u_shekel),
            "ackley": ("2d_ackley", wl.ackley(20, 16, 40), po.u_ackley),
            "burgers": ("1dt_burger", wl.burgers(64, 20, 10), po.u_burgers)}
    for name, (case, w, u_or) in snap.items():
        X_v, U, U_struct, U_nn = ref_snapshots(ref, case, w)
        X_v2, U2, Us2, _ = po.create_snapshots(w["x_mesh"], 1, w["n_t"], u_or, w["mu"], w["t_min"], w["t_max"])
        err = cmp.rel_err(U2, U)
        assert err < 1e-14 and np.array_equal(X_v, X_v2), (name, err)
        if w["n_t"] > 0:
            assert cmp.rel_err(Us2, U_struct) < 1e-14   # numba exp vs numpy exp: <= 1 ulp
        np.savez_compressed(os.path.join(GOLD, f"snap_{name}.npz"), x_mesh=w["x_mesh"], mu=w["mu"],
                            n_t=w["n_t"], t_min=w["t_min"], t_max=w["t_max"], X_v=X_v, U=U,
                            U_struct=U_struct if w["n_t"] > 0 else np.zeros(0))
        print(f"snap_{name}: restatement vs reference rel err {err:.2e}")

    # ---- perform_pod on the three fields (U regenerated from the stored inputs by the tests)
    pods = {"shekel_c1": ("1d_shekel", wl.shekel(400, 1000), po.u_shekel, 1e-10),           # C1, wide
            "shekel_tall": ("1d_shekel", wl.shekel(400, 300), po.u_shekel, 1e-10),
            "ackley_64": ("2d_ackley", wl.ackley(64, 64, 256), po.u_ackley, 1e-10),
            "burgers_1step": ("1dt_burger", wl.burgers(256, 100, 12), po.u_burgers, 1e-8)}
    for name, (case, w, u_or, eps) in pods.items():
        _, U, _, _ = ref_snapshots(ref, case, w)
        _, U2, _, _ = po.create_snapshots(w["x_mesh"], 1, w["n_t"], u_or, w["mu"], w["t_min"], w["t_max"])
        assert cmp.rel_err(U2, U) < 1e-14
        V = perform_pod(U, eps, 0, False)
        V2 = po.perform_pod(U2, eps)
        D, _ = po.pod_spectrum(U)
        assert V.shape == V2.shape, (name, V.shape, V2.shape)
        ang = cmp.largest_principal_angle(V2, V)
        assert ang < 1e-9, (name, ang)
        v = (V.T.dot(U)).T
        U_pod = V.dot(v.T)
        pod_sig = np.stack((U, U_pod), axis=-1).std(-1).mean(-1)
        np.savez_compressed(os.path.join(GOLD, f"pod_{name}.npz"), x_mesh=w["x_mesh"], mu=w["mu"],
                            n_t=w["n_t"], t_min=w["t_min"], t_max=w["t_max"], eps=eps, sigma=D, V=V,
                            v=v, pod_sig=pod_sig)
        print(f"pod_{name}: U {U.shape} n_L {V.shape[1]} restatement angle {ang:.2e}")

    # ---- two-step POD (perform_fast_pod)
    w = wl.burgers(256, 100, 24)
    _, U, U_struct, _ = ref_snapshots(ref, "1dt_burger", w)
    for eps in (1e-4, 1e-8):
        V = perform_fast_pod(U_struct, eps, eps)
        V2, ranks = po.perform_fast_pod(U_struct, eps, eps, return_ranks=True)
        assert V.shape == V2.shape
        ang = cmp.largest_principal_angle(V2, V)
        assert ang < 1e-9, ang
        np.savez_compressed(os.path.join(GOLD, f"fastpod_burgers_{eps:g}.npz"), x_mesh=w["x_mesh"], mu=w["mu"],
                            n_t=w["n_t"], t_min=w["t_min"], t_max=w["t_max"], eps=eps, ranks=ranks, V=V)
        print(f"fastpod eps={eps:g}: ranks sum {ranks.sum()} n_L {V.shape[1]} restatement angle {ang:.2e}")

    # ---- C2 at full size (160 000 x 1000): spectrum, rank, projection and a row sample of V
    w = wl.ackley(400, 400, 1000)
    _, U, _, _ = ref_snapshots(ref, "2d_ackley", w)
    V = perform_pod(U, 1e-10, 0, False)
    D = np.linalg.svd(U, compute_uv=False)
    v = (V.T.dot(U)).T
    rows = np.arange(0, U.shape[0], 97)
    np.savez_compressed(os.path.join(GOLD, "pod_ackley_c2_full.npz"), eps=1e-10, sigma=D, n_L=V.shape[1],
                        rows=rows, V_rows=V[rows], v=v, U_rows=U[rows][:, ::7])
    print(f"pod_ackley_c2_full: n_L {V.shape[1]} sigma ratios {D[V.shape[1]-1]/D[0]:.3e} {D[V.shape[1]]/D[0]:.3e}")


if __name__ == "__main__":
    main()
