"""Seeded synthetic workloads C1..C5 (BASELINE.json:configs).

Host-side harness data only: meshes, parameter samples and field parameters
that tests, bench.py and the golden-fixture script feed to BOTH the oracle and
the CUDA path.  Nothing here computes on the hot path.
"""
import numpy as np


def create_linear_mesh(x_min, x_max, n_x, y_min=0, y_max=0, n_y=0):
    """Tensor-grid mesh with a 1-based id in column 0 (reference poduqnn/mesh.py:11-42).

    2-D node order is iy * n_x + ix (np.meshgrid 'xy', mesh.py:36-39).
    """
    x = np.linspace(x_min, x_max, n_x)
    if n_y > 0:
        y = np.linspace(y_min, y_max, n_y)
        X, Y = np.meshgrid(x, y)
        n_xyz = n_x * n_y
        return np.column_stack((np.arange(1, n_xyz + 1, dtype=np.float64),
                                X.reshape(n_xyz), Y.reshape(n_xyz)))
    return np.column_stack((np.arange(1, n_x + 1, dtype=np.float64), x))


# Shekel parameter box, experiments/1d_shekel/hyperparams.py:43-47
_SHEKEL_MU_MEAN = np.hstack((1/10 * np.array([1, 2, 2, 4, 4]),
                             1. * np.array([4, 1, 8, 6, 3])))
SHEKEL_MU_MIN = _SHEKEL_MU_MEAN * (1 - np.sqrt(3)/10)
SHEKEL_MU_MAX = _SHEKEL_MU_MEAN * (1 + np.sqrt(3)/10)


def _uniform(seed, n_s, mu_min, mu_max):
    mu_min, mu_max = np.asarray(mu_min, float), np.asarray(mu_max, float)
    R = np.random.default_rng(seed).random((n_s, mu_min.shape[0]))
    return mu_min + (mu_max - mu_min) * R


def shekel(n_x=400, n_s=1000, seed=1111):
    """C1: Shekel u(x; mu), x in [0, 10], d = 10 (1d_shekel/hyperparams.py)."""
    return dict(field="shekel", n_v=1, n_t=0, t_min=0., t_max=0.,
                x_mesh=create_linear_mesh(0., 10., n_x),
                mu=_uniform(seed, n_s, SHEKEL_MU_MIN, SHEKEL_MU_MAX), eps=1e-10)


def ackley(n_x=400, n_y=400, n_s=1000, seed=2222):
    """C2: Ackley on [-5, 5]^2, mu in [-1, 1]^3 (2d_ackley/hyperparams.py)."""
    return dict(field="ackley", n_v=1, n_t=0, t_min=0., t_max=0.,
                x_mesh=create_linear_mesh(-5., 5., n_x, -5., 5., n_y),
                mu=_uniform(seed, n_s, [-1.] * 3, [1.] * 3), eps=1e-10)


def burgers(n_x=256, n_t=100, n_s=300, seed=3333, eps=1e-4):
    """C3: Burgers, x in [0, 1.5], t in [1, 5], mu in [0.001, 0.01] (1dt_burger/hyperparams.py)."""
    return dict(field="burgers", n_v=1, n_t=n_t, t_min=1., t_max=5.,
                x_mesh=create_linear_mesh(0., 1.5, n_x),
                mu=_uniform(seed, n_s, [0.001], [0.01]), eps=eps, eps_init=eps)


def dct_lowrank_host(n_h, n_s, r=256, decay=0.25, dtype=np.float64):
    """C4/C5 known-answer matrix on the host: U = sum_k s_k phi_k psi_k^T.

    phi_k, psi_k are DCT-II vectors (exactly orthonormal), s_k = 2**(-decay*k),
    so the singular values are exactly s_k and the left basis exactly phi_k
    (SURVEY.md §8d).  The device generator (pb_dct_lowrank) fills the same
    matrix directly in HBM for sizes the host cannot hold.
    """
    i = (np.arange(n_h) + .5)[:, None]
    j = (np.arange(n_s) + .5)[:, None]
    k = (np.arange(r) + 1.)[None, :]
    phi = np.sqrt(2. / n_h) * np.cos(np.pi * i * k / n_h)
    psi = np.sqrt(2. / n_s) * np.cos(np.pi * j * k / n_s)
    s = 2. ** (-decay * np.arange(r))
    return (phi * s).dot(psi.T).astype(dtype), s, phi


def dct_rank(s, eps):
    """Closed-form truncation rank of the DCT known-answer matrix (pod.py:22-32 rule)."""
    lam = s ** 2
    c = np.cumsum(lam) / np.sum(lam)
    return int(np.argmax(c >= 1 - eps) + 1)
