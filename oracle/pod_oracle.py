"""numpy restatement of the reference's POD / snapshot / projection path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  float64 throughout, as the
reference (podnnmodel.py:56-57).  Citations are relative to /root/reference.
"""
import numpy as np


# --------------------------------------------------------------------------
# mesh (poduqnn/mesh.py:11-42) -- needed to build x_mesh for the fields
# --------------------------------------------------------------------------
def create_linear_mesh(x_min, x_max, n_x, y_min=0, y_max=0, n_y=0):
    """1-D / 2-D tensor grid; column 0 is the 1-based node id (mesh.py:11-42).

    2-D node order is ``iy * n_x + ix`` (np.meshgrid default 'xy' indexing,
    mesh.py:36-39).
    """
    x = np.linspace(x_min, x_max, n_x).reshape((n_x, 1))
    if n_y > 0:
        y = np.linspace(y_min, y_max, n_y).reshape((n_y, 1))
        X, Y = np.meshgrid(x, y)
        n_xyz = n_x * n_y
        idx = np.arange(1, n_xyz + 1).reshape((n_xyz, 1))
        return np.hstack((idx, X.reshape((n_xyz, 1)), Y.reshape((n_xyz, 1))))
    idx = np.arange(1, n_x + 1).reshape((n_x, 1))
    return np.hstack((idx, x))


# --------------------------------------------------------------------------
# analytic fields u(X, t, mu)
# --------------------------------------------------------------------------
def u_shekel(X, _, mu):
    """experiments/1d_shekel/hyperparams.py:55-67."""
    x = X[0]
    sep = int(mu.shape[0] / 2)
    bet = mu[:sep]
    gam = mu[sep:]
    u_sum = np.zeros_like(x)
    for i in range(len(bet)):
        i_sum = (x - gam[i])**2
        u_sum += 1 / (bet[i] + i_sum)
    return u_sum.reshape((1, u_sum.shape[0]))


def u_ackley(X, _, mu):
    """experiments/2d_ackley/hyperparams.py:51-57."""
    x = X[0]
    y = X[1]
    u_0 = - 20*(1+.1*mu[2])*np.exp(-.2*(1+.1*mu[1])*np.sqrt(.5*(x**2+y**2))) \
          - np.exp(.5*(np.cos(2*np.pi*(1+.1*mu[0])*x) + np.cos(2*np.pi*(1+.1*mu[0])*y))) \
          + 20 + np.exp(1)
    return u_0.reshape((1, u_0.shape[0]))


def u_burgers(X, t, mu):
    """experiments/1dt_burger/hyperparams.py:45-55 (relies on inf -> 0)."""
    x = X[0]
    mu = mu[0]
    with np.errstate(over="ignore"):
        if t == 1.:
            res = x / (1 + np.exp(1/(4*mu)*(x**2 - 1/4)))
        else:
            t0 = np.exp(1 / (8*mu))
            res = (x/t) / (1 + np.sqrt(t/t0)*np.exp(x**2/(4*mu*t)))
    return res.reshape((1, x.shape[0]))


FIELDS = {"shekel": u_shekel, "ackley": u_ackley, "burgers": u_burgers}


# --------------------------------------------------------------------------
# snapshot loops (poduqnn/acceleration.py), noise-free branch
# --------------------------------------------------------------------------
def loop_u(u, n_h, X_v, U, U_no_noise, X, mu_lhs):
    """acceleration.py:13-30 with u_noise = x_noise = 0."""
    for i in range(mu_lhs.shape[0]):
        X_v[i, :] = mu_lhs[i]
        U_i = u(X, 0, X_v[i, :]).reshape((n_h,))
        U[:, i] = U_i
        U_no_noise[:, i] = U_i
    return X_v, U, U, U_no_noise


def linspace_numba_parallel(start, stop, num):
    """np.linspace as numba evaluates it inside the reference's @jit(parallel=True) loop_u_t
    (acceleration.py:38): start + (i * (stop - start)) / (num - 1), which differs from numpy's
    start + i * step in the last bit for some i (pinned against the reference in gen_golden.py)."""
    if num == 1:
        return np.array([float(start)])
    return start + (np.arange(num) * (float(stop) - float(start))) / (num - 1)


def loop_u_t(u, n_t, n_v, n_xyz, n_h, X_v, U, U_no_noise, U_struct, X, mu_lhs,
             t_min, t_max):
    """acceleration.py:34-70 with u_noise = x_noise = 0."""
    t = linspace_numba_parallel(t_min, t_max, n_t)
    tT = t.reshape((n_t, 1))
    for i in range(mu_lhs.shape[0]):
        s = n_t * i
        e = n_t * (i + 1)
        mu_i = mu_lhs[i, :]
        X_v[s:e, :] = np.hstack((tT, np.ones_like(tT)*mu_i))
        Ui = np.zeros((n_v, n_xyz, n_t))
        for j in range(n_t):
            Ui[:, :, j] = u(X, t[j], mu_i)
        U[:, s:e] = Ui.reshape((n_h, n_t))
        U_no_noise[:, s:e] = U[:, s:e]
        U_struct[:, :, i] = U[:, s:e]
    return X_v, U, U_struct, U_no_noise


def create_snapshots(x_mesh, n_v, n_t, u, mu_lhs, t_min=0, t_max=0):
    """podnnmodel.py:89-116 (noise-free)."""
    n_s = mu_lhs.shape[0]
    n_xyz = x_mesh.shape[0]
    n_h = n_v * n_xyz
    has_t = n_t > 0
    n_st = n_s * (n_t if has_t else 1)
    n_d = mu_lhs.shape[1] + (1 if has_t else 0)
    X = x_mesh[:, 1:].T
    X_v = np.zeros((n_st, n_d))
    U = np.zeros((n_h, n_st))
    U_nn = np.zeros((n_h, n_st))
    if has_t:
        U_struct = np.zeros((n_h, n_t, n_s))
        return loop_u_t(u, n_t, n_v, n_xyz, n_h, X_v, U, U_nn, U_struct, X,
                        mu_lhs, t_min, t_max)
    return loop_u(u, n_h, X_v, U, U_nn, X, mu_lhs)


# --------------------------------------------------------------------------
# POD (poduqnn/pod.py)
# --------------------------------------------------------------------------
def pod_spectrum(U):
    """pod.py:16-23: thin SVD, lambdas = D**2, Z = ZT.T."""
    _, D, ZT = np.linalg.svd(U, full_matrices=False)
    return D, ZT.T


def truncation_rank(lambdas, eps, n_st):
    """pod.py:22-32: sequential fp64 accumulation, descending order.

    ``sum_lambdas`` is a plain left-to-right sum (numba's np.sum), and the
    search stops at the first i with sum_trunc / sum_lambdas >= 1 - eps.  The
    reference bounds the loop by n_st, not len(lambdas) (pod.py:28); this
    restatement clamps to len(lambdas) (the out-of-bounds read is a bug, not
    behaviour to reproduce).
    """
    sum_lambdas = 0.
    for lam in lambdas:
        sum_lambdas += lam
    n_L = 0
    acc = 0.
    for i in range(min(n_st, len(lambdas))):
        acc += lambdas[i]
        n_L += 1
        if acc / sum_lambdas >= (1 - eps):
            break
    return n_L


def perform_pod(U, eps=0., n_L=0, verbose=True):
    """pod.py:7-47."""
    n_h, n_st = U.shape
    D, Z = pod_spectrum(U)
    lambdas = D**2
    if n_L == 0:
        n_L = truncation_rank(lambdas, eps, n_st)
    U = np.ascontiguousarray(U)
    V = np.zeros((n_h, n_L))
    for i in range(n_L):
        V[:, i] = U.dot(np.ascontiguousarray(Z[:, i])) / np.sqrt(lambdas[i])
    return np.ascontiguousarray(V)


def perform_fast_pod(U, eps, eps_init, return_ranks=False):
    """pod.py:51-68: per-sample POD (eps_init), unweighted concat, global POD (eps)."""
    n_s = U.shape[-1]
    T_list = []
    for k in range(n_s):
        T_list.append(perform_pod(U[:, :, k], eps=eps_init, n_L=0, verbose=False))
    U_f = np.concatenate(T_list, axis=1)
    V = perform_pod(U_f, eps=eps, n_L=0, verbose=True)
    if return_ranks:
        return V, np.array([T.shape[1] for T in T_list])
    return V


# --------------------------------------------------------------------------
# projection / reconstruction / POD error (poduqnn/podnnmodel.py)
# --------------------------------------------------------------------------
def project_to_v(V, U):
    """podnnmodel.py:435-437 and :247-248."""
    return (V.T.dot(U)).T


def project_to_U(V, v):
    """podnnmodel.py:431-433."""
    return V.dot(v.T)


def pod_sig(U, U_pod, chunk=4096):
    """podnnmodel.py:252: np.stack((U, U_pod), -1).std(-1).mean(-1).

    Chunked over rows only to bound the temporaries; the arithmetic per row is
    the reference's (population std of the two values, then the row mean).
    """
    out = np.empty(U.shape[0])
    for r in range(0, U.shape[0], chunk):
        out[r:r+chunk] = np.stack((U[r:r+chunk], U_pod[r:r+chunk]),
                                  axis=-1).std(-1).mean(-1)
    return out


def restruct(U, n_v, n_xyz, n_t):
    """podnnmodel.py:382-402."""
    if n_t > 0:
        n_s = int(U.shape[-1] / n_t)
        U_struct = np.zeros((n_v, n_xyz, n_t, n_s))
        for i in range(n_s):
            U_struct[:, :, :, i] = U[:, n_t*i:n_t*(i+1)].reshape((n_v, n_xyz, n_t))
        return U_struct
    n_s = U.shape[-1]
    U_struct = np.zeros((n_v, n_xyz, n_s))
    for i in range(n_s):
        U_struct[:, :, i] = U[:, i].reshape((n_v, n_xyz))
    return U_struct


def destruct(U_struct, n_t):
    """podnnmodel.py:404-421."""
    n_v, n_xyz = U_struct.shape[0], U_struct.shape[1]
    n_h = n_v * n_xyz
    n_s = U_struct.shape[-1]
    if n_t > 0:
        U = np.zeros((n_h, n_t * n_s))
        for i in range(n_s):
            U[:, n_t*i:n_t*(i+1)] = U_struct[:, :, :, i].reshape((n_h, n_t))
        return U
    U = np.zeros((n_h, n_s))
    for i in range(n_s):
        U[:, i] = U_struct[:, :, i].reshape((n_h,))
    return U
