"""Snapshot-evaluation loops -- same names and argument order as the reference's
poduqnn/acceleration.py (loop_u :13-30, loop_u_t :34-70), evaluated by CUDA kernels that
write the snapshot matrix directly in HBM (csrc/fields.cu).

Differences a caller can observe: `u` must be a registered field (fields.resolve), and the
noise branches (u_noise / x_noise > 0, numba's unseeded RNG in the reference) are not built:
they raise NotImplementedError.
"""
import ctypes as C

import numpy as np

from . import _lib
from .fields import resolve


def time_grid(t_min, t_max, n_t):
    """The reference's time steps: np.linspace as numba's parallel JIT evaluates it
    (start + (i * (stop - start)) / (n - 1); last-bit different from numpy for some i)."""
    if n_t == 1:
        return np.array([float(t_min)])
    return t_min + (np.arange(n_t) * (float(t_max) - float(t_min))) / (n_t - 1)


def snapshots_device(ctx, u, X, mu_lhs, n_t=0, t_min=0., t_max=0., row0=0, n_rows=None, out=None, ackley_path=None):
    """Fill (and return) the device snapshot matrix for nodes [row0, row0+n_rows).

    X is (dim, n_xyz) as the reference passes it (x_mesh[:, 1:].T).  Returns
    (U_dev (n_rows, n_st), X_v host (n_st, n_d), t host or None).  `ackley_path` forces one of the three Ackley
    kernels ("plain" | "indexed" | "grouped"; None = pick by the mesh's repeated coordinates / radii).
    """
    f = resolve(u)
    X = np.ascontiguousarray(X, dtype=np.float64)
    mu = np.ascontiguousarray(mu_lhs, dtype=np.float64)
    dim, n_xyz = X.shape
    n_s, n_mu = mu.shape
    has_t = n_t > 0
    if has_t != f.has_t:
        raise ValueError(f"field '{f.name}' is {'time-dependent' if f.has_t else 'steady'}; got n_t={n_t}")
    n_rows = n_xyz - row0 if n_rows is None else n_rows
    n_st = n_s * (n_t if has_t else 1)
    t = time_grid(t_min, t_max, n_t) if has_t else None        # acceleration.py:38
    Xd, mud = ctx.to_device(X), ctx.to_device(mu)
    td = ctx.to_device(t) if has_t else None
    U = out if out is not None else ctx.alloc(n_rows, n_st)
    if f.name == "ackley" and dim == 2 and n_mu == 3:
        # tensor grids repeat coordinates: tabulate the separable term per distinct x / y (fields.cu)
        ux, ix = np.unique(X[0], return_inverse=True)
        uy, iy = np.unique(X[1], return_inverse=True)
        if 4 * (len(ux) + len(uy)) <= n_xyz and ackley_path != "plain":
            uxd, uyd = ctx.to_device(ux), ctx.to_device(uy)
            ixd, iyd = ctx.to_device_i32(ix), ctx.to_device_i32(iy)
            # symmetric meshes repeat the radius of the first term: order this call's nodes by r so that one exp
            # serves a whole run of equal radii (fields.cu, ackley_grouped_kernel); same bits as the indexed kernel
            xs, ys = X[0, row0:row0 + n_rows], X[1, row0:row0 + n_rows]
            r = np.sqrt(.5 * (xs ** 2 + ys ** 2))
            order = np.argsort(r, kind="stable")
            rs = r[order]
            n_runs = 1 + int(np.count_nonzero(rs[1:] != rs[:-1])) if n_rows > 0 else 0
            if ackley_path == "grouped" or (ackley_path is None and 2 * n_runs <= n_rows):
                permd, rsd = ctx.to_device_i32(order + row0), ctx.to_device(rs)
                ctx.check(_lib.lib().pb_snapshots_ackley_grouped(ctx.h, n_xyz, len(ux), uxd.ptr, ixd.ptr, len(uy), uyd.ptr,
                                                                 iyd.ptr, n_s, mud.ptr, row0, n_rows, permd.ptr, rsd.ptr,
                                                                 U.ptr, U.ld))
                return U, mu.copy(), None
            ctx.check(_lib.lib().pb_snapshots_ackley_indexed(ctx.h, n_xyz, Xd.ptr, len(ux), uxd.ptr, ixd.ptr, len(uy),
                                                             uyd.ptr, iyd.ptr, n_s, mud.ptr, row0, n_rows, U.ptr, U.ld))
            return U, mu.copy(), None
    ctx.check(_lib.lib().pb_snapshots(ctx.h, f.field_id, dim, n_xyz, Xd.ptr, n_s, n_mu, mud.ptr,
                                      n_t if has_t else 0, td.ptr if has_t else None,
                                      row0, n_rows, U.ptr, U.ld))
    if has_t:   # X_v rows [t_j, mu_i...], acceleration.py:53
        X_v = np.hstack((np.tile(t, n_s)[:, None], np.repeat(mu, n_t, axis=0)))
    else:       # acceleration.py:19
        X_v = mu.copy()
    return U, X_v, t


def _no_noise(u_noise, x_noise):
    if u_noise > 0. or x_noise > 0.:
        raise NotImplementedError("noisy snapshots (u_noise / x_noise > 0) are outside the CUDA hot path")


def loop_u(u, n_h, X_v, U, U_no_noise, X, mu_lhs, u_noise=0., x_noise=0., *, ctx=None):
    """Reference signature (acceleration.py:13): fills the caller's X_v, U, U_no_noise."""
    _no_noise(u_noise, x_noise)
    ctx = ctx or _lib.default_context()
    Ud, Xv, _ = snapshots_device(ctx, u, X, mu_lhs)
    if Ud.shape != (n_h, mu_lhs.shape[0]):
        raise ValueError(f"n_h={n_h} does not match the mesh ({Ud.shape[0]} nodes, n_v=1)")
    X_v[...] = Xv
    U[...] = Ud.download()
    U_no_noise[...] = U
    return X_v, U, U, U_no_noise


def loop_u_t(u, n_t, n_v, n_xyz, n_h, X_v, U, U_no_noise, U_struct, X, mu_lhs, t_min, t_max,
             u_noise=0., x_noise=0., *, ctx=None):
    """Reference signature (acceleration.py:34-35): fills X_v, U, U_no_noise, U_struct."""
    _no_noise(u_noise, x_noise)
    if n_v != 1:
        raise NotImplementedError("the registered analytic fields have n_v = 1")
    ctx = ctx or _lib.default_context()
    Ud, Xv, _ = snapshots_device(ctx, u, X, mu_lhs, n_t, t_min, t_max)
    n_s = mu_lhs.shape[0]
    X_v[...] = Xv
    U[...] = Ud.download()
    U_no_noise[...] = U
    Us = ctx.alloc(n_h, n_t, n_s)       # U_struct[:, :, i] = U[:, s:e], acceleration.py:69
    ctx.check(_lib.lib().pb_matrix_to_struct(ctx.h, Ud.ptr, Ud.ld, n_h, n_t, n_s, Us.ptr))
    U_struct[...] = Us.download()
    return X_v, U, U_struct, U_no_noise
