"""numpy prototype of the TSQR-mode arithmetic (csrc/qr.cu), written the way the kernels compute it.

Development aid, not on any product or test path.  Checks, on the ill-conditioned workloads, that
  * the panel factorisation with ONE reduction per column (partial dots below the diagonal row + the published
    diagonal row, LAPACK dlarfg formulas) and the tau / striu(Y^T Y) forward substitution (no explicit T) give a
    backward-stable R,
  * SVD(R) reproduces sigma and the right singular vectors of the tall matrix to the LAPACK floor,
  * a norm-sorted column order + second small QR leaves a factor on which one-sided Jacobi converges quickly.
"""
import sys

import numpy as np

sys.path.insert(0, "/root/repo")
from oracle import pod_oracle as po            # noqa: E402
from oracle import compare                     # noqa: E402
from pod_uqnn_b200 import workloads            # noqa: E402


def panel_factor(W, J, nbw, parts=7):
    """Householder QR of columns [J, J + nbw) of W over rows >= J, column by column; `parts` row groups stand in
    for the CTAs (partial sums are formed per group, then added in group order).  Returns tau (nbw)."""
    rows = W.shape[0]
    tau = np.zeros(nbw)
    bounds = np.linspace(J, rows, parts + 1).astype(int)
    for c in range(nbw):
        d = J + c
        g = np.zeros(nbw - c)
        for p in range(parts):
            lo, hi = max(bounds[p], d + 1), bounds[p + 1]
            if hi > lo:
                g += W[lo:hi, J + c] @ W[lo:hi, J + c:J + nbw]
        diag = W[d, J + c:J + nbw].copy() if d < rows else np.zeros(nbw - c)
        alpha = diag[0]
        xnorm2 = g[0]
        if xnorm2 == 0.:
            tau[c] = 0.
            continue
        beta = -np.copysign(np.sqrt(alpha * alpha + xnorm2), alpha)
        t = (beta - alpha) / beta
        scale = 1. / (alpha - beta)
        w = diag[1:] + g[1:] * scale
        tau[c] = t
        W[d, J + c] = beta
        W[d, J + c + 1:J + nbw] -= t * w
        v = W[d + 1:, J + c] * scale
        W[d + 1:, J + c] = v
        W[d + 1:, J + c + 1:J + nbw] -= t * np.outer(v, w)
    return tau


def explicit_y(W, J, nbw):
    rows = W.shape[0]
    Y = np.zeros((rows, nbw))
    for c in range(nbw):
        d = J + c
        if d < rows:
            Y[d, c] = 1.
            Y[d + 1:, c] = W[d + 1:, J + c]
    return Y


def apply_block(Y, tau, C):
    """C <- (I - Y T Y^T)^T C with T^-1 = diag(1/tau) + striu(Y^T Y), by forward substitution."""
    S = Y.T @ Y
    Z = Y.T @ C
    X = np.zeros_like(Z)
    for i in range(Y.shape[1]):
        X[i] = tau[i] * (Z[i] - S[:i, i] @ X[:i])
    C -= Y @ X


def qr_r(A, NB=64, nb=16):
    """R (m x m upper triangular) of A (rows x m), blocked as the GPU code: outer panels of NB, inner panels of nb."""
    rows0, m = A.shape
    rows = max(rows0, m)
    W = np.zeros((rows, m))
    W[:rows0] = A
    for J in range(0, m, NB):
        nbo = min(NB, m - J)
        taus = np.zeros(nbo)
        for i0 in range(0, nbo, nb):
            w = min(nb, nbo - i0)
            taus[i0:i0 + w] = panel_factor(W, J + i0, w)
            if i0 + w < nbo:
                Yin = explicit_y(W, J + i0, w)
                C = W[:, J + i0 + w:J + nbo]
                apply_block(Yin, taus[i0:i0 + w], C)
        if J + nbo < m:
            Y = explicit_y(W, J, nbo)
            apply_block(Y, taus, W[:, J + nbo:])
    return np.triu(W[:m])


def jacobi_sweeps(L, tol=None, max_sweeps=40):
    """one-sided Jacobi on the columns of L (cyclic by rows, vectorised per round-robin round); returns sweeps."""
    X = L.copy()
    n = X.shape[1]
    if n % 2:
        X = np.hstack([X, np.zeros((X.shape[0], 1))]); n += 1
    tol = tol or max(4., np.sqrt(X.shape[0])) * np.finfo(float).eps
    idx = list(range(n))
    for sweep in range(max_sweeps):
        rot = 0
        for rnd in range(n - 1):
            p = np.array([idx[i] for i in range(n // 2)])
            q = np.array([idx[n - 1 - i] for i in range(n // 2)])
            xp, xq = X[:, p], X[:, q]
            app = np.sum(xp * xp, 0); aqq = np.sum(xq * xq, 0); apq = np.sum(xp * xq, 0)
            act = (apq * apq > tol * tol * app * aqq) & (apq != 0)
            rot += int(act.sum())
            if act.any():
                d = aqq - app
                h = np.sqrt(d * d + 4 * apq * apq)
                c2 = .5 + .5 * np.abs(d) / np.where(h > 0, h, 1)
                c = np.sqrt(c2)
                s = np.where(d >= 0, apq, -apq) / np.where(h > 0, h, 1) / c
                c = np.where(act, c, 1.); s = np.where(act, s, 0.)
                X[:, p], X[:, q] = c * xp - s * xq, s * xp + c * xq
            idx = [idx[0]] + [idx[-1]] + idx[1:-1]
        if rot == 0:
            return sweep + 1, X
    return max_sweeps, X


def check(name, A, eps=1e-10, jac=False):
    rows, m = A.shape
    R = qr_r(A)
    G = A.T @ A
    print(f"{name}: {rows} x {m}  |R^T R - A^T A| / |A^T A| = {np.abs(R.T @ R - G).max() / np.abs(G).max():.2e}")
    Rl = np.linalg.qr(A, mode="r")
    sr = np.linalg.svd(R, compute_uv=False)
    s0 = np.linalg.svd(A, compute_uv=False)
    print(f"   sigma err vs svd(A): {np.abs(sr - s0).max() / s0[0]:.2e}   (LAPACK qr: {np.abs(np.linalg.svd(Rl, compute_uv=False) - s0).max() / s0[0]:.2e})")
    # rank + basis
    lam = s0 ** 2
    nL = int(np.argmax(np.cumsum(lam) / lam.sum() >= 1 - eps) + 1)
    _, s2, Zt = np.linalg.svd(R)
    V = A @ Zt[:nL].T / s2[:nL]
    _, _, Zt0 = np.linalg.svd(A, full_matrices=False)
    V0 = A @ Zt0[:nL].T / s0[:nL]
    print(f"   n_L = {nL}, angle(V_tsqr, V_svd) = {compare.largest_principal_angle(V, V0):.2e}")
    if jac:
        nrm = np.linalg.norm(R, axis=0)
        order = np.argsort(-nrm, kind="stable")
        sw_plain, _ = jacobi_sweeps(R.T)
        R2 = qr_r(R[:, order])
        sw_sorted, X = jacobi_sweeps(R2.T)
        sj = np.sort(np.linalg.norm(X, axis=0))[::-1][:m]
        print(f"   Jacobi sweeps on R^T: {sw_plain}; after norm-sorted re-QR: {sw_sorted}; sigma err {np.abs(sj - s0).max() / s0[0]:.2e}")


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    check("random", rng.standard_normal((700, 130)), jac=True)
    w = workloads.shekel(n_x=400, n_s=300)
    U = po.create_snapshots(w["x_mesh"], 1, 0, po.FIELDS["shekel"], w["mu"])[1]
    check("shekel 400x300 (tall)", U, jac=True)
    check("shekel^T 300x400 -> pad", U.T[:, :200].copy(), jac=False)
    w = workloads.ackley(n_x=48, n_y=48, n_s=160)
    U = po.create_snapshots(w["x_mesh"], 1, 0, po.FIELDS["ackley"], w["mu"])[1]
    check("ackley 2304x160", U, jac=True)
    A, s, phi = workloads.dct_lowrank_host(3000, 192, r=64)
    check("dct 3000x192 r=64", A, jac=True)


def qrcp_noswap(R1, tol_rel):
    """Householder QR with column pivoting of the small square factor, WITHOUT moving columns: step c eliminates the
    remaining column of largest norm below row c.  Rows of the result = the factor vectors, already in the original
    column order.  Stops when the largest remaining norm <= tol_rel * first pivot norm.  Returns (W[:r], r)."""
    W = R1.copy()
    m = W.shape[0]
    done = np.zeros(m, bool)
    first = None
    r = m
    for c in range(m):
        cn2 = np.where(done, -1., np.sum(W[c:] ** 2, axis=0))
        p = int(np.argmax(cn2))
        if first is None:
            first = cn2[p]
        if not (cn2[p] > tol_rel * tol_rel * first) or cn2[p] <= 0.:
            r = c
            break
        x = W[c:, p].copy()
        alpha = x[0]
        xn2 = float(x[1:] @ x[1:])
        if xn2 == 0.:
            done[p] = True
            continue
        beta = -np.copysign(np.sqrt(alpha * alpha + xn2), alpha)
        tau = (beta - alpha) / beta
        scale = 1. / (alpha - beta)
        act = ~done
        act[p] = False
        g = x[1:] @ W[c + 1:, act]
        w = W[c, act] + g * scale
        W[c, act] -= tau * w
        W[c + 1:, act] -= tau * np.outer(x[1:] * scale, w)
        W[c, p] = beta
        W[c + 1:, p] = 0.
        done[p] = True
    return W[:r].copy(), r


def check2(name, A, eps=1e-10, full=False):
    rows, m = A.shape
    R1 = qr_r(A)
    tol = 0. if full else m * np.finfo(float).eps
    F, r = qrcp_noswap(R1, tol)
    sw, X = jacobi_sweeps(F.T)
    nrm = np.linalg.norm(X, axis=0)
    order = np.argsort(-nrm)
    s0 = np.linalg.svd(A, compute_uv=False)
    sj = np.zeros(m); sj[:min(r, m)] = nrm[order][:min(r, m)]
    lam = s0 ** 2
    nL = int(np.argmax(np.cumsum(lam) / lam.sum() >= 1 - eps) + 1)
    Z = X[:, order[:nL]] / nrm[order[:nL]]
    V = A @ Z / nrm[order[:nL]]
    _, _, Zt0 = np.linalg.svd(A, full_matrices=False)
    V0 = A @ Zt0[:nL].T / s0[:nL]
    lam2 = sj ** 2
    nL2 = int(np.argmax(np.cumsum(lam2) / lam2.sum() >= 1 - eps) + 1)
    print(f"{name}: r = {r} of {m}, Jacobi sweeps {sw}, sigma err {np.abs(sj - s0)[:max(r, 1)].max() / s0[0]:.2e}, "
          f"n_L {nL2} (ref {nL}), angle {compare.largest_principal_angle(V, V0):.2e}")


if __name__ == "__main__":
    print("---- QRCP (no swap) + Jacobi")
    rng = np.random.default_rng(0)
    check2("random", rng.standard_normal((700, 130)))
    w = workloads.shekel(n_x=400, n_s=300)
    U = po.create_snapshots(w["x_mesh"], 1, 0, po.FIELDS["shekel"], w["mu"])[1]
    check2("shekel 400x300", U)
    check2("shekel 400x300 full", U, full=True)
    w = workloads.ackley(n_x=48, n_y=48, n_s=160)
    U = po.create_snapshots(w["x_mesh"], 1, 0, po.FIELDS["ackley"], w["mu"])[1]
    check2("ackley 2304x160", U)
    check2("ackley 2304x160 full", U, full=True)
    A, s, phi = workloads.dct_lowrank_host(3000, 192, r=64)
    check2("dct 3000x192 r=64", A)
