"""numpy prototype of the device algorithm (dev aid, not shipped, not the oracle):
Gram -> pivoted Cholesky (rank revealing) -> one-sided Jacobi on L -> eigenpairs;
'accurate' mode: rotate (Y = U W_r), second Gram, Jacobi again."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from oracle import pod_oracle as po, compare as cmp
from pod_uqnn_b200 import workloads as wl

def pivoted_cholesky(G, tol_rel=None):
    n = G.shape[0]
    d = np.diag(G).copy()
    L = np.zeros((n, n))
    piv = []
    dmax0 = d.max()
    tol = (tol_rel if tol_rel is not None else n * np.finfo(float).eps) * dmax0
    for k in range(n):
        p = int(np.argmax(d))
        if d[p] <= tol:
            break
        piv.append(p)
        col = G[:, p] - L[:, :k].dot(L[p, :k])
        col /= np.sqrt(d[p])
        L[:, k] = col
        d -= col**2
        d[p] = -np.inf  # never pick again
    r = len(piv)
    return L[:, :r], piv

def onesided_jacobi(A, tol=1e-15, max_sweeps=30):
    """columns of A orthogonalised in place; returns sweeps"""
    n, r = A.shape
    for sweep in range(max_sweeps):
        rotated = 0
        for p in range(r - 1):
            for q in range(p + 1, r):
                a = A[:, p].dot(A[:, p]); b = A[:, q].dot(A[:, q]); c = A[:, p].dot(A[:, q])
                if abs(c) <= tol * np.sqrt(a * b):
                    continue
                rotated += 1
                zeta = (b - a) / (2 * c)
                t = np.sign(zeta) / (abs(zeta) + np.sqrt(1 + zeta**2)) if zeta != 0 else 1.0
                cs = 1 / np.sqrt(1 + t**2); sn = cs * t
                Ap = A[:, p].copy()
                A[:, p] = cs * Ap - sn * A[:, q]
                A[:, q] = sn * Ap + cs * A[:, q]
        if rotated == 0:
            return sweep + 1
    return max_sweeps

def eig_psd(G):
    L, piv = pivoted_cholesky(G)
    A = L.copy()
    sweeps = onesided_jacobi(A)
    sig = np.linalg.norm(A, axis=0)
    order = np.argsort(-sig)
    sig = sig[order]
    W = A[:, order] / sig
    return sig**2, W, sweeps, L.shape[1]

def pod_gram(U, eps):
    tall = U.shape[0] >= U.shape[1]
    G = U.T.dot(U) if tall else U.dot(U.T)
    lam, W, sweeps, r = eig_psd(G)
    lam_full = np.zeros(G.shape[0]); lam_full[:r] = lam
    n_L = po.truncation_rank(lam_full, eps, U.shape[1])
    if tall:
        V = U.dot(W[:, :n_L]) / np.sqrt(lam[:n_L])
    else:
        V = W[:, :n_L]
    return V, np.sqrt(lam_full), n_L, (sweeps, r), W, lam

def pod_accurate(U, eps):
    tall = U.shape[0] >= U.shape[1]
    A = U if tall else U.T           # A is tall: n x m
    V1, sig1, n_L1, info, W, lam = pod_gram(U, eps)
    r = W.shape[1]
    Y = A.dot(W)                     # n x r, nearly orthogonal columns
    G2 = Y.T.dot(Y)
    # Jacobi on scaled G2 = D (I+E) D: do one-sided Jacobi on chol(G2)^T ... prototype: scale first
    d = np.sqrt(np.diag(G2))
    A2 = G2 / np.outer(d, d)
    lam2s, W2s = np.linalg.eigh(A2)   # placeholder for Jacobi on well-conditioned matrix
    # better: one-sided Jacobi on Y's Cholesky factor
    L2, piv2 = pivoted_cholesky(G2, tol_rel=0.0)
    B = L2.copy()
    sw = onesided_jacobi(B)
    s = np.linalg.norm(B, axis=0)
    order = np.argsort(-s); s = s[order]
    W2 = B[:, order] / s              # eigenvectors of G2
    lam_full = np.zeros(min(U.shape)); lam_full[:len(s)] = s**2
    n_L = po.truncation_rank(lam_full, eps, U.shape[1])
    Z = W.dot(W2[:, :n_L])           # right sing vecs of A
    if tall:
        V = Y.dot(W2[:, :n_L]) / s[:n_L]
    else:
        V = Z
    return V, np.sqrt(lam_full), n_L, (info, sw)

def report(name, U, eps):
    t = time.time(); Vr = po.perform_pod(U, eps); D, _ = po.pod_spectrum(U); tr = time.time() - t
    for mode, fn in (("gram", pod_gram), ("accurate", pod_accurate)):
        t = time.time(); out = fn(U, eps); tm = time.time() - t
        V, sig, n_L, info = out[0], out[1], out[2], out[3]
        nn = min(len(sig), len(D))
        print(f"{name:28s} {mode:9s} n_L {n_L} (ref {Vr.shape[1]}) sigerr(all) {cmp.sigma_error(sig, D):.2e} "
              f"sigerr(kept) {cmp.sigma_error(sig[:n_L], D[:n_L]):.2e} "
              f"angle {cmp.largest_principal_angle(V, Vr) if n_L == Vr.shape[1] else float('nan'):.2e} info {info} t {tm:.1f}s ref {tr:.1f}s")

if __name__ == "__main__":
    w = wl.shekel(400, 300)
    X_v, U, _, _ = po.create_snapshots(w["x_mesh"], 1, 0, po.u_shekel, w["mu"])
    report("shekel 400x300", U, 1e-10)
    w = wl.ackley(48, 48, 160)
    X_v, U, _, _ = po.create_snapshots(w["x_mesh"], 1, 0, po.u_ackley, w["mu"])
    report("ackley 2304x160", U, 1e-10)
    w = wl.burgers(256, 100, 3)
    X_v, U, Us, _ = po.create_snapshots(w["x_mesh"], 1, 100, po.u_burgers, w["mu"], 1., 5.)
    for eps in (1e-4, 1e-8, 1e-10):
        report(f"burgers 256x300 eps={eps}", U, eps)
        report(f"burgers sample0 256x100 eps={eps}", Us[:, :, 0], eps)
