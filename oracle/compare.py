"""Comparison helpers for the parity tests (TEST INFRASTRUCTURE).

Tolerances come from BASELINE.json:north_star: singular values to 1e-10
relative to sigma_max, basis / coefficients up to per-mode sign with subspace
angle <= 1e-8, identical truncation rank, ensemble mean/variance 1e-6 relative.
"""
import numpy as np

TOL_SIGMA_REL = 1e-10
TOL_ANGLE = 1e-8
TOL_ENSEMBLE_REL = 1e-6


def sigma_error(sig, sig_ref):
    """max |sigma - sigma_ref| / sigma_max over the compared modes."""
    n = min(len(sig), len(sig_ref))
    return float(np.max(np.abs(sig[:n] - sig_ref[:n])) / sig_ref[0])


def largest_principal_angle(V, V_ref):
    """Largest principal angle between span(V) and span(V_ref) (radians).

    Both are orthonormalised first (the reference basis V = U z / sigma is only
    orthonormal to ~eps*sigma_max/sigma_min).  Uses the
    sine-based formula, accurate for small angles:
    sin(theta_max) = || (I - Q_ref Q_ref^T) Q ||_2.
    """
    Q, _ = np.linalg.qr(V)
    Qr, _ = np.linalg.qr(V_ref)
    R = Q - Qr.dot(Qr.T.dot(Q))
    # second projection pass removes the O(eps) residue of the first
    R = R - Qr.dot(Qr.T.dot(R))
    s = np.linalg.svd(R, compute_uv=False)
    return float(np.arcsin(min(1., s[0])))


def align_signs(V, V_ref):
    """Flip columns of V so that <V[:, i], V_ref[:, i]> >= 0; returns (V_aligned, signs)."""
    signs = np.sign(np.sum(V * V_ref, axis=0))
    signs[signs == 0] = 1.
    return V * signs, signs


def max_mode_error(V, V_ref):
    """Per-mode max-abs difference after sign alignment (diagnostic only: for
    clustered singular values individual modes rotate inside the cluster while
    the subspace is well defined)."""
    Va, _ = align_signs(V, V_ref)
    return np.max(np.abs(Va - V_ref), axis=0)


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), np.finfo(float).tiny))
