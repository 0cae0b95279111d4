"""Validation metrics used by the training log (reference poduqnn/metrics.py:11-30), numpy only."""
import numpy as np
from numpy.linalg import norm


def re(U, U_pred):
    """Relative error of one sample, metrics.py:11-13."""
    return norm(U - U_pred) / norm(U)


def re_max(U, U_pred):
    """metrics.py:16-18."""
    return norm(U - U_pred) / max(norm(U), norm(U_pred))


def re_s(U, U_pred, div_max=False):
    """Mean relative error over the columns, metrics.py:21-30."""
    n_s = U.shape[1]
    err = 0.
    for i in range(n_s):
        err += re_max(U[:, i], U_pred[:, i]) if div_max else re(U[:, i], U_pred[:, i])
    return err / n_s
