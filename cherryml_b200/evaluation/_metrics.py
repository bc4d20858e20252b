"""Distances between two rate matrices, as the reference reports them when it compares a learned
matrix with the truth (``cherryml/evaluation/_metrics.py:14-125``; its plotting helpers are not
part of this package).  All of them look at off-diagonal entries only, and only at entries a
``mask_matrix`` (1 = keep) allows."""
from typing import List, Optional

import numpy as np


def _kept_off_diagonal(num_states: int, mask_matrix: Optional[np.ndarray]) -> np.ndarray:
    keep = ~np.eye(num_states, dtype=bool)
    if mask_matrix is not None:
        keep &= np.asarray(mask_matrix) != 0
    return keep


def _log_ratios(y: np.ndarray, y_hat: np.ndarray, mask_matrix: Optional[np.ndarray]) -> np.ndarray:
    y, y_hat = np.asarray(y), np.asarray(y_hat)
    if y.shape != y_hat.shape:
        raise ValueError(
            f"y and y_hat should have the same shape. Shapes are: y.shape={y.shape}, y_hat.shape={y_hat.shape}"
        )
    assert y.ndim == 2 and y.shape[0] == y.shape[1]
    out = np.zeros(y.shape)
    keep = _kept_off_diagonal(y.shape[0], mask_matrix)
    out[keep] = np.log(y[keep] / y_hat[keep])
    return out


def l_infty_norm(y, y_hat, mask_matrix=None) -> float:
    """Largest absolute log ratio (reference :39-46)."""
    return np.max(np.abs(_log_ratios(y, y_hat, mask_matrix)))


def rmse(y, y_hat, mask_matrix=None) -> float:
    """Root mean square of the log ratios.  With a mask the divisor is ``mask.sum() - num_states``,
    i.e. the reference assumes the mask's diagonal is set (:49-64)."""
    num_states = np.asarray(y).shape[0]
    lr = _log_ratios(y, y_hat, mask_matrix)
    n = np.asarray(mask_matrix).sum().sum() - num_states if mask_matrix is not None else num_states * (num_states - 1)
    return np.sqrt(np.sum(lr * lr) / n)


def mre(y, y_hat, mask_matrix=None) -> float:
    """Max relative error (reference :67-75)."""
    return np.exp(l_infty_norm(y, y_hat, mask_matrix)) - 1


def relative_error(y: float, y_hat: float) -> float:
    assert y > 0
    assert y_hat > 0
    return y / y_hat - 1 if y > y_hat else y_hat / y - 1


def relative_errors(y, y_hat, mask_matrix=None) -> List[float]:
    """``max(y, y_hat) / min(y, y_hat) - 1`` for the kept off-diagonal entries in row-major order;
    a mask keeps the entries equal to 1 (reference :90-108)."""
    y, y_hat = np.asarray(y), np.asarray(y_hat)
    num_states = y.shape[0]
    keep = ~np.eye(num_states, dtype=bool)
    if mask_matrix is not None:
        keep &= np.asarray(mask_matrix) == 1
    return [relative_error(a, b) for a, b in zip(y[keep], y_hat[keep])]


def mean_relative_error(y, y_hat, mask_matrix=None) -> float:
    """Average relative error (reference :111-125)."""
    return np.mean(relative_errors(y=y, y_hat=y_hat, mask_matrix=mask_matrix))
