"""Rate-matrix distances (reference cherryml/evaluation/_metrics.py).  The expected values were
produced by the unmodified reference functions on the same seeded matrices (bit-equal when both
were run side by side in the build container)."""
import numpy as np
import pytest

from cherryml_b200 import evaluation as ev


def _Q(rng, n):
    a = rng.uniform(0.01, 2, (n, n))
    np.fill_diagonal(a, 0)
    np.fill_diagonal(a, -a.sum(1))
    return a


REFERENCE = {  # (n, masked) -> l_infty_norm, rmse, mre, mean_relative_error
    (4, False): (3.108962467096752, 1.1349593693782083, 21.397793893124437, 2.731311035708711),
    (20, False): (4.824143652895586, 1.33083786797605, 123.4798247902293, 3.8092763664986693),
    (6, True): (3.4268767320405193, 1.2626057315532444, 29.780357161797, 3.270188170150071),
    (20, True): (4.91131806141794, 1.3678109224336215, 134.8183133613528, 4.238312362686482),
}


def test_distances_equal_the_reference_values():
    rng = np.random.default_rng(1)
    for (n, masked), want in REFERENCE.items():
        y, y_hat = _Q(rng, n), _Q(rng, n)
        mask = None
        if masked:
            mask = (rng.random((n, n)) < 0.6).astype(int)
            np.fill_diagonal(mask, 1)
            y, y_hat = y * mask, y_hat * mask
        got = (ev.l_infty_norm(y, y_hat, mask), ev.rmse(y, y_hat, mask), ev.mre(y, y_hat, mask),
               ev.mean_relative_error(y, y_hat, mask))
        assert got == want
        errs = ev.relative_errors(y, y_hat, mask)
        assert len(errs) == (int(mask.sum()) - n if masked else n * (n - 1))
        assert abs(max(errs) - got[2]) < 1e-9 * got[2]  # max relative error = exp(max |log ratio|) - 1


def test_identical_matrices_and_bad_shapes():
    rng = np.random.default_rng(2)
    q = _Q(rng, 5)
    assert ev.l_infty_norm(q, q) == 0 and ev.rmse(q, q) == 0 and ev.mre(q, q) == 0 and ev.mean_relative_error(q, q) == 0
    assert ev.relative_error(2.0, 1.0) == 1.0 and ev.relative_error(1.0, 4.0) == 3.0
    with pytest.raises(ValueError, match="same shape"):
        ev.l_infty_norm(q, _Q(rng, 4))
