from ._engine import FitEngine, random_theta, theta_from_initialization  # noqa: F401
from ._bench import bench_fit  # noqa: F401
from ._jtt_ipw import jtt_ipw, jtt_ipw_from_counts  # noqa: F401
from ._quantized_transitions_mle import RateMatrixLearner, quantized_transitions_mle  # noqa: F401


def _smoke_fit(counts, grid) -> str:
    """Used by __graft_entry__.smoke(): 30 Adam epochs on freshly counted LG matrices, checked
    against the fp64 oracle (loss trace within 1e-6 relative)."""
    import numpy as np
    import torch

    from oracle.fit_oracle import fit_oracle

    c = counts.cpu().numpy()
    S = c.shape[-1]
    theta0 = random_theta(S, seed=0)
    eng = FitEngine(np.asarray(sorted(grid)), counts, theta0, num_epochs=30, device=counts.device)
    eng.run()
    res = eng.results()
    ref = fit_oracle(sorted(grid), c, num_epochs=30, dtype=torch.float64)
    rel = float(np.max(np.abs(res["loss"] - ref["loss"]) / np.abs(ref["loss"])))
    assert rel < 1e-6, f"fit loss trace differs from the oracle: rel {rel}"
    return f"fit ok (30 epochs, loss {res['loss'][0]:.6f} -> {res['loss'][-1]:.6f}, max rel diff vs oracle {rel:.1e})"
