"""``quantized_transitions_mle``: drop-in stage function for the rate-matrix fit.

Same name, keyword arguments, defaults, caching behaviour and output files as the
reference's ``cherryml/estimation/_quantized_transitions_mle.py:35-122`` (+ the
``RateMatrixLearner`` it drives, ``_ratelearn/ratelearner.py:34-184``); the optimisation
itself runs on the GPU in fp64 (``FitEngine``).

Differences a user can observe, all documented in DESIGN.md:
* ``device`` is accepted for compatibility; the computation always runs on a CUDA device
  (``"cuda:1"`` etc. select which); there is no CPU path.
* matrices are written with fp64 digits (the reference prints its fp32 tensors);
* ``df_res.txt`` has the reference's columns; ``time`` is interpolated from the total
  because epochs are never synchronised with the host individually;
* ``training_plot.png`` is written only if matplotlib is importable;
* when the counts were produced by ``count_transitions`` in this process they are taken
  from device memory instead of re-parsing ``result.txt`` (same values).
"""
import logging
import os
import time
from typing import List, Optional

import numpy as np
import torch

from .. import caching
from ..io import read_count_matrices_array, read_mask_matrix, read_rate_matrix, write_rate_matrix
from ._engine import FitEngine, random_theta, theta_from_initialization

logger = logging.getLogger(__name__)


def _cuda_device(device: str) -> torch.device:
    if isinstance(device, str) and device.startswith("cuda"):
        return torch.device(device)
    return torch.device("cuda", torch.cuda.current_device())


class RateMatrixLearner:
    """Counterpart of the reference's ``RateMatrixLearner`` (same constructor and methods)."""

    def __init__(
        self,
        branches: List[float],
        mats,
        states: List[str],
        output_dir: Optional[str],
        stationnary_distribution: Optional[str],
        device: str,
        mask=None,
        rate_matrix_parameterization: str = "pande_reversible",
        initialization: Optional[np.ndarray] = None,
        skip_writing_to_output_dir: bool = False,
    ):
        if rate_matrix_parameterization != "pande_reversible":
            raise NotImplementedError(
                f"rate_matrix_parameterization={rate_matrix_parameterization!r}: only "
                "'pande_reversible' (the one CherryML's pipelines use) is implemented"
            )
        if stationnary_distribution is not None:
            raise NotImplementedError("a fixed stationary distribution is not supported")
        self.branches = [float(b) for b in branches]
        self.mats = mats
        self.states = list(states)
        self.output_dir = None if skip_writing_to_output_dir else output_dir
        self.skip_writing_to_output_dir = skip_writing_to_output_dir
        self.device = device
        self.initialization = initialization
        if isinstance(mask, str):
            mask = np.loadtxt(mask)
        self.mask = None if mask is None else np.asarray(mask, dtype=np.float64)
        self.trained = False
        self.df_res = None
        self.Q_dict = None

    def train(self, lr=1e-1, num_epochs=2000, do_adam: bool = True, loss_normalization: bool = False,
              return_best_iter: bool = True):
        start = time.time()
        S = len(self.states)
        mask = np.ones((S, S)) if self.mask is None else self.mask
        if self.initialization is not None:
            theta0 = theta_from_initialization(np.asarray(self.initialization, dtype=np.float64), mask)
        else:
            theta0 = random_theta(S, seed=0)
        counts = self.mats if isinstance(self.mats, torch.Tensor) else np.stack(
            [np.asarray(m, dtype=np.float64) for m in self.mats])
        engine = FitEngine(
            times=np.asarray(self.branches), counts=counts, theta0=theta0, mask=mask, num_epochs=num_epochs,
            learning_rate=lr, do_adam=do_adam, loss_normalization=loss_normalization, best_mode=0,
            device=_cuda_device(self.device),
        )
        if self.initialization is not None:
            # the reference asserts that the parameterisation reproduces the initialisation
            # to 3 decimals (rate.py:89-91)
            np.testing.assert_almost_equal(engine.Q[0].cpu().numpy(), self.initialization, decimal=3)
        engine.run()
        res = engine.results()
        total = time.time() - start
        Q_dict = {k: v for k, v in res.items() if k.startswith("Q_")}
        if num_epochs > 0:
            Q_dict["result"] = (res["Q_best"] if return_best_iter else res["Q_last"]).copy()
        self.Q_dict = Q_dict
        import pandas as pd

        n = len(res["loss"])
        self.df_res = pd.DataFrame({
            "nuc_norm": np.zeros(n), "frob_norm": np.zeros(n), "loss": res["loss"],
            "time": total * (np.arange(n) + 1) / max(n, 1), "epoch": np.arange(n),
            "frob_norm_diag": np.zeros(n), "frob_norm_offdiag": np.zeros(n),
        })
        self.engine = engine
        self.trained = True
        if not self.skip_writing_to_output_dir:
            self.process_results()

    def process_results(self):
        os.makedirs(self.output_dir, exist_ok=True)
        for key, value in self.Q_dict.items():
            write_rate_matrix(value, self.states, os.path.join(self.output_dir, key + ".txt"))
        self.df_res.to_csv(os.path.join(self.output_dir, "df_res.txt"))
        try:
            import matplotlib

            matplotlib.use("Agg")
            import matplotlib.pyplot as plt

            plt.subplots(figsize=(5, 4))
            self.df_res.loss.plot()
            plt.xscale("log")
            plt.ylabel("Negative likelihood", fontsize=13)
            plt.xlabel("# of iterations", fontsize=13)
            plt.tight_layout()
            plt.savefig(os.path.join(self.output_dir, "training_plot.png"))
            plt.close()
        except Exception:  # matplotlib is optional
            pass

    def get_learnt_rate_matrix(self):
        if not self.trained:
            raise ValueError("Model should be trained first!")
        import pandas as pd

        return pd.DataFrame(self.Q_dict["result"], columns=self.states, index=self.states)


@caching.cached_computation(
    output_dirs=["output_rate_matrix_dir"],
    exclude_args=["device", "OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS"],
    write_extra_log_files=True,
)
def quantized_transitions_mle(
    count_matrices_path: str,
    initialization_path: Optional[str],
    mask_path: Optional[str],
    output_rate_matrix_dir: Optional[str],
    stationary_distribution_path: Optional[str] = None,
    rate_matrix_parameterization: str = "pande_reversible",
    device: str = "cpu",
    learning_rate: float = 1e-1,
    num_epochs: int = 2000,
    do_adam: bool = True,
    loss_normalization: bool = True,
    OMP_NUM_THREADS: Optional[int] = 1,
    OPENBLAS_NUM_THREADS: Optional[int] = 1,
    return_best_iter: bool = True,
):
    start_time = time.time()
    logger.info("Starting")
    assert device in ["cpu", "cuda"] or str(device).startswith("cuda:")
    from ..counting import device_result

    resident = device_result(os.path.dirname(count_matrices_path))
    if resident is not None and os.path.basename(count_matrices_path) == "result.txt":
        q, states, counts = resident
    else:
        q, states, counts = read_count_matrices_array(count_matrices_path)
    mask = read_mask_matrix(mask_path).to_numpy() if mask_path is not None else None
    initialization = read_rate_matrix(initialization_path).to_numpy() if initialization_path is not None else None
    learner = RateMatrixLearner(
        branches=[float(x) for x in q], mats=counts, states=states, output_dir=output_rate_matrix_dir,
        stationnary_distribution=stationary_distribution_path, mask=mask,
        rate_matrix_parameterization=rate_matrix_parameterization, device=device,
        initialization=initialization,
    )
    learner.train(lr=learning_rate, num_epochs=num_epochs, do_adam=do_adam,
                  loss_normalization=loss_normalization, return_best_iter=return_best_iter)
    logger.info("Done!")
    with open(os.path.join(output_rate_matrix_dir, "profiling.txt"), "w") as f:
        f.write(
            f"Total time: {time.time() - start_time} seconds with "
            f"{OPENBLAS_NUM_THREADS} OPENBLAS_NUM_THREADS and {OMP_NUM_THREADS}"
            " OMP_NUM_THREADS\n"
        )
