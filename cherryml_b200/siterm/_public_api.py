"""SiteRM public API: ``learn_site_specific_rate_matrices`` / ``learn_site_rate_matrices`` with
the reference's arguments, defaults and returned dictionaries
(``cherryml/_siterm_public_api.py:21-171``, ``_siterm/_learn_site_rate_matrix.py:1109-1281``).

With ``tree=None`` the tree and the site rates come from FastCherries on the GPU
(``cherryml_b200.phylogeny_estimation.fast_cherries``, 20 rate categories, 50 iterations, as
the reference calls it at :1196-1225).  With a tree, site rates are the maximisers over
``site_rate_grid`` of ``log prior + sum over cherries (both directions) of log expm(rate * t *
Q)[x, y]`` (:387-533; the reference's Cython ``compute_optimal_site_rates``), evaluated here as
one gather-and-sum over the device table of log transition probabilities.  The rate matrices
are then learnt by ``estimate_site_specific_rate_matrices_given_tree_and_site_rates``.
"""
import os
import tempfile
import time
from typing import Dict, List, Optional

import numpy as np
import torch

from ..io import Tree, read_site_rates, read_tree, write_msa, write_rate_matrix
from ..markov_chain import expm_batched
from ..phylogeny_estimation import fast_cherries
from ._site_specific import estimate_site_specific_rate_matrices_given_tree_and_site_rates, get_cherry_transitions

QUANTIZATION_GRID_CENTER = 0.03
QUANTIZATION_GRID_STEP = 1.1
QUANTIZATION_GRID_NUM_STEPS = 64


def get_standard_site_rate_grid(num_site_rates: int = 20) -> List[float]:
    """Geometric grid from 1/n to n (reference ``_learn_site_rate_matrix.py:933-941``)."""
    n = num_site_rates
    return [n ** (-1.0 + 2.0 * (n - i) / (n - 1.0)) for i in range(1, n + 1)][::-1]


def get_standard_site_rate_prior(num_site_rates: int = 20) -> List[float]:
    """Density of gamma(shape 3, scale 1/3) on the standard grid (:944-952): 13.5 r^2 e^{-3r}."""
    from scipy.stats import gamma

    return [float(gamma.pdf(r, a=3.0, scale=1.0 / 3.0)) for r in get_standard_site_rate_grid(num_site_rates)]


def estimate_site_rates(tree: Tree, leaf_states: Dict[str, str], site_rate_grid: List[float],
                        site_rate_prior: List[float], rate_matrix, device="cuda") -> List[float]:
    """Per-site maximiser of the cherries' composite likelihood over the rate grid
    (``_estimate_site_rates`` / ``_estimate_site_rates_fast``).  Ties go to the larger rate, like
    the reference's ``sorted(...)[-1]``."""
    if len(site_rate_grid) != len(site_rate_prior):
        raise ValueError(
            "site_rate_grid and site_rate_prior should have the same length. "
            f"You provided: site_rate_grid='{site_rate_grid}' and site_rate_prior='{site_rate_prior}'."
        )
    states = list(rate_matrix.columns)
    S = len(states)
    num_sites = len(next(iter(leaf_states.values())))
    if len(site_rate_grid) == 1:
        return [site_rate_grid[0]] * num_sites
    for seq in leaf_states.values():
        low = [ch for ch in seq if ch.islower()]
        if low:
            raise ValueError(f"Lowercase state found: '{low[0]}' . Did you forget to make it uppercase?")
    cherries = get_cherry_transitions(tree, leaf_states)
    lut = np.full(256, S, dtype=np.int64)
    for i, ch in enumerate(states):
        lut[ord(ch)] = i
    enc = lambda s: lut[np.frombuffer(s.encode("latin-1"), dtype=np.uint8)]  # noqa: E731
    xa = np.stack([enc(x) for x, _, _ in cherries])  # [C, L]
    xb = np.stack([enc(y) for _, y, _ in cherries])
    t = np.array([tt for _, _, tt in cherries], dtype=np.float64)
    R, C = len(site_rate_grid), len(cherries)
    exponents = (np.array(site_rate_grid, dtype=np.float64)[:, None] * t[None, :]).reshape(-1)
    dev = torch.device(device if str(device).startswith("cuda") else "cuda")
    logp = torch.log(expm_batched(rate_matrix.to_numpy(dtype=np.float64), exponents, dev)).reshape(R, C, S, S)
    padded = torch.zeros((R, C, S + 1, S + 1), dtype=torch.float64, device=dev)  # unknown residue: contributes 0
    padded[:, :, :S, :S] = logp
    a = torch.from_numpy(xa).to(dev)
    b = torch.from_numpy(xb).to(dev)
    flat = padded.reshape(R, C, (S + 1) * (S + 1))
    fwd = torch.gather(flat, 2, (a * (S + 1) + b)[None].expand(R, C, num_sites))
    bwd = torch.gather(flat, 2, (b * (S + 1) + a)[None].expand(R, C, num_sites))
    # the reference walks cherries then reversed cherries, adding one term at a time
    ll = torch.cumsum(torch.cat([fwd, bwd], dim=1), dim=1)[:, -1, :]  # [R, L]
    ll = ll + torch.log(torch.tensor(site_rate_prior, dtype=torch.float64, device=dev))[:, None]
    ll_h = ll.cpu().numpy()
    grid = np.array(site_rate_grid, dtype=np.float64)
    out = []
    for site in range(num_sites):
        best = max(range(R), key=lambda r: (ll_h[r, site], grid[r]))
        out.append(float(grid[best]))
    return out


def learn_site_rate_matrices(
    tree: Optional[Tree],
    leaf_states: Dict[str, str],
    alphabet: List[str],
    regularization_rate_matrix,
    regularization_strength: float,
    use_vectorized_implementation: bool,
    vectorized_implementation_device: str = "cpu",
    vectorized_implementation_num_cores: int = 1,
    site_rate_grid: List[float] = [2.0 ** i for i in range(-10, 10)],
    site_rate_prior: List[float] = [1.0 for i in range(-10, 10)],
    alphabet_for_site_rate_estimation: Optional[List[str]] = None,
    rate_matrix_for_site_rate_estimation=None,
    num_epochs: int = 100,
    use_fast_site_rate_implementation: bool = False,
    quantization_grid_num_steps: int = QUANTIZATION_GRID_NUM_STEPS,
    just_run_fast_cherries: bool = False,
) -> Dict:
    prof: Dict = {}
    st = time.time()
    if alphabet_for_site_rate_estimation is None:
        alphabet_for_site_rate_estimation = alphabet[:]
    if rate_matrix_for_site_rate_estimation is None:
        rate_matrix_for_site_rate_estimation = regularization_rate_matrix.copy()
    assert list(rate_matrix_for_site_rate_estimation.columns) == alphabet_for_site_rate_estimation
    assert list(regularization_rate_matrix.columns) == alphabet
    device = vectorized_implementation_device
    device = device if str(device).startswith("cuda") else "cuda"
    prof["time_init_learn_site_rate_matrices"] = time.time() - st

    st = time.time()
    site_rates_fast_cherries = None
    if tree is None:
        with tempfile.TemporaryDirectory() as tmp:
            rm_path = os.path.join(tmp, "rate_matrix.txt")
            write_rate_matrix(rate_matrix_for_site_rate_estimation.to_numpy(),
                              list(rate_matrix_for_site_rate_estimation.columns), rm_path)
            msa_dir = os.path.join(tmp, "msa_dir")
            os.makedirs(msa_dir)
            write_msa(leaf_states, os.path.join(msa_dir, "family_0.txt"))
            fast_cherries(
                msa_dir=msa_dir, families=["family_0"], rate_matrix_path=rm_path, num_rate_categories=20,
                max_iters=50, num_processes=1, verbose=False, output_tree_dir=os.path.join(tmp, "tree"),
                output_site_rates_dir=os.path.join(tmp, "site_rates"),
                output_likelihood_dir=os.path.join(tmp, "lls"), device=device,
            )
            tree = read_tree(os.path.join(tmp, "tree", "family_0.txt"))
            site_rates_fast_cherries = read_site_rates(os.path.join(tmp, "site_rates", "family_0.txt"))
    elif just_run_fast_cherries:
        raise ValueError("If just_run_fast_cherries is True, then tree must be None.")
    time_estimate_tree = time.time() - st

    st = time.time()
    if site_rates_fast_cherries is not None:
        site_rates = site_rates_fast_cherries
    else:
        site_rates = estimate_site_rates(tree, leaf_states, site_rate_grid[:], site_rate_prior[:],
                                         rate_matrix_for_site_rate_estimation, device)
    time_estimate_site_rate = time.time() - st

    learnt, learnt_prof = None, {}
    if not just_run_fast_cherries:
        st = time.time()
        step = QUANTIZATION_GRID_STEP ** (QUANTIZATION_GRID_NUM_STEPS / quantization_grid_num_steps)
        points = [QUANTIZATION_GRID_CENTER * step ** i
                  for i in range(-quantization_grid_num_steps, quantization_grid_num_steps + 1, 1)]
        learnt_prof["time_build_quantization_points"] = time.time() - st
        r = estimate_site_specific_rate_matrices_given_tree_and_site_rates(
            tree=tree, site_rates=site_rates, msa=leaf_states, alphabet=alphabet,
            regularization_strength=regularization_strength,
            regularization_rate_matrix=regularization_rate_matrix.to_numpy(), quantization_points=points,
            optimization_num_epochs=num_epochs, transitions_strategy="cherry++", include_reverse_transitions=True,
            rate_matrix_parameterization="pande_reversible", log_dir=None, plot_site_specific_rate_matrices=0,
            use_vectorized_cherryml_implementation=use_vectorized_implementation,
            vectorized_cherryml_implementation_device=device,
            vectorized_cherryml_implementation_num_cores=vectorized_implementation_num_cores,
        )
        learnt = r["res"]
        learnt_prof.update({k: v for k, v in r.items() if k.startswith("time_")})
        learnt_prof["time_build_pandas_return"] = 0.0
    return {
        "learnt_rate_matrices": learnt,
        "learnt_site_rates": site_rates,
        "learnt_tree": tree,
        "time_estimate_tree": time_estimate_tree,
        "time_estimate_site_rate": time_estimate_site_rate,
        **prof,
        **learnt_prof,
    }


def learn_site_specific_rate_matrices(
    tree: Optional[Tree],
    msa: Dict[str, str],
    alphabet: List[str],
    regularization_rate_matrix,
    regularization_strength: float = 0.5,
    device: str = "cpu",
    num_rate_categories: int = 20,
    alphabet_for_site_rate_estimation: Optional[List[str]] = None,
    rate_matrix_for_site_rate_estimation=None,
    num_epochs: int = 100,
    quantization_grid_num_steps: int = 64,
    use_vectorized_implementation: bool = True,
    just_run_fast_cherries: bool = False,
) -> Dict:
    """Learn a rate matrix per site of an MSA (SiteRM).  ``device`` is accepted for compatibility;
    everything runs on a CUDA device."""
    return learn_site_rate_matrices(
        tree=tree, leaf_states=msa, alphabet=alphabet, regularization_rate_matrix=regularization_rate_matrix,
        regularization_strength=regularization_strength, use_vectorized_implementation=use_vectorized_implementation,
        vectorized_implementation_device=device, vectorized_implementation_num_cores=1,
        site_rate_grid=get_standard_site_rate_grid(num_site_rates=num_rate_categories),
        site_rate_prior=get_standard_site_rate_prior(num_site_rates=num_rate_categories),
        alphabet_for_site_rate_estimation=alphabet_for_site_rate_estimation,
        rate_matrix_for_site_rate_estimation=rate_matrix_for_site_rate_estimation, num_epochs=num_epochs,
        use_fast_site_rate_implementation=True, quantization_grid_num_steps=quantization_grid_num_steps,
        just_run_fast_cherries=just_run_fast_cherries,
    )
