"""``count_transitions`` / ``count_co_transitions``: drop-in stage functions.

Same names, keyword arguments, defaults, output files and caching behaviour as the
reference's ``cherryml/counting/_count_transitions.py:201-379`` and
``_count_co_transitions.py:227-412``; the work itself is the sm_100a counting kernels.

``use_cpp_implementation`` keeps its meaning as the choice of numerical "personality":
``True`` (default) reproduces the C++ binary (branch lengths rounded through float32 as
by ``std::stof``, ``_count_transitions.cpp:247``; ``result.txt`` in the C++ writer's
layout with 6 significant digits, ``.cpp:524-548``), ``False`` reproduces the Python
implementation (fp64 branch lengths, pandas-style ``result.txt``).  Both run on the GPU.
``num_processes`` sizes the host ingest thread pool; the ``cpp_command_line_*`` arguments
are accepted and ignored.
Extra keyword arguments (all excluded from the cache key): ``device``,
``process_group`` (torch.distributed group: families are striped over ranks exactly like
the reference stripes them over MPI ranks, ``.cpp:624-629``, and the raw integer
histograms are all-reduced), ``result_style`` (override the writer layout), ``ingest``
(``"native"``: the library's multithreaded C++ parser/encoder with ``num_processes`` threads,
default; ``"python"``: the pure-Python one, same output).
"""
import logging
import os
import time
from typing import Dict, List, Optional, Tuple, Union

from collections import OrderedDict

import numpy as np
import torch

from .. import caching
from ..io import write_count_matrices_array
from ..utils import get_process_args
from ._device import count_batch, count_batches_streamed
from ._ingest import build_co_batch, build_co_batch_native, build_lg_batch, build_lg_batch_native

logger = logging.getLogger(__name__)

# In-process hand-off to the fit: output dir -> (grid fp64 [K], states, counts fp64 device tensor, stamp of
# result.txt).  At most _MAX_DEVICE_RESULTS entries are kept (a co-transition tensor is 165 MB of HBM), the
# least recently used one is dropped first, and an entry is only handed out while result.txt on disk is still
# the file this process wrote (size and mtime): a replaced file wins over the resident tensor.
_MAX_DEVICE_RESULTS = 2
_DEVICE_RESULTS: "OrderedDict[str, Tuple[np.ndarray, List[str], torch.Tensor, Tuple[int, int]]]" = OrderedDict()


def _stamp(path: str):
    try:
        st = os.stat(path)
    except OSError:
        return None
    return (st.st_size, st.st_mtime_ns)


def device_result(output_count_matrices_dir: str):
    """Counts of a ``count_*`` call made in this process, still resident on the device: ``(grid, states,
    counts)``, or None if there are none or ``result.txt`` in that directory is no longer the file this
    process wrote."""
    key = os.path.realpath(output_count_matrices_dir)
    entry = _DEVICE_RESULTS.get(key)
    if entry is None:
        return None
    if entry[3] is None or _stamp(os.path.join(key, "result.txt")) != entry[3]:
        del _DEVICE_RESULTS[key]
        return None
    _DEVICE_RESULTS.move_to_end(key)
    return entry[:3]


def clear_device_results() -> None:
    """Drop every resident count tensor (frees the HBM they hold)."""
    _DEVICE_RESULTS.clear()


def round_like_the_cpp_writer(counts: torch.Tensor) -> torch.Tensor:
    """The values a reader gets back from a ``result.txt`` written in the C++ program's format (``%g``: six
    significant digits, reference counting/_count_transitions.cpp:560-577): what the reference's fit is trained
    on when counting ran with ``use_cpp_implementation=True``.  Counts are multiples of 0.25, so the decimal
    rounding (round-half-even on the exact value, as printf does) is done exactly in integers."""
    n = torch.round(counts * 4.0).to(torch.int64)  # exact: counts are multiples of 0.25
    out = counts.clone()
    # 1e4 <= v < 1e5: one decimal survives (v * 10 = 5 n / 2)
    sel = (counts >= 1e4) & (counts < 1e5)
    if bool(sel.any()):
        m = 5 * n[sel]
        q, r = torch.div(m, 2, rounding_mode="floor"), m % 2
        q = q + ((r == 1) & (q % 2 == 1)).to(torch.int64)  # r == 1 is an exact tie
        out[sel] = q.to(torch.float64) / 10.0
    # 10^e <= v < 10^(e+1), e >= 5: multiples of D = 10^(e-5) survive (v / D = n / (4 D))
    lo, d = 1e5, 1
    while bool((counts >= lo).any()):
        sel = (counts >= lo) & (counts < lo * 10)
        if bool(sel.any()):
            q, r = torch.div(n[sel], 4 * d, rounding_mode="floor"), n[sel] % (4 * d)
            up = (2 * r > 4 * d) | ((2 * r == 4 * d) & (q % 2 == 1))
            out[sel] = ((q + up.to(torch.int64)) * d).to(torch.float64)
        lo, d = lo * 10, d * 10
    return out


def _rank_world(process_group):
    if process_group is None:
        return 0, 1
    import torch.distributed as dist

    return dist.get_rank(process_group), dist.get_world_size(process_group)


def _ingest_threads(num_processes) -> int:
    """``num_processes`` is the reference's degree of host parallelism for counting (MPI ranks);
    here it sizes the ingest thread pool, capped by the host's cores."""
    try:
        n = int(num_processes)
    except (TypeError, ValueError):
        n = 1
    return max(1, min(n, os.cpu_count() or 1))


def _finish(counts_dev, grid, states, out_dir, style, start_time, num_processes, rank, process_group=None):
    if rank == 0:
        counts = counts_dev.cpu().numpy()
        write_count_matrices_array(
            grid.tolist(), states, counts, os.path.join(out_dir, "result.txt"), style
        )
        with open(os.path.join(out_dir, "profiling.txt"), "w") as f:
            f.write(
                f"Total time: {time.time() - start_time} seconds with "
                f"{num_processes} processes.\n"
            )
    if process_group is not None:
        import torch.distributed as dist

        dist.barrier(process_group)  # result.txt exists before any rank returns
    # hand-off to the fit: with the C++ writer's format the file holds six significant digits, and the
    # reference trains on the file, so the resident copy is rounded the same way
    key = os.path.realpath(out_dir)
    resident, grid_res = counts_dev, grid
    if style == "cpp":  # the quantization points go through "%g" as well
        resident = round_like_the_cpp_writer(counts_dev)
        grid_res = np.array([float("%g" % x) for x in grid])
    _DEVICE_RESULTS[key] = (grid_res, list(states), resident, _stamp(os.path.join(key, "result.txt")))
    _DEVICE_RESULTS.move_to_end(key)
    while len(_DEVICE_RESULTS) > _MAX_DEVICE_RESULTS:
        _DEVICE_RESULTS.popitem(last=False)


@caching.cached_computation(
    exclude_args=[
        "num_processes",
        "use_cpp_implementation",
        "cpp_command_line_prefix",
        "cpp_command_line_suffix",
        "device",
        "process_group",
        "result_style",
        "ingest",
        "families_per_batch",
    ],
    output_dirs=["output_count_matrices_dir"],
    write_extra_log_files=True,
)
def count_transitions(
    tree_dir: str,
    msa_dir: str,
    site_rates_dir: str,
    families: List[str],
    amino_acids: List[str],
    quantization_points: List[Union[str, float]],
    edge_or_cherry: str,
    output_count_matrices_dir: Optional[str] = None,
    num_processes: int = 1,
    use_cpp_implementation: bool = True,
    cpp_command_line_prefix: str = "",
    cpp_command_line_suffix: str = "",
    device: str = "cuda",
    process_group=None,
    result_style: Optional[str] = None,
    ingest: str = "native",
    families_per_batch: int = 1024,
) -> None:
    """Count single-site transitions into a ``K x S x S`` tensor (see module docstring).  More than
    ``families_per_batch`` families are streamed: batches are parsed and encoded by the library's
    host threads into pooled page-locked buffers while the previous batch is uploaded and counted."""
    if edge_or_cherry.startswith("cherry++__"):
        edge_or_cherry = "cherry++"
    start_time = time.time()
    logger.info(f"Starting on {len(families)} families")
    os.makedirs(output_count_matrices_dir, exist_ok=True)
    quantization_points = [float(q) for q in quantization_points]
    rank, world = _rank_world(process_group)
    my_families = get_process_args(rank, world, list(families))
    # In-memory route (reference _cherry.py:279-336): the FastCherries stage of this process has just estimated
    # the cherries of exactly these families from exactly this MSA directory and still holds the encoded
    # residues and its results on the device -- count on them instead of parsing the files it wrote.
    # Same counts, bit for bit (tests/test_gpu_fast_cherries.py); cherries only, one process.
    if process_group is None and edge_or_cherry in ("cherry", "cherry++") and ingest == "native":
        from ..phylogeny_estimation import _fast_cherries as _fc

        hand = _fc.take_handoff(tree_dir, site_rates_dir, msa_dir, list(families), list(amino_acids))
        if hand is not None:
            from ..phylogeny_estimation._pipeline import lg_batch_from_fast_cherries
            from ._device import count_raw, sorted_grid, symmetrize

            S = len(amino_acids)
            dev_batch = lg_batch_from_fast_cherries(hand["fams"], hand["out"], hand["grid"], hand["cats"], S,
                                                    bool(use_cpp_implementation), hand["device"],
                                                    n_threads=_ingest_threads(num_processes))
            grid = sorted_grid(quantization_points)
            with torch.cuda.device(dev_batch.msa.device):
                grid_dev = torch.from_numpy(grid).to(dev_batch.msa.device)
                counts = symmetrize(count_raw(dev_batch, grid_dev, int(grid.size), S), "lg", int(grid.size), S, False)
            del hand, dev_batch
            style = result_style or ("cpp" if use_cpp_implementation else "python")
            _finish(counts, np.array(sorted(quantization_points)), list(amino_acids),
                    output_count_matrices_dir, style, start_time, num_processes, rank, process_group)
            logger.info("Done! (counted on the resident FastCherries results)")
            return
    if ingest == "native" and len(my_families) > families_per_batch > 0:
        def build(fams):
            return build_lg_batch_native(
                tree_dir, msa_dir, site_rates_dir, fams, amino_acids, edge_or_cherry,
                float32_branch_lengths=bool(use_cpp_implementation), n_threads=_ingest_threads(num_processes),
                pinned=True)

        chunks = [my_families[i: i + families_per_batch] for i in range(0, len(my_families), families_per_batch)]
        counts = count_batches_streamed(build, chunks, "lg", quantization_points, len(amino_acids),
                                        directed=(edge_or_cherry == "edge"), device=device,
                                        process_group=process_group)
        style = result_style or ("cpp" if use_cpp_implementation else "python")
        _finish(counts, np.array(sorted(quantization_points)), list(amino_acids),
                output_count_matrices_dir, style, start_time, num_processes, rank, process_group)
        logger.info("Done!")
        return
    if ingest == "native":
        batch = build_lg_batch_native(
            tree_dir, msa_dir, site_rates_dir, my_families, amino_acids, edge_or_cherry,
            float32_branch_lengths=bool(use_cpp_implementation), n_threads=_ingest_threads(num_processes),
        )
    elif ingest == "python":
        batch = build_lg_batch(
            tree_dir, msa_dir, site_rates_dir, my_families, amino_acids, edge_or_cherry,
            float32_branch_lengths=bool(use_cpp_implementation),
        )
    else:
        raise ValueError(f"Unknown ingest: {ingest!r}")
    counts = count_batch(
        batch, quantization_points, len(amino_acids), directed=(edge_or_cherry == "edge"),
        device=device, process_group=process_group,
    )
    style = result_style or ("cpp" if use_cpp_implementation else "python")
    _finish(counts, np.array(sorted(quantization_points)), list(amino_acids),
            output_count_matrices_dir, style, start_time, num_processes, rank, process_group)
    logger.info("Done!")


@caching.cached_computation(
    exclude_args=[
        "num_processes",
        "use_cpp_implementation",
        "cpp_command_line_prefix",
        "cpp_command_line_suffix",
        "device",
        "process_group",
        "result_style",
        "ingest",
        "families_per_batch",
    ],
    output_dirs=["output_count_matrices_dir"],
    write_extra_log_files=True,
)
def count_co_transitions(
    tree_dir: str,
    msa_dir: str,
    contact_map_dir: str,
    families: List[str],
    amino_acids: List[str],
    quantization_points: List[Union[str, float]],
    edge_or_cherry: str,
    minimum_distance_for_nontrivial_contact: int,
    output_count_matrices_dir: Optional[str] = None,
    num_processes: int = 1,
    use_cpp_implementation: bool = True,
    cpp_command_line_prefix: str = "",
    cpp_command_line_suffix: str = "",
    device: str = "cuda",
    process_group=None,
    result_style: Optional[str] = None,
    ingest: str = "native",
    families_per_batch: int = 1024,
) -> None:
    """Count transitions of contacting site pairs into a ``K x S^2 x S^2`` tensor (streamed in
    batches of ``families_per_batch`` families like ``count_transitions``)."""
    if edge_or_cherry.startswith("cherry++__"):
        edge_or_cherry = "cherry++"
    start_time = time.time()
    os.makedirs(output_count_matrices_dir, exist_ok=True)
    quantization_points = [float(q) for q in quantization_points]
    rank, world = _rank_world(process_group)
    my_families = get_process_args(rank, world, list(families))
    if ingest == "native" and len(my_families) > families_per_batch > 0:
        def build(fams):
            return build_co_batch_native(
                tree_dir, msa_dir, contact_map_dir, fams, amino_acids, edge_or_cherry,
                minimum_distance_for_nontrivial_contact, float32_branch_lengths=bool(use_cpp_implementation),
                n_threads=_ingest_threads(num_processes), pinned=True)

        chunks = [my_families[i: i + families_per_batch] for i in range(0, len(my_families), families_per_batch)]
        counts = count_batches_streamed(build, chunks, "co", quantization_points, len(amino_acids),
                                        directed=(edge_or_cherry == "edge"), device=device,
                                        process_group=process_group)
        pair_states = [a + b for a in amino_acids for b in amino_acids]
        style = result_style or ("cpp" if use_cpp_implementation else "python")
        _finish(counts, np.array(sorted(quantization_points)), pair_states,
                output_count_matrices_dir, style, start_time, num_processes, rank, process_group)
        logger.info("Done!")
        return
    if ingest == "native":
        batch = build_co_batch_native(
            tree_dir, msa_dir, contact_map_dir, my_families, amino_acids, edge_or_cherry,
            minimum_distance_for_nontrivial_contact, float32_branch_lengths=bool(use_cpp_implementation),
            n_threads=_ingest_threads(num_processes),
        )
    elif ingest == "python":
        batch = build_co_batch(
            tree_dir, msa_dir, contact_map_dir, my_families, amino_acids, edge_or_cherry,
            minimum_distance_for_nontrivial_contact, float32_branch_lengths=bool(use_cpp_implementation),
        )
    else:
        raise ValueError(f"Unknown ingest: {ingest!r}")
    counts = count_batch(
        batch, quantization_points, len(amino_acids), directed=(edge_or_cherry == "edge"),
        device=device, process_group=process_group,
    )
    pair_states = [a + b for a in amino_acids for b in amino_acids]
    style = result_style or ("cpp" if use_cpp_implementation else "python")
    _finish(counts, np.array(sorted(quantization_points)), pair_states,
            output_count_matrices_dir, style, start_time, num_processes, rank, process_group)
    logger.info("Done!")
