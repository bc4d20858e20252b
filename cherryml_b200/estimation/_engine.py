"""Device-side fit engine: owns the training state on the GPU and drives the C ABI.

PyTorch allocates the buffers and provides the stream; every computation (Q(theta), the
matrix exponentials, the loss, its gradient, Adam, best-iterate bookkeeping) happens in the
CUDA library.  Nothing is copied to the host until ``results()``.

Mirrors, for the boundary: ``RateMatrix`` (reference estimation/_ratelearn/rate.py:31-188),
``train_quantization`` (trainer.py:118-243) and the per-site batched variant
(_siterm/_cherryml_vectorized.py:107-402).
"""
import ctypes
import math
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .. import _lib


class FitArgs(ctypes.Structure):
    """``cherry_fit_args`` of include/cherryml_b200.h."""

    _fields_ = [
        ("S", ctypes.c_int), ("K", ctypes.c_int), ("n_problems", ctypes.c_int),
        ("t", ctypes.c_void_p), ("C", ctypes.c_void_p), ("mask", ctypes.c_void_p), ("sumC", ctypes.c_void_p),
        ("theta", ctypes.c_void_p), ("adam_m", ctypes.c_void_p), ("adam_v", ctypes.c_void_p),
        ("Q", ctypes.c_void_p), ("Q_best", ctypes.c_void_p), ("Q_last", ctypes.c_void_p),
        ("best_loss", ctypes.c_void_p),
        ("loss_trace", ctypes.c_void_p), ("loss_trace_epochs", ctypes.c_int),
        ("snapshots", ctypes.c_void_p), ("n_snapshots", ctypes.c_int),
        ("dQ_part", ctypes.c_void_p), ("loss_part", ctypes.c_void_p),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("epoch_counter", ctypes.c_void_p), ("status_flag", ctypes.c_void_p),
        ("lr_pi", ctypes.c_double), ("lr_upper", ctypes.c_double), ("beta1", ctypes.c_double),
        ("beta2", ctypes.c_double), ("eps", ctypes.c_double),
        ("do_adam", ctypes.c_int), ("loss_normalization", ctypes.c_int), ("best_mode", ctypes.c_int),
    ]


def solve_stationary_dist(rate_matrix: np.ndarray) -> np.ndarray:
    """Stationary distribution as the reference computes it (rate.py:9-17)."""
    eigvals, eigvecs = np.linalg.eig(rate_matrix.transpose())
    index = np.argmin(np.abs(eigvals.real))
    pi = eigvecs.real[:, index]
    return pi / sum(pi)


def theta_from_initialization(init: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """Invert Q(theta) for an initial rate matrix (reference rate.py:61-91): returns the
    parameter vector ``[log pi, softplus^-1(upper triangle of D^1/2 Q D^-1/2)]``."""
    S = init.shape[0]
    pi = solve_stationary_dist(init)
    if np.any(np.abs(pi) < 1e-8):
        raise ValueError("Stationary distribution of initialization is degenerate.")
    if np.any(np.abs(mask * init - init) > 1e-8):
        raise ValueError("initialization not compatible with mask")
    sym = (np.sqrt(pi)[:, None] * init) / np.sqrt(pi)[None, :]
    iu = np.triu_indices(S, k=1)
    with np.errstate(divide="ignore"):
        upper = np.log(np.exp(sym[iu]) - 1)
    return np.concatenate([np.log(pi), upper])


def random_theta(S: int, seed: int = 0) -> np.ndarray:
    """The reference's no-initialisation start: uniform pi, upper = 0.01 * randn drawn from
    torch's CPU generator right after ``torch.manual_seed(seed)`` (ratelearner.py:77,
    rate.py:51-53), in fp32 like the reference, then widened."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    n_upper = S * (S - 1) // 2
    upper = (0.01 * torch.randn(n_upper, generator=gen, dtype=torch.float32)).double().numpy()
    return np.concatenate([np.log(np.ones(S) / S), upper])


def assign_buckets(times: Sequence[float], world_size: int, rate_scale: float = 1.0) -> list:
    """Partition the K time buckets over ``world_size`` ranks, balanced by estimated cost.

    Cost model: a bucket costs one unit (Taylor evaluation, loss,
    adjoint) plus one unit per squaring, ``s_k = max(0, ceil(log2(t_k * rate_scale / 1.09)))``
    with ``rate_scale`` ~ max |Q_ii| of the starting point.  Longest-processing-time greedy,
    ties broken by bucket index, so every rank computes the same partition.  Returns
    ``world_size`` sorted index arrays."""
    t = np.asarray(times, dtype=np.float64).reshape(-1)
    with np.errstate(divide="ignore"):
        s = np.maximum(0.0, np.ceil(np.log2(np.maximum(t * rate_scale, 1e-300) / 1.09)))
    cost = 1.0 + s
    order = sorted(range(len(t)), key=lambda k: (-cost[k], k))
    load = [0.0] * world_size
    parts = [[] for _ in range(world_size)]
    for k in order:
        r = min(range(world_size), key=lambda i: (load[i], len(parts[i]), i))
        parts[r].append(k)
        load[r] += cost[k]
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


class FitEngine:
    """Training state of ``n_problems`` independent rate-matrix fits on one GPU.

    With ``process_group`` (torch.distributed, one process per GPU) the K buckets are
    sharded over the ranks (``assign_buckets``); theta and the optimiser state are
    replicated and every epoch exchanges ONE all-reduce of ``[dL/dQ | loss]``
    (``cherry_fit_epoch_local`` / ``cherry_fit_epoch_update``)."""

    def __init__(
        self,
        times: np.ndarray,           # [P, K] or [K]
        counts,                      # [P, K, S, S] or [K, S, S]; numpy or a CUDA tensor
        theta0: np.ndarray,          # [P, n_theta] or [n_theta]
        mask: Optional[np.ndarray] = None,
        num_epochs: int = 100,
        learning_rate: float = 0.1,
        lr_upper: Optional[float] = None,
        do_adam: bool = True,
        loss_normalization: bool = True,
        best_mode: int = 0,
        device="cuda",
        betas=(0.9, 0.999),
        eps: float = 1e-8,
        process_group=None,
        rate_scale: float = 1.0,
    ):
        self.lib = _lib.load()
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.CherryError("the fit runs on CUDA devices only (no CPU fallback)")
        self.device = device
        times = np.asarray(times, dtype=np.float64)
        if times.ndim == 1:
            times = times[None, :]
        P, K = times.shape
        if isinstance(counts, torch.Tensor):
            C = counts.to(device=device, dtype=torch.float64)
        else:
            C = torch.from_numpy(np.ascontiguousarray(counts, dtype=np.float64)).to(device)
        if C.dim() == 3:
            C = C[None]
        S = C.shape[-1]
        if tuple(C.shape) != (P, K, S, S):
            raise ValueError(f"counts shape {tuple(C.shape)} does not match times {times.shape}")
        self.process_group = process_group
        self.bucket_index = np.arange(K)
        if process_group is not None:
            import torch.distributed as dist

            world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
            if K < world:
                raise ValueError(f"{K} buckets cannot be sharded over {world} ranks (run replicas instead)")
            if P != 1 and np.any(times != times[0]):
                raise ValueError("bucket sharding needs the same time grid for every problem")
            self.bucket_index = assign_buckets(times[0], world, rate_scale)[rank]
            sel = torch.from_numpy(self.bucket_index).to(device)
            times = times[:, self.bucket_index]
            C = C.index_select(1, sel)
            K = len(self.bucket_index)
        theta0 = np.asarray(theta0, dtype=np.float64)
        if theta0.ndim == 1:
            theta0 = theta0[None, :]
        n_theta = S + S * (S - 1) // 2
        if theta0.shape != (P, n_theta):
            raise ValueError(f"theta0 shape {theta0.shape}, expected {(P, n_theta)}")
        self.S, self.K, self.P, self.num_epochs = S, K, P, int(num_epochs)
        f64 = dict(dtype=torch.float64, device=device)
        self.C = C.contiguous()
        self.t = torch.from_numpy(times).to(device).contiguous()
        self.mask = torch.from_numpy(
            np.ones((S, S)) if mask is None else np.ascontiguousarray(mask, dtype=np.float64)
        ).to(device)
        self.sumC = self.C.sum(dim=(1, 2, 3)).contiguous()
        if process_group is not None:
            dist.all_reduce(self.sumC, op=dist.ReduceOp.SUM, group=process_group)
        self.packed = torch.zeros(P * S * S + P, **f64) if process_group is not None else None
        self.theta = torch.from_numpy(theta0).to(device).contiguous()
        self.adam_m = torch.zeros_like(self.theta)
        self.adam_v = torch.zeros_like(self.theta)
        self.Q = torch.zeros((P, S, S), **f64)
        self.Q_best = torch.zeros((P, S, S), **f64)
        self.Q_last = torch.zeros((P, S, S), **f64)
        self.best_loss = torch.full((P,), float("inf"), **f64)
        self.loss_trace = torch.full((max(1, self.num_epochs), P), float("nan"), **f64)
        self.n_snapshots = (int(math.floor(math.log2(self.num_epochs))) + 1) if (self.num_epochs >= 1 and P == 1) else 0
        self.snapshots = torch.zeros((max(1, self.n_snapshots), S, S), **f64)
        # S <= 32: one gradient piece per bucket (reduced in fixed order by the update kernel);
        # larger S: the large path reduces internally and leaves the total in dQ_part[0]
        self.dQ_part = torch.zeros((P * K if S <= 32 else P, S, S), **f64)
        self.loss_part = torch.zeros((P * K,), **f64)
        nbytes = ctypes.c_size_t(0)
        _lib.check(self.lib.cherry_fit_workspace_bytes(S, K, P, ctypes.byref(nbytes)), "cherry_fit_workspace_bytes")
        self.workspace = torch.empty(max(8, nbytes.value), dtype=torch.uint8, device=device)
        self.epoch_counter = torch.zeros((P,), dtype=torch.int32, device=device)
        self.status_flag = torch.zeros((1,), dtype=torch.int32, device=device)
        a = FitArgs()
        a.S, a.K, a.n_problems = S, K, P
        a.t, a.C, a.mask, a.sumC = (_lib.ptr(x) for x in (self.t, self.C, self.mask, self.sumC))
        a.theta, a.adam_m, a.adam_v = (_lib.ptr(x) for x in (self.theta, self.adam_m, self.adam_v))
        a.Q, a.Q_best, a.Q_last, a.best_loss = (
            _lib.ptr(x) for x in (self.Q, self.Q_best, self.Q_last, self.best_loss))
        a.loss_trace, a.loss_trace_epochs = _lib.ptr(self.loss_trace), max(1, self.num_epochs)
        a.snapshots, a.n_snapshots = (_lib.ptr(self.snapshots) if self.n_snapshots else None), self.n_snapshots
        a.dQ_part, a.loss_part = _lib.ptr(self.dQ_part), _lib.ptr(self.loss_part)
        a.workspace, a.workspace_bytes = _lib.ptr(self.workspace), nbytes.value
        a.epoch_counter, a.status_flag = _lib.ptr(self.epoch_counter), _lib.ptr(self.status_flag)
        a.lr_pi = float(learning_rate)
        a.lr_upper = float(learning_rate if lr_upper is None else lr_upper)
        a.beta1, a.beta2, a.eps = float(betas[0]), float(betas[1]), float(eps)
        a.do_adam, a.loss_normalization, a.best_mode = int(do_adam), int(loss_normalization), int(best_mode)
        self.args = a
        self.epochs_done = 0
        with torch.cuda.device(device):
            _lib.check(self.lib.cherry_fit_init(ctypes.byref(a), _lib.current_stream_ptr()), "cherry_fit_init")
            if best_mode == 1:
                self.Q_best.copy_(self.Q)  # the per-site variant starts from the initial Q

    def run(self, num_epochs: Optional[int] = None) -> None:
        """Enqueue epochs on the current stream (no host synchronisation)."""
        n = self.num_epochs - self.epochs_done if num_epochs is None else int(num_epochs)
        if n <= 0:
            return
        if self.process_group is not None:
            return self._run_sharded(n)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream()
            if stream.cuda_stream == 0:
                # the legacy default stream cannot be captured into a graph: use a side stream
                side = torch.cuda.Stream()
                side.wait_stream(stream)
                with torch.cuda.stream(side):
                    rc = self.lib.cherry_fit_run(ctypes.byref(self.args), n, side.cuda_stream)
                stream.wait_stream(side)
            else:
                rc = self.lib.cherry_fit_run(ctypes.byref(self.args), n, stream.cuda_stream)
        _lib.check(rc, "cherry_fit_run")
        self.epochs_done += n

    def _run_sharded(self, n: int) -> None:
        import torch.distributed as dist

        with torch.cuda.device(self.device):
            a, packed = ctypes.byref(self.args), _lib.ptr(self.packed)
            for _ in range(n):
                st = _lib.current_stream_ptr()
                _lib.check(self.lib.cherry_fit_epoch_local(a, packed, st), "cherry_fit_epoch_local")
                dist.all_reduce(self.packed, op=dist.ReduceOp.SUM, group=self.process_group)
                _lib.check(self.lib.cherry_fit_epoch_update(a, packed, st), "cherry_fit_epoch_update")
        self.epochs_done += n

    @property
    def symmetric_form(self) -> bool:
        """True when the epochs of this (large-S) fit run on the symmetric form of the reversible model
        (include/cherryml_b200.h cherry_fit_symmetric_form)."""
        return bool(self.lib.cherry_fit_symmetric_form(ctypes.byref(self.args)))

    def loss_and_grad(self):
        """(loss [P], dL/dQ [P,S,S]) at the current Q, normalised like the training loss."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.cherry_fit_loss_grad(ctypes.byref(self.args), _lib.current_stream_ptr()),
                       "cherry_fit_loss_grad")
        scale = (1.0 / self.sumC) if self.args.loss_normalization else torch.ones_like(self.sumC)
        if self.S <= 32:
            loss = self.loss_part.view(self.P, self.K).sum(dim=1) * scale
            grad = self.dQ_part.view(self.P, self.K, self.S, self.S).sum(dim=1) * scale.view(-1, 1, 1)
        else:
            loss = self.loss_part.view(self.P, self.K).sum(dim=1) * scale
            grad = self.dQ_part[: self.P] * scale.view(-1, 1, 1)
        self.check_status()
        return loss, grad

    def check_status(self) -> None:
        if int(self.status_flag.item()) != 0:
            raise _lib.CherryError(f"fit kernel reported status {int(self.status_flag.item())} (1/2: rate matrix norm beyond the supported range, 3: internal scheduling timeout)")

    def results(self) -> Dict[str, np.ndarray]:
        """Synchronise and fetch: loss trace, Q snapshots (single problem), best and last Q."""
        torch.cuda.synchronize(self.device)
        self.check_status()
        out: Dict[str, np.ndarray] = {}
        n = self.epochs_done
        trace = self.loss_trace[:n].cpu().numpy()
        out["loss_per_problem"] = trace
        out["loss"] = trace.sum(axis=1) if self.P > 1 else trace[:, 0]
        if self.P == 1:
            snaps = self.snapshots.cpu().numpy()
            for j in range(self.n_snapshots):
                if (1 << j) <= n:
                    out[f"Q_{1 << j}"] = snaps[j].copy()
            out["Q_best"] = self.Q_best[0].cpu().numpy()
            out["Q_last"] = self.Q_last[0].cpu().numpy()
        else:
            out["Q_best"] = self.Q_best.cpu().numpy()
            out["Q_last"] = self.Q_last.cpu().numpy()
        out["best_loss"] = self.best_loss.cpu().numpy()
        return out

