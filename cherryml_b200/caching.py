"""Filesystem memoisation with the calling convention of the reference's
``cherryml/caching`` (``_cached_computation.py:150-369``, key derivation
``_common.py:114-180``).

Only what decorates the hot-path stage functions is provided: ``cached_computation``.
Semantics kept: keyword-only calls (``CacheUsageError`` otherwise); with no cache dir the
wrapped function is called as is and its ``None`` is returned; with a cache dir the output
dirs default to ``<cache>/<func name>/<sha512 key>/<output_dir name>``, the call is skipped
when ``result.txt`` + ``result.success`` exist, ``result.txt`` is made read-only afterwards,
and the dict ``{output_dir_name: path}`` is returned.  The key is computed from the same
strings in the same way, so a cache directory written by the reference is found again;
arguments that this implementation adds (device lists, compat switches) are listed in
``exclude_args`` by the stage functions and therefore do not change the key.
"""
import hashlib
import logging
import os
import stat
from functools import wraps
from inspect import signature
from typing import List, Optional

logger = logging.getLogger(__name__)

_CACHE_DIR: Optional[str] = None
_USE_HASH = True
_HASH_LEN = 64  # the reference's package __init__ sets 64 / 0 levels (cherryml/__init__.py:18-19)
_DIR_LEVELS = 0
_READ_ONLY = False


class CacheUsageError(Exception):
    pass


def set_cache_dir(cache_dir: Optional[str]) -> None:
    global _CACHE_DIR
    _CACHE_DIR = cache_dir


def get_cache_dir() -> Optional[str]:
    return _CACHE_DIR


def set_use_hash(use_hash: bool) -> None:
    global _USE_HASH
    _USE_HASH = use_hash


def set_hash_len(hash_len: int) -> None:
    if hash_len > 128:
        raise ValueError(f"The maximum allowed hash length is 128. You requested: {hash_len}")
    global _HASH_LEN
    _HASH_LEN = hash_len


def set_dir_levels(dir_levels: int) -> None:
    global _DIR_LEVELS
    _DIR_LEVELS = dir_levels


def set_read_only(read_only: bool) -> None:
    global _READ_ONLY
    _READ_ONLY = read_only


def set_log_level(log_level: int) -> None:
    logger.setLevel(log_level)


def _hash_all(xs: List[str]) -> str:
    inner = "".join(hashlib.sha512(x.encode("utf-8")).hexdigest() for x in xs)
    res = hashlib.sha512(inner.encode("utf-8")).hexdigest()[:_HASH_LEN]
    if _DIR_LEVELS:
        res = "/".join(res[:_DIR_LEVELS]) + "/" + res[_DIR_LEVELS:]
    return res


def _caching_dir(func, unhashed: List[str], kwargs, cache_dir: str, use_hash: Optional[bool] = None) -> str:
    binding = signature(func).bind(**kwargs)
    binding.apply_defaults()
    items = [(k, v) for k, v in binding.arguments.items() if k not in unhashed]
    if _USE_HASH if use_hash is None else use_hash:
        key = _hash_all(sum(([f"{k}", f"{v}"] for k, v in items), []))
        return os.path.join(cache_dir, func.__name__, key)
    return os.path.join(cache_dir, func.__name__, *[f"{k}_{v}" for k, v in items])


def _call_caching_dir(func, exclude_args, exclude_args_if_default, output_dirs, parallel_arg, kwargs,
                      cache_dir: str, use_hash: Optional[bool] = None):
    """Cache directory of one call of a decorated function and the arguments left out of its key:
    the decorator's ``exclude_args``, the parallel argument, the output directories, and every
    ``exclude_args_if_default`` argument that has its default value (reference
    ``_cached_computation.py:37-82``, ``_cached_parallel_computation.py:45-91``)."""
    params = signature(func).parameters
    unhashed = list(exclude_args) + ([parallel_arg] if parallel_arg is not None else []) + list(output_dirs)
    bound = dict(kwargs)
    for od in output_dirs:
        bound[od] = None
    binding = signature(func).bind(**bound)
    binding.apply_defaults()
    for arg in exclude_args_if_default:
        if binding.arguments[arg] == params[arg].default:
            unhashed.append(arg)
    return _caching_dir(func, unhashed, bound, cache_dir, use_hash), unhashed


def _make_read_only(path: str) -> None:
    os.chmod(path, stat.S_IRUSR | stat.S_IRGRP | stat.S_IROTH)


def _write_unhashed_dir_log(out_dir: str, func, unhashed: List[str], kwargs, cache_dir: str) -> None:
    """``_unhashed_output_dir.log``: the cache directory this call would have without hashing
    (reference ``_cached_computation.py:96-129``); written once, then read-only."""
    log = os.path.join(out_dir, "_unhashed_output_dir.log")
    if os.path.exists(log):
        return
    plain = {k: (None if k in unhashed else v) for k, v in kwargs.items()}
    try:
        text = _caching_dir(func, unhashed, plain, cache_dir, use_hash=False)
    except Exception:  # unprintable arguments must not break the computation
        return
    with open(log, "w") as f:
        f.write(text)
    _make_read_only(log)


def cached_computation(
    exclude_args: List[str] = [],
    exclude_args_if_default: List[str] = [],
    output_dirs: List[str] = [],
    write_extra_log_files: bool = False,
):
    def decorator(func):
        params = signature(func).parameters
        named = list(exclude_args) + list(exclude_args_if_default) + list(output_dirs)
        for arg in named:
            if arg not in params:
                raise CacheUsageError(
                    f"{arg} is not an argument to '{func.__name__}'. Fix the "
                    f"arguments of the caching decorator."
                )
        if len(set(named)) != len(named):
            raise CacheUsageError(
                "All the function arguments specified in the caching decorator for "
                f"'{func.__name__}' should be distinct. You provided: {named} "
            )

        @wraps(func)
        def wrapper(*args, **kwargs):
            if len(args) > 0:
                raise CacheUsageError(
                    f"Please call {func.__name__} with keyword arguments only. "
                    f"Positional arguments are not allowed for caching reasons."
                )
            cache_dir = get_cache_dir()
            if cache_dir is None:
                return func(**kwargs)
            for od in output_dirs:  # output dirs may be required parameters: bind them as None
                kwargs.setdefault(od, None)
            func_dir, unhashed = _call_caching_dir(func, exclude_args, exclude_args_if_default, output_dirs, None,
                                                   kwargs, cache_dir)
            for od in output_dirs:
                if kwargs.get(od) is None:
                    kwargs[od] = os.path.join(func_dir, od)
            res = {od: kwargs[od] for od in output_dirs}

            def token(od, name):
                return os.path.join(kwargs[od], name)

            computed = all(
                os.path.exists(token(od, "result.txt")) and os.path.exists(token(od, "result.success"))
                for od in output_dirs
            )
            for od in output_dirs:
                os.makedirs(kwargs[od], exist_ok=True)
                if write_extra_log_files:
                    log = token(od, "_function_binding.log")
                    if not os.path.exists(log):
                        b = signature(func).bind(**kwargs)
                        b.apply_defaults()
                        with open(log, "w") as f:
                            f.write(str(b))
                        _make_read_only(log)
                    _write_unhashed_dir_log(kwargs[od], func, unhashed, kwargs, cache_dir)
            if not computed:
                if _READ_ONLY:
                    raise CacheUsageError("Cache is in read only mode! Will not call function.")
                for od in output_dirs:
                    for name in ("result.txt", "result.success"):
                        p = token(od, name)
                        if os.path.exists(p):
                            os.chmod(p, 0o666)
                            os.remove(p)
                func(**kwargs)
                for od in output_dirs:
                    if not os.path.exists(token(od, "result.txt")):
                        raise CacheUsageError(
                            f"function {func.__name__} should have created and written "
                            f"output to {token(od, 'result.txt')} but the file does not exist."
                        )
                for od in output_dirs:
                    _make_read_only(token(od, "result.txt"))
                    with open(token(od, "result.success"), "w") as f:
                        f.write("SUCCESS\n")
            return res

        wrapper.caching_dir = lambda cache_dir, use_hash=None, **kwargs: _call_caching_dir(
            func, exclude_args, exclude_args_if_default, output_dirs, None, kwargs, cache_dir, use_hash)[0]
        return wrapper

    return decorator


def secure_parallel_output(output_dir: str, parallel_arg: str) -> None:
    """Mark ``<output_dir>/<parallel_arg>.txt`` as done from inside a parallel stage: read-only
    plus a success token (reference ``_cached_parallel_computation.py:15-19``).  Stages here do not
    need it -- the decorator does this for every item after the function returns."""
    _make_read_only(os.path.join(output_dir, parallel_arg + ".txt"))
    with open(os.path.join(output_dir, parallel_arg + ".success"), "w") as f:
        f.write("SUCCESS\n")


def cached_parallel_computation(
    parallel_arg: str,
    exclude_args: List[str] = [],
    exclude_args_if_default: List[str] = [],
    output_dirs: List[str] = [],
    write_extra_log_files: bool = False,
):
    """Per-item caching (reference ``caching/_cached_parallel_computation.py:162-440``): the call
    is keyed on everything except ``parallel_arg``; item ``v`` counts as computed when every
    output dir holds ``v.txt`` and ``v.success``; the function is called with the remaining
    items only, its outputs are made read-only and given success tokens."""

    def decorator(func):
        params = signature(func).parameters
        named = list(exclude_args) + list(exclude_args_if_default) + [parallel_arg] + list(output_dirs)
        for arg in named:
            if arg not in params:
                raise CacheUsageError(
                    f"{arg} is not an argument to '{func.__name__}'. Fix the "
                    f"arguments of the caching decorator."
                )
        if len(set(named)) != len(named):
            raise CacheUsageError(
                "All the function arguments specified in the caching decorator for "
                f"'{func.__name__}' should be distinct. You provided: {named} "
            )

        @wraps(func)
        def wrapper(*args, **kwargs):
            if len(args) > 0:
                raise CacheUsageError(
                    f"Please call {func.__name__} with keyword arguments only. "
                    f"Positional arguments are not allowed for caching reasons."
                )
            kwargs[parallel_arg] = sorted(set(kwargs[parallel_arg]))
            cache_dir = get_cache_dir()
            if cache_dir is None:
                return func(**kwargs)
            given_dirs = {od: kwargs.get(od) for od in output_dirs}
            func_dir, unhashed = _call_caching_dir(func, exclude_args, exclude_args_if_default, output_dirs,
                                                   parallel_arg, kwargs, cache_dir)
            for od in output_dirs:
                kwargs[od] = given_dirs[od] if given_dirs[od] is not None else os.path.join(func_dir, od)
            res = {od: kwargs[od] for od in output_dirs}

            def paths(od, v):
                return os.path.join(kwargs[od], v + ".txt"), os.path.join(kwargs[od], v + ".success")

            # With a process group (this package's multi-GPU extension of the per-family stages) every rank
            # enters this wrapper: rank 0 alone decides what is left to do, cleans stale outputs, verifies and
            # tokenises; the list is broadcast and barriers keep a late rank from removing files that a faster
            # rank has already written.  (One node: the ranks share the cache directory.)
            group = kwargs.get("process_group") if "process_group" in params else None
            rank = 0
            if group is not None:
                import torch.distributed as dist

                rank = dist.get_rank(group)
            todo = None
            if rank == 0:
                todo = [
                    v for v in kwargs[parallel_arg]
                    if not all(os.path.exists(p) for od in output_dirs for p in paths(od, v))
                ]
            if group is not None:
                box = [todo]
                dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0), group=group)
                todo = box[0]
            kwargs[parallel_arg] = todo
            for od in output_dirs:
                os.makedirs(kwargs[od], exist_ok=True)
                if write_extra_log_files and rank == 0:
                    log = os.path.join(kwargs[od], "_function_binding.log")
                    if not os.path.exists(log):
                        logged = dict(kwargs)
                        for k in unhashed:
                            if k in logged:
                                logged[k] = None
                        b = signature(func).bind(**logged)
                        b.apply_defaults()
                        with open(log, "w") as f:
                            f.write(str(b))
                        _make_read_only(log)
                    _write_unhashed_dir_log(kwargs[od], func, unhashed, kwargs, cache_dir)
            if todo:
                if _READ_ONLY:
                    raise CacheUsageError("Cache is in read only mode! Will not call function.")
                if rank == 0:
                    for v in todo:
                        for od in output_dirs:
                            for p in paths(od, v):
                                if os.path.exists(p):
                                    os.chmod(p, 0o666)
                                    os.remove(p)
                if group is not None:
                    dist.barrier(group)  # stale outputs are gone before any rank writes
                func(**kwargs)
                if group is not None:
                    dist.barrier(group)  # every rank's outputs exist
                if rank == 0:
                    for od in output_dirs:
                        for v in todo:
                            out, _ = paths(od, v)
                            if not os.path.exists(out):
                                raise CacheUsageError(
                                    f"function {func.__name__} should have created and written "
                                    f"output to {out} but the file does not exist."
                                )
                    for od in output_dirs:
                        for v in todo:
                            out, tok = paths(od, v)
                            _make_read_only(out)
                            with open(tok, "w") as f:
                                f.write("SUCCESS\n")
                if group is not None:
                    dist.barrier(group)  # tokens exist before any rank returns
            return res

        wrapper.caching_dir = lambda cache_dir, use_hash=None, **kwargs: _call_caching_dir(
            func, exclude_args, exclude_args_if_default, output_dirs, parallel_arg, kwargs, cache_dir, use_hash)[0]
        return wrapper

    return decorator
