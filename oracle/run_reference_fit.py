"""Run the UNMODIFIED reference ``quantized_transitions_mle`` once and report its timings
(BENCH INFRASTRUCTURE: the reference arm of the fit).

    python oracle/run_reference_fit.py --counts result.txt --init init.txt --device cpu|cuda \
        --epochs E --threads T --out DIR

Prints one JSON line: wall seconds of the whole stage call (file parsing, tensors, E epochs, result
files), the trainer's own per-epoch clock (``df_res.txt`` column ``time``: seconds since the first
epoch started, trainer.py:146-213), the losses of the first and last epoch and the torch / device used.
Runs in its own process so that the stand-ins of oracle/ref_package.py never enter the bench process.
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--counts", required=True)
    ap.add_argument("--init", default=None)
    ap.add_argument("--mask", default=None)
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--epochs", type=int, required=True)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--out", required=True)
    ap.add_argument("--lr", type=float, default=0.1)
    args = ap.parse_args()
    if args.threads <= 0:
        args.threads = os.cpu_count() or 1
    if os.environ.get("CHERRY_REF_FIT_DUMP_AFTER"):  # debugging aid: where is it if it stalls
        import faulthandler

        faulthandler.dump_traceback_later(int(os.environ["CHERRY_REF_FIT_DUMP_AFTER"]), exit=True)
    import torch

    torch.set_num_threads(args.threads)
    from oracle.ref_package import import_reference

    import_reference()
    from cherryml.estimation import quantized_transitions_mle

    # Library start-up is not the reference's cost: torch imports its compiler stack lazily the first
    # time an optimizer is built (several seconds), CUDA creates its context and cuBLAS / cuSOLVER
    # handles on first use.  One throw-away Adam step on a tiny matrix_exp warms all of it.
    w = (0.1 * torch.ones(4, 4)).requires_grad_(True)
    opt = torch.optim.Adam([w], lr=0.1)
    torch.log(torch.matrix_exp(w)).sum().backward()
    opt.step()
    if args.device == "cuda":
        torch.matrix_exp(0.1 * torch.ones(4, 4, device="cuda"))
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    quantized_transitions_mle(
        count_matrices_path=args.counts, initialization_path=args.init, mask_path=args.mask,
        output_rate_matrix_dir=args.out, stationary_distribution_path=None,
        rate_matrix_parameterization="pande_reversible", device=args.device, learning_rate=args.lr,
        num_epochs=args.epochs, do_adam=True, OMP_NUM_THREADS=args.threads, OPENBLAS_NUM_THREADS=args.threads,
    )
    wall = time.perf_counter() - t0
    import pandas as pd

    df = pd.read_csv(os.path.join(args.out, "df_res.txt"))
    tcol = df["time"].to_numpy()
    out = {
        "wall_seconds": wall, "epochs": int(args.epochs), "device": args.device, "threads": int(args.threads),
        "train_seconds": float(tcol[-1]),
        "seconds_per_epoch": float((tcol[-1] - tcol[0]) / (len(tcol) - 1)) if len(tcol) > 1 else float(tcol[-1]),
        "first_epoch_seconds": float(tcol[0]),
        "loss_first": float(df["loss"].iloc[0]), "loss_last": float(df["loss"].iloc[-1]),
        "torch": torch.__version__,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
