"""Instruction histogram of the hot kernels from the shipped library (cuobjdump -sass; no GPU needed):

    python profiles/sass_histogram.py > profiles/r02_sass_histogram.txt

For every kernel listed: the count of the mnemonics that identify the path it takes (DMMA = FP64 tensor
pipe, ATOMS.POPC.INC = warp-aggregated shared-memory reductions, LDGSTS = cp.async, RED/ATOMG = L2
reductions, SYNCS/ARRIVES = mbarrier) and the ten most frequent mnemonics.
"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cherryml_b200", "csrc", "libcherryml_b200.so")
KERNELS = ["count_lg_kernel", "count_co_sorted_kernel", "bucket_table_tiles_kernel", "chain_dataflow_kernel",
           "squaring_dataflow_kernel", "gemm_tasks_kernel", "taylor_fused_kernel", "taylor_fused_smem_kernel", "loss_grad_kernel",
           "accumulate_M_kernel", "expm_loss_grad_small",
           "fit_update_small", "fc_pair_kernel", "fc_ble_kernel"]
MARKERS = ["DMMA", "ATOMS.POPC.INC", "ATOMS", "LDGSTS", "RED", "ATOMG", "SYNCS", "IDP.4A", "LDG", "LDS", "STS", "BAR", "SHFL",
           "DFMA", "MUFU"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", sass)
    for blk in blocks[1:]:
        name = blk.split("\n", 1)[0].strip()
        if not any(k in name for k in KERNELS):
            continue
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        ops = collections.Counter()
        for line in blk.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                ops[m.group(1)] += 1
        total = sum(ops.values())
        short = re.sub(r"\(.*", "", demangled.replace("(anonymous namespace)::", "").replace("void ", ""))
        print(f"== {short}   ({total} SASS instructions)")
        marks = []
        for mk in MARKERS:
            n = sum(v for k, v in ops.items() if k == mk or k.startswith(mk + "."))
            if n:
                marks.append(f"{mk} {n}")
        print("   markers: " + ", ".join(marks))
        base = collections.Counter()
        for k, v in ops.items():
            base[k.split(".")[0]] += v
        print("   top: " + ", ".join(f"{k} {v}" for k, v in base.most_common(10)))


if __name__ == "__main__":
    main()
