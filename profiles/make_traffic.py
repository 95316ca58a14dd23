#!/usr/bin/env python
"""DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every hot kernel, keyed by bench.py's kernel labels.
Inputs: an ncu launch list of `profiles/kbench.py ... --call-log calls.json` taken with
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv
and the call log of the SAME command line.  The C-ABI calls that launch a main kernel are matched, in order, with the main-kernel
rows of the launch list.   usage: python profiles/make_traffic.py LAUNCHES.csv CALLS.json OUT.json [source note]"""
import csv
import json
import re
import sys

MAIN_KERNEL = re.compile(r"dwconv7_v2_kernel|dwconv7_wgrad_v2_kernel|gemm_tn_tc|gemm_wgrad_pair_kernel|gemm_wgrad_tc_kernel|mlp_fused_fwd_kernel|ln_bwd_|ln_fwd_")
MAIN_CALL = {"cnx_dwconv7_ln_fwd": 1, "cnx_dwconv7_ln_fwd_x3": 1, "cnx_dwconv7_dgrad": 1, "cnx_dwconv7_wgrad": 1, "cnx_gemm_bias_gelu_fwd": 1,
             "cnx_gemm_bias_gelu_fwd_x3": 1, "cnx_gemm_bias_scale_residual_fwd": 1, "cnx_gemm_dgrad_gelu_bwd": 1, "cnx_gemm_plain": 1,
             "cnx_gemm_wgrad": 1, "cnx_mlp_fused_fwd": 1, "cnx_ln_bwd": 1, "cnx_ln_fwd": 1, "cnx_gemm_dgrad_gelu_recompute_bwd": 1}


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) >= 15 and r[0].isdigit()]
    per_id = {}
    order = []
    for r in rows:
        kid = int(r[0])
        if kid not in per_id:
            per_id[kid] = {"name": r[4]}
            order.append(kid)
        val = float(r[14].replace(",", ""))
        unit = r[13]
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
        per_id[kid][r[12]] = val * mult
    kernels = [per_id[k] for k in order if MAIN_KERNEL.search(per_id[k]["name"])]
    calls = [c for c in json.load(open(sys.argv[2])) if c[0] in MAIN_CALL]
    if len(kernels) != len(calls):
        print(f"warning: {len(kernels)} main-kernel launches vs {len(calls)} main calls; matching the common prefix", file=sys.stderr)
    out = {}
    for k, (name, label) in zip(kernels, calls):
        out[label] = int(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0))
    note = sys.argv[4] if len(sys.argv) > 4 else sys.argv[1]
    json.dump({"source": note, "kernels": out}, open(sys.argv[3], "w"), indent=1)
    print(f"{sys.argv[3]}: {len(out)} labels")


if __name__ == "__main__":
    main()
