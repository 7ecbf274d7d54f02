"""Summarises an `ncu --set full` capture of one cfg-2 step (tools/prof_step.py 148) into profiles/r02_ncu_traffic.json:
DRAM bytes and ALU-pipe lane operations per evaluated cell of the aggregation kernels, utilisation of every kernel.
usage: ncu -i REPORT.ncu-rep --page raw --csv > profiles/r02_ncu_full_final_raw.csv; python tools/ncu_summary.py"""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RAW = os.path.join(ROOT, "profiles", "r02_ncu_full_final_raw.csv")
rows = list(csv.reader(open(RAW)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cells = 687 * 480 * 64 * 148
names = {"k_sgbm_prefilter": "sgbm_prefilter", "k_sgbm_vsum": "sgbm_vsum", "k_sgbm_h1": "sgbm_h1", "k_sweep": "sgbm_td",
         "k_sgbm_h2_wta": "sgbm_h2_wta", "k_median3": "median3", "k_ccl_rows": "ccl_rows", "k_ccl_vmerge": "ccl_vmerge",
         "k_ccl_flatten": "ccl_flatten", "k_ccl_apply": "ccl_apply"}


def num(r, k):
    v = r[ix[k]].replace(",", "")
    return float(v) if v else 0.0


def scaled(r, k, table):
    return num(r, k) * table.get(units[ix[k]], 1)


out = {"source": "ncu --set full --clock-control none --import-source on, tools/prof_step.py 148 (cfg 2, 148 frames, third step): "
                 "profiles/r02_ncu_full_final_raw.csv",
       "cells_per_launch": cells, "dram_bytes_per_cell": {}, "alu_pipe_lane_ops_per_cell": {},
       "alu_pipe_note": "ALU-pipe warp instructions = sm__pipe_alu_cycles_active (pct of peak, 2 warp instructions per clock per SM) "
                        "x sm__cycles_active x 148 SMs; x 32 lanes / evaluated cells",
       "utilisation": {}}
B = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
T = {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}
for r in rows[2:]:
    key = next((v for k, v in names.items() if k in r[ix["Kernel Name"]]), None)
    if not key:
        continue
    db = scaled(r, "dram__bytes_read.sum", B) + scaled(r, "dram__bytes_write.sum", B)
    alu = num(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active") / 100 * num(r, "sm__cycles_active.avg") * 2 * 148 * 32
    if key in ("sgbm_vsum", "sgbm_h1", "sgbm_td", "sgbm_h2_wta"):
        out["dram_bytes_per_cell"][key] = db / cells
        out["alu_pipe_lane_ops_per_cell"][key] = alu / cells
    out["utilisation"][key] = {
        "ms": scaled(r, "gpu__time_duration.sum", T),
        "issue_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "alu_pipe_pct": num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": num(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "smem_wavefront_pct": num(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "dram_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "regs": num(r, "launch__registers_per_thread"), "warp_instr": num(r, "smsp__inst_executed.sum"), "dram_bytes": db}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json"), "w"), indent=1)
print(json.dumps({k: out[k] for k in ("dram_bytes_per_cell", "alu_pipe_lane_ops_per_cell")}, indent=1))
tot = sum(v["ms"] for v in out["utilisation"].values())
for k, v in out["utilisation"].items():
    print("%-16s %6.3f ms  share %4.1f %%  issue %4.1f  alu %4.1f  fma %4.1f  smem %4.1f  dram %4.1f  regs %3d" % (
        k, v["ms"], 100 * v["ms"] / tot, v["issue_pct"], v["alu_pipe_pct"], v["fma_pipe_pct"], v["smem_wavefront_pct"], v["dram_pct"], v["regs"]))
