#!/usr/bin/env python3
"""Summarises an `ncu --set full` report (read here, no GPU needed) into a small JSON + markdown:
    python profiles/summarise.py gpurun_out/gl_iter_r01.ncu-rep profiles/r01_gl_iter_v1 --frame-iters 43776
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.avg", "lts__t_bytes.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_red.sum",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_warps", "sm__maximum_warps_per_active_cycle_pct"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    ksel, alg_bytes = [], 18436
    if "--kernel" in sys.argv:          # regex on the kernel name when a report holds several kernels
        ksel = ["-k", "regex:" + sys.argv[sys.argv.index("--kernel") + 1]]
    if "--alg-bytes" in sys.argv:
        alg_bytes = int(sys.argv[sys.argv.index("--alg-bytes") + 1])
    frame_iters = None
    if "--frame-iters" in sys.argv:
        frame_iters = int(sys.argv[sys.argv.index("--frame-iters") + 1])
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep] + ksel + ["--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    res = {"report": rep, "kernel": data[0][col["Kernel Name"]], "launches_captured": len(data), "metrics": {}}
    for k in KEYS:
        if k in col:
            res["metrics"][k] = {"unit": units[col[k]], "values": [r[col[k]] for r in data]}
    unit_scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    rd = float(data[0][col["dram__bytes_read.sum"]]) * unit_scale.get(units[col["dram__bytes_read.sum"]], 1.0)
    wr = float(data[0][col["dram__bytes_write.sum"]]) * unit_scale.get(units[col["dram__bytes_write.sum"]], 1.0)
    res["dram_bytes_per_launch"] = rd + wr
    if frame_iters:
        res["frame_iters_per_launch"] = frame_iters
        res["dram_bytes_per_frame_iter"] = (rd + wr) / frame_iters
        res["algorithmic_bytes_per_frame_iter"] = alg_bytes
    # stall reasons + opcode mix from the source page
    src = list(csv.reader(io.StringIO(ncu(["-i", rep] + ksel + ["--page", "source", "--csv", "--print-source", "sass"]))))
    hidx = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    h = src[hidx[0]]
    end = hidx[1] - 1 if len(hidx) > 1 else len(src)
    body = src[hidx[0] + 1:end]
    c2 = {n: i for i, n in enumerate(h)}
    stalls = collections.Counter()
    ops = collections.Counter()
    for r in body:
        for n in h:
            if n.startswith("stall_") and "Not Issued" not in n:
                try:
                    stalls[n] += int(r[c2[n]])
                except (ValueError, IndexError):
                    pass
        w = r[c2["Source"]].split()
        if w:
            o = (w[1] if w[0].startswith("@") and len(w) > 1 else w[0]).split(".")[0]
            try:
                ops[o] += int(r[c2["Instructions Executed"]])
            except ValueError:
                pass
    tot = sum(stalls.values()) or 1
    res["sass_lines"] = len(body)
    res["stall_samples_pct"] = {k: round(100.0 * v / tot, 2) for k, v in stalls.most_common(12)}
    ti = sum(ops.values()) or 1
    res["opcode_mix_pct"] = {k: round(100.0 * v / ti, 2) for k, v in ops.most_common(20)}
    res["warp_instructions"] = ti
    json.dump(res, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu summary: {res['kernel']}\n\nreport: `{rep}` ({res['launches_captured']} launches captured)\n\n")
        f.write("| metric | unit | value(s) |\n|---|---|---|\n")
        for k, v in res["metrics"].items():
            f.write(f"| {k} | {v['unit']} | {', '.join(v['values'])} |\n")
        f.write(f"\nDRAM bytes per launch: {res['dram_bytes_per_launch']:.4g}")
        if frame_iters:
            f.write(f" = {res['dram_bytes_per_frame_iter']:.0f} B per frame-iteration (algorithmic {alg_bytes:,})")
        f.write(f"\n\nSASS lines: {res['sass_lines']}, warp instructions executed: {ti}\n\n## stall samples (%)\n\n")
        for k, v in res["stall_samples_pct"].items():
            f.write(f"* {k}: {v}\n")
        f.write("\n## opcode mix (% of executed warp instructions)\n\n")
        for k, v in res["opcode_mix_pct"].items():
            f.write(f"* {k}: {v}\n")
    print("wrote", out + ".json", out + ".md")


if __name__ == "__main__":
    main()
