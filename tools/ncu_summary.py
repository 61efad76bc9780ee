#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of the step kernel into tracked files under profiles/.

  python tools/ncu_summary.py gpurun_out/r02x/prof.ncu-rep --tag r02_step_kernel --env-id StraightMimicWalker \
         --num-envs 4096 --integrator rk4

Writes profiles/<tag>.json (the metrics DESIGN.md / profiles/*.md quote) and records the kernel's DRAM traffic per launch
in profiles/step_kernel_traffic.json, which bench.py reads for `roofline.traffic` (so that the number in the bench line
is always the one of a committed capture, never a literal).
"""
import argparse
import csv
import io
import json
import os
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_per_block",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "smsp__warps_active.avg.per_cycle_active": "warps_active_per_scheduler",
    "smsp__warps_eligible.avg.per_cycle_active": "warps_eligible_per_scheduler",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_instruction",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum": "thread_ffma",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum": "thread_fadd",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum": "thread_fmul",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--tag", required=True)
    ap.add_argument("--env-id", default="StraightMimicWalker")
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--integrator", default="rk4")
    ap.add_argument("--command", default="")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = {"source": os.path.relpath(args.rep, REPO), "command": args.command, "launches": []}
    for vals in rows[2:]:
        rec = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""}
        stalls = {}
        for i, h in enumerate(hdr):
            if i >= len(vals) or vals[i] == "":
                continue
            if h in WANT:
                x = float(vals[i].replace(",", ""))
                x *= UNIT_SCALE.get(units[i].split("/")[0], 1.0)
                rec[WANT[h]] = x
            elif h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                name = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
                v = float(vals[i])
                if v >= 0.05:
                    stalls[name] = round(v, 3)
        rec["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
        if "dram_read" in rec and "dram_write" in rec:
            rec["dram_bytes"] = rec["dram_read"] + rec["dram_write"]
        out["launches"].append(rec)
    os.makedirs(os.path.join(REPO, "profiles"), exist_ok=True)
    with open(os.path.join(REPO, "profiles", args.tag + ".json"), "w") as f:
        json.dump(out, f, indent=1)
    step = [r for r in out["launches"] if "mimic_step" in r.get("kernel", "")]
    if step:
        path = os.path.join(REPO, "profiles", "step_kernel_traffic.json")
        table = {"captures": []}
        if os.path.exists(path):
            with open(path) as f:
                table = json.load(f)
        table["captures"] = [c for c in table["captures"] if not (c["env_id"] == args.env_id and c["num_envs"] == args.num_envs
                                                                   and c["integrator"] == args.integrator)]
        table["captures"].append({"env_id": args.env_id, "num_envs": args.num_envs, "integrator": args.integrator,
                                  "dram_bytes": step[0]["dram_bytes"], "source": "profiles/" + args.tag + ".json",
                                  "kernel": step[0]["kernel"]})
        with open(path, "w") as f:
            json.dump(table, f, indent=1)
    for r in out["launches"]:
        print(json.dumps({k: v for k, v in r.items() if k != "stalls_per_issue"}))
        print("  stalls/issue:", r["stalls_per_issue"])


if __name__ == "__main__":
    main()
