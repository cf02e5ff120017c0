#!/usr/bin/env python
"""Turns one `ncu --set full` capture (.ncu-rep) into the text summary kept under profiles/ and, for captures of the
tracking kernel launched by bench.py, refreshes the per-unit figures in profiles/roofline_traffic.json that bench.py reads.

  python tools/ncu_summary.py gpurun_out/r02_track_ws_bench.ncu-rep profiles/r02_track_ws_bench.summary.txt \
      --command "python bench.py --steps 1 --warmup 3 --seconds 2 --no-cpu --no-also --no-acq" --track packed2 --streams 64 --blocks 3906
"""
import argparse
import csv
import io
import json
import os
import subprocess

KEYS = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_allocated", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
]
STALLS = ["long_scoreboard", "math_pipe_throttle", "not_selected", "wait", "short_scoreboard", "no_instruction", "barrier",
          "branch_resolving", "dispatch_stall", "mio_throttle"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--command", default="")
    ap.add_argument("--note", default="")
    ap.add_argument("--track", default=None, choices=[None, "packed2", "int8"])
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--blocks", type=int, default=3906)
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    lines = [f"kernel: {d.get('Kernel Name')}, grid {d.get('Grid Size')} x block {d.get('Block Size')}"]
    if a.note:
        lines.append("        " + a.note)
    if a.command:
        lines.append(f"command: {a.command}  (ncu --set full --clock-control none --import-source on, one launch)")
    for k in KEYS:
        if k in d:
            lines.append(f"{k} [{u.get(k, '')}] = {d[k]}")
    for s in STALLS:
        k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
        if k in d:
            lines.append(f"{k} = {d[k]}")

    def num(k):
        return float(d[k].replace(",", "")) if k in d and d[k] not in ("", "n/a") else None

    def in_bytes(k):
        v = num(k)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u.get(k, "byte"), 1)
        return None if v is None else v * scale

    if a.track:
        ss = a.streams * a.blocks * 8192  # stream-samples of the launch
        cs = ss * 12
        wi = num("smsp__inst_executed.sum") / cs
        tr = (in_bytes("dram__bytes_read.sum") + in_bytes("dram__bytes_write.sum")) / ss
        alg = 0.5 if a.track == "packed2" else 2.0
        lines.append(f"derived: {wi:.4f} warp-instructions per channel-sample ({32 * wi:.1f} thread-instructions); "
                     f"{num('smsp__inst_executed.sum') / (a.streams * 12 * a.blocks):.0f} warp-instructions per channel-block")
        lines.append(f"derived: DRAM traffic {tr:.4f} B per stream-sample (algorithmic {alg} B): reads "
                     f"{in_bytes('dram__bytes_read.sum') / (ss * alg):.3f} x the record, writes {in_bytes('dram__bytes_write.sum') / 1e6:.1f} MB")
        p = os.path.join(os.path.dirname(os.path.abspath(a.out)), "roofline_traffic.json")
        j = json.load(open(p)) if os.path.exists(p) else {}
        j[f"track_dram_bytes_per_stream_sample_{a.track}"] = tr
        j[f"track_warp_inst_per_channel_sample_{a.track}"] = wi
        j.setdefault("captures", {})[a.track] = f"{os.path.basename(a.out)} ({a.command}; {a.streams} streams per GPU)"
        json.dump(j, open(p, "w"), indent=1)
    open(a.out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
