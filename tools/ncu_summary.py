"""Summaries kept under profiles/ from one `ncu --set full --import-source on` report.

  ncu -i X.ncu-rep --page raw --csv > raw.csv
  ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv
  python tools/ncu_summary.py counters raw.csv            > profiles/..._ncu.txt
  python tools/ncu_summary.py lines src.csv KERNEL [N]    > profiles/..._hot_lines.txt
  python tools/ncu_summary.py json raw.csv TRAJECTORIES KERNEL [KERNEL ...] > profiles/r2_kernel_counters.json
      (per-trajectory DRAM bytes and executed FP64 instructions of each KERNEL, first matching launch:
       what bench.py reads for roofline.traffic and the FP64 fraction)
"""
import csv, sys, collections

COUNTERS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_op_shared_st.sum",
]


def _num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def as_json(path, trajectories, kernels):
    import json
    out = {}
    for kernel in kernels:
        out[kernel] = one_kernel(path, kernel, trajectories)
    print(json.dumps(out, indent=1))


def one_kernel(path, kernel, trajectories):
    rows = list(csv.reader(open(path)))
    head, units = rows[0], rows[1]
    for row in rows[2:]:
        if kernel not in row[head.index("Kernel Name")]:
            continue

        def get(name, scale=True):
            if name not in head:
                return None
            i = head.index(name)
            v = _num(row[i])
            if v is None:
                return None
            u = units[i].lower()
            if scale and u.startswith("mbyte"):
                v *= 1e6
            elif scale and u.startswith("gbyte"):
                v *= 1e9
            elif scale and u.startswith("kbyte"):
                v *= 1e3
            return v
        dram = (get("dram__bytes_read.sum") or 0.0) + (get("dram__bytes_write.sum") or 0.0)
        dfma = get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum")
        dmul = get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum")
        dadd = get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum")
        out = {"kernel": row[head.index("Kernel Name")][:80], "trajectories_per_launch": trajectories,
               "gpu_time": "%s %s" % (row[head.index("gpu__time_duration.sum")], units[head.index("gpu__time_duration.sum")]),
               "dram_bytes_per_trajectory": dram / trajectories,
               "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
               "warp_instructions_per_trajectory": (get("smsp__inst_executed.sum") or 0.0) / trajectories,
               "source": "ncu --set full --clock-control none (cold-cache, serialised replays)"}
        if None not in (dfma, dmul, dadd):
            out["fp64_instructions_per_trajectory"] = (dfma + dmul + dadd) / trajectories
            out["fp64_flops_per_trajectory"] = (2 * dfma + dmul + dadd) / trajectories
        return out
    raise SystemExit("kernel %s not in %s" % (kernel, path))



def counters(path):
    rows = list(csv.reader(open(path)))
    head, units = rows[0], rows[1]
    for row in rows[2:]:
        print("=" * 60)
        for name in ["Kernel Name"] + COUNTERS:
            if name in head:
                i = head.index(name)
                print("%-86s %s %s" % (name, row[i], units[i]))


def lines(path, kernel, top=40):
    rows = list(csv.reader(open(path)))
    cur, hdr, out, take = None, None, [], False
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            take = kernel in r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            iex, ith, ism = (hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"),
                             hdr.index("# Samples"))
            continue
        if take and hdr and r[0] != "":
            try:
                out.append((cur, int(r[0]), r[1].strip(), int(r[iex]), int(r[ith]), int(r[ism])))
            except ValueError:
                pass
    tot, ts = sum(o[3] for o in out), sum(o[5] for o in out)
    print("kernel %s: warp instructions %d, stall samples %d" % (kernel, tot, ts))
    print("file line | share of warp instructions | active lanes | share of stall samples | source")
    for o in sorted(out, key=lambda o: -o[3])[:top]:
        print("%-20s %4d exec %5.1f%% act %4.1f smp %5.1f%% | %s" % (
            o[0], o[1], 100 * o[3] / tot, o[4] / max(o[3], 1), 100 * o[5] / max(ts, 1), o[2][:100]))


if __name__ == "__main__":
    if sys.argv[1] == "counters":
        counters(sys.argv[2])
    elif sys.argv[1] == "json":
        as_json(sys.argv[2], int(sys.argv[3]), sys.argv[4:])
    else:
        lines(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
