"""Summarise an .ncu-rep (raw page + source page) into text for profiles/.  Runs on the CPU box."""
import csv
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return dict(zip(rows[0], zip(rows[2], rows[1])))


def source(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    return hdr, rows[2:]


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "sm__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum", "sm__cycles_elapsed.max"]
STALLS = ["barrier", "branch_resolving", "dispatch_stall", "long_scoreboard", "math_pipe_throttle", "no_instruction",
          "not_selected", "selected", "short_scoreboard", "wait", "mio_throttle", "lg_throttle"]


def main(rep):
    d = raw(rep)
    print(f"# {rep}")
    for k in KEYS:
        if k in d:
            print(f"{k:85s} {d[k][0]:>20s} {d[k][1]}")
    print("-- warp stall reasons (warps per issue-active cycle) --")
    for s in STALLS:
        k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
        if k in d:
            print(f"  {s:22s} {d[k][0]}")
    hdr, rows = source(rep)
    iE, iT, iN = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    iS = hdr.index("Source")
    rec = []
    for r in rows:
        try:
            rec.append((r[iS], int(r[iE]), int(r[iT]), int(r[iN] or 0)))
        except (ValueError, IndexError):
            pass
    tot = sum(x[1] for x in rec) or 1
    tots = sum(x[3] for x in rec) or 1
    print(f"-- SASS: {len(rec)} instructions, {tot} warp-instructions executed, {sum(x[2] for x in rec) / tot:.2f} threads/instruction")
    for lo, hi in ((0, 4), (4, 8), (8, 16), (16, 24), (24, 31.5), (31.5, 33)):
        s = [x for x in rec if x[1] > 0 and lo <= x[2] / x[1] < hi]
        print(f"  lanes active [{lo:>4},{hi:>4}): {sum(x[1] for x in s) / tot * 100:6.2f} % of executed, {sum(x[3] for x in s) / tots * 100:6.2f} % of samples")
    ops = {}
    for x in rec:
        op = x[0].split()[0] if x[0].split() else "?"
        if op.startswith("@"):
            op = x[0].split()[1]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + x[1]
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:14]
    print("-- executed warp-instructions by opcode: " + ", ".join(f"{k} {v / tot * 100:.1f}%" for k, v in top))


if __name__ == "__main__":
    main(sys.argv[1])
