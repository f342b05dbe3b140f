"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.
    python tools/ncu_summary.py gpurun_out/foo.ncu-rep profiles/r01_foo.txt [rows_per_launch]
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    units = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit_row = rows[0], rows[1]
    lines = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, unit_row))
        lines.append(f"kernel: {d.get('Kernel Name', '?')}")
        for k in KEYS:
            if k in d:
                lines.append(f"  {k:75s} {d[k]} {u[k]}")
        stalls = {h: float(d[h]) for h in hdr if "warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")}
        lines.append("  warp stall reasons (warps per issue-active cycle):")
        for h, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]:
            lines.append(f"    {h.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {v:.3f}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 2:
        h = srows[1]
        ix = {n: i for i, n in enumerate(h)}
        by_op, st_op, tot, samp = collections.Counter(), collections.Counter(), 0, 0
        for r in srows[2:]:
            if len(r) < len(h):
                continue
            toks = r[ix["Source"]].split()
            if not toks:
                continue
            op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            try:  # reports with several kernels repeat the header rows per kernel
                n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
            except ValueError:
                continue
            by_op[op] += n
            st_op[op] += s
            tot += n
            samp += s
        lines.append(f"  SASS: {len(srows) - 2} static instructions, {tot} warp-instructions executed"
                     + (f" = {tot / units:.0f} per row" if units else ""))
        for op, n in by_op.most_common(14):
            per = f"{n / units:8.1f}/row" if units else f"{n:12d}"
            lines.append(f"    {op:10s} {per} {100 * n / tot:5.1f}% of instructions, {100 * st_op[op] / max(samp, 1):5.1f}% of stall samples")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
