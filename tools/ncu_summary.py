"""Summarise an .ncu-rep (raw page + SASS source page) into the few numbers DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--sass-chunks 40]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sectors_op_read.sum",
    "lts__t_sectors_op_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "smsp__inst_executed_pipe_fp64", "sm__inst_executed_pipe_alu",
    "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_lsu", "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_uniform",
    "smsp__average_warps_issue_stalled",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    chunk = int(sys.argv[sys.argv.index("--sass-chunks") + 1]) if "--sass-chunks" in sys.argv else 0
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("== kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
        for h, u, v in zip(hdr, units, vals):
            if any(h.startswith(k) for k in KEYS) and ".max" not in h and ".min" not in h and "peak_sustained" not in h.split("pct_of_")[0]:
                print(f"{h} [{u}] = {v}")
    if chunk:
        rows = page(rep, "source")
        hdr, data = rows[1], rows[2:]
        i_s, i_n, i_e = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        tot = sum(int(r[i_e]) for r in data) or 1
        ts = sum(int(r[i_n]) for r in data) or 1
        print(f"== SASS: {len(data)} instructions, {tot} warp-instructions executed, {ts} samples")
        for i in range(0, len(data), chunk):
            blk = data[i:i + chunk]
            e = sum(int(r[i_e]) for r in blk)
            s = sum(int(r[i_n]) for r in blk)
            ops = {}
            for r in blk:
                tok = r[i_s].split()
                op = tok[1] if tok and tok[0].startswith("@") else (tok[0] if tok else "?")
                ops[op] = ops.get(op, 0) + 1
            top = ", ".join(f"{k}x{v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:6])
            print(f"{i:5d} exec {e / tot * 100:5.1f}% samples {s / ts * 100:5.1f}%  {top}")


if __name__ == "__main__":
    main()
