"""Markdown summary of an ncu report (`ncu --set full ... -o X`), read here without a GPU:
    python scripts/ncu_summary.py gpurun_out/X.ncu-rep "title" [cells per launch] [algorithmic bytes per cell] > profiles/...md
One section per captured launch: duration, DRAM bytes, pipe utilisation, issue activity, registers,
shared-memory bank conflicts, the top stall reasons, and (when the cell count is given) warp
instructions per cell and DRAM traffic / algorithmic bytes."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
cells = float(sys.argv[3]) if len(sys.argv) > 3 else None
bpc = float(sys.argv[4]) if len(sys.argv) > 4 else 8.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
print("# %s\n" % title)
print("Read with `ncu -i %s --page raw --csv` (scripts/ncu_summary.py).  ncu replays each kernel ~40x with cold caches and\n"
      "serialised launches: durations here are NOT benchmark numbers (bench.py's CUDA-event timings are).\n" % rep.split("/")[-1])


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


for r in data:
    name = r[col["Kernel Name"]] if "Kernel Name" in col else "?"
    print("## %s\n" % name)
    print("| metric | value |\n|---|---|")
    vals = {}
    for m in WANT:
        if m in col:
            v, u = r[col[m]], units[col[m]]
            vals[m] = (num(v), u)
            print("| %s | %s %s |" % (m, v, u))
    stalls = []
    for n, i in col.items():
        if n.startswith("smsp__average_warp") and "issue_stalled" in n and n.endswith("_per_warp_active.pct") is False and "ratio" in n:
            x = num(r[i])
            if x is not None:
                stalls.append((x, n.split("issue_stalled_")[1].split("_per")[0].replace(".ratio", "")))
    stalls.sort(reverse=True)
    if stalls:
        print("| top stall reasons (warp-cycles per issued instruction) | %s |" % ", ".join("%s %.2f" % (n, x) for x, n in stalls[:5]))
    if cells:
        rd, wr = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")

        def to_bytes(v):
            if v is None or v[0] is None:
                return None
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(v[1], 1)
            return v[0] * scale
        b = (to_bytes(rd) or 0) + (to_bytes(wr) or 0)
        if b:
            print("| dram bytes per launch | %.0f |" % b)
            print("| DRAM traffic / algorithmic bytes (%g B/cell) | %.3f |" % (bpc, b / (cells * bpc)))
        ins = vals.get("smsp__inst_executed.sum")
        if ins and ins[0]:
            print("| warp instructions x 32 lanes / cells | %.1f |" % (ins[0] * 32 / cells))
    print()
