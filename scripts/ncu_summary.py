"""Key metrics of an `ncu --set full` report as a markdown table.  usage: python scripts/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "duration (us)"),
        ("sm__cycles_elapsed.max", "SM cycles"),
        ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (realtime)"),
        ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor-memory (TMEM) active %"),
        ("dram__bytes_read.sum", "DRAM read (MB)"),
        ("dram__bytes_write.sum", "DRAM write (MB)"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU wavefronts %"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem tensor-core wavefronts %"),
        ("launch__registers_per_thread", "registers/thread"),
        ("launch__grid_size", "grid"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem (KB)"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %")]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    recs = [dict(zip(h, r)) for r in rows[2:]]
    # every tensor-pipe metric the report holds (the name differs between ncu versions / chips; an empty row told us so in round 1)
    extra = [c for c in h if ("pipe_tensor" in c or "tensor_op" in c or "inst_executed_pipe_uniform" in c) and c not in dict(KEYS)]
    names = [r["Kernel Name"].replace("void ", "").split("(")[0][:60] + " grid " + r.get("launch__grid_size", "") for r in recs]
    print("| metric | " + " | ".join(f"launch {i}" for i in range(len(recs))) + " |")
    print("|---|" + "---:|" * len(recs))
    print("| kernel | " + " | ".join(f"`{n}`" for n in names) + " |")
    for k, label in KEYS + [(c, c) for c in extra]:
        vals = []
        for r in recs:
            v = r.get(k, "")
            try:
                vals.append(f"{float(v):.1f}")
            except ValueError:
                vals.append(v)
        print(f"| {label} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
