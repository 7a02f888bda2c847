"""Turn .ncu-rep captures (gpurun_out/) into the markdown tables kept under profiles/.
    python profiles/summarize.py gpurun_out/prof_k2.ncu-rep [...]      (needs the ncu CLI; no GPU)
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe % (active)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "fp64 pipe % (elapsed)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fp32 fma pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__inst_executed_op_local_ld.sum", "local loads"),
    ("smsp__inst_executed_op_local_st.sum", "local stores"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        print("### %s\n" % path.split("/")[-1])
        names = [r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("brdfgpu::", "") for r in data]
        print("| metric | " + " | ".join("%s #%d" % (n, i) for i, n in enumerate(names)) + " |")
        print("|---|" + "---|" * len(data))
        for key, label in WANT:
            if key not in col:
                continue
            vals = []
            for r in data:
                v = r[col[key]]
                try:
                    v = "%.4g" % float(v.replace(",", ""))
                except ValueError:
                    pass
                vals.append("%s %s" % (v, units[col[key]]))
            print("| %s | %s |" % (label, " | ".join(vals)))
        print()


if __name__ == "__main__":
    main()
