"""Turn `ncu --page raw --csv` exports into the small per-kernel tables committed under profiles/.

    python profiles/summarize.py gpurun_out/prof_conv_raw.csv > profiles/r1_conv_tc2.md
"""
import csv
import sys

COLS = [
    ("Kernel Name", "kernel", None),
    ("launch__grid_size", "grid", None),
    ("launch__registers_per_thread", "regs", None),
    ("gpu__time_duration.sum", "time_us", 1.0),
    ("dram__bytes_read.sum", "dram_rd_MB", 1.0),
    ("dram__bytes_write.sum", "dram_wr_MB", 1.0),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 1.0),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%", 1.0),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst", 1.0),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%", 1.0),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%", 1.0),
    ("lts__t_bytes.sum", "l2_bytes", 1.0),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(h, n) for h, n, _ in COLS if h in idx]
    # also pick up any tensor-pipe metric names that exist in this ncu version
    extra = [h for h in hdr if "pipe_tensor" in h and h not in dict(cols)]
    print("| " + " | ".join(n for _, n in cols) + " | " + " | ".join(e.replace("sm__", "") for e in extra[:3]) + " |")
    print("|" + "---|" * (len(cols) + len(extra[:3])))
    for r in rows[2:]:
        out = []
        for h, n in cols:
            v = r[idx[h]]
            if n == "kernel":
                v = v.replace("<unnamed>::", "").split("(")[0][:46]
            else:
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:.1f}" if f < 1e6 else f"{f:.3g}"
                except ValueError:
                    pass
            out.append(v)
        for e in extra[:3]:
            out.append(r[idx[e]][:10])
        print("| " + " | ".join(out) + " |")
    print()
    print("units: " + ", ".join(f"{n}={units[idx[h]]}" for h, n in cols if units[idx[h]]))


if __name__ == "__main__":
    main(sys.argv[1])
