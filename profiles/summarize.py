"""Turn `ncu --page raw --csv` exports into the small per-kernel tables committed under profiles/.

    python profiles/summarize.py gpurun_out/prof_conv_raw.csv > profiles/r1_conv_tc2.md
    python profiles/summarize.py raw.csv --labels gpurun_out/launch_order.txt [--only conv|wgrad|bn|loss|other]

--labels: the C-ABI calls of the profiled step in launch order (profiles/run_step.py).  Kernels are matched to calls family by
family in order (k-th convolution kernel <-> k-th wsl_conv_tc2 call, ...), which names the layer of every row.
"""
import csv
import sys

# kernel-name fragment -> C-ABI entry points whose launches it belongs to (one kernel per call unless noted)
FAMILIES = [
    (("conv_row_kernel", "conv_tc2_kernel"), ("wsl_conv_tc2",)),
    (("conv_tc_kernel",), ("wsl_conv_tc",)),
    (("wgrad_tc3_kernel",), ("wsl_wgrad_tc3",)),
    (("wgrad_tc_kernel",), ("wsl_wgrad_tc",)),
    (("bn_act_fwd_kernel",), ("wsl_bn_act_fwd",)),
    (("bn_bwd_reduce_kernel",), ("wsl_bn_bwd", "wsl_bn_bwd_first")),
    (("bn_bwd_apply_kernel",), ("wsl_bn_bwd",)),
    (("bn_bwd_apply_first_kernel",), ("wsl_bn_bwd_first",)),
    (("bn_finalize_kernel",), ("wsl_bn_finalize",)),
    (("upsample2x_fwd_kernel",), ("wsl_upsample2x_fwd",)),
    (("upsample2x_bwd_kernel",), ("wsl_upsample2x_bwd",)),
    (("conv_first_kernel",), ("wsl_conv_first",)),
    (("chan_scale_kernel",), ("wsl_chan_scale",)),
    (("channel_sum_kernel",), ("wsl_channel_sum",)),
]
GROUPS = {"conv": ("conv_row", "conv_tc2", "conv_tc_k", "conv_first"), "wgrad": ("wgrad",), "bn": ("bn_",),
          "loss": ("gatedcrf", "softmax_pce", "head_bwd", "mumford", "pdice", "mix_argmax"), "up": ("upsample",)}


def load_labels(path):
    calls = [ln.rstrip("\n").split("\t") for ln in open(path)]
    queues = {}
    for frags, entries in FAMILIES:
        q = [(c[1], c[2]) for c in calls if c[0] in entries]
        for f in frags:
            queues[f] = list(q)
    return queues


def label_for(kernel, queues):
    for f, q in queues.items():
        if f in kernel:
            if q:
                kind, layer = q.pop(0)
                return f"{layer} ({kind})" if layer != "-" else ""
            return ""
    return ""

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


def main(path, labels=None, only=None):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(h, n) for h, n, _ in COLS if h in idx]
    queues = load_labels(labels) if labels else None
    extra = []
    print("| " + ("layer | " if queues else "") + " | ".join(n for _, n in cols) + " |")
    print("|" + "---|" * (len(cols) + (1 if queues else 0)))
    for r in rows[2:]:
        kname = r[idx["Kernel Name"]]
        lab = label_for(kname, queues) if queues else None
        if only and not any(f in kname for f in GROUPS[only]):
            continue
        out = [lab] if queues else []
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
    a = sys.argv[1:]
    labels = a[a.index("--labels") + 1] if "--labels" in a else None
    only = a[a.index("--only") + 1] if "--only" in a else None
    main(a[0], labels, only)
