"""One eager, single-stream training step of the benchmarked configuration (unet_cct, pCE + GatedCRF, 64 x 1 x 256 x 256, bf16) for
ncu: three warm-up steps, then ONE step between cudaProfilerStart / Stop (run ncu with --profile-from-start off).  Writes
gpurun_out/launch_order.txt: every C-ABI call of that step in launch order with its layer label, so the anonymous ncu rows can be
named (profiles/summarize.py --labels).

    ncu --set full --clock-control none --profile-from-start off -o /tmp/r2_step python profiles/run_step.py [--precision bf16]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from wsl4mis_b200 import _lib  # noqa: E402
from wsl4mis_b200.engine import TrainStep  # noqa: E402
from wsl4mis_b200.networks.unet import UNet, UNet_CCT  # noqa: E402


class Recorder:
    """stands in for _lib.Profiler: records (entry point, label) per call and forwards the launch"""

    def __init__(self):
        self.rows, self.meta = [], None

    def record(self, name, fn, vals):
        rc = fn(*vals)
        if rc != 0:
            raise RuntimeError(f"{name} failed ({rc}): {_lib.LIB.last_error()}")
        self.rows.append((name, self.meta))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="unet_cct")
    ap.add_argument("--variant", default="pce_gatedcrf")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(2022)
    model = (UNet_CCT if args.model == "unet_cct" else UNet)(1, 4).to(dev).set_precision(args.precision)
    step = TrainStep(model, args.variant, graph=False)
    step.ex.multi_stream = False
    g = torch.Generator().manual_seed(2022)
    img = torch.rand(args.batch, 1, 256, 256, generator=g).to(dev)
    lab = torch.full((args.batch, 256, 256), 4, dtype=torch.uint8)
    m = torch.rand(args.batch, 256, 256, generator=g) < 0.03
    lab[m] = torch.randint(0, 4, (int(m.sum()),), generator=g, dtype=torch.uint8)
    lab = lab.to(dev)
    for _ in range(3):
        step(img, lab)
    torch.cuda.synchronize()
    rec = Recorder()
    _lib.PROFILE = rec
    torch.cuda.profiler.start()
    step(img, lab)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    _lib.PROFILE = None
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "launch_order.txt"), "w") as f:
        for name, meta in rec.rows:
            f.write(f"{name}\t{meta[0] if meta else '-'}\t{meta[1] if meta else '-'}\n")
    print(f"profiled one step: {len(rec.rows)} C-ABI calls")


if __name__ == "__main__":
    main()
