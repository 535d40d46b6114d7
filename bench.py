#!/usr/bin/env python
"""Benchmark of the WSL4MIS hot path: images/sec of the unet_cct pCE+GatedCRF training step
(train_weakly_supervised_pCE_GatedCRFLoss_2D.py:111-130 replayed on main_seg) at 256x256,
batch 64 per GPU, synthetic data, random-init weights.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # CPU restatement of the reference step on the host cores

Prints ONE JSON line (rank 0).  value = whole-job images/sec with inputs resident in HBM;
e2e = same step driven through the public API from pinned HOST buffers (H2D of the batch and D2H of the loss
inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "images/sec unet_cct pCE+GatedCRF 256x256 bs64"
HW = 256


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (weak scaling)")
    ap.add_argument("--model", default="unet_cct", choices=["unet", "unet_cct"])
    ap.add_argument("--variant", default="pce_gatedcrf")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-kernel-table", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=2, help="images per CPU-baseline step")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs: skip the CPU baseline leg")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: skip the host-buffer leg")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


# ----------------------------------------------------------------------------------------------
# CPU baseline: the oracle restatement of the reference step (the ONLY place bench.py touches oracle/)
# ----------------------------------------------------------------------------------------------
def cpu_step_time(model_name, variant, n_img, steps, warmup):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import wsl_oracle as O
    O.CRF_IMPL = "unfold"          # time the reference's own (materialising) GatedCRF formulation
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    try:  # honour a cgroup CPU quota (oversubscribing a quota-limited container is catastrophically slow)
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            cores = max(1, min(cores, int(float(q) / float(per))))
    except Exception:
        pass
    cores = min(cores, 64)   # torch CPU ops stop scaling (and start thrashing) far below 128 threads on these sizes
    torch.set_num_threads(cores)
    cct = model_name == "unet_cct"
    decs = ("main_decoder", "aux_decoder1") if cct else ("decoder",)
    p = O.synth_params(1, 4, decs, 2022)
    image, label = O.synth_batch(n_img, HW, HW, seed=2022)
    moms = {}
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        loss, grads, _ = O.full_step(p, image, label, variant, cct)
        trainable = {k: p[k] for k in grads}
        O.sgd_step(trainable, grads, moms, 0.01)
        p.update(trainable)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = sum(times) / len(times)
    return {"value": n_img / t, "unit": "images/sec", "cores": cores, "kind": "port",
            "sample": f"{n_img} images of 1x{HW}x{HW} per step, {steps} steps after {warmup} warm-up, torch {torch.__version__} fp32, "
                      f"oracle/wsl_oracle.py full_step + sgd_step ({model_name}, {variant})"}, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    cb, t = cpu_step_time(args.model, args.variant, args.cpu_sample, steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "images/sec", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.model} {args.variant} train step, {args.cpu_sample}x1x{HW}x{HW} per step on host CPU",
                       "note": "oracle port of the reference step (the reference is pure PyTorch; /root/reference is absent on the GPU box)"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from wsl4mis_b200 import _lib
    from wsl4mis_b200.engine import TrainStep
    from wsl4mis_b200.networks.unet import UNet, UNet_CCT

    N = args.batch
    torch.manual_seed(2022)
    model = (UNet_CCT if args.model == "unet_cct" else UNet)(1, 4).to(dev)
    step = TrainStep(model, args.variant, base_lr=0.01, max_iterations=30000, graph=not args.no_graph, world_size=world)

    # synthetic batch: image ~ U[0,1), ~3 % scribble pixels; different per rank
    g = torch.Generator().manual_seed(2022 + rank)
    img_h = torch.rand(N, 1, HW, HW, generator=g).pin_memory()
    lab_h = torch.full((N, HW, HW), 4, dtype=torch.uint8)
    m = torch.rand(N, HW, HW, generator=g) < 0.03
    lab_h[m] = torch.randint(0, 4, (int(m.sum()),), generator=g, dtype=torch.uint8)
    lab_h = lab_h.pin_memory()
    img_d, lab_d = img_h.to(dev), lab_h.to(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (>= 3: two eager steps allocate everything, the third captures the graph) ----
    W = max(args.warmup, 3)
    for _ in range(W):
        loss = step(img_d, lab_d)
    barrier()
    first_loss = float(loss.item())

    # ---- timed region 1: device-resident inputs ----
    K = args.steps
    sampler = ClockSampler(local) if rank == 0 else None
    c0 = _lib.COUNTERS["launch_calls"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(K):
        loss = step(img_d, lab_d)
    ev1.record()
    barrier()
    ms_dev = ev0.elapsed_time(ev1)
    launches_eager = _lib.COUNTERS["launch_calls"] - c0

    # ---- timed region 2: end to end from pinned host memory, loss read back every step ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    bufs = [step.input_buffers(0), step.input_buffers(1)] if step.graph_enabled else [(img_d, lab_d), (img_d.clone(), lab_d.clone())]
    bufs[1][0].copy_(img_d)
    bufs[1][1].copy_(lab_d)
    step(*bufs[1])                                   # captures the second graph outside the timed region
    torch.cuda.synchronize()

    def prefetch(i):
        """H2D of one batch from pinned host memory on the copy stream (overlaps the previous step's graph)"""
        with torch.cuda.stream(copy_stream):
            bufs[i][0].copy_(img_h, non_blocking=True)
            bufs[i][1].copy_(lab_h, non_blocking=True)
            copied[i].record(copy_stream)

    loss_h = torch.zeros(2, dtype=torch.float32).pin_memory()
    read_ev = [torch.cuda.Event(), torch.cuda.Event()]
    barrier()
    e0.record()
    n_e2e = 0 if args.skip_e2e else K
    if n_e2e:
        prefetch(0)
    lv = float("nan")
    for k in range(n_e2e):
        cur = k % 2
        main_stream.wait_event(copied[cur])          # this step's inputs have landed
        loss = step(*bufs[cur])
        loss_h[cur:cur + 1].copy_(loss.detach().reshape(1), non_blocking=True)   # D2H of this step's result ...
        read_ev[cur].record()
        if k + 1 < n_e2e:
            if k >= 1:
                copy_stream.wait_event(read_ev[1 - cur])   # step k-1, the last reader of that buffer pair, has finished
            prefetch((k + 1) % 2)                    # next step's H2D runs under this step's graph
        if k >= 1:                                   # ... consumed on the host one step later (software pipelining:
            read_ev[1 - cur].synchronize()           # the launch of step k is never stalled by reading step k-1)
            lv = float(loss_h[1 - cur])
    if n_e2e:
        read_ev[(n_e2e - 1) % 2].synchronize()
        lv = float(loss_h[(n_e2e - 1) % 2])
    e1.record()
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 1e-6)
    if args.skip_e2e:
        lv = loss.item()
    clocks = sampler.stop() if sampler is not None else None

    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()

    # ---- per-kernel table (eager, CUDA events around every C-ABI launch), rank 0, outside the timed regions ----
    table, roof, launches_per_step = None, None, None
    if rank == 0 and not args.no_kernel_table:
        peaks = load_peaks()
        prof = _lib.Profiler()
        eager = TrainStep(model, args.variant, graph=False, world_size=1)
        eager.ex.multi_stream = False        # serialise: per-kernel event times are only meaningful without overlap
        eager(img_d, lab_d)
        torch.cuda.synchronize()
        _lib.PROFILE = prof
        n_prof = 2
        for _ in range(n_prof):
            eager(img_d, lab_d)
        _lib.PROFILE = None
        agg = prof.table()
        byname = {}
        for (name, meta), (cnt, ms) in agg.items():
            d = byname.setdefault(name, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            d["launches"] += cnt
            d["ms"] += ms
            if meta is not None:
                d["flops"] += meta[2] * cnt
                d["bytes"] += meta[3] * cnt
        detail = sorted(((name, meta[0], meta[1], cnt, ms / cnt, meta[2] / (ms / cnt * 1e-3) / 1e12, meta[3] / (ms / cnt * 1e-3) / 1e9)
                         for (name, meta), (cnt, ms) in agg.items() if meta is not None), key=lambda r: -r[3] * r[4])
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "kernel_detail.txt"), "w") as f:
            f.write("entry kind layer launches avg_ms TFLOP/s algorithmic_GB/s\n")
            for r in detail:
                f.write(f"{r[0]} {r[1]} {r[2]} {r[3] // n_prof} {r[4]:.4f} {r[5]:.1f} {r[6]:.0f}\n")
        tot = sum(d["ms"] for d in byname.values())
        launches_per_step = sum(d["launches"] for d in byname.values()) // n_prof
        table = {k: {"launches_per_step": v["launches"] // n_prof, "ms_per_step": round(v["ms"] / n_prof, 4),
                     "share": round(v["ms"] / tot, 4),
                     **({"tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2)} if v["flops"] else {})}
                 for k, v in sorted(byname.items(), key=lambda kv: -kv[1]["ms"])}
        top = max(byname.items(), key=lambda kv: kv[1]["ms"])
        name, v = top
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if tj.get("kernel") == name:
                traffic = {"dram_MB_per_launch": round(tj["dram_bytes_per_launch_MB"], 1), "source": tj["source"]}
        if v["flops"] > 0:
            ach = v["flops"] / (v["ms"] * 1e-3) / 1e12
            roof = {"kernel": name, "bound": "tensor", "achieved": round(ach, 2), "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                    "frac": round(ach / peaks["tf_sustained"], 4), "traffic": traffic,
                    "algorithmic_MB_per_launch": round(v["bytes"] / v["launches"] / 1e6, 1),
                    "peak_source": peaks["src"] + " bf16 sustained (kernel timed inside a long step)",
                    "avg_launch_ms": round(v["ms"] / v["launches"], 4), "algorithmic_flops_per_launch": v["flops"] / v["launches"]}
        else:
            byts = 36.0 * N * HW * HW if "gatedcrf" in name else 0.0
            ach = byts * (v["launches"]) / (v["ms"] * 1e-3) / 1e9 if byts else 0.0
            roof = {"kernel": name, "bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": round(ach / peaks["hbm_gbs"], 4), "traffic": None, "peak_source": peaks["src"]}

    if rank == 0:
        cb = None if args.skip_cpu else cpu_step_time(args.model, args.variant, args.cpu_sample, 2, 1)[0]
        imgs = N * world * K
        line = {
            "metric": METRIC, "value": imgs / (ms_dev * 1e-3), "unit": "images/sec", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.model} {args.variant} train step (fwd+loss+bwd+SGD), {N}x1x{HW}x{HW} per GPU, "
                                   f"loss on main_seg (SURVEY F7)", "global_batch": N * world, "parallelism": f"dp{world}",
                       "cuda_graph": step.graph_enabled, "bn": "per-rank batch statistics (stock DDP semantics)",
                       "l2": "per-step working set (~6 GB of activations) >> 126 MB L2; no explicit flush needed"},
            "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "images/sec", "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": int(img_h.numel() * 4 + lab_h.numel()), "d2h_bytes_per_step": 4},
            "gpu_launches": (step.launches_per_step * K) if step.graph_enabled else launches_eager,
            "gpu_launches_per_step": step.launches_per_step,
            "clocks": clocks, "roofline": roof, "cpu_baseline": cb, "kernels": table,
            "loss_first": first_loss, "loss_last": float(lv),
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
