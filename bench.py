"""Benchmark of the 3D U-Net train step (BASELINE.json metric: UNet3d train voxels/sec on (B,1,128,128,128)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one full train step of BASELINE.json configs[1] on one synthetic batch per GPU:
``UNet3d(1, 2, depth=4, initial_features=32, final_activation="Sigmoid")``, bf16 autocast, (4,1,128,128,128) patches,
DiceLoss, zero_grad + forward + loss + backward (+ one gradient all-reduce for N > 1) + AdamW step.
Weak scaling: the per-GPU batch is fixed.  ``value`` = voxels/s of the whole job with inputs resident in HBM;
``e2e`` = the same with pinned-host inputs copied H2D and the loss read back D2H inside every timed step.
``--impl reference`` times the reference's CPU arithmetic for this path (the oracle port: /root/reference does not
exist on the GPU box and is pure Python over torch, so "the reference compiled here" does not apply) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL_KW = dict(in_channels=1, out_channels=2, depth=4, initial_features=32, final_activation="Sigmoid")
BATCH, PATCH = 4, (128, 128, 128)
METRIC = "UNet3d train voxels/sec on (B,1,128,128,128)"
NCU_FULL_CSV = "profiles/r01_kernels_ncu_full.csv"
# backend launch label -> __global__ function it launches (csrc/)
KERNEL_OF = {"first:fwd": "conv3d_first_kernel", "first:wgrad": "conv3d_first_wgrad_kernel", "ds:fwd": "conv3d_umma_ds_kernel",
             "ds:dgrad": "conv3d_umma_ds_kernel", "plain:fwd": "conv3d_umma_kernel", "plain:dgrad": "conv3d_umma_kernel",
             "thin:fwd": "conv3d_umma_kernel", "cs:wgrad": "conv3d_wgrad_cs_kernel", "umma:wgrad": "conv3d_wgrad_umma_kernel",
             "thin:wgrad": "conv3d_wgrad_umma_kernel", "direct:fwd": "conv3d_direct_kernel", "direct:dgrad": "conv3d_direct_kernel",
             "direct:wgrad": "conv3d_wgrad_direct_kernel", "smallcin:wgrad": "conv3d_wgrad_smallcin_kernel"}


def synthetic_batch(batch, patch, seed):
    """Per-sample standardised random volume (like transform/raw.py:40-65) and binary 2-channel targets."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((batch, 1) + tuple(patch), generator=g)
    x = (x - x.mean(dim=(1, 2, 3, 4), keepdim=True)) / (x.std(dim=(1, 2, 3, 4), keepdim=True) + 1e-7)
    t = (torch.rand((batch, 2) + tuple(patch), generator=g) > 0.5).float()
    return x, t


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for nm, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        smax = max((float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(self.samples)}


def reference_root():
    """The UNMODIFIED reference package: baseline/_ref (offline `pip install --no-deps --target baseline/_ref` of
    /root/reference; git-ignored, travels to the GPU box) or /root/reference in the build container."""
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(cand, "torch_em", "model", "unet.py")):
            return cand
    return None


def load_reference_modules():
    """The reference's torch-only modules model/unet.py and loss/dice.py, loaded by file path (``import torch_em`` needs
    imageio/skimage/... which this image lacks).  -> (unet module, dice module) or (None, None)."""
    root = reference_root()
    if root is None:
        return None, None
    import importlib.util
    mods = []
    for name, rel in (("_ref_unet", "model/unet.py"), ("_ref_dice", "loss/dice.py")):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, "torch_em", rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


def reference_cpu_steps(patch, steps, warmup, threads):
    """The reference's OWN modules (UNet3d + DiceLoss from baseline/_ref, unmodified) on the host cores, fp32, the trainer's
    step (default_trainer.py:805-831: zero_grad, forward, loss, backward, AdamW step).  None if the package is absent."""
    import torch
    ru, rd = load_reference_modules()
    if ru is None:
        return None
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = ru.UNet3d(**MODEL_KW)
    loss_fn = rd.DiceLoss()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    x, t = synthetic_batch(1, patch, 0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = loss_fn(model(x), t)
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    vox = patch[0] * patch[1] * patch[2]
    return vox * len(times) / sum(times), sum(times) / len(times)


def gpu_reference_steps(dev, batch, patch, steps, warmup, budget_s=150.0):
    """The reference's OWN UNet3d + DiceLoss (baseline/_ref, unmodified torch.nn modules -> cuDNN / ATen) on the SAME GPU in
    the SAME run, bf16 autocast like the trainer (default_trainer.py:789-794): the denominator of north_star's ">= 1.5x the
    reference's cuDNN train step".  Variants: as written (NCDHW), channels_last_3d, torch.compile (the trainer's default,
    default_trainer.py:541).  Each variant is bounded in wall time; a variant that fails is reported as an error string."""
    import torch
    ru, rd = load_reference_modules()
    if ru is None:
        return {"unavailable": "baseline/_ref (pip install --target of the reference) not present"}
    torch.backends.cudnn.benchmark = True
    x, t = synthetic_batch(batch, patch, seed=1)
    x, t = x.to(dev), t.to(dev)
    vox = batch * patch[0] * patch[1] * patch[2]
    out = {"source": "baseline/_ref/torch_em/model/unet.py + loss/dice.py (unmodified reference modules), torch.nn -> cuDNN, "
                     "bf16 autocast, cudnn.benchmark=True, same batch/patch, same box, same run",
           "steps": steps, "warmup": warmup}

    def run(variant):
        torch.manual_seed(0)
        model = ru.UNet3d(**MODEL_KW).to(dev)
        xx = x
        if variant == "channels_last_3d":
            model = model.to(memory_format=torch.channels_last_3d)
            xx = x.contiguous(memory_format=torch.channels_last_3d)
        fwd = torch.compile(model) if variant == "compiled" else model
        loss_fn = rd.DiceLoss()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3)

        def step():
            opt.zero_grad()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = loss_fn(fwd(xx), t)
            loss.backward()
            opt.step()
            return loss

        t0 = time.perf_counter()
        for _ in range(warmup):
            step()
            torch.cuda.synchronize()
            if time.perf_counter() - t0 > budget_s:
                raise TimeoutError(f"warm-up exceeded {budget_s:.0f} s")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    best = None
    for variant in ("eager_ncdhw", "channels_last_3d", "compiled"):
        try:
            ms = run(variant)
            out[variant + "_ms"] = ms
            best = ms if best is None else min(best, ms)
        except Exception as e:  # noqa: BLE001 -- a failing variant must not lose the bench line
            out[variant + "_error"] = f"{type(e).__name__}: {str(e)[:200]}"
        torch.cuda.empty_cache()
    if best is not None:
        out["best_ms"] = best
        out["value"] = vox / (best * 1e-3)
        out["unit"] = "voxels/s"
    return out


def cpu_reference_steps(patch, steps, warmup, threads):
    """The reference's arithmetic for this path on the host cores: oracle U-Net + Dice + AdamW, fp32."""
    import torch
    from oracle import dice as odice
    from oracle import unet as ounet
    torch.set_num_threads(threads)
    sd = ounet.init_state_dict(1, 2, [2] * MODEL_KW["depth"], initial_features=MODEL_KW["initial_features"], seed=0)
    params = [v.requires_grad_(True) for v in sd.values()]
    opt = torch.optim.AdamW(params, lr=1e-3)
    x, t = synthetic_batch(1, patch, 0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        y = ounet.unet3d_forward(x, sd, [2] * MODEL_KW["depth"], final_activation="Sigmoid")
        loss = odice.dice_loss(y, t)
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    vox = patch[0] * patch[1] * patch[2]
    return vox * len(times) / sum(times), sum(times) / len(times)


def workload_name(batch, patch):
    return (f"UNet3d(1,2,depth=4,initial_features=32,Sigmoid)+DiceLoss+AdamW train step, ({batch},1,{patch[0]},{patch[1]},{patch[2]}) "
            "per GPU (configs[1])")


def cpu_arm(patch, steps, warmup):
    """CPU baseline on the host cores: the unmodified reference modules when baseline/_ref is present ("reference"), else the
    oracle port ("port").  One (1,1,*patch) sample of the bench's batch per step."""
    threads = os.cpu_count() or 1
    r = reference_cpu_steps(patch, steps, warmup, threads)
    kind = "reference"
    if r is None:
        r = cpu_reference_steps(patch, steps, warmup, threads)
        kind = "port"
    vps, sec = r
    impl = "torch_em.model.UNet3d + torch_em.loss.DiceLoss from baseline/_ref (unmodified)" if kind == "reference" else "oracle port"
    sample = (f"{steps} train steps on ONE (1,1,{patch[0]},{patch[1]},{patch[2]}) patch of the bench batch (same model, fp32, "
              f"{impl}, {threads} host threads, {sec:.2f} s/step)")
    return vps, sec, {"value": vps, "unit": "voxels/s", "cores": threads, "kind": kind, "sample": sample}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    patch = tuple(args.patch)
    # bounded sample: one patch of the batch per step; the step count is capped so the arm ends within a few minutes
    steps, warmup = min(args.steps, 10), min(args.warmup, 2)
    vps, sec, cpu = cpu_arm(patch, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": vps, "unit": "voxels/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.batch, patch), "sample": cpu["sample"],
                   "global_batch": args.batch * args.gpus, "parallelism": f"dp{args.gpus}"},
        "cpu_baseline": cpu,
        "e2e": {"value": vps, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import torch_em_b200 as tb
    from torch_em_b200.backend import default_backend

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    batch, patch = args.batch, tuple(args.patch)
    torch.manual_seed(0)
    model = tb.UNet3d(**MODEL_KW).to(dev)
    loss_fn = tb.DiceLoss()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    if world > 1:
        tb.distributed.broadcast_parameters(model)
        tb.distributed.sync_gradients(model)
    xh, th = synthetic_batch(batch, patch, seed=1 + rank)
    xh, th = xh.pin_memory(), th.pin_memory()
    xd, td = xh.to(dev), th.to(dev)

    def step(x, t):
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pred = model(x)
            loss = loss_fn(pred, t)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step(xd, td)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    tb.reset_launch_count()
    ms = timed(lambda: step(xd, td), args.steps)
    launches = tb.launch_count() // args.steps

    # End to end: every step copies ITS OWN batch from pinned host memory and reads the loss back.  Like a DataLoader with
    # pin_memory + non_blocking copies, the copy of step i+1's batch is issued on a side stream while step i computes, so
    # K timed steps contain K host->device copies (the first batch is staged before the timed region, the copy issued by
    # the last step belongs to the step after it).
    copy_stream = torch.cuda.Stream(device=dev)
    staged = {}

    def stage_next():
        with torch.cuda.stream(copy_stream):
            staged["x"] = xh.to(dev, non_blocking=True)
            staged["t"] = th.to(dev, non_blocking=True)

    def e2e_step():
        cur = torch.cuda.current_stream(dev)
        cur.wait_stream(copy_stream)
        x, t = staged["x"], staged["t"]
        x.record_stream(cur)
        t.record_stream(cur)
        stage_next()
        return step(x, t).item()

    stage_next()
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.finish() if sampler else None

    # per-kernel-family device time (CUDA events around every conv launch) for the roofline of the dominant kernel
    B = default_backend()
    B.start_timing()
    nroof = 2
    for _ in range(nroof):
        step(xd, td)
    torch.cuda.synchronize()
    fam = B.stop_timing()

    vox = batch * patch[0] * patch[1] * patch[2]
    value = world * vox / (ms * 1e-3)
    e2e = world * vox / (ms_e2e * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tens_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    def ncu_traffic(kernel_substr):
        """DRAM bytes (read + write) of one launch of the kernel from the committed `ncu --set full` capture, or None."""
        try:
            import csv
            rows = list(csv.reader(open(os.path.join(ROOT, NCU_FULL_CSV))))
            hdr, units = rows[0], rows[1]
            ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            for r in rows[2:]:
                if kernel_substr in r[ik]:
                    return float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        except Exception:
            pass
        return None

    roofline = None
    if fam:
        # fam: "<kernel>:<direction>" -> (launches, ms, flops) over nroof steps.  The dominant KERNEL (one __global__ function)
        # is the roofline subject; families_ms_per_step lists every conv kernel.
        kern = {}
        for k, (n, tot_ms, work) in fam.items():
            name = KERNEL_OF.get(k, k)
            a_ = kern.setdefault(name, [0, 0.0, 0.0])
            a_[0] += n; a_[1] += tot_ms; a_[2] += work
        top = max(kern, key=lambda k: kern[k][1])
        n, tot_ms, work = kern[top]
        achieved = work / (tot_ms * 1e-3) / 1e12
        traffic = ncu_traffic(top)
        roofline = {"bound": "tensor", "kernel": top, "achieved": achieved, "peak": tens_peak, "unit": "TFLOP/s",
                    "frac": achieved / tens_peak,
                    # per-launch DRAM bytes (read + write) of this kernel's LARGEST launch from the committed ncu --set full
                    # capture (not measured in this run); achieved / avg_launch_ms average over all of its launches
                    "traffic": traffic, "traffic_source": NCU_FULL_CSV if traffic is not None else None,
                    "launches_per_step": n // nroof,
                    "avg_launch_ms": tot_ms / n, "share_of_step": tot_ms / nroof / ms, "peak_source": peak_src,
                    "families_ms_per_step": {k: v[1] / nroof for k, v in sorted(fam.items())},
                    "families_tflops": {k: v[2] / (v[1] * 1e-3) / 1e12 for k, v in sorted(fam.items()) if v[1] > 0}}
    cpu = None
    gpu_ref = None
    if world == 1 and not args.no_cpu_baseline:
        _, _, cpu = cpu_arm((64, 64, 64) if args.quick_cpu else patch, 3 if args.quick_cpu else 4, 1)
    if world == 1 and not args.no_gpu_reference:
        del xd, td
        torch.cuda.empty_cache()
        gpu_ref = gpu_reference_steps(dev, batch, patch, steps=5, warmup=3)
    from torch_em_b200.util.flops import conv_flops_train
    flops = conv_flops_train(1, 2, [2] * MODEL_KW["depth"], patch, batch, MODEL_KW["initial_features"])
    line = {
        "metric": METRIC, "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(batch, patch),
                   "e2e": "per step: H2D of the batch (pinned, side stream, overlapped with the previous step) + train step + loss.item()",
                   "global_batch": batch * world, "parallelism": f"dp{world}", "l2": "inputs and activations larger than L2 (no flush needed)",
                   "conv_tflop_per_step": flops / 1e12, "step_tflops": flops / (ms * 1e-3) / 1e12},
        "e2e": {"value": e2e, "unit": "voxels/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": (xh.numel() + th.numel()) * 4,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "gpu_reference": gpu_ref,
        "vs_gpu_reference": (value / gpu_ref["value"]) if gpu_ref and gpu_ref.get("value") else None,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--patch", type=int, nargs=3, default=list(PATCH))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the same-box cuDNN arm (reference torch.nn modules)")
    ap.add_argument("--quick-cpu", action="store_true", help="CPU baseline on a 64^3 patch instead of one patch of the batch")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
