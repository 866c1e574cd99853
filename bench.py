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

CREMI_OFFSETS = [[-1, 0, 0], [0, -1, 0], [0, 0, -1], [-2, 0, 0], [0, -3, 0], [0, 0, -3],
                 [-3, 0, 0], [0, -9, 0], [0, 0, -9], [-4, 0, 0], [0, -27, 0], [0, 0, -27]]      # cli.py:85-90
# BASELINE.json configs.  "cfg2" (configs[1]) is the one the metric is quoted on and what the driver runs; the others are
# selected with --config and are builder-run evidence (profiles/).
CONFIGS = {
    "cfg2": dict(label="configs[1]", model="UNet3d",
                 model_kw=dict(in_channels=1, out_channels=2, depth=4, initial_features=32, final_activation="Sigmoid"),
                 batch=4, patch=(128, 128, 128), dtype="bf16", loss="dice",
                 metric="UNet3d train voxels/sec on (B,1,128,128,128)"),
    "cfg3": dict(label="configs[2]", model="AnisotropicUNet",
                 model_kw=dict(in_channels=1, out_channels=12, scale_factors=[[1, 2, 2], [1, 2, 2], [2, 2, 2], [2, 2, 2]],
                               initial_features=32, final_activation="Sigmoid"),
                 batch=2, patch=(64, 256, 256), dtype="bf16", loss="affinity",
                 metric="AnisotropicUNet + affinity loss train voxels/sec on (2,1,64,256,256)"),
    "cfg4": dict(label="configs[3]", model="UNet3d",
                 model_kw=dict(in_channels=1, out_channels=2, depth=5, initial_features=64, final_activation="Sigmoid"),
                 batch=1, patch=(128, 128, 128), dtype="f32", loss="boundary",
                 metric="UNet3d depth=5 f=64 fp32 boundary-target train voxels/sec on (1,1,128,128,128) per GPU"),
    "cfg5": dict(label="configs[4]", model="UNet3d",
                 model_kw=dict(in_channels=1, out_channels=2, depth=4, initial_features=32, final_activation="Sigmoid"),
                 volume=(512, 512, 512), block_shape=(128, 128, 128), halo=(32, 32, 32), dtype="bf16",
                 metric="UNet3d tiled prediction output voxels/sec on a (512,512,512) volume, blocks 128^3 + halo 32"),
}
MODEL_KW = CONFIGS["cfg2"]["model_kw"]
BATCH, PATCH = CONFIGS["cfg2"]["batch"], CONFIGS["cfg2"]["patch"]
METRIC = CONFIGS["cfg2"]["metric"]
NCU_FULL_CSV = "profiles/r02_kernels_ncu_full.csv"
# backend launch label -> __global__ function it launches (csrc/)
KERNEL_OF = {"first:fwd": "conv3d_first_kernel", "first:wgrad": "conv3d_first_wgrad_kernel", "ds:fwd": "conv3d_umma_ds_kernel",
             "ds:dgrad": "conv3d_umma_ds_kernel", "plain:fwd": "conv3d_umma_kernel", "plain:dgrad": "conv3d_umma_kernel",
             "thin:fwd": "conv3d_umma_kernel", "cs:wgrad": "conv3d_wgrad_cs_kernel", "umma:wgrad": "conv3d_wgrad_umma_kernel",
             "thin:wgrad": "conv3d_wgrad_umma_kernel", "direct:fwd": "conv3d_direct_kernel", "direct:dgrad": "conv3d_direct_kernel",
             "direct:wgrad": "conv3d_wgrad_direct_kernel", "smallcin:wgrad": "conv3d_wgrad_smallcin_kernel",
             "tf32:fwd": "conv3d_umma_kernel<float> (TF32)", "tf32:dgrad": "conv3d_umma_kernel<float> (TF32)",
             "split3:wgrad": "split_bf16 + 3 x conv3d_wgrad_{cs,umma}_kernel (bf16 hi/lo)",
             "h16:fwd": "conv3d_umma_kernel<__half, float", "h16:dgrad": "conv3d_umma_kernel<__half, float",     # (substrings of the ncu names)
             "h16:wgrad": "conv3d_wgrad_{cs,umma}_kernel (h16: fp16 operands)", "h16ds:fwd": "conv3d_umma_ds_kernel"}


def synthetic_batch(batch, patch, seed, kind="dice", out_channels=2):
    """Per-sample standardised random volume (like transform/raw.py:40-65) and the host-side target of the config:
    "dice": binary (B, C, ...) float targets; "affinity" / "boundary": int64 instance labels (B, ...) -- a jittered block
    labelling with ~10 % background -- from which the targets are computed on the GPU inside the step."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((batch, 1) + tuple(patch), generator=g)
    x = (x - x.mean(dim=(1, 2, 3, 4), keepdim=True)) / (x.std(dim=(1, 2, 3, 4), keepdim=True) + 1e-7)
    if kind == "dice":
        t = (torch.rand((batch, out_channels) + tuple(patch), generator=g) > 0.5).float()
    else:
        D, H, W = patch
        gd, gh, gw = max(D // 16, 1), max(H // 32, 1), max(W // 32, 1)
        ids = torch.randperm(gd * gh * gw, generator=g).reshape(gd, gh, gw) + 1
        ids[torch.rand(ids.shape, generator=g) < 0.1] = 0
        lab = ids.repeat_interleave(-(-D // gd), 0)[:D].repeat_interleave(-(-H // gh), 1)[:, :H].repeat_interleave(-(-W // gw), 2)[:, :, :W]
        t = torch.stack([torch.roll(lab, shifts=(b, 3 * b, 5 * b), dims=(0, 1, 2)) for b in range(batch)]).to(torch.int64).contiguous()
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
    """The reference's torch-only modules model/unet.py, loss/dice.py and loss/wrapper.py, loaded by file path (``import
    torch_em`` needs imageio/skimage/... which this image lacks).  -> (unet, dice, wrapper) or (None, None, None)."""
    root = reference_root()
    if root is None:
        return None, None, None
    import importlib.util
    mods = []
    for name, rel in (("_ref_unet", "model/unet.py"), ("_ref_dice", "loss/dice.py"), ("_ref_wrapper", "loss/wrapper.py")):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, "torch_em", rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


def reference_model_and_loss(cfg):
    """The reference's own model and loss for a config (unmodified classes).  The affinity / boundary targets of cfg3 / cfg4 are
    produced in the reference by CPU dataloader workers (label.py); for the arms below they are precomputed once and kept
    next to the input, i.e. the reference arm is NOT charged for them."""
    ru, rd, rw = load_reference_modules()
    if ru is None:
        return None, None
    model = getattr(ru, cfg["model"])(**cfg["model_kw"])
    loss = rd.DiceLoss()
    if cfg["loss"] == "affinity":
        loss = rw.LossWrapper(rd.DiceLoss(), transform=rw.ApplyAndRemoveMask(masking_method="multiply"))   # cli.py:263-267
    return model, loss


def reference_targets(cfg, labels_or_target, device):
    """Float targets the reference's loss consumes, computed once with the GPU label kernels (outside every timed region)."""
    import torch
    import torch_em_b200 as tb
    if cfg["loss"] == "dice":
        return labels_or_target.to(device)
    if not torch.cuda.is_available():
        from oracle import labels as olabels                                  # CPU reference arm: numpy restatement
        lab = labels_or_target.numpy()
        if cfg["loss"] == "affinity":
            t = [olabels.affinity_targets(l, CREMI_OFFSETS, ignore_label=0, add_mask=True) for l in lab]
        else:
            t = [olabels.boundary_targets(l, add_binary_target=True) for l in lab]
        import numpy as np
        return torch.from_numpy(np.stack(t)).to(device)
    lab = labels_or_target.to("cuda")
    if cfg["loss"] == "affinity":
        t = tb.AffinityTransform(CREMI_OFFSETS, ignore_label=0, add_mask=True)(lab)
    else:
        t = tb.BoundaryTransform(add_binary_target=True)(lab)
    return t.to(device)


def reference_cpu_steps(cfg, patch, steps, warmup, threads):
    """The reference's OWN modules (baseline/_ref, unmodified) on the host cores, fp32, the trainer's step
    (default_trainer.py:805-831: zero_grad, forward, loss, backward, AdamW step).  None if the package is absent."""
    import torch
    model, loss_fn = reference_model_and_loss(cfg)
    if model is None:
        return None
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    x, t = synthetic_batch(1, patch, 0, cfg["loss"], cfg["model_kw"]["out_channels"])
    t = reference_targets(cfg, t, "cpu")
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


def gpu_reference_steps(cfg, dev, batch, patch, steps, warmup, budget_s=150.0):
    """The reference's OWN model + loss (baseline/_ref, unmodified torch.nn modules -> cuDNN / ATen) on the SAME GPU in the SAME
    run, under the precision the trainer would use for the config (bf16 autocast, default_trainer.py:789-794; fp32 with
    torch's default TF32 convolutions for cfg4): the denominator of north_star's ">= 1.5x the reference's cuDNN train step".
    Variants: as written (NCDHW), channels_last_3d, torch.compile (the trainer's default, default_trainer.py:541).  Each variant
    is bounded in wall time; a variant that fails is reported as an error string."""
    import torch
    model0, _ = reference_model_and_loss(cfg)
    if model0 is None:
        return {"unavailable": "baseline/_ref (pip install --target of the reference) not present"}
    del model0
    torch.backends.cudnn.benchmark = True
    bf16 = cfg["dtype"] == "bf16"
    x, t = synthetic_batch(batch, patch, 1, cfg["loss"], cfg["model_kw"]["out_channels"])
    x = x.to(dev)
    t = reference_targets(cfg, t, dev)
    vox = batch * patch[0] * patch[1] * patch[2]
    out = {"source": f"baseline/_ref torch_em.model.{cfg['model']} + reference loss (unmodified modules), torch.nn -> cuDNN, "
                     + ("bf16 autocast" if bf16 else "fp32 (TF32 convolutions: torch default)")
                     + ", cudnn.benchmark=True, same batch/patch, same box, same run; targets precomputed (not charged)",
           "steps": steps, "warmup": warmup}

    def run(variant):
        torch.manual_seed(0)
        model, loss_fn = reference_model_and_loss(cfg)
        model = model.to(dev)
        xx = x
        if variant == "channels_last_3d":
            model = model.to(memory_format=torch.channels_last_3d)
            xx = x.contiguous(memory_format=torch.channels_last_3d)
        fwd = torch.compile(model) if variant == "compiled" else model
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3)

        def step():
            opt.zero_grad()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
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


def oracle_cpu_steps(patch, steps, warmup, threads):
    """Fallback CPU arm when baseline/_ref is absent: the oracle port of the cfg2 step (oracle U-Net + Dice + AdamW, fp32)."""
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


def workload_name(cfg, batch, patch):
    kw = ",".join(f"{k}={v}" for k, v in cfg["model_kw"].items() if k not in ("in_channels", "final_activation"))
    loss = {"dice": "DiceLoss", "affinity": "affinity loss (masked Dice on 12-offset affinity targets from labels)",
            "boundary": "BoundaryTransform(add_binary_target) targets + DiceLoss"}[cfg["loss"]]
    return (f"{cfg['model']}({kw},Sigmoid)+{loss}+AdamW train step, ({batch},1,{patch[0]},{patch[1]},{patch[2]}) per GPU "
            f"({cfg['label']})")


def cpu_arm(cfg, patch, steps, warmup):
    """CPU baseline on the host cores: the unmodified reference modules when baseline/_ref is present ("reference"), else the
    oracle port ("port", cfg2 only).  One (1,1,*patch) sample of the bench's batch per step."""
    threads = os.cpu_count() or 1
    r = reference_cpu_steps(cfg, patch, steps, warmup, threads)
    kind = "reference"
    if r is None:
        r = oracle_cpu_steps(patch, steps, warmup, threads)
        kind = "port"
    vps, sec = r
    impl = f"torch_em.model.{cfg['model']} + reference loss from baseline/_ref (unmodified)" if kind == "reference" else "oracle port"
    sample = (f"{steps} train steps on ONE (1,1,{patch[0]},{patch[1]},{patch[2]}) patch of the bench batch (same model, fp32, "
              f"{impl}, {threads} host threads, {sec:.2f} s/step)")
    return vps, sec, {"value": vps, "unit": "voxels/s", "cores": threads, "kind": kind, "sample": sample}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    if args.config == "cfg5":
        return run_reference_cfg5(args)
    patch = tuple(args.patch or cfg["patch"])
    batch = args.batch or cfg["batch"]
    # bounded sample: one patch of the batch per step; the step count is capped so the arm ends within a few minutes
    steps, warmup = min(args.steps, 10), min(args.warmup, 2)
    vps, sec, cpu = cpu_arm(cfg, patch, steps, warmup)
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": vps, "unit": "voxels/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg, batch, patch), "sample": cpu["sample"],
                   "global_batch": batch * args.gpus, "parallelism": f"dp{args.gpus}"},
        "cpu_baseline": cpu,
        "e2e": {"value": vps, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def load_peaks():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    return peaks


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


def conv_roofline(fam, nroof, ms, tens_peak, peak_src):
    """fam: "<kernel>:<direction>" -> (launches, ms, flops) over nroof steps.  The dominant KERNEL (one __global__ function) is the
    roofline subject; families_ms_per_step lists every conv kernel."""
    if not fam:
        return None
    kern = {}
    for k, (n, tot_ms, work) in fam.items():
        a_ = kern.setdefault(KERNEL_OF.get(k, k), [0, 0.0, 0.0])
        a_[0] += n; a_[1] += tot_ms; a_[2] += work
    top = max(kern, key=lambda k: kern[k][1])
    n, tot_ms, work = kern[top]
    achieved = work / (tot_ms * 1e-3) / 1e12
    traffic = ncu_traffic(top)
    return {"bound": "tensor", "kernel": top, "achieved": achieved, "peak": tens_peak, "unit": "TFLOP/s",
            "frac": achieved / tens_peak,
            # per-launch DRAM bytes (read + write) of ONE representative launch of this kernel from the committed ncu --set full
            # capture (shape in that file's first column; not measured in this run); achieved / avg_launch_ms average all launches
            "traffic": traffic, "traffic_source": NCU_FULL_CSV if traffic is not None else None,
            "launches_per_step": n // nroof, "avg_launch_ms": tot_ms / n, "share_of_step": tot_ms / nroof / ms,
            "peak_source": peak_src,
            "families_ms_per_step": {k: v[1] / nroof for k, v in sorted(fam.items())},
            "families_tflops": {k: v[2] / (v[1] * 1e-3) / 1e12 for k, v in sorted(fam.items()) if v[1] > 0}}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import torch_em_b200 as tb
    from torch_em_b200.backend import default_backend

    cfg = CONFIGS[args.config]
    if args.config == "cfg5":
        return run_ours_cfg5(args)
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

    batch, patch = args.batch or cfg["batch"], tuple(args.patch or cfg["patch"])
    bf16 = cfg["dtype"] == "bf16"
    torch.manual_seed(0)
    model = getattr(tb, cfg["model"])(**cfg["model_kw"]).to(dev)
    if cfg["loss"] == "affinity":
        loss_fn = tb.AffinityLoss(CREMI_OFFSETS, ignore_label=0)          # target + mask computed inside the loss kernels
    else:
        loss_fn = tb.DiceLoss()
    boundary = tb.BoundaryTransform(add_binary_target=True) if cfg["loss"] == "boundary" else None
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    if world > 1:
        tb.distributed.broadcast_parameters(model)
        tb.distributed.sync_gradients(model)
    xh, th = synthetic_batch(batch, patch, 1 + rank, cfg["loss"], cfg["model_kw"]["out_channels"])
    xh, th = xh.pin_memory(), th.pin_memory()
    xd, td = xh.to(dev), th.to(dev)

    def step(x, t):
        opt.zero_grad()
        if boundary is not None:
            t = boundary(t)                                               # labels -> [foreground, boundary] on the GPU, every step
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
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

    peaks = load_peaks()
    if bf16:
        tens_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    elif any(k.startswith("h16") for k in fam):
        tens_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured sustained 16-bit tensor rate (fp32 activations run the h16 path: fp16 operand copies, kind::f16 MMAs, fp32 accumulate)"
    else:
        tens_peak = peaks.get("bf16_tflops_sustained", 1400.0) / 2
        peak_src = "half of the measured sustained bf16 rate (TF32 tensor peak = bf16 / 2; fp32 runs the TF32 tensor path)"
    roofline = conv_roofline(fam, nroof, ms, tens_peak, peak_src)
    cpu = None
    gpu_ref = None
    if world == 1 and not args.no_cpu_baseline:
        _, _, cpu = cpu_arm(cfg, (64, 64, 64) if args.quick_cpu else patch, 3 if args.quick_cpu else 4, 1)
    if world == 1 and not args.no_gpu_reference:
        del xd, td
        torch.cuda.empty_cache()
        gpu_ref = gpu_reference_steps(cfg, dev, batch, patch, steps=5, warmup=3)
    from torch_em_b200.util.flops import conv_flops_train
    mk = cfg["model_kw"]
    sfs = mk.get("scale_factors", [2] * mk.get("depth", 4))
    flops = conv_flops_train(1, mk["out_channels"], sfs, patch, batch, mk["initial_features"])
    line = {
        "metric": cfg["metric"], "value": value, "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
        "config": {"workload": workload_name(cfg, batch, patch),
                   "e2e": "per step: H2D of the batch (pinned, side stream, overlapped with the previous step) + train step + loss.item()",
                   "global_batch": batch * world, "parallelism": f"dp{world}", "l2": "inputs and activations larger than L2 (no flush needed)",
                   "conv_tflop_per_step": flops / 1e12, "step_tflops": flops / (ms * 1e-3) / 1e12,
                   "precision": ("bf16 autocast: bf16 activations and operands, fp32 accumulation / statistics / parameters" if bf16 else
                                 "fp32 activations, statistics, parameters and gradients; convolutions TF32-class, as torch / cuDNN run fp32 "
                                 "convolutions by default (allow_tf32=True, what the reference arm uses): fp16 operand copies (TF32's 11-bit "
                                 "significand, exact power-of-two range scaling per tensor), fp32 accumulation")},
        "e2e": {"value": e2e, "unit": "voxels/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": xh.numel() * xh.element_size() + th.numel() * th.element_size(), "d2h_bytes_per_step": 4},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "gpu_reference": gpu_ref,
        "vs_gpu_reference": (value / gpu_ref["value"]) if gpu_ref and gpu_ref.get("value") else None,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---- cfg5: tiled inference ----------------------------------------------------------------------------------------------------
def cfg5_volume(shape, seed=0):
    import numpy as np
    return np.random.default_rng(seed).random(shape, dtype="float32")


def reference_predict_with_halo():
    """The reference's own ``predict_with_halo`` (baseline/_ref, unmodified) with ``bioimage_cpp.utils.Blocking`` -- a third-party
    class absent from this image -- supplied by our restatement of it (a C-order block grid; it only enumerates the blocks)."""
    root = reference_root()
    if root is None:
        return None, None
    sys.path.insert(0, ROOT)
    from tests import ref_harness
    te = ref_harness.import_torch_em()
    import bioimage_cpp
    from torch_em_b200.util import Blocking
    bioimage_cpp.utils.Blocking = Blocking
    import torch_em.util.prediction as rp
    return te, rp


def run_reference_cfg5(args):
    """CPU arm of cfg5: the reference's predict_with_halo + UNet3d on the host cores, on a bounded sample (ONE haloed block)."""
    import numpy as np
    import torch
    cfg = CONFIGS["cfg5"]
    te, rp = reference_predict_with_halo()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    bs, halo = cfg["block_shape"], cfg["halo"]
    vol = cfg5_volume(tuple(b for b in bs))                           # one inner block; its halo is mirrored
    torch.manual_seed(0)
    import torch_em.model.unet as ru
    model = ru.UNet3d(**cfg["model_kw"]).eval()
    t0 = time.perf_counter()
    out = rp.predict_with_halo(vol, model, ["cpu"], bs, halo, disable_tqdm=True)
    sec = time.perf_counter() - t0
    vps = float(np.prod(bs)) / sec
    sample = (f"ONE {bs} block + halo {halo} of the volume (1 of 64), reference predict_with_halo + torch_em.model.UNet3d from "
              f"baseline/_ref (unmodified), fp32, {threads} host threads, {sec:.1f} s")
    line = {"impl": "reference", "metric": cfg["metric"], "value": vps, "unit": "voxels/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["metric"], "sample": sample},
            "cpu_baseline": {"value": vps, "unit": "voxels/s", "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": vps, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "out_shape": list(out.shape)}
    print(json.dumps(line), flush=True)


def run_ours_cfg5(args):
    """Tiled prediction of a 512^3 volume (64 blocks of 128^3 + halo 32 -> 192^3 network inputs).  A "step" is one whole-volume
    prediction through the public API from a HOST numpy array to a HOST numpy array (the H2D of the volume and the D2H of the
    result inside the timed region, overlapped with the block loop on side streams); ``value`` = the device work only (CUDA
    events on the compute stream from the arrival of the first batch's rows to the last scatter)."""
    import numpy as np
    import torch
    import torch_em_b200 as tb
    from torch_em_b200.backend import default_backend
    from torch_em_b200.util import predict_with_halo
    cfg = CONFIGS["cfg5"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    shape, bs, halo = tuple(args.patch or cfg["volume"]), cfg["block_shape"], cfg["halo"]
    vol = cfg5_volume(shape)
    torch.manual_seed(0)
    model = tb.UNet3d(**cfg["model_kw"]).to(dev).eval()
    n_vox = float(np.prod(shape))
    steps, warmup = max(1, min(args.steps, 5)), max(2, min(args.warmup, 3))

    def run():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return predict_with_halo(vol, model, [0], bs, halo, disable_tqdm=True)

    from torch_em_b200.util import prediction as P
    for _ in range(warmup):
        out = run()
    sampler = ClockSampler(0)
    sampler.start()
    tb.reset_launch_count()
    B = default_backend()
    B.start_timing()
    loop_ms, h2d_ms, d2h_ms = [], [], []
    t0 = time.perf_counter()
    for _ in range(steps):
        out = run()
        loop_ms.append(P.last_timing["loop_ms"]); h2d_ms.append(P.last_timing["h2d_ms"]); d2h_ms.append(P.last_timing["d2h_ms"])
    torch.cuda.synchronize()
    sec_e2e = (time.perf_counter() - t0) / steps
    fam = B.stop_timing()
    launches = tb.launch_count() // steps
    clocks = sampler.finish()
    conv_ms = sum(v[1] for v in fam.values()) / steps
    conv_flops = sum(v[2] for v in fam.values()) / steps
    # device-resident figure: the block loop alone (CUDA events inside predict_with_halo, between the H2D of the volume and the
    # D2H of the result)
    ms_dev = sum(loop_ms) / steps
    h2d, d2h = vol.nbytes, out.nbytes
    peaks = load_peaks()
    tens_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    roofline = conv_roofline(fam, steps, ms_dev, tens_peak, "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback")
    # I/O kernels alone: gather + standardize + scatter of one block, HBM GB/s against the measured copy bandwidth
    io = cfg5_io_kernels(dev, vol, bs, halo, out.shape[0], peaks.get("hbm_gbs", 6500.0))
    gpu_ref = None
    if not args.no_gpu_reference:
        gpu_ref = cfg5_gpu_reference(cfg, dev, vol, bs, halo)
    kept_frac = float(np.prod(bs)) / float(np.prod([b_ + 2 * h_ for b_, h_ in zip(bs, halo)]))
    line = {
        "metric": cfg["metric"], "value": n_vox / (ms_dev * 1e-3), "unit": "voxels/s", "n_gpus": 1, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"UNet3d(depth=4,initial_features=32,Sigmoid) inference, predict_with_halo over a {shape} float32 volume, "
                               f"block {bs} halo {halo} ({cfg['label']})",
                   "kept_fraction_of_computed_voxels": kept_frac, "blocks": int(np.prod([-(-s // b_) for s, b_ in zip(shape, bs)])),
                   "value": "block loop on the device (gather, standardize, forward, scatter), volume and output resident in HBM",
                   "e2e": "host numpy volume -> predict_with_halo -> host numpy result; the volume goes up in row ranges ahead of the blocks "
                          "that read them and finished output rows go down behind them (side streams): h2d_ms / d2h_ms are the "
                          "EXPOSED parts (first batch's rows up, last rows down)",
                   "l2": "volume, activations and output larger than L2",
                   "conv_tflop_per_step": conv_flops / 1e12, "conv_ms_per_step": conv_ms},
        "e2e": {"value": n_vox / sec_e2e, "unit": "voxels/s", "ms_per_step": sec_e2e * 1e3, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "h2d_ms": sum(h2d_ms) / steps, "d2h_ms": sum(d2h_ms) / steps, "block_loop_ms": ms_dev,
                "blocks_per_forward": P.last_timing.get("batch_size")},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "io_kernels": io, "cpu_baseline": None,
        "gpu_reference": gpu_ref,
        "vs_gpu_reference": (n_vox / sec_e2e / gpu_ref["value"]) if gpu_ref and gpu_ref.get("value") else None,
    }
    print(json.dumps(line), flush=True)


def cfg5_io_kernels(dev, vol, bs, halo, n_out, hbm_peak):
    """CUDA-event timing of the three tiling kernels on ONE block (inputs far larger than what stays in L2 between launches
    because the volume is 512 MB): algorithmic bytes / time against the measured HBM copy bandwidth."""
    import ctypes
    import numpy as np
    import torch
    from torch_em_b200._lib import call
    big = tuple(b + 2 * h for b, h in zip(bs, halo))
    nvox_big, nvox_in = int(np.prod(big)), int(np.prod(bs))
    dv = torch.from_numpy(vol[None]).to(dev)
    inp = torch.empty((1, 1) + big, dtype=torch.float32, device=dev)
    stats = torch.zeros((1, 2), dtype=torch.float64, device=dev)
    pred = torch.rand((1, n_out) + big, dtype=torch.float32, device=dev)
    outv = torch.zeros((n_out,) + vol.shape, dtype=torch.float32, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())                              # noqa: E731
    begin = (ctypes.c_int * 3)(*[bs[i] - halo[i] for i in range(3)])
    obeg, oshp = (ctypes.c_int * 3)(*bs), (ctypes.c_int * 3)(*bs)
    D, H, W = vol.shape
    runs = {
        "gather_blocks": (lambda: call("b200em_gather_blocks", vp(dv), 7, 1, D, H, W, begin, 1, *big, vp(inp), vp(stats), st), nvox_big * 8),
        "standardize_blocks": (lambda: call("b200em_standardize_blocks", vp(inp), 1, nvox_big, vp(stats), 1e-7, st), nvox_big * 8),
        "scatter_blocks": (lambda: call("b200em_scatter_blocks", vp(pred), n_out, *big, *halo, obeg, oshp, 1, vp(outv), D, H, W, 0, n_out,
                                        None, st), nvox_in * n_out * 8),
    }
    res = {}
    for name, (fn, nbytes) in runs.items():
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 100.0
        res[name] = {"us": us, "algorithmic_bytes": nbytes, "gbs": nbytes / us / 1e3, "frac_of_hbm": nbytes / us / 1e3 / hbm_peak}
    return res


def cfg5_gpu_reference(cfg, dev, vol, bs, halo):
    """The reference's own predict_with_halo + UNet3d (unmodified, baseline/_ref) on the same GPU: fp32 (TF32 convolutions, torch
    default -- the reference has no autocast at inference, prediction.py:252-275) and under bf16 autocast."""
    import numpy as np
    import torch
    te, rp = reference_predict_with_halo()
    if rp is None:
        return {"unavailable": "baseline/_ref not present"}
    import torch_em.model.unet as ru
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    model = ru.UNet3d(**cfg["model_kw"]).to(dev).eval()
    out = {"source": "baseline/_ref torch_em.util.prediction.predict_with_halo + torch_em.model.UNet3d (unmodified; Blocking supplied "
                     "by our restatement), same volume, same box, same run, host numpy -> host numpy"}
    best = None
    for name, ctx in (("fp32_tf32", lambda: torch.autocast("cuda", enabled=False)), ("bf16_autocast", lambda: torch.autocast("cuda", dtype=torch.bfloat16))):
        try:
            with ctx():
                rp.predict_with_halo(vol[:256, :256, :256], model, [0], bs, halo, disable_tqdm=True)      # warm-up (cuDNN autotune)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                rp.predict_with_halo(vol, model, [0], bs, halo, disable_tqdm=True)
                torch.cuda.synchronize()
                sec = time.perf_counter() - t0
            out[name + "_s"] = sec
            best = sec if best is None else min(best, sec)
        except Exception as e:  # noqa: BLE001
            out[name + "_error"] = f"{type(e).__name__}: {str(e)[:200]}"
    if best is not None:
        out["best_s"] = best
        out["value"] = float(np.prod(vol.shape)) / best
        out["unit"] = "voxels/s"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS), help="BASELINE.json config (default: configs[1], the metric's)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's)")
    ap.add_argument("--patch", type=int, nargs=3, default=None, help="patch (cfg5: volume) shape (default: the config's)")
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
