#!/usr/bin/env python
"""Benchmark of the DUSty-v2 G+D training step (BASELINE.json metric:
"dusty_v2 G+D train images/sec @64x512").

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU path (oracle port)

One "step" = one training iteration of the reference's Trainer.step (G step, D step, R1 every
16th iteration, EMA, ADA controller) at per-GPU batch 64 on synthetic 64x512 range images with
random-init weights.  N>1 is launched by torchrun (one rank per GPU, NCCL, weak scaling:
per-GPU batch fixed).  Rank 0 prints ONE JSON line.

Timing: W warm-up steps, then exactly K steps bracketed by barrier + synchronize, CUDA events
on the launching stream, max over ranks.  The working set of a step (GBs of activations)
is far larger than the 126 MB L2, so no explicit flush is needed between steps; the isolated
kernel measurements for the roofline block do flush L2 between launches.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "tests"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "dusty_v2 G+D train images/sec @64x512"       # --arch dusty_v1 / vanilla substitute their name
UNIT = "images/s"
H, W = 64, 512


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch-per-gpu", type=int, default=64)
    ap.add_argument("--global-batch", type=int, default=None,
                    help="strong scaling: fix the job's batch and give each GPU global/N of it")
    ap.add_argument("--arch", default="dusty_v2", choices=["dusty_v2", "dusty_v1", "vanilla"])
    ap.add_argument("--ada-p", type=float, default=None, help="pin the ADA probability")
    ap.add_argument("--cpu-batch", type=int, default=2, help="batch of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--per-step", action="store_true", help="print the device time of every timed step (stderr)")
    ap.add_argument("--no-cudnn-benchmark", action="store_true")
    ap.add_argument("--no-cuda-graphs", action="store_true")
    ap.add_argument("--ref-kernels-only", action="store_true",
                    help="time the reference's own CUDA kernels (oracle/_ref) beside ours, print JSON, exit")
    ap.add_argument("--no-ref-kernels", action="store_true")
    ap.add_argument("--ncu-window", action="store_true",
                    help="bracket the timed region with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    return ap.parse_args()


# ------------------------------------------------------------------------------ helpers
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled in-process every
    10 ms (the timed region of the default run is ~0.4 s: `nvidia-smi -lms` delivers two samples
    in that time), nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        self.rows, self.proc, self.nvml = [], None, None
        self.sm, self.mx, self.reasons, self._stop = [], [], set(), False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                for name, bit in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop = True
            self.thread.join(timeout=1.0)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None,
                    "sm_min_mhz": min(self.sm) if self.sm else None,
                    "sm_max_mhz": max(self.mx) if self.mx else None, "samples": len(self.sm),
                    "reasons": sorted(self.reasons), "source": "nvml, 10 ms polling"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons), "source": "nvidia-smi -lms 50"}


def synthetic_batches(n, batch, seed, pinned=False, device=None):
    """Synthetic reals: depth = 1.45 + 78.55*U(0,1), mask ~ Bernoulli(0.85)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        depth = 1.45 + (80 - 1.45) * torch.rand(batch, 1, H, W, generator=g)
        mask = (torch.rand(batch, 1, H, W, generator=g) < 0.85).float()
        if pinned:
            depth, mask = depth.pin_memory(), mask.pin_memory()
        if device is not None:
            depth, mask = depth.to(device), mask.to(device)
        out.append({"depth": depth, "mask": mask})
    return out


def cycle(pool):
    i = 0
    while True:
        yield pool[i % len(pool)]
        i += 1


# ------------------------------------------------------------------------------ CPU baseline
def cpu_reference_run(args, steps, warmup, batch):
    """Times the oracle's restatement of Trainer.step on the host cores.  Returns
    (images_per_s, seconds_per_step, cores, description)."""
    from oracle import dusty_oracle as O
    with O.fir_impl("library"):      # the reference's CPU path runs its FIRs as depthwise convolutions
        return _cpu_reference_run(args, steps, warmup, batch)


def _cpu_reference_run(args, steps, warmup, batch):
    import numpy as np
    import torch

    from dusty_gan_v2_b200.gans.augment.adaptive_augment import AdaptiveAugment
    from dusty_gan_v2_b200.gans.coords import CoordBridge
    from dusty_gan_v2_b200.gans.models.builder import build_discriminator, build_generator
    from dusty_gan_v2_b200.presets import preset
    from oracle import dusty_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    np.random.seed(0)
    cfg = preset(args.arch, batch_size=batch)
    G = build_generator(cfg.model.generator)            # parameter init only (CPU, no kernels)
    D = build_discriminator(cfg.model.discriminator)
    nograd = ("ema_var", "w_avg", "kernel", "pe.", "raydrop_const")
    sdG = {k: v.clone().requires_grad_(not any(t in k for t in nograd)) for k, v in G.state_dict().items()}
    sdD = {k: v.clone().requires_grad_("kernel" not in k) for k, v in D.state_dict().items()}
    cb = CoordBridge(H, W, 1.45, 80.0, os.path.join(ROOT, "data/coords/kitti_raw.npy"))
    angle = cb.angle.repeat_interleave(batch, dim=0)
    ada = AdaptiveAugment(p_init=args.ada_p or 0.0, **cfg.training.augment.policy)
    ada.generator = torch.Generator().manual_seed(7)
    pool = synthetic_batches(2, batch, seed=2)
    g = torch.Generator().manual_seed(1)

    def draws():
        r = {}
        for tag in ("g", "d"):
            r[f"z_{tag}"] = torch.randn(batch, 512, generator=g)
            r[f"shift_{tag}"] = torch.rand(batch, generator=g)
            r[f"u_{tag}"] = torch.rand(batch, 1, H, W, generator=g)
        for tag in ("g_fake", "d_real", "d_fake", "r1"):
            r[f"keep_{tag}"] = torch.bernoulli(torch.full((batch, 1, H, W), 0.5), generator=g)
            r[f"Ginv_{tag}"] = torch.inverse(ada.sample_affine(batch, H, W))
            r[f"C_{tag}"] = ada.sample_color(batch)
        return r

    lazy = 16 / 17.0
    optG = torch.optim.Adam([v for v in sdG.values() if v.requires_grad], lr=0.002, betas=(0.0, 0.99))
    optD = torch.optim.Adam([v for v in sdD.values() if v.requires_grad], lr=0.002 * lazy,
                            betas=(0.0, 0.99 ** lazy))

    def apply(opt, sd, grads):
        for k, gr in grads.items():
            sd[k].grad = gr
        opt.step()
        opt.zero_grad(set_to_none=True)

    def one(it):
        b = pool[it % len(pool)]
        x_real = O.fetch_reals(b["depth"], b["mask"], 1.45, 80.0)
        r = O.train_iteration(sdG, sdD, x_real, angle, draws(), with_r1=(it % 16 == 0), arch=args.arch)
        apply(optG, sdG, r["grads_G"])
        apply(optD, sdD, r["grads_D"])
        if "grads_R1" in r:
            apply(optD, sdD, r["grads_R1"])

    for i in range(warmup):
        one(i + 1)
    per_it = []
    for i in range(steps):
        t0 = time.perf_counter()
        one(i)
        per_it.append((i % 16 == 0, time.perf_counter() - t0))
    r1 = [t for is_r1, t in per_it if is_r1]
    plain = [t for is_r1, t in per_it if not is_r1]
    if steps % 16 == 0 or not r1 or not plain:
        spt = sum(t for _, t in per_it) / steps          # the sample already holds R1 at its 1/16 share
        mix = f"{steps} steps"
    else:
        # short sample: iteration 0 carries the lazy R1 pass; weight it as the training loop does
        spt = (15 * sum(plain) / len(plain) + sum(r1) / len(r1)) / 16
        mix = (f"{len(plain)} plain + {len(r1)} R1 iteration(s) timed, combined at the training loop's "
               f"15:1 ratio ({sum(plain) / len(plain):.2f} s / {sum(r1) / len(r1):.2f} s)")
    desc = (f"oracle port of {args.arch} Trainer.step (G step + D step + R1 on every 16th step, ADA p="
            f"{args.ada_p or 0.0}, warm-up dropout 0.5), fp32, batch {batch}, {mix}, "
            f"Adam updates included (all three phases use the pre-step weights)")
    return batch / spt, spt, cores, desc


# ------------------------------------------------------------------------------ kernel roofline
def kernel_rooflines(device, peaks):
    """Isolated CUDA-event timings of this repo's kernels at the step's real shapes (B=64,
    bf16), L2 flushed between launches.  `achieved` uses ALGORITHMIC bytes / flops (DESIGN.md
    section 4).  Returns (dominant_entry, list_of_entries, conv_family_summary): the dominant
    entry is the kernel with the largest share of the training step's device time (its isolated
    duration x its launches per plain iteration) -- a discriminator convolution."""
    import torch

    import dusty_gan_v2_b200.functional as DF
    from dusty_gan_v2_b200 import _cabi as K
    from dusty_gan_v2_b200.gans.models.ops import Pad, Resample
    B = 64
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    def timeit(fn, reps=7):
        """Median device time of ONE launch: L2 flushed (256 MB write) before EVERY timed launch,
        so `achieved` is against DRAM-cold inputs like the ncu `traffic` figure; a spin kernel
        between the flush and the start event gives the host time to finish enqueueing `fn`
        (Python, tensor-map encoding) before the device reaches it -- without it a 20-50 us
        kernel's event interval measures the host wrapper."""
        fn()
        torch.cuda.synchronize()
        if os.environ.get("DUSTY_KB_ONCE"):      # one launch per kernel: for `ncu --set full`
            return 1.0
        ts = []
        for _ in range(reps):
            flush.zero_()
            torch.cuda._sleep(600_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        ts.sort()
        return ts[len(ts) // 2]

    bf = torch.bfloat16
    hbm = peaks.get("hbm_gbs", 6650.0)
    tflops = peaks.get("bf16_tflops", 1590.0)
    src = "measured" if "hbm_gbs" in peaks else "fallback"
    out = []

    # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the same kernels
    # from the committed `ncu --set full` capture (profiles/ncu_traffic.json, written by
    # tools/ncu_traffic.py from the .ncu-rep); None where the capture has no such launch.
    try:
        traffic_db = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["kernels"]
    except Exception:
        traffic_db = {}

    def traffic_of(name):
        e = traffic_db.get(name)
        return None if e is None else int(e["dram_bytes"])

    def mem_entry(name, fn, bytes_):
        t = timeit(fn)
        out.append({"kernel": name, "bound": "hbm", "achieved": bytes_ / t / 1e9, "peak": hbm,
                    "unit": "GB/s", "frac": bytes_ / t / 1e9 / hbm, "traffic": traffic_of(name),
                    "algorithmic_bytes": bytes_, "ms": t * 1e3, "peak_source": src})

    def tc_entry(name, fn, flops, bytes_):
        """A contraction is reported against the roof that bounds it: the larger of
        flops / bf16 peak and algorithmic bytes / HBM peak (thin layers: O = 32 moves 438 MB
        for 77 GFLOP, i.e. 67 us of HBM time against 47 us of tensor time)."""
        t = timeit(fn)
        tf, gb = flops / t / 1e12, bytes_ / t / 1e9
        hbm_bound = bytes_ / (hbm * 1e9) > flops / (tflops * 1e12)
        e = {"kernel": name, "bound": "hbm" if hbm_bound else "tensor",
             "achieved": gb if hbm_bound else tf, "peak": hbm if hbm_bound else tflops,
             "unit": "GB/s" if hbm_bound else "TFLOP/s",
             "frac": gb / hbm if hbm_bound else tf / tflops, "traffic": traffic_of(name),
             "algorithmic_bytes": bytes_, "algorithmic_flops": flops, "ms": t * 1e3, "peak_source": src,
             "tflops": tf, "tensor_frac": tf / tflops, "hbm_gbs": gb, "hbm_frac": gb / hbm}
        out.append(e)
        return e

    # ---- memory-bound kernels
    x = torch.randn(B, 32, H, W, device=device, dtype=bf)
    b = torch.zeros(32, device=device)
    mem_entry("bias_act_fwd[64,32,64,512]bf16", lambda: DF._bias_act_raw(x, b, None, 3, 0, 0.2, 1.41), 2 * x.numel() * 2)
    y = DF._bias_act_raw(x, b, None, 3, 0, 0.2, 1.41)
    gx = torch.empty_like(x)
    db = torch.zeros(32, device=device)
    mem_entry("bias_act_bwd[64,32,64,512]bf16",
              lambda: K.call("dusty_bias_act_bwd", K.ptr(x), K.ptr(y), K.ptr(gx), K.ptr(db), B, 32, H * W,
                             0.2, 1.41, K.BF16, K.stream_of(x)), 3 * x.numel() * 2)
    up = Resample(up=2).to(device)
    h = torch.randn(B, 64, 32, 256, device=device, dtype=bf)
    mem_entry("resample_up2[64,64,32,256]bf16", lambda: up(h), 5 * h.numel() * 2)
    gup = torch.randn(B, 64, 64, 512, device=device, dtype=bf)
    taps = tuple(up.kernel.tolist())
    mem_entry("resample_up2_adjoint[64,64,64,512]bf16", lambda: DF._resample4_raw(gup, taps, 2, True),
              5 * h.numel() * 2)
    mem_entry("resample_up2+sumsq[64,64,32,256]bf16", lambda: DF.up2_with_sumsq(h, taps), 5 * h.numel() * 2)
    blur = Resample().to(device)
    mem_entry("resample_blur[64,32,64,512]bf16", lambda: blur(x), 2 * x.numel() * 2)
    btaps = tuple(blur.kernel.tolist())
    mem_entry("resample_blur_adjoint[64,32,64,512]bf16", lambda: DF._resample4_raw(x, btaps, 1, True),
              2 * x.numel() * 2)
    pad = Pad(1, ring=True)
    mem_entry("pad_ring1[64,32,64,512]bf16", lambda: pad(x), (x.numel() + B * 32 * 66 * 514) * 2)
    gp = torch.randn(B, 32, 66, 514, device=device, dtype=bf)
    mem_entry("pad_ring1_adjoint[64,32,66,514]bf16",
              lambda: DF._pad_raw(gp, (1, 1, 1, 1), (K.PAD_REPLICATE, K.PAD_CIRCULAR), True, (H, W)),
              (x.numel() + gp.numel()) * 2)
    # NHWC variants (what the discriminator trunk actually runs)
    CL = torch.channels_last
    xc = x.contiguous(memory_format=CL)
    mem_entry("bias_act_fwd_nhwc[64,32,64,512]bf16", lambda: DF._bias_act_raw(xc, b, None, 3, 0, 0.2, 1.41),
              2 * x.numel() * 2)
    yc = DF._bias_act_raw(xc, b, None, 3, 0, 0.2, 1.41)
    gxc = torch.empty_like(xc)
    mem_entry("bias_act_bwd_nhwc[64,32,64,512]bf16",
              lambda: K.call("dusty_bias_act_bwd_cl", K.ptr(xc), K.ptr(yc), K.ptr(gxc), K.ptr(db), B * H * W, 32,
                             0.2, 1.41, K.BF16, K.stream_of(xc)), 3 * x.numel() * 2)
    mem_entry("resample_blur_nhwc[64,32,64,512]bf16", lambda: blur(xc), 2 * x.numel() * 2)
    mem_entry("blur+pad_fused_nhwc[64,32,64,512]bf16", lambda: DF.blur_pad_cl(xc, btaps),
              (x.numel() + B * 32 * 66 * 514) * 2)
    gpc = gp.contiguous(memory_format=CL)
    mem_entry("blur+pad_fused_adjoint_nhwc[64,32,66,514]bf16", lambda: DF._BlurPadCL.apply(gpc, btaps, True),
              (x.numel() + gp.numel()) * 2)
    mem_entry("pad_ring1_nhwc[64,32,64,512]bf16", lambda: pad(xc), (x.numel() + gp.numel()) * 2)
    mem_entry("pad_ring1_adjoint_nhwc[64,32,66,514]bf16",
              lambda: DF._pad_raw(gpc, (1, 1, 1, 1), (K.PAD_REPLICATE, K.PAD_CIRCULAR), True, (H, W)),
              (x.numel() + gp.numel()) * 2)
    gdc = torch.randn(B, 32, H // 2, W // 2, device=device, dtype=bf).contiguous(memory_format=CL)
    dxc = torch.empty_like(xc)
    mem_entry("residual_fork_bwd_nhwc[64,32,64,512]bf16",
              lambda: K.call("dusty_residual_fork_bwd_cl", K.ptr(gpc), K.ptr(gdc), K.ptr(dxc), btaps[0], btaps[1],
                             btaps[2], btaps[3], B, H, W, 32, K.BF16, K.stream_of(xc)),
              (gpc.numel() + gdc.numel() + dxc.numel()) * 2)
    del gdc, dxc
    # discriminator stem (BlurVH + 1x1 conv 2->32 + bias/lrelu) and residual tail, fused kernels
    xs = torch.tanh(torch.randn(B, 1, H, W, device=device))
    ws_ = torch.randn(32, 2, device=device) / 1.4
    bs_ = torch.zeros(32, device=device)
    ys_ = torch.empty((B, 32, H, W), dtype=bf, device=device, memory_format=CL)
    mem_entry("stem_fwd[64,1->32,64,512]bf16",
              lambda: K.call("dusty_stem_fwd", K.ptr(xs), K.ptr(ws_), K.ptr(bs_), K.ptr(ys_), B, H, W, 32, 0.25, 0.5,
                             0.25, 0.2, 1.41, K.F32, K.stream_of(xs)), xs.numel() * 4 + ys_.numel() * 2)
    dvh_ = torch.empty(B, 2, H, W, device=device)
    dwb_ = torch.empty(32, 3, device=device)
    mem_entry("stem_bwd[64,1->32,64,512]bf16",
              lambda: K.call("dusty_stem_bwd", K.ptr(xc), K.ptr(ys_), K.ptr(xs), K.ptr(ws_), K.ptr(dvh_), K.ptr(dwb_),
                             B, H, W, 32, 0.25, 0.5, 0.25, 0.2, 1.41, K.F32, K.stream_of(xs)),
              xs.numel() * 4 + 2 * ys_.numel() * 2 + dvh_.numel() * 4)
    sk_ = torch.randn(B, 64, 32, 256, device=device, dtype=bf).contiguous(memory_format=CL)
    pr_ = torch.randn(B, 64, 32, 256, device=device, dtype=bf).contiguous(memory_format=CL)
    b64 = torch.zeros(64, device=device)
    mem_entry("residual_tail_fwd[64,64,32,256]bf16", lambda: DF.residual_tail(pr_, b64, sk_), 3 * sk_.numel() * 2)
    del xs, ys_, dvh_, sk_, pr_
    del xc, yc, gxc, gpc
    hd = torch.randn(B, 32, H, W, device=device, dtype=bf)
    wh = (torch.randn(B, 2, 32, device=device) / 6).to(bf)
    bh = torch.zeros(2, device=device)
    mem_entry("heads_fwd[O=2,C=32,64x512]bf16", lambda: DF.modconv_bmm(wh, hd, None, bh, 1, 0.0, 1.0),
              (hd.numel() + B * 2 * H * W) * 2)
    gh_ = torch.randn(B, 2, H, W, device=device, dtype=bf)
    gwh_ = torch.empty(B, 2, 32, device=device)
    mem_entry("heads_dw[O=2,C=32,64x512]bf16",
              lambda: K.call("dusty_modconv_bwd_dw", K.ptr(gh_), K.ptr(hd), None, K.ptr(gwh_), B, 2, 32, 0, 1,
                             H * W, K.BF16, 0, 0, K.stream_of(hd)), (hd.numel() + gh_.numel()) * 2)
    del gh_
    del hd
    ang = torch.rand(B, 2, H, W, device=device)
    fr = torch.randn(256, 2, device=device)
    ph = torch.rand(256, device=device)
    mem_entry("fourier[64,512,64,512]bf16", lambda: DF.fourier_features(ang, fr, ph, bf),
              B * H * W * (8 + 512 * 2))
    mem_entry("sumsq[64,64,64,512]bf16", lambda: DF.sumsq_total(gup), gup.numel() * 2)
    lg = torch.randn(B, 1, H, W, device=device)
    im = torch.tanh(torch.randn(B, 1, H, W, device=device))
    u = torch.rand(B, 1, H, W, device=device)
    mem_entry("gumbel_raydrop_fwd[64,1,64,512]f32", lambda: DF.raydrop_count(lg, im, u, -1.0, 1.0),
              lg.numel() * 4 * 6)
    trig = torch.rand(4, H * W, device=device)
    inv = torch.rand(B, 1, H, W, device=device)
    mem_entry("point_project[64,1,64,512]f32", lambda: DF.point_project(inv, trig, 1.45, 80.0),
              inv.numel() * 4 * 4)
    k12 = torch.randn(1, 12, device=device)
    from dusty_gan_v2_b200.gans.models.ops.upfirdn2d.upfirdn2d import upfirdn2d
    xa = torch.randn(B, 1, 76, 524, device=device)
    mem_entry("upfirdn2d_ada_up_x[64,1,76,524]f32", lambda: upfirdn2d(xa, k12, up=(2, 1), pad=(6, 5, 0, 0)),
              xa.numel() * 4 * 3)
    del x, y, gx, h, gup, gp
    # ---- contractions (tcgen05): level-4 conv1 (HBM/L2 bound: N = 32) and level-1 conv1
    def contraction(tag, O_, C1, C2, hh, ww, shared):
        P = hh * ww
        wb = (torch.randn(B, O_, C1 + C2, device=device) / (C1 + C2) ** 0.5).to(bf)
        x1 = torch.randn(B, C1, hh, ww, device=device, dtype=bf)
        a = torch.rand(1 if shared else B, 2, hh, ww, device=device)
        x2 = DF.fourier_features(a, fr, ph, bf)
        bias = torch.zeros(O_, device=device)
        flops = 2.0 * B * O_ * (C1 + C2) * P
        byts = (x1.numel() + B * O_ * P + wb.numel() + x2.numel()) * 2
        e = tc_entry(f"modconv_fwd[{tag} B=64,O={O_},K={C1}+{C2},P={P},pe={'shared' if shared else 'per-sample'}]bf16",
                     lambda: DF.modconv_bmm(wb, x1, x2, bias, 3, 0.2, 1.41), flops, byts)
        g = torch.randn(B, O_, hh, ww, device=device, dtype=bf)
        gw = torch.empty(B, O_, C1 + C2, device=device)
        pe_tag = "shared" if shared else "per-sample"
        tc_entry(f"modconv_dw[{tag},pe={pe_tag}]bf16",
                 lambda: K.call("dusty_modconv_bwd_dw", K.ptr(g), K.ptr(x1), K.ptr(x2), K.ptr(gw), B, O_, C1, C2,
                                x2.shape[0], P, K.BF16, 0, 0, K.stream_of(g)),
                 flops, (x1.numel() + g.numel() + x2.numel() * (1 if shared else 1)) * 2 + gw.numel() * 4)
        gx1 = torch.empty_like(x1)
        tc_entry(f"modconv_dx[{tag},pe={pe_tag}]bf16",
                 lambda: K.call("dusty_modconv_bwd_dx", K.ptr(wb), K.ptr(g), K.ptr(gx1), B, O_, C1, C1 + C2, P,
                                K.BF16, K.BF16, 0, None, None, K.stream_of(g)),
                 2.0 * B * O_ * C1 * P, (g.numel() + gx1.numel() + wb.numel()) * 2)
        return e

    contraction("L4.conv1", 32, 64, 512, H, W, True)
    contraction("L4.conv1", 32, 64, 512, H, W, False)
    contraction("L1.conv1", 256, 512, 512, 8, 64, True)
    contraction("L2.conv1", 128, 256, 512, 16, 128, True)
    torch.cuda.empty_cache()

    # ---- discriminator convolutions (tcgen05 implicit GEMM): every layer x (fprop, dgrad, wgrad).
    # Per plain training iteration the trunk runs 3 forward, 3 data-gradient and 2 filter-gradient
    # passes in units of a 64-sample batch (G step: D frozen on 64 fakes; D step: 128 stacked
    # real + fake with parameter gradients) -- the weights of the time shares below.
    CLF = torch.channels_last
    per_iter = {"fprop": 3, "dgrad": 3, "wgrad": 2}
    layers = []
    for i in range(4):
        C_ = 32 << i
        hh, ww = H >> i, W >> i
        layers.append((f"RB{i}.conv1 3x3 s1 {C_}->{C_}", C_, C_, hh + 2, ww + 2, 3, 1))
        layers.append((f"RB{i}.conv2 3x3 s2 {C_}->{2 * C_}", C_, 2 * C_, hh + 2, ww + 2, 3, 2))
        layers.append((f"RB{i}.skip 1x1 s1 {C_}->{2 * C_} (decimated)", C_, 2 * C_, hh // 2, ww // 2, 1, 1))
    layers.append(("EP.conv 3x3 s1 512->512", 512, 512, 6, 34, 3, 1))
    conv_entries, fam_ms, fam_flops = [], 0.0, 0.0
    for name, C_, O_, hh, ww, k_, s_ in layers:
        xx = torch.randn(B, C_, hh, ww, device=device).to(bf).contiguous(memory_format=CLF)
        wt = (torch.randn(O_, C_, k_, k_, device=device) / (C_ * k_ * k_) ** 0.5).to(bf)
        st_ = (s_, s_)
        ho, wo = (hh - k_) // s_ + 1, (ww - k_) // s_ + 1
        yy = torch.empty((B, O_, ho, wo), dtype=bf, device=device, memory_format=CLF)
        gy = torch.randn(B, O_, ho, wo, device=device).to(bf).contiguous(memory_format=CLF)
        wtco = DF.filter_tco(wt)
        flops = 2.0 * yy.numel() * C_ * k_ * k_
        nb = {"fprop": (xx.numel() + yy.numel() + wt.numel()) * 2,
              "dgrad": (xx.numel() + yy.numel() + wt.numel()) * 2,
              "wgrad": (xx.numel() + yy.numel()) * 2 + wt.numel() * 4}
        fns = {"fprop": lambda: DF.conv2d_fprop_tc(xx, wt, st_),
               "dgrad": lambda: DF.conv2d_dgrad_tc(gy, wt, st_, (hh, ww), wtco),
               "wgrad": lambda: DF.conv2d_wgrad_tc(gy, xx, st_, wt.shape, torch.float32)}
        for op in ("fprop", "dgrad", "wgrad"):
            e = tc_entry(f"conv_{op}[{name} @B=64]bf16", fns[op], flops, nb[op])
            e["launches_per_iteration"] = per_iter[op]
            e["step_ms_share"] = e["ms"] * per_iter[op]
            conv_entries.append(e)
            fam_ms += e["ms"] * per_iter[op]
            fam_flops += flops * per_iter[op]
        del xx, wt, yy, gy, wtco
    sustained = peaks.get("bf16_tflops_sustained", 1400.0)
    family = {"what": "all D convolutions of one plain training iteration (3 fprop + 3 dgrad + 2 wgrad passes "
                      "in 64-sample units), isolated launches, L2 flushed before each",
              "ms_per_iteration": fam_ms, "tflop_per_iteration": fam_flops / 1e12,
              "tflops": fam_flops / (fam_ms * 1e-3) / 1e12,
              "frac_of_bf16_sustained": fam_flops / (fam_ms * 1e-3) / 1e12 / sustained,
              "frac_of_bf16_burst": fam_flops / (fam_ms * 1e-3) / 1e12 / tflops}
    dom = max(conv_entries, key=lambda e: e["step_ms_share"])
    return dom, out, family


# ------------------------------------------------------------------------------ main
# ------------------------------------------------------- reference's own CUDA kernels (oracle/_ref)
def reference_kernel_baseline():
    """Kernel-level reference arm: the REFERENCE'S OWN two CUDA kernels (fused_bias_act_kernel.cu,
    upfirdn2d_kernel.cu; sm_100 builds made in place from /root/reference by oracle/build_ref.py,
    shipped as oracle/_ref/*.so) timed beside ours on the same tensors at the training step's
    shapes.  Device time of CUDA-graph replays (the 20-40 us of host wrapper cost per call on
    either side would hide kernels of this size), L2 flushed before each replay.  Returns a dict."""
    import torch

    if not torch.cuda.is_available():
        return {"unavailable": "no CUDA device"}
    from oracle import build_ref      # the checker / reference arm; never imported by the package
    import dusty_gan_v2_b200.functional as DF
    from dusty_gan_v2_b200.gans.models.ops.upfirdn2d.upfirdn2d import upfirdn2d
    fused = build_ref.load_built("dusty_ref_fused")
    ufd = build_ref.load_built("dusty_ref_upfirdn2d")
    if fused is None or ufd is None:
        return {"unavailable": "oracle/_ref not built (python oracle/build_ref.py in the build container)"}
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, reps=6, inner=4):
        """Median device time per call in us.  The `inner` calls are captured into a CUDA graph
        and replayed, so that the host cost of either side's Python / pybind wrapper (20-40 us,
        more than some of these kernels take) stays out of the number."""
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        graph = None
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=side):
                    for _ in range(inner):
                        fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = gr
        except Exception as e:          # fall back to eager launches
            print(f"graph capture failed ({type(e).__name__}); eager timing", file=sys.stderr)
            torch.cuda.synchronize()
        best = []
        for _ in range(reps):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if graph is not None:
                graph.replay()
            else:
                for _ in range(inner):
                    fn()
            e1.record()
            torch.cuda.synchronize()
            best.append(e0.elapsed_time(e1) / inner)
        best.sort()
        return best[len(best) // 2] * 1e3          # us

    rows = []

    def row(name, ref_fn, our_fn, bytes_):
        r, o = timeit(ref_fn), timeit(our_fn)
        rows.append({"kernel": name, "reference_us": round(r, 1), "ours_us": round(o, 1),
                     "speedup": round(r / o, 2), "algorithmic_bytes": bytes_,
                     "reference_gbs": round(bytes_ / r / 1e3, 1), "ours_gbs": round(bytes_ / o / 1e3, 1)})

    B, H, W = 64, 64, 512
    for dt_ref, dt_our, tag in ((torch.float32, torch.float32, "f32"), (torch.float16, torch.bfloat16, "f16|bf16")):
        x = torch.randn(B, 32, H, W, device=dev)
        b = torch.randn(32, device=dev)
        xr, br, xo, bo = x.to(dt_ref), b.to(dt_ref), x.to(dt_our), b.to(dt_our)
        er, eo = xr.new_empty(0), xo.new_empty(0)
        es = x.element_size() if dt_ref == torch.float32 else 2
        row(f"bias_act_fwd[64,32,64,512]{tag}", lambda: fused.fused_bias_act(xr, br, er, 3, 0, 0.2, 1.41),
            lambda: DF.fused_bias_act(xo, bo, eo, 3, 0, 0.2, 1.41), 2 * x.numel() * es)
        yr = fused.fused_bias_act(xr, br, er, 3, 0, 0.2, 1.41)
        yo = DF.fused_bias_act(xo, bo, eo, 3, 0, 0.2, 1.41)
        # the reference's backward = the gradient kernel + a separate ATen reduction for db
        # (fused_act.py:28-40); ours produces dx and db in one pass
        row(f"bias_act_bwd+db[64,32,64,512]{tag}",
            lambda: fused.fused_bias_act(xr, er, yr, 3, 1, 0.2, 1.41).sum((0, 2, 3)),
            lambda: DF._BiasActBackward.apply(xo, yo, True, 0.2, 1.41), 3 * x.numel() * es)
        del x, xr, xo, yr, yo
    # AdaptiveAugment's four separable passes (adaptive_augment.py:497-545), 1-channel fp32
    k = torch.tensor([0.015404109327027373, 0.0034907120842174702, -0.11799011114819057,
                      -0.048311742585633, 0.4910559419267466, 0.787641141030194, 0.3379294217276218,
                      -0.07263752278646252, -0.021060292512300564, 0.04472490177066578,
                      0.0017677118642428036, -0.007800708325034148], device=dev)
    for name, hw, ks, up, down, pad in (("ada_up_x", (76, 524), (1, 12), (2, 1), (1, 1), (6, 5, 0, 0)),
                                        ("ada_up_y", (76, 1048), (12, 1), (1, 2), (1, 1), (0, 0, 6, 5)),
                                        ("ada_down_x", (140, 1036), (1, 12), (1, 1), (2, 1), (-1, -1, 0, 0)),
                                        ("ada_down_y", (140, 512), (12, 1), (1, 1), (1, 2), (0, 0, -1, -1))):
        x = torch.randn(B, 1, hw[0], hw[1], device=dev)
        kk = k.reshape(ks).contiguous()
        x4 = x.reshape(B, hw[0], hw[1], 1)
        a = (up[0], up[1], down[0], down[1], pad[0], pad[1], pad[2], pad[3])
        out = ufd.upfirdn2d(x4, kk, *a)
        row(f"upfirdn2d_{name}[64,1,{hw[0]},{hw[1]}]f32", lambda: ufd.upfirdn2d(x4, kk, *a),
            lambda: upfirdn2d(x, kk, up=up, down=down, pad=pad), (x.numel() + out.numel()) * 4)
    res = {"device": torch.cuda.get_device_name(0), "note": "reference kernels = sm_100 builds of the "
           "reference's own .cu files (oracle/_ref); median of 6 graph replays of 4 launches (device time, no host "
           "wrapper cost), L2 flushed before each replay", "rows": rows}
    return res


def generator_forward_cpu_baseline(sdG, z, angle, batch, reps=3):
    """CPU leg of BASELINE config 1 (tools/config1_bench.py --cpu): the oracle's generator forward
    (eval, psi = 1) on the host cores, best of `reps` after one warm-up."""
    import torch

    from oracle import dusty_oracle as O
    ang = angle.expand(batch, -1, -1, -1) if angle.shape[0] != batch else angle
    u = torch.rand(batch, 1, ang.shape[-2], ang.shape[-1], generator=torch.Generator().manual_seed(3))
    times = []
    with torch.no_grad(), O.fir_impl("library"):
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            O.generator(sdG, z, ang, u)
            times.append(time.perf_counter() - t0)
    best = min(times[1:])
    return {"images_per_s": batch / best, "ms_per_forward": best * 1e3, "batch": batch,
            "cores": torch.get_num_threads(), "kind": "port"}


def inversion_cpu_baseline(sdG, z0, angle, depth, mask, latent_type, min_depth=1.45, max_depth=80.0):
    """CPU leg of BASELINE config 5 (tools/inversion_bench.py --cpu): the oracle's restatement of
    one stage-1 inversion step (objective, backward to the latent, Adam) on the host cores for a
    bounded sample (the batch handed in); best of two steps."""
    import torch

    from oracle import dusty_oracle as O
    t_depth, t_inv = O.inversion_targets(depth, mask, min_depth, max_depth)
    z = torch.nn.Parameter(z0.clone())
    opt = torch.optim.Adam([z], lr=5e-2)
    times = []
    for _ in range(2):
        t0 = time.perf_counter()
        with O.fir_impl("library"):
            _, loss = O.inversion_forward(sdG, z, angle, t_depth, t_inv, mask, min_depth, max_depth,
                                          latent_type)
            opt.zero_grad(set_to_none=True)
            loss.backward(gradient=torch.ones_like(loss))
        opt.step()
        times.append(time.perf_counter() - t0)
    n = int(z0.shape[0])
    return {"target_iterations_per_s": n / min(times), "batch": n, "cores": torch.get_num_threads(),
            "kind": "port"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.ref_kernels_only:
        if rank == 0:
            print(json.dumps(reference_kernel_baseline()))
        return

    if args.impl == "reference":
        if rank != 0:
            return
        ips, spt, cores, desc = cpu_reference_run(args, args.steps, args.warmup, args.cpu_batch)
        line = {"impl": "reference", "metric": METRIC.replace("dusty_v2", args.arch), "value": ips, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": spt * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"{args.arch} G+D training step (nsgan + lazy R1), 64x512, CPU sample "
                                       f"batch {args.cpu_batch}", "global_batch": args.cpu_batch,
                           "parallelism": "cpu"},
                "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
                "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    import dusty_gan_v2_b200 as pkg
    from dusty_gan_v2_b200.gans.trainer import Trainer
    from dusty_gan_v2_b200.presets import preset

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    pkg.set_precision(args.precision)
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark   # the reference sets it (gans/utils.py:29-30)
    if args.precision == "bf16":                 # fp32 epilogue GEMMs of D on the tensor cores
        torch.backends.cuda.matmul.allow_tf32 = True
    torch.manual_seed(0 + rank)
    import numpy as np
    np.random.seed(0 + rank)

    B = args.batch_per_gpu
    if args.global_batch is not None:
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} does not divide over {world} GPUs")
        B = args.global_batch // world
    cfg = preset(args.arch, batch_size=B * world)
    if args.ada_p is not None:
        cfg.training.augment.p_init = args.ada_p
        cfg.training.augment.p_target = None
    pool_dev = synthetic_batches(4, B, seed=2 + rank, device=device)
    tr = Trainer(cfg, cycle(pool_dev), device=device, rank=rank, world_size=world,
                 angle_file=os.path.join(ROOT, "data/coords/kitti_raw.npy"),
                 cuda_graphs=not args.no_cuda_graphs)
    tr.A.generator = torch.Generator().manual_seed(100 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_events = []

    def run(first_it, n, trace=False):
        last = None
        for i in range(n):
            last = tr.step(first_it + i)
            if trace:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                step_events.append(ev)
        return last

    # warm-up: iteration 0 carries an R1 step, so every phase is warmed
    run(0, args.warmup)
    gp = max(tr.gp_every, 1)
    start = ((args.warmup + gp - 1) // gp) * gp          # timed region starts on an R1 iteration
    # the iterations between the requested warm-up and that R1 iteration run untimed as well (they
    # would otherwise be skipped): lazily created handles, allocator growth and NCCL channel set-up
    # of a young process otherwise land in the timed window (seen at N = 2: 15.4-17.6 ms / step for
    # the same build, while the later end-to-end window was stable at 15.4)
    run(args.warmup, start - args.warmup)
    # (the sampler is started BEFORE the barrier: NVML initialisation on rank 0 alone, between the
    # barrier and the first timed step, made the other ranks wait for it inside their first
    # collective -- up to 80 ms inside their timed window, and the window is the max over ranks)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    n0 = pkg.launch_count()
    g0 = tr.graph_replayed_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.ncu_window:
        torch.cuda.profiler.start()
    e0.record()
    run(start, args.steps, trace=True)
    e1.record()
    barrier()
    if args.ncu_window:
        torch.cuda.profiler.stop()
    launches = pkg.launch_count() - n0 + (tr.graph_replayed_launches - g0)
    prev, per = e0, []
    for ev in step_events:
        per.append(prev.elapsed_time(ev))
        prev = ev
    if args.per_step and rank == 0:
        print("device ms per timed step:", [round(v, 2) for v in per], file=sys.stderr)
    # the training loop's long-run mix (one R1 iteration in `gp_every`) from this rank's per-step
    # device times: the timed window starts ON an R1 iteration, so a window that is not a multiple
    # of gp_every steps carries more than its share of them (e.g. 2 in 20) -- `value` is the
    # window's own throughput, this is what the same step times give at the 15:1 mix
    is_r1 = [bool(tr.gp_every) and (start + i) % gp == 0 for i in range(args.steps)]
    plain = sorted(v for v, r in zip(per, is_r1) if not r)
    r1t = sorted(v for v, r in zip(per, is_r1) if r)
    mix = None
    if plain:
        p_med = plain[len(plain) // 2]
        r_med = r1t[len(r1t) // 2] if r1t else p_med
        mix_ms = ((gp - 1) * p_med + r_med) / gp if tr.gp_every else p_med
        mix = {"plain_ms": round(p_med, 3), "r1_ms": round(r_med, 3), "ms_per_step": round(mix_ms, 3),
               "images_per_s_per_gpu": round(B / (mix_ms * 1e-3), 1),
               "what": f"rank-0 median step times combined at the loop's {gp - 1}:1 plain : R1 mix"}
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clk = clocks.stop() if clocks else None
    value = args.steps * B * world / (ms * 1e-3)
    r1_steps = len([i for i in range(start, start + args.steps) if tr.gp_every and i % tr.gp_every == 0])

    # ---- end to end: pinned host batches in, step scalars out, every step
    e2e = None
    if not args.no_e2e:
        pool_host = synthetic_batches(4, B, seed=2 + rank, pinned=True)
        tr.batch_iter = cycle(pool_host)
        h2d = sum(t.numel() * t.element_size() for t in pool_host[0].values())
        # same R1 mix as the device-timed window: start on an R1 iteration too
        first = start + args.steps
        e2e_start = ((first + 2 + gp - 1) // gp) * gp
        run(first, e2e_start - first)
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        # every step's scalars are copied to pinned host memory and READ on the host inside the
        # timed region; the read of step i happens after step i+1 has been enqueued (two pinned
        # slots + events), so the host never drains the device queue between steps
        slots = [None, None]
        seen = []
        for i in range(args.steps):
            packed = tr.step(e2e_start + i)
            if slots[i & 1] is None:
                slots[i & 1] = (torch.empty(32, dtype=packed.dtype, pin_memory=True), torch.cuda.Event())
            buf, ev = slots[i & 1]
            buf[:packed.numel()].copy_(packed, non_blocking=True)    # (R1 iterations carry more scalars)
            ev.record()
            d2h = packed.numel() * packed.element_size()
            if i > 0:
                pbuf, pev = slots[(i - 1) & 1]
                pev.synchronize()
                seen.append(float(pbuf[0]))
        slots[(args.steps - 1) & 1][1].synchronize()
        seen.append(float(slots[(args.steps - 1) & 1][0][0]))
        assert len(seen) == args.steps
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=device)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": args.steps * B * world / float(dt.item()), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "d2h": "async copy of the step's scalars to pinned memory, read one step later",
               "r1_steps_timed": len([i for i in range(e2e_start, e2e_start + args.steps)
                                      if tr.gp_every and i % tr.gp_every == 0])}
        tr.batch_iter = cycle(pool_dev)

    line = {"metric": METRIC.replace("dusty_v2", args.arch), "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if args.global_batch is None else "strong", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.arch} full G+D training step (nsgan + R1 every 16th step, ADA, "
                                   f"warm-up dropout, EMA, Adam), 64x512 range images, random-init weights",
                       "global_batch": B * world, "per_gpu_batch": B, "parallelism": f"dp{world}",
                       "r1_steps_timed": r1_steps, "steady_state_mix": mix, "ada_p": "adaptive from 0.0" if args.ada_p is None else args.ada_p,
                       "l2": "no flush: per-step working set (GBs) >> 126 MB L2",
                       "warmup_extra": f"{start - args.warmup} further untimed iterations up to the R1 iteration the "
                                       "timed window starts on",
                       "dense_convs": "every dense (transposed) convolution on own kernels: tcgen05 implicit GEMM "
                                      "(fprop / dgrad / wgrad, CTA pairs for the deep layers) for bf16 NHWC, "
                                      "split-bf16 operands on the same kernels in fp32 mode, CUDA-core family "
                                      "for odd shapes; no library convolution; D's linears on own GEMMs "
                                      "(cuBLAS only for the mapping / style linears)",
                       "augment": "AdaptiveAugment as a device-side op (transforms drawn on the device)",
                       "side_streams": "generator weight bank, discriminator filter bank, next batch's upload"},
            "clocks": clk, "gpu_launches": launches, "e2e": e2e}

    if rank == 0 and world == 1:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        if not args.no_roofline:
            del tr
            torch.cuda.empty_cache()
            dom, kernels, family = kernel_rooflines(device, peaks)
            line["roofline"] = {k: dom[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")}
            for k in ("kernel", "peak_source", "algorithmic_bytes", "algorithmic_flops", "tflops",
                      "tensor_frac", "hbm_gbs", "hbm_frac", "ms", "launches_per_iteration", "step_ms_share"):
                line["roofline"][k] = dom[k]
            line["roofline"]["selected_by"] = ("largest share of the step's device time among all kernels timed "
                                               "below (isolated duration x launches per plain iteration)")
            line["roofline"]["conv_family"] = family
            line["kernels"] = kernels
        if not args.no_cpu_baseline:
            ips, spt, cores, desc = cpu_reference_run(args, 2, 1, args.cpu_batch)
            line["cpu_baseline"] = {"value": ips, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": desc}
        if not args.no_ref_kernels and not args.no_roofline:
            # kernel-level reference arm in a child process with a hard time limit: nothing it
            # does (a failed graph capture, a missing oracle/_ref) can cost the line above
            try:
                cp = subprocess.run([sys.executable, os.path.abspath(__file__), "--ref-kernels-only"],
                                    capture_output=True, text=True, timeout=120, cwd=ROOT)
                line["ref_kernels"] = json.loads(cp.stdout.strip().splitlines()[-1])
            except Exception as e:
                line["ref_kernels"] = {"unavailable": f"{type(e).__name__}: {str(e)[:120]}"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
