#!/usr/bin/env python
"""bench.py -- FocusRAFT inference throughput on the B200 correlation path.

Contract (one JSON line on stdout from rank 0):
  python bench.py --gpus N --steps K --warmup W            # this repo's arm
  python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm (rank 0 only)

A "step" = one FocusRAFT forward (12 refinement iterations, test mode) over one batch of 8
synthetic KITTI-shaped pairs (376x1248) per GPU: BASELINE.json configs[1].  Work shards by
image pair: every rank runs its own batch, there is no data-path collective (weak scaling).

  value     pairs/s, inputs already resident in HBM (max time over ranks, CUDA events)
  e2e       pairs/s through the public API with PINNED HOST inputs: H2D of images+mask and
            D2H of the full-resolution flow inside the timed region, every step; the copies run on a
            second stream so they overlap the neighbouring steps' compute
  roofline  the dominant kernel of the hot path (the per-iteration lookup, 12 launches/step):
            algorithmic bytes (2904 B/query) / mean launch duration measured with
            CUDA events around every launch of the timed steps, vs the measured HBM peak
  stock_gpu_hot_path  (N=1) the reference's own op sequence for the hot path through stock PyTorch CUDA kernels on
            this GPU, next to this repo's kernels: correlation-path milliseconds per step
  cpu_baseline  the reference's CPU PyTorch path (oracle/corr_torch_cpu.py restates corr.py with
            the same ATen ops; host model = this repo's PyTorch rewrite, validated against the
            reference in tests/test_host_model.py) on a bounded sample: batch 1, same shape
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout at
# communicator init, cuDNN / torch may warn), so the process-level stdout is pointed at stderr for the whole run and the
# JSON line is written to the saved original descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


METRIC = "FF-RAFT pairs/sec @376x1248, 12 iters"
UNIT = "pairs/s"
H, W, ITERS, BATCH = 376, 1248, 12, 8
LOOKUP_BYTES_PER_QUERY = 4 * 100 * 4 + 324 * 4 + 8  # 2904 B


# ------------------------------------------------------------------------------ helpers
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()  # the exact PID we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nme, v in zip(names, r[4:8]):
                    if v.strip().lower() == "active":
                        reasons.add(nme)
            except Exception:
                continue
        if sm:
            sm.sort()
            out.update({"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)})
        return out


def dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime

        # a short watchdog: a rank that falls out of step must fail the run in minutes, not hold 8 GPUs for ten
        dist.init_process_group(backend="nccl" if torch.cuda.is_available() else "gloo", init_method="env://",
                                timeout=datetime.timedelta(seconds=180))
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist

        dist.barrier()


def max_over_ranks(x: float, world: int, device) -> float:
    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def make_model(device, channels_last: bool):
    from focusflow_official_b200.host import FocusRAFT

    torch.manual_seed(1234)  # GLOBAL.SEED of every reference config
    model = FocusRAFT()
    from weights import fill_state_dict

    sd = model.state_dict()
    fill_state_dict(sd, seed=1234)  # random init with a contractive flow head (see tests/weights.py)
    model.load_state_dict(sd)
    model = model.to(device).eval()
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    return model


class StockTorchCorrBlock:
    """The reference's CorrBlock (corr.py:12-60 + utils.py:57-71) written with the stock PyTorch ops it calls, for
    the 'what if only the hot path were stock' measurement.  Measurement only; never used by the product."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4):
        import torch.nn.functional as F

        self.num_levels, self.radius = num_levels, radius
        b, d, h, w = fmap1.shape
        n = h * w
        corr = torch.matmul(fmap1.view(b, d, n).transpose(1, 2), fmap2.view(b, d, n)).view(b, h, w, 1, h, w)
        corr = (corr / torch.sqrt(torch.tensor(d).float())).reshape(b * n, 1, h, w)
        self.corr_pyramid = [corr]
        for _ in range(num_levels - 1):
            corr = F.avg_pool2d(corr, 2, stride=2)
            self.corr_pyramid.append(corr)

    def __call__(self, coords):
        import torch.nn.functional as F

        r = self.radius
        coords = coords.permute(0, 2, 3, 1)
        b, h, w, _ = coords.shape
        out = []
        for i, corr in enumerate(self.corr_pyramid):
            dx = torch.linspace(-r, r, 2 * r + 1, device=coords.device)
            delta = torch.stack(torch.meshgrid(dx, dx, indexing="ij"), axis=-1)
            cl = coords.reshape(b * h * w, 1, 1, 2) / 2 ** i + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
            hh, ww = corr.shape[-2:]
            xg, yg = cl.split([1, 1], dim=-1)
            grid = torch.cat([2 * xg / (ww - 1) - 1, 2 * yg / (hh - 1) - 1], dim=-1)
            out.append(F.grid_sample(corr, grid, align_corners=True).view(b, h, w, -1))
        return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def stock_gpu_corr_path(b, device):
    """SURVEY 8(d) 'stock' GPU baseline: the reference's own op sequence for the hot path (corr.py:13-60 +
    utils.py:57-71: torch.matmul + divide, 3x avg_pool2d, per level meshgrid + grid_sample + glue) on THIS GPU,
    TF32 matmul as the reference configures it, same fmap shape as the step.  Stock PyTorch CUDA kernels only --
    none of this repo's code.  Returns (build_ms, lookup_ms), medians of 5 after 2 warm-ups."""
    import torch.nn.functional as F

    h, w, d = H // 8, W // 8, 256
    n = h * w
    g = torch.Generator(device=device)
    g.manual_seed(1234)
    f1 = torch.randn(b, d, h, w, device=device, generator=g) * 4.4
    f2 = torch.randn(b, d, h, w, device=device, generator=g) * 4.4
    ys, xs = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
    coords = torch.stack([xs, ys], 0).float()[None].repeat(b, 1, 1, 1) + torch.randn(b, 2, h, w, device=device, generator=g) * 3
    state = {}

    def build():
        corr = torch.matmul(f1.view(b, d, n).transpose(1, 2), f2.view(b, d, n)).view(b, h, w, 1, h, w)
        corr = (corr / torch.sqrt(torch.tensor(d).float())).reshape(b * n, 1, h, w)
        pyr = [corr]
        for _ in range(3):
            corr = F.avg_pool2d(corr, 2, stride=2)
            pyr.append(corr)
        state["pyr"] = pyr

    def lookup():
        r = 4
        c = coords.permute(0, 2, 3, 1)
        out = []
        for i, corr in enumerate(state["pyr"]):
            dx = torch.linspace(-r, r, 2 * r + 1, device=device)
            delta = torch.stack(torch.meshgrid(dx, dx, indexing="ij"), axis=-1)
            cl = c.reshape(b * n, 1, 1, 2) / 2 ** i + delta.view(1, 2 * r + 1, 2 * r + 1, 2)
            hh, ww = corr.shape[-2:]
            xg, yg = cl.split([1, 1], dim=-1)
            grid = torch.cat([2 * xg / (ww - 1) - 1, 2 * yg / (hh - 1) - 1], dim=-1)
            out.append(F.grid_sample(corr, grid, align_corners=True).view(b, h, w, -1))
        return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()

    res = []
    with torch.no_grad():
        for fn in (build, lookup):
            ts = []
            for i in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                if i >= 2:
                    ts.append(e0.elapsed_time(e1))
            ts.sort()
            res.append(ts[len(ts) // 2])
    state.clear()
    torch.cuda.empty_cache()
    return res[0], res[1]


# ------------------------------------------------------------------------------ CPU arm
def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def load_reference_model(device):
    """The UNMODIFIED reference FF_RAFT_FUSION (ff_raft.py:75-160) from /root/reference or its verbatim copy under
    baseline/_ref (oracle/install_reference.py), with the same deterministic weights the B200 arm uses.
    Returns (model, kind): kind "reference", or ("port", this repo's host model + oracle/corr_torch_cpu.py) when the
    reference sources are not installed."""
    from oracle import reference_loader as RL
    from weights import fill_state_dict

    if RL.reference_root() is not None:
        model, _, _ = RL.load_ff_raft()
        sd = model.state_dict()
        fill_state_dict(sd, seed=1234)
        model.load_state_dict(sd, strict=True)
        return model.to(device).eval(), "reference"
    from oracle.corr_torch_cpu import TorchCorrBlock

    model = make_model(device, False)
    model.flow_net.corr_block = TorchCorrBlock
    return model, "port"


def cpu_reference_run(steps: int, warmup: int, batch: int = 1, config: int = 2):
    """The reference's own CPU path on all host threads, on a bounded sample of the workload (`batch` pair(s) per step):
    config 2 / 4 = FF_RAFT_FUSION.forward (ff_raft.py:134, test mode) at that config's shape and iteration count;
    config 5 = one training step (train.py:296-328: forward, the reference's MixLoss, backward, clip, AdamW)."""
    from weights import synthetic_pair, train_inputs

    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it can get
    torch.set_num_threads(host_threads())
    model, kind = load_reference_model("cpu")
    if config == 5:
        from focusflow_official_b200.host import TrainStep

        model.train()
        if kind == "reference":      # the reference's own loss module
            from oracle import reference_loader as RL

            sys.path.insert(0, RL.reference_root())
            from losses import build_losses as ref_losses
        step = TrainStep(model, iters=12)
        if kind == "reference":
            step.loss_function = ref_losses("MixLoss", gamma=0.8, max_flow=400, kernel_size=1, sigma=0.01, lamda=1)
        data = train_inputs(batch, 368, 496, 777)
        fn = lambda: step(*data)
        ctx = torch.enable_grad()
    else:
        wl = WORKLOADS[config]
        hh = wl["H"] + sum(wl["pad"])
        im1, im2, m1, m2 = synthetic_pair(batch, hh, wl["W"], seed=1234)
        fn = lambda: model(im1, im2, m1, m2, raft_iters=wl["iters"], test_mode=True)
        ctx = torch.no_grad()
    with ctx:
        for _ in range(warmup):
            fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, torch.get_num_threads(), kind


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    if args.config == 3:
        emit({"impl": "reference", "unavailable": "the reference's FF-PWC has no CPU path: correlation.py:320-321 raises "
                                                  "NotImplementedError and backwarp calls .cuda() (ff_pwcnet.py:33)"})
        return
    steps = max(5, min(args.steps, 8)) if args.config == 2 else max(2, min(args.steps, 3))
    warmup = max(1, min(args.warmup, 2)) if args.config == 2 else 1
    val, sec, cores, kind = cpu_reference_run(steps, warmup, config=args.config)
    what = {2: "FocusRAFT inference, KITTI shape 376x1248, 12 iters, CPU", 4: "FocusRAFT inference, Sintel shape 440x1024, 32 iters, CPU",
            5: "FocusRAFT training step (MixLoss, AdamW), 368x496, 12 iters, CPU"}[args.config]
    metric = {2: METRIC, 4: WORKLOADS[4]["metric"], 5: "FF-RAFT training pairs/sec @368x496, 12 iters, MixLoss"}[args.config]
    line = {
        "impl": "reference", "metric": metric, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": what, "batch": 1,
                   "note": "bounded sample: 1 pair per step on the host cores",
                   "code": "unmodified reference FF_RAFT_FUSION (baseline/_ref or /root/reference)" if kind == "reference"
                   else "this repo's PyTorch host model + oracle/corr_torch_cpu.py (reference sources not installed)"},
        "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{steps} step(s) of 1 pair after {warmup} warm-up: {what}"},
        "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def reference_on_this_gpu(b, device, steps=3, warmup=2):
    """SURVEY 8(d) 'stock' baseline: the UNMODIFIED reference model (its own CorrBlock: cuBLAS TF32 matmul, avg_pool2d,
    grid_sample + glue; its own encoder / update block code) on THIS GPU, same batch, same inputs, same weights,
    TF32 and cudnn.benchmark as the reference configures them (common.py:20-27).  None if the sources are absent."""
    from oracle import reference_loader as RL
    from weights import synthetic_pair

    if RL.reference_root() is None:
        return None
    model, _ = load_reference_model(device)
    im1, im2, m1, m2 = (t.to(device) for t in synthetic_pair(b, H, W, seed=1234))
    with torch.no_grad():
        for _ in range(warmup):
            model(im1, im2, m1, m2, raft_iters=ITERS, test_mode=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            model(im1, im2, m1, m2, raft_iters=ITERS, test_mode=True)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model
    torch.cuda.empty_cache()
    return {"what": "unmodified reference FF_RAFT_FUSION on this GPU (TF32, cudnn.benchmark), same batch / weights / inputs",
            "ms_per_step": round(ms, 3), "pairs_per_s": round(b / (ms * 1e-3), 3), "steps": steps}


# ------------------------------------------------------------------------------ GPU arm
def build_roofline(build_ms, b, tiled, peaks, h=None, w=None, storage_bytes=4):
    """CorrBlock.__init__ (operand pre-pass + GEMM [+ pyramid]) against BOTH rooflines: HBM (SURVEY 8(d) bytes) and the
    tensor pipe (2*B*N^2*D flop against the measured bf16 burst / sustained cuBLAS throughput)."""
    if not build_ms:
        return None
    h, w, d = h or H // 8, w or W // 8, 256
    n = h * w
    lv = [(h >> i) * (w >> i) for i in range(4)]
    vol = b * (2 * n * d * 4 + n * n * storage_bytes)                  # operands in, level 0 out
    pyr_w = storage_bytes * b * n * sum(lv[1:])                        # levels 1..3 out
    algo = vol + pyr_w + (0 if tiled else 4 * b * n * lv[0])           # the standalone pyramid re-reads level 0
    gbs = algo / (build_ms * 1e-3) / 1e9
    flops = 2.0 * b * n * n * d
    tf = flops / (build_ms * 1e-3) / 1e12
    out = {"kernels": "operand_amax + operand_prepass + volume_gemm (pyramid fused in the epilogue)" if tiled
           else "operand_prepass + volume_gemm + pyramid", "algorithmic_bytes": int(algo), "ms": round(build_ms, 4),
           "achieved": round(gbs, 1), "unit": "GB/s", "frac": round(gbs / peaks["hbm_gbs"], 4),
           "flops": flops, "tflops": round(tf, 1), "tensor_frac": round(tf / peaks["bf16_tflops"], 4),
           "bound": "hbm write (arithmetic intensity %.0f flop/B < ridge %.0f)" % (flops / algo, peaks["bf16_tflops"] * 1e3 / peaks["hbm_gbs"])}
    if peaks.get("bf16_tflops_sustained"):
        out["tensor_frac_sustained"] = round(tf / peaks["bf16_tflops_sustained"], 4)
    return out


class LaunchMeter:
    """Counts this repo's kernel launches and times every lookup launch with CUDA events."""

    def __init__(self):
        self.launches = 0
        self.lookup_events = []
        self.build_events = []
        self.enabled = False
        self.tiled = False
        self.nhwc = False
        self.fused_conv = False
        self.storage_bytes = 4

    def install(self):
        from focusflow_official_b200 import corr as C

        meter = self
        raw_lookup, raw_build = C._lookup_raw, C._volume_pyramid_raw
        raw_lookup_t, raw_build_t = C._lookup_tiled_raw, C._volume_pyramid_tiled_raw

        def lookup(*a, **k):
            if not meter.enabled:
                return raw_lookup(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = raw_lookup(*a, **k)
            e1.record()
            meter.lookup_events.append((e0, e1))
            meter.launches += 1
            return out

        def build(f1, f2, nl, prec, *a, **k):
            if not meter.enabled:
                return raw_build(f1, f2, nl, prec, *a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = raw_build(f1, f2, nl, prec, *a, **k)
            e1.record()
            meter.build_events.append((e0, e1))
            meter.launches += {0: 4, 1: 2}.get(prec, 3)  # [amax +] operand pre-pass + GEMM + pyramid (fp32: SGEMM + pyramid)
            return out

        def lookup_t(*a, **k):
            if not meter.enabled:
                return raw_lookup_t(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = raw_lookup_t(*a, **k)
            e1.record()
            meter.lookup_events.append((e0, e1))
            meter.launches += 1
            meter.tiled = True
            meter.nhwc = bool(out.dim() == 4 and out.stride(1) == 1 and out.shape[1] > 1)
            return out

        def build_t(f1, f2, nl, prec, *a, **k):
            if not meter.enabled:
                return raw_build_t(f1, f2, nl, prec, *a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = raw_build_t(f1, f2, nl, prec, *a, **k)
            e1.record()
            meter.build_events.append((e0, e1))
            # [max|fmap| for the fp16 block scaling +] operand pre-pass + GEMM with the pyramid fused into its epilogue
            meter.launches += 3 if prec == 0 else 2
            meter.storage_bytes = out[0].element_size()
            return out

        raw_lookup_conv = C.CorrBlock.lookup_conv

        def lookup_conv(blk, coords, conv):
            if not meter.enabled:
                return raw_lookup_conv(blk, coords, conv)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = raw_lookup_conv(blk, coords, conv)
            e1.record()
            meter.lookup_events.append((e0, e1))
            meter.launches += 1
            meter.tiled = meter.nhwc = meter.fused_conv = True
            return out

        C._lookup_raw, C._volume_pyramid_raw = lookup, build
        C._lookup_tiled_raw, C._volume_pyramid_tiled_raw = lookup_t, build_t
        C.CorrBlock.lookup_conv = lookup_conv

    def mean_ms(self, events):
        if not events:
            return None
        return sum(a.elapsed_time(b) for a, b in events) / len(events)


WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on (weak scaling: 8 pairs per GPU)
    2: dict(name="FocusRAFT inference batch {b}/GPU, KITTI shape 376x1248, 12 iters, B200", H=376, W=1248, pad=(0, 0), iters=12,
            batch=8, total=None, scaling="weak", metric=METRIC),
    # BASELINE.json configs[3]: Sintel 436x1024 (InputPadder 'sintel' -> 440x1024, utils.py:9-16), 32 iterations
    # (evaluate.py:62), 64 pairs in total sharded by pair over the ranks (strong scaling), sub-batches of 8
    4: dict(name="FocusRAFT inference, Sintel shape 436x1024 (padded to 440), 32 iters, 64 pairs sharded over {w} GPU(s)", H=436,
            W=1024, pad=(2, 2), iters=32, batch=8, total=64, scaling="strong", metric="FF-RAFT pairs/sec @436x1024, 32 iters"),
}
PWC_LEVELS = [(32, 112, 256), (64, 56, 128), (96, 28, 64), (128, 14, 32), (196, 7, 16)]     # config 3, 436x1024 -> 448x1024


def pwc_level_roofline(device, peaks, batch=16, iters=10):
    """The five cost volumes of one FF-PWC forward at config 3 (B = 16): CUDA events around every launch, L2 flushed
    between launches.  Algorithmic bytes 4*B*H*W*(2C+81), flops 162*C*B*H*W per level."""
    import focusflow_official_b200 as ff

    flush = torch.empty(512 << 20, dtype=torch.uint8, device=device)
    levels, tot_ms, tot_b = [], 0.0, 0.0
    for c, hh, ww in PWC_LEVELS:
        one = torch.randn(batch, c, hh, ww, device=device)
        two = torch.randn(batch, c, hh, ww, device=device)
        for _ in range(3):
            ff.correlation_leaky(one, two, 0.1)
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ff.correlation_leaky(one, two, 0.1)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        byts = 4.0 * batch * hh * ww * (2 * c + 81)
        tot_ms += ms
        tot_b += byts
        levels.append({"C": c, "H": hh, "W": ww, "ms": round(ms, 4), "GBps": round(byts / ms / 1e6, 1),
                       "frac": round(byts / ms / 1e6 / peaks["hbm_gbs"], 4), "TFLOPs_fp32": round(162.0 * c * batch * hh * ww / ms / 1e9, 2)})
    del flush
    return {"kernel": "pwc81_tma_kernel (ffcorr_pwc81_f32, leaky_relu fused)", "workload": "config 3: five levels of one FF-PWC forward, B=16, 436x1024",
            "bound": "hbm (level 2) / fp32 FMA (levels 3-6)", "algorithmic_bytes": int(tot_b), "ms": round(tot_ms, 4),
            "achieved": round(tot_b / tot_ms / 1e6, 1), "unit": "GB/s", "frac": round(tot_b / tot_ms / 1e6 / peaks["hbm_gbs"], 4),
            "levels": levels}


def run_gpu_arm(args, rank, world, local):
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device for the B200 arm (no CPU fallback); "
                           "use --impl reference for the CPU baseline")
    from focusflow_official_b200 import _lib
    from focusflow_official_b200.sharding import shard_range
    from weights import synthetic_pair

    wl = WORKLOADS[args.config]
    H, W, ITERS = wl["H"], wl["W"], wl["iters"]
    PH = H + sum(wl["pad"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    _lib.lib()  # fail loudly if the extension is missing
    # reference run settings: ALLOW_TF32 true, cudnn benchmark (common.py:20-27)
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True

    peaks, peak_src = load_peaks()
    model = make_model(device, args.channels_last)
    if args.update_cl:
        model.flow_net.update_block.to(memory_format=torch.channels_last)
        model.flow_net.update_channels_last = True
    if args.cnet_cl:
        # host plumbing: the context encoder (eval-mode BatchNorm has an NHWC kernel) in channels_last -> no NCHW<->NHWC
        # transposes around its convolutions (+3.2 % pairs/s).  The feature encoder stays NCHW: F.instance_norm has no
        # channels_last path, and a var_mean formulation of it gains nothing more (NOTES).
        model.flow_net.cnet.to(memory_format=torch.channels_last)
    if args.storage:
        model.flow_net.corr_storage = args.storage
    if args.fuse_convc1:
        model.flow_net.fuse_convc1 = True
    meter = LaunchMeter()
    meter.install()

    b = args.batch
    if wl["total"] is None:
        my_pairs = b                                           # weak scaling: a fixed batch per GPU
    else:
        lo, hi = shard_range(wl["total"], rank, world)         # strong scaling: this rank's slice of the 64 pairs
        my_pairs = hi - lo
    sub_batches = [min(b, my_pairs - s) for s in range(0, my_pairs, b)]
    nb = max(sub_batches)
    im1_h, im2_h, m1_h, _ = synthetic_pair(nb, H, W, seed=1234 + rank)
    pin = lambda t: t.pin_memory()
    im1_h, im2_h, m1_h = pin(im1_h), pin(im2_h), pin(m1_h)
    im1, im2, m1 = (t.to(device, non_blocking=True) for t in (im1_h, im2_h, m1_h))
    flow_host = torch.empty((nb, 2, H, W), dtype=torch.float32).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in (im1_h, im2_h, m1_h)) * len(sub_batches)
    d2h = flow_host.numel() * flow_host.element_size() * len(sub_batches)
    pad = wl["pad"]

    def forward(a, c, m):
        if pad != (0, 0):                                      # InputPadder 'sintel': replicate rows (utils.py:9-16)
            a, c, m = (torch.nn.functional.pad(t, [0, 0, pad[0], pad[1]], mode="replicate") for t in (a, c, m))
        up = model(a, c, m, None, raft_iters=ITERS, test_mode=True)[1]
        return up[:, :, pad[0]:PH - pad[1]] if pad != (0, 0) else up

    def step_resident():
        out = None
        for nbk in sub_batches:
            out = forward(im1[:nbk], im2[:nbk], m1[:nbk])
        return out

    # e2e: every (sub-)batch uploads ITS inputs from pinned host memory and downloads ITS result, all inside the timed
    # region.  The copies run on a second stream so that the upload of batch i+1 and the download of batch i-1
    # overlap the compute of batch i (two device input slots, events in both directions).
    copy_stream = torch.cuda.Stream(device=device)
    hosts = (im1_h, im2_h, m1_h)
    slots = [[torch.empty(t.shape, dtype=t.dtype, device=device) for t in hosts] for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]        # slot uploaded
    consumed = [torch.cuda.Event() for _ in range(2)]     # slot read by the forward that used it
    pipe = {"i": 0, "primed": False}

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for dst, src in zip(slots[slot], hosts):
                dst.copy_(src, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e():
        main = torch.cuda.current_stream()
        up = None
        for nbk in sub_batches:
            slot = pipe["i"] & 1
            if not pipe["primed"]:
                upload(slot)
                pipe["primed"] = True
            upload(slot ^ 1)                               # the next batch's inputs (same synthetic data, fresh copy)
            main.wait_event(ready[slot])
            up = forward(slots[slot][0][:nbk], slots[slot][1][:nbk], slots[slot][2][:nbk])
            consumed[slot].record(main)
            done = torch.cuda.Event()
            done.record(main)
            up.record_stream(copy_stream)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                flow_host[:nbk].copy_(up, non_blocking=True)
            pipe["i"] += 1
        return up

    def drain_copies():
        torch.cuda.current_stream().wait_stream(copy_stream)

    def timed(fn, steps, warmup, sample_clocks=False, finalize=None):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            if finalize is not None:
                finalize()
            torch.cuda.synchronize()
            barrier(world)
            sampler = ClockSampler(local) if sample_clocks and rank == 0 else None
            if sampler:
                sampler.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            meter.enabled = True
            e0.record()
            for _ in range(steps):
                fn()
            if finalize is not None:
                finalize()                                 # e.g. the last download must land before the clock stops
            e1.record()
            torch.cuda.synchronize()
            meter.enabled = False
            barrier(world)
            clocks = sampler.stop() if sampler else None
        ms = max_over_ranks(e0.elapsed_time(e1), world, device)
        return ms, clocks

    ms_res, clocks = timed(step_resident, args.steps, args.warmup, sample_clocks=True)
    lookup_ms = meter.mean_ms(meter.lookup_events)
    build_ms = meter.mean_ms(meter.build_events)
    n_lookups = len(meter.lookup_events)
    launches = meter.launches
    meter.lookup_events, meter.build_events = [], []
    ms_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2), finalize=drain_copies)

    total_pairs = (world * b if wl["total"] is None else wl["total"]) * args.steps
    value = total_pairs / (ms_res * 1e-3)
    e2e_value = total_pairs / (ms_e2e * 1e-3)

    h8, w8 = PH // 8, W // 8
    n_query = sub_batches[0] * h8 * w8
    # SURVEY 8d per-query bytes; with the opt-in fp16 storage the window reads are 2-byte elements
    algo_bytes = n_query * (LOOKUP_BYTES_PER_QUERY if meter.storage_bytes == 4 else 4 * 100 * 2 + 324 * 4 + 8)
    if meter.fused_conv:      # the fused lookup + convc1 writes 256 outputs per query instead of the 324 samples
        algo_bytes = n_query * (4 * 100 * 4 + 256 * 4 + 8)
    achieved = algo_bytes / (lookup_ms * 1e-3) / 1e9 if lookup_ms else None
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "lookup_traffic.json")
    if os.path.exists(tpath) and args.config == 2 and b == BATCH and not meter.fused_conv:
        try:
            tj = json.load(open(tpath))
            key = "nhwc" if meter.nhwc else "nchw"
            traffic = tj.get("dram_bytes_per_launch_" + key, tj.get("dram_bytes_per_launch"))
            traffic_src = "static: " + tj.get("source", "ncu capture under profiles/") + " (not re-measured in this run)"
        except Exception:
            traffic = None

    if rank != 0:
        return
    stock = None
    if world == 1 and lookup_ms and build_ms and args.config == 2 and not args.no_stock:
        try:
            sb, sl = stock_gpu_corr_path(b, device)
            # the whole model with ONLY the correlation block swapped for the stock one (same host code, same inputs)
            ours_block = model.flow_net.corr_block
            model.flow_net.corr_block = StockTorchCorrBlock
            try:
                ms_stock, _ = timed(step_resident, 2, 1)
            finally:
                model.flow_net.corr_block = ours_block
            stock_pairs = world * b * 2 / (ms_stock * 1e-3)
            stock = {"what": "the reference's op sequence for the hot path through stock PyTorch CUDA kernels on this GPU "
                             "(TF32 matmul), same fmap shape", "build_ms": round(sb, 4), "lookup_ms": round(sl, 4),
                     "corr_path_ms_per_step": round(sb + ITERS * sl, 3),
                     "this_repo_corr_path_ms_per_step": round(build_ms + ITERS * lookup_ms, 3),
                     "model_pairs_per_s_with_stock_corr_block": round(stock_pairs, 3),
                     "model_pairs_per_s_with_this_repo": round(value, 3)}
            try:
                stock["reference_model"] = reference_on_this_gpu(b, device)
            except Exception as exc:
                stock["reference_model"] = {"failed": str(exc)[:200]}
        except Exception as exc:
            stock = {"failed": str(exc)[:200]}
    pwc = None
    if world == 1 and args.config == 2 and not args.no_pwc:
        try:
            pwc = pwc_level_roofline(device, peaks)
        except Exception as exc:
            pwc = {"failed": str(exc)[:200]}
    cpu = None
    if not args.no_cpu_baseline and args.config == 2:
        try:
            v, sec, cores, kind = cpu_reference_run(steps=3, warmup=1)
            cpu = {"value": round(v, 4), "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": "1 warm-up + 3 timed forwards of 1 pair, 376x1248, 12 iters (FF_RAFT_FUSION on the host cores)"}
        except Exception as exc:  # keep the GPU numbers even if the CPU arm fails
            cpu = {"value": None, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "none", "sample": f"failed: {exc}"}

    build = build_roofline(build_ms, sub_batches[0], meter.tiled, peaks, h8, w8, meter.storage_bytes)
    storage = getattr(model.flow_net, "corr_storage", None) or "fp32"
    line = {
        "metric": wl["metric"], "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_res / args.steps, 3), "higher_is_better": True,
        "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32 (fp16 GEMM operands, fp32 accumulate; TF32 host convs)", "data": "synthetic",
        "config": {"workload": wl["name"].format(b=b, w=world), "batch_per_gpu": b, "pairs_this_rank_per_step": my_pairs, "iters": ITERS,
                   "corr_precision": "fp16 operands, fp32 accumulate", "pyramid_layout": "tiled 4x4" if meter.tiled else "row-major",
                   "pyramid_storage": storage, "lookup_output": ("none: convc1 + ReLU fused into the lookup kernel" if meter.fused_conv else
                                     "channels_last (NHWC), written by the kernel" if meter.nhwc else "NCHW"),
                   "host_convs": "PyTorch/cuDNN TF32 (reference ALLOW_TF32)", "channels_last": bool(args.channels_last), "update_block_channels_last": bool(args.update_cl),
                   "context_encoder_channels_last": bool(args.cnet_cl),
                   "l2": "per-step working set (2.3 GB pyramid + activations) >> 126 MB L2, no flush needed",
                   "sharding": "by image pair, no data-path collective",
                   "e2e_pipeline": "copy stream: H2D of batch i+1 and D2H of batch i-1 overlap the compute of batch i"},
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": "lookup_convc1_kernel (ffcorr_lookup_convc1_tiled_f32: lookup + convc1 + ReLU)" if meter.fused_conv else
                               ("lookup_tiled_nhwc_kernel<4> (ffcorr_lookup_tiled_f32, out_channels_last)" if meter.nhwc else
                                "lookup_tiled_stream_kernel<4> (ffcorr_lookup_tiled_f32)") if meter.tiled else "lookup_kernel<4> (ffcorr_lookup_f32)",
                     "bound": "hbm",
                     "achieved": round(achieved, 1) if achieved else None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": round(achieved / peaks["hbm_gbs"], 4) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes,
                     "launch_ms": round(lookup_ms, 5) if lookup_ms else None, "launches_timed": n_lookups,
                     "share_of_step": round(lookup_ms * n_lookups / ms_res, 4) if lookup_ms else None,
                     "volume_plus_pyramid_ms": round(build_ms, 4) if build_ms else None,
                     "build": build, "pwc": pwc},
        "cpu_baseline": cpu,
        "stock_gpu_hot_path": stock,
    }
    emit(line)


# ------------------------------------------------------------------------------ config 3: FF-PWC forward
def run_pwc_arm(args, rank, world, local):
    """BASELINE configs[2]: FF-PWC forward (five 9x9 cost volumes per pair), Sintel shape 436x1024 (resized to 448x1024
    by the model, ff_pwcnet.py:390-402), batch 16 per GPU.  Weak scaling by pair, no collective."""
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --config 3 needs a CUDA device (the reference's FF-PWC has no CPU path either)")
    from focusflow_official_b200 import _lib
    from focusflow_official_b200 import correlation as CORR
    from focusflow_official_b200.host import FocusPWC
    from focusflow_official_b200.host import focuspwc as FP
    from weights import PWC_GAINS, fill_state_dict, synthetic_pair

    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    _lib.lib()
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    peaks, peak_src = load_peaks()
    HH, WW, b = 436, 1024, 16 if args.batch == BATCH else args.batch
    model = FocusPWC()
    sd = model.state_dict()
    fill_state_dict(sd, seed=99, gains=PWC_GAINS)
    model.load_state_dict(sd)
    model = model.to(device).eval()
    im1_h, im2_h, m1_h, _ = synthetic_pair(b, HH, WW, seed=1234 + rank)
    im1_h, im2_h, m1_h = (t.pin_memory() for t in (im1_h, im2_h, m1_h))
    im1, im2, m1 = (t.to(device) for t in (im1_h, im2_h, m1_h))
    flow_host = torch.empty((b, 2, HH, WW), dtype=torch.float32).pin_memory()

    # count and time this repo's launches: the five cost volumes and the four backwarps of a forward
    ev = {"corr": [], "warp": [], "on": False, "n": 0}
    raw_fwd, raw_warp = CORR._launch_fwd, FP.backwarp

    def fwd(one, two, leaky):
        if not ev["on"]:
            return raw_fwd(one, two, leaky)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = raw_fwd(one, two, leaky)
        e1.record()
        ev["corr"].append((e0, e1, one.shape))
        ev["n"] += 1
        return out

    def warp(x, fl, sc=1.0):
        if not ev["on"]:
            return raw_warp(x, fl, sc)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = raw_warp(x, fl, sc)
        e1.record()
        ev["warp"].append((e0, e1, x.shape))
        ev["n"] += 1
        return out

    CORR._launch_fwd, FP.backwarp = fwd, warp

    def step_resident():
        return model(im1, im2, m1, None, test_mode=True)

    def step_e2e():
        a, c, m = (t.to(device, non_blocking=True) for t in (im1_h, im2_h, m1_h))
        out = model(a, c, m, None, test_mode=True)
        flow_host.copy_(out, non_blocking=True)
        return out

    def timed(fn, steps, warmup, clocks=False):
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            barrier(world)
            sampler = ClockSampler(local) if clocks and rank == 0 else None
            if sampler:
                sampler.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev["on"] = True
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ev["on"] = False
            barrier(world)
            ck = sampler.stop() if sampler else None
        return max_over_ranks(e0.elapsed_time(e1), world, device), ck

    ms_res, clocks = timed(step_resident, args.steps, args.warmup, clocks=True)
    corr_ev, warp_ev, launches = ev["corr"], ev["warp"], ev["n"]
    ev["corr"], ev["warp"] = [], []
    ms_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    if rank != 0:
        return
    per_level = {}
    for e0, e1, shp in corr_ev:
        per_level.setdefault(tuple(shp), []).append(e0.elapsed_time(e1))
    levels, tot_ms, tot_b = [], 0.0, 0.0
    for shp, ts in sorted(per_level.items(), key=lambda kv: -kv[0][2]):
        bb, c, hh, ww = shp
        ms = sum(ts) / len(ts)
        byts = 4.0 * bb * hh * ww * (2 * c + 81)
        tot_ms += ms
        tot_b += byts
        levels.append({"C": c, "H": hh, "W": ww, "ms": round(ms, 4), "GBps": round(byts / ms / 1e6, 1),
                       "frac": round(byts / ms / 1e6 / peaks["hbm_gbs"], 4)})
    warp_ms = sum(a.elapsed_time(c) for a, c, _ in warp_ev) / max(1, args.steps)
    pairs = world * b * args.steps
    line = {
        "metric": "FF-PWC pairs/sec @436x1024, batch 16; 9x9 cost volume GB/s vs HBM peak", "value": round(pairs / (ms_res * 1e-3), 3),
        "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_res / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (TF32 host convs)", "data": "synthetic",
        "config": {"workload": f"config 3: FF-PWC forward batch {b}/GPU, Sintel shape 436x1024 (model resizes to 448x1024), B200",
                   "cost_volumes_per_forward": 5, "backwarps_per_forward": 4, "l2": "activations of a batch-16 forward >> 126 MB L2"},
        "e2e": {"value": round(pairs / (ms_e2e * 1e-3), 3), "unit": UNIT, "ms_per_step": round(ms_e2e / args.steps, 3),
                "h2d_bytes_per_step": sum(t.numel() * 4 for t in (im1_h, im2_h, m1_h)), "d2h_bytes_per_step": flow_host.numel() * 4},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"kernel": "pwc81_tma_kernel (ffcorr_pwc81_f32, leaky_relu fused), five levels of one forward", "bound": "hbm",
                     "achieved": round(tot_b / tot_ms / 1e6, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": round(tot_b / tot_ms / 1e6 / peaks["hbm_gbs"], 4), "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_forward": int(tot_b), "ms_per_forward": round(tot_ms, 4), "levels": levels,
                     "share_of_step": round(tot_ms * args.steps / ms_res, 4), "backwarp_ms_per_forward": round(warp_ms, 4)},
        "cpu_baseline": {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "none",
                         "sample": "the reference's FF-PWC has no CPU path (correlation.py:320-321 raises, ff_pwcnet.py:33 calls .cuda())"},
    }
    emit(line)


# ------------------------------------------------------------------------------ config 5: training step
def run_train_arm(args, rank, world, local):
    """BASELINE configs[4]: one FocusRAFT training step (train.py:296-328) with MixLoss, batch 8 per GPU, 368x496,
    12 iterations, stock DDP gradient all-reduce over NCCL when world > 1.  The correlation path runs this repo's
    forward AND backward kernels.  value = pairs/s of whole training steps (forward + loss + backward + all-reduce +
    clip + AdamW + scheduler)."""
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --config 5 needs a CUDA device")
    from focusflow_official_b200 import _lib
    from focusflow_official_b200 import corr as C
    from focusflow_official_b200.host import TrainStep, parallel_model

    from weights import train_inputs

    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    _lib.lib()
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    peaks, peak_src = load_peaks()
    HH, WW, IT, b = 368, 496, 12, args.batch
    model = make_model(device, False).train()
    if args.update_cl:
        model.flow_net.update_block.to(memory_format=torch.channels_last)
        model.flow_net.update_channels_last = True
    ddp = parallel_model(model, device, rank if world > 1 else -1, local)
    step = TrainStep(ddp, world_size=world, iters=IT)
    host = [t.pin_memory() for t in train_inputs(b, HH, WW, 777 + rank)]
    batch = [t.to(device) for t in host]
    loss_host = torch.empty(1).pin_memory()

    # CUDA events around this repo's launches: forward build + lookups, and the two backward entry points
    ev = {"fwd": [], "bwd": [], "on": False, "n": 0}

    def wrap(fn, key, nlaunch):
        def inner(*a, **k):
            if not ev["on"]:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            ev[key].append((e0, e1))
            ev["n"] += nlaunch
            return out
        return inner

    C._volume_pyramid_tiled_raw = wrap(C._volume_pyramid_tiled_raw, "fwd", 2)
    C._volume_pyramid_raw = wrap(C._volume_pyramid_raw, "fwd", 3)
    C._lookup_tiled_raw = wrap(C._lookup_tiled_raw, "fwd", 1)
    C._lookup_raw = wrap(C._lookup_raw, "fwd", 1)
    C._Lookup.backward = staticmethod(wrap(C._Lookup.backward, "bwd", 1))
    C._VolumePyramid.backward = staticmethod(wrap(C._VolumePyramid.backward, "bwd", 4))

    last = {"loss": None}

    def one_step(from_host=False):
        data = [t.to(device, non_blocking=True) for t in host] if from_host else batch
        loss, _ = step(*data)
        if from_host:
            loss_host.copy_(loss.reshape(1), non_blocking=True)
        last["loss"] = loss
        return loss

    def timed(fn, steps, warmup, clocks=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        barrier(world)
        sampler = ClockSampler(local) if clocks and rank == 0 else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev["on"] = True
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ev["on"] = False
        barrier(world)
        ck = sampler.stop() if sampler else None
        return max_over_ranks(e0.elapsed_time(e1), world, device), ck

    ms_res, clocks = timed(one_step, args.steps, args.warmup, clocks=True)
    fwd_ms = sum(a.elapsed_time(c) for a, c in ev["fwd"]) / args.steps
    bwd_ms = sum(a.elapsed_time(c) for a, c in ev["bwd"]) / args.steps
    launches = ev["n"]
    ev["fwd"], ev["bwd"] = [], []
    ms_e2e, _ = timed(lambda: one_step(True), args.steps, 1)
    exposed = None
    if world > 1:   # the same steps without the gradient all-reduce: the difference is the exposed communication time
        with ddp.no_sync():
            ms_nosync, _ = timed(one_step, args.steps, 1)
        exposed = (ms_res - ms_nosync) / args.steps
    if rank != 0:
        return
    pairs = world * b * args.steps
    nparam = sum(p.numel() for p in model.parameters())
    line = {
        "metric": "FF-RAFT training pairs/sec @368x496, 12 iters, MixLoss", "value": round(pairs / (ms_res * 1e-3), 3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_res / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (fp16 forward GEMM operands, tf32 backward GEMMs; TF32 host convs)",
        "data": "synthetic",
        "config": {"workload": f"config 5: FocusRAFT training step, batch {b}/GPU, 368x496, 12 iters, MixLoss, AdamW + OneCycleLR, "
                               f"{'stock DDP (NCCL all-reduce of %.1f MB fp32 gradients)' % (nparam * 4 / 1e6) if world > 1 else 'single GPU'}",
                   "batch_per_gpu": b, "iters": IT, "batch_norm": "training mode (chairs stage, train.py:192)",
                   "l2": "activations of a 12-iteration training step >> 126 MB L2"},
        "e2e": {"value": round(pairs / (ms_e2e * 1e-3), 3), "unit": UNIT, "ms_per_step": round(ms_e2e / args.steps, 3),
                "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host), "d2h_bytes_per_step": 4},
        "gpu_launches": launches, "clocks": clocks,
        "train": {"native_corr_forward_ms_per_step": round(fwd_ms, 3), "native_corr_backward_ms_per_step": round(bwd_ms, 3),
                  "native_corr_share_of_step": round((fwd_ms + bwd_ms) * args.steps / ms_res, 4),
                  "allreduce_exposed_ms_per_step": None if exposed is None else round(exposed, 3),
                  "allreduce_payload_bytes": nparam * 4, "loss": float(last["loss"].item())},
        "roofline": {"kernel": "correlation path of a training step: fused build + 12 tiled lookups forward; 12 lookup_bwd + pyramid_bwd + 2 tf32 GEMMs backward",
                     "bound": "hbm", "achieved": None, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src},
        "cpu_baseline": None,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 2 = KITTI inference (the headline, default), 3 = FF-PWC forward batch 16, "
                         "4 = Sintel 32-iteration inference, 64 pairs sharded over the ranks, 5 = training step with DDP")
    ap.add_argument("--storage", default=None, choices=[None, "fp32", "fp16"], help="pyramid storage of the CorrBlock (opt-in fp16)")
    ap.add_argument("--fuse-convc1", action="store_true", help="lookup + convc1 + ReLU in one kernel (CorrBlock.lookup_conv)")
    ap.add_argument("--no-stock", action="store_true", help="skip the stock-PyTorch / reference-on-this-GPU comparison")
    ap.add_argument("--no-pwc", action="store_true", help="skip the config-3 cost-volume roofline entry")
    ap.add_argument("--no-cudnn-benchmark", dest="no_cudnn_benchmark", action="store_true", default=False,
                    help="heuristic cuDNN algorithm choice instead of timing-based autotuning (profiler runs: the "
                         "autotuner mis-times kernels under ncu and picks different engines)")
    ap.add_argument("--channels-last", dest="channels_last", action="store_true", default=False)
    ap.add_argument("--cnet-nchw", dest="cnet_cl", action="store_false", default=True,
                    help="keep the context encoder in NCHW instead of channels_last (host plumbing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--update-nchw", dest="update_cl", action="store_false", default=True,
                    help="run the GRU update block in NCHW instead of channels_last (host plumbing)")
    args = ap.parse_args()
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference_arm(args, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return
    rank, world, local = dist_setup(args.gpus)
    try:
        if args.config == 3:
            run_pwc_arm(args, rank, world, local)
        elif args.config == 5:
            run_train_arm(args, rank, world, local)
        else:
            run_gpu_arm(args, rank, world, local)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
