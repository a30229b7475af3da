#!/usr/bin/env python
"""Benchmark of the Real-SRGD sampling hot path on B200 (see DESIGN.md "Measurement").

A *step* is one pass of the hot path over one batch: the conditional U-Net denoise (2x batch under
classifier-free guidance) + the fused posterior update, at the times of the shipped 250-step
linear-logSNR schedule.  `--workload` selects the BASELINE.json configuration (default = configs[1]):

    sample16   configs[1]  sample(): 16 synthetic 64x64-LR tiles (256x256) per GPU, label 0, scale 1.0
    cfg32      configs[2]  sample(): 32 tiles per GPU, class_cond_scale 3.0 (64-row U-Net batch), test_label swept 0,1,2
    tiled512   configs[3]  tiled_sample(): ONE 512x512-LR image (2304^2 canvas, 81 / 64 tiles per step), tiles sharded
                           over the GPUs with one all-gather per step (exact mode; strong scaling)
    tiled128   configs[4]  tiled_sample(): `--images` 128x128-LR images per GPU advancing together (768^2 canvases,
                           9 / 4 tiles per image and step)
    sweep128   configs[4]  the same for 1, 2, 4 ... 256 images per GPU (one line, `sweep` array)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference      # CPU arm: the oracle port of the reference on the host cores

Prints ONE JSON line (rank 0).  `value` = SR images/sec over all GPUs with inputs resident in HBM (K timed
steps, CUDA events, max over ranks); `e2e` = the same through the reference-facing API with pinned HOST
buffers copied in and out every step, timed in segments that ALTERNATE with the device-resident segments (same
clock / power state); `roofline` = the tcgen05 conv kernel's achieved TFLOP/s (CUDA events around every conv
launch, srgd_profile_*) against the measured bf16 peak; `cpu_baseline` = the oracle port of the reference on
the host cores (bounded sample); `gpu_eager_reference` = the same oracle through stock PyTorch (cuDNN / cuBLAS)
on the same B200, TF32 and bf16-autocast -- the bar SURVEY.md section 2.1 names.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SAMPLE_STEPS = 250                    # shipped schedule (conf yaml:17)
TILE = 256                            # image_size (conf yaml:31): one 64x64 LR image = one tile
# algorithmic work per tile-NFE measured on the reference module (SURVEY.md §8d / BASELINE.md §2)
TOTAL_GFLOP_PER_TILE_NFE = 793.8
WORKLOADS = {
    "sample16": dict(kind="sample", batch=16, ccs=1.0, config=1),
    "cfg32": dict(kind="sample", batch=32, ccs=3.0, config=2),
    "tiled512": dict(kind="tiled", lr=512, images=1, shard=True, config=3),
    "tiled128": dict(kind="tiled", lr=128, images=16, shard=False, config=4),
    "sweep128": dict(kind="sweep", lr=128, images=16, shard=False, config=4),
}
SWEEP_BATCHES = (1, 2, 4, 8, 16, 32, 64, 128, 256)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 250 = one full schedule; tiled: 20)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="sample16")
    ap.add_argument("--batch", type=int, default=None, help="override the tiles per GPU of a sample() workload")
    ap.add_argument("--images", type=int, default=None, help="override the images per GPU of tiled128")
    ap.add_argument("--class_cond_scale", type=float, default=None)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--cpu_steps", type=int, default=2, help="timed CPU-baseline steps (one tile each)")
    ap.add_argument("--cpu_budget_s", type=float, default=150.0, help="--impl reference: wall-clock bound of the run")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--no_gpu_eager", action="store_true")
    ap.add_argument("--dump_launches", type=str, default=None,
                    help="write the per-launch CUDA-event table of one profiled step (kind, ms, TFLOP/s, GB/s) here")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16_tflops=p.get("bf16_tflops_sustained", p.get("bf16_tflops")), hbm_gbs=p.get("hbm_gbs"),
                    bf16_tflops_burst=p.get("bf16_tflops"), source="MEASURED_PEAKS.json (sustained bf16)")
    return dict(bf16_tflops=1400.0, hbm_gbs=6650.0, bf16_tflops_burst=None, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = sorted(float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.startswith("Active")})
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    power_w=pw[len(pw) // 2] if pw else None, samples=len(self.rows))


def synth_lr(lr_size, idx):
    """Synthetic LR image exactly as BASELINE.md §3: RandomState(71+idx) uint8, PIL bicubic x4 -> [3,4h,4w] in [0,1]."""
    import numpy as np
    from PIL import Image
    lr = np.random.RandomState(71 + idx).randint(0, 256, (lr_size, lr_size, 3), dtype=np.uint8)
    hr = Image.fromarray(lr, mode="RGB").resize((4 * lr_size, 4 * lr_size), resample=Image.BICUBIC)
    return torch.from_numpy(np.array(hr, dtype=np.uint8)).permute(2, 0, 1).float().div(255.)


def synth_inputs(batch, first=0, lr_size=64):
    return torch.stack([synth_lr(lr_size, first + i) for i in range(batch)])


def tile_nfe_per_image(w, ccs):
    """Tile forwards one finished image costs: 250 steps x tiles per step (x2 under guidance)."""
    g = 2 if ccs != 1.0 else 1
    if w["kind"] == "sample":
        return SAMPLE_STEPS * g
    from srgd_b200.tiling import TilePlan
    return TilePlan(4 * w["lr"], 4 * w["lr"]).tiles_per_image(SAMPLE_STEPS) * g


def describe(name, w, batch, images, ccs):
    if w["kind"] == "sample":
        lab = "test_label swept 0,1,2" if name == "cfg32" else "label 0"
        return (f"sample(): {batch} synthetic 64x64-LR tiles (256x256) per GPU, {lab}, class_cond_scale {ccs}, "
                f"{SAMPLE_STEPS}-step linear-logSNR schedule, dim128 U-Net")
    n = "1, 2, 4 ... 256" if w["kind"] == "sweep" else str(images)
    c = 4 * w["lr"] + 256                                  # canvas = HR size rounded up to 256 + one tile of padding
    return (f"tiled_sample(): {n} synthetic {w['lr']}x{w['lr']}-LR image(s) per "
            f"{'job, tiles sharded over the GPUs' if w['shard'] else 'GPU'} ({c}x{c} canvas, alternating tile grids), "
            f"label 0, class_cond_scale {ccs}, {SAMPLE_STEPS}-step linear-logSNR schedule, dim128 U-Net")


# ------------------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute oracle/): cpu_baseline of the product line and `--impl reference`
# ------------------------------------------------------------------------------------------------------------
def cpu_tile_steps(ccs, n_steps, warm, budget_s=None):
    """p_sample of ONE 256x256 tile through the oracle port on all host threads; returns the timed steps."""
    from oracle import srgd_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    spec = O.UnetSpec()
    sd = O.make_state_dict(spec, 1234)
    cond = synth_inputs(1) * 2 - 1
    g = torch.Generator().manual_seed(71)
    x = torch.randn(1, 3, TILE, TILE, generator=g)
    steps = torch.linspace(1., 0., SAMPLE_STEPS + 1)
    label = torch.tensor([0])
    times = []
    t_start = time.perf_counter()
    with torch.inference_mode():
        for i in range(warm + n_steps):
            t0 = time.perf_counter()
            x, _ = O.p_sample(sd, spec, x, steps[i % SAMPLE_STEPS], cond, label, 1.0, ccs, steps[i % SAMPLE_STEPS + 1],
                              generator=g)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
            # bounded: stop when the next step would overrun the budget (the line reports the steps really timed)
            if budget_s is not None and times and (time.perf_counter() - t_start) + 1.5 * dt > budget_s:
                break
    return times, threads


def cpu_baseline(w, ccs, n_steps, warm, budget_s=None):
    times, threads = cpu_tile_steps(ccs, n_steps, warm, budget_s)
    s_per_step = sum(times) / len(times)
    nfe = 2 if ccs != 1.0 else 1
    per_image = tile_nfe_per_image(w, ccs)
    img_per_s = nfe / s_per_step / per_image
    return dict(value=img_per_s, unit="images/s", cores=threads, kind="port",
                sample=f"{len(times)} timed p_sample steps (+{warm} warm-up) of ONE 256x256 tile, fp32, "
                       f"class_cond_scale {ccs} ({nfe} U-Net forward(s) per step)",
                derivation=f"images/s = tile forwards per second / {per_image} tile forwards per finished image",
                s_per_step=s_per_step, steps_timed=len(times), unet_steps_per_sec=nfe / s_per_step)


def gpu_eager_reference(dev, ccs, w, batch=16, warm=2, timed=3):
    """The reference's algorithm through stock PyTorch (cuDNN / cuBLAS / ATen) on THIS B200: the oracle port with its
    weights on the GPU, p_sample on `batch` tiles, as the reference runs on a GPU (TF32 allowed, cudnn.benchmark:
    inference.py:52-56) and under bf16 autocast.  Reported as images/s of the bench workload."""
    from oracle import srgd_oracle as O
    spec = O.UnetSpec()
    sd = {k: v.to(dev) for k, v in O.make_state_dict(spec, 1234).items()}
    g = torch.Generator().manual_seed(3)
    x = torch.randn(batch, 3, TILE, TILE, generator=g).to(dev)
    cond = (torch.rand(batch, 3, TILE, TILE, generator=g) * 2 - 1).to(dev)
    label = torch.tensor([0], device=dev)
    steps = torch.linspace(1., 0., SAMPLE_STEPS + 1).to(dev)
    per_image = tile_nfe_per_image(w, ccs)
    nfe = batch * (2 if ccs != 1.0 else 1)
    out = {}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    try:
        for name, ctx in (("tf32", None), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.inference_mode():
                for k in range(warm + timed):
                    if k == warm:
                        torch.cuda.synchronize()
                        e0.record()
                    i = 100 + k
                    if ctx is None:
                        O.p_sample(sd, spec, x, steps[i], cond, label, 1.0, ccs, steps[i + 1])
                    else:
                        with ctx:
                            O.p_sample(sd, spec, x, steps[i], cond, label, 1.0, ccs, steps[i + 1])
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / timed
            out[name] = dict(ms_per_step=ms, unet_steps_per_sec=nfe / (ms * 1e-3),
                             value=nfe / (ms * 1e-3) / per_image, unit="images/s")
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = old
        del sd
        torch.cuda.empty_cache()
    out["what"] = (f"oracle port of the reference through stock PyTorch on the same B200: p_sample on {batch} 256x256 "
                   f"tiles, class_cond_scale {ccs}, {timed} timed steps (+{warm} warm-up), CUDA events")
    return out


# ------------------------------------------------------------------------------------------------------------
# workloads on the product path
# ------------------------------------------------------------------------------------------------------------
class SampleWorkload:
    """sample()'s loop body: p_sample on B tiles (configs[1], configs[2])."""

    def __init__(self, diff, dev, rank, batch, ccs, sweep_labels):
        self.diff, self.dev, self.B, self.ccs = diff, dev, batch, ccs
        self.cond01 = synth_inputs(batch, first=rank * batch)
        self.cond = (self.cond01 * 2 - 1).to(dev)
        self.labels = [torch.tensor([k], device=dev) for k in ((0, 1, 2) if sweep_labels else (0,))]
        self.steps = torch.linspace(1., 0., SAMPLE_STEPS + 1)
        self.nfe_per_step = batch * (2 if ccs != 1.0 else 1)
        self.images = batch
        torch.manual_seed(71 + rank)
        self.img = torch.randn(batch, 3, TILE, TILE, device=dev)
        self.h2d_bytes = 2 * batch * 3 * TILE * TILE * 4
        self.d2h_bytes = batch * 3 * TILE * TILE * 4
        self._e2e = None

    def run(self, first, count):
        d = self.diff
        for k in range(count):
            i = (first + k) % SAMPLE_STEPS
            self.img, _ = d.p_sample(self.img, self.steps[i], self.cond, self.labels[(first + k) % len(self.labels)],
                                     1.0, self.ccs, self.steps[i + 1])

    def result(self):
        return self.diff._finalize(self.img)

    # -- end to end: pinned host x / cond in, img_next out, every step; copies double-buffered on a second stream --
    def _e2e_setup(self):
        B, dev = self.B, self.dev
        shape = (B, 3, TILE, TILE)
        e = self._e2e = dict()
        e["x_host"] = torch.randn(shape).pin_memory()
        e["c_host"] = (self.cond01 * 2 - 1).pin_memory()
        e["r_host"] = [torch.empty(shape).pin_memory() for _ in range(3)]
        e["xd"] = [torch.empty(shape, device=dev) for _ in range(2)]
        e["cd"] = [torch.empty(shape, device=dev) for _ in range(2)]
        e["od"] = [torch.empty(shape, device=dev) for _ in range(3)]       # results staged for the copy stream
        e["copy"] = torch.cuda.Stream(device=dev)
        e["h2d"] = [torch.cuda.Event() for _ in range(2)]
        e["comp"] = [torch.cuda.Event() for _ in range(2)]
        e["staged"] = [torch.cuda.Event() for _ in range(3)]
        e["d2h"] = [torch.cuda.Event() for _ in range(3)]

    def run_e2e(self, first, count):
        """Step k+1's inputs travel while step k computes and step k's result travels while step k+1 computes (what a
        serving loop around p_sample does); the host waits for (= can read) the result of step k-2 before it issues
        step k+1.  All buffers are preallocated and every cross-stream hand-over is an event: no record_stream, so the
        caching allocator never has to defer or re-allocate blocks between the two streams."""
        if self._e2e is None:
            self._e2e_setup()
        e, d = self._e2e, self.diff
        cur, cp = torch.cuda.current_stream(), e["copy"]

        def issue_h2d(slot):
            with torch.cuda.stream(cp):
                cp.wait_event(e["comp"][slot])                 # the step that last read this slot has finished
                e["xd"][slot].copy_(e["x_host"], non_blocking=True)
                e["cd"][slot].copy_(e["c_host"], non_blocking=True)
                e["h2d"][slot].record(cp)

        for ev in e["comp"]:
            ev.record(cur)
        for ev in e["d2h"]:
            ev.record(cp)
        issue_h2d(0)
        for k in range(count):
            slot, o3 = k & 1, k % 3
            i = (first + k) % SAMPLE_STEPS
            cur.wait_event(e["h2d"][slot])
            o, _ = d.p_sample(e["xd"][slot], self.steps[i], e["cd"][slot], self.labels[(first + k) % len(self.labels)],
                              1.0, self.ccs, self.steps[i + 1])
            e["comp"][slot].record(cur)
            cur.wait_event(e["d2h"][o3])                       # the previous user of this staging buffer has left
            e["od"][o3].copy_(o)                               # 12.6 MB device copy on the compute stream
            e["staged"][o3].record(cur)
            if k + 1 < count:
                issue_h2d(slot ^ 1)
            with torch.cuda.stream(cp):
                cp.wait_event(e["staged"][o3])
                e["r_host"][o3].copy_(e["od"][o3], non_blocking=True)
                e["d2h"][o3].record(cp)
            if k >= 2:
                e["d2h"][(k - 2) % 3].synchronize()            # result of step k-2 is in host memory
        for ev in e["d2h"]:
            ev.synchronize()
        cur.wait_stream(cp)

    e2e_call = ("ConditionalContinuousTimeGaussianDiffusionSR.p_sample with pinned host x/cond in, img_next out, every "
                "step; copies double-buffered on a second stream")


class TiledWorkload:
    """tiled_sample()'s loop body (srgd_b200/tiled.py run_tiled) on `images` canvases (configs[3], configs[4])."""

    def __init__(self, diff, dev, rank, world, lr, images, ccs, shard):
        import torch.nn.functional as F
        from srgd_b200.tiled import CudaTiledOps
        from srgd_b200.tiling import TilePlan
        self.diff, self.dev, self.ccs, self.shard, self.rank = diff, dev, ccs, shard, rank
        self.images = images
        first = 0 if shard else rank * images                 # sharded: every rank holds the same image
        cond = synth_inputs(images, first=first, lr_size=lr) * 2 - 1
        self.plan = plan = TilePlan(4 * lr, 4 * lr)
        cond = F.pad(cond.to(dev), plan.canvas_pad, mode="reflect")
        it, ib, il, ir = plan.inner
        self.cond_canvas = torch.zeros_like(cond)
        self.cond_canvas[:, :, it:ib, il:ir] = cond[:, :, it:ib, il:ir]
        torch.manual_seed(71 + (0 if shard else rank))
        self.img = torch.randn((1,) + tuple(cond.shape[1:]), device=dev).expand(images, -1, -1, -1).contiguous()
        self.ops = CudaTiledOps(diff)
        self.steps = torch.linspace(1., 0., SAMPLE_STEPS + 1)
        self.label = torch.tensor([0], device=dev)
        g = 2 if ccs != 1.0 else 1
        n0, n1 = len(plan.grids[0]), len(plan.grids[1])
        self.tiles_per_step = (n0 + n1) / 2.0 * images        # mean of the two grids (K is kept even)
        self.nfe_per_step = self.tiles_per_step * g / (world if shard else 1)   # per rank
        self.world = world if shard else 1
        self.batch_size = 8                                   # the CLI default (shapes the noise draws)
        self.h2d_bytes = 2 * self.img.numel() * 4
        self.d2h_bytes = self.img.numel() * 4
        self._host = None

    def _steps(self, first, count):
        from srgd_b200.tiled import run_tiled
        done = 0
        while done < count:                                   # contiguous index ranges inside [0, 249): never the
            i0 = (first + done) % (SAMPLE_STEPS - 2)          # no-noise last step, parity continues across the wrap
            n = min(count - done, SAMPLE_STEPS - 2 - i0)
            self.img, _ = run_tiled(self.ops, self.img, self.cond_canvas, self.plan, self.steps, i0 + n,
                                    self.batch_size, self.label, 1.0, 0, self.ccs, 0, generation_start_steps=i0,
                                    shard=self.shard)
            done += n

    def run(self, first, count):
        self._steps(first, count)

    def result(self):
        t, b, l, r = self.plan.crop
        return self.diff._finalize(self.img[:, :, t:b, l:r].contiguous())

    def run_e2e(self, first, count):
        """Per step: state canvas and condition canvas host -> device (every rank needs them), one sampling step, next
        canvas device -> host.  The result is replicated in the tile-sharded workload, so rank 0 alone reads it back;
        the read-back runs on a second stream from a staging copy and overlaps the next step."""
        if self._host is None:
            self._host = dict(x=self.img.cpu().pin_memory(), c=self.cond_canvas.cpu().pin_memory(),
                              r=torch.empty(self.img.shape).pin_memory(), stage=torch.empty_like(self.img),
                              copy=torch.cuda.Stream(device=self.dev), staged=torch.cuda.Event(),
                              done=torch.cuda.Event())
        h = self._host
        cur, cp = torch.cuda.current_stream(), h["copy"]
        reads_back = (not self.shard) or self.rank == 0
        h["done"].record(cp)
        for k in range(count):
            self.img.copy_(h["x"], non_blocking=True)
            self.cond_canvas.copy_(h["c"], non_blocking=True)
            self._steps(first + k, 1)
            if reads_back:
                cur.wait_event(h["done"])                      # the previous read-back has left the staging canvas
                h["stage"].copy_(self.img)
                h["staged"].record(cur)
                with torch.cuda.stream(cp):
                    cp.wait_event(h["staged"])
                    h["r"].copy_(h["stage"], non_blocking=True)
                    h["done"].record(cp)
        h["done"].synchronize()
        cur.wait_stream(cp)
        cur.synchronize()

    e2e_call = ("one sampling step of tiled_sample()'s loop (srgd_b200.tiled.run_tiled) per call with the pinned host "
                "state canvas and condition canvas copied in (every rank) and the next canvas copied out (rank 0 in the "
                "tile-sharded workload: the result is replicated), every step; the read-back overlaps the next step")


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    name = args.workload
    w = dict(WORKLOADS[name])
    ccs = args.class_cond_scale if args.class_cond_scale is not None else w.get("ccs", 1.0)
    batch = args.batch if args.batch is not None else w.get("batch")
    images = args.images if args.images is not None else w.get("images")
    if args.steps is None:
        args.steps = SAMPLE_STEPS if w["kind"] == "sample" else 20
    if w["kind"] != "sample" and args.steps % 2:
        args.steps += 1                                       # equal numbers of aligned-grid and shifted-grid steps
    workload = describe(name, w, batch, images, ccs)
    lr = w.get("lr", 64)
    metric = f"SR images/sec ({lr}x{lr} LR -> {4 * lr}x{4 * lr}, 250 sampling steps)"
    scaling = "strong" if w.get("shard") else "weak"

    if args.impl == "reference":
        # The reference's own algorithm on the box's host cores (the oracle port: the reference is Python and does not
        # travel to the GPU box).  Every step is a bounded sample of the workload -- ONE 256x256 tile through p_sample --
        # and the line reports the steps that were really timed.
        if rank != 0:
            return
        cb = cpu_baseline(w, ccs, max(1, args.steps), max(1, min(args.warmup, 2)), budget_s=args.cpu_budget_s)
        line = dict(metric=metric, value=cb["value"], unit="images/s", n_gpus=args.gpus, steps=cb["steps_timed"],
                    warmup=max(1, min(args.warmup, 2)), requested_steps=args.steps, requested_warmup=args.warmup,
                    ms_per_step=cb["s_per_step"] * 1e3, higher_is_better=True, scaling=scaling, vs_baseline=None,
                    dtype="f32", data="synthetic", impl="reference",
                    config=dict(workload=workload, sample="each timed step = p_sample on ONE 256x256 tile of that "
                                                          "workload (bounded sample; all host threads)",
                                derivation=cb["derivation"], weights="random-init seed 1234"),
                    cpu_baseline=dict(value=cb["value"], unit="images/s", cores=cb["cores"], kind="port",
                                      sample=cb["sample"]),
                    unet_steps_per_sec=cb["unet_steps_per_sec"],
                    e2e=dict(value=cb["value"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import torch.distributed as dist
    import model as M
    from srgd_b200 import _lib, arch, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: the product arm needs an sm_100 (B200) CUDA device -- there is no CPU fallback "
                         "(`--impl reference` times the reference's algorithm on the host cores)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec = arch.UnetSpec()
    unet = M.ConditionalSRUnet(dim=128, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=TILE, num_sample_steps=SAMPLE_STEPS)
    diff.load_state_dict(arch.seeded_state_dict(spec, 1234), strict=True)
    diff = diff.eval().to(dev)
    diff.progress = False
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return ms

    def make(images_override=None):
        if w["kind"] == "sample":
            return SampleWorkload(diff, dev, rank, batch, ccs, sweep_labels=(name == "cfg32"))
        return TiledWorkload(diff, dev, rank, world, lr, images_override or images, ccs, w["shard"])

    def timed(fn):
        """fn() between a barrier + synchronize on both sides, CUDA events on the launching stream -> ms (this rank)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    with torch.inference_mode():
        if w["kind"] == "sweep":
            sweep = []
            for nimg in SWEEP_BATCHES:
                wl = make(nimg)
                wl.run(100, 2)
                ms = max_over_ranks(timed(lambda: wl.run(102, 4))) / 4
                g = 2 if ccs != 1.0 else 1
                sweep.append(dict(images_per_gpu=nimg, ms_per_step=ms,
                                  images_per_sec=world * nimg / (ms * 1e-3 * SAMPLE_STEPS),
                                  unet_steps_per_sec=world * wl.tiles_per_step * g / (ms * 1e-3)))
                del wl
                torch.cuda.empty_cache()
        wl = make()
        first = 100 - args.warmup
        wl.run(first, args.warmup)                            # untimed warm-up steps
        out = wl.result()
        if world > 1 and not w.get("shard"):                  # warm-up covers the gather too (NCCL channel set-up)
            sharding.gather_rows(out, [out.shape[0]] * world, dst=0)
        wl.run_e2e(first, 2)                                  # ... and the host buffers / copy stream of the e2e leg

        # ---- K device-resident steps and K end-to-end steps, in alternating segments ----
        K = args.steps
        rounds = 2 if K >= 4 else 1
        seg = [K // rounds + (1 if r < K % rounds else 0) for r in range(rounds)]
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        launches0 = lib.srgd_launch_count()
        res_ms = e2e_ms = 0.0
        res_launches = 0
        pos = 100
        for r in range(rounds):
            l0 = lib.srgd_launch_count()
            if r == rounds - 1:
                def last_segment():
                    wl.run(pos, seg[r])
                    fin = wl.result()                         # clamp + [0,1] of the finished images ...
                    if world > 1 and not w.get("shard"):      # ... and the only collective of the image-sharded
                        sharding.gather_rows(fin, [fin.shape[0]] * world, dst=0)   # workloads: one final gather
                res_ms += timed(last_segment)
            else:
                res_ms += timed(lambda: wl.run(pos, seg[r]))
            res_launches += lib.srgd_launch_count() - l0
            e2e_ms += timed(lambda: wl.run_e2e(pos, seg[r]))
            pos += seg[r]
        clocks = sampler.summary() if sampler else None
        wl_bytes, wl_call = (wl.h2d_bytes, wl.d2h_bytes), wl.e2e_call
        elapsed_ms = max_over_ranks(res_ms)
        e2e_ms = max_over_ranks(e2e_ms)

        # ---- per-kernel-kind device time (CUDA events around every launch), rank 0 only ----
        prof = None
        if rank == 0 or w.get("shard"):                       # sharded steps contain a collective: all ranks take part
            prof_steps = 2 if w["kind"] != "sample" else 3
            torch.cuda.synchronize()
            _lib.check(lib.srgd_profile_begin())
            wl.run(100, prof_steps)
            _lib.check(lib.srgd_profile_end())
            prof = {k: dict(v, ms=v["ms"] / prof_steps, launches=v["launches"] / prof_steps,
                            flops=v["flops"] / prof_steps, bytes=v["bytes"] / prof_steps)
                    for k, v in _lib.profile_report().items()}
            recs_all = _lib.profile_records() if rank == 0 else []
            if args.dump_launches and rank == 0:
                recs = _lib.profile_records()
                per = len(recs) // prof_steps
                with open(args.dump_launches, "w") as f:
                    f.write("# one warm step of the bench workload, CUDA events around every library launch\n")
                    f.write("idx kind ms tflops gbs\n")
                    for i, (kind, ms, fl, by) in enumerate(recs[-per:]):
                        f.write(f"{i} {kind} {ms:.4f} {fl / (ms * 1e-3) / 1e12 if ms > 0 else 0:.1f} "
                                f"{by / (ms * 1e-3) / 1e9 if ms > 0 else 0:.0f}\n")
        # ---- the exchange step of the tile-sharded workload, timed alone: one all_gather_into_tensor of the largest
        # per-rank tile stack (what run_tiled issues once per step), CUDA events, max over ranks ----
        exchange = None
        if w.get("shard") and world > 1:
            from srgd_b200.sharding import shard_range
            n_even = len(wl.plan.grids[0])
            width = max(hi - lo for lo, hi in (shard_range(n_even, world, r) for r in range(world)))
            send = torch.zeros(images, width, 3, TILE, TILE, device=dev)
            recv = torch.empty((world,) + tuple(send.shape), device=dev)
            for _ in range(3):
                dist.all_gather_into_tensor(recv.flatten(0, 1), send)
            ex_ms = max_over_ranks(timed(lambda: [dist.all_gather_into_tensor(recv.flatten(0, 1), send)
                                                  for _ in range(20)])) / 20
            exchange = dict(collective="all_gather_into_tensor (NCCL), once per sampling step",
                            bytes_per_rank=send.numel() * 4, bytes_gathered=recv.numel() * 4, ms_standalone=ex_ms,
                            note="issued asynchronously; the odd steps' full-canvas noise draw overlaps it")
            del send, recv
        eager = None
        if rank == 0 and world == 1 and not args.no_gpu_eager:
            del wl
            torch.cuda.empty_cache()
            try:
                eager = gpu_eager_reference(dev, ccs, w)
            except Exception as ex:                           # the checker's leg must never take the product line down
                eager = dict(error=f"{type(ex).__name__}: {ex}"[:300])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    g = 2 if ccs != 1.0 else 1
    ms_per_step = elapsed_ms / K
    if w["kind"] == "sample":
        nfe_per_step_job = world * batch * g
        nfe_rank = batch * g
        images_job = world * batch
    else:
        from srgd_b200.tiling import TilePlan
        plan = TilePlan(4 * lr, 4 * lr)
        tiles = (len(plan.grids[0]) + len(plan.grids[1])) / 2.0 * images
        nfe_per_step_job = tiles * g * (1 if w["shard"] else world)
        nfe_rank = tiles * g / (world if w["shard"] else 1)
        images_job = images * (1 if w["shard"] else world)
    nfe_per_sec = nfe_per_step_job / (ms_per_step * 1e-3)
    img_per_sec = images_job / (ms_per_step * 1e-3 * SAMPLE_STEPS)
    e2e_img_per_sec = images_job / (e2e_ms / K * 1e-3 * SAMPLE_STEPS)
    pk = peaks()
    conv = prof["conv_igemm"]
    # FLOPs actually executed by the conv launches (2*M*N*K credited per launch by the library); the 1x1 convs of
    # the fused LinearAttention blocks are not in this kernel
    conv_tflops = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
    step_kernel_ms = sum(v["ms"] for v in prof.values())
    # DRAM traffic of the same kernel from the committed `ncu --set full` capture of the default workload at HEAD
    # (profiles/ncu_full_conv_latest.json): dram__bytes_read.sum + dram__bytes_write.sum averaged over the conv launches
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_full_conv_latest.json")
    if os.path.exists(tpath) and name == "sample16" and batch == 16 and ccs == 1.0:
        try:
            kk = json.load(open(tpath))
            rows = kk["kernels"] if isinstance(kk, dict) else kk
            traffic = sum(o["dram_bytes"] for o in rows) / max(len(rows), 1)
            traffic_src = kk.get("source") if isinstance(kk, dict) else None
        except Exception:
            traffic = None
    roof = dict(bound="tensor", kernel="conv_igemm_kernel / conv_igemm_t_kernel (tcgen05 implicit GEMM, every "
                                       "conv3x3 / 1x1 / 7x7 launch of a step)",
                achieved=conv_tflops, peak=pk["bf16_tflops"], unit="TFLOP/s", frac=conv_tflops / pk["bf16_tflops"],
                traffic=traffic,
                traffic_unit="DRAM bytes per conv launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, "
                             "profiles/ncu_full_conv_latest.json" + (f": {traffic_src}" if traffic_src else "") + ")",
                achieved_per_launch_flops=conv["flops"] / max(conv["launches"], 1), peak_source=pk["source"],
                frac_of_burst_peak=(conv_tflops / pk["bf16_tflops_burst"]) if pk.get("bf16_tflops_burst") else None,
                algorithmic=f"{conv['flops'] / 1e9 / nfe_rank:.2f} GFLOP per tile-NFE (2*M*N*K of every conv launch) x "
                            f"{nfe_rank:g} tile-NFE per step and GPU",
                launches_per_step=conv["launches"], kernel_ms_per_step=conv["ms"],
                share_of_step=conv["ms"] / step_kernel_ms if step_kernel_ms > 0 else None,
                share_note="share of the summed per-launch CUDA-event times of one step (profiled steps run without "
                           "launch overlap)")
    hbm = {}
    for k in ("gn_apply", "sampler_step", "linear_attention", "norm_misc"):
        if prof[k]["ms"] > 0:
            gbs = prof[k]["bytes"] / (prof[k]["ms"] * 1e-3) / 1e9
            hbm[k] = dict(ms_per_step=prof[k]["ms"], gb_per_s=gbs, frac_of_hbm_peak=gbs / pk["hbm_gbs"],
                          launches_per_step=prof[k]["launches"])
            # the same over the launches that move more than the 126 MB L2 holds (the others are latency-bound)
            big = [(ms, by) for kind, ms, fl, by in recs_all if kind == k and by >= 2.56e8 and ms > 0]
            if big:
                gb = sum(b for _, b in big) / (sum(m for m, _ in big) * 1e-3) / 1e9
                hbm[k]["launches_over_256MB"] = dict(count_per_step=len(big) / prof_steps, gb_per_s=gb,
                                                     frac_of_hbm_peak=gb / pk["hbm_gbs"])
    cfg = dict(workload=workload, baseline_config=f"BASELINE.json configs[{w['config']}]", tile=TILE,
               timing="inputs (activations of a step >> 126 MB L2) larger than L2; no explicit flush; device-resident "
                      f"and end-to-end steps timed in {rounds} alternating segment pair(s)",
               weights="random-init seed 1234 (shipped .pth is a Git-LFS pointer)")
    if w["kind"] == "sample":
        cfg.update(per_gpu_batch=batch, global_batch=batch * world)
    else:
        cfg.update(images=images_job, tiles_per_step_mean=nfe_per_step_job / g, noise_minibatch=8,
                   sharding=("tiles of one image over the ranks, one all_gather_into_tensor per step"
                             if w["shard"] else "images over the ranks, one final gather"))
    line = dict(
        metric=metric, value=img_per_sec, unit="images/s", n_gpus=world, steps=K, warmup=args.warmup,
        ms_per_step=ms_per_step, higher_is_better=True, scaling=scaling, vs_baseline=None, dtype="bf16",
        data="synthetic", config=cfg,
        unet_steps_per_sec=nfe_per_sec, tensor_tflops_whole_step=TOTAL_GFLOP_PER_TILE_NFE * nfe_per_sec / 1e3,
        e2e=dict(value=e2e_img_per_sec, unit="images/s", h2d_bytes_per_step=wl_bytes[0], d2h_bytes_per_step=wl_bytes[1],
                 steps=K, ms_per_step=e2e_ms / K, call=wl_call),
        gpu_launches=int(res_launches),
        roofline=roof, hbm_kernels=hbm,
        kernel_ms_per_step={k: round(v["ms"], 4) for k, v in prof.items()},
        clocks=clocks,
    )
    if exchange is not None:
        exchange["share_of_step"] = exchange["ms_standalone"] / ms_per_step
        line["exchange"] = exchange
    if w["kind"] == "sweep":
        line["sweep"] = sweep
    if eager is not None:
        line["gpu_eager_reference"] = eager
    if not args.no_cpu_baseline:
        cb = cpu_baseline(w, ccs, args.cpu_steps, 1)
        line["cpu_baseline"] = dict(value=cb["value"], unit="images/s", cores=cb["cores"], kind="port",
                                    sample=cb["sample"], derivation=cb["derivation"])
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
