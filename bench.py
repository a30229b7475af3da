#!/usr/bin/env python
"""Benchmark of the Real-SRGD sampling hot path on B200 (see DESIGN.md "Measurement").

A *step* is one pass of the hot path over one batch: `p_sample` = conditional U-Net denoise
(2x batch under CFG) + fused posterior update, on `--batch` 256x256 tiles (= 64x64 LR images), with
the actual 250-step linear-logSNR schedule times.  Default workload = BASELINE.json configs[1]:
batch 16, label 0, class_cond_scale 1.0, bf16 kernels, K = 250 steps = one full sampling schedule.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--class_cond_scale S]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference      # CPU arm: the oracle port on the host cores

Prints ONE JSON line (rank 0).  `value` = SR images/sec (64x64 LR -> 256x256, 250 steps) over all
GPUs with inputs resident in HBM; `e2e` = the same through the reference-facing API with pinned
host buffers copied in/out every step (double-buffered on a second stream, as many steps as `value`); `roofline` = the tcgen05 conv kernel's achieved TFLOP/s
(CUDA events around every conv launch, srgd_profile_*) against the measured bf16 peak;
`cpu_baseline` = the oracle port of the reference on the host cores (bounded sample).
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
CONV_GFLOP_PER_TILE_NFE = 789.35      # conv3x3 705.45 + conv1x1 78.97 + conv7x7 4.93
TOTAL_GFLOP_PER_TILE_NFE = 793.8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=SAMPLE_STEPS)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--class_cond_scale", type=float, default=1.0)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--cpu_steps", type=int, default=2, help="timed CPU-baseline steps (B=1 tile each)")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--dump_launches", type=str, default=None,
                    help="write the per-launch CUDA-event table of one profiled step (kind, ms, TFLOP/s, GB/s) here")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16_tflops=p.get("bf16_tflops_sustained", p.get("bf16_tflops")), hbm_gbs=p.get("hbm_gbs"),
                    source="MEASURED_PEAKS.json (sustained bf16)")
    return dict(bf16_tflops=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


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
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.startswith("Active")})
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.rows))


def synth_inputs(batch, seed=71):
    """Synthetic LR images exactly as BASELINE.md §3: RandomState(71+idx) uint8 64x64x3, PIL bicubic x4."""
    import numpy as np
    from PIL import Image
    conds = []
    for idx in range(batch):
        lr = np.random.RandomState(seed + idx).randint(0, 256, (64, 64, 3), dtype=np.uint8)
        hr = Image.fromarray(lr, mode="RGB").resize((TILE, TILE), resample=Image.BICUBIC)
        conds.append(torch.from_numpy(np.array(hr, dtype=np.uint8)).permute(2, 0, 1).float().div(255.))
    return torch.stack(conds)           # [B,3,256,256] in [0,1]


def cpu_reference_arm(args, n_steps, warm):
    """The reference's own algorithm on the host cores: the oracle port (oracle/srgd_oracle.py, pinned
    to the unmodified reference by tests/golden) -- the reference itself is Python and cannot travel."""
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
    with torch.inference_mode():
        for i in range(warm + n_steps):
            t0 = time.perf_counter()
            x, _ = O.p_sample(sd, spec, x, steps[i], cond, label, 1.0, args.class_cond_scale, steps[i + 1],
                              generator=g)
            if i >= warm:
                times.append(time.perf_counter() - t0)
    nfe_per_step = 2 if args.class_cond_scale != 1.0 else 1
    s_per_step = sum(times) / len(times)
    img_per_s = 1.0 / (s_per_step * SAMPLE_STEPS)          # one tile (= one 64x64-LR image) per step
    return dict(value=img_per_s, unit="images/s", cores=threads, kind="port",
                sample=f"{n_steps} timed p_sample steps (+{warm} warm-up) of ONE 256x256 tile, fp32, "
                       f"class_cond_scale {args.class_cond_scale}, extrapolated x{SAMPLE_STEPS} steps",
                s_per_step=s_per_step, unet_steps_per_sec=nfe_per_step / s_per_step)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    workload = (f"sample(): {args.batch} synthetic 64x64-LR tiles (256x256) per GPU, label 0, "
                f"class_cond_scale {args.class_cond_scale}, {SAMPLE_STEPS}-step linear-logSNR schedule, dim128 U-Net")
    metric = "SR images/sec (64x64 LR -> 256x256, 250 sampling steps)"

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference_arm(args, max(1, min(args.steps, 3)), 1)
        line = dict(metric=metric, value=cb["value"], unit="images/s", n_gpus=args.gpus, steps=args.steps,
                    warmup=args.warmup, ms_per_step=cb["s_per_step"] * 1e3, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                    config=dict(workload=workload, note="CPU arm: each step is a bounded sample (1 tile)"),
                    cpu_baseline=dict(value=cb["value"], unit="images/s", cores=cb["cores"], kind="port",
                                      sample=cb["sample"]),
                    unet_steps_per_sec=cb["unet_steps_per_sec"],
                    e2e=dict(value=cb["value"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    import torch.distributed as dist
    from oracle import srgd_oracle as O           # only for the deterministic random-init weights + CPU arm
    import model as M
    from srgd_b200 import _lib, sharding

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec = O.UnetSpec()
    unet = M.ConditionalSRUnet(dim=128, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=TILE, num_sample_steps=SAMPLE_STEPS)
    diff.load_state_dict(O.make_state_dict(spec, 1234), strict=True)
    diff = diff.eval().to(dev)
    diff.progress = False
    lib = _lib.load()

    B = args.batch
    cond01 = synth_inputs(B, seed=71 + rank * B)
    cond = (cond01 * 2 - 1).to(dev)
    label = torch.tensor([0], device=dev)
    steps = torch.linspace(1., 0., SAMPLE_STEPS + 1)
    ccs = args.class_cond_scale
    nfe_per_step = B * (2 if ccs != 1.0 else 1)

    def run_steps(img, first, count):
        for k in range(count):
            i = (first + k) % SAMPLE_STEPS
            img, _ = diff.p_sample(img, steps[i], cond, label, 1.0, ccs, steps[i + 1])
        return img

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.manual_seed(71 + rank)
    img = torch.randn(B, 3, TILE, TILE, device=dev)
    with torch.inference_mode():
        img = run_steps(img, 0, args.warmup)
        if world > 1:                              # warm-up covers every op of the timed region, the gather included
            sharding.gather_rows(diff._finalize(img), [B] * world, dst=0)   # (first use sets up NCCL's P2P channels)
        barrier()
        launches_before = 0
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        img = run_steps(img, args.warmup, args.steps)
        out = diff._finalize(img)
        if world > 1:                              # the only collective: final gather of finished images
            gathered = sharding.gather_rows(out, [B] * world, dst=0)
        e1.record()
        barrier()
        elapsed_ms = e0.elapsed_time(e1)
        clocks = sampler.summary() if sampler else None
        step_launches = diff.last_step_launches
        if world > 1:
            t = torch.tensor([elapsed_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            elapsed_ms = float(t)

        # ---- end-to-end through the public API with HOST buffers (pinned), every step ----
        # Every step's x / cond come from pinned host memory and its img_next goes back to pinned host memory inside
        # the timed region.  The copies run on a second stream, double-buffered: step k+1's inputs travel while step k
        # computes and step k's result travels while step k+1 computes (what a serving loop around p_sample does);
        # the host waits for (= can read) the result of step k-2 before it issues step k+1.
        e2e_steps = max(3, args.steps)                    # as long as the device-resident run: same clock / power state
        x_host = torch.randn(B, 3, TILE, TILE).pin_memory()
        c_host = (cond01 * 2 - 1).pin_memory()
        r_host = [torch.empty(B, 3, TILE, TILE).pin_memory() for _ in range(3)]
        xd = [torch.empty(B, 3, TILE, TILE, device=dev) for _ in range(2)]
        cd = [torch.empty(B, 3, TILE, TILE, device=dev) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream()
        h2d_done = [torch.cuda.Event() for _ in range(2)]
        comp_done = [torch.cuda.Event() for _ in range(2)]
        d2h_done = [torch.cuda.Event() for _ in range(3)]

        def issue_h2d(slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(comp_done[slot])          # the step that last read this slot has finished
                xd[slot].copy_(x_host, non_blocking=True)
                cd[slot].copy_(c_host, non_blocking=True)
                h2d_done[slot].record(copy_stream)

        # One process per GPU: pipelined copies were measured erratic with several ranks on one box (15.1-18.3 ms per
        # step at 2 GPUs against 14.9 device-resident, cause not found), so ranks > 1 keep the plain sequence
        # copy in -> step -> copy out -> sync, which costs the ~0.8 ms of PCIe time per step but is stable.
        pipelined = (world == 1) if not os.environ.get("SRGD_E2E_PIPELINED") else os.environ["SRGD_E2E_PIPELINED"] == "1"
        barrier()
        for ev in comp_done:
            ev.record(cur)
        e0.record()
        if not pipelined:
            for k in range(e2e_steps):
                i = (args.warmup + k) % SAMPLE_STEPS
                xd[0].copy_(x_host, non_blocking=True)
                cd[0].copy_(c_host, non_blocking=True)
                o, _ = diff.p_sample(xd[0], steps[i], cd[0], label, 1.0, ccs, steps[i + 1])
                r_host[0].copy_(o, non_blocking=True)
                cur.synchronize()
        else:
            issue_h2d(0)
        for k in range(e2e_steps if pipelined else 0):
            slot = k & 1
            i = (args.warmup + k) % SAMPLE_STEPS
            cur.wait_event(h2d_done[slot])
            o, _ = diff.p_sample(xd[slot], steps[i], cd[slot], label, 1.0, ccs, steps[i + 1])
            comp_done[slot].record(cur)
            o.record_stream(copy_stream)
            if k + 1 < e2e_steps:
                issue_h2d(slot ^ 1)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(comp_done[slot])
                r_host[k % 3].copy_(o, non_blocking=True)
                d2h_done[k % 3].record(copy_stream)
            if k >= 2:
                d2h_done[(k - 2) % 3].synchronize()              # result of step k-2 is in host memory (its buffer is
                                                                 # rewritten by step k+1); two steps stay queued
        for ev in d2h_done:
            ev.synchronize()
        cur.wait_stream(copy_stream)
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        if os.environ.get("SRGD_BENCH_DEBUG"):
            print(f"\n[rank {rank}] e2e {e2e_ms / e2e_steps:.3f} ms/step, device-resident {elapsed_ms / args.steps:.3f}\n",
                  file=sys.stderr, flush=True)
        if world > 1:
            t = torch.tensor([e2e_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t)

        # ---- per-kernel-kind device time (CUDA events around every launch), rank 0 only ----
        prof = None
        if rank == 0:
            prof_steps = 3
            torch.cuda.synchronize()
            _lib.check(lib.srgd_profile_begin())
            img = run_steps(img, 100, prof_steps)
            _lib.check(lib.srgd_profile_end())
            prof = {k: dict(v, ms=v["ms"] / prof_steps, launches=v["launches"] // prof_steps,
                            flops=v["flops"] / prof_steps, bytes=v["bytes"] / prof_steps)
                    for k, v in _lib.profile_report().items()}
            if args.dump_launches:
                recs = _lib.profile_records()
                per = len(recs) // prof_steps
                with open(args.dump_launches, "w") as f:
                    f.write("# one warm step of the bench workload, CUDA events around every library launch\n")
                    f.write("idx kind ms tflops gbs\n")
                    for i, (kind, ms, fl, by) in enumerate(recs[-per:]):
                        f.write(f"{i} {kind} {ms:.4f} {fl / (ms * 1e-3) / 1e12 if ms > 0 else 0:.1f} "
                                f"{by / (ms * 1e-3) / 1e9 if ms > 0 else 0:.0f}\n")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = elapsed_ms / args.steps
    nfe_per_sec = world * nfe_per_step / (ms_per_step * 1e-3)
    tiles_per_image = SAMPLE_STEPS * (2 if ccs != 1.0 else 1)
    img_per_sec = nfe_per_sec / tiles_per_image
    e2e_img_per_sec = world * nfe_per_step / (e2e_ms / e2e_steps * 1e-3) / tiles_per_image
    pk = peaks()
    conv = prof["conv_igemm"]
    # FLOPs actually executed by the conv launches (2*M*N*K credited per launch by the library); the 1x1 convs of
    # the fused LinearAttention blocks are not in this kernel any more
    conv_tflops = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
    # DRAM traffic of the same kernel from the committed `ncu --set full` capture of this workload (profiles/):
    # dram__bytes_read.sum + dram__bytes_write.sum averaged over the conv launches of one step
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_full_conv_latest.json")
    if os.path.exists(tpath) and args.batch == 16 and ccs == 1.0:
        try:
            kk = json.load(open(tpath))
            traffic = sum(o["dram_bytes"] for o in kk) / max(len(kk), 1)
        except Exception:
            traffic = None
    roof = dict(bound="tensor", kernel="conv_igemm_kernel (tcgen05 implicit GEMM, all conv3x3/1x1/7x7 launches of a step)",
                achieved=conv_tflops, peak=pk["bf16_tflops"], unit="TFLOP/s", frac=conv_tflops / pk["bf16_tflops"],
                traffic=traffic, traffic_unit="DRAM bytes per conv launch (ncu dram__bytes_read+write, profiles/ncu_full_conv_latest.json)",
                achieved_per_launch_flops=conv["flops"] / max(conv["launches"], 1), peak_source=pk["source"],
                algorithmic=f"{conv['flops'] / 1e9 / nfe_per_step:.2f} GFLOP per tile-NFE (2*M*N*K of every conv launch) "
                            f"x {nfe_per_step} tile-NFE per step",
                launches_per_step=conv["launches"], kernel_ms_per_step=conv["ms"],
                share_of_step=conv["ms"] / ms_per_step)
    hbm = {}
    for k in ("gn_apply", "sampler_step", "linear_attention", "norm_misc"):
        if prof[k]["ms"] > 0:
            gbs = prof[k]["bytes"] / (prof[k]["ms"] * 1e-3) / 1e9
            hbm[k] = dict(ms_per_step=prof[k]["ms"], gb_per_s=gbs, frac_of_hbm_peak=gbs / pk["hbm_gbs"],
                          launches_per_step=prof[k]["launches"])
    line = dict(
        metric=metric, value=img_per_sec, unit="images/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
        data="synthetic",
        config=dict(workload=workload, per_gpu_batch=B, global_batch=B * world, tile=TILE,
                    timing="inputs (activations of a step >> 126 MB L2) larger than L2; no explicit flush",
                    weights="random-init seed 1234 (shipped .pth is a Git-LFS pointer)"),
        unet_steps_per_sec=nfe_per_sec, tensor_tflops_whole_step=TOTAL_GFLOP_PER_TILE_NFE * nfe_per_sec / 1e3,
        e2e=dict(value=e2e_img_per_sec, unit="images/s", h2d_bytes_per_step=2 * B * 3 * TILE * TILE * 4,
                 d2h_bytes_per_step=B * 3 * TILE * TILE * 4, steps=e2e_steps,
                 call="ConditionalContinuousTimeGaussianDiffusionSR.p_sample with pinned host x/cond in, img_next out; "
                      + ("copies double-buffered on a second stream" if pipelined else "copy in, step, copy out, sync")),
        gpu_launches=int(step_launches) * args.steps + 1,
        roofline=roof, hbm_kernels=hbm,
        kernel_ms_per_step={k: round(v["ms"], 4) for k, v in prof.items()},
        clocks=clocks,
    )
    if not args.no_cpu_baseline:
        cb = cpu_reference_arm(args, args.cpu_steps, 1)
        line["cpu_baseline"] = dict(value=cb["value"], unit="images/s", cores=cb["cores"], kind="port", sample=cb["sample"])
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
