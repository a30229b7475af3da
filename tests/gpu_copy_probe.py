"""Probe (not a test): pinned host<->device copy rates per rank, alone and under a running p_sample step (torchrun)."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import srgd_oracle as O  # noqa: E402
import model as M  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
spec = O.UnetSpec()
unet = M.ConditionalSRUnet(dim=128, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=256, num_sample_steps=250)
diff.load_state_dict(O.make_state_dict(spec, 1234), strict=True)
diff = diff.eval().to(dev)
B = 16
h = torch.randn(B, 3, 256, 256).pin_memory()
d = torch.empty(B, 3, 256, 256, device=dev)
cond = torch.rand(B, 3, 256, 256, device=dev)
label = torch.tensor([0], device=dev)
steps = torch.linspace(1., 0., 251)
cs = torch.cuda.Stream()


def copies(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(cs):
        e0.record(cs)
        for _ in range(n):
            d.copy_(h, non_blocking=True)
        e1.record(cs)
    return e0, e1


with torch.inference_mode():
    x = torch.randn(B, 3, 256, 256, device=dev)
    for k in range(3):
        x, _ = diff.p_sample(x, steps[100 + k], cond, label, 1.0, 1.0, steps[101 + k])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = copies(20)
    torch.cuda.synchronize()
    alone = 20 * h.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(10):
        x, _ = diff.p_sample(x, steps[110 + k], cond, label, 1.0, 1.0, steps[111 + k])
    e0, e1 = copies(20)
    torch.cuda.synchronize()
    under = 20 * h.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    print(f"rank {rank}: H2D {alone:.1f} GB/s alone, {under:.1f} GB/s while p_sample runs "
          f"(10 steps + copies took {(time.perf_counter() - t0) * 1e3:.1f} ms)", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
