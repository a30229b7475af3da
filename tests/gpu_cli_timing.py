"""CLI-level measurements (not a pytest test; run on a B200):

    python tests/gpu_cli_timing.py [--images 24] [--steps 250]

  f-3  start-up: get_model() + .to(cuda) + first forward, (a) first start of a checkpoint (torch.load of the 550 MB fp32
       state dict, load_state_dict, repack, ingest-cache write) and (b) a later start (ingest cache hit: no torch.load,
       no repack), each in a FRESH python process so that nothing is warm except the OS page cache;
  f-2  the directory loop: wall time of inference.main over N images against N x the sampling time of one image
       (the pipeline hides decode + bicubic + PNG encode + copies behind the sampling of the neighbouring images).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STARTUP = r'''
import sys, time, logging, json
t0 = time.perf_counter()
import torch
sys.path.insert(0, %(root)r)
import config, model as M
t_import = time.perf_counter() - t0
torch.cuda.init(); torch.zeros(1, device="cuda"); torch.cuda.synchronize()
t1 = time.perf_counter()
conf = config.load_config(%(yaml)r); conf.ckpt_path = %(ckpt)r
ema = M.get_model(conf, logging.getLogger("t"))
t2 = time.perf_counter()
sr = ema.module.eval().to("cuda")
torch.cuda.synchronize()
t3 = time.perf_counter()
x = torch.zeros(1, 3, 256, 256, device="cuda")
with torch.inference_mode():
    sr.p_sample(x, torch.tensor(0.5), x, torch.tensor([0], device="cuda"), 1.0, 1.0, torch.tensor(0.496))
torch.cuda.synchronize()
t4 = time.perf_counter()
print(json.dumps(dict(import_torch_s=t_import, get_model_s=t2 - t1, to_cuda_s=t3 - t2, first_step_incl_pack_s=t4 - t3,
                      ready_s=t4 - t1, cached=sr.model._cached_pack is not None)))
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=40)
    ap.add_argument("--steps", type=int, default=250)
    a = ap.parse_args()
    import numpy as np
    import torch
    from PIL import Image
    from srgd_b200 import arch, weights
    tmp = tempfile.mkdtemp(prefix="srgd_cli_")
    yaml_path = os.path.join(tmp, "c.yaml")
    open(yaml_path, "w").write("model: conditional_continuous\nnoise_schedule: linear\nunet_dim: 128\nimage_size: 256\n"
                               "num_sample_steps: 250\nlearned_sinusoidal_cond: true\nlearned_sinusoidal_dim: 32\n")
    ckpt = os.path.join(tmp, "w.pth")
    torch.save({"ema_model": arch.seeded_state_dict(arch.UnetSpec(), 1234, init="torch")}, ckpt)
    out = dict(checkpoint_bytes=os.path.getsize(ckpt))
    # ---- f-3: start-up in fresh processes ----
    code = STARTUP % dict(root=ROOT, yaml=yaml_path, ckpt=ckpt)
    runs = []
    for k in range(3):                       # 0: no cache (writes it), 1 and 2: cache hit
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stderr[-2000:])
            raise SystemExit(1)
        runs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    out["startup_first"] = runs[0]
    out["startup_cached"] = runs[2]
    out["ingest_cache_bytes"] = os.path.getsize(weights.pack_cache_path(ckpt))
    # ---- f-2: directory loop vs N x sampling ----
    import inference
    in_dir, out_dir = os.path.join(tmp, "in"), os.path.join(tmp, "out")
    os.makedirs(in_dir)
    for i in range(a.images):
        lr = np.random.RandomState(71 + i).randint(0, 256, (64, 64, 3), dtype=np.uint8)
        Image.fromarray(lr, mode="RGB").save(os.path.join(in_dir, f"im{i:03d}.png"))
    argv = ["-c", yaml_path, "-m", ckpt, "--input_dir", in_dir, "--output_dir", out_dir, "--num_sample_steps",
            str(a.steps), "--test_label", "0", "--seed", "71"]
    inference.main(argv[:7] + [os.path.join(tmp, "warm")] + argv[8:9] + ["4"] + argv[10:])      # warm-up: 4 steps
    t0 = time.perf_counter()
    inference.main(argv)
    wall = time.perf_counter() - t0
    # sampling alone: the same tiled_sample call on a resident condition, N times
    import logging
    import config
    import model as M
    conf = config.load_config(yaml_path)
    conf.num_sample_steps, conf.ckpt_path = a.steps, ckpt
    sr = M.get_model(conf, logging.getLogger("t")).module.eval().to("cuda")
    sr.progress = False
    cond = torch.rand(1, 3, 256, 256, device="cuda")
    label = torch.tensor([0], device="cuda")
    with torch.inference_mode():
        sr.tiled_sample(batch_size=8, condition_x=cond, class_label=label, num_sample_steps=4)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.images):
            inference.seed_everything(71)
            sr.tiled_sample(batch_size=8, condition_x=cond, class_label=label, num_sample_steps=a.steps)
        torch.cuda.synchronize()
        sampling = time.perf_counter() - t0
    out["directory_loop"] = dict(images=a.images, steps=a.steps, cli_wall_s=wall, n_x_sampling_s=sampling,
                                 overhead_pct=100.0 * (wall / sampling - 1.0),
                                 note="cli_wall_s includes get_model + first-use packing of that process")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
