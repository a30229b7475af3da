"""Command-line super-resolution with the srgd_b200 CUDA path.

Same flags, defaults, file handling and output naming as the reference CLI (reference:
inference.py:21-44 flags, 47-56 seeding, 59-98 per-image pipeline, 108-142 directory loop), so
`python inference.py -c <yaml> -m <ckpt> --class_cond_scale F --test_label I --seed I --input_dir D
--output_dir D` is a drop-in.  Differences: the model runs on an sm_100 GPU only (explicit error
otherwise) and `--world_size/--rank` (or torchrun's env) shard the sorted file list across GPUs,
the multi-GPU analogue of the reference's manual --start_index/--end_index split.
"""
import glob
import logging
import os
import random
from argparse import ArgumentParser

import numpy as np
import torch
from PIL import Image

from config import load_config
from model import get_model

logger = logging.getLogger("srgd_b200")


def parse_args(argv=None):
    ap = ArgumentParser(description="Real-SRGD x4 super-resolution (B200-native sampling path)")
    ap.add_argument('-c', '--conf', required=True, help='Path to config file')
    ap.add_argument('-m', '--ckpt_path', type=str, required=True)
    ap.add_argument('--input_dir', type=str, required=True)
    ap.add_argument('--output_dir', type=str, required=True)
    for name, typ, default in (('batch_size', int, 8), ('num_sample_steps', int, 250), ('interpolation', str, 'bicubic'),
                               ('cond_scale', float, 1.0), ('class_cond_scale', float, 1.0),
                               ('guidance_start_steps', int, 0), ('class_guidance_start_steps', int, 0),
                               ('generation_start_steps', int, 0), ('start_index', int, 0), ('end_index', int, None),
                               ('test_label', int, None), ('seed', int, 71), ('backend', str, 'ddp')):
        ap.add_argument('--' + name, type=typ, default=default)
    ap.add_argument('--no_amp', dest='amp', action='store_false')
    ap.add_argument('--no_dpmpp_solver', dest='use_dpmpp_solver', action='store_false')
    # multi-GPU sharding of the file list (defaults come from torchrun's environment)
    ap.add_argument('--world_size', type=int, default=int(os.environ.get('WORLD_SIZE', 1)))
    ap.add_argument('--rank', type=int, default=int(os.environ.get('RANK', 0)))
    # throughput extension: consecutive files of equal size are sampled together (their tiles share denoiser batches);
    # every image still sees the noise stream of its own reseeded run, so the outputs do not change
    ap.add_argument('--images_per_batch', type=int, default=1)
    # large images under torchrun: every rank works on EVERY image and denoises a contiguous range of its tiles, with
    # one all-gather per step (tiled_sample's exact mode: the image does not depend on the number of GPUs)
    ap.add_argument('--shard_tiles', action='store_true')
    return ap.parse_args(argv)


def seed_everything(seed):
    """Per-image reseed of every generator the sampling loop draws from (inference.py:47-56)."""
    random.seed(seed)
    os.environ['PYTHONHASHSEED'] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


def _to_unit_tensor(image: Image.Image) -> torch.Tensor:
    """PIL RGB -> float [1,3,H,W] in [0,1] (what torchvision's ToTensor does for uint8 images)."""
    arr = np.array(image, dtype=np.uint8)          # a writable copy (torch warns about read-only arrays)
    return torch.from_numpy(arr).permute(2, 0, 1).unsqueeze(0).float().div(255.)


def _to_image(t: torch.Tensor) -> Image.Image:
    """float [3,H,W] in [0,1] -> PIL RGB with ToPILImage's mul(255).byte() truncation (inference.py:94)."""
    arr = t.detach().mul(255).byte().permute(1, 2, 0).cpu().numpy()
    return Image.fromarray(arr, mode='RGB')


def _prepare(images, scale=4):
    """PIL bicubic pre-upscale + ToTensor of equally sized images (inference.py:66-74) -> float [N,3,4h,4w] in [0,1],
    in pinned host memory when CUDA is present (an asynchronous H2D copy then overlaps the previous image's work)."""
    width, height = images[0].size
    assert all(im.size == (width, height) for im in images)
    # the reference maps both 'bicubic' and 'lanczos' to bicubic (inference.py:66-69)
    cond = torch.cat([_to_unit_tensor(im.resize((width * scale, height * scale), resample=Image.BICUBIC))
                      for im in images])
    if torch.cuda.is_available():
        try:
            cond = cond.pin_memory()
        except RuntimeError:
            pass
    return cond


def _sample(condition_x, sr_model, batch_size, test_label, cond_scale, guidance_start_steps, class_cond_scale,
            class_guidance_start_steps, generation_start_steps, num_sample_steps, enable_amp, seed, shard_tiles=False):
    """seed + tiled_sample (inference.py:77-92); returns the device result as uint8 [N,H,W,3] with ToPILImage's
    mul(255).byte() truncation (inference.py:94) -- a quarter of the float bytes to bring back."""
    condition_x = condition_x.to(sr_model.device, non_blocking=True)
    label = None if test_label is None else torch.tensor([test_label], dtype=torch.long, device=sr_model.device)
    seed_everything(seed)
    kw = dict(shard_tiles=True) if shard_tiles else {}
    with torch.inference_mode():
        output = sr_model.tiled_sample(batch_size=batch_size, condition_x=condition_x, class_label=label,
                                       cond_scale=cond_scale, guidance_start_steps=guidance_start_steps,
                                       class_cond_scale=class_cond_scale,
                                       class_guidance_start_steps=class_guidance_start_steps,
                                       generation_start_steps=generation_start_steps,
                                       num_sample_steps=num_sample_steps, amp=enable_amp, **kw)
        return output.detach().mul(255).byte().permute(0, 2, 3, 1).contiguous()


def sr_target_images(images, sr_model, scale=4, batch_size=8, test_label=2, cond_scale=1.0, guidance_start_steps=0,
                     class_cond_scale=1.0, class_guidance_start_steps=0, generation_start_steps=0,
                     num_sample_steps=250, enable_amp=False, interpolation='bicubic', seed=71, shard_tiles=False):
    """The per-image pipeline of the reference (inference.py:59-98) for a list of equally sized PIL images."""
    width, height = images[0].size
    u8 = _sample(_prepare(images, scale), sr_model, batch_size, test_label, cond_scale, guidance_start_steps,
                 class_cond_scale, class_guidance_start_steps, generation_start_steps, num_sample_steps, enable_amp,
                 seed, shard_tiles)
    outs = [Image.fromarray(a, mode='RGB') for a in u8.cpu().numpy()]
    assert all(o.size == (width * 4, height * 4) for o in outs)
    return outs


def sr_target_image(image, sr_model, scale=4, batch_size=8, test_label=2, cond_scale=1.0, guidance_start_steps=0,
                    class_cond_scale=1.0, class_guidance_start_steps=0, generation_start_steps=0,
                    num_sample_steps=250, enable_amp=False, interpolation='bicubic', seed=71):
    """Same name, arguments and defaults as the reference's per-image function (inference.py:59-98)."""
    return sr_target_images([image], sr_model, scale=scale, batch_size=batch_size, test_label=test_label,
                            cond_scale=cond_scale, guidance_start_steps=guidance_start_steps,
                            class_cond_scale=class_cond_scale, class_guidance_start_steps=class_guidance_start_steps,
                            generation_start_steps=generation_start_steps, num_sample_steps=num_sample_steps,
                            enable_amp=enable_amp, interpolation=interpolation, seed=seed)[0]


def try_open_image(image_path):
    try:
        return Image.open(image_path).convert('RGB')
    except (IOError, SyntaxError):
        return None


def _work_groups(files, output_dir, images_per_batch):
    """The reference's directory walk (inference.py:120-142: sorted files, skip existing outputs, skip unreadable
    files) as a generator of groups [(image, save_path), ...] of consecutive equally sized inputs."""
    pending = []
    for path in files:
        save_path = os.path.join(output_dir, os.path.basename(path).replace('.png', '_out.png'))
        if os.path.exists(save_path):           # doubles as resume-after-crash (inference.py:126)
            print('skip')
            continue
        image = try_open_image(path)
        if image is None:
            print('Invalid image or unable to open image:', path)
            continue
        if pending and (image.size != pending[0][0].size or len(pending) >= images_per_batch):
            yield pending
            pending = []
        pending.append((image, save_path))
    if pending:
        yield pending


def batch_sr_target_images(input_dir, output_dir, sr_model, scale=4, batch_size=8, test_label=2, cond_scale=1.0,
                           guidance_start_steps=0, class_cond_scale=1.0, class_guidance_start_steps=0,
                           generation_start_steps=0, num_sample_steps=250, start_index=0, end_index=None,
                           enable_amp=False, interpolation='bicubic', seed=71, world_size=1, rank=0,
                           images_per_batch=1, shard_tiles=False):
    """The directory loop (inference.py:108-142) as a three-stage pipeline, so the GPU never waits for PIL:

      loader thread : open + RGB convert + bicubic x4 + ToTensor into pinned memory, one group ahead of the sampler
      main thread   : H2D (async from pinned), reseed, tiled_sample, mul(255).byte() on the device, D2H of the uint8
                      result into a pinned buffer on a side stream
      saver thread  : waits for that copy's event, PNG-encodes and writes <name>_out.png

    The order of sampling, the per-image reseed and the pixels are those of the sequential loop."""
    import queue
    import threading
    print(f"save images at: {output_dir}")
    os.makedirs(output_dir, exist_ok=True)
    files = sorted(glob.glob(f"{input_dir}/*"))[start_index:end_index]
    if not shard_tiles:
        files = files[rank::world_size]         # images over ranks; with shard_tiles every rank works on every image
    use_cuda = sr_model.device.type == 'cuda'
    ready: "queue.Queue" = queue.Queue(maxsize=2)
    to_save: "queue.Queue" = queue.Queue(maxsize=4)
    failure = []

    def loader():
        try:
            for group in _work_groups(files, output_dir, images_per_batch):
                ready.put((group, _prepare([im for im, _ in group], scale)))
        except BaseException as e:              # surfaced by the main thread
            failure.append(e)
        finally:
            ready.put(None)

    def saver():
        while True:
            item = to_save.get()
            if item is None:
                return
            try:
                host, event, group = item
                if event is not None:
                    event.synchronize()
                for arr, (im, save_path) in zip(host.numpy(), group):
                    out = Image.fromarray(arr, mode='RGB')
                    assert out.size == (im.size[0] * 4, im.size[1] * 4)
                    out.save(save_path)
            except BaseException as e:
                failure.append(e)

    threads = [threading.Thread(target=loader, daemon=True), threading.Thread(target=saver, daemon=True)]
    for t in threads:
        t.start()
    copy_stream = torch.cuda.Stream(device=sr_model.device) if use_cuda else None
    try:
        while not failure:
            item = ready.get()
            if item is None:
                break
            group, cond = item
            u8 = _sample(cond, sr_model, batch_size, test_label, cond_scale, guidance_start_steps, class_cond_scale,
                         class_guidance_start_steps, generation_start_steps, num_sample_steps, enable_amp, seed,
                         shard_tiles)
            if shard_tiles and rank != 0:
                continue                         # every rank holds the image; rank 0 writes it
            if use_cuda:
                host = torch.empty(u8.shape, dtype=torch.uint8).pin_memory()
                done = torch.cuda.Event()
                copy_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(copy_stream):
                    host.copy_(u8, non_blocking=True)
                    done.record(copy_stream)
                u8.record_stream(copy_stream)
                to_save.put((host, done, group))
            else:
                to_save.put((u8, None, group))
    finally:
        to_save.put(None)
        threads[1].join()
    if failure:
        raise failure[0]


def main(argv=None):
    args = parse_args(argv)
    logging.basicConfig(level=logging.INFO)
    conf = load_config(args.conf)
    conf.num_sample_steps = args.num_sample_steps
    conf.ckpt_path = args.ckpt_path
    if not torch.cuda.is_available():
        raise RuntimeError("srgd_b200 needs an sm_100 (B200) CUDA device; there is no CPU fallback. "
                           "Use the reference implementation for CPU inference.")
    local_rank = int(os.environ.get('LOCAL_RANK', args.rank % max(1, torch.cuda.device_count())))
    torch.cuda.set_device(local_rank)
    shard = bool(args.shard_tiles) and args.world_size > 1
    if shard:                                   # one large image at a time, its tiles split over the GPUs of the box
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    ema_model = get_model(conf, logger)
    sr_model = ema_model.module.eval().to(torch.device('cuda', local_rank))
    print(args)
    batch_sr_target_images(args.input_dir, args.output_dir, sr_model, scale=4, batch_size=args.batch_size,
                           test_label=args.test_label, cond_scale=args.cond_scale,
                           guidance_start_steps=args.guidance_start_steps, class_cond_scale=args.class_cond_scale,
                           class_guidance_start_steps=args.class_guidance_start_steps,
                           generation_start_steps=args.generation_start_steps,
                           num_sample_steps=args.num_sample_steps, start_index=args.start_index,
                           end_index=args.end_index, enable_amp=args.amp, interpolation=args.interpolation,
                           seed=args.seed, world_size=args.world_size, rank=args.rank,
                           images_per_batch=max(1, args.images_per_batch), shard_tiles=shard)


if __name__ == '__main__':
    main()
