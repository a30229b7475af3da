"""One process driving two GPUs (needs >= 2 devices; skipped otherwise): a model that lives on cuda:1 while the current
device is cuda:0 -- the library keeps per-device kernel attributes and every call runs under the tensor's device."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import srgd_oracle as O  # noqa: E402  (checker only: deterministic weights)
import model as M  # noqa: E402


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_the_second_device_while_the_first_is_current():
    spec = O.UnetSpec(dim=64)
    sd = O.make_state_dict(spec, 22, init="torch")

    def build(dev):
        unet = M.ConditionalSRUnet(dim=64, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
        d = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=64, num_sample_steps=250)
        d.load_state_dict(sd, strict=True)
        d = d.eval().to(dev)
        d.progress = False
        return d

    torch.cuda.set_device(0)
    d0, d1 = build("cuda:0"), build("cuda:1")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 3, 64, 128, generator=g)
    cond = torch.rand(3, 3, 64, 128, generator=g) * 2 - 1
    noise = torch.randn(3, 3, 64, 128, generator=g)
    outs = []
    for d, dev in ((d0, "cuda:0"), (d1, "cuda:1"), (d0, "cuda:0"), (d1, "cuda:1")):      # interleaved on purpose
        assert torch.cuda.current_device() == 0
        img, x0 = d.p_sample(x.to(dev), torch.tensor(0.6), cond.to(dev), torch.tensor([1], device=dev), 1.0, 2.0,
                             torch.tensor(0.596), noise=noise.to(dev))
        assert img.device == torch.device(dev)
        outs.append(img.cpu())
    assert torch.equal(outs[0], outs[2]) and torch.equal(outs[1], outs[3])
    assert torch.equal(outs[0], outs[1])                     # same kernels, same inputs: bit-identical across devices
    # a whole tiled run on the second device, then moving a model between devices re-packs its weights
    c01 = torch.rand(1, 3, 200, 264, generator=g)
    torch.manual_seed(3)
    a = d1.tiled_sample(batch_size=4, condition_x=c01.to("cuda:1"), class_label=torch.tensor([0], device="cuda:1"),
                        num_sample_steps=4)
    d0.rng_device = d1.rng_device = "cpu"
    torch.manual_seed(3)
    b1 = d1.tiled_sample(batch_size=4, condition_x=c01.to("cuda:1"), class_label=torch.tensor([0], device="cuda:1"),
                         num_sample_steps=4).cpu()
    moved = d1.to("cuda:0")
    torch.manual_seed(3)
    b0 = moved.tiled_sample(batch_size=4, condition_x=c01.to("cuda:0"), class_label=torch.tensor([0], device="cuda:0"),
                            num_sample_steps=4).cpu()
    assert a.device == torch.device("cuda:1") and torch.isfinite(a).all()
    assert torch.equal(b0, b1)
