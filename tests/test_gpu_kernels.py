"""GPU parity tests of the individual CUDA kernels (C-ABI calls) against plain PyTorch fp32 math on
the same bf16-rounded inputs (kernel-level) -- run with `pytest -m gpu` on a B200.

Tolerances: kernels accumulate in fp32 and round the stored result to bf16 once, so the bound is
bf16 rounding of the result (2^-8 relative) plus fp32 summation-order noise: 1e-2 * max|ref| abs.
The sampler kernels are fp32 end to end and must match torch's op-by-op fp32 evaluation exactly.
"""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from srgd_b200 import _lib  # noqa: E402
import gpu_util as G  # noqa: E402
from oracle import srgd_oracle as O  # noqa: E402  (checker only)


@pytest.fixture(scope="module")
def lib():
    l = _lib.load()
    _lib.check(l.srgd_device_check(0), "device_check")
    return l


@pytest.fixture(autouse=True)
def _release_kept_tensors():
    yield
    G.release()


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-6))


# ------------------------------------------------------------------------------------------------
# sampler
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("step,scale,has_noise", [(0, 1.0, True), (0, 3.0, True), (120, 3.0, True), (249, 3.0, False),
                                                  (249, 1.0, False), (60, 1.0, True)])
def test_sampler_step_exact(lib, step, scale, has_noise):
    g = torch.Generator().manual_seed(step)
    n = 3 * 64 * 64 * 2 + 3            # odd tail exercises the scalar path
    x, ec, en, z = (torch.randn(n, generator=g) for _ in range(4))
    steps = torch.linspace(1., 0., 251)
    s = O.step_scalars(steps[step], steps[step + 1])
    eps = en + (ec - en) * scale if scale != 1.0 else ec
    mean, var, x0 = O.posterior_update(x, eps, s)
    ref = mean + var.sqrt() * z if has_noise else mean
    sc = _lib.StepScalars(float(s["alpha"]), float(s["sigma"]), float(s["alpha_next"]), float(s["c"]),
                          float(s["var"].sqrt()) if has_noise else 0.0, scale, 1)
    xd, ecd, end_, zd = (t.cuda() for t in (x, ec, en, z))
    img, x0d = torch.empty_like(xd), torch.empty_like(xd)
    _lib.check(lib.srgd_sampler_step(G.P(xd), G.P(ecd), G.P(end_) if scale != 1.0 else None,
                                     G.P(zd) if has_noise else None, G.P(img), G.P(x0d), n,
                                     C.byref(sc), G.stream()))
    torch.cuda.synchronize()
    # bit-exact against torch's separately rounded fp32 ops
    assert torch.equal(x0d.cpu(), x0), float((x0d.cpu() - x0).abs().max())
    assert torch.equal(img.cpu(), ref), float((img.cpu() - ref).abs().max())


def test_q_sample_and_finalize(lib):
    g = torch.Generator().manual_seed(3)
    x0, z = torch.randn(1000, generator=g), torch.randn(1000, generator=g)
    out = torch.empty(1000, device="cuda")
    x0d, zd = x0.cuda(), z.cuda()          # keep references: ctypes pointers do not own the storage
    _lib.check(lib.srgd_q_sample(G.P(x0d), G.P(zd), G.P(out), 1000, 0.6, 0.8, G.stream()))
    torch.testing.assert_close(out.cpu(), x0 * 0.6 + z * 0.8, rtol=1e-6, atol=1e-6)
    _lib.check(lib.srgd_q_sample(None, G.P(zd), G.P(out), 1000, 0.6, 0.8, G.stream()))
    torch.testing.assert_close(out.cpu(), z * 0.8, rtol=0, atol=0)
    img = (x0 * 2).cuda()
    _lib.check(lib.srgd_finalize_image(G.P(img), G.P(out), 1000, G.stream()))
    torch.testing.assert_close(out.cpu(), ((x0 * 2).clamp(-1, 1) + 1) * 0.5, rtol=0, atol=0)


def test_tile_gather_scatter_renoise(lib):
    """Batched tile orchestration kernels of tiled_sample (model.py:3364-3380, 3392-3396)."""
    g = torch.Generator().manual_seed(3)
    H, W, T = 96, 160, 32
    canvas = torch.randn(1, 3, H, W, generator=g).cuda()
    coords = [(0, 0), (32, 64), (64, 128), (16, 4), (64, 0)]
    tc = _lib.TileCoords()
    tc.n = len(coords)
    for k, (y, x) in enumerate(coords):
        tc.yx[k][0], tc.yx[k][1] = y, x
    tiles = torch.full((len(coords), 3, T, T), float("nan"), device="cuda")
    _lib.check(lib.srgd_gather_tiles(G.P(canvas), G.P(tiles), C.byref(tc), 3, H, W, T, G.stream()))
    ref = torch.cat([canvas[:, :, y:y + T, x:x + T] for y, x in coords], 0)
    assert torch.equal(tiles, ref)
    dst = torch.zeros_like(canvas)
    expect = torch.zeros_like(canvas)
    for k, (y, x) in enumerate(coords):
        expect[:, :, y:y + T, x:x + T] = tiles[k]
    _lib.check(lib.srgd_scatter_tiles(G.P(dst), G.P(tiles), C.byref(tc), 3, H, W, T, G.stream()))
    assert torch.equal(dst, expect)
    noise = torch.randn(1, 3, H, W, generator=g).cuda()
    img = canvas.clone()
    _lib.check(lib.srgd_renoise_outside(G.P(img), G.P(noise), 3, H, W, 16, 80, 32, 128, 0.37, G.stream()))
    want = noise * 0.37
    want[:, :, 16:80, 32:128] = canvas[:, :, 16:80, 32:128]
    assert torch.equal(img, want)
    tc.yx[0][1] = 130                                                       # tile sticks out of the canvas
    assert lib.srgd_gather_tiles(G.P(canvas), G.P(tiles), C.byref(tc), 3, H, W, T, G.stream()) == -1


# ------------------------------------------------------------------------------------------------
# convolutions
# ------------------------------------------------------------------------------------------------
CONV_CASES = [
    # B, H, W, [Cin...], Cout, ksize
    (2, 32, 32, [128], 128, 3),
    (1, 64, 64, [128], 128, 3),
    (2, 16, 16, [256, 128], 256, 3),      # concat of two sources, BN=128/256
    (2, 16, 16, [128], 384, 1),           # qkv-like 1x1
    (3, 8, 8, [512], 1024, 3),            # tile spans 2 samples (TN=2), odd batch -> masked rows
    (1, 40, 24, [128], 128, 3),           # partial tiles in x and y
    (2, 32, 32, [64], 64, 3),             # BN=64 instance
    (1, 32, 32, [1024, 512], 1024, 1),    # long-K 1x1 over a concat
    (1, 24, 40, [128], 128, 3),           # odd number of pixel tiles (channels-as-M pairs them)
    (5, 16, 16, [128], 384, 1),
    (1, 6, 256, [128], 128, 3),           # W % 256 == 0: halo-reuse variant (row-shifted tcgen05 operand descriptors)
    (2, 3, 512, [64, 128], 128, 3),       # halo reuse over a concat, two 256-pixel tiles per row
    (1, 2, 256, [64], 384, 3),            # halo reuse, three 128-channel slabs
    (1, 160, 128, [64], 256, 3),          # 160 tiles on 148 SMs: the 12 tail tiles run as 128-column halves
    (1, 152, 128, [128], 512, 1),         # tail split with two n-tiles per pixel tile (304 tiles, 8 split)
]


@pytest.fixture(params=["auto", "n", "t"])
def conv_variant(request):
    """auto: channels-as-M kernel for 128-wide Cout slabs and for short-K convs (halo-reuse variant where it
    applies); n: force the pixels-as-M kernel; t: channels-as-M without halo reuse."""
    import os
    if request.param != "auto":
        os.environ["SRGD_CONV_VARIANT"] = request.param
    yield request.param
    os.environ.pop("SRGD_CONV_VARIANT", None)


@pytest.mark.parametrize("B,H,W,cins,Cout,ks", CONV_CASES)
@pytest.mark.parametrize("direct", [False, True])
def test_conv_matches_torch(lib, conv_variant, B, H, W, cins, Cout, ks, direct):
    if direct and conv_variant != "auto":
        pytest.skip("direct path has no variants")
    g = torch.Generator().manual_seed(B * 1000 + H + Cout)
    xs = [G.bf16_round(torch.randn(B, c, H, W, generator=g)) for c in cins]
    cin = sum(cins)
    w = G.bf16_round(torch.randn(Cout, cin, ks, ks, generator=g) / math.sqrt(cin * ks * ks))
    bias = torch.randn(Cout, generator=g)
    ref = F.conv2d(torch.cat(xs, 1), w, bias, padding=ks // 2)
    out = torch.zeros(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    d = G.plain_conv_desc([G.nhwc_bf16(x) for x in xs], G.pack_conv_weight(w), B, H, W, Cout, ks, out,
                          bias=bias.cuda())
    G.run_conv(d, direct)
    got = G.to_nchw_f32(out)
    assert rel_err(got, ref) < 1e-2, f"rel err {rel_err(got, ref)}"


# (B, H, W, cins, Cout, ks): launches whose last wave of 128 x 256 tiles is at most half full on 148 SMs, so the K loop
# of those tiles is split across idle SMs when a split-K workspace is passed (conv_igemm.cu WorkItem)
SPLITK_CASES = [
    (1, 32, 32, [1024], 1024, 3),         # 32 tiles -> 4 K parts each (128 work items), the one-tile CLI shape
    (16, 32, 32, [1024], 1024, 3),        # 512 tiles = 3 full waves + 68 tiles split in two (the bench shape)
    (1, 32, 32, [1024, 512], 1024, 3),    # concat: two sources with different k-block counts per phase
    (2, 64, 64, [256], 256, 3),           # 64 tiles x 2 parts, GroupNorm partial records from the main part
    (1, 64, 64, [512], 512, 3),           # 64 tiles (two n-tiles per pixel tile)
    (3, 32, 32, [512], 512, 1),           # short K: 8 k-blocks -> no split (falls back to whole tiles)
]


@pytest.mark.parametrize("B,H,W,cins,Cout,ks", SPLITK_CASES)
def test_conv_splitk_matches_unsplit_and_torch(lib, B, H, W, cins, Cout, ks):
    g = torch.Generator().manual_seed(B * 77 + H + Cout + ks)
    xs = [G.bf16_round(torch.randn(B, c, H, W, generator=g)) for c in cins]
    cin = sum(cins)
    w = G.bf16_round(torch.randn(Cout, cin, ks, ks, generator=g) / math.sqrt(cin * ks * ks))
    bias = torch.randn(Cout, generator=g)
    res = G.bf16_round(torch.randn(B, Cout, H, W, generator=g))
    ref = F.conv2d(torch.cat(xs, 1), w, bias, padding=ks // 2)
    xd, wd, rd = [G.nhwc_bf16(x) for x in xs], G.pack_conv_weight(w), G.nhwc_bf16(res)
    ws = G.splitk_workspace()
    n_rec = lib.srgd_conv_m_tiles(B, H, W)
    outs, parts = [], []
    # unsplit, split, and split again on the SAME workspace (the kernel must have re-armed its flags)
    for sk in (None, ws, ws):
        out = torch.zeros(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
        part = torch.zeros(n_rec, 8, 2, device="cuda")
        d = G.plain_conv_desc(xd, wd, B, H, W, Cout, ks, out, bias=bias.cuda(), gn_partials=part, splitk_ws=sk)
        G.run_conv(d)
        outs.append(out)
        parts.append(part)
    assert rel_err(G.to_nchw_f32(outs[1]), ref) < 1e-2
    assert torch.equal(outs[1], outs[2]) and torch.equal(parts[1], parts[2])          # deterministic, flags re-armed
    # split vs unsplit: the same products summed in another fp32 order, then one bf16 rounding
    assert float((outs[0].float() - outs[1].float()).abs().max()) <= 2e-2 * float(ref.abs().max())
    torch.testing.assert_close(parts[0], parts[1], rtol=2e-4, atol=1e-2)
    assert int(ws[:4096].view(torch.int32).abs().sum()) == 0                          # flags left at rest
    # epilogue variants through the split path: SiLU + residual, no GroupNorm records
    out = torch.zeros(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    d = G.plain_conv_desc(xd, wd, B, H, W, Cout, ks, out, bias=bias.cuda(), residual=rd, act=1, splitk_ws=ws)
    G.run_conv(d)
    assert rel_err(G.to_nchw_f32(out), F.silu(ref) + res) < 1e-2


def test_conv_epilogue_variants(lib, conv_variant):
    g = torch.Generator().manual_seed(5)
    B, H, W, Cin, Cout = 2, 16, 16, 128, 512
    x = G.bf16_round(torch.randn(B, Cin, H, W, generator=g))
    w = G.bf16_round(torch.randn(Cout, Cin, 1, 1, generator=g) / math.sqrt(Cin))
    bias = torch.randn(Cout, generator=g)
    rs = torch.rand(B * H * W, generator=g) + 0.5
    res = G.bf16_round(torch.randn(B, Cout, H, W, generator=g))
    xd, wd = G.nhwc_bf16(x), G.pack_conv_weight(w)
    # row scale + bias + residual
    ref = F.conv2d(x, w) * rs.reshape(B, 1, H, W) + bias.reshape(1, -1, 1, 1) + res
    for direct in (False, True):
        out = torch.zeros(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
        d = G.plain_conv_desc([xd], wd, B, H, W, Cout, 1, out, bias=bias.cuda(), row_scale=rs.cuda(),
                              residual=G.nhwc_bf16(res))
        G.run_conv(d, direct)
        assert rel_err(G.to_nchw_f32(out), ref) < 1e-2, ("rowscale/residual", direct)
    # SiLU + PixelShuffle(2): weights packed (i j c') as weights.py does
    ref = F.pixel_shuffle(F.silu(F.conv2d(x, w, bias)), 2)
    w_ps = w.reshape(Cout // 4, 4, Cin).permute(1, 0, 2).reshape(Cout, Cin, 1, 1)
    b_ps = bias.reshape(Cout // 4, 4).permute(1, 0).reshape(Cout)
    for direct in (False, True):
        out = torch.zeros(B, 2 * H, 2 * W, Cout // 4, device="cuda", dtype=torch.bfloat16)
        d = G.plain_conv_desc([xd], G.pack_conv_weight(w_ps), B, H, W, Cout, 1, out, bias=b_ps.cuda(), act=1,
                              out_mode=_lib.OUT_PIXEL_SHUFFLE)
        G.run_conv(d, direct)
        assert rel_err(G.to_nchw_f32(out), ref) < 1e-2, ("pixelshuffle", direct)


def test_conv_downsample_as_strided_sources(lib):
    g = torch.Generator().manual_seed(6)
    B, H, W, Cin, Cout = 2, 32, 32, 128, 256
    x = G.bf16_round(torch.randn(B, Cin, H, W, generator=g))
    w = G.bf16_round(torch.randn(Cout, 4 * Cin, 1, 1, generator=g) / math.sqrt(4 * Cin))
    bias = torch.randn(Cout, generator=g)
    ref = F.conv2d(F.pixel_unshuffle(x, 2), w, bias)            # 'b c (h p1) (w p2) -> b (c p1 p2) h w'
    wp = w.reshape(Cout, Cin, 2, 2).permute(0, 2, 3, 1).reshape(Cout, 4 * Cin).contiguous().cuda().bfloat16()
    xd = G.nhwc_bf16(x)
    Ho, Wo = H // 2, W // 2
    srcs, phases = [], []
    for i in range(4):
        p1, p2 = i >> 1, i & 1
        srcs.append((xd.data_ptr() + (p1 * W + p2) * Cin * 2, H * W * Cin, 2 * W * Cin, 2 * Cin, Ho, Wo, Cin))
        phases.append((i, 0, 0, i * Cin))
    for direct in (False, True):
        out = torch.zeros(B, Ho, Wo, Cout, device="cuda", dtype=torch.bfloat16)
        d = G.conv_desc(srcs, phases, wp, 4 * Cin, B, Ho, Wo, Cout, out, bias=bias.cuda())
        G.run_conv(d, direct)
        assert rel_err(G.to_nchw_f32(out), ref) < 1e-2, direct


def test_init_conv_pack_and_7tap(lib):
    g = torch.Generator().manual_seed(7)
    B, H, W, Cout = 2, 32, 48, 128
    x, cond = torch.randn(B, 3, H, W, generator=g), torch.rand(B, 3, H, W, generator=g) * 2 - 1
    w = torch.randn(Cout, 6, 7, 7, generator=g) / math.sqrt(6 * 49)
    bias = torch.randn(Cout, generator=g)
    ref = F.conv2d(G.bf16_round(torch.cat((x, cond), 1)), G.bf16_round(w), bias, padding=3)
    pk = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.srgd_pack_input(G.P(x.cuda()), G.P(cond.cuda()), B, B, G.P(pk), B, H, W, G.stream()))
    wi = torch.zeros(Cout, 7, 64)
    wi[:, :, :42] = w.permute(0, 2, 3, 1).reshape(Cout, 7, 42)
    wp = wi.reshape(Cout, 7 * 64).cuda().bfloat16().contiguous()
    out = torch.zeros(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    d = G.conv_desc([(pk, H * W * 64, W * 64, 64, H, W, 64)], [(0, i - 3, 0, i * 64) for i in range(7)], wp, 7 * 64,
                    B, H, W, Cout, out, bias=bias.cuda())
    G.run_conv(d)
    assert rel_err(G.to_nchw_f32(out), ref) < 1e-2
    # null condition rows (x_self_cond=None -> zeros) and the CFG row mapping b % Bx
    pk2 = torch.empty(2 * B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.srgd_pack_input(G.P(x.cuda()), G.P(cond.cuda()), B, B, G.P(pk2), 2 * B, H, W,
                                   G.stream()))
    torch.cuda.synchronize()
    assert torch.equal(pk2[:B], pk)
    ch = torch.arange(64, device="cuda")
    is_cond = ((ch % 6) >= 3) & (ch < 42)
    assert float(pk2[B:][..., is_cond].float().abs().max()) == 0.0
    assert torch.equal(pk2[B:][..., ~is_cond], pk[..., ~is_cond])


# ------------------------------------------------------------------------------------------------
# GroupNorm / RMSNorm
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W,C", [(2, 32, 32, 128), (3, 8, 8, 1024), (1, 64, 64, 256), (2, 16, 16, 64), (2, 4, 256, 128),
                                     (1, 160, 128, 256)])      # last: tail-split tiles (160 tiles on 148 SMs)
def test_groupnorm_fused_stats_and_apply(lib, conv_variant, B, H, W, C):
    g = torch.Generator().manual_seed(C + H)
    Cin = 128
    x = G.bf16_round(torch.randn(B, Cin, H, W, generator=g))
    w = G.bf16_round(torch.randn(C, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9))
    bias = torch.randn(C, generator=g) * 0.5
    conv = F.conv2d(x, w, bias, padding=1)
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    ss = torch.randn(B, 2 * C + 7, generator=g) * 0.3             # row stride != 2C on purpose
    resid = G.bf16_round(torch.randn(B, C, H, W, generator=g))
    mt = lib.srgd_conv_m_tiles(B, H, W)                           # 64-byte records: one per (M tile, sample slot)
    part = torch.full((mt * 8 * 2,), float("nan"), device="cuda")
    out = torch.zeros(B, H, W, C, device="cuda", dtype=torch.bfloat16)
    d = G.plain_conv_desc([G.nhwc_bf16(x)], G.pack_conv_weight(w), B, H, W, C, 3, out, bias=bias.cuda(),
                          gn_partials=part)
    G.run_conv(d)
    stats = torch.empty(B * 8 * 2, device="cuda")
    _lib.check(lib.srgd_groupnorm_finalize(G.P(part), G.P(stats), B, H, W, C, G.stream()))
    grp = conv.reshape(B, 8, -1)
    ref_mean, ref_rstd = grp.mean(-1), (grp.var(-1, unbiased=False) + 1e-5).rsqrt()
    st = stats.cpu().reshape(B, 8, 2)
    torch.testing.assert_close(st[..., 0], ref_mean, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(st[..., 1], ref_rstd, rtol=1e-3, atol=1e-4)
    # stand-alone statistics kernel agrees (computed from the bf16-rounded conv output)
    stats2 = torch.empty_like(stats)
    _lib.check(lib.srgd_groupnorm_stats(G.P(out), G.P(stats2), B, H, W, C, G.stream()))
    torch.testing.assert_close(stats2.cpu().reshape(B, 8, 2)[..., 0], ref_mean, rtol=1e-2, atol=2e-3)
    # apply: GN affine, (scale+1, shift), SiLU, + residual -- reference evaluated on the bf16 conv output
    conv_b = G.to_nchw_f32(out)
    scale, shift = ss[:, :C], ss[:, C:2 * C]
    y = F.group_norm(conv_b, 8, gamma, beta, eps=1e-5)
    y = F.silu(y * (scale[:, :, None, None] + 1) + shift[:, :, None, None]) + resid
    yd = torch.empty_like(out)
    want_inv = C in (128, 256) and (H * W) % 4 == 0                          # fused RMSNorm statistic (model.py:207)
    inv = torch.full((B * H * W,), float("nan"), device="cuda") if want_inv else None
    _lib.check(lib.srgd_groupnorm_apply(G.P(out), B, G.P(stats), None, G.P(gamma.cuda()), G.P(beta.cuda()),
                                        G.P(ss.cuda()), 2 * C + 7, G.P(G.nhwc_bf16(resid)), G.P(yd), G.P(inv),
                                        B, H, W, C, G.stream()))
    torch.cuda.synchronize()
    assert rel_err(G.to_nchw_f32(yd), y) < 1.5e-2
    if want_inv:
        ref_inv = 1.0 / yd.float().reshape(B * H * W, C).norm(dim=1).clamp(min=1e-12)
        torch.testing.assert_close(inv, ref_inv, rtol=1e-5, atol=1e-7)
    # product path: no finalize launch, every block folds the partial records itself -- bit-identical result
    yd2 = torch.empty_like(out)
    _lib.check(lib.srgd_groupnorm_apply(G.P(out), B, None, G.P(part), G.P(gamma.cuda()), G.P(beta.cuda()),
                                        G.P(ss.cuda()), 2 * C + 7, G.P(G.nhwc_bf16(resid)), G.P(yd2), None,
                                        B, H, W, C, G.stream()))
    torch.cuda.synchronize()
    assert torch.equal(yd2, yd)
    assert lib.srgd_groupnorm_apply(G.P(out), B, G.P(stats), G.P(part), G.P(gamma.cuda()), G.P(beta.cuda()), None, 0,
                                    None, G.P(yd2), None, B, H, W, C, G.stream()) != 0      # exactly one of the two


def test_groupnorm_apply_final_fused_conv(lib):
    """Last GroupNorm pass + residual fused with the final 1x1 conv 128 -> 3 (model.py:674-675, 724-725)."""
    g = torch.Generator().manual_seed(9)
    B, H, W, C = 2, 16, 24, 128
    x = G.bf16_round(torch.randn(B, C, H, W, generator=g) * 1.5 + 0.2)
    res = G.bf16_round(torch.randn(B, C, H, W, generator=g))
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    w3, b3 = torch.randn(3, C, generator=g) / math.sqrt(C), torch.randn(3, generator=g) * 0.1
    xd = G.nhwc_bf16(x)
    stats = torch.empty(B * 8 * 2, device="cuda")
    _lib.check(lib.srgd_groupnorm_stats(G.P(xd), G.P(stats), B, H, W, C, G.stream()))
    eps = torch.full((B, 3, H, W), float("nan"), device="cuda")
    _lib.check(lib.srgd_groupnorm_apply_final(G.P(xd), G.P(stats), None, G.P(gamma.cuda()), G.P(beta.cuda()),
                                              G.P(G.nhwc_bf16(res)), G.P(w3.cuda().contiguous()), G.P(b3.cuda()), G.P(eps),
                                              B, H, W, C, G.stream()))
    torch.cuda.synchronize()
    ref = F.conv2d(F.silu(F.group_norm(x, 8, gamma, beta, eps=1e-5)) + res, w3.reshape(3, C, 1, 1), b3)
    assert rel_err(eps.cpu(), ref) < 5e-3


@pytest.mark.parametrize("C", [64, 128, 256, 512, 1024])
def test_rmsnorm_kernels(lib, C):
    g = torch.Generator().manual_seed(C)
    M = 777
    x = G.bf16_round(torch.randn(M, C, generator=g) * 2)
    gain = 1 + 0.1 * torch.randn(C, generator=g)
    res = G.bf16_round(torch.randn(M, C, generator=g))
    inv = torch.empty(M, device="cuda")
    xd = x.cuda().bfloat16()
    _lib.check(lib.srgd_pixel_inv_norm(G.P(xd), G.P(inv), M, C, G.stream()))
    torch.testing.assert_close(inv.cpu(), 1.0 / x.norm(dim=1).clamp(min=1e-12), rtol=1e-5, atol=1e-7)
    ref = F.normalize(x, dim=1) * gain * math.sqrt(C) + res
    y = torch.empty_like(xd)
    _lib.check(lib.srgd_rmsnorm_residual(G.P(xd), G.P(gain.cuda()), G.P(res.cuda().bfloat16()),
                                         G.P(y), M, C, G.stream()))
    torch.cuda.synchronize()
    assert rel_err(y.float().cpu(), ref) < 1e-2


# ------------------------------------------------------------------------------------------------
# attention cores
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,N", [(2, 1024), (1, 4096), (3, 200)])
def test_linear_attention_core(lib, B, N):
    g = torch.Generator().manual_seed(N)
    qkv = G.bf16_round(torch.randn(B, N, 384, generator=g) * 1.5)
    q, k, v = (t.reshape(B, N, 4, 32).permute(0, 2, 3, 1) for t in qkv.chunk(3, dim=-1))   # b h d n
    qs, ks = q.softmax(dim=-2) * 32 ** -0.5, k.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", ks, v)
    ref = torch.einsum("bhde,bhdn->bhen", ctx, qs).permute(0, 3, 1, 2).reshape(B, N, 128)
    out = torch.empty(B, N, 128, device="cuda", dtype=torch.bfloat16)
    wsb = lib.srgd_linear_attention_workspace(B, N, 4)
    ws = torch.empty(wsb, device="cuda", dtype=torch.uint8)
    _lib.check(lib.srgd_linear_attention(G.P(qkv.cuda().bfloat16()), G.P(out), B, N, 4, G.P(ws), wsb,
                                         G.stream()))
    torch.cuda.synchronize()
    assert rel_err(out.float().cpu(), ref) < 1e-2


# the last four are the bench-scale instances: many 128-pixel tiles per thread block (ring-slot reuse, the x slot handed
# back by the epilogue, staged output rows), which the small shapes -- one or two tiles per block -- never reach
@pytest.mark.parametrize("B,H,W,C", [(2, 32, 32, 128), (1, 64, 128, 128), (3, 16, 8, 256), (2, 64, 64, 256),
                                     (20, 16, 16, 128), (16, 256, 256, 128), (16, 128, 128, 128), (16, 128, 128, 256),
                                     (1, 256, 256, 128)])
def test_linear_attention_block_fused(lib, B, H, W, C):
    """Whole LinearAttention block + residual on tcgen05 (linattn_fused.cu) vs the oracle's
    _linear_attention (model.py:287-324) + x (model.py:703)."""
    g = torch.Generator().manual_seed(B * 1000 + H + C)
    N = H * W
    assert lib.srgd_linear_attention_block_supported(N, C, 4) == 1
    x = G.bf16_round(torch.randn(B, C, H, W, generator=g) * 1.3)
    sd = {"a.norm.g": (1 + 0.1 * torch.randn(1, C, 1, 1, generator=g)),
          "a.to_qkv.weight": torch.randn(384, C, 1, 1, generator=g) * (2.0 / math.sqrt(C)),
          "a.to_out.0.weight": torch.randn(C, 128, 1, 1, generator=g) * (1.0 / math.sqrt(128)),
          "a.to_out.0.bias": torch.randn(C, generator=g) * 0.1,
          "a.to_out.1.g": (1 + 0.1 * torch.randn(1, C, 1, 1, generator=g))}
    # the kernel sees bf16 weights with the pre-norm gain folded in (srgd_b200/weights.py)
    wq = G.bf16_round(sd["a.to_qkv.weight"].reshape(384, C) * (sd["a.norm.g"].reshape(1, C) * math.sqrt(C)))
    wo = G.bf16_round(sd["a.to_out.0.weight"].reshape(C, 128))
    ref_sd = dict(sd)
    ref_sd["a.norm.g"] = torch.ones(1, C, 1, 1) / math.sqrt(C) * 1.0          # gain already inside wq
    ref_sd["a.to_qkv.weight"] = wq.reshape(384, C, 1, 1)
    ref_sd["a.to_out.0.weight"] = wo.reshape(C, 128, 1, 1)
    if B * N > 100000:                       # big instances: the fp32 oracle on the GPU (strict fp32, TF32 off)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        ref = (O._linear_attention({k: v.cuda() for k, v in ref_sd.items()}, "a", x.cuda(), 4, 32) + x.cuda()).cpu()
    else:
        ref = O._linear_attention(ref_sd, "a", x, 4, 32) + x
    xd = G.nhwc_bf16(x)
    wsb = lib.srgd_linear_attention_block_workspace(B, N, C, 4)
    ws = torch.empty(wsb, device="cuda", dtype=torch.uint8)
    outs = []
    for _ in range(2):
        out = torch.empty_like(xd)
        _lib.check(lib.srgd_linear_attention_block(G.P(xd), None, G.P(wq.cuda().bfloat16()), G.P(wo.cuda().bfloat16()),
                                                   G.P(sd["a.to_out.0.bias"].cuda()),
                                                   G.P(sd["a.to_out.1.g"].reshape(-1).contiguous().cuda()), G.P(out), B, N,
                                                   C, 4, G.P(ws), wsb, G.stream()), "linear_attention_block")
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1]), "the fused LinearAttention block is not deterministic"
    got = G.to_nchw_f32(out)
    err = (got - ref).abs()
    print(f"fused LA block B={B} {H}x{W} C={C}: max {float(err.max()):.4f} rms {float(err.pow(2).mean().sqrt()):.5f} "
          f"ref rms {float(ref.pow(2).mean().sqrt()):.3f}")
    assert rel_err(got, ref) < 1.5e-2


def test_linear_attention_block_reference_jumps(lib):
    """Online softmax over n with a moving exponent reference: k logits that grow by hundreds of nats from tile to
    tile force the 'redo with the exact max' and 'rescale the TMEM context' paths of la_ctx_pp_kernel."""
    g = torch.Generator().manual_seed(5)
    B, H, W, C = 1, 32, 32, 128
    N = H * W
    x = torch.randn(B, C, H, W, generator=g)
    ramp = torch.linspace(0.0005, 12.0, N).reshape(1, 1, H, W)                # later pixels point harder along e0
    x[:, 0:1] = x[:, 0:1] * 0.1 + ramp * 8
    x = G.bf16_round(x)
    wqkv = torch.randn(384, C, generator=g) * (2.0 / math.sqrt(C))
    wqkv[128:256, 0] += torch.linspace(-3, 3, 128) * 36                       # k rows with huge gain on channel 0
    sd = {"a.norm.g": torch.ones(1, C, 1, 1) / math.sqrt(C), "a.to_qkv.weight": G.bf16_round(wqkv).reshape(384, C, 1, 1),
          "a.to_out.0.weight": G.bf16_round(torch.randn(C, 128, generator=g) / math.sqrt(128)).reshape(C, 128, 1, 1),
          "a.to_out.0.bias": torch.randn(C, generator=g) * 0.1,
          "a.to_out.1.g": (1 + 0.1 * torch.randn(1, C, 1, 1, generator=g))}
    ref = O._linear_attention(sd, "a", x, 4, 32) + x
    k = torch.nn.functional.conv2d(torch.nn.functional.normalize(x, dim=1), sd["a.to_qkv.weight"])[:, 128:256]
    assert float(k.reshape(128, -1).max(dim=1).values.max() - k.reshape(128, -1)[:, :64].max(dim=1).values.min()) > 60
    xd = G.nhwc_bf16(x)
    out = torch.empty_like(xd)
    wsb = lib.srgd_linear_attention_block_workspace(B, N, C, 4)
    ws = torch.empty(wsb, device="cuda", dtype=torch.uint8)
    _lib.check(lib.srgd_linear_attention_block(G.P(xd), None, G.P(sd["a.to_qkv.weight"].reshape(384, C).cuda().bfloat16()),
                                               G.P(sd["a.to_out.0.weight"].reshape(C, 128).cuda().bfloat16()),
                                               G.P(sd["a.to_out.0.bias"].cuda()),
                                               G.P(sd["a.to_out.1.g"].reshape(-1).contiguous().cuda()), G.P(out), B, N, C, 4,
                                               G.P(ws), wsb, G.stream()), "linear_attention_block")
    torch.cuda.synchronize()
    got = G.to_nchw_f32(out)
    assert torch.isfinite(got).all()
    print(f"reference-jump LA block: max err {float((got - ref).abs().max()):.4f}, ref rms {float(ref.pow(2).mean().sqrt()):.3f}")
    assert rel_err(got, ref) < 2e-2


def test_linear_attention_block_unsupported_shapes(lib):
    assert lib.srgd_linear_attention_block_supported(64, 128, 4) == 0       # N % 128 != 0
    assert lib.srgd_linear_attention_block_supported(1024, 512, 4) == 0     # C = 512 takes the unfused path
    assert lib.srgd_linear_attention_block_workspace(2, 64, 128, 4) == 0


@pytest.mark.parametrize("B,N", [(2, 1024), (1, 64), (2, 100)])
def test_full_attention_core(lib, B, N):
    g = torch.Generator().manual_seed(N + 1)
    qkv = G.bf16_round(torch.randn(B, N, 384, generator=g) * 1.5)
    q, k, v = (t.reshape(B, N, 4, 32).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1))   # b h n d
    ref = O._attend(q, k, v).permute(0, 2, 1, 3).reshape(B, N, 128)
    out = torch.empty(B, N, 128, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.srgd_attention(G.P(qkv.cuda().bfloat16()), G.P(out), B, N, 4, G.stream()))
    torch.cuda.synchronize()
    assert rel_err(out.float().cpu(), ref) < 1e-2


@pytest.mark.parametrize("B,N", [(2, 1024), (1, 128), (3, 256), (16, 1024)])
def test_full_attention_tcgen05(lib, B, N):
    """tcgen05 flash attention (attention_tc.cu) vs the oracle's Attend restatement."""
    g = torch.Generator().manual_seed(N + B)
    assert lib.srgd_attention_tc_supported(N, 4) == 1 and lib.srgd_attention_tc_supported(100, 4) == 0
    qkv = G.bf16_round(torch.randn(B, N, 384, generator=g) * 1.5)
    q, k, v = (t.reshape(B, N, 4, 32).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1))   # b h n d
    ref = O._attend(q, k, v).permute(0, 2, 1, 3).reshape(B, N, 128)
    out = torch.empty(B, N, 128, device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.srgd_attention_tc(G.P(qkv.cuda().bfloat16()), G.P(out), B, N, 4, G.stream()))
    torch.cuda.synchronize()
    err = (out.float().cpu() - ref).abs()
    print(f"tcgen05 attention B={B} N={N}: max {float(err.max()):.4f} rms {float(err.pow(2).mean().sqrt()):.5f}")
    assert rel_err(out.float().cpu(), ref) < 1e-2


# ------------------------------------------------------------------------------------------------
# small dense layers / embeddings / final conv
# ------------------------------------------------------------------------------------------------
def test_embedding_kernels(lib):
    g = torch.Generator().manual_seed(11)
    B = 5
    lsnr = torch.tensor([-10.0, -3.3, 0.1, 4.4, 9.2])
    wts = torch.randn(16, generator=g)
    feats = torch.empty(B, 33, device="cuda")
    _lib.check(lib.srgd_fourier_features(G.P(lsnr.cuda()), G.P(wts.cuda()), G.P(feats), B, 16, G.stream()))
    fr = lsnr[:, None] * wts[None, :] * 2 * math.pi
    ref = torch.cat((lsnr[:, None], fr.sin(), fr.cos()), -1)
    torch.testing.assert_close(feats.cpu(), ref, rtol=0, atol=2e-5)
    for act, fn in ((0, lambda t: t), (1, F.silu), (2, F.gelu)):
        x, w, b = torch.randn(B, 512, generator=g), torch.randn(300, 512, generator=g) / 22, torch.randn(300, generator=g)
        y = torch.empty(B, 300, device="cuda")
        _lib.check(lib.srgd_dense_rows(G.P(x.cuda()), G.P(w.cuda()), G.P(b.cuda()), G.P(y), B, 300,
                                       512, act, 0, G.stream()))
        torch.testing.assert_close(y.cpu(), F.linear(fn(x), w, b), rtol=1e-4, atol=1e-4)
    t = torch.randn(B, 512, generator=g)
    table = torch.randn(3, 512, generator=g)
    labels = torch.tensor([0, -1, 2, 1, -1], dtype=torch.int32)
    td = t.cuda()
    _lib.check(lib.srgd_add_class_rows(G.P(td), G.P(table.cuda()), G.P(labels.cuda()), B, 512, 3, G.stream()))
    ref = t.clone()
    for i, l in enumerate(labels.tolist()):
        if l >= 0:
            ref[i] += table[l]
    torch.testing.assert_close(td.cpu(), ref, rtol=0, atol=0)


def test_final_conv(lib):
    g = torch.Generator().manual_seed(12)
    B, H, W, Cc = 2, 16, 24, 128
    h = G.bf16_round(torch.randn(B, Cc, H, W, generator=g))
    w, b = torch.randn(3, Cc, generator=g) / 11, torch.randn(3, generator=g)
    eps = torch.empty(B, 3, H, W, device="cuda")
    _lib.check(lib.srgd_final_conv(G.P(G.nhwc_bf16(h)), G.P(w.cuda()), G.P(b.cuda()), G.P(eps), B, H,
                                   W, Cc, 3, G.stream()))
    torch.testing.assert_close(eps.cpu(), F.conv2d(h, w.reshape(3, Cc, 1, 1), b), rtol=1e-4, atol=1e-4)
