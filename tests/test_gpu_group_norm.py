"""GPU parity of the fused GroupNorm (+ SiLU) kernels (`csrc/unet_ops.cu`, through the C ABI) against PyTorch's own
`F.group_norm` / `F.silu` evaluated in fp32 on the same 16-bit inputs -- the op pair the reference's UNet runs
(`torch.nn.GroupNorm(32, C)` + SiLU inside diffusers' ResnetBlock2D / Transformer2DModel, driven by
pipeline_guided_attention.py:583-743).  Floating point: the bound is the output dtype's rounding (stated per test).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# (n, C, h, w, groups): every GroupNorm shape of the SD-1.4 UNet at 64x64 latents (incl. the concatenated skip inputs of
# the up blocks), batch 1 (guidance passes) and 2 (CFG), the SD-2.x 96x96 latent sizes, the tiny test UNet (8 groups,
# groups narrower than a 128-bit vector), and ragged pixel counts
SHAPES = [
    (1, 320, 64, 64, 32), (2, 320, 64, 64, 32), (1, 640, 32, 32, 32), (1, 1280, 16, 16, 32), (1, 1280, 8, 8, 32),
    (2, 2560, 8, 8, 32), (1, 2560, 16, 16, 32), (1, 1920, 16, 16, 32), (1, 1920, 32, 32, 32), (1, 1280, 32, 32, 32),
    (1, 960, 32, 32, 32), (1, 960, 64, 64, 32), (1, 640, 64, 64, 32), (8, 320, 64, 64, 32),
    (1, 320, 96, 96, 32), (1, 1280, 12, 12, 32), (1, 640, 24, 24, 32),
    (1, 32, 64, 64, 8), (2, 96, 32, 32, 8), (1, 192, 8, 8, 8), (1, 256, 16, 16, 8), (1, 64, 5, 7, 8), (3, 128, 1, 1, 8),
]
TOL = {torch.float16: (2e-3, 2e-3), torch.bfloat16: (1.6e-2, 1.6e-2)}      # (rtol, atol) on O(1) outputs


def _inputs(n, c, h, w, dtype, seed=0, offset=0.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = (torch.randn(n, c, h, w, device=DEV, generator=g) * 1.5 + offset).to(dtype)
    x = x.contiguous(memory_format=torch.channels_last)
    wgt = (1.0 + 0.3 * torch.randn(c, device=DEV, generator=g)).to(dtype)
    b = (0.2 * torch.randn(c, device=DEV, generator=g)).to(dtype)
    dy = torch.randn(n, c, h, w, device=DEV, generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    return x, wgt, b, dy


def _reference(x, wgt, b, groups, eps, silu, dy):
    xr = x.float().detach().requires_grad_(True)
    y = F.group_norm(xr, groups, wgt.float(), b.float(), eps)
    if silu:
        y = F.silu(y)
    (dx,) = torch.autograd.grad(y, xr, dy.float())
    return y.detach(), dx


@pytest.mark.parametrize("silu", [False, True])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_group_norm_matches_torch(shape, dtype, silu):
    from guided_attention_b200 import ops
    n, c, h, w, groups = shape
    x, wgt, b, dy = _inputs(n, c, h, w, dtype, seed=n + c + h)
    assert ops.group_norm_supported(x, wgt, b, groups)
    xq = x.detach().requires_grad_(True)
    y = ops.group_norm(xq, wgt, b, groups, 1e-5, silu=silu)
    (dx,) = torch.autograd.grad(y, xq, dy)
    y_ref, dx_ref = _reference(x, wgt, b, groups, 1e-5, silu, dy)
    rtol, atol = TOL[dtype]
    assert y.dtype == dtype and y.shape == x.shape and y.is_contiguous(memory_format=torch.channels_last)
    assert torch.allclose(y.float(), y_ref, rtol=rtol, atol=atol), float((y.float() - y_ref).abs().max())
    # the gradient is a difference of O(1) terms scaled by rstd: bound it against its own largest entry
    err = float((dx.float() - dx_ref).abs().max() / dx_ref.abs().max().clamp_min(1e-6))
    assert err < (4e-3 if dtype == torch.float16 else 3e-2), err


def test_group_norm_is_bit_stable_and_accepts_channels_first():
    from guided_attention_b200 import ops
    x, wgt, b, dy = _inputs(2, 640, 32, 32, torch.float16, seed=3)
    outs = []
    for _ in range(3):
        xq = x.detach().requires_grad_(True)
        y = ops.group_norm(xq, wgt, b, 32, 1e-5, silu=True)
        (dx,) = torch.autograd.grad(y, xq, dy)
        outs.append((y.clone(), dx.clone()))
    assert all(torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) for o in outs[1:])
    # a channels-first (NCHW-contiguous) input gives the same values
    xc = x.contiguous(memory_format=torch.contiguous_format)
    assert not xc.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(ops.group_norm(xc, wgt, b, 32, 1e-5, silu=True), outs[0][0])


def test_group_norm_large_mean_offset():
    """Statistics are merged with Chan's update over per-chunk (mean, M2): a mean 40x the standard deviation must not
    cost more than the fp16 rounding of the INPUT already does."""
    from guided_attention_b200 import ops
    x, wgt, b, dy = _inputs(1, 320, 64, 64, torch.float16, seed=5, offset=60.0)
    y = ops.group_norm(x, wgt, b, 32, 1e-5, silu=False)
    y_ref, _ = _reference(x, wgt, b, 32, 1e-5, False, dy)
    assert torch.allclose(y.float(), y_ref, rtol=3e-3, atol=3e-3), float((y.float() - y_ref).abs().max())


def test_group_norm_unsupported_shapes_are_declined():
    from guided_attention_b200 import ops
    assert ops.group_norm_ws_bytes(1, 64, 36, 6) < 0          # channels not a multiple of 8
    assert ops.group_norm_ws_bytes(1, 64, 40, 8) < 0          # 5-channel groups: a 128-bit vector would touch three
    assert ops.group_norm_ws_bytes(1, 64, 320, 32) > 0
    x = torch.randn(1, 320, 8, 8, device=DEV)                 # fp32 stays on PyTorch's op
    assert not ops.group_norm_supported(x, torch.ones(320, device=DEV), torch.zeros(320, device=DEV), 32)
    w = torch.ones(320, device=DEV, dtype=torch.float16, requires_grad=True)
    assert not ops.group_norm_supported(x.half(), w, torch.zeros(320, device=DEV, dtype=torch.float16), 32)


def test_unet_with_fused_norms_matches_stock_unet():
    """The whole (tiny, fp16) UNet forward + backward to the latents with the norms routed through the fused kernels vs
    the same weights on PyTorch's GroupNorm + SiLU."""
    from guided_attention_b200 import ptp_utils
    from guided_attention_b200.substrate import UNetConfig, build_unet
    cfg = UNetConfig.tiny(sample_size=32)
    stock = build_unet(cfg, seed=0, dtype=torch.float16, device=DEV).requires_grad_(False)
    fused = build_unet(cfg, seed=0, dtype=torch.float16, device=DEV).requires_grad_(False)
    assert ptp_utils.register_fused_norms(fused) == 61
    g = torch.Generator(device=DEV).manual_seed(1)
    lat = torch.randn(2, 4, 32, 32, device=DEV, generator=g).half()
    emb = torch.randn(2, 77, cfg.cross_attention_dim, device=DEV, generator=g).half()
    res = []
    for net in (stock, fused):
        z = lat.clone().requires_grad_(True)
        out = net(z, 321, encoder_hidden_states=emb).sample
        (gz,) = torch.autograd.grad(out.float().square().sum(), z)
        res.append((out.float(), gz.float()))
    cos_o = float(F.cosine_similarity(res[0][0].flatten(), res[1][0].flatten(), dim=0))
    cos_g = float(F.cosine_similarity(res[0][1].flatten(), res[1][1].flatten(), dim=0))
    rel_o = float((res[0][0] - res[1][0]).abs().max() / res[0][0].abs().max())
    assert cos_o > 0.9999 and cos_g > 0.999 and rel_o < 2e-2, (cos_o, cos_g, rel_o)
    assert np.isfinite(res[1][1].cpu().numpy()).all()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 320, 64, 64, 32), (2, 1280, 16, 16, 32), (1, 640, 24, 24, 32), (2, 96, 8, 8, 8)],
                         ids=lambda s: "x".join(map(str, s)))
def test_group_norm_with_shift_matches_torch(shape, dtype):
    """`shift` (n, C): the conv bias + time-embedding projection the ResNet block adds in front of norm2."""
    from guided_attention_b200 import ops
    n, c, h, w, groups = shape
    x, wgt, b, dy = _inputs(n, c, h, w, dtype, seed=11)
    shift = (0.7 * torch.randn(n, c, device=DEV, generator=torch.Generator(device=DEV).manual_seed(2))).to(dtype)
    xq = x.detach().requires_grad_(True)
    y = ops.group_norm(xq, wgt, b, groups, 1e-5, silu=True, shift=shift)
    (dx,) = torch.autograd.grad(y, xq, dy)
    xr = x.float().detach().requires_grad_(True)
    y_ref = F.silu(F.group_norm(xr + shift.float()[:, :, None, None], groups, wgt.float(), b.float(), 1e-5))
    (dx_ref,) = torch.autograd.grad(y_ref, xr, dy.float())
    rtol, atol = TOL[dtype]
    assert torch.allclose(y.float(), y_ref, rtol=rtol, atol=atol), float((y.float() - y_ref).abs().max())
    err = float((dx.float() - dx_ref).abs().max() / dx_ref.abs().max())
    assert err < (4e-3 if dtype == torch.float16 else 3e-2), err


@pytest.mark.parametrize("with_b", [False, True])
def test_add_bias_residual_is_exactly_the_rounded_sum(with_b):
    from guided_attention_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(4)
    a = torch.randn(2, 640, 17, 9, device=DEV, generator=g).half().contiguous(memory_format=torch.channels_last)
    b = torch.randn(2, 640, 17, 9, device=DEV, generator=g).half() if with_b else None     # channels-first on purpose
    bias = torch.randn(640, device=DEV, generator=g).half()
    aq = a.detach().requires_grad_(True)
    bq = b.detach().requires_grad_(True) if with_b else None
    out = ops.add_bias_residual(aq, bias, bq)
    ref = a.float() + bias.float()[None, :, None, None] + (b.float() if with_b else 0.0)
    assert torch.equal(out, ref.half())                       # one rounding of the fp32 sum
    gout = torch.randn_like(out)
    grads = torch.autograd.grad(out, [aq] + ([bq] if with_b else []), gout)
    assert all(torch.equal(gr, gout) for gr in grads)


def test_patched_convolutions_match_cudnn_with_bias():
    """1x1 convolutions run as `F.linear` on the channels-last view, the others as bias-free cuDNN + one bias pass."""
    from guided_attention_b200 import ptp_utils
    torch.manual_seed(0)
    for (cin, cout, k, stride) in ((320, 320, 1, 1), (640, 320, 1, 1), (320, 640, 3, 1), (320, 320, 3, 2), (4, 320, 3, 1)):
        conv = torch.nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2).to(DEV).half().requires_grad_(False)
        conv = conv.to(memory_format=torch.channels_last)
        x = torch.randn(2, cin, 32, 32, device=DEV).half().contiguous(memory_format=torch.channels_last)
        xq = x.clone().requires_grad_(True)
        ref = conv(xq)
        (gref,) = torch.autograd.grad(ref.float().square().sum(), xq)
        holder = torch.nn.Sequential(conv)
        ptp_utils.register_fused_norms(holder)
        xq2 = x.clone().requires_grad_(True)
        got = conv(xq2)
        (ggot,) = torch.autograd.grad(got.float().square().sum(), xq2)
        assert got.shape == ref.shape
        assert float((got.float() - ref.float()).abs().max() / ref.float().abs().max()) < 4e-3, (cin, cout, k)
        assert float((ggot.float() - gref.float()).abs().max() / gref.float().abs().max()) < 1e-2, (cin, cout, k)


def test_fused_resnet_block_matches_stock_block():
    from guided_attention_b200 import ptp_utils
    from guided_attention_b200.substrate.unet import ResnetBlock2D
    for (cin, cout) in ((320, 320), (960, 640)):
        torch.manual_seed(1)
        stock = ResnetBlock2D(cin, cout, 1280, 32).to(DEV).half().requires_grad_(False).to(memory_format=torch.channels_last)
        fused = ResnetBlock2D(cin, cout, 1280, 32).to(DEV).half().requires_grad_(False).to(memory_format=torch.channels_last)
        fused.load_state_dict(stock.state_dict())
        ptp_utils.register_fused_norms(fused)
        assert fused.fused_forward is not None
        x = torch.randn(2, cin, 32, 32, device=DEV).half().contiguous(memory_format=torch.channels_last)
        temb = torch.randn(2, 1280, device=DEV).half()
        res = []
        for blk in (stock, fused):
            xq = x.clone().requires_grad_(True)
            out = blk(xq, temb)
            (gx,) = torch.autograd.grad(out.float().square().sum(), xq)
            res.append((out.float(), gx.float()))
        assert float((res[0][0] - res[1][0]).abs().max() / res[0][0].abs().max()) < 5e-3
        assert float(F.cosine_similarity(res[0][1].flatten(), res[1][1].flatten(), dim=0)) > 0.9999


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 4096, 1280), (2, 1024, 2560), (1, 256, 5120), (3, 7, 64)],
                         ids=lambda s: "x".join(map(str, s)))
def test_geglu_matches_torch(shape, dtype):
    """`h * gelu(gate)` of the feed-forward (exact erf GELU) and its gradient vs PyTorch in fp32 on the same inputs."""
    from guided_attention_b200 import ops
    b, n, inner = shape
    g = torch.Generator(device=DEV).manual_seed(inner)
    proj = (1.5 * torch.randn(b, n, 2 * inner, device=DEV, generator=g)).to(dtype)
    d_out = torch.randn(b, n, inner, device=DEV, generator=g).to(dtype)
    assert ops.geglu_supported(proj)
    pq = proj.detach().requires_grad_(True)
    out = ops.geglu(pq)
    (dp,) = torch.autograd.grad(out, pq, d_out)
    pr = proj.float().detach().requires_grad_(True)
    h, gate = pr.chunk(2, dim=-1)
    ref = h * F.gelu(gate)
    (dp_ref,) = torch.autograd.grad(ref, pr, d_out.float())
    rtol, atol = TOL[dtype]
    assert out.shape == ref.shape and out.dtype == dtype
    assert torch.allclose(out.float(), ref, rtol=rtol, atol=atol), float((out.float() - ref).abs().max())
    assert torch.allclose(dp.float(), dp_ref, rtol=rtol, atol=4 * atol), float((dp.float() - dp_ref).abs().max())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 4096, 320), (2, 1024, 640), (1, 256, 1280), (2, 64, 1280), (3, 5, 64), (1, 7, 2048)],
                         ids=lambda s: "x".join(map(str, s)))
def test_layer_norm_matches_torch(shape, dtype):
    from guided_attention_b200 import ops
    b, n, c = shape
    g = torch.Generator(device=DEV).manual_seed(c)
    x = (1.3 * torch.randn(b, n, c, device=DEV, generator=g) + 0.5).to(dtype)
    wgt = (1.0 + 0.3 * torch.randn(c, device=DEV, generator=g)).to(dtype)
    bias = (0.2 * torch.randn(c, device=DEV, generator=g)).to(dtype)
    dy = torch.randn(b, n, c, device=DEV, generator=g).to(dtype)
    assert ops.layer_norm_supported(x, wgt, bias)
    xq = x.detach().requires_grad_(True)
    y = ops.layer_norm(xq, wgt, bias, 1e-5)
    (dx,) = torch.autograd.grad(y, xq, dy)
    xr = x.float().detach().requires_grad_(True)
    y_ref = F.layer_norm(xr, (c,), wgt.float(), bias.float(), 1e-5)
    (dx_ref,) = torch.autograd.grad(y_ref, xr, dy.float())
    rtol, atol = TOL[dtype]
    assert torch.allclose(y.float(), y_ref, rtol=rtol, atol=atol), float((y.float() - y_ref).abs().max())
    err = float((dx.float() - dx_ref).abs().max() / dx_ref.abs().max())
    assert err < (4e-3 if dtype == torch.float16 else 3e-2), err
