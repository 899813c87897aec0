"""The 2-CTA (cta_group::2, M = 256) form of the implicit-GEMM fprop kernel: it is only selected for layers with at least
74 pair tiles, i.e. at batch sizes the small unit tests of test_igemm_gpu.py never reach, so these cases run the BASELINE
layer shapes at batch 16-149 against torch's fp32 convolution of the same bf16 operands (models/pix2pix.py:63-69,99-105,
models/wrapper.py:229-233).  PAI_NO_CTA_PAIR=1 would route the same calls through the 1-CTA kernel."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).bfloat16()


def _nchw(t):
    return t.float().permute(0, 3, 1, 2)


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("n,h,w,cin,cout", [(24, 128, 128, 64, 128),     # enc1 / D1: n_tile 128
                                             (149, 16, 16, 64, 512),      # odd number of 128-pixel tiles, 2 n-tiles of 256
                                             (16, 64, 64, 128, 256),      # enc2 / D2
                                             (40, 32, 32, 256, 512)])     # enc3 / D3
def test_pair_conv_fprop(n, h, w, cin, cout):
    from pai_b200 import ops
    x = _rand((n, h, w, cin), 1)
    wt = _rand((cout, cin, 4, 4), 2, 0.05)
    bias = torch.randn(cout, device="cuda")
    wp = ops.pack_conv_weight(wt.float())
    ref = _nhwc(F.conv2d(_nchw(x), wt.float(), bias, stride=2, padding=1))
    scale = max(1.0, ref.abs().max().item())
    y = ops.conv4x4_fprop(x, wp, cout, stride=2, bias=bias, out_f32=True)
    assert (y - ref).abs().max().item() < 2e-3 * scale
    y2 = ops.conv4x4_fprop(x, wp, cout, stride=2, bias=bias, act=ops.ACT_LEAKY, slope=0.2)
    assert (y2.float() - F.leaky_relu(ref, 0.2)).abs().max().item() < 1e-2 * scale
    raw, part = ops.conv4x4_fprop_bnstats(x, wp, cout, bias=bias)
    assert (raw.float() - ref).abs().max().item() < 1e-2 * scale
    f = raw.float().reshape(-1, cout)
    got = part.sum(0)
    assert torch.allclose(got[:cout], f.sum(0), rtol=1e-3, atol=0.5)
    assert torch.allclose(got[cout:], (f * f).sum(0), rtol=1e-3, atol=0.5)


@pytest.mark.parametrize("n,h,w,cin,cout", [(32, 16, 16, 1024, 256),     # dec4: 4 phases, n_tile 256
                                             (16, 64, 64, 256, 64),       # dec6: phase-fused tile
                                             (16, 32, 32, 512, 128),      # dec5
                                             (19, 32, 32, 128, 64)])      # fused, odd tile count
def test_pair_convT_fprop(n, h, w, cin, cout):
    from pai_b200 import ops
    x = _rand((n, h, w, cin), 3)
    wt = _rand((cin, cout, 4, 4), 4, 0.05)
    bias = torch.randn(cout, device="cuda")
    wp = ops.pack_convT_weight(wt.float())
    ref = _nhwc(F.conv_transpose2d(_nchw(x), wt.float(), bias, stride=2, padding=1))
    scale = max(1.0, ref.abs().max().item())
    y = ops.convT4x4s2_fprop(x, wp, cout, bias=bias, act=ops.ACT_RELU)
    assert (y.float() - F.relu(ref)).abs().max().item() < 1e-2 * scale
    raw, part = ops.convT4x4s2_fprop_bnstats(x, wp, cout, bias=bias)
    assert (raw.float() - ref).abs().max().item() < 1e-2 * scale
    f = raw.float().reshape(-1, cout)
    got = part.sum(0)
    assert torch.allclose(got[:cout], f.sum(0), rtol=1e-3, atol=0.5)
    assert torch.allclose(got[cout:], (f * f).sum(0), rtol=1e-3, atol=0.5)
    # into a concat slot of a wider buffer
    buf = torch.zeros(n, 2 * h, 2 * w, 2 * cout, dtype=torch.bfloat16, device="cuda")
    ops.convT4x4s2_fprop(x, wp, cout, bias=bias, out=buf[..., :cout])
    assert (buf[..., :cout].float() - ref).abs().max().item() < 1e-2 * scale and buf[..., cout:].abs().max().item() == 0


@pytest.mark.parametrize("n,h,w,cout,cin", [(16, 64, 64, 128, 64), (16, 32, 32, 256, 128), (40, 16, 16, 512, 256)])
def test_pair_dgrad_with_fused_activation_backward(n, h, w, cout, cin):
    from pai_b200 import ops
    gy = _rand((n, h, w, cout), 41)
    wt = torch.randn(cout, cin, 4, 4, device="cuda") * 0.05
    wd = ops.pack_convT_weight(wt)
    saved = F.leaky_relu(_rand((n, 2 * h, 2 * w, cin), 42).float(), 0.2).bfloat16()
    ref = _nhwc(F.conv_transpose2d(_nchw(gy), wt.bfloat16().float(), None, stride=2, padding=1))
    want = ref * torch.where(saved.float() > 0, 1.0, 0.2)
    got, part = ops.conv4x4_dgrad_act(gy, wd, cin, saved, slope=0.2, want_colsum=True)
    scale = max(1.0, want.abs().max().item())
    assert (got.float() - want).abs().max().item() < 1e-2 * scale
    sums = got.float().reshape(-1, cin).sum(0)
    assert torch.allclose(part.sum(0)[:cin], sums, rtol=2e-2, atol=1.0)
