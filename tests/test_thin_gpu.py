"""Single-pass kernels of the degenerate layers (csrc/thin.cu) against torch's fp32 convolutions of the SAME bf16-rounded
operands (the kernels multiply bf16 x bf16 exactly and accumulate in fp32, so only the summation order and the bf16
rounding of the stored output differ).  Reference layers: enc0 models/pix2pix.py:141-147, D0 models/wrapper.py:229,
dec7 models/pix2pix.py:186-195, PatchGAN head models/wrapper.py:233."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ACT_NONE, ACT_LEAKY, ACT_RELU, ACT_TANH = 0, 1, 2, 3


def _q(t):
    return t.bfloat16().float()


def _thin_in_pack(w):          # the engine's packer: [C, cin, 4, 4] -> bf16 [C, 64], column = tap*cin + j
    from pai_b200 import engine
    return engine._pad_cols(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1))


@pytest.mark.parametrize("cin,cout,n,size,two", [(1, 64, 3, 64, True), (2, 64, 2, 256, False), (1, 128, 2, 256, False),
                                                 (2, 64, 5, 24, True), (1, 64, 64, 256, True)])
def test_thin_conv_fprop(cin, cout, n, size, two):
    from pai_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(cin * 100 + cout + n)
    planes = [torch.randn(n, size, size, device="cuda", generator=g) for _ in range(cin)]
    w = torch.randn(cout, cin, 4, 4, device="cuda", generator=g) * 0.1
    b = torch.randn(cout, device="cuda", generator=g)
    o = size // 2
    # outputs as channel slots of wider buffers (ld > cout), like the concat buffers of the U-Net
    buf1 = torch.full((n, o, o, cout + 8), 7.0, dtype=torch.bfloat16, device="cuda")
    buf2 = torch.full((n, o, o, 2 * cout), 9.0, dtype=torch.bfloat16, device="cuda")
    out1 = buf1[..., :cout]
    out2 = buf2[..., cout:] if two else None
    ops.thin_conv_fprop(planes, _thin_in_pack(w), cout, b, out1, ACT_LEAKY, out2, ACT_NONE, slope=0.2)
    x = torch.stack(planes, 1)
    ref = F.conv2d(_q(x), _q(w), b, stride=2, padding=1).permute(0, 2, 3, 1)
    want1 = F.leaky_relu(ref, 0.2)
    assert torch.allclose(out1.float(), want1, rtol=1e-2, atol=1e-2), (out1.float() - want1).abs().max()
    assert (buf1[..., cout:] == 7.0).all()
    if two:
        assert torch.allclose(out2.float(), ref, rtol=1e-2, atol=1e-2), (out2.float() - ref).abs().max()
        assert (buf2[..., :cout] == 9.0).all()


@pytest.mark.parametrize("cin,c,n,size", [(1, 64, 2, 256), (2, 64, 3, 256), (1, 128, 2, 256), (1, 64, 64, 256), (2, 64, 1, 128)])
def test_thin_conv_wgrad(cin, c, n, size):
    from pai_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(cin * 7 + c + n)
    planes = [torch.randn(n, size, size, device="cuda", generator=g) for _ in range(cin)]
    o = size // 2
    wide = torch.randn(n, o, o, c + 64, device="cuda", generator=g).bfloat16()
    u = wide[..., :c]
    assert ops.thin_wgrad_ok(u, planes)
    dw = ops.thin_conv_wgrad(u, planes)                       # [c, 16*cin], column = tap*cin + j
    x = _q(torch.stack(planes, 1)).requires_grad_(False)
    wz = torch.zeros(c, cin, 4, 4, device="cuda", requires_grad=True)
    y = F.conv2d(x, wz, None, stride=2, padding=1)
    y.backward(u.float().permute(0, 3, 1, 2))
    want = wz.grad.permute(0, 2, 3, 1).reshape(c, 16 * cin)
    scale = want.abs().max().item()
    assert (dw - want).abs().max().item() < 2e-3 * scale + 1e-3, ((dw - want).abs().max().item(), scale)


@pytest.mark.parametrize("c,n,h,act,ld_extra", [(128, 2, 128, ACT_TANH, 0), (64, 3, 128, ACT_NONE, 64), (128, 64, 128, ACT_TANH, 0),
                                                (64, 1, 5, ACT_NONE, 0), (256, 2, 16, ACT_TANH, 0)])
def test_thin_convT_plane(c, n, h, act, ld_extra):
    from pai_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(c + n + h)
    wide = (torch.randn(n, h, 128, c + ld_extra, device="cuda", generator=g) * 0.5).bfloat16()
    x = wide[..., :c]
    assert ops.thin_plane_ok(x)
    w = torch.randn(c, 1, 4, 4, device="cuda", generator=g) * 0.05
    b = torch.randn(1, device="cuda", generator=g) * 0.1
    w_taps = w[:, 0].reshape(c, 16).t().contiguous().bfloat16()
    out = ops.thin_convT_plane(x, w_taps, b, act)
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), _q(w), b, stride=2, padding=1)[:, 0]
    if act == ACT_TANH:
        ref = torch.tanh(ref)
    assert out.shape == ref.shape
    assert torch.allclose(out, ref, rtol=1e-4, atol=2e-4), (out - ref).abs().max()


def test_head_conv_as_one_gemm_plus_gather():
    from pai_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    n, c = 5, 512
    x = (torch.randn(n, 16, 16, c, device="cuda", generator=g)).bfloat16()
    w = torch.randn(1, c, 4, 4, device="cuda", generator=g) * 0.02
    taps = w[0].reshape(c, 16).t().contiguous().bfloat16()
    part = ops.pointwise_gemm(x, taps, 16, out_f32=True)
    out = ops.col2im4x4s1(part)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), _q(w), None, stride=1, padding=1)[:, 0]
    assert out.shape == ref.shape == (n, 15, 15)
    assert torch.allclose(out, ref, rtol=1e-4, atol=5e-4), (out - ref).abs().max()
