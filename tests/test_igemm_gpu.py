"""GPU parity of the tcgen05 implicit-GEMM kernels (through the C-ABI) against PyTorch's own
fp32 convolution of the same bf16-rounded operands -- the arithmetic the reference's nn.Conv2d /
nn.ConvTranspose2d perform (models/pix2pix.py:63-69,99-105; models/wrapper.py:229-233)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _ops():
    from pai_b200 import ops
    return ops


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).bfloat16()


def _nchw(t):
    return t.float().permute(0, 3, 1, 2)


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


CONV_CASES = [
    # n, h, w, cin, cout, stride
    (2, 128, 128, 64, 128, 2),    # enc1
    (2, 16, 16, 512, 512, 2),     # enc4
    (3, 4, 4, 512, 512, 2),       # enc6
    (2, 2, 2, 512, 512, 2),       # enc7 (1x1 output, zero padding on every side)
    (2, 64, 64, 128, 256, 2),     # enc2 / D2
    (2, 16, 16, 512, 1, 1),       # D4: stride 1, cout 1 -> 15x15
    (1, 32, 32, 64, 64, 2),
]


@pytest.mark.parametrize("n,h,w,cin,cout,stride", CONV_CASES)
def test_conv_fprop(n, h, w, cin, cout, stride):
    ops = _ops()
    x = _rand((n, h, w, cin), 1)
    wt = _rand((cout, cin, 4, 4), 2, 0.05)
    bias = torch.randn(cout, device="cuda")
    wp = ops.pack_conv_weight(wt.float())
    y = ops.conv4x4_fprop(x, wp, cout, stride=stride, bias=bias, out_f32=True)
    ref = _nhwc(F.conv2d(_nchw(x), wt.float(), bias, stride=stride, padding=1))
    torch.cuda.synchronize()
    err = (y - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err
    # bf16 output + fused LeakyReLU
    y2 = ops.conv4x4_fprop(x, wp, cout, stride=stride, bias=bias, act=ops.ACT_LEAKY, slope=0.2)
    ref2 = F.leaky_relu(ref, 0.2)
    assert (y2.float() - ref2).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())


CONVT_CASES = [
    # n, h, w, cin, cout
    (2, 1, 1, 512, 512),      # dec0
    (2, 4, 4, 1024, 512),     # dec2
    (2, 16, 16, 1024, 256),   # dec4
    (2, 64, 64, 256, 64),     # dec6
    (1, 128, 128, 128, 1),    # dec7 (cout 1 padded to 16, tanh)
]


@pytest.mark.parametrize("n,h,w,cin,cout", CONVT_CASES)
def test_convT_fprop(n, h, w, cin, cout):
    ops = _ops()
    x = _rand((n, h, w, cin), 3)
    wt = _rand((cin, cout, 4, 4), 4, 0.05)
    bias = torch.randn(cout, device="cuda")
    wp = ops.pack_convT_weight(wt.float())
    y = ops.convT4x4s2_fprop(x, wp, cout, bias=bias, out_f32=True)
    ref = _nhwc(F.conv_transpose2d(_nchw(x), wt.float(), bias, stride=2, padding=1))
    torch.cuda.synchronize()
    err = (y - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err
    y2 = ops.convT4x4s2_fprop(x, wp, cout, bias=bias, act=ops.ACT_TANH)
    assert (y2.float() - torch.tanh(ref)).abs().max().item() < 1e-2


def test_convT_into_concat_slot_and_channel_slice_input():
    """Zero-copy skip concat: the decoder writes into channels [0,C) of a 2C-wide buffer and the next
    decoder reads the whole buffer; a channel slice is also a valid input (pixel stride > channels)."""
    ops = _ops()
    n, h, w, cin, cout = 2, 8, 8, 128, 64
    x_wide = _rand((n, h, w, 2 * cin), 5)
    x = x_wide[..., cin:]                       # slice view, ld = 2*cin
    wt = _rand((cin, cout, 4, 4), 6, 0.05)
    wp = ops.pack_convT_weight(wt.float())
    buf = torch.zeros(n, 2 * h, 2 * w, 2 * cout, dtype=torch.bfloat16, device="cuda")
    ops.convT4x4s2_fprop(x, wp, cout, out=buf[..., :cout])
    ref = _nhwc(F.conv_transpose2d(_nchw(x), wt.float(), None, stride=2, padding=1))
    assert (buf[..., :cout].float() - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())
    assert buf[..., cout:].abs().max().item() == 0


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 128, 128, 64, 128), (2, 16, 16, 512, 512), (4, 2, 2, 512, 512),
                                             (2, 32, 32, 256, 512)])
def test_conv_dgrad_is_convT_fprop(n, h, w, cin, cout):
    """dL/dx of Conv2d(4,2,1) == ConvT fprop of dL/dy with the conv weight read as [in=Cout,out=Cin]."""
    ops = _ops()
    gy = _rand((n, h // 2, w // 2, cout), 7)
    wt = _rand((cout, cin, 4, 4), 8, 0.05)
    wp = ops.pack_convT_weight(wt.float())
    gx = ops.convT4x4s2_fprop(gy, wp, cin, out_f32=True)
    ref = _nhwc(F.conv_transpose2d(_nchw(gy), wt.float(), None, stride=2, padding=1))
    assert (gx - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("n,h,w,cin,cout,splitk", [(2, 128, 128, 64, 128, 0), (2, 16, 16, 512, 512, 0),
                                                    (4, 2, 2, 512, 512, 1), (2, 64, 64, 128, 256, 3)])
def test_conv_wgrad(n, h, w, cin, cout, splitk):
    ops = _ops()
    x = _rand((n, h, w, cin), 9)
    gy = _rand((n, h // 2, w // 2, cout), 10)
    dw = ops.conv4x4_wgrad(x, gy, stride=2, splitk=splitk)
    got = dw.permute(1, 2, 0).reshape(cout, cin, 4, 4)
    xr = _nchw(x).requires_grad_(False)
    wref = torch.zeros(cout, cin, 4, 4, device="cuda", requires_grad=True)
    F.conv2d(xr, wref, None, stride=2, padding=1).backward(_nchw(gy))
    ref = wref.grad
    assert (got - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 1, 1, 512, 512), (2, 16, 16, 1024, 256), (2, 64, 64, 256, 64)])
def test_convT_wgrad(n, h, w, cin, cout):
    ops = _ops()
    x = _rand((n, h, w, cin), 11)
    gy = _rand((n, 2 * h, 2 * w, cout), 12)
    dw = ops.convT4x4s2_wgrad(x, gy)
    got = dw.permute(1, 2, 0).reshape(cin, cout, 4, 4)
    wref = torch.zeros(cin, cout, 4, 4, device="cuda", requires_grad=True)
    F.conv_transpose2d(_nchw(x), wref, None, stride=2, padding=1).backward(_nchw(gy))
    ref = wref.grad
    assert (got - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())
