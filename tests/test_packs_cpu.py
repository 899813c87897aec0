"""Host-side operand packers (pai_b200/ops.py, pai_b200/layers.py) checked on CPU: the GEMM each pack is built for is
emulated with plain fp32 tensor arithmetic (same index math as the kernels: SURVEY.md Appendix B) and compared with
torch's Conv2d / ConvTranspose2d / grouped Conv2d -- so a layout mistake shows up without a GPU."""
import torch
import torch.nn.functional as F

from pai_b200 import layers, ops

_TD = ((0, -1), (1, 0))        # input offset d of tap t for output parity p (csrc/pai_api.cu kTd)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def test_conv4x4_stride2_pack_is_the_im2col_gemm_operand():
    torch.manual_seed(0)
    n, cin, cout, h, w = 2, 8, 24, 8, 12
    x, wt, b = torch.randn(n, cin, h, w), torch.randn(cout, cin, 4, 4), torch.randn(cout)
    pack = ops.pack_conv_weight(wt).float()                          # [cout_pad, 16*cin], column (ky*4+kx)*cin + ci
    assert pack.shape == (ops.padded_cout(cout), 16 * cin) and float(pack[cout:].abs().max()) == 0
    xp = F.pad(_nhwc(x), (0, 0, 1, 1, 1, 1))                         # zero padding 1 (the TMA out-of-bounds fill)
    cols = torch.stack([xp[:, ky:ky + h:2, kx:kx + w:2, :] for ky in range(4) for kx in range(4)], dim=3)
    y = cols.reshape(n, h // 2, w // 2, 16 * cin) @ pack[:cout].t() + b
    ref = _nhwc(F.conv2d(x, wt.bfloat16().float(), b, stride=2, padding=1))
    assert torch.allclose(y, ref, atol=1e-4)


def test_convT4x4_stride2_pack_four_sub_pixel_phases():
    torch.manual_seed(1)
    n, cin, cout, h, w = 2, 8, 16, 5, 6
    x, wt = torch.randn(n, cin, h, w), torch.randn(cin, cout, 4, 4)
    pack = ops.pack_convT_weight(wt).float()                         # [4 phases, cout_pad, 4*cin]
    xp = F.pad(_nhwc(x), (0, 0, 1, 1, 1, 1))
    y = torch.zeros(n, 2 * h, 2 * w, cout)
    for py in range(2):
        for px in range(2):
            taps = [xp[:, 1 + _TD[py][ty]:1 + _TD[py][ty] + h, 1 + _TD[px][tx]:1 + _TD[px][tx] + w, :]
                    for ty in range(2) for tx in range(2)]           # out[2a+py, 2b+px] reads in[a + d_y, b + d_x]
            cols = torch.cat(taps, dim=3)                            # [n, h, w, 4*cin], column (ty*2+tx)*cin + ci
            y[:, py::2, px::2, :] = cols @ pack[py * 2 + px, :cout].t()
    ref = _nhwc(F.conv_transpose2d(x, wt.bfloat16().float(), None, stride=2, padding=1))
    assert torch.allclose(y, ref, atol=1e-4)


def test_conv3x3_packs_forward_and_data_gradient():
    torch.manual_seed(2)
    n, cin, cout, h, w = 1, 8, 16, 6, 7
    x, wt = torch.randn(n, cin, h, w), torch.randn(cout, cin, 3, 3)
    pf = ops.pack_conv3x3_weight(wt).float()                         # [cout_pad, 9*cin]

    def conv3x3(inp_nhwc, pack, c_out):
        xp = F.pad(inp_nhwc, (0, 0, 1, 1, 1, 1))
        hh, ww = inp_nhwc.shape[1], inp_nhwc.shape[2]
        cols = torch.cat([xp[:, ky:ky + hh, kx:kx + ww, :] for ky in range(3) for kx in range(3)], dim=3)
        return cols @ pack[:c_out].t()

    y = conv3x3(_nhwc(x), pf, cout)
    wq = wt.bfloat16().float()
    assert torch.allclose(y, _nhwc(F.conv2d(x, wq, None, padding=1)), atol=1e-4)
    gy = torch.randn(n, cout, h, w)
    pd = ops.pack_conv3x3_weight_dgrad(wt).float()                   # [cin_pad, 9*cout]: flipped taps, in/out swapped
    gx = conv3x3(_nhwc(gy), pd, cin)
    xr = x.clone().requires_grad_(True)
    F.conv2d(xr, wq, None, padding=1).backward(gy)
    assert torch.allclose(gx, _nhwc(xr.grad), atol=1e-4)


def test_grouped_conv_as_block_diagonal_dense_blocks():
    torch.manual_seed(3)
    wt = torch.randn(128, 4, 3, 3)
    dense = layers._g4_dense_blocks(wt)                              # [2, 64, 64, 3, 3]
    x = torch.randn(2, 128, 6, 6)
    ref = F.conv2d(x, wt, padding=1, groups=32)
    got = torch.cat([F.conv2d(x[:, 64 * k:64 * k + 64], dense[k], padding=1) for k in range(2)], dim=1)
    assert torch.equal(ref, got) or torch.allclose(ref, got, atol=1e-5)
    blocks = dense.view(2, 16, 4, 16, 4, 3, 3).permute(0, 1, 3, 2, 4, 5, 6)          # [nb, g_out, g_in, 4, 4, 3, 3]
    mask = torch.arange(16)[:, None] != torch.arange(16)[None, :]
    assert float(blocks[:, mask].abs().max()) == 0                                    # only the group diagonal is set


def test_padded_carriers_keep_the_padding_zero():
    torch.manual_seed(4)
    wt, b = torch.randn(16, 64, 1, 1), torch.randn(16)
    wp = layers._pack_p1_f(wt, 64).float()                           # [pad64(16), 64]
    assert wp.shape == (64, 64) and float(wp[16:].abs().max()) == 0
    assert float(layers._pad_b(b, 64)[16:].abs().max()) == 0
    w3 = layers._pad_w(torch.randn(16, 16, 3, 3), 64)
    assert w3.shape == (64, 64, 3, 3) and float(w3[16:].abs().max()) == 0 and float(w3[:, 16:].abs().max()) == 0
