"""GPU unit parity of the BatchNorm / activation / degenerate-layer kernels (C-ABI through
pai_b200.ops) against plain PyTorch fp32 formulations of the same reference ops
(nn.BatchNorm2d, LeakyReLU/ReLU, Conv2d / ConvTranspose2d with 1-2 channel operands)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False


def _ops():
    from pai_b200 import ops
    return ops


def _rand(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype)


@pytest.mark.parametrize("n,h,w,c", [(2, 64, 64, 128), (8, 2, 2, 512), (2, 128, 128, 64), (3, 5, 7, 256)])
def test_batchnorm_forward_backward(n, h, w, c):
    ops = _ops()
    x = _rand((n, h, w, c), 1, 2.0) + 0.5
    gamma = torch.rand(c, device="cuda") + 0.5
    beta = torch.randn(c, device="cuda")
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    m = n * h * w
    sums = ops.bn_stats(x)
    ss = ops.bn_finalize(sums, m, c, gamma, beta, rm, rv, training=True)
    wide = torch.zeros(n, h, w, 2 * c, dtype=torch.bfloat16, device="cuda")
    o1 = torch.empty(n, h, w, c, dtype=torch.bfloat16, device="cuda")
    ops.bn_apply_act(x, ss, o1, ops.ACT_LEAKY, wide[..., c:], ops.ACT_RELU, slope=0.2)
    # reference
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    rm_r, rv_r = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    z = F.batch_norm(xr, rm_r, rv_r, gr, br, True, 0.1, 1e-5)
    y1, y2 = F.leaky_relu(z, 0.2), F.relu(z)
    tol = 2e-2 * max(1.0, z.abs().max().item())
    assert (o1.float().permute(0, 3, 1, 2) - y1).abs().max().item() < tol
    assert (wide[..., c:].float().permute(0, 3, 1, 2) - y2).abs().max().item() < tol
    assert wide[..., :c].abs().max().item() == 0
    assert torch.allclose(rm, rm_r, atol=1e-4, rtol=1e-4) and torch.allclose(rv, rv_r, atol=1e-4, rtol=1e-3)
    # backward
    g1 = _rand((n, h, w, c), 2)
    g2w = _rand((n, h, w, 2 * c), 3)
    g2 = g2w[..., c:]
    (y1 * g1.float().permute(0, 3, 1, 2) + y2 * g2.float().permute(0, 3, 1, 2)).sum().backward()
    s2 = ops.bn_bwd_reduce(x, ss, g1, ops.ACT_LEAKY, g2, ops.ACT_RELU, slope=0.2)
    dx = torch.empty_like(x)
    ops.bn_bwd_apply(x, ss, g1, ops.ACT_LEAKY, g2, ops.ACT_RELU, s2, gamma, dx, slope=0.2)
    ref = xr.grad.permute(0, 2, 3, 1)
    err = (dx.float() - ref).abs()
    tol = 2e-2 * max(1.0, ref.abs().max().item())
    # an activation input within one ulp of 0 may land on the other side of the kink: allow isolated outliers
    assert (err > tol).float().mean().item() < 1e-5 and err.mean().item() < 2e-3 * max(1.0, ref.abs().max().item())
    assert torch.allclose(s2[:c], br.grad, rtol=2e-2, atol=2e-2 * br.grad.abs().max().item())
    assert torch.allclose(s2[c:], gr.grad, rtol=2e-2, atol=2e-2 * gr.grad.abs().max().item())
    # eval mode uses the running statistics
    ss_e = ops.bn_finalize(None, m, c, gamma, beta, rm, rv, training=False)
    ops.bn_apply_act(x, ss_e, o1, ops.ACT_NONE)
    ze = F.batch_norm(x.float().permute(0, 3, 1, 2), rm, rv, gamma, beta, False, 0.1, 1e-5)
    assert (o1.float().permute(0, 3, 1, 2) - ze).abs().max().item() < 2e-2 * max(1.0, ze.abs().max().item())


def test_activation_backward_without_norm_and_colsum():
    ops = _ops()
    x = _rand((2, 16, 16, 64), 4)
    g = _rand((2, 16, 16, 64), 5)
    s = ops.bn_bwd_reduce(x, None, g, ops.ACT_LEAKY, slope=0.2)
    dx = torch.empty_like(x)
    ops.bn_bwd_apply(x, None, g, ops.ACT_LEAKY, None, ops.ACT_NONE, s, None, dx, slope=0.2)
    ref = g.float() * torch.where(x.float() > 0, 1.0, 0.2)
    assert (dx.float() - ref).abs().max().item() < 1e-2 * ref.abs().max().item()
    assert torch.allclose(s[:64], ref.sum((0, 1, 2)), rtol=1e-3, atol=1e-2)
    assert torch.allclose(ops.colsum(x), x.float().sum((0, 1, 2)), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("cin,c,act", [(1, 64, 1), (2, 64, 1), (1, 128, 0)])
def test_smallc_conv_fprop_and_wgrad(cin, c, act):
    ops = _ops()
    n, h, w = 2, 64, 48
    planes = [torch.randn(n, h, w, device="cuda") for _ in range(cin)]
    wt = torch.randn(c, cin, 4, 4, device="cuda") * 0.1
    bias = torch.randn(c, device="cuda")
    wtm = wt.permute(0, 2, 3, 1).reshape(c, 16, cin).contiguous()
    o1 = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device="cuda")
    o2 = torch.empty(n, h // 2, w // 2, 2 * c, dtype=torch.bfloat16, device="cuda")
    ops.smallc_conv_fprop(planes, wtm, bias, o1, act, o2[..., c:], ops.ACT_RELU, stride=2)
    xin = torch.stack(planes, 1)
    ref = F.conv2d(xin, wt, bias, stride=2, padding=1)
    r1 = F.leaky_relu(ref, 0.2) if act == 1 else ref
    assert (o1.float().permute(0, 3, 1, 2) - r1).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())
    assert (o2[..., c:].float().permute(0, 3, 1, 2) - F.relu(ref)).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())
    # wgrad
    gy = _rand((n, h // 2, w // 2, c), 7)
    dw = ops.smallc_conv_wgrad(gy, planes, stride=2)
    wr = wt.clone().requires_grad_(True)
    F.conv2d(xin, wr, None, stride=2, padding=1).backward(gy.float().permute(0, 3, 1, 2))
    got = dw.view(c, 4, 4, cin).permute(0, 3, 1, 2)
    assert (got - wr.grad).abs().max().item() < 2e-3 * max(1.0, wr.grad.abs().max().item())


def test_smallc_as_convT_cout1_gradients():
    """dec7: ConvTranspose2d(128, 1, 4, 2, 1) data / weight gradient."""
    ops = _ops()
    n, h, w, cin = 2, 32, 32, 128
    x = _rand((n, h, w, cin), 8)
    wt = (torch.randn(cin, 1, 4, 4, device="cuda") * 0.1).requires_grad_(True)
    g = torch.randn(n, 2 * h, 2 * w, device="cuda")
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.conv_transpose2d(xr, wt, None, stride=2, padding=1).backward(g.view(n, 1, 2 * h, 2 * w))
    dx = torch.empty_like(x)
    ops.smallc_conv_fprop([g], wt.detach().reshape(cin, 16, 1), None, dx, ops.ACT_NONE, stride=2)
    ref = xr.grad.permute(0, 2, 3, 1)
    assert (dx.float() - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())
    dw = ops.smallc_conv_wgrad(x, [g], stride=2)
    assert (dw.view(cin, 1, 4, 4) - wt.grad).abs().max().item() < 2e-3 * max(1.0, wt.grad.abs().max().item())


def test_smallc_as_stride1_cout1_gradients():
    """D4: Conv2d(512, 1, 4, 1, 1, bias=False) data / weight gradient (flipped taps)."""
    ops = _ops()
    n, h, w, cin = 3, 16, 16, 512
    x = _rand((n, h, w, cin), 9)
    wt = (torch.randn(1, cin, 4, 4, device="cuda") * 0.1).requires_grad_(True)
    g = torch.randn(n, h - 1, w - 1, device="cuda")
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.conv2d(xr, wt, None, stride=1, padding=1).backward(g.view(n, 1, h - 1, w - 1))
    dx = torch.empty_like(x)
    ops.smallc_conv_fprop([g], wt.detach()[0].reshape(cin, 16, 1), None, dx, ops.ACT_NONE, stride=1, flip=True)
    ref = xr.grad.permute(0, 2, 3, 1)
    assert (dx.float() - ref).abs().max().item() < 1e-2 * max(1.0, ref.abs().max().item())
    dw = ops.smallc_conv_wgrad(x, [g], stride=1, flip=True)
    assert (dw.view(cin, 4, 4, 1).permute(3, 0, 1, 2) - wt.grad).abs().max().item() < 2e-3 * max(1.0, wt.grad.abs().max().item())


def test_thin_conv_as_im2col_plus_pointwise_gemm():
    """enc0 / D0: Conv2d(1|2 -> 64, 4, 2, 1) == im2col4x4 + pointwise GEMM (two fused outputs), and its wgrad."""
    ops = _ops()
    n, h, w, c = 2, 64, 64, 64
    for cin in (1, 2):
        planes = [torch.randn(n, h, w, device="cuda") for _ in range(cin)]
        wt = torch.randn(c, cin, 4, 4, device="cuda") * 0.1
        bias = torch.randn(c, device="cuda")
        col = ops.im2col4x4(planes, h // 2, w // 2, stride=2)
        wp = torch.zeros(c, 64, dtype=torch.bfloat16, device="cuda")
        wp[:, :16 * cin] = wt.permute(0, 2, 3, 1).reshape(c, -1)
        o2 = torch.zeros(n, h // 2, w // 2, 2 * c, dtype=torch.bfloat16, device="cuda")
        o1 = ops.pointwise_gemm(col, wp, c, bias=bias, act=ops.ACT_LEAKY, out2=o2[..., c:], act2=ops.ACT_NONE,
                                k_valid=16 * cin)        # im2col leaves the padding columns unwritten
        xin = torch.stack(planes, 1).bfloat16().float()
        ref = F.conv2d(xin, wt.bfloat16().float(), bias, stride=2, padding=1)
        tol = 1e-2 * max(1.0, ref.abs().max().item())
        assert (o1.float().permute(0, 3, 1, 2) - F.leaky_relu(ref, 0.2)).abs().max().item() < tol
        assert (o2[..., c:].float().permute(0, 3, 1, 2) - ref).abs().max().item() < tol
        assert o2[..., :c].abs().max().item() == 0
        gy = _rand((n, h // 2, w // 2, c), 21)
        dw = ops.pointwise_wgrad(gy, col)
        wr = wt.clone().requires_grad_(True)
        F.conv2d(xin, wr, None, stride=2, padding=1).backward(gy.float().permute(0, 3, 1, 2))
        got = dw[:, :16 * cin].reshape(c, 4, 4, cin).permute(0, 3, 1, 2)
        assert (got - wr.grad).abs().max().item() < 3e-3 * max(1.0, wr.grad.abs().max().item())


def test_thin_convT_as_pointwise_gemm_plus_col2im():
    """dec7: ConvTranspose2d(128 -> 1, 4, 2, 1) + Tanh forward, data gradient and weight gradient."""
    ops = _ops()
    n, h, w, cin = 2, 32, 32, 128
    x = _rand((n, h, w, cin), 22)
    wt = (torch.randn(cin, 1, 4, 4, device="cuda") * 0.05)
    bias = torch.randn(1, device="cuda") * 0.1
    wq = wt.bfloat16().float().requires_grad_(True)
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    ref = torch.tanh(F.conv_transpose2d(xr, wq, bias, stride=2, padding=1))
    part = ops.pointwise_gemm(x, wt[:, 0].reshape(cin, 16).t().contiguous().bfloat16(), 16, out_f32=True)
    y = ops.col2im4x4s2(part, bias, ops.ACT_TANH)
    assert (y - ref[:, 0]).abs().max().item() < 2e-3
    g = torch.randn(n, 2 * h, 2 * w, device="cuda")
    F.conv_transpose2d(xr, wq, None, stride=2, padding=1).backward(g.view(n, 1, 2 * h, 2 * w))
    gcol = ops.im2col4x4([g], h, w, stride=2)
    wd = torch.zeros(cin, 64, dtype=torch.bfloat16, device="cuda")
    wd[:, :16] = wt[:, 0].reshape(cin, 16)
    dx = ops.pointwise_gemm(gcol, wd, cin, k_valid=16)
    refdx = xr.grad.permute(0, 2, 3, 1)
    assert (dx.float() - refdx).abs().max().item() < 2e-2 * max(1.0, refdx.abs().max().item())
    dw = ops.pointwise_wgrad(x, gcol)[:, :16].reshape(cin, 1, 4, 4)
    assert (dw - wq.grad).abs().max().item() < 1e-2 * max(1.0, wq.grad.abs().max().item())


@pytest.mark.parametrize("a,b", [(128, 64), (512, 256), (1, 512), (96, 40)])
def test_fused_adam_matches_torch_and_repacks(a, b):
    """FusedAdam (pai_adam_pack_conv4x4 + pai_adam_multi) vs torch.optim.Adam with the reference's
    hyper-parameters (models/wrapper.py:98-111), and the packs it rewrites vs the Python packers."""
    from pai_b200 import engine, ops
    from pai_b200.optim import FusedAdam
    torch.manual_seed(3)
    hyper = dict(lr=2e-4, betas=(0.5, 0.999), eps=1e-7)
    w = torch.nn.Parameter((torch.randn(a, b, 4, 4, device="cuda") * 0.02))
    bias = torch.nn.Parameter(torch.randn(a, device="cuda"))
    thin = torch.nn.Parameter(torch.randn(64, 1, 4, 4, device="cuda") * 0.02)
    ref = [torch.nn.Parameter(t.detach().clone()) for t in (w, bias, thin)]
    opt, opt_ref = FusedAdam([w, bias, thin], **hyper), torch.optim.Adam(ref, **hyper)
    # populate the pack cache the way the engine does
    engine._fprop_pack(w)
    if a % 8 == 0:
        engine._dgrad_pack(w)
    engine._thin_in_pack(thin)
    for it in range(3):
        for p, r in zip((w, bias, thin), ref):
            g = torch.randn_like(p) * (10.0 ** (-it))
            p.grad, r.grad = g.clone(), g.clone()
        opt.step()
        opt_ref.step()
        for p, r in zip((w, bias, thin), ref):
            assert torch.allclose(p, r, rtol=1e-5, atol=1e-7), (it, (p - r).abs().max())
        assert torch.equal(engine._fprop_pack(w), ops.pack_conv_weight(w.detach()))
        if a % 8 == 0:
            assert torch.equal(engine._dgrad_pack(w), ops.pack_convT_weight(w.detach()))
        assert torch.equal(engine._thin_in_pack(thin), engine._pad_cols(thin.detach().permute(0, 2, 3, 1).reshape(64, -1)))
    sd, sd_ref = opt.state_dict(), opt_ref.state_dict()
    assert sd["state"].keys() == sd_ref["state"].keys()
    for k in sd["state"]:
        assert torch.allclose(sd["state"][k]["exp_avg_sq"], sd_ref["state"][k]["exp_avg_sq"], rtol=1e-4, atol=1e-12)
        assert float(sd["state"][k]["step"]) == float(sd_ref["state"][k]["step"])


@pytest.mark.parametrize("with_g2", [False, True])
def test_act_bwd_fused_bias_gradient(with_g2):
    """pai_act_bwd (layers without BatchNorm) == separate reduce + apply == autograd of LeakyReLU / identity."""
    ops = _ops()
    n, h, w, c = 2, 16, 16, 64
    pre = _rand((n, h, w, c), 11)
    x = F.leaky_relu(pre.float(), 0.2).bfloat16()          # the saved activation has the sign of the pre-activation
    g1, g2 = _rand((n, h, w, c), 12), (_rand((n, h, w, 2 * c), 13)[..., c:] if with_g2 else None)
    dx = torch.empty(n, h, w, c, dtype=torch.bfloat16, device="cuda")
    sums = ops.act_bwd(x, g1, ops.ACT_LEAKY, g2, ops.ACT_NONE, dx, slope=0.2)
    want = g1.float() * torch.where(pre.float() > 0, 1.0, 0.2)
    if with_g2:
        want = want + g2.float()
    assert (dx.float() - want).abs().max().item() < 2e-2
    assert torch.allclose(sums[:c], want.reshape(-1, c).sum(0), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("kind,n,h,w,cin,cout", [("conv", 16, 64, 64, 64, 128), ("conv", 4, 64, 96, 128, 256),
                                                 ("convT", 8, 32, 32, 256, 64), ("convT", 4, 32, 48, 512, 128)])
def test_bn_statistics_from_the_gemm_epilogue(kind, n, h, w, cin, cout):
    """pai_conv4x4_fprop_bnstats / pai_convT4x4s2_fprop_bnstats: same raw output as the plain convolution, and the
    per-CTA partial sums add up to what pai_bn_stats computes from that output (sum, sum of squares per channel);
    convT 256 -> 64 takes the phase-fused tile path."""
    ops = _ops()
    x = _rand((n, h, w, cin), 31)
    bias = torch.randn(cout, device="cuda")
    if kind == "conv":
        wt = torch.randn(cout, cin, 4, 4, device="cuda") * 0.05
        wp = ops.pack_conv_weight(wt)
        ref = ops.conv4x4_fprop(x, wp, cout, stride=2, bias=bias)
        raw, part = ops.conv4x4_fprop_bnstats(x, wp, cout, bias=bias)
    else:
        wt = torch.randn(cin, cout, 4, 4, device="cuda") * 0.05
        wp = ops.pack_convT_weight(wt)
        ref = ops.convT4x4s2_fprop(x, wp, cout, bias=bias)
        raw, part = ops.convT4x4s2_fprop_bnstats(x, wp, cout, bias=bias)
    assert torch.equal(raw, ref)
    want = ops.bn_stats(ref)
    got = part.sum(0)
    assert torch.allclose(got[:cout], want[:cout], rtol=1e-4, atol=1e-2), (got[:cout] - want[:cout]).abs().max()
    assert torch.allclose(got[cout:], want[cout:], rtol=1e-4, atol=1e-2)
    f = ref.float().reshape(-1, cout)
    assert torch.allclose(got[:cout], f.sum(0), rtol=1e-3, atol=5e-2)
    assert torch.allclose(got[cout:], (f * f).sum(0), rtol=1e-3, atol=5e-2)


@pytest.mark.parametrize("n,h,w,cout,cin", [(8, 32, 32, 128, 64), (4, 32, 32, 256, 128), (4, 16, 48, 512, 256)])
def test_dgrad_with_fused_activation_backward(n, h, w, cout, cin):
    """pai_conv4x4_dgrad_act == pai_convT4x4s2_fprop (data gradient) followed by pai_act_bwd (LeakyReLU backward +
    bias-gradient column sums) of the layer below; cin = 64 takes the phase-fused tile path."""
    ops = _ops()
    gy = _rand((n, h, w, cout), 41)
    wt = torch.randn(cout, cin, 4, 4, device="cuda") * 0.05          # Conv2d weight of the layer being differentiated
    wd = ops.pack_convT_weight(wt)                                     # its data-gradient operand
    saved = F.leaky_relu(_rand((n, 2 * h, 2 * w, cin), 42).float(), 0.2).bfloat16()
    dh = ops.convT4x4s2_fprop(gy, wd, cin)
    want = torch.empty_like(dh)
    want_sums = ops.act_bwd(saved, dh, ops.ACT_LEAKY, None, ops.ACT_NONE, want, slope=0.2)
    got, part = ops.conv4x4_dgrad_act(gy, wd, cin, saved, slope=0.2, want_colsum=True)
    # the unfused path rounds the data gradient to bf16 before masking, the fused one masks the fp32 accumulator
    assert (got.float() - want.float()).abs().max().item() <= 2e-2 * max(1.0, want.float().abs().max().item())
    assert torch.allclose(part.sum(0)[:cin], want_sums[:cin], rtol=2e-2, atol=0.5)
    got2, none = ops.conv4x4_dgrad_act(gy, wd, cin, saved, slope=0.2, want_colsum=False)
    assert none is None and torch.equal(got2, got)


def test_dropout2d_layer_node_scales_whole_channels():
    """layers.dropout2d (Attention / Res U-Net decoders): train mode zeroes whole (sample, channel) planes with
    probability p and scales the rest by 1 / (1 - p), the backward applies the same mask; eval mode is the identity."""
    from pai_b200 import layers as L
    torch.manual_seed(0)
    x = _rand((4, 8, 8, 64), 51).requires_grad_(True)
    mod = torch.nn.Dropout2d(0.5).train()
    y = L.dropout2d(x, mod)
    ratio = (y.float() / x.detach().float()).reshape(4, 64, 64)            # [n, pixels, c]
    per_plane = ratio.mean(1)
    assert torch.all((per_plane.abs() < 1e-6) | ((per_plane - 2.0).abs() < 2e-2)), per_plane
    kept = per_plane.abs() > 1e-6
    assert 0.2 < kept.float().mean().item() < 0.8
    assert torch.all((ratio - per_plane[:, None, :]).abs() < 2e-2)          # one factor per (sample, channel)
    y.backward(torch.ones_like(y))
    g = x.grad.float().reshape(4, 64, 64).mean(1)
    assert torch.allclose(g, per_plane.round(), atol=1e-2)
    assert L.dropout2d(x, mod.eval()) is x


@pytest.mark.parametrize("n,h,w,c,drop", [(64, 8, 8, 512, True), (8, 2, 2, 512, False), (3, 5, 7, 256, True), (64, 4, 4, 64, False)])
def test_small_layer_batchnorm_one_launch_equals_the_separate_kernels(n, h, w, c, drop):
    """pai_bn_small_fwd / pai_bn_small_bwd (one block per 8 channels, everything in one launch) against the chain they
    replace: bn_stats -> bn_finalize -> bn_apply_act (-> scale_channels) and scale_channels -> bn_bwd_reduce ->
    bn_bwd_apply.  Same arithmetic, different summation order: outputs equal to a bf16 ulp, sums to fp32 rounding."""
    ops = _ops()
    x = _rand((n, h, w, c), 11, 2.0) + 0.5
    gamma = torch.rand(c, device="cuda") + 0.5
    beta = torch.randn(c, device="cuda")
    mask = ops.dropout2d_mask(n, c, 0.5, x.device) if drop else None
    m = n * h * w
    assert ops.bn_small_ok(x, 3)
    # separate kernels
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    ss = ops.bn_finalize(ops.bn_stats(x), m, c, gamma, beta, rm, rv, training=True)
    wide = torch.zeros(n, h, w, 2 * c, dtype=torch.bfloat16, device="cuda")
    o1 = torch.empty_like(x)
    ops.bn_apply_act(x, ss, o1, ops.ACT_LEAKY, wide[..., c:], ops.ACT_RELU, slope=0.2)
    if drop:
        ops.scale_channels(o1, mask, out=o1)
    # one launch
    rm2, rv2 = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    wide2 = torch.zeros_like(wide)
    p1 = torch.empty_like(x)
    ss2 = ops.bn_small_fwd(x, gamma, beta, rm2, rv2, p1, ops.ACT_LEAKY, wide2[..., c:], ops.ACT_RELU, slope=0.2, mask=mask)
    assert torch.allclose(ss2, ss, rtol=1e-4, atol=1e-5)
    assert torch.allclose(rm2, rm, rtol=1e-5, atol=1e-6) and torch.allclose(rv2, rv, rtol=1e-5, atol=1e-6)
    for got, want in ((p1, o1), (wide2, wide)):
        d = (got.float() - want.float()).abs()
        assert (d > 2.0 ** -7 * want.float().abs()).float().mean().item() < 1e-4      # at most a bf16 ulp, and rarely
        assert d.max().item() <= 2.0 ** -6 * want.float().abs().max().item()
    # backward: two gradients (encoder) and one gradient + dropout (decoder)
    g1 = _rand((n, h, w, c), 12)
    g2 = _rand((n, h, w, 2 * c), 13)[..., c:]
    for two in (True, False):
        ga = g1.clone()
        if not two and drop:
            ops.scale_channels(ga, mask, out=ga)
        a2, t2 = (g2, ops.ACT_RELU) if two else (None, ops.ACT_NONE)
        s_ref = ops.bn_bwd_reduce(x, ss, ga, ops.ACT_LEAKY, a2, t2, slope=0.2)
        dx_ref = torch.empty_like(x)
        ops.bn_bwd_apply(x, ss, ga, ops.ACT_LEAKY, a2, t2, s_ref, gamma, dx_ref, slope=0.2)
        dx = torch.empty_like(x)
        s_got = ops.bn_small_bwd(x, ss, g1, ops.ACT_LEAKY, a2, t2, gamma, dx, slope=0.2, mask=None if two else mask)
        scale = s_ref.abs().max().item()
        assert torch.allclose(s_got, s_ref, rtol=1e-3, atol=1e-4 * scale)
        d = (dx.float() - dx_ref.float()).abs()
        assert d.max().item() <= 2.0 ** -6 * dx_ref.float().abs().max().item() + 1e-6
