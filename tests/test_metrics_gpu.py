"""GPU parity of the SSIM / PSNR / RMSE kernel (C-ABI pai_ssim_psnr_fwd / _bwd through
pai_b200.metrics) against the oracle: the torchmetrics-0.11.4 restatement (oracle/torchmetrics_port.py),
the independent fp64 C formulation (oracle/ssim_ref.c) and the fixtures of tests/golden/metrics_ref.npz.  Those fixtures
were produced by the reference's own call sites (models/utils.py:38-47, report.py:78-96,188-217) running ON that same
restatement (oracle/shim/torchmetrics; the real torchmetrics 0.11.4 cannot be installed here): they pin the kernel to the
port, NOT to torchmetrics itself -- the metric half of the oracle stays "parity unpinned" (DESIGN.md section 2), backed by
closed-form answers and the fp64 C formulation.  Tolerance for SSIM/PSNR: 1e-4 (BASELINE.json north_star)."""
import ctypes
import math
import os

import numpy as np
import pytest
import torch

import pix2pix_port as port
import torchmetrics_port as tm

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _m():
    from pai_b200 import metrics
    return metrics


def test_golden_report_metrics(golden_dir):
    gz = np.load(os.path.join(golden_dir, "metrics_ref.npz"))
    pred, tgt = port.synthetic_eval_pairs(8, seed=4321)
    r = _m().report_metrics(pred, tgt, want_maps=True)          # host tensors in, host tensors out
    assert r["ssim"].device.type == "cpu"
    assert np.abs(r["ssim"].numpy() - gz["ssim_per_image"]).max() < TOL
    assert np.abs(r["psnr"].numpy() - gz["psnr_per_image"]).max() < TOL
    assert np.allclose(r["mse"].numpy(), gz["mse_per_image"], rtol=1e-4)
    assert np.abs(r["depth_ssim"].numpy() - gz["depth_ssim"]).max() < TOL
    assert float(r["rmse"]) == pytest.approx(float(gz["rmse_global"]), rel=1e-4)
    d = np.abs(r["ssim_maps"][:, :, ::8, ::8].numpy() - gz["ssim_map_sub"])
    assert d.max() < 2e-3 and d.mean() < 2e-5          # per-pixel fp32 cancellation noise, see oracle tests
    m = _m()
    assert float(m.ssim(pred.cuda(), tgt.cuda())) == pytest.approx(float(gz["ssim"]), abs=TOL)
    assert float(m.psnr(pred.cuda(), tgt.cuda())) == pytest.approx(float(gz["psnr"]), abs=TOL)
    assert float(m.rmse(pred.cuda(), tgt.cuda())) == pytest.approx(float(gz["rmse"]), abs=1e-6)


def test_golden_train_metrics_with_denormalize(golden_dir):
    gz = np.load(os.path.join(golden_dir, "metrics_ref.npz"))
    xn, tn = port.synthetic_pairs(4, seed=77)
    s, p, r = _m().train_metrics(xn.cuda(), tn.cuda(), denormalize=True)
    assert float(s) == pytest.approx(float(gz["train_ssim"]), abs=TOL)
    assert float(p) == pytest.approx(float(gz["train_psnr"]), abs=TOL)
    assert float(r) == pytest.approx(float(gz["train_rmse"]), abs=1e-6)


@pytest.mark.parametrize("shape", [(3, 1, 256, 256), (2, 3, 64, 96), (5, 1, 37, 53), (1, 1, 11 + 1, 300), (2, 1, 128, 16)])
def test_against_oracle_port_and_fp64(shape, ssim_ref_lib):
    g = torch.Generator().manual_seed(sum(shape))
    base = torch.rand(shape, generator=g)
    pred = (base + 0.1 * torch.randn(shape, generator=g)).clamp(0, 1)
    m = _m()
    want_s, want_full = tm.structural_similarity_index_measure(pred, base, data_range=1.0, reduction="none",
                                                                return_full_image=True)
    r = m.report_metrics(pred, base, want_maps=True, want_depth=False)
    assert np.abs(r["ssim"].numpy() - want_s.numpy()).max() < TOL
    assert np.abs(r["ssim_maps"].numpy() - want_full.numpy()).max() < 5e-3
    assert float(m.ssim(pred.cuda(), base.cuda())) == pytest.approx(float(want_s.mean()), abs=TOL)
    assert float(m.psnr(pred.cuda(), base.cuda())) == pytest.approx(
        float(tm.peak_signal_noise_ratio(pred, base, data_range=1.0)), abs=TOL)
    assert float(m.rmse(pred.cuda(), base.cuda())) == pytest.approx(
        float(tm.mean_squared_error(pred, base, squared=False)), abs=1e-6)
    # independent fp64 formulation, plane by plane
    n, c, h, w = shape
    p32 = np.ascontiguousarray(pred.numpy().reshape(n * c, h, w))
    t32 = np.ascontiguousarray(base.numpy().reshape(n * c, h, w))
    s = np.zeros(n * c)
    e = np.zeros(n * c)
    assert ssim_ref_lib.ssim_ref_f64(p32.ctypes.data_as(ctypes.c_void_p), t32.ctypes.data_as(ctypes.c_void_p),
                                     n * c, h, w, s.ctypes.data_as(ctypes.c_void_p),
                                     e.ctypes.data_as(ctypes.c_void_p), None) == 0
    assert np.abs(r["ssim"].numpy() - s.reshape(n, c).mean(1)).max() < TOL
    assert np.allclose(r["mse"].numpy(), e.reshape(n, c).sum(1) / (c * h * w), rtol=1e-5)


def test_known_answers():
    m = _m()
    x = torch.rand(2, 1, 64, 64, device="cuda")
    assert float(m.ssim(x, x)) == pytest.approx(1.0, abs=1e-6)
    assert float(m.rmse(x, x)) == 0.0
    assert math.isinf(float(m.psnr(x, x)))
    a, b = 0.3, 0.7
    p = torch.full((1, 1, 32, 32), a, device="cuda")
    t = torch.full((1, 1, 32, 32), b, device="cuda")
    assert float(m.ssim(p, t)) == pytest.approx((2 * a * b + 1e-4) / (a * a + b * b + 1e-4), abs=5e-4)
    assert float(m.psnr(p, t)) == pytest.approx(10 * math.log10(1 / 0.16), abs=TOL)


def test_bf16_inputs_accumulate_in_fp32():
    pred, tgt = port.synthetic_eval_pairs(4, seed=3)
    pb, tb = pred.bfloat16(), tgt.bfloat16()
    want = tm.structural_similarity_index_measure(pb.float(), tb.float(), data_range=1.0)
    got = _m().ssim(pb.cuda(), tb.cuda())
    assert float(got) == pytest.approx(float(want), abs=TOL)


def test_errors():
    m = _m()
    with pytest.raises(RuntimeError):
        m.ssim(torch.rand(1, 1, 32, 32, device="cuda"), torch.rand(1, 1, 32, 31, device="cuda"))
    with pytest.raises(ValueError):
        m.ssim(torch.rand(1, 32, 32, device="cuda"), torch.rand(1, 32, 32, device="cuda"))
    with pytest.raises(RuntimeError):
        m.ssim(torch.rand(1, 1, 8, 8, device="cuda"), torch.rand(1, 1, 8, 8, device="cuda"))   # smaller than the window


def test_golden_loss_gradient(golden_dir):
    """ssim+psnr loss (models/wrapper.py:59-63) gradient w.r.t. the normalised prediction."""
    gz = np.load(os.path.join(golden_dir, "metrics_ref.npz"))
    xn, tn = port.synthetic_pairs(4, seed=77)
    m = _m()
    xr = xn.cuda().requires_grad_(True)
    s, p, _ = m.train_metrics(xr, tn.cuda(), denormalize=True)
    loss = -(30 * s + p)
    loss.backward()
    assert float(loss.detach()) == pytest.approx(float(gz["sp_loss"]), abs=3e-3)       # 30*1e-4 (ssim) + 1e-4 (psnr)
    g = xr.grad.cpu()
    ref = gz["sp_grad_sub"]
    assert np.abs(g[:, :, ::8, ::8].numpy() - ref).max() < 1e-3 * np.abs(ref).max()
    assert float(g.double().norm()) == pytest.approx(float(gz["sp_grad_norm"]), rel=1e-3)


@pytest.mark.parametrize("shape,denorm", [((2, 1, 64, 80), False), ((2, 2, 48, 48), True), ((1, 1, 256, 256), True)])
def test_gradient_against_autograd_through_oracle(shape, denorm):
    g = torch.Generator().manual_seed(5 + sum(shape))
    if denorm:
        t = torch.rand(shape, generator=g) * 2.4 - 1.2           # exercises both clamp sides
        x = (t + 0.3 * torch.randn(shape, generator=g))
    else:
        t = torch.rand(shape, generator=g)
        x = (t + 0.1 * torch.randn(shape, generator=g)).clamp(0, 1)
    xo = x.clone().requires_grad_(True)
    po, to = (port.denormalize(xo), port.denormalize(t)) if denorm else (xo, t)
    lo = -(30 * port.ssim(po, to) + port.psnr(po, to)) + 3.0 * port.rmse(po, to)
    lo.backward()
    m = _m()
    xg = x.cuda().requires_grad_(True)
    s, p, r = m.train_metrics(xg, t.cuda(), denormalize=denorm)
    lg = -(30 * s + p) + 3.0 * r
    lg.backward()
    assert float(lg.detach()) == pytest.approx(float(lo.detach()), abs=3e-3)
    ref = xo.grad
    assert (xg.grad.cpu() - ref).abs().max().item() < 2e-3 * ref.abs().max().item()


def test_full_size_properties():
    """Size-independent properties at the report.py sweep's shape (chunks of 256^2 pairs)."""
    m = _m()
    g = torch.Generator(device="cuda").manual_seed(1)
    base = torch.rand(512, 1, 256, 256, device="cuda", generator=g)
    noisy = (base + 0.05 * torch.randn(base.shape, device="cuda", generator=g)).clamp(0, 1)
    r_same = m.report_metrics(base, base)
    assert (r_same["ssim"] - 1).abs().max().item() < 1e-6 and r_same["mse"].abs().max().item() == 0
    r = m.report_metrics(noisy, base, chunk=64)
    r2 = m.report_metrics(noisy, base, chunk=512)
    assert torch.allclose(r["mse"], r2["mse"], rtol=1e-6) and (r["ssim"] - r2["ssim"]).abs().max().item() < 1e-6   # chunk-invariant (fp32 atomics: order-dependent last bit)
    rs = m.report_metrics(base, noisy)                      # SSIM and MSE are symmetric in (pred, target)
    assert (r["ssim"] - rs["ssim"]).abs().max().item() < 1e-5
    assert torch.allclose(r["mse"], rs["mse"], rtol=1e-6)
    assert float(r["rmse"]) == pytest.approx(math.sqrt(float(r["mse"].double().mean())), rel=1e-5)
    # depth bands average back to rows 16d+5..16d+10 of the full map
    rm = m.report_metrics(noisy[:8], base[:8], want_maps=True)
    fm = rm["ssim_maps"]
    for d in (0, 7, 15):
        band = fm[:, 0, 16 * d + 5:16 * d + 11, 5:251].reshape(8, -1).mean(-1)
        assert float(band.mean()) == pytest.approx(float(rm["depth_ssim"][d, 0]), abs=1e-5)


@pytest.mark.parametrize("dtype,denorm,h", [(torch.float32, False, 256), (torch.float32, True, 256),
                                            (torch.bfloat16, False, 256), (torch.float32, False, 128)])
def test_streaming_sweep_kernel_against_oracle(dtype, denorm, h):
    """Large batches of 256-wide images take the persistent warp-specialised streaming kernel
    (csrc/ssim.cu, namespace stream; n >= 592).  Checked against the torchmetrics restatement on images spread over
    the CTAs' image sequences (first / second / last image of a CTA, last image overall) and against the
    row-batched kernel (PAI_SSIM_NO_STREAM) on every image: per-image SSIM, 16 depth bands, squared error."""
    m = _m()
    n = 700
    g = torch.Generator(device="cuda").manual_seed(11)
    base = torch.rand(n, 1, h, 256, device="cuda", generator=g)
    # smooth structure so that SSIM is not degenerate: average neighbouring rows
    base = (base + base.roll(1, 2) + base.roll(1, 3)) / 3
    pred = (base + 0.05 * torch.randn(n, 1, h, 256, device="cuda", generator=g)).clamp_(0, 1)
    if denorm:
        base, pred = 2 * base - 1, 2 * pred - 1
    base, pred = base.to(dtype), pred.to(dtype)
    bands = h == 256                      # depth bands need h / 16 > 10 rows
    s1, e1, b1, _ = m._launch_fwd(pred, base, denorm, bands, False)
    os.environ["PAI_SSIM_NO_STREAM"] = "1"
    try:
        s0, e0, b0, _ = m._launch_fwd(pred, base, denorm, bands, False)
    finally:
        del os.environ["PAI_SSIM_NO_STREAM"]
    torch.cuda.synchronize()
    win = (h - 10) * 246
    assert float((s1 - s0).abs().max()) / win < 1e-6                  # same arithmetic, different summation order
    if bands:
        assert float((b1 - b0).abs().max()) / ((h // 16 - 10) * 246) < 1e-6
    assert torch.allclose(e1, e0, rtol=1e-5, atol=1e-6)
    idx = [0, 1, 147, 148, 149, 295, 296, 591, 592, n - 1]
    pf, bf = pred[idx].float().cpu(), base[idx].float().cpu()
    if denorm:
        pf, bf = port.denormalize(pf), port.denormalize(bf)
    want = tm.structural_similarity_index_measure(pf, bf, data_range=1.0, reduction="none")
    assert float((s1[idx].cpu() / win - want).abs().max()) < TOL
    want_sse = ((pf - bf) ** 2).flatten(1).sum(1)
    assert torch.allclose(e1[idx].cpu(), want_sse, rtol=1e-4)
    if bands:
        want_depth = torch.stack([port.depth_ssim(pf[i:i + 1], bf[i:i + 1])[:, 0] for i in range(len(idx))])
        got_depth = b1[idx].cpu() / (6 * 246)
        assert float((got_depth - want_depth).abs().max()) < TOL
