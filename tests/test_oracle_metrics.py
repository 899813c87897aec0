"""Pins the CPU oracle of the torchmetrics-0.11.4 path (oracle/torchmetrics_port.py) with
known answers, an independent fp64 C formulation (oracle/ssim_ref.c) and the fixtures the
real reference produced (oracle/gen_golden.py)."""
import ctypes
import math
import os

import numpy as np
import pytest
import torch

import pix2pix_port as port
import torchmetrics_port as tm


def _c_ssim(lib, p, t, want_full=False):
    n, _, h, w = p.shape
    p32 = np.ascontiguousarray(p.numpy().reshape(n, h, w), dtype=np.float32)
    t32 = np.ascontiguousarray(t.numpy().reshape(n, h, w), dtype=np.float32)
    s = np.zeros(n)
    e = np.zeros(n)
    full = np.zeros((n, h, w)) if want_full else None
    rc = lib.ssim_ref_f64(p32.ctypes.data_as(ctypes.c_void_p), t32.ctypes.data_as(ctypes.c_void_p),
                          n, h, w, s.ctypes.data_as(ctypes.c_void_p), e.ctypes.data_as(ctypes.c_void_p),
                          full.ctypes.data_as(ctypes.c_void_p) if want_full else None)
    assert rc == 0
    return s, e, full


def test_identical_images():
    x = torch.rand(2, 1, 64, 64)
    assert float(tm.structural_similarity_index_measure(x, x, data_range=1.0)) == pytest.approx(1.0, abs=1e-6)
    assert float(tm.mean_squared_error(x, x, squared=False)) == 0.0
    assert math.isinf(float(tm.peak_signal_noise_ratio(x, x, data_range=1.0)))


def test_constant_images_closed_form(ssim_ref_lib):
    a, b = 0.3, 0.7
    p = torch.full((1, 1, 32, 32), a)
    t = torch.full((1, 1, 32, 32), b)
    want = (2 * a * b + 1e-4) / (a * a + b * b + 1e-4)
    # fp32 upstream dataflow: the 121 window taps sum to 1+eps, which the contrast term amplifies
    # by (a-b)^2/c2 = 178x on constant images, so the fp32 port only holds ~5e-4 here; the fp64
    # formulation holds the closed form tightly.
    assert float(tm.structural_similarity_index_measure(p, t, data_range=1.0)) == pytest.approx(want, abs=5e-4)
    s_c, e_c, _ = _c_ssim(ssim_ref_lib, p, t)
    assert s_c[0] == pytest.approx(want, abs=2e-7)   # inputs are float32(0.3), float32(0.7)
    assert e_c[0] == pytest.approx(0.16 * 32 * 32, rel=1e-6)
    assert float(tm.peak_signal_noise_ratio(p, t, data_range=1.0)) == pytest.approx(10 * math.log10(1 / 0.16), abs=1e-4)
    assert float(tm.mean_squared_error(p, t)) == pytest.approx(0.16, abs=1e-6)


def test_gaussian_window():
    g = tm._gaussian_1d(11, 1.5, torch.float64, "cpu")[0]
    assert g.shape == (11,) and float(g.sum()) == pytest.approx(1.0, abs=1e-12)
    assert torch.allclose(g, g.flip(0))
    assert float(g[5] / g[4]) == pytest.approx(math.exp(0.5 / 2.25), rel=1e-12)


def test_against_independent_fp64(ssim_ref_lib):
    pred, tgt = port.synthetic_eval_pairs(4, seed=11)
    s_c, e_c, full_c = _c_ssim(ssim_ref_lib, pred, tgt, want_full=True)
    s, full = tm.structural_similarity_index_measure(pred, tgt, data_range=1.0, reduction="none", return_full_image=True)
    assert np.abs(s.numpy() - s_c).max() < 5e-6
    # per-pixel: fp32 E[x^2]-mu^2 cancellation in flat regions costs up to a few 1e-4 on single
    # pixels of the fp32 upstream dataflow; the means above are what the metric tolerance is about
    d = np.abs(full.numpy()[:, 0] - full_c)
    assert d.max() < 2e-3 and d.mean() < 2e-5
    mse = np.array([float(tm.mean_squared_error(a, b)) for a, b in zip(pred, tgt)])
    assert np.allclose(mse, e_c / (256 * 256), rtol=1e-5)


def test_depth_bands_equal_full_map_rows():
    """report.depth_ssim band d == mean of full-map rows 16d+5..16d+10, cols 5..250."""
    pred, tgt = port.synthetic_eval_pairs(3, seed=5)
    d = port.depth_ssim(pred, tgt)
    _, full = tm.structural_similarity_index_measure(pred, tgt, data_range=1.0, reduction="none", return_full_image=True)
    for k in range(16):
        band = full[:, 0, 16 * k + 5:16 * k + 11, 5:251].reshape(3, -1).mean(-1)
        assert float(band.mean()) == pytest.approx(float(d[k, 0]), abs=2e-6)
        assert float(band.std()) == pytest.approx(float(d[k, 1]), abs=2e-6)


def test_psnr_is_batch_global():
    pred, tgt = port.synthetic_eval_pairs(4, seed=9)
    mse = float(tm.mean_squared_error(pred, tgt))
    assert float(tm.peak_signal_noise_ratio(pred, tgt, data_range=1.0)) == pytest.approx(10 * math.log10(1 / mse), abs=1e-4)


def test_errors():
    with pytest.raises(RuntimeError):
        tm.structural_similarity_index_measure(torch.rand(1, 1, 32, 32), torch.rand(1, 1, 32, 31), data_range=1.0)
    with pytest.raises(ValueError):
        tm.structural_similarity_index_measure(torch.rand(1, 32, 32), torch.rand(1, 32, 32), data_range=1.0)


def test_golden_metrics(golden_dir):
    gz = np.load(os.path.join(golden_dir, "metrics_ref.npz"))
    pred, tgt = port.synthetic_eval_pairs(8, seed=4321)
    r = port.report_metrics(pred, tgt)
    assert np.abs(r["ssim"].numpy() - gz["ssim_per_image"]).max() < 1e-6
    assert np.abs(r["psnr"].numpy() - gz["psnr_per_image"]).max() < 1e-4
    assert np.allclose(r["mse"].numpy(), gz["mse_per_image"], rtol=1e-6)
    assert np.abs(r["depth_ssim"].numpy() - gz["depth_ssim"]).max() < 1e-6
    assert float(r["rmse"]) == pytest.approx(float(gz["rmse_global"]), rel=1e-6)
    assert np.abs(r["ssim_maps"][:, :, ::8, ::8].numpy() - gz["ssim_map_sub"]).max() < 1e-6
    assert float(port.ssim(pred, tgt)) == pytest.approx(float(gz["ssim"]), abs=1e-6)


def test_golden_loss_gradient(golden_dir, ssim_ref_lib):
    """Closed-form fp64 gradient (Appendix A) against autograd through the reference's loss."""
    gz = np.load(os.path.join(golden_dir, "metrics_ref.npz"))
    xn, tn = port.synthetic_pairs(4, seed=77)
    p = port.denormalize(xn)
    t = port.denormalize(tn)
    n, _, h, w = p.shape
    p32 = np.ascontiguousarray(p.numpy().reshape(n, h, w))
    t32 = np.ascontiguousarray(t.numpy().reshape(n, h, w))
    g = np.zeros((n, h, w))
    rc = ssim_ref_lib.ssim_psnr_grad_ref_f64(
        p32.ctypes.data_as(ctypes.c_void_p), t32.ctypes.data_as(ctypes.c_void_p),
        n, h, w, ctypes.c_double(-30.0), ctypes.c_double(-1.0), g.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    # chain through denormalize: 0.5 * [0 <= 0.5x+0.5 <= 1]  (clamp passes gradient at the bounds)
    v = (xn * 0.5 + 0.5).numpy().reshape(n, h, w)
    gx = 0.5 * g * ((v >= 0) & (v <= 1))
    ref = gz["sp_grad_sub"][:, 0]
    got = gx[:, ::8, ::8]
    assert np.abs(got - ref).max() < 1e-3 * np.abs(ref).max()
    assert np.linalg.norm(gx) == pytest.approx(float(gz["sp_grad_norm"]), rel=1e-3)
