"""report.py's evaluation loop on the GPU (pai_b200.report.evaluate) against the oracle's CPU restatement of
report.py:72-104,144-146,188-217 (oracle/pix2pix_port.report_metrics / depth_ssim): identity "model" and a frozen
drop-in Pix2Pix.  Tolerance 1e-4 on SSIM / PSNR statistics (north star), maps as in test_metrics_gpu."""
import numpy as np
import pytest
import torch

import pix2pix_port as port

pytestmark = pytest.mark.gpu


def test_evaluate_identity_matches_report_py_restatement():
    from pai_b200 import report
    x, t = port.synthetic_pairs(6, seed=99)                    # in [-1, 1] like the dataloader's batches
    batches = [(x[:4], t[:4]), (x[4:], t[4:])]
    res = report.evaluate(lambda v: v, batches, want_maps=True, device=torch.device("cuda"))
    pd, td = port.denormalize(x), port.denormalize(t)
    want = port.report_metrics(pd, td, chunk=64, want_maps=True)
    assert np.abs(res["ssim"].cpu().numpy() - want["ssim"].numpy()).max() < 1e-4
    assert np.abs(res["psnr"].cpu().numpy() - want["psnr"].numpy()).max() < 1e-4
    assert float(res["ssim_stat"]) == pytest.approx(float(want["ssim"].mean()), abs=1e-4)
    assert float(res["psnr_stat"]) == pytest.approx(float(want["psnr"].mean()), abs=1e-4)
    assert float(res["rmse_stat"]) == pytest.approx(float(want["rmse"]), abs=1e-6)
    depth = port.depth_ssim(pd, td)
    assert np.abs(res["depth_ssim"].cpu().numpy() - depth.numpy()).max() < 1e-4
    assert res["ssim_maps_uint8"].dtype == torch.uint8 and res["ssim_maps_uint8"].shape == (6, 1, 256, 256)
    csv = report.depth_csv(res["depth_ssim"].cpu())
    assert csv.splitlines()[0] == "depth,mean,std" and len(csv.splitlines()) == 17
    assert res["parameter_count"] == 0


def test_evaluate_frozen_dropin_model():
    from models.pix2pix import Pix2Pix
    from pai_b200 import report
    torch.manual_seed(0)
    m = Pix2Pix(in_channels=1, out_channels=1, dropout=0.0, loss_type="ssim").cuda()
    m.freeze()
    x, t = port.synthetic_pairs(3, seed=5)
    res = report.evaluate(m, [(x, t)], want_maps=False)
    assert res["preds"].shape == (3, 1, 256, 256) and float(res["preds"].min()) >= 0 and float(res["preds"].max()) <= 1
    assert res["ssim"].shape == (3,) and res["ssim_maps"] is None and res["parameter_count"] == 54_413_313
    # the same predictions through the oracle's metric restatement
    want = port.report_metrics(res["preds"].cpu(), res["targets"].cpu(), chunk=64, want_maps=False)
    assert np.abs(res["ssim"].cpu().numpy() - want["ssim"].numpy()).max() < 1e-4
    assert np.abs(res["psnr"].cpu().numpy() - want["psnr"].numpy()).max() < 1e-4
