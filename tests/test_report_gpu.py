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


def test_eval_batchnorm_folding_matches_unfolded_and_oracle(monkeypatch):
    """Eval-mode inference (report.py:26-43,63): BatchNorm folded into the GEMM operands, both consumers of an encoder
    written by one epilogue.  Must equal the unfolded kernels (same bf16 operands up to the folded scale) and the CPU
    oracle within the north star's 1e-2 -- also AFTER training steps moved the running statistics (the folded packs are
    stamped with them)."""
    from models.pix2pix import Pix2Pix
    from pai_b200 import engine
    torch.manual_seed(3)
    m = Pix2Pix(in_channels=1, out_channels=1, dropout=0.0, loss_type="ssim+psnr").cuda()
    x, t = port.synthetic_pairs(8, seed=21)
    xc, tc = x.cuda(), t.cuda()
    m.train()
    for i in range(2):                      # running statistics and weights away from their initial values
        m.training_step((xc, tc), i)
    m.eval()
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        y_fold = m(xc)
        monkeypatch.setattr(engine, "FOLD_EVAL_BN", False)
        y_plain = m(xc)
        monkeypatch.setattr(engine, "FOLD_EVAL_BN", True)
        yo = port.unet_forward(sd, x, training=False)
    assert float((y_fold - y_plain).abs().max()) < 1e-2
    assert float((y_fold.cpu() - yo).abs().max()) < 1e-2, float((y_fold.cpu() - yo).abs().max())
    # one more training step changes weights AND running statistics: the folded operands must follow
    m.train()
    m.training_step((xc, tc), 2)
    m.eval()
    sd2 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        y2 = m(xc)
        yo2 = port.unet_forward(sd2, x, training=False)
    assert float((y2.cpu() - yo2).abs().max()) < 1e-2
    assert float((y2 - y_fold).abs().max()) > 0


def test_uint8_conversion_and_hot_colormap_on_the_device():
    """``to_int`` (models/utils.py:12) of a RAW SSIM map (negative values included, report.py:134-135) must equal
    torchvision's host conversion bit for bit; the "afmhot" images (report.py:220-233; matplotlib is not installed here,
    so the colormap is pinned to its published definition: 256-entry table of clip(2v, 2v-0.5, 2v-1) at v = i/255,
    index min(int(256 x), 255))."""
    from torchvision.transforms import ConvertImageDtype
    from pai_b200 import report
    g = torch.Generator().manual_seed(1)
    raw = torch.rand(2, 1, 64, 64, generator=g) * 1.6 - 0.4          # SSIM-map-like values in [-0.4, 1.2)
    raw[0, 0, 0, :4] = torch.tensor([0.0, 1.0, -1e-4, 0.999999])
    want = ConvertImageDtype(torch.uint8)(raw)                        # what report.py computes on the host
    got = report._to_int(raw.cuda()).cpu()
    assert torch.equal(got, want)
    with pytest.raises(RuntimeError):                                 # no host branch: the product path is the kernel
        report._to_int(raw)
    img = torch.rand(3, 1, 32, 48, generator=g)
    img[0, 0, 0, :3] = torch.tensor([0.0, 1.0, 0.5])
    idx = (img * 256).to(torch.int64).clamp(max=255)
    v = idx.float() / 255.0
    rgb = torch.cat([(2 * v).clamp(0, 1), (2 * v - 0.5).clamp(0, 1), (2 * v - 1).clamp(0, 1)], 1)
    want_hot = ConvertImageDtype(torch.uint8)(rgb)
    got_hot = report.hot_images(img.cuda()).cpu()
    assert got_hot.shape == (3, 3, 32, 48) and torch.equal(got_hot, want_hot)
