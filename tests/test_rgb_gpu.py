"""GPU parity of the drop-in models built with the REFERENCE'S OWN DEFAULT CONSTRUCTORS -- ``Pix2Pix()`` = 3 -> 3
channels and ``Discriminator()`` with a 6-plane first convolution (models/pix2pix.py:25-27, models/wrapper.py:34,225) --
against fixtures produced by the unmodified reference (tests/golden/pix2pix_rgb_ref.npz, oracle/gen_golden_rgb.py).
The image channels travel in a zero-padded 64-channel carrier through the ordinary implicit-GEMM kernels
(pai_b200/engine.py: ``_carrier``); everything between the first and the last convolution is the grayscale path.
Tolerances as in tests/test_pix2pix_gpu.py (BASELINE.json north_star: 1e-2 eval-mode at bf16)."""
import os

import numpy as np
import pytest
import torch

import pix2pix_port as port

pytestmark = pytest.mark.gpu


def rgb_pairs(n, seed=900):
    xs, ts = zip(*[port.synthetic_pairs(n, seed=seed + c) for c in range(3)])
    return torch.cat(xs, 1), torch.cat(ts, 1)


def _build(loss_type="gan", seed=0):
    from models.pix2pix import Pix2Pix
    torch.manual_seed(seed)
    return Pix2Pix(dropout=0.0, loss_type=loss_type).cuda()      # reference defaults: 3 -> 3 channels, Discriminator()


@pytest.fixture(scope="module")
def gz(golden_dir):
    return np.load(os.path.join(golden_dir, "pix2pix_rgb_ref.npz"))


def test_default_constructors_give_the_references_state_dict(gz):
    m = _build()
    sd = m.state_dict()
    keys = sorted(sd.keys())
    assert keys == list(gz["state_keys"])
    assert tuple(sd["unet.encoders.0.weight"].shape) == (64, 3, 4, 4)
    assert tuple(sd["discriminator.discriminator.0.block.0.weight"].shape)[1] == 6
    cs = np.array([[float(sd[k].cpu().double().sum()), float(sd[k].cpu().double().abs().sum())] for k in keys])
    assert np.allclose(cs, gz["state_checksums"], rtol=1e-5, atol=1e-6)


def test_eval_forward_and_discriminator(gz):
    m = _build()
    x, target = rgb_pairs(2)
    m.eval()
    with torch.no_grad():
        y = m(x.cuda())
        logits = m.discriminator(x.cuda(), target.cuda())
    assert y.shape == (2, 3, 256, 256) and y.dtype == torch.float32
    assert np.abs(y.cpu()[:, :, ::4, ::4].numpy() - gz["gen_eval_sub"]).max() < 1e-2
    assert np.abs(logits.cpu().numpy() - gz["disc_logits"]).max() < 1e-2


def test_train_forward_losses_and_gradients(gz):
    m = _build()
    x, target = rgb_pairs(2)
    x, target = x.cuda(), target.cuda()
    m.train()
    y = m(x)
    d = np.abs(y.detach().cpu()[:, :, ::4, ::4].numpy() - gz["gen_train_sub"])
    assert d.max() < 3e-2 and d.mean() < 5e-3, (d.max(), d.mean())
    loss = m.loss(x, y, target)
    assert float(loss.detach()) == pytest.approx(float(gz["gan_gloss0"]), rel=2e-2)
    loss.backward()
    named = dict(m.named_parameters())
    bad = []
    for k, want in zip(gz["grad_keys"], gz["grad_norms"]):
        g = named[str(k)].grad
        assert g is not None, k
        got = float(g.double().norm())
        if want < 1e-4:       # conv biases in front of a BatchNorm: mathematically zero gradient
            assert got < 1e-3, (k, got)
        elif abs(got - want) > 0.08 * want:
            bad.append((str(k), got, float(want)))
    assert not bad, bad
    # the discriminator's own loss (fake = the detached prediction): every gradient incl. the 6-plane first convolution
    m.zero_grad(set_to_none=True)
    dl = m.discriminator_loss(m.discriminator(x, y.detach()), m.discriminator(x, target))
    assert float(dl.detach()) == pytest.approx(float(gz["d_loss0"]), rel=2e-2)
    dl.backward()
    dnamed = dict(m.discriminator.named_parameters())
    for k, want in zip(gz["d_grad_keys"], gz["d_grad_norms"]):
        got = float(dnamed[str(k)].grad.double().norm())
        assert got == pytest.approx(float(want), rel=0.08), (k, got, want)


def test_three_gan_training_steps_against_reference_logs(gz):
    m = _build()
    m.train()
    x, target = rgb_pairs(2)
    batch = (x.cuda(), target.cuda())
    for _ in range(3):
        m.training_step(batch, 0)
    tol = {"d_loss": 0.05, "loss": 0.05, "train_ssim": 0.03, "train_psnr": 0.03, "train_rmse": 0.03}
    for k, vals in m.logged.items():
        got = np.array([float(v) for v in vals])
        want = gz["gan_log_" + k]
        assert np.allclose(got, want, rtol=tol[k], atol=2e-3), (k, got, want)


def test_step_graph_replays_the_rgb_step():
    """``enable_step_graph()`` on the 3-channel model: the carrier packs are rebuilt inside the capture after every Adam
    update (they are not among the packs the fused optimizer rewrites in place); replayed steps must log what eager steps
    log (first steps closely, later ones within the chaos of batch-2 GAN training)."""
    x, target = rgb_pairs(2)
    batch = (x.cuda(), target.cuda())
    logs = {}
    for mode in ("eager", "graph"):
        m = _build(seed=1).train()
        if mode == "graph":
            m.enable_step_graph(warmup=2)
        for i in range(5):
            m.training_step(batch, i)
        torch.cuda.synchronize()
        if mode == "graph":
            assert m.__dict__["_pai_step_graph"].replays == 3
        logs[mode] = {k: np.array([float(v) for v in vals]) for k, vals in m.logged.items()}
    for k in logs["eager"]:
        a, g = logs["eager"][k], logs["graph"][k]
        assert np.allclose(a[:3], g[:3], rtol=1e-2, atol=2e-3), (k, a, g)
        assert np.allclose(a, g, rtol=0.15, atol=5e-3), (k, a, g)
