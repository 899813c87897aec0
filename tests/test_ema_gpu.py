"""callbacks/ema.py on the GPU: the multi-tensor EMA kernel (pai_ema_multi) against torch_ema's update rule evaluated in
fp64 on the host, over the whole Pix2Pix GAN parameter set (reference: callbacks/ema.py:24-34, decay 0.9999 main.py:131)
and around real training steps (store / copy_to / restore swap for validation)."""
import pytest
import torch

import pix2pix_port as port

pytestmark = pytest.mark.gpu


def test_ema_kernel_matches_update_rule_over_the_gan_parameters():
    from callbacks.ema import EMACallback
    from models.pix2pix import Pix2Pix
    from models.utils import init_weights
    from models.wrapper import Discriminator
    from pai_b200 import lib
    torch.manual_seed(0)
    m = Pix2Pix(in_channels=1, out_channels=1, dropout=0.0, loss_type="gan")
    m.discriminator = Discriminator(in_channels=1)
    m.discriminator.apply(init_weights)
    m = m.cuda().train()
    cb = EMACallback(0.9999)
    cb.on_fit_start(None, m)
    params = [p for p in m.parameters() if p.requires_grad]
    ref = [p.detach().double().cpu() for p in params]
    x, t = (a.cuda() for a in port.synthetic_pairs(2, seed=3))
    before = lib.launches
    for step in range(1, 4):
        m.training_step((x, t), step)
        cb.on_train_batch_end(None, m)
        d = min(0.9999, (1 + step) / (10 + step))
        ref = [s - (1 - d) * (s - p.detach().double().cpu()) for s, p in zip(ref, params)]
    assert lib.launches > before
    assert cb.ema.num_updates == 3 and len(cb.ema.shadow_params) == len(params)
    for s, r, p in zip(cb.ema.shadow_params, ref, params):
        assert torch.allclose(s.double().cpu(), r, rtol=1e-6, atol=1e-7)
    # validation runs on the averaged weights, training resumes on the live ones
    live = [p.detach().clone() for p in params]
    cb.on_validation_start(None, m)
    assert all(torch.equal(p.detach(), s) for p, s in zip(params, cb.ema.shadow_params))
    with torch.no_grad():
        m.eval()
        y = m(x)
        m.train()
    assert torch.isfinite(y).all()
    cb.on_validation_end(None, m)
    assert all(torch.equal(p.detach(), q) for p, q in zip(params, live))
