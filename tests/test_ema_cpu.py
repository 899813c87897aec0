"""callbacks/ema.py mirror: the moving-average arithmetic (torch_ema's documented update rule with its warm-up of the
decay), the store / copy_to / restore swap around validation and the checkpoint round trip -- device independent, run on
CPU."""
import torch

from callbacks.ema import EMACallback, ExponentialMovingAverage


def test_update_rule_with_decay_warm_up():
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    frozen = torch.nn.Parameter(torch.ones(2), requires_grad=False)
    ema = ExponentialMovingAverage(list(net.parameters()) + [frozen], decay=0.9999)
    assert len(ema.shadow_params) == 4                      # only parameters that require grad
    ref = [p.detach().clone().double() for p in net.parameters()]
    for step in range(1, 6):
        with torch.no_grad():
            for p in net.parameters():
                p.add_(torch.randn_like(p))
        ema.update()
        d = min(0.9999, (1 + step) / (10 + step))
        ref = [s - (1 - d) * (s - p.detach().double()) for s, p in zip(ref, net.parameters())]
        for s, r in zip(ema.shadow_params, ref):
            assert torch.allclose(s.double(), r, atol=1e-6)
    assert ema.num_updates == 5


def test_callback_swaps_weights_for_validation_and_round_trips():
    net = torch.nn.Linear(3, 3)
    cb = EMACallback(0.5)
    cb.on_fit_start(None, net)
    with torch.no_grad():
        net.weight.add_(1.0)
    live = net.weight.detach().clone()
    cb.on_train_batch_end(None, net)
    shadow = cb.ema.shadow_params[0].clone()
    assert not torch.equal(shadow, live)
    cb.on_validation_start(None, net)
    assert torch.equal(net.weight.detach(), shadow)          # validation runs on the averaged weights
    cb.on_validation_end(None, net)
    assert torch.equal(net.weight.detach(), live)            # training resumes on the live weights
    state = cb.on_save_checkpoint(None, net, {})
    cb2 = EMACallback(0.9)
    cb2.on_fit_start(None, net)
    cb2.on_load_checkpoint(None, net, state)
    assert cb2.ema.decay == 0.5 and cb2.ema.num_updates == 1 and torch.equal(cb2.ema.shadow_params[0], shadow)
