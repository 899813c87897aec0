"""GPU parity of the drop-in ``models`` package (Pix2Pix U-Net + PatchGAN on the B200 kernels)
against the oracle: fixtures produced by the UNMODIFIED reference (tests/golden/pix2pix_ref.npz, made by
oracle/gen_golden.py) and the travelling CPU restatement oracle/pix2pix_port.py.

Tolerances (BASELINE.json north_star): generator outputs 1e-2 max-abs at bf16 in eval mode; train-mode
BatchNorm over tiny populations makes 1e-2 marginal even for the reference against itself
(2.07e-2), so the train-mode bound is 3e-2 max-abs / 5e-3 mean-abs."""
import os

import numpy as np
import pytest
import torch

import pix2pix_port as port

pytestmark = pytest.mark.gpu


def _build(loss_type, seed=0):
    from models.pix2pix import Pix2Pix
    from models.wrapper import Discriminator
    from models.utils import init_weights
    torch.manual_seed(seed)
    m = Pix2Pix(in_channels=1, out_channels=1, dropout=0.0, loss_type=loss_type)
    if loss_type == "gan":
        m.discriminator = Discriminator(in_channels=1)
        m.discriminator.apply(init_weights)
    return m.cuda()


@pytest.fixture(scope="module")
def gz(golden_dir):
    return np.load(os.path.join(golden_dir, "pix2pix_ref.npz"))


def test_state_dict_is_the_references(gz):
    m = _build("gan")
    sd = m.state_dict()
    keys = sorted(sd.keys())
    assert keys == list(gz["state_keys"])
    cs = np.array([[float(sd[k].cpu().double().sum()), float(sd[k].cpu().double().abs().sum())] for k in keys])
    # bit-exact in the container that produced the fixture (tests/test_oracle_models.py); other hosts may
    # differ in the last bit of the vectorised CPU normal_()
    assert np.allclose(cs, gz["state_checksums"], rtol=1e-5, atol=1e-6)


def test_eval_forward_and_discriminator(gz):
    m = _build("gan")
    x, target = port.synthetic_pairs(2, seed=1234)
    m.eval()
    with torch.no_grad():
        y = m(x.cuda())
        logits = m.discriminator(x.cuda(), target.cuda())
    assert y.shape == (2, 1, 256, 256) and y.dtype == torch.float32
    assert np.abs(y.cpu()[:, :, ::4, ::4].numpy() - gz["gen_eval_sub"]).max() < 1e-2
    assert float(y.mean()) == pytest.approx(float(gz["gen_eval_stats"][0]), abs=2e-3)
    assert np.abs(logits.cpu().numpy() - gz["disc_logits"]).max() < 1e-2


def test_train_forward_and_gradients(gz):
    m = _build("gan")
    x, target = port.synthetic_pairs(2, seed=1234)
    m.train()
    y = m(x.cuda())
    d = np.abs(y.detach().cpu()[:, :, ::4, ::4].numpy() - gz["gen_train_sub"])
    assert d.max() < 3e-2 and d.mean() < 5e-3, (d.max(), d.mean())
    loss = m.loss(x.cuda(), y, target.cuda())
    assert float(loss.detach()) == pytest.approx(float(gz["gan_gloss0"]), rel=2e-2)
    loss.backward()
    named = dict(m.named_parameters())
    bad = []
    for k, want in zip(gz["grad_keys"], gz["grad_norms"]):
        k = str(k)
        g = named[k].grad
        assert g is not None, k
        got = float(g.double().norm())
        if want < 1e-4:       # conv biases in front of a BatchNorm: mathematically zero gradient
            assert got < 1e-3, (k, got)
            continue
        if abs(got - want) > 0.08 * want:
            bad.append((k, got, float(want)))
    assert not bad, bad


@pytest.mark.parametrize("n", [1, 3])
def test_forward_backward_against_oracle_port(n):
    """Same weights, same inputs: every parameter gradient of the ssim+psnr loss vs the CPU oracle."""
    m = _build("ssim+psnr", seed=3)
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    x, target = port.synthetic_pairs(n, seed=50 + n)
    tr = port.OracleTrainer(sd, "ssim+psnr")
    yo = port.unet_forward(tr.sd, x, training=True)
    lo = port.generator_loss(tr.sd, "ssim+psnr", x, yo, target)
    lo.backward()
    m.train()
    y = m(x.cuda())
    loss = m.loss(x.cuda(), y, target.cuda())
    loss.backward()
    d = (y.detach().cpu() - yo.detach()).abs()
    assert d.max().item() < 4e-2 and d.mean().item() < 6e-3, (d.max().item(), d.mean().item())
    assert float(loss.detach()) == pytest.approx(float(lo.detach()), rel=3e-2)
    named = dict(m.named_parameters())
    for k in tr.g_keys:
        go = tr.sd[k].grad
        g = named[k].grad.cpu()
        if float(go.norm()) < 1e-4:       # bias in front of a BatchNorm: rounding noise only
            assert float(g.norm()) < 1e-3, k
            continue
        cos = float((g.double() * go.double()).sum() / (g.double().norm() * go.double().norm() + 1e-30))
        # The deep layers (BatchNorm over N*4 .. N*64 values) sit at the bf16 noise floor: the reference against ITSELF
        # (fp32 vs bf16 autocast, oracle/bf16_selfcheck.py 1) gives 0.925 .. 0.955 there at batch 1, 0.985 on level 2 and
        # 0.998+ on the outer layers.  This repo over 10 runs (profiles/scripts/grad_noise.py; split-K and BatchNorm
        # atomics reorder sums from run to run): batch 1: deep 0.888 .. 0.957, level 2 0.980 .. 0.983, outer >= 0.989;
        # batch 3: deep >= 0.954, level 2 0.996, outer >= 0.993.  The bounds sit 0.02 - 0.03 under those minima.
        deep = any(f"encoders.{i}." in k for i in (3, 4, 5, 6, 7)) or any(f"decoders.{j}." in k for j in (0, 1, 2, 3))
        mid = "encoders.2." in k
        if n == 1:
            bound = 0.86 if deep else 0.96 if mid else 0.98
        else:
            bound = 0.93 if deep else 0.98
        assert cos > bound, (k, cos)
        assert float(g.norm()) == pytest.approx(float(go.norm()), rel=0.12), k
    # running statistics advanced exactly like the reference's BatchNorm
    for k, v in m.state_dict().items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert torch.allclose(v.cpu(), tr.sd[k], rtol=3e-2, atol=3e-3), k
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(tr.sd[k])


@pytest.mark.parametrize("loss_type,prefix", [("gan", "gan_log_"), ("ssim+psnr", "sp_log_")])
def test_three_training_steps_against_reference_logs(gz, loss_type, prefix):
    m = _build(loss_type)
    m.train()
    x, target = port.synthetic_pairs(2, seed=1234)
    batch = (x.cuda(), target.cuda())
    for _ in range(3):
        m.training_step(batch, 0)
    tol = {"d_loss": 0.05, "loss": 0.05, "train_ssim": 0.03, "train_psnr": 0.03, "train_rmse": 0.03}
    for k, vals in m.logged.items():
        got = np.array([float(v) for v in vals])
        want = gz[prefix + k]
        assert np.allclose(got, want, rtol=tol[k], atol=2e-3), (k, got, want)


def test_frozen_generator_builds_no_graph_and_discriminator_dgrad_only():
    m = _build("gan")
    m.train()
    x, target = port.synthetic_pairs(1, seed=7)
    x, target = x.cuda(), target.cuda()
    opt_g, opt_d = m.optimizers()
    m.toggle_optimizer(opt_d)
    pred = m.unet(x)
    assert not pred.requires_grad
    m.untoggle_optimizer(opt_d)
    m.toggle_optimizer(opt_g)
    pred = m.unet(x)
    assert pred.requires_grad
    m.loss(x, pred, target).backward()
    assert all(p.grad is None for p in m.discriminator.parameters())
    assert all(p.grad is not None for p in m.unet.parameters())
    m.untoggle_optimizer(opt_g)


def test_unsupported_configurations_fail_loudly():
    """No CPU / PyTorch fallback: what the kernels cannot run raises (3-channel models run: tests/test_rgb_gpu.py)."""
    from models.pix2pix import Pix2Pix
    m3 = Pix2Pix(in_channels=3, out_channels=3, dropout=0.0, loss_type="mse").cuda()
    with pytest.raises(RuntimeError):
        m3(torch.zeros(1, 1, 256, 256, device="cuda"))          # wrong number of image channels
    with pytest.raises(RuntimeError):
        m3(torch.zeros(1, 3, 250, 256, device="cuda"))          # not divisible by 2^8
    m1 = Pix2Pix(in_channels=1, out_channels=1, dropout=0.0, loss_type="mse")
    with pytest.raises(RuntimeError):
        m1(torch.zeros(1, 1, 256, 256))                         # CPU tensors: there is no host path


def test_train_mode_dropout2d_matches_reference_arithmetic(monkeypatch):
    """The reference's default constructor has ``dropout=0.5``: Dropout2d after the BatchNorm of decoders 0-2
    (models/pix2pix.py:107,176-183).  With the SAME (sample, channel) masks the fused engine must reproduce the
    reference arithmetic (restated here on the oracle port's layer functions): forward 3e-2 max-abs at bf16 like the
    other train-mode checks, weight-gradient norms within 12 %.  RNG streams themselves cannot match (CPU vs device)."""
    import torch.nn.functional as F
    from models.pix2pix import Pix2Pix
    from pai_b200 import ops
    torch.manual_seed(2)
    m = Pix2Pix(in_channels=1, out_channels=1, dropout=0.5, loss_type="mse").cuda().train()
    assert [b.dropout for b in list(m.unet.decoders)[:-1]] == [0.5, 0.5, 0.5, 0, 0, 0, 0]
    sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    n = 3
    x, target = port.synthetic_pairs(n, seed=31)
    g = torch.Generator().manual_seed(9)
    masks = [torch.bernoulli(torch.full((n, 512), 0.5), generator=g) * 2.0 for _ in range(3)]
    calls = []

    def fixed_mask(nn_, c, p, device):
        assert (nn_, c, p) == (n, 512, 0.5)
        calls.append(len(calls))
        return masks[len(calls) - 1].to(device).contiguous()

    monkeypatch.setattr(ops, "dropout2d_mask", fixed_mask)
    y = m(x.cuda())
    assert len(calls) == 3
    loss = F.mse_loss(y, target.cuda())
    loss.backward()

    # reference arithmetic with the same masks (oracle/pix2pix_port.py layer by layer + Dropout2d)
    ref = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    h, feats = x, []
    for i in range(8):
        if i == 0:
            h = F.conv2d(h, ref["unet.encoders.0.weight"], ref["unet.encoders.0.bias"], 2, 1)
        else:
            pfx = f"unet.encoders.{i}.encode"
            h = F.conv2d(F.leaky_relu(h, 0.2), ref[pfx + ".1.weight"], ref[pfx + ".1.bias"], 2, 1)
            if pfx + ".2.weight" in ref:
                h = port._bn(ref, pfx + ".2", h, True)
        feats.append(h)
    feats.pop()
    for j in range(8):
        if j:
            h = torch.cat([h, feats.pop()], 1)
        if j < 7:
            pfx = f"unet.decoders.{j}.decode"
            h = F.conv_transpose2d(F.relu(h), ref[pfx + ".1.weight"], ref[pfx + ".1.bias"], 2, 1)
            h = port._bn(ref, pfx + ".2", h, True)
            if j < 3:
                h = h * masks[j][:, :, None, None]
        else:
            h = F.conv_transpose2d(h, ref["unet.decoders.7.weight"], ref["unet.decoders.7.bias"], 2, 1)
    yo = torch.tanh(h)
    F.mse_loss(yo, target).backward()
    d = (y.detach().cpu() - yo.detach()).abs()
    assert d.max().item() < 4e-2 and d.mean().item() < 6e-3, (d.max().item(), d.mean().item())
    named = dict(m.named_parameters())
    for k in ("unet.decoders.0.decode.1.weight", "unet.decoders.2.decode.1.weight", "unet.decoders.5.decode.1.weight",
              "unet.encoders.2.encode.1.weight", "unet.encoders.7.encode.1.weight"):
        got, want = float(named[k].grad.norm()), float(ref[k].grad.norm())
        assert got == pytest.approx(want, rel=0.12), (k, got, want)
    # eval mode: Dropout2d is the identity
    m.eval()
    with torch.no_grad():
        y1, y2 = m(x.cuda()), m(x.cuda())
    # (split-K layers accumulate with atomics: equal up to summation order, which can flip bf16 roundings downstream)
    assert torch.allclose(y1, y2, atol=1e-2) and len(calls) == 3


@pytest.mark.parametrize("loss_type", ["gan", "ssim+psnr"])
def test_step_graph_matches_eager_training(loss_type):
    """``enable_step_graph()``: the whole training step replayed as one CUDA graph gives the same logged losses /
    metrics as eager launches (same kernels, same order), and FusedAdam's device-side step counter round-trips through
    ``state_dict()``.

    Training from a random init with batch 2 is chaotic: the gradient atomics sum in a different order in every run
    and the differences grow step by step (the reference's own GAN d_loss jumps 1.16 -> 2.65 -> 1.22 within steps
    5-8, tests/golden/curves_ref.npz).  The yardstick is therefore the run-to-run spread of two EAGER runs: the graph
    run may deviate from eager run A by at most 5x what eager run B does (floor 2 %; a stale buffer or a wrong step
    count shows up as tens of percent)."""
    data = [tuple(t.cuda() for t in port.synthetic_pairs(2, seed=900 + i)) for i in range(3)]
    nsteps = 5
    runs = {}
    for mode in ("eager_a", "eager_b", "graph"):
        m = _build(loss_type, seed=1).train()
        if mode == "graph":
            m.enable_step_graph(warmup=2)
        for i in range(nsteps):
            m.training_step(data[i % 3], i)
        torch.cuda.synchronize()
        if mode == "graph":
            assert m.__dict__["_pai_step_graph"].replays == nsteps - 2
        opts = m.optimizers()
        opt_g = opts[0] if isinstance(opts, list) else opts
        steps = {float(st["step"]) for st in opt_g.state_dict()["state"].values()}
        assert steps == {float(nsteps)}, (mode, steps)
        nbt = {k: int(v) for k, v in m.state_dict().items() if k.endswith("num_batches_tracked")}
        runs[mode] = ({k: np.array([float(v) for v in vals]) for k, vals in m.logged.items()}, nbt)
    (la, na), (lb, nb), (lg, ng) = runs["eager_a"], runs["eager_b"], runs["graph"]
    assert na == nb == ng
    assert la.keys() == lg.keys()
    for k in la:
        assert len(la[k]) == len(lg[k]) == nsteps
        scale = np.abs(la[k]) + 1e-3
        # the eager warm-up steps (0, 1) and the first replay must agree closely; later ones within the eager spread
        assert np.allclose(la[k][:3], lg[k][:3], rtol=5e-3, atol=1e-3), (k, la[k], lg[k])
        dev_graph = float((np.abs(lg[k] - la[k]) / scale).max())
        dev_eager = float((np.abs(lb[k] - la[k]) / scale).max())
        assert dev_graph <= max(5 * dev_eager, 2e-2), (k, dev_graph, dev_eager, la[k], lb[k], lg[k])


def test_step_graph_recaptures_and_optimizer_state_round_trips():
    """The graph bakes in pointers and hyper-parameters: changing the learning rate or loading a state dict must drop it
    and capture again (never replay stale arguments), and FusedAdam's device-side step counter must survive
    ``state_dict()`` -> ``load_state_dict()`` into a fresh optimizer exactly like torch.optim.Adam's per-parameter steps."""
    data = tuple(t.cuda() for t in port.synthetic_pairs(2, seed=77))
    m = _build("ssim+psnr", seed=4).train()
    m.enable_step_graph(warmup=1)
    runner = m.__dict__["_pai_step_graph"]
    for i in range(4):
        m.training_step(data, i)
    assert runner.replays == 3
    opt = m.optimizers()
    import copy
    sd = copy.deepcopy(opt.state_dict())      # like torch's, state_dict() hands out the LIVE per-parameter state
    model_sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    assert {float(st["step"]) for st in sd["state"].values()} == {4.0}
    # new learning rate -> signature mismatch -> one eager warm-up call, then a fresh capture
    w_before = m.unet.encoders[3].encode[1].weight.detach().clone()
    for group in opt.param_groups:
        group["lr"] = 0.0
    for i in range(3):
        m.training_step(data, i)
    torch.cuda.synchronize()
    assert runner.replays == 3 + 2
    assert torch.equal(m.unet.encoders[3].encode[1].weight.detach(), w_before), "lr = 0 must freeze the weights"
    assert {float(st["step"]) for st in opt.state_dict()["state"].values()} == {7.0}
    # a fresh model + optimizer resumes from the saved state: same step count, then the same update as the original
    m2 = _build("ssim+psnr", seed=4).train()
    m2.load_state_dict(model_sd)
    opt2 = m2.optimizers()
    opt2.load_state_dict(sd)
    m2.training_step(data, 0)
    torch.cuda.synchronize()
    steps2 = {float(st["step"]) for st in opt2.state_dict()["state"].values()}
    assert steps2 == {5.0}, steps2
    moved = (m2.unet.encoders[3].encode[1].weight.detach() - model_sd["unet.encoders.3.encode.1.weight"]).abs().max().item()
    assert 0 < moved < 1e-3        # one Adam step of lr 2e-4 with the restored moments


def test_batched_discriminator_passes_equal_the_two_reference_calls(monkeypatch):
    """training_step evaluates D(x, target) and D(x, pred) as ONE pass over 2N samples (models/wrapper.py of this repo);
    the PatchGAN has no BatchNorm, so the logged d_loss and every discriminator weight after the update must equal those
    of the reference's two separate calls (models/wrapper.py:121-122 of the reference) up to bf16 summation order."""
    import models.wrapper as W
    x, target = port.synthetic_pairs(4, seed=77)
    batch = (x.cuda(), target.cuda())
    out = {}
    for batched in (True, False):
        monkeypatch.setattr(W, "BATCH_DISCRIMINATOR_PASSES", batched)
        m = _build("gan", seed=5).train()
        m.training_step(batch, 0)
        torch.cuda.synchronize()
        out[batched] = (float(m.logged["d_loss"][-1]), float(m.logged["loss"][-1]),
                        {k: v.detach().float().clone() for k, v in m.discriminator.named_parameters()})
    (da, la, pa), (db, lb, pb) = out[True], out[False]
    assert da == pytest.approx(db, rel=2e-3) and la == pytest.approx(lb, rel=1e-2)
    for k in pa:
        # The first Adam step moves every weight by lr * sign-like(g) = +-2e-4: an element whose gradient is rounding
        # noise may flip (difference 2 lr), everything else must agree to a small fraction of the step
        d = (pa[k] - pb[k]).abs()
        assert d.max().item() <= 4.1e-4 and d.mean().item() < 2e-5, (k, d.max().item(), d.mean().item())
