"""CPU emulation: which bf16 rounding sites dominate the train-mode generator output error?"""
import sys, itertools
sys.path.insert(0, 'oracle')
import torch, torch.nn.functional as F
import pix2pix_port as port
torch.set_num_threads(8)

def r(t, on): return t.bfloat16().float() if on else t

def fwd(sd, x, sites, hi_levels=()):
    """sites: set of {'w','in','raw','act'}; hi_levels: encoder/decoder level ids kept in fp32 ('e5','d0',...)"""
    L = 8
    h = x
    feats = []
    def keep(tag): return tag in hi_levels
    for i in range(L):
        tag = f'e{i}'
        on = not keep(tag)
        if i == 0:
            w = sd["unet.encoders.0.weight"]; b = sd["unet.encoders.0.bias"]
            h = F.conv2d(r(h, 'in' in sites and on), r(w, 'w' in sites and on), b, 2, 1)
            h = r(h, 'raw' in sites and on)
        else:
            p = f"unet.encoders.{i}.encode"
            a = r(F.leaky_relu(h, 0.2), 'act' in sites and on)
            h = F.conv2d(a, r(sd[p + ".1.weight"], 'w' in sites and on), sd[p + ".1.bias"], 2, 1)
            if 'raw_after_stats' in sites and on and p + ".2.weight" in sd:
                # stats from fp32, normalise the bf16-rounded raw
                mean = h.mean((0, 2, 3), keepdim=True); var = h.var((0, 2, 3), unbiased=False, keepdim=True)
                hq = h.bfloat16().float()
                h = (hq - mean) / torch.sqrt(var + 1e-5) * sd[p + ".2.weight"].view(1, -1, 1, 1) + sd[p + ".2.bias"].view(1, -1, 1, 1)
            else:
                h = r(h, 'raw' in sites and on)
                if p + ".2.weight" in sd:
                    h = F.batch_norm(h, None, None, sd[p + ".2.weight"], sd[p + ".2.bias"], True, 0.1, 1e-5)
        feats.append(h)
    feats.pop()
    for i in range(L):
        tag = f'd{i}'
        on = not keep(tag)
        if i:
            h = torch.cat([h, feats.pop()], 1)
        if i < L - 1:
            p = f"unet.decoders.{i}.decode"
            a = r(F.relu(h), 'act' in sites and on)
            h = F.conv_transpose2d(a, r(sd[p + ".1.weight"], 'w' in sites and on), sd[p + ".1.bias"], 2, 1)
            h = r(h, 'raw' in sites and on)
            h = F.batch_norm(h, None, None, sd[p + ".2.weight"], sd[p + ".2.bias"], True, 0.1, 1e-5)
        else:
            p = f"unet.decoders.{i}"
            a = r(h, 'act' in sites and on)
            h = F.conv_transpose2d(a, r(sd[p + ".weight"], 'w' in sites and on), sd[p + ".bias"], 2, 1)
    return torch.tanh(h)

if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    sd = port.init_state(5, loss_type="ssim")
    x, t = port.synthetic_pairs(n, seed=640)
    with torch.no_grad():
        ref = fwd(sd, x, set())
        def rep(name, sites, hi=()):
            y = fwd(sd, x, sites, hi)
            d = (y - ref).abs()
            print(f"{name:50s} max {d.max():.3e} mean {d.mean():.3e}", flush=True)
        rep('all (w,act,raw)', {'w', 'act', 'raw', 'in'})
        rep('w only', {'w'})
        rep('act only', {'act', 'in'})
        rep('raw only', {'raw'})
        rep('w+act (raw fp32)', {'w', 'act', 'in'})
        deep = ('e4', 'e5', 'e6', 'e7', 'd0', 'd1', 'd2', 'd3')
        rep('all, deep levels (<=16x16 out) fp32', {'w', 'act', 'raw', 'in'}, deep)
        deeper = ('e5', 'e6', 'e7', 'd0', 'd1', 'd2')
        rep('all, levels <=8x8 fp32', {'w', 'act', 'raw', 'in'}, deeper)
        outer = ('e0', 'd7')
        rep('all, e0+d7 fp32', {'w', 'act', 'raw', 'in'}, outer)
        rep('all, e0,e1,d6,d7 fp32', {'w', 'act', 'raw', 'in'}, ('e0', 'e1', 'd6', 'd7'))
        rep('all but e0..e3,d4..d7 fp32 (only deep bf16)', {'w', 'act', 'raw', 'in'}, ('e0','e1','e2','e3','d4','d5','d6','d7'))

def fwd2(sd, x, mode):
    """all sites bf16; raw handling mode: 'plain' | 'stats32' (stats from fp32, normalise rounded) | 'center' (round raw - mean)"""
    L = 8
    q = lambda t: t.bfloat16().float()
    def bn(h, g, b):
        mean = h.mean((0, 2, 3), keepdim=True); var = h.var((0, 2, 3), unbiased=False, keepdim=True)
        if mode == 'plain':
            hq = q(h); mean = hq.mean((0, 2, 3), keepdim=True); var = hq.var((0, 2, 3), unbiased=False, keepdim=True)
            return (hq - mean) / torch.sqrt(var + 1e-5) * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
        if mode == 'stats32':
            return (q(h) - mean) / torch.sqrt(var + 1e-5) * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
        if mode == 'center':
            return q(h - mean) / torch.sqrt(var + 1e-5) * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
        if mode == 'ratio':
            print('   |mean|/std per channel: median %.2f max %.2f' % (float((mean.abs() / var.sqrt()).median()), float((mean.abs() / var.sqrt()).max())))
            return (h - mean) / torch.sqrt(var + 1e-5) * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
    h = x; feats = []
    for i in range(L):
        if i == 0:
            h = q(F.conv2d(q(h), q(sd["unet.encoders.0.weight"]), sd["unet.encoders.0.bias"], 2, 1))
        else:
            p = f"unet.encoders.{i}.encode"
            h = F.conv2d(q(F.leaky_relu(h, 0.2)), q(sd[p + ".1.weight"]), sd[p + ".1.bias"], 2, 1)
            h = bn(h, sd[p + ".2.weight"], sd[p + ".2.bias"]) if p + ".2.weight" in sd else q(h)
        feats.append(h)
    feats.pop()
    for i in range(L):
        if i: h = torch.cat([h, feats.pop()], 1)
        if i < L - 1:
            p = f"unet.decoders.{i}.decode"
            h = F.conv_transpose2d(q(F.relu(h)), q(sd[p + ".1.weight"]), sd[p + ".1.bias"], 2, 1)
            h = bn(h, sd[p + ".2.weight"], sd[p + ".2.bias"])
        else:
            p = f"unet.decoders.{i}"
            h = F.conv_transpose2d(q(h), q(sd[p + ".weight"]), sd[p + ".bias"], 2, 1)
    return torch.tanh(h)

if __name__ == '__main__':
    with torch.no_grad():
        for mode in ('plain', 'stats32', 'center', 'ratio'):
            y = fwd2(sd, x, mode); d = (y - ref).abs()
            print(f"fwd2 {mode:10s} max {d.max():.3e} mean {d.mean():.3e}", flush=True)
