"""Seeded synthetic inputs and weights for the BASELINE configurations: parameter inventories in Gluon's
collect_params() order, He-normal / uniform weight generators, synthetic frames and caption sources.

Shared by the product-side benchmark (`bench.py`, `tools/`), which needs weights and inputs but must not touch the CPU oracle,
and by the oracle and the tests, which need the SAME tensors to compare against.  No arithmetic of the hot path lives here."""
import math

import torch

DENSE_CFG = (6, 12, 24, 16)
GROWTH, BOTTLENECK = 32, 128
RESNET_CH = (64, 128, 256, 512)


# parameter inventories (canonical order == Gluon's collect_params() order; see include/tennis_b200.h)
def _bn_names(prefix):
    return [prefix + ".gamma", prefix + ".beta", prefix + ".running_mean", prefix + ".running_var"]


def densenet121_param_shapes():
    shapes = [("conv0.weight", (64, 3, 7, 7))] + [(n, (64,)) for n in _bn_names("bn0")]
    c = 64
    for b, nl in enumerate(DENSE_CFG):
        for l in range(nl):
            p = "block%d.layer%d" % (b + 1, l + 1)
            shapes += [(n, (c,)) for n in _bn_names(p + ".bn1")]
            shapes += [(p + ".conv1.weight", (BOTTLENECK, c, 1, 1))]
            shapes += [(n, (BOTTLENECK,)) for n in _bn_names(p + ".bn2")]
            shapes += [(p + ".conv2.weight", (GROWTH, BOTTLENECK, 3, 3))]
            c += GROWTH
        if b < 3:
            p = "trans%d" % (b + 1)
            shapes += [(n, (c,)) for n in _bn_names(p + ".bn")]
            shapes += [(p + ".conv.weight", (c // 2, c, 1, 1))]
            c //= 2
    shapes += [(n, (c,)) for n in _bn_names("bn5")]
    return shapes


def resnet18_v2_param_shapes():
    shapes = [(n, (3,)) for n in _bn_names("bn_data")]
    shapes += [("conv0.weight", (64, 3, 7, 7))] + [(n, (64,)) for n in _bn_names("bn0")]
    cin = 64
    for s, c in enumerate(RESNET_CH):
        for b in range(2):
            p = "stage%d.block%d" % (s + 1, b + 1)
            shapes += [(n, (cin,)) for n in _bn_names(p + ".bn1")]
            shapes += [(p + ".conv1.weight", (c, cin, 3, 3))]
            shapes += [(n, (c,)) for n in _bn_names(p + ".bn2")]
            shapes += [(p + ".conv2.weight", (c, c, 3, 3))]
            if b == 0 and cin != c:
                shapes += [(p + ".downsample.weight", (c, cin, 1, 1))]
            cin = c
    shapes += [(n, (512,)) for n in _bn_names("bn_final")]
    return shapes


PARAM_SHAPES = {"densenet121": densenet121_param_shapes, "resnet18_v2": resnet18_v2_param_shapes}


def synthetic_params(arch, seed=1234, dtype=torch.float32):
    """Seeded synthetic weights: He-normal convs, BN gamma~U(.5,1.5), beta~N(0,.1),
    running_mean~N(0,.1), running_var~U(.5,1.5) so activations stay O(1) through the depth."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in PARAM_SHAPES[arch]():
        if name.endswith(".weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif name.endswith(".gamma") or name.endswith(".running_var"):
            t = torch.rand(shape, generator=g) + 0.5
        else:
            t = torch.randn(shape, generator=g) * 0.1
        if name.startswith("bn_data") and (name.endswith(".gamma") or name.endswith(".beta")):
            # BatchNorm(scale=False, center=False): gamma fixed to 1, beta fixed to 0 (A.2)
            t = torch.ones(shape) if name.endswith(".gamma") else torch.zeros(shape)
        out[name] = t.to(dtype)
    return out


def flatten_params(arch, params):
    return torch.cat([params[n].reshape(-1).float() for n, _ in PARAM_SHAPES[arch]()]).contiguous()



def rnn_param_shapes(cell, D, H, bidirectional=True):
    G = 3 if cell == "gru" else 4
    shapes = []
    for d in (["l0", "r0"] if bidirectional else ["l0"]):
        shapes += [(d + "_i2h_weight", (G * H, D)), (d + "_h2h_weight", (G * H, H)), (d + "_i2h_bias", (G * H,)),
                   (d + "_h2h_bias", (G * H,))]
    return shapes


def synthetic_rnn_params(cell, D, H, seed=4321, bidirectional=True, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in rnn_param_shapes(cell, D, H, bidirectional):
        fan = shape[1] if len(shape) == 2 else H
        out[name] = ((torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan)).to(dtype)
    return out


def synthetic_frames(n, size=224, seed=100):
    """Config-1 style input: u8 pixels ~ U{0..255} -> ToTensor -> Normalize (train.py:142-147)."""
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (n, size, size, 3), generator=g, dtype=torch.uint8)
    return u8, normalize_u8(u8)


def structured_clips(B, T, size=224, seed=300):
    """Clips whose content differs from clip to clip (a smooth random colour field per clip, drifting over its frames, plus
    pixel noise): unlike i.i.d. noise frames, their features -- and the detector's logits -- spread over the classes, so an
    argmax comparison between two implementations is a real test (VERDICT r1 weak #1c).  Returns (u8 (B,T,H,W,3), fp32
    normalised (B,T,3,H,W))."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand(B, 1, 3, 7, 7, generator=g) * 1.6 - 0.3
    drift = (torch.rand(B, T, 3, 7, 7, generator=g) - 0.5) * 0.5
    field = torch.nn.functional.interpolate((low + drift).reshape(B * T, 3, 7, 7), size=(size, size), mode="bilinear",
                                            align_corners=False)
    noise = (torch.rand(B * T, 3, size, size, generator=g) - 0.5) * 0.2
    u8 = ((field + noise).clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    x = normalize_u8(u8)
    return u8.reshape(B, T, size, size, 3), x.reshape(B, T, 3, size, size)


def normalize_u8(u8):
    mean = torch.tensor([0.485, 0.456, 0.406])
    std = torch.tensor([0.229, 0.224, 0.225])
    x = u8.float().div(255.0)
    return ((x - mean) / std).permute(0, 3, 1, 2).contiguous()


# ----------------------------------------------------------------------------------------------------------------------
# captioner (GNMT) parameters and sources
def cell_param_shapes(cell, in_dim, H):
    G = 3 if cell == "gru" else 4
    return [("i2h_weight", (G * H, in_dim)), ("h2h_weight", (G * H, H)), ("i2h_bias", (G * H,)), ("h2h_bias", (G * H,))]


def gnmt_param_shapes(cell="lstm", H=128, D_src=1024, E=100, V=254, num_layers=2, num_bi_layers=1):
    """Structural names follow Gluon."""
    shapes = []
    in_dim = D_src
    for i in range(num_layers):
        if i < num_bi_layers:
            for side in ("l_cell", "r_cell"):
                shapes += [("encoder.rnn_cells.%d.%s.%s" % (i, side, n), s) for n, s in cell_param_shapes(cell, in_dim, H)]
            in_dim = 2 * H
        else:
            shapes += [("encoder.rnn_cells.%d.%s" % (i, n), s) for n, s in cell_param_shapes(cell, in_dim, H)]
            in_dim = H
    for i in range(num_layers):
        din = E + H if i == 0 else 2 * H
        shapes += [("decoder.rnn_cells.%d.%s" % (i, n), s) for n, s in cell_param_shapes(cell, din, H)]
    shapes += [("decoder.attention_cell.proj_query.weight", (H, H))]
    shapes += [("tgt_embed.weight", (V, E)), ("tgt_proj.weight", (V, H)), ("tgt_proj.bias", (V,))]
    return shapes


def synthetic_gnmt_params(seed=10000, scale=0.1, **kw):
    """model.initialize(init=Uniform(0.1)) style weights (train_gnmt.py:231), seeded."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in gnmt_param_shapes(**kw):
        out[name] = (torch.rand(shape, generator=g) * 2 - 1) * scale
    return out


def synthetic_sources(B, T, D, seed=100, min_len=None):
    """Config-4 style inputs: non-negative features N(0,1)*0.5 clipped at 0, valid_length ~ U{min_len..T}."""
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(B, T, D, generator=g) * 0.5).clamp(min=0)
    lo = min_len if min_len is not None else max(1, T // 4)
    vl = torch.randint(lo, T + 1, (B,), generator=g).float()
    vl[0] = T
    for b in range(B):
        x[b, int(vl[b]):] = 0
    return x, vl
