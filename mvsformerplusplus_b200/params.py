"""Parameter containers with the reference's state-dict key names and shapes.

The modules built here hold weights only (their ``forward`` is never called): the arithmetic runs in
the CUDA library.  Key names/shapes follow the reference so ``load_state_dict(strict=True)`` of the
hot-path subset of a reference checkpoint works unchanged (reference: test.py:212-220; key inventory
pinned in tests/golden/hotpath_state_dict_keys.txt, produced from models/FMT.py:140-152,
models/cost_volume.py:21-49, models/module.py:367-408,453-504,602-629).
"""
import torch
import torch.nn as nn

from .config import stage_list


class Bag(nn.Module):
    """Named container; never executed."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the hot path runs in the CUDA library")


def _conv_bn2d(cin, cout):
    m = Bag()
    m.conv = nn.Conv2d(cin, cout, 3, padding=1, bias=False)
    m.bn = nn.BatchNorm2d(cout)
    return m


def _conv_bn3d(cin, cout, stride):
    m = Bag()
    m.conv = nn.Conv3d(cin, cout, 3, stride=stride, padding=1, bias=False)
    m.bn = nn.BatchNorm3d(cout)
    return m


def _deconv_bn3d_named(cin, cout, stride, opad):
    m = Bag()
    m.conv = nn.ConvTranspose3d(cin, cout, 3, stride=stride, padding=1, output_padding=opad, bias=False)
    m.bn = nn.BatchNorm3d(cout)
    return m


def _deconv_bn3d_seq(cin, cout, stride, opad):
    return nn.Sequential(nn.ConvTranspose3d(cin, cout, 3, stride=stride, padding=1, output_padding=opad, bias=False),
                         nn.BatchNorm3d(cout), nn.ReLU(inplace=True))


class _LN3D(Bag):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))


def _cross_block(dim, hidden):
    b = Bag()
    b.norm1 = nn.LayerNorm(dim)
    b.attn = Bag()
    b.attn.q_proj = nn.Linear(dim, dim, bias=False)
    b.attn.k_proj = nn.Linear(dim, dim, bias=False)
    b.attn.v_proj = nn.Linear(dim, dim, bias=False)
    b.attn.proj = nn.Linear(dim, dim, bias=True)
    b.ls1 = Bag()
    b.ls1.gamma = nn.Parameter(torch.ones(dim))
    b.norm2 = nn.LayerNorm(dim)
    b.mlp = Bag()
    b.mlp.fc1 = nn.Linear(dim, hidden)
    b.mlp.fc2 = nn.Linear(hidden, dim)
    b.ls2 = Bag()
    b.ls2.gamma = nn.Parameter(torch.ones(dim))
    return b


def build_fmt(fmt_cfg):
    d = fmt_cfg["d_model"]
    bc = fmt_cfg.get("base_channel", 8)
    m = Bag()
    m.FMT = Bag()
    m.FMT.layers = nn.ModuleList([_cross_block(d, 4 * d) for _ in fmt_cfg["layer_names"]])
    iv = fmt_cfg.get("init_values", 1.0)
    for blk in m.FMT.layers:
        nn.init.constant_(blk.ls1.gamma, iv)
        nn.init.constant_(blk.ls2.gamma, iv)
    m.dim_reduction_1 = nn.Conv2d(bc * 8, bc * 4, 1, bias=False)
    m.dim_reduction_2 = nn.Conv2d(bc * 4, bc * 2, 1, bias=False)
    m.dim_reduction_3 = nn.Conv2d(bc * 2, bc, 1, bias=False)
    m.smooth_1 = nn.Conv2d(bc * 4, bc * 4, 3, padding=1, bias=False)
    m.smooth_2 = nn.Conv2d(bc * 2, bc * 2, 3, padding=1, bias=False)
    m.smooth_3 = nn.Conv2d(bc, bc, 3, padding=1, bias=False)
    return m


def _costreg_unet(c, kind):
    """kind 'CostRegNet' (stride 2 in D,H,W, named deconvs, 3^3 prob without bias) or
    'CostRegNet3D' (stride (1,2,2), Sequential deconvs, 1^3 prob with bias)."""
    m = Bag()
    m.kind = kind
    s = 2 if kind == "CostRegNet" else (1, 2, 2)
    m.conv1 = _conv_bn3d(c, 2 * c, s)
    m.conv2 = _conv_bn3d(2 * c, 2 * c, 1)
    m.conv3 = _conv_bn3d(2 * c, 4 * c, s)
    m.conv4 = _conv_bn3d(4 * c, 4 * c, 1)
    m.conv5 = _conv_bn3d(4 * c, 8 * c, s)
    m.conv6 = _conv_bn3d(8 * c, 8 * c, 1)
    if kind == "CostRegNet":
        m.conv7 = _deconv_bn3d_named(8 * c, 4 * c, 2, 1)
        m.conv9 = _deconv_bn3d_named(4 * c, 2 * c, 2, 1)
        m.conv11 = _deconv_bn3d_named(2 * c, c, 2, 1)
        m.prob = nn.Conv3d(c, 1, 3, padding=1, bias=False)
    else:
        m.conv7 = _deconv_bn3d_seq(8 * c, 4 * c, (1, 2, 2), (0, 1, 1))
        m.conv9 = _deconv_bn3d_seq(4 * c, 2 * c, (1, 2, 2), (0, 1, 1))
        m.conv11 = _deconv_bn3d_seq(2 * c, c, (1, 2, 2), (0, 1, 1))
        m.prob = nn.Conv3d(c, 1, 1)
    return m


def _costreg_transformer(c, tc):
    mid = tc["mid_channel"]
    dr = tuple(tc["down_rate"])
    m = Bag()
    m.kind = "PureTransformerCostReg"
    m.pe_proj = nn.Conv3d(c * 3, c, 1, 1, bias=False)
    m.down = nn.Sequential(nn.Conv3d(c, mid, kernel_size=dr, stride=dr), _LN3D(mid))
    layers = []
    for _ in range(tc["layer_num"]):
        b = Bag()
        b.gamma1 = nn.Parameter(torch.tensor(1.0))
        b.gamma2 = nn.Parameter(torch.tensor(1.0))
        b.attn = Bag()
        b.attn.qkv = nn.Linear(mid, 3 * mid, bias=False)
        b.attn.proj = nn.Linear(mid, mid, bias=True)
        b.norm1 = nn.LayerNorm(mid)
        b.ffn = Bag()
        b.ffn.linear1 = nn.Linear(mid, int(mid * tc["mlp_ratio"]))
        b.ffn.linear2 = nn.Linear(int(mid * tc["mlp_ratio"]), mid)
        b.norm2 = nn.LayerNorm(mid)
        layers.append(b)
    m.attention_layers = nn.ModuleList(layers)
    m.up = nn.Sequential(nn.ConvTranspose3d(mid, c, kernel_size=dr, stride=dr), _LN3D(c))
    m.prob = nn.Conv3d(c, 1, 1)
    return m


def build_stage(args, ndepth, stage_idx):
    c = stage_list(args["base_ch"], stage_idx)
    m = Bag()
    m.vis = nn.Sequential(_conv_bn2d(1, 16), _conv_bn2d(16, 16), _conv_bn2d(16, 8), nn.Conv2d(8, 1, 1), nn.Sigmoid())
    t = args.get("cost_reg_type", ["Normal"] * 4)[stage_idx]
    if t == "PureTransformerCostReg":
        m.cost_reg = _costreg_transformer(c, args["transformer_config"][stage_idx])
    elif ndepth <= args.get("model_th", 8):
        m.cost_reg = _costreg_unet(c, "CostRegNet3D")
    else:
        m.cost_reg = _costreg_unet(c, "CostRegNet")
    return m


def build_hotpath_params(args):
    root = Bag()
    root.FMT_module = build_fmt(args["FMT_config"])
    root.fusions = nn.ModuleList([build_stage(args, args["ndepths"][i], i) for i in range(len(args["ndepths"]))])
    return root
