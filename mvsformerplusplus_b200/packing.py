"""Packs a reference state_dict (hot-path keys) into the flat fp32 weight layouts libmvsf_b200 documents in
include/mvsf_b200.h / csrc/*.cu.  BatchNorm (eval mode, eps 1e-5) is folded into the preceding conv in fp64.
This runs once at install time; nothing here is on the per-frame path."""
import torch

BN_EPS = 1e-5


def _d(t):
    return t.detach().double().cpu()


def _fold_bn(sd, p):
    scale = _d(sd[p + "weight"]) / torch.sqrt(_d(sd[p + "running_var"]) + BN_EPS)
    shift = _d(sd[p + "bias"]) - _d(sd[p + "running_mean"]) * scale
    return scale, shift


def _cat(parts, pad_to=4):
    flat = torch.cat([x.reshape(-1).double() for x in parts])
    if flat.numel() % pad_to:
        flat = torch.cat([flat, torch.zeros(pad_to - flat.numel() % pad_to, dtype=torch.float64)])
    return flat.float().contiguous()


def pack_vis(sd, p):
    """p = 'fusions.{s}.vis.'  -> w1[9][16] b1[16] w2[16 ic][9][16 oc] b2[16] w3[16][9][8] b3[8] w4[8] b4[1]"""
    parts = []
    for i, (cin, cout) in enumerate([(1, 16), (16, 16), (16, 8)]):
        w = _d(sd[f"{p}{i}.conv.weight"])  # [cout, cin, 3, 3]
        scale, shift = _fold_bn(sd, f"{p}{i}.bn.")
        w = w * scale.view(-1, 1, 1, 1)
        parts.append(w.permute(1, 2, 3, 0).reshape(cin, 9, cout))  # [ic][tap][oc]
        parts.append(shift)
    parts.append(_d(sd[p + "3.weight"]).reshape(8))
    parts.append(_d(sd[p + "3.bias"]).reshape(1))
    out = _cat(parts)
    assert out.numel() == 3652, out.numel()  # 3649 + pad
    return out


def pack_costreg_unet(sd, p):
    """p = 'fusions.{s}.cost_reg.'  -> (kind, flat) ; kind 0 = CostRegNet, 1 = CostRegNet3D"""
    is3d = (p + "conv7.0.weight") in sd
    parts = []
    for name in ("conv1", "conv2", "conv3", "conv4", "conv5", "conv6"):
        w = _d(sd[f"{p}{name}.conv.weight"])  # [cout, cin, 3,3,3]
        scale, shift = _fold_bn(sd, f"{p}{name}.bn.")
        w = w * scale.view(-1, 1, 1, 1, 1)
        parts.append(w.permute(2, 3, 4, 1, 0).reshape(27, w.shape[1], w.shape[0]))  # [tap][ci][co]
        parts.append(shift)
    for name in ("conv7", "conv9", "conv11"):
        if is3d:
            w = _d(sd[f"{p}{name}.0.weight"])  # [cin, cout, 3,3,3]
            scale, shift = _fold_bn(sd, f"{p}{name}.1.")
        else:
            w = _d(sd[f"{p}{name}.conv.weight"])
            scale, shift = _fold_bn(sd, f"{p}{name}.bn.")
        w = w * scale.view(1, -1, 1, 1, 1)
        parts.append(w.permute(2, 3, 4, 0, 1).reshape(27, w.shape[0], w.shape[1]))  # [tap][ci][co]
        parts.append(shift)
    pw = _d(sd[p + "prob.weight"])
    if is3d:
        parts += [pw.reshape(8), _d(sd[p + "prob.bias"]).reshape(1)]
    else:
        parts.append(pw[0].permute(1, 2, 3, 0).reshape(27, 8))  # [tap][ci]
    return (1 if is3d else 0), _cat(parts)


def pack_costreg_tr(sd, p, layers):
    """p = 'fusions.{s}.cost_reg.' ; layout documented in csrc/costreg_tr.cu"""
    parts = [_d(sd[p + "pe_proj.weight"]).reshape(8, 24)]
    wd = _d(sd[p + "down.0.weight"])  # [64, 8, 2,4,4] -> [64][kd][kh][kw][ci]
    parts += [wd.permute(0, 2, 3, 4, 1).reshape(64, 256), _d(sd[p + "down.0.bias"]),
              _d(sd[p + "down.1.weight"]), _d(sd[p + "down.1.bias"])]
    for i in range(layers):
        q = f"{p}attention_layers.{i}."
        parts += [_d(sd[q + "attn.qkv.weight"]), _d(sd[q + "attn.proj.weight"]), _d(sd[q + "attn.proj.bias"]),
                  _d(sd[q + "gamma1"]).reshape(1).expand(64), _d(sd[q + "norm1.weight"]), _d(sd[q + "norm1.bias"]),
                  _d(sd[q + "ffn.linear1.weight"]), _d(sd[q + "ffn.linear1.bias"]),
                  _d(sd[q + "ffn.linear2.weight"]), _d(sd[q + "ffn.linear2.bias"]),
                  _d(sd[q + "gamma2"]).reshape(1).expand(64), _d(sd[q + "norm2.weight"]), _d(sd[q + "norm2.bias"])]
    wu = _d(sd[p + "up.0.weight"])  # [64 ci, 8 co, 2,4,4] -> [kd][kh][kw][co][ci] = [256][64]
    parts += [wu.permute(2, 3, 4, 1, 0).reshape(256, 64), _d(sd[p + "up.0.bias"]).repeat(32),
              _d(sd[p + "up.1.weight"]), _d(sd[p + "up.1.bias"]),
              _d(sd[p + "prob.weight"]).reshape(8), _d(sd[p + "prob.bias"]).reshape(1)]
    return _cat(parts, pad_to=8)  # multiple of 8 floats: the fp16 hi/lo copies stay 16-byte aligned


def pack_fmt(sd, p="FMT_module."):
    """layout documented in csrc/fmt.cu"""
    parts = []
    for i in range(4):
        q = f"{p}FMT.layers.{i}."
        parts += [_d(sd[q + "norm1.weight"]), _d(sd[q + "norm1.bias"]),
                  torch.cat([_d(sd[q + "attn.q_proj.weight"]), _d(sd[q + "attn.k_proj.weight"]),
                             _d(sd[q + "attn.v_proj.weight"])], 0),
                  _d(sd[q + "attn.proj.weight"]), _d(sd[q + "attn.proj.bias"]), _d(sd[q + "ls1.gamma"]),
                  _d(sd[q + "norm2.weight"]), _d(sd[q + "norm2.bias"]),
                  _d(sd[q + "mlp.fc1.weight"]), _d(sd[q + "mlp.fc1.bias"]),
                  _d(sd[q + "mlp.fc2.weight"]), _d(sd[q + "mlp.fc2.bias"]), _d(sd[q + "ls2.gamma"])]
    for k in (1, 2, 3):
        w = _d(sd[f"{p}dim_reduction_{k}.weight"])
        parts.append(w.reshape(w.shape[0], w.shape[1]))
    for k in (1, 2, 3):
        w = _d(sd[f"{p}smooth_{k}.weight"])  # [co, ci, 3, 3] -> [tap][ci][co]
        parts.append(w.permute(2, 3, 1, 0).reshape(9, w.shape[1], w.shape[0]))
    return _cat(parts, pad_to=8)
