"""Host side of the B200 hot path: Python/PyTorch mirrors of the reference seams that
enqueue libmvsf_b200 kernels.  PyTorch is used for device memory, streams and module/state-dict plumbing only.

  StageNet.forward(features, proj_matrices, depth_values, tmp, position3d=None)   <- models/cost_volume.py:51-133
  FMT_with_pathway.forward(features)                                              <- models/FMT.py:164-206
  HotPathNet.forward_features(features, proj_matrices, depth_values, tmp)         <- DINOv2_mvsformer_model.py:117-179
  install(model)  rebinds the two module seams of a reference-constructed DINOv2MVSNet        (test.py drop-in)

Tensors crossing the seams keep the reference's logical shapes ([B,V,C,H,W] features, [B,D,H,W] volumes).  Feature
maps produced by FMT_with_pathway are channels-last in memory (a permuted view), which StageNet consumes
without a copy; any NCHW-contiguous input is converted by a transpose kernel.
"""
import ctypes
import math

import torch
import torch.nn as nn

from . import _lib, packing
from .config import load_args, stage_list, validate_args
from .params import build_fmt, build_stage


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (the B200 hot path has no CPU fallback)")


def to_nhwc(x):
    """[N,C,H,W] (any strides) -> contiguous [N,H,W,C] buffer; zero-copy when x is already channels-last."""
    n, c, h, w = x.shape
    xp = x.permute(0, 2, 3, 1)
    if x.dtype == torch.float32 and xp.is_contiguous():
        return xp
    x = _f32c(x)
    out = torch.empty((n, h, w, c), device=x.device, dtype=torch.float32)
    L = _lib.lib()
    _lib.check(L.mvsf_nchw_to_nhwc(_ptr(x), _ptr(out), n, c, h * w, _stream()), "nchw_to_nhwc")
    return out


def to_nchw(x_nhwc):
    n, h, w, c = x_nhwc.shape
    out = torch.empty((n, c, h, w), device=x_nhwc.device, dtype=torch.float32)
    L = _lib.lib()
    _lib.check(L.mvsf_nhwc_to_nchw(_ptr(x_nhwc), _ptr(out), n, c, h * w, _stream()), "nhwc_to_nchw")
    return out


def split_weights_f16(flat):
    """fp32 weight blob on the device -> fp16 [hi | lo] blob for the tcgen05 GEMMs (install time, once)."""
    L = _lib.lib()
    n = flat.numel()
    assert n % 8 == 0
    out = torch.empty(2 * n, device=flat.device, dtype=torch.float16)
    _lib.check(L.mvsf_split_weights_f16(_ptr(flat), _ptr(out), ctypes.c_size_t(n), _stream()), "split_weights_f16")
    return out


def pack_unet_tc(kind, flat):
    """fp32 U-Net weight blob (packing.pack_costreg_unet) on the device -> fp16 hi/lo weight slabs of the tcgen05
    implicit-GEMM convolutions (install time, once)."""
    L = _lib.lib()
    need = ctypes.c_size_t(0)
    _lib.check(L.mvsf_costreg_unet_tc_bytes(ctypes.byref(need)), "costreg_unet_tc_bytes")
    out = torch.empty(need.value // 2, device=flat.device, dtype=torch.float16)
    _lib.check(L.mvsf_costreg_unet_pack_tc(kind, _ptr(flat), _ptr(out), ctypes.c_size_t(need.value), _stream()),
               "costreg_unet_pack_tc")
    return out


@torch.no_grad()
def homo_warping_3D_with_mask(src_fea, src_proj, ref_proj, depth_values):
    """Drop-in for the reference's finest seam, models/warping.py:69-109 (called at cost_volume.py:72):
    src_fea [B,C,H,W], src_proj / ref_proj [B,4,4] (already composed K@E), depth_values [B,D,H,W] or [B,D]
    -> (warped_src_fea [B,C,D,H,W] fp32, mask [B,D,H,W] bool; True where the sample falls outside the source image or
    behind the camera).  The hot path never materialises this volume (the warp is fused with the correlation);
    this standalone op exists for seam-level parity and for callers of the reference function."""
    _require_cuda(src_fea, "homo_warping_3D_with_mask(src_fea)")
    B, C, H, W = src_fea.shape
    D = depth_values.shape[1]
    dev = src_fea.device
    L = _lib.lib()
    st = _stream()
    depth_values = _f32c(depth_values.to(dev))
    if depth_values.dim() == 2:  # warping.py:73-74
        depth_values = depth_values.view(B, D, 1, 1).expand(B, D, H, W).contiguous()
    sp, rp = _f32c(src_proj.to(dev)), _f32c(ref_proj.to(dev))
    homs = torch.empty((B, 12), device=dev, dtype=torch.float32)
    _lib.check(L.mvsf_homography_from_proj(_ptr(sp), _ptr(rp), B, _ptr(homs), st), "homography_from_proj")
    src = to_nhwc(src_fea)
    warped = torch.empty((B, C, D, H, W), device=dev, dtype=torch.float32)
    mask = torch.empty((B, D, H, W), device=dev, dtype=torch.uint8)
    for b in range(B):
        _lib.check(L.mvsf_homo_warp(_ptr(src[b]), _ptr(homs[b]), _ptr(depth_values[b]), _ptr(warped[b]), _ptr(mask[b]),
                                    C, D, H, W, st), "homo_warp")
    return warped, mask.bool()


class _PackedMixin:
    """Packs the module's parameters for the CUDA library on first use / after load_state_dict."""

    def _invalidate(self, *a, **k):
        self._packed = None

    def repack(self):
        """Drop the packed (BN-folded, fp16-split) weight blobs; they are rebuilt on the next forward.  The cache is
        invalidated automatically by load_state_dict() and .to()/.cuda(); call this after in-place parameter edits
        (param.data.copy_, optimiser steps)."""
        self._packed = None

    def _init_packing(self):
        self._packed = None
        self.register_load_state_dict_post_hook(lambda m, k: m._invalidate())

    def _apply(self, fn, *a, **k):  # .to()/.cuda() move parameters: repack lazily
        self._packed = None
        return super()._apply(fn, *a, **k)


# =====================================================================================================
class StageNet(_PackedMixin, nn.Module):
    """Drop-in for the reference StageNet (models/cost_volume.py:21-133): same constructor arguments, parameter
    names, forward signature and output dict; eval-mode arithmetic (BatchNorm folded, depth_type 'ce')."""

    def __init__(self, args, ndepth, stage_idx):
        super().__init__()
        self.args = args
        self.fusion_type = args.get("fusion_type", "cnn")
        if self.fusion_type != "cnn":
            raise NotImplementedError(f"Not implemented fusion type: {self.fusion_type}.")
        self.ndepth = ndepth
        self.stage_idx = stage_idx
        self.cost_reg_type = args.get("cost_reg_type", ["Normal"] * 4)[stage_idx]
        self.depth_type = stage_list(args["depth_type"], stage_idx)
        bag = build_stage(args, ndepth, stage_idx)
        self.vis = bag.vis
        self.cost_reg = bag.cost_reg
        # per-view group correlations spilled between the two cost-volume passes; above this many bytes the stage
        # recomputes the gather in pass B instead (no spill buffer), so large inputs degrade instead of running out of memory
        self.corr_spill_budget_bytes = int(args.get("corr_spill_budget_bytes", 8 << 30))
        self._init_packing()

    # ---- packing
    def _pack(self, device):
        if self._packed is not None and self._packed["device"] == device:
            return self._packed
        sd = {k: v for k, v in self.state_dict().items()}
        pk = {"device": device, "vis": packing.pack_vis(sd, "vis.").to(device)}
        if self.cost_reg_type == "PureTransformerCostReg":
            tc = self.args["transformer_config"][self.stage_idx]
            if tuple(tc["down_rate"]) != (2, 4, 4) or tc["mid_channel"] != 64 or tc["num_heads"] != 4 or tc["mlp_ratio"] != 4:
                raise NotImplementedError("transformer regulariser: only the shipped geometry (down_rate (2,4,4), "
                                          "mid 64, 4 heads, mlp_ratio 4) is implemented")
            pk["kind"] = "tr"
            pk["layers"] = tc["layer_num"]
            pk["reg"] = packing.pack_costreg_tr(sd, "cost_reg.", tc["layer_num"]).to(device)
            pk["reg16"] = split_weights_f16(pk["reg"])
        else:
            kind, flat = packing.pack_costreg_unet(sd, "cost_reg.")
            pk["kind"] = kind
            pk["reg"] = flat.to(device)
            pk["reg_tc"] = pack_unet_tc(kind, pk["reg"])
        self._packed = pk
        return pk

    def _softmax_scale(self, n_tokens):
        tc = self.args["transformer_config"][self.stage_idx]
        scale = (tc["mid_channel"] // tc["num_heads"]) ** -0.5
        if tc.get("softmax_scale", None) is not None:  # attention.py:158-161
            scale *= math.log(n_tokens, tc["train_avg_length"])
        return scale

    # ---- one sample
    def _forward_one(self, feat_nhwc, proj, depth_values, tmp, position3d, pk, keep=False):
        L = _lib.lib()
        st = _stream()
        V, H, W, C = feat_nhwc.shape
        D = depth_values.shape[0]
        dev = feat_nhwc.device
        G = stage_list(self.args["base_ch"], self.stage_idx)
        if G > C:
            raise AssertionError("G must <= C!")
        f32 = dict(device=dev, dtype=torch.float32)
        homs = torch.empty((V - 1) * 12, **f32)
        kinv = torch.empty(9, **f32)
        _lib.check(L.mvsf_compose_geometry(_ptr(proj), V, _ptr(homs), _ptr(kinv), st), "compose_geometry")
        entropy = torch.empty((V - 1, H, W), **f32)
        vis = torch.empty((V - 1, H, W), **f32)
        volume = torch.empty((D, H, W, G), **f32)
        # spill plan (pass A stores the per-view group correlations, the aggregation streams them) unless the buffer would
        # exceed the budget: then both passes gather (TMA-staged window kernels at C = 8 / 16), no intermediate buffer
        two_gathers = L.mvsf_warp_corr_plan(C, G, D, H, W, V, ctypes.c_size_t(self.corr_spill_budget_bytes)) == 1
        if not two_gathers:
            # pass A also stores the per-view group correlations; the view aggregation then streams them (no second gather)
            corr = torch.empty((V - 1, D, H, W, G), **f32)
            _lib.check(L.mvsf_warp_corr_entropy_store(_ptr(feat_nhwc), _ptr(homs), _ptr(depth_values), _ptr(entropy),
                                                      _ptr(corr), V, C, G, D, H, W, st), "warp_corr_entropy_store")
            _lib.check(L.mvsf_vis_cnn(_ptr(entropy), _ptr(pk["vis"]), _ptr(vis), V - 1, H, W, st), "vis_cnn")
            _lib.check(L.mvsf_corr_aggregate(_ptr(corr), _ptr(vis), _ptr(volume), V, G, D, H, W, st), "corr_aggregate")
            del corr
        else:
            _lib.check(L.mvsf_warp_corr_entropy(_ptr(feat_nhwc), _ptr(homs), _ptr(depth_values), _ptr(entropy),
                                                V, C, G, D, H, W, st), "warp_corr_entropy")
            _lib.check(L.mvsf_vis_cnn(_ptr(entropy), _ptr(pk["vis"]), _ptr(vis), V - 1, H, W, st), "vis_cnn")
            _lib.check(L.mvsf_warp_corr_aggregate(_ptr(feat_nhwc), _ptr(homs), _ptr(depth_values), _ptr(vis),
                                                  _ptr(volume), V, C, G, D, H, W, st), "warp_corr_aggregate")
        kept = dict(entropy=entropy, vis_weight=vis, volume_mean=volume.clone() if keep else None) if keep else None
        logits = torch.empty((D, H, W), **f32)
        need = ctypes.c_size_t(0)
        if pk["kind"] == "tr":
            _lib.check(L.mvsf_costreg_tr_workspace_bytes(G, D, H, W, ctypes.byref(need)), "costreg_tr_workspace_bytes")
            ws = torch.empty(need.value // 4 + 4, **f32)
            n_tok = (D // 2) * (H // 4) * (W // 4)
            _lib.check(L.mvsf_costreg_tr_forward(_ptr(volume), _ptr(position3d), _ptr(pk["reg"]), _ptr(pk["reg16"]),
                                                 ctypes.c_size_t(pk["reg"].numel()), _ptr(logits),
                                                 _ptr(ws), ctypes.c_size_t(ws.numel() * 4), G, D, H, W, pk["layers"],
                                                 float(self._softmax_scale(n_tok)), st), "costreg_tr_forward")
        else:
            _lib.check(L.mvsf_costreg_unet_workspace_bytes(pk["kind"], G, D, H, W, ctypes.byref(need)),
                       "costreg_unet_workspace_bytes")
            ws = torch.empty(need.value // 4 + 4, **f32)
            _lib.check(L.mvsf_costreg_unet_forward(pk["kind"], _ptr(volume), _ptr(pk["reg"]), _ptr(pk["reg_tc"]),
                                                   _ptr(logits), _ptr(ws),
                                                   ctypes.c_size_t(ws.numel() * 4), G, D, H, W, st),
                       "costreg_unet_forward")
        prob = torch.empty((D, H, W), **f32)
        depth = torch.empty((H, W), **f32)
        conf = torch.empty((H, W), **f32)
        _lib.check(L.mvsf_softargmax(_ptr(logits), _ptr(depth_values), float(tmp), _ptr(prob), _ptr(depth), _ptr(conf),
                                     D, H, W, st), "softargmax")
        return depth, prob, conf, logits, kept

    @torch.no_grad()
    def forward(self, features, proj_matrices, depth_values, tmp, position3d=None, keep_intermediates=False):
        if self.training:
            raise NotImplementedError("B200 hot path implements the eval-mode forward (test.py); call .eval()")
        if self.depth_type != "ce":
            raise NotImplementedError("depth_type must be 'ce'")
        _require_cuda(features, "StageNet.forward(features)")
        B, V, C, H, W = features.shape
        if V != proj_matrices.shape[1]:
            raise AssertionError("Different number of images and projection matrices")
        pk = self._pack(features.device)
        proj_matrices = _f32c(proj_matrices)
        depth_values = _f32c(depth_values)
        if depth_values.dim() == 2:  # [B,D] -> per-pixel hypotheses (warping.py:73 accepts both)
            depth_values = depth_values.view(B, -1, 1, 1).expand(B, depth_values.shape[1], H, W).contiguous()
        if position3d is not None:
            position3d = _f32c(position3d)
        outs = []
        for b in range(B):
            feat = to_nhwc(features[b])
            p3 = position3d[b] if position3d is not None else None
            outs.append(self._forward_one(feat, proj_matrices[b], depth_values[b], tmp, p3, pk, keep_intermediates))
        stack = (lambda i: torch.stack([o[i] for o in outs], 0)) if B > 1 else (lambda i: outs[0][i].unsqueeze(0))
        out = {"depth": stack(0), "prob_volume": stack(1), "photometric_confidence": stack(2),
               "depth_values": depth_values, "prob_volume_pre": stack(3)}
        if keep_intermediates:
            for k in ("entropy", "vis_weight", "volume_mean"):
                out[k] = torch.stack([o[4][k] for o in outs], 0)
        return out


# =====================================================================================================
class FMT_with_pathway(_PackedMixin, nn.Module):
    """Drop-in for the reference FMT_with_pathway (models/FMT.py:140-206)."""

    def __init__(self, base_channel=8, **kwargs):
        super().__init__()
        cfg = dict(kwargs)
        cfg["base_channel"] = base_channel
        if cfg.get("d_model", 64) != 64 or cfg.get("nhead", 4) != 4 or base_channel != 8:
            raise NotImplementedError("FMT: only the shipped geometry (d_model 64, 4 heads, base_channel 8)")
        if cfg.get("attention_type", "Linear") != "Linear":
            raise NotImplementedError("Unkown attention type", cfg.get("attention_type"))
        self.cfg = cfg
        bag = build_fmt(cfg)
        self.FMT = bag.FMT
        for k in (1, 2, 3):
            setattr(self, f"dim_reduction_{k}", getattr(bag, f"dim_reduction_{k}"))
            setattr(self, f"smooth_{k}", getattr(bag, f"smooth_{k}"))
        self.pe_dict = {}
        self._init_packing()

    def _pack(self, device):
        if self._packed is None or self._packed["device"] != device:
            sd = {"FMT_module." + k: v for k, v in self.state_dict().items()}
            w = packing.pack_fmt(sd).to(device)
            self._packed = {"device": device, "w": w, "w16": split_weights_f16(w)}
        return self._packed

    def _pe(self, H, W, device):
        """PositionEncodingSineNorm table (position_encoding.py:61-74) as [H*W, 64], cached per shape like the
        reference's pe_dict (constant per resolution; built with the same torch ops as the reference)."""
        key = (H, W, str(device))
        if key not in self.pe_dict:
            d_model, max_shape = 64, (128, 128)
            pe = torch.zeros((d_model, H, W))
            ypos = torch.ones((H, W)).cumsum(0).float().unsqueeze(0) * max_shape[0] / H
            xpos = torch.ones((H, W)).cumsum(1).float().unsqueeze(0) * max_shape[1] / W
            div = torch.exp(torch.arange(0, d_model // 2, 2).float() * (-math.log(10000.0) / (d_model // 2)))[:, None, None]
            pe[0::4] = torch.sin(xpos * div)
            pe[1::4] = torch.cos(xpos * div)
            pe[2::4] = torch.sin(ypos * div)
            pe[3::4] = torch.cos(ypos * div)
            self.pe_dict[key] = pe.permute(1, 2, 0).reshape(H * W, d_model).contiguous().to(device)
        return self.pe_dict[key]

    @torch.no_grad()
    def forward(self, features):
        f1 = features["stage1"]
        _require_cuda(f1, "FMT_with_pathway.forward(features)")
        B, V, C, H1, W1 = f1.shape
        assert C == 64, "FMT d_model must equal the stage-1 channel count"
        L = _lib.lib()
        pk = self._pack(f1.device)
        pe = self._pe(H1, W1, f1.device)
        f32 = dict(device=f1.device, dtype=torch.float32)
        need = ctypes.c_size_t(0)
        _lib.check(L.mvsf_fmt_workspace_bytes(V, H1, W1, ctypes.byref(need)), "fmt_workspace_bytes")
        ws = torch.empty(need.value // 4 + 4, **f32)
        outs = {k: [] for k in ("stage1", "stage2", "stage3", "stage4")}
        for b in range(B):
            ins = [_f32c(features[f"stage{k}"][b]) for k in (1, 2, 3, 4)]
            for k, (c, sc) in enumerate(((64, 1), (32, 2), (16, 4), (8, 8))):
                if tuple(ins[k].shape) != (V, c, H1 * sc, W1 * sc):
                    raise AssertionError(f"stage{k + 1} features must be [V,{c},{H1 * sc},{W1 * sc}], got {tuple(ins[k].shape)}")
            o = [torch.empty((V, H1 * sc, W1 * sc, c), **f32) for c, sc in ((64, 1), (32, 2), (16, 4), (8, 8))]
            _lib.check(L.mvsf_fmt_forward(_ptr(ins[0]), _ptr(ins[1]), _ptr(ins[2]), _ptr(ins[3]), _ptr(pe), _ptr(pk["w"]),
                                          _ptr(pk["w16"]), ctypes.c_size_t(pk["w"].numel()),
                                          _ptr(o[0]), _ptr(o[1]), _ptr(o[2]), _ptr(o[3]), _ptr(ws),
                                          ctypes.c_size_t(ws.numel() * 4), V, H1, W1, _stream()), "fmt_forward")
            for k in range(4):
                outs[f"stage{k + 1}"].append(o[k])
        # logical [B,V,C,H,W]; channels-last in memory (StageNet consumes it without a copy)
        return {k: (v[0].unsqueeze(0) if B == 1 else torch.stack(v, 0)).permute(0, 1, 4, 2, 3) for k, v in outs.items()}


# =====================================================================================================
class HotPathNet(nn.Module):
    """FMT + 4-stage cascade from the FPN feature pyramid to the output dict: the part of
    DINOv2MVSNet.forward after feature extraction (DINOv2_mvsformer_model.py:117-179).  Sub-module names
    (FMT_module, fusions) and parameter names equal the reference's, so the hot-path subset of a reference
    checkpoint loads with load_state_dict(strict=True)."""

    def __init__(self, args):
        super().__init__()
        self.args = validate_args(load_args(args))
        a = self.args
        self.ndepths = a["ndepths"]
        self.depth_interals_ratio = a["depth_interals_ratio"]
        self.cost_reg_type = a.get("cost_reg_type", ["Normal"] * 4)
        self.use_pe3d = a.get("use_pe3d", False)
        self.FMT_module = FMT_with_pathway(**a["FMT_config"])
        self.fusions = nn.ModuleList([StageNet(a, self.ndepths[i], i) for i in range(len(self.ndepths))])

    @torch.no_grad()
    def forward_features(self, features, proj_matrices, depth_values, tmp=(5.0, 5.0, 5.0, 1.0), run_fmt=True,
                         keep_intermediates=False):
        return cascade_forward(self.FMT_module if run_fmt else None, self.fusions, self.args, features, proj_matrices,
                               depth_values, tmp, keep_intermediates)

    forward = forward_features


def cascade_forward(fmt_module, fusions, args, features, proj_matrices, depth_values, tmp, keep_intermediates=False):
    """DINOv2_mvsformer_model.py:117-179 with the element-wise glue (hypothesis scheduling, 3-D positions, confidence
    averaging) as CUDA kernels.  `fusions[i]` must be this package's StageNet."""
    L = _lib.lib()
    ndepths, ratios = args["ndepths"], args["depth_interals_ratio"]
    if fmt_module is not None:
        features = fmt_module.forward(features)
    last = features[f"stage{len(ndepths)}"]
    _require_cuda(last, "features")
    B, Hf, Wf = last.shape[0], last.shape[3], last.shape[4]
    dev = last.device
    f32 = dict(device=dev, dtype=torch.float32)
    depth_values = _f32c(depth_values.to(dev))
    Dn = depth_values.shape[1]
    prob_maps = torch.empty((B, Hf, Wf), **f32)
    stats = torch.zeros(8, **f32)   # x/y extents of the 3-D positions + depth range, shared by the whole batch
    have_extents = False            # the reference computes them on the first stage that builds the PE and reuses them
    outputs, so = {}, None
    for s in range(len(ndepths)):
        pm = _f32c(proj_matrices[f"stage{s + 1}"].to(dev))
        f = features[f"stage{s + 1}"]
        _, V, C, H, W = f.shape
        D = ndepths[s]
        ds = torch.empty((B, D, H, W), **f32)
        st = _stream()
        for b in range(B):
            if s == 0:
                _lib.check(L.mvsf_init_inverse_range(_ptr(depth_values[b]), Dn, _ptr(ds[b]), D, H, W, st),
                           "init_inverse_range")
            else:
                _lib.check(L.mvsf_schedule_inverse_range(_ptr(so["depth"][b]), _ptr(so["depth_values"][b]),
                                                         so["depth_values"].shape[1], float(ratios[s]), _ptr(ds[b]),
                                                         D, H, W, st), "schedule_inverse_range")
        p3d = None
        if args["cost_reg_type"][s] != "Normal" and args.get("use_pe3d", False):
            # position_encoding.py:138-161: extents (first PE stage only) and depth_values.min()/max() are reductions
            # over the WHOLE batch (DINOv2_mvsformer_model.py:152-160)
            p3d = torch.empty((B, 3, D, H, W), **f32)
            kinvs = torch.empty((B, 9), **f32)
            homs = torch.empty((V - 1) * 12, **f32)
            for b in range(B):
                _lib.check(L.mvsf_compose_geometry(_ptr(pm[b]), V, _ptr(homs), _ptr(kinvs[b]), st), "compose_geometry")
                if not have_extents:
                    _lib.check(L.mvsf_position3d(_ptr(kinvs[b]), _ptr(ds[b]), None, 0, _ptr(stats), 2 if b == 0 else 3,
                                                 None, D, H, W, st), "position3d(extents)")
            if not have_extents:
                _lib.check(L.mvsf_position3d(None, None, _ptr(depth_values), B * Dn, _ptr(stats), 4, None, D, H, W, st),
                           "position3d(finalize)")
                have_extents = True
            for b in range(B):
                _lib.check(L.mvsf_position3d(_ptr(kinvs[b]), _ptr(ds[b]), None, 0, _ptr(stats), 5, _ptr(p3d[b]),
                                             D, H, W, st), "position3d(normalise)")
        so = fusions[s].forward(f, pm, ds, tmp=tmp[s], position3d=p3d, keep_intermediates=keep_intermediates)
        outputs[f"stage{s + 1}"] = so
        conf = so["photometric_confidence"]
        for b in range(B):
            _lib.check(L.mvsf_conf_accumulate(_ptr(conf[b]), H, W, _ptr(prob_maps[b]), Hf, Wf, 1.0 / len(ndepths),
                                              1 if s == 0 else 0, st), "conf_accumulate")
        outputs.update(so)
    outputs["refined_depth"] = so["depth"]
    outputs["photometric_confidence"] = prob_maps
    outputs["features"] = features
    return outputs


# =====================================================================================================
def install(model, args=None):
    """Rebinds the hot-path seams of a reference-constructed DINOv2MVSNet (models/networks/DINOv2_mvsformer_model.py)
    to the CUDA path: model.FMT_module and model.fusions[i] are replaced by this package's modules carrying the
    same weights (state_dict round trip, strict).  The rest of the model (ViT, FPN) is untouched.  Returns model."""
    args = validate_args(load_args(args if args is not None else model.args))
    dev = next(model.parameters()).device
    fmt = FMT_with_pathway(**args["FMT_config"])
    fmt.load_state_dict(model.FMT_module.state_dict(), strict=True)
    model.FMT_module = fmt.to(dev).eval()
    new = []
    for i, old in enumerate(model.fusions):
        st = StageNet(args, args["ndepths"][i], i)
        st.load_state_dict(old.state_dict(), strict=True)
        new.append(st.to(dev).eval())
    model.fusions = nn.ModuleList(new)
    return model
