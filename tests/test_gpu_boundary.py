"""Drop-in boundary on the GPU: install() on a model object shaped like the reference's DINOv2MVSNet,
driven the way test.py drives it (test.py:209-251: strict state-dict load, .to(device), .eval(),
`with torch.cuda.amp.autocast(dtype=torch.bfloat16): model.forward(...)`), and the finest seam
homo_warping_3D_with_mask with the reference's Python signature (models/warping.py:69-109)."""
import pytest
import torch
import torch.nn as nn

from tests.common import TMP, load_golden, max_abs, rec, rel_linf

pytestmark = pytest.mark.gpu


class StubMVSNet(nn.Module):
    """What install() touches of models/networks/DINOv2_mvsformer_model.py:DINOv2MVSNet: `args`, `FMT_module`, `fusions`
    (parameter containers with the reference's key names) and the part of forward() after feature extraction
    (:117-179), restated with the reference's own glue ops (oracle restatements of init_inverse_range /
    schedule_inverse_range / get_position_3d evaluated with torch on the host, F.interpolate for the confidence)."""

    def __init__(self, args):
        super().__init__()
        from mvsformerplusplus_b200.params import build_hotpath_params
        self.args = args
        p = build_hotpath_params(args)
        self.FMT_module, self.fusions = p.FMT_module, p.fusions
        self.ndepths, self.ratios = args["ndepths"], args["depth_interals_ratio"]

    def forward(self, features, proj_matrices, depth_values, tmp):
        import torch.nn.functional as F
        from oracle import hotpath as O
        dev = depth_values.device
        features = self.FMT_module.forward(features)                                     # :117
        Hf, Wf = features["stage4"].shape[-2:]
        prob_maps = torch.zeros(depth_values.shape[0], Hf, Wf, device=dev)
        outputs, so = {}, {}
        hmin = hmax = wmin = wmax = None
        for s in range(4):
            pm = proj_matrices[f"stage{s + 1}"]
            f = features[f"stage{s + 1}"]
            B, V, C, H, W = f.shape
            if s == 0:
                ds = O.init_inverse_range(depth_values.cpu(), self.ndepths[s], H, W)
            else:
                ds = O.schedule_inverse_range(so["depth"].float().cpu(), so["depth_values"].float().cpu(), self.ndepths[s],
                                              self.ratios[s], H, W)
            p3d = None
            if self.args["cost_reg_type"][s] != "Normal" and self.args["use_pe3d"]:
                p3d, hmin, hmax, wmin, wmax = O.get_position_3d(B, H, W, pm[:, 0, 1, :3, :3].cpu(), ds, depth_values.min().cpu(),
                                                                depth_values.max().cpu(), hmin, hmax, wmin, wmax)
                p3d = p3d.to(dev)
            so = self.fusions[s].forward(f, pm, ds.to(dev), tmp=tmp[s], position3d=p3d)   # :164
            outputs[f"stage{s + 1}"] = so
            conf = so["photometric_confidence"]
            if conf.shape[1] != Hf or conf.shape[2] != Wf:
                conf = F.interpolate(conf.unsqueeze(1), [Hf, Wf], mode="nearest").squeeze(1)
            prob_maps += conf
            outputs.update(so)
        outputs["refined_depth"] = so["depth"]
        outputs["photometric_confidence"] = prob_maps / 4
        return outputs


def test_install_on_reference_shaped_model_under_bf16_autocast():
    from mvsformerplusplus_b200 import hotpath, synth
    from mvsformerplusplus_b200.config import default_args
    dev = torch.device("cuda:0")
    args = default_args()
    torch.manual_seed(0)
    model = StubMVSNet(args)
    sd = synth.randomize_state_dict(model, seed=31)
    keys_before = sorted(model.state_dict().keys())
    model = model.to(dev).eval()
    with pytest.raises(RuntimeError, match="parameter container"):   # before install() the stub cannot compute anything
        model.fusions[0].forward(None, None, None, tmp=1.0)
    hotpath.install(model)                                            # the one line INTEGRATION.md adds to test.py
    assert isinstance(model.FMT_module, hotpath.FMT_with_pathway) and all(isinstance(f, hotpath.StageNet) for f in model.fusions)
    assert sorted(model.state_dict().keys()) == keys_before           # checkpoint-compatible before and after
    for k, v in model.state_dict().items():
        assert torch.equal(v.cpu(), sd[k]), k
    V, H, W = 3, 64, 96
    feats = {k: v.to(dev) for k, v in synth.make_features(V, H, W, seed=5).items()}
    proj = {k: v.to(dev) for k, v in synth.make_proj_matrices(V, H, W, theta_step=0.12).items()}
    dv = synth.make_depth_values(48, 425.0, 2.65 * 4).to(dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):   # test.py:250
        out = model.forward(feats, proj, dv, TMP)
        # under autocast the FPN hands the hot path bf16 feature maps: they must be accepted (and promoted) too
        out_bf = model.forward({k: v.bfloat16() for k, v in feats.items()}, proj, dv, TMP)
    net = hotpath.HotPathNet(args).eval()
    net.load_state_dict(sd, strict=True)
    want = net.to(dev).forward_features(feats, proj, dv, TMP)
    e = dict(depth_rel=rel_linf(out["refined_depth"].cpu(), want["refined_depth"].cpu()),
             conf=max_abs(out["photometric_confidence"].cpu(), want["photometric_confidence"].cpu()),
             prob4=max_abs(out["stage4"]["prob_volume"].cpu(), want["stage4"]["prob_volume"].cpu()))
    rec("install_stub_autocast_bf16", **e)
    assert out["refined_depth"].dtype == torch.float32 and out["refined_depth"].shape == (1, H, W)
    assert e["depth_rel"] < 1e-4 and e["conf"] < 1e-4 and e["prob4"] < 1e-4
    assert torch.isfinite(out_bf["refined_depth"]).all() and out_bf["refined_depth"].dtype == torch.float32
    for k in ("refined_depth", "photometric_confidence"):   # outputs survive test.py's tensor2numpy (utils.py:63-74)
        assert out[k].detach().cpu().numpy().dtype.name == "float32"


def test_batch_of_two_reduces_extents_over_the_batch():
    """B = 2 with different depth ranges per sample: the 3-D position extents and depth_values.min()/max() are reductions
    over the whole batch (position_encoding.py:152-157, DINOv2_mvsformer_model.py:156), not per sample."""
    from mvsformerplusplus_b200 import hotpath, synth
    from mvsformerplusplus_b200.config import default_args
    from oracle import hotpath as O
    dev = torch.device("cuda:0")
    args = default_args()
    torch.manual_seed(0)
    net = hotpath.HotPathNet(args).eval()
    sd = synth.randomize_state_dict(net, seed=19)
    net = net.to(dev)
    V, H, W = 3, 64, 96
    feats = synth.make_features(V, H, W, seed=8, batch=2)
    proj = synth.make_proj_matrices(V, H, W, batch=2, theta_step=0.12)
    dv = torch.cat([synth.make_depth_values(48, 425.0, 2.65 * 4), synth.make_depth_values(48, 500.0, 2.2 * 4)], 0)
    out = net.forward_features({k: v.to(dev) for k, v in feats.items()}, {k: v.to(dev) for k, v in proj.items()}, dv.to(dev), TMP)
    with torch.no_grad():
        want = O.hotpath_forward(feats, proj, dv, sd, args, tmp=TMP)
    e = dict(depth_rel=rel_linf(out["refined_depth"].cpu(), want["refined_depth"]),
             conf=max_abs(out["photometric_confidence"].cpu(), want["photometric_confidence"]),
             prob1=max_abs(out["stage1"]["prob_volume"].cpu(), want["stage1"]["prob_volume"]))
    rec("batch2_cascade", **e)
    assert e["depth_rel"] < 1e-3 and e["conf"] < 1e-4 and e["prob1"] < 1e-4


def test_first_pe_stage_is_not_stage_one():
    """cost_reg_type[0] == 'Normal' with the transformer (and its 3-D PE) on stage 2: the extents come from the first stage
    that builds the PE (height_min is None there, DINOv2_mvsformer_model.py:152-160)."""
    from mvsformerplusplus_b200 import hotpath, synth
    from mvsformerplusplus_b200.config import default_args
    from oracle import hotpath as O
    dev = torch.device("cuda:0")
    args = default_args()
    args["cost_reg_type"] = ["Normal", "PureTransformerCostReg", "Normal", "Normal"]
    args["transformer_config"] = [args["transformer_config"][0]] * 2
    args["ndepths"] = [8, 8, 8, 4]
    torch.manual_seed(0)
    net = hotpath.HotPathNet(args).eval()
    sd = synth.randomize_state_dict(net, seed=23)
    net = net.to(dev)
    V, H, W = 3, 64, 64
    feats = synth.make_features(V, H, W, seed=3)
    proj = synth.make_proj_matrices(V, H, W, theta_step=0.12)
    dv = synth.make_depth_values(48, 425.0, 2.65 * 4)
    out = net.forward_features({k: v.to(dev) for k, v in feats.items()}, {k: v.to(dev) for k, v in proj.items()}, dv.to(dev), TMP)
    with torch.no_grad():
        want = O.hotpath_forward(feats, proj, dv, sd, args, tmp=TMP)
    e = dict(depth_rel=rel_linf(out["refined_depth"].cpu(), want["refined_depth"]),
             prob2=max_abs(out["stage2"]["prob_volume"].cpu(), want["stage2"]["prob_volume"]))
    rec("pe_on_stage2", **e)
    assert e["depth_rel"] < 1e-3 and e["prob2"] < 1e-4


def test_homo_warping_seam_python_signature():
    from mvsformerplusplus_b200.hotpath import homo_warping_3D_with_mask
    dev = torch.device("cuda:0")
    g, _ = load_golden("warp_seam")   # produced by the reference's own homo_warping_3D_with_mask (oracle/gen_golden.py)
    warped, mask = homo_warping_3D_with_mask(g["src"].to(dev), g["src_proj"].to(dev), g["ref_proj"].to(dev),
                                             g["depth_values"].to(dev))
    assert warped.shape == g["warped"].shape and mask.shape == g["mask"].shape and mask.dtype == torch.bool
    e = max_abs(warped.cpu(), g["warped"])
    mm = float((mask.cpu() != g["mask"]).float().mean())
    rec("homo_warp_seam_python", abs=e, mask_mismatch=mm)
    assert e < 2e-4 and mm < 0.01
    # [B, D] hypotheses (warping.py:73-74) broadcast over the image
    dv2 = g["depth_values"][:, :, 0, 0].contiguous()
    w2, m2 = homo_warping_3D_with_mask(g["src"].to(dev), g["src_proj"].to(dev), g["ref_proj"].to(dev), dv2.to(dev))
    B, D, H, W = g["depth_values"].shape
    w3, _ = homo_warping_3D_with_mask(g["src"].to(dev), g["src_proj"].to(dev), g["ref_proj"].to(dev),
                                      dv2.view(B, D, 1, 1).expand(B, D, H, W).to(dev))
    assert torch.equal(w2, w3)


def test_stage_without_spill_buffer_matches_spill_plan():
    """corr_spill_budget_bytes = 0 makes StageNet run the cost volume as two gathers (window kernels at C = 8 / 16, the
    L1-gather kernels at C = 32 / 64) instead of spilling the per-view correlations: same outputs, no [(V-1),D,H,W,8] buffer."""
    from mvsformerplusplus_b200 import hotpath, synth
    from mvsformerplusplus_b200.config import default_args
    dev = torch.device("cuda:0")
    V, H, W = 4, 96, 128
    feats = {k: v.to(dev) for k, v in synth.make_features(V, H, W, seed=2).items()}
    proj = {k: v.to(dev) for k, v in synth.make_proj_matrices(V, H, W, theta_step=0.12).items()}
    dv = synth.make_depth_values(48, 425.0, 2.65 * 4).to(dev)
    outs = []
    for budget in (None, 0):
        args = default_args()
        if budget is not None:
            args["corr_spill_budget_bytes"] = budget
        torch.manual_seed(0)
        net = hotpath.HotPathNet(args).eval()
        synth.randomize_state_dict(net, seed=29)
        outs.append(net.to(dev).forward_features(feats, proj, dv, TMP))
    e = dict(depth_rel=rel_linf(outs[1]["refined_depth"].cpu(), outs[0]["refined_depth"].cpu()),
             prob4=max_abs(outs[1]["stage4"]["prob_volume"].cpu(), outs[0]["stage4"]["prob_volume"].cpu()),
             prob1=max_abs(outs[1]["stage1"]["prob_volume"].cpu(), outs[0]["stage1"]["prob_volume"].cpu()))
    rec("two_gather_plan_vs_spill_plan", **e)
    assert e["depth_rel"] < 1e-4 and e["prob4"] < 5e-5 and e["prob1"] < 5e-5
