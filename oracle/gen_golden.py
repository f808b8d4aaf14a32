"""ORACLE - TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by executing the REFERENCE's own
modules (imported read-only from /root/reference, which exists only in the build container) on seeded
synthetic inputs.  The fixtures pin oracle/hotpath.py (tests/test_oracle_golden.py) and, on the GPU box,
the CUDA path (tests/test_gpu_parity.py).  Re-run:  python oracle/gen_golden.py

The reference has no tests/golden vectors of its own, so these reference-executed outputs
are the pin.  Reference code is imported, never copied.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MVSF_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)

from mvsformerplusplus_b200 import synth  # noqa: E402
from mvsformerplusplus_b200.config import default_args  # noqa: E402


def import_reference():
    import models.dino.layers.attention as A
    A.FLASH_AVAILABLE = False  # CPU: SDPA fallback with the same scale (attention.py:82-96,144-146)
    from models.FMT import FMT_with_pathway
    from models.cost_volume import StageNet
    from models.module import init_inverse_range, schedule_inverse_range
    from models.position_encoding import get_position_3d
    from models.warping import homo_warping_3D_with_mask
    return dict(FMT=FMT_with_pathway, StageNet=StageNet, init_inverse_range=init_inverse_range,
                schedule_inverse_range=schedule_inverse_range, get_position_3d=get_position_3d,
                homo_warp=homo_warping_3D_with_mask)


class RefHotPath(nn.Module):
    """Reference modules under the reference's attribute names (DINOv2_mvsformer_model.py:41,53)."""

    def __init__(self, R, args):
        super().__init__()
        self.FMT_module = R["FMT"](**args["FMT_config"])
        self.fusions = nn.ModuleList([R["StageNet"](args, args["ndepths"][i], i) for i in range(len(args["ndepths"]))])


def reference_hotpath(R, model, args, features, proj_matrices, depth_values, tmp, run_fmt=True):
    """Glue of DINOv2_mvsformer_model.py:117-179, calling the reference's functions/modules only."""
    cap = {}

    def hook_vis(s):
        def fn(mod, inp, out):
            cap.setdefault(f"stage{s + 1}.entropy", []).append(inp[0][:, 0].clone())
            cap.setdefault(f"stage{s + 1}.vis_weight", []).append(out[:, 0].clone())
        return fn

    def hook_reg(s):
        def fn(mod, inp):
            cap[f"stage{s + 1}.volume_mean"] = inp[0].clone()
        return fn

    hs = []
    for s, st in enumerate(model.fusions):
        hs.append(st.vis.register_forward_hook(hook_vis(s)))
        hs.append(st.cost_reg.register_forward_pre_hook(hook_reg(s)))
    with torch.no_grad():
        if run_fmt:
            features = model.FMT_module.forward(features)
        ndepths, ratios = args["ndepths"], args["depth_interals_ratio"]
        Hf, Wf = features[f"stage{len(ndepths)}"].shape[-2:]
        B = depth_values.shape[0]
        prob_maps = torch.zeros(B, Hf, Wf)
        outputs, so = {}, {}
        hmin = hmax = wmin = wmax = None
        for s in range(len(ndepths)):
            pm = proj_matrices[f"stage{s + 1}"]
            f = features[f"stage{s + 1}"]
            _, _, C, H, W = f.shape
            if s == 0:
                ds = R["init_inverse_range"](depth_values, ndepths[s], f.device, f.dtype, H, W)
            else:
                ds = R["schedule_inverse_range"](so["depth"], so["depth_values"], ndepths[s], ratios[s], H, W)
            p3d = None
            if args["cost_reg_type"][s] != "Normal" and args["use_pe3d"]:
                p3d, hmin, hmax, wmin, wmax = R["get_position_3d"](
                    B, H, W, pm[:, 0, 1, :3, :3], ds, depth_min=depth_values.min(), depth_max=depth_values.max(),
                    height_min=hmin, height_max=hmax, width_min=wmin, width_max=wmax, normalize=True)
            so = model.fusions[s].forward(f, pm, ds, tmp=tmp[s], position3d=p3d)
            outputs[f"stage{s + 1}"] = so
            conf = so["photometric_confidence"]
            if conf.shape[1] != Hf or conf.shape[2] != Wf:
                conf = F.interpolate(conf.unsqueeze(1), [Hf, Wf], mode="nearest").squeeze(1)
            prob_maps += conf
        outputs["refined_depth"] = so["depth"]
        outputs["photometric_confidence"] = prob_maps / len(ndepths)
    for h in hs:
        h.remove()
    for k in list(cap):
        if isinstance(cap[k], list):
            cap[k] = torch.stack(cap[k], 1)
    outputs["features"] = features
    outputs["captured"] = cap
    return outputs


CASES = {
    # name: (V, H, W, feature seed, weight seed)
    "hotpath_v3_96x128": dict(V=3, H=96, W=128, fseed=1234, wseed=7, numdepth=192),
    "hotpath_v4_64x96": dict(V=4, H=64, W=96, fseed=99, wseed=11, numdepth=48),
}


def make_case(c, args):
    feats = synth.make_features(c["V"], c["H"], c["W"], seed=c["fseed"])
    proj = synth.make_proj_matrices(c["V"], c["H"], c["W"], theta_step=0.12)
    dv = synth.make_depth_values(c["numdepth"], 425.0, 2.65 * 192 / c["numdepth"])
    return feats, proj, dv


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    R = import_reference()
    args = default_args()
    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    tmp = [5.0, 5.0, 5.0, 1.0]

    for name, c in CASES.items():
        torch.manual_seed(0)
        model = RefHotPath(R, args).eval()
        synth.randomize_state_dict(model, seed=c["wseed"])
        feats, proj, dv = make_case(c, args)
        out = reference_hotpath(R, model, args, feats, proj, dv, tmp)
        blob = {}
        for s in range(4):
            so = out[f"stage{s + 1}"]
            for k in ("depth", "photometric_confidence", "depth_values", "prob_volume_pre"):  # prob_volume = softmax(pre)
                blob[f"stage{s + 1}.{k}"] = so[k][0].numpy()
            blob[f"stage{s + 1}.entropy"] = out["captured"][f"stage{s + 1}.entropy"][0].numpy()
            blob[f"stage{s + 1}.vis_weight"] = out["captured"][f"stage{s + 1}.vis_weight"][0].numpy()
            if s < 2:
                blob[f"stage{s + 1}.volume_mean"] = out["captured"][f"stage{s + 1}.volume_mean"][0].numpy()
        blob["refined_depth"] = out["refined_depth"][0].numpy()
        blob["photometric_confidence"] = out["photometric_confidence"][0].numpy()
        blob["fmt.stage1"] = out["features"]["stage1"][0].numpy()
        # full-resolution FMT outputs are large: keep view 1 (first source view) of stages 2-4 only
        for k in ("stage2", "stage3", "stage4"):
            blob[f"fmt.{k}.view1"] = out["features"][k][0, 1].numpy()
        blob["meta"] = np.frombuffer(json.dumps(c).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **blob)
        print(name, "refined_depth range", float(out["refined_depth"].min()), float(out["refined_depth"].max()),
              "stage4 conf mean", float(out["stage4"]["photometric_confidence"].mean()))

    # ---- warp seam alone, including out-of-image taps and per-pixel hypotheses (warping.py:69-109)
    g = torch.Generator().manual_seed(5)
    H, W, C, D = 12, 16, 8, 4
    src = torch.randn(1, C, H, W, generator=g)
    pm = synth.make_proj_matrices(3, H * 8, W * 8, theta_step=0.5)["stage1"]  # wide baseline -> taps leave the image
    from oracle.hotpath import compose_projection
    refp, srcp = compose_projection(pm[:, 0]), compose_projection(pm[:, 2])
    dvals = 425.0 + 500.0 * torch.rand(1, D, H, W, generator=g)
    warped, mask = R["homo_warp"](src, srcp, refp, dvals)
    np.savez_compressed(os.path.join(out_dir, "warp_seam.npz"), src=src.numpy(), src_proj=srcp.numpy(),
                        ref_proj=refp.numpy(), depth_values=dvals.numpy(), warped=warped.numpy(), mask=mask.numpy())
    print("warp_seam: fraction masked", float(mask.float().mean()))


if __name__ == "__main__":
    main()
