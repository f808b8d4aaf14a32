"""CPU-side tests: the C-ABI library builds/loads and exports every symbol include/mvsf_b200.h declares (no compute
calls without a GPU), host logic (config schema, state-dict key compatibility, weight packing) and loud failure
without a CUDA device."""
import os
import re

import pytest
import torch

from mvsformerplusplus_b200 import packing, synth
from mvsformerplusplus_b200.config import default_args, load_args, validate_args
from mvsformerplusplus_b200.params import build_hotpath_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mvsformerplusplus_b200.build import build
    build()
    from mvsformerplusplus_b200 import _lib
    return _lib.lib()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "mvsf_b200.h")).read()
    names = sorted(set(re.findall(r"\b(mvsf_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mvsf_b200.h but not exported"
    assert lib.mvsf_abi_version() == 1


def test_bad_arguments_are_rejected_without_touching_the_gpu(lib):
    import ctypes
    from mvsformerplusplus_b200 import _lib
    need = ctypes.c_size_t(0)
    rc = lib.mvsf_costreg_tr_workspace_bytes(8, 31, 16, 16, ctypes.byref(need))  # D not a multiple of 2
    assert rc == -1 and b"down_rate" in lib.mvsf_last_error()
    assert lib.mvsf_costreg_unet_workspace_bytes(0, 8, 12, 16, 16, ctypes.byref(need)) == -1  # CostRegNet needs D % 8 == 0
    assert lib.mvsf_costreg_unet_workspace_bytes(1, 8, 4, 1152, 1536, ctypes.byref(need)) == 0
    # fp16 hi|lo buffers, 4 bytes per element: the split input (n0) + two buffers per level
    n0, n1, n2, n3 = 4 * 1152 * 1536 * 8, 4 * 576 * 768 * 16, 4 * 288 * 384 * 32, 4 * 144 * 192 * 64
    assert need.value == 4 * (n0 + 2 * (n1 + n2 + n3))
    assert lib.mvsf_costreg_unet_tc_bytes(ctypes.byref(need)) == 0 and need.value % 16 == 0 and need.value > 0
    with pytest.raises(RuntimeError, match="status -1"):
        _lib.check(lib.mvsf_warp_corr_entropy(None, None, None, None, 5, 8, 8, 4, 16, 16, None), "warp_corr_entropy")


def test_state_dict_keys_match_reference_inventory():
    m = build_hotpath_params(default_args())
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    ref = {}
    for line in open(os.path.join(ROOT, "tests", "golden", "hotpath_state_dict_keys.txt")):
        k, s = line.strip().split(" ", 1)
        ref[k] = eval(s)
    assert mine == ref


def test_hotpath_modules_have_reference_keys_and_fail_loudly_on_cpu():
    from mvsformerplusplus_b200.hotpath import HotPathNet
    net = HotPathNet(default_args()).eval()
    ref = [l.split(" ", 1)[0] for l in open(os.path.join(ROOT, "tests", "golden", "hotpath_state_dict_keys.txt"))]
    assert sorted(net.state_dict().keys()) == sorted(ref)
    feats = synth.make_features(3, 64, 96)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net.forward_features(feats, synth.make_proj_matrices(3, 64, 96), synth.make_depth_values(48), [5., 5., 5., 1.])


def test_config_schema_and_reference_errors():
    a = load_args({"arch": {"args": {"ndepths": [16, 8, 8, 4]}}})
    assert a["ndepths"] == [16, 8, 8, 4] and a["base_ch"] == [8, 8, 8, 8]
    validate_args(a)
    bad = default_args(); bad["fusion_type"] = "mean"
    with pytest.raises(NotImplementedError, match="Not implemented fusion type"):
        validate_args(bad)
    bad = default_args(); bad["FMT_config"]["attention_type"] = "FLASH2"
    with pytest.raises(NotImplementedError):
        validate_args(bad)


def test_packing_layout_sizes_and_bn_folding():
    m = build_hotpath_params(default_args()).eval()
    sd = synth.randomize_state_dict(m, seed=4)
    assert packing.pack_vis(sd, "fusions.0.vis.").numel() == 3652
    k0, f0 = packing.pack_costreg_unet(sd, "fusions.1.cost_reg.")
    k1, f1 = packing.pack_costreg_unet(sd, "fusions.3.cost_reg.")
    assert (k0, k1) == (0, 1) and f0.numel() == 290800 and f1.numel() == 290596
    assert packing.pack_costreg_tr(sd, "fusions.0.cost_reg.", 6).numel() == 332960
    assert packing.pack_fmt(sd).numel() == 214464
    # folded conv+bias reproduces conv -> BatchNorm(eval)
    x = torch.randn(1, 1, 9, 9)
    w = sd["fusions.0.vis.0.conv.weight"]
    bn = torch.nn.BatchNorm2d(16).eval()
    bn.load_state_dict({k.split("bn.")[1]: v for k, v in sd.items() if k.startswith("fusions.0.vis.0.bn.")})
    want = bn(torch.nn.functional.conv2d(x, w, padding=1))
    flat = packing.pack_vis(sd, "fusions.0.vis.")
    wf = flat[:144].view(9, 16).t().reshape(16, 1, 3, 3)
    got = torch.nn.functional.conv2d(x, wf, flat[144:160], padding=1)
    assert float((got - want).abs().max()) < 1e-5
