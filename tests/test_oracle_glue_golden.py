"""The glue goldens (tests/golden/dn_model_glue_*.npz) were produced by the REFERENCE's own get_outputs / get_loss_dict
code executed with gsplat served by oracle/gsplat_ref.py (see tests/golden/make_golden_model.py).  Here:
  * oracle/dn_ref.get_outputs (the restatement of that glue) must reproduce the reference's outputs;
  * the product's host-side model (dn_splatter_b200.DNSplatterModel), run on the CPU proxy, must reproduce the
    reference's loss dict and parameter gradients — i.e. the mirrored control flow is the reference's control flow.
The CUDA kernels are compared with the same files in tests/test_gpu_model.py."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import dn_ref
from tests.cpu_proxy import cpu_proxy

FILES = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "dn_model_glue_*.npz")))
PARAMS = ("means", "quats", "scales", "opacities", "features_dc", "features_rest")


def load(f):
    z = np.load(f)
    d = {k: (torch.from_numpy(z[k]) if z[k].shape != () or z[k].dtype.kind != "U" else str(z[k])) for k in z.files}
    d["cfg"] = json.loads(str(z["cfg_json"]))
    return d


def model_from_golden(z, device="cpu"):
    from dn_splatter_b200.cameras import Cameras
    from dn_splatter_b200.dn_model import DNSplatterModelConfig
    from dn_splatter_b200.losses import DepthLossType

    cfg = dict(z["cfg"])
    if "depth_loss_type" in cfg:
        cfg["depth_loss_type"] = DepthLossType(cfg["depth_loss_type"])
    m = DNSplatterModelConfig(random_init=True, num_random=16, background_color="black", ssim_lambda=0.0, **cfg).setup(device=device)
    m.load_gaussians({k: z["in_" + k] for k in PARAMS})
    m.background_color = z["background"].clone()
    m.step = int(z["step"])
    m.train()
    fx, fy, cx, cy, W, H = [float(v) for v in z["cam_intr"]]
    cam = Cameras(z["cam_c2w"][None].to(device), fx, fy, cx, cy, int(W), int(H), metadata={"cam_idx": 3})
    batch = {k[len("batch_"):]: v.clone().to(device) for k, v in z.items() if k.startswith("batch_")}
    return m, cam, batch


def test_goldens_exist():
    assert len(FILES) == 7


@pytest.mark.parametrize("f", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_get_outputs_reproduces_reference_glue(f):
    z = load(f)
    fx, fy, cx, cy, W, H = [float(v) for v in z["cam_intr"]]
    p = {k: z["in_" + k] for k in PARAMS}
    out = dn_ref.get_outputs(p, z["cam_c2w"], fx, fy, cx, cy, int(W), int(H), z["background"],
                             sh_degree=min(int(z["step"]) // 1000, 3), rasterize_mode=z["cfg"].get("rasterize_mode", "classic"))
    # same primitives, same operation order: differences are re-association noise of a few ulp at most
    for k in ("rgb", "depth", "normal", "surface_normal", "accumulation"):
        torch.testing.assert_close(out[k], z["out_" + k], rtol=1e-5, atol=1e-6, msg=lambda m: f"{k}: {m}")
    torch.testing.assert_close(out["gauss_normals"], z["out_gauss_normals"], rtol=1e-6, atol=1e-7)
    assert torch.equal(out["info"]["radii"], z["out_radii"])


@pytest.mark.parametrize("f", FILES, ids=[os.path.basename(f) for f in FILES])
def test_host_model_reproduces_reference_loss_dict_and_gradients(f):
    z = load(f)
    with cpu_proxy():
        m, cam, batch = model_from_golden(z)
        if bool(z["eval"]):
            m.eval()
        out = m.get_outputs(cam)
        m.train()
        for k in ("rgb", "depth", "normal", "surface_normal", "accumulation"):
            torch.testing.assert_close(out[k], z["out_" + k], rtol=1e-5, atol=1e-6)
        ld = m.get_loss_dict(out, batch)
        torch.testing.assert_close(ld["main_loss"], z["out_main_loss"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(ld["scale_reg"].reshape(()), z["out_scale_reg"].reshape(()), rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(out["normal"], z["out_normal_after_loss"], rtol=1e-5, atol=1e-6)  # quirk B11
        (ld["main_loss"] + ld["scale_reg"]).backward()
        for k in PARAMS:
            got, want = m.gauss_params[k].grad, z["grad_" + k]
            rel = float((got - want).norm() / (want.norm() + 1e-30))
            assert rel < 1e-4, (k, rel)
        assert m.camera_idx == 3
        torch.testing.assert_close(m.normals.detach(), z["out_gauss_normals"], rtol=1e-6, atol=1e-7)
