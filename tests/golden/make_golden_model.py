"""Generates tests/golden/dn_model_glue_*.npz by executing the REFERENCE's own
`DNSplatterModel.get_outputs` and `DNSplatterModel.get_loss_dict`
(/root/reference/dn_splatter/dn_model.py:404-612 and :614-729, unmodified, imported from where they lie)
with the three things that are absent from this container replaced:

  * gsplat 1.0.0's `rasterization` / `rasterize_gaussians` / `quat_to_rotmat`  -> oracle/gsplat_ref.py (the restatement)
  * nerfstudio 1.1.3's `SplatfactoModel` base class (`get_gt_img`, the parent `get_loss_dict`, the parameter
    properties) and `get_viewmat`                                              -> restated below per SURVEY.md A7
  * every other third-party import (torchmetrics, torchvision, open3d, ...)    -> inert auto-stubs

So everything BETWEEN the gsplat calls — background blend and clamp, the detached-max depth fill, the per-Gaussian
normal construction (argmin / one-hot / flip / c2w), which tensors are detached, the white-background normal pass,
normalise + remap, surface normals from the detached depth, gt clamping, mask handling, normal_supervision="depth",
how DNRegularization is wired and weighted — is the reference's code, not a restatement.  The committed .npz files pin
oracle/dn_ref.get_outputs (tests/test_oracle_glue_golden.py), the host-side model on the CPU proxy, and the CUDA path
(tests/test_gpu_model.py::test_model_matches_reference_glue_goldens).

Run only where /root/reference exists:   python tests/golden/make_golden_model.py
"""
import dataclasses
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))

from oracle import dn_ref, gsplat_ref as G  # noqa: E402


# ---------------------------------------------------------------- inert stubs for absent third-party packages
class _Any(type):
    def __getattr__(cls, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _mk(n)


def _mk(name):
    return _Any(name, (torch.nn.Module,), {"__init__": lambda self, *a, **k: torch.nn.Module.__init__(self)})


class _StubModule(types.ModuleType):
    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        v = _mk(n)
        setattr(self, n, v)
        return v


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    roots = ("nerfstudio", "gsplat", "torchvision", "torchmetrics", "open3d", "cv2", "tyro", "viser", "rich", "pymeshlab")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.roots and name not in sys.modules:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, m):
        pass


# ---------------------------------------------------------------- nerfstudio 1.1.3 pieces, restated [EXT]
@dataclasses.dataclass
class SplatfactoModelConfig:
    sh_degree: int = 3
    sh_degree_interval: int = 1000
    rasterize_mode: str = "classic"
    ssim_lambda: float = 0.2
    use_scale_regularization: bool = False
    max_gauss_ratio: float = 10.0
    num_downscales: int = 2
    resolution_schedule: int = 3000
    refine_every: int = 100
    reset_alpha_every: int = 30
    warmup_length: int = 500
    background_color: str = "random"


class SplatfactoModel(torch.nn.Module):
    means = property(lambda self: self.gauss_params["means"])
    scales = property(lambda self: self.gauss_params["scales"])
    quats = property(lambda self: self.gauss_params["quats"])
    features_dc = property(lambda self: self.gauss_params["features_dc"])
    features_rest = property(lambda self: self.gauss_params["features_rest"])
    opacities = property(lambda self: self.gauss_params["opacities"])
    device = property(lambda self: torch.device("cpu"))

    def _get_downscale_factor(self):
        if self.training:
            return 2 ** max(self.config.num_downscales - self.step // self.config.resolution_schedule, 0)
        return 1

    def _downscale_if_required(self, image):
        assert self._get_downscale_factor() == 1
        return image

    def get_gt_img(self, image):
        if image.dtype == torch.uint8:
            image = image.float() / 255.0
        return self._downscale_if_required(image).to(self.device)

    def composite_with_background(self, image, background):
        assert image.shape[2] == 3
        return image

    def _get_background_color(self):
        return self._fixed_background

    def get_loss_dict(self, outputs, batch, metrics_dict=None):
        gt_img = self.composite_with_background(self.get_gt_img(batch["image"]), outputs["background"])
        pred_img = outputs["rgb"]
        if "mask" in batch:
            mask = self._downscale_if_required(batch["mask"]).to(self.device)
            assert mask.shape[:2] == gt_img.shape[:2] == pred_img.shape[:2]
            gt_img = gt_img * mask
            pred_img = pred_img * mask
        Ll1 = torch.abs(gt_img - pred_img).mean()
        assert self.config.ssim_lambda == 0.0, "goldens keep SSIM (torchmetrics, absent) out of the picture"
        if self.config.use_scale_regularization and self.step % 10 == 0:
            scale_exp = torch.exp(self.scales)
            scale_reg = torch.maximum(scale_exp.amax(dim=-1) / scale_exp.amin(dim=-1),
                                      torch.tensor(self.config.max_gauss_ratio)) - self.config.max_gauss_ratio
            scale_reg = 0.1 * scale_reg.mean()
        else:
            scale_reg = torch.tensor(0.0)
        return {"main_loss": (1 - self.config.ssim_lambda) * Ll1, "scale_reg": scale_reg}


def _get_viewmat(optimized_camera_to_world):
    assert optimized_camera_to_world.shape == (1, 3, 4)
    return dn_ref.get_viewmat(optimized_camera_to_world[0])[None]


# ---------------------------------------------------------------- gsplat 1.0.0 call signatures over the restatement
_BIN = {}


def _rasterization(means, quats, scales, opacities, colors, viewmats, Ks, width, height, tile_size, packed, near_plane,
                   far_plane, render_mode, sh_degree, sparse_grad, absgrad, rasterize_mode):
    assert viewmats.shape == (1, 4, 4) and Ks.shape == (1, 3, 3) and render_mode == "RGB+ED"
    assert not packed and not sparse_grad and absgrad
    render, alpha, info = G.rasterization(means, quats, scales, opacities, colors, viewmats[0], Ks[0], width, height,
                                          tile_size, near_plane=near_plane, far_plane=far_plane, sh_degree=sh_degree,
                                          rasterize_mode=rasterize_mode)
    out = {k: info[k][None] for k in ("means2d", "radii", "depths", "conics", "tiles_per_gauss")}
    return render[None], alpha[None], out


def _rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height, img_width, block_width,
                         background=None, return_alpha=False):
    assert background is None and not return_alpha and opacity.dim() == 2
    _, _, flatten_ids, offsets, _ = G.isect_tiles(xys, radii, depths, block_width, img_width, img_height)
    return G.rasterize_gaussians_legacy(xys, conics, colors, opacity.squeeze(-1), img_height, img_width, block_width,
                                        offsets, flatten_ids)


def install():
    sys.meta_path.insert(0, _Finder())
    sf = types.ModuleType("nerfstudio.models.splatfacto")
    sf.SplatfactoModel, sf.SplatfactoModelConfig, sf.get_viewmat = SplatfactoModel, SplatfactoModelConfig, _get_viewmat
    sf.RGB2SH = lambda rgb: (rgb - 0.5) / 0.28209479177387814
    sf.SH2RGB = lambda sh: sh * 0.28209479177387814 + 0.5
    sf.random_quat_tensor = None
    sf.num_sh_bases = G.num_sh_bases
    import nerfstudio.models  # noqa: F401  (auto-stub package)

    sys.modules["nerfstudio.models.splatfacto"] = sf
    pkg = types.ModuleType("dn_splatter")
    pkg.__path__ = [os.path.join(REF, "dn_splatter")]
    sys.modules["dn_splatter"] = pkg
    import dn_splatter.dn_model as M

    M.rasterization, M.rasterize_gaussians, M.quat_to_rotmat = _rasterization, _rasterize_gaussians, G.quat_to_rotmat
    torch.Tensor.cuda = lambda self, *a, **k: self  # dn_model.py:474 calls .cuda() on the intrinsics
    return M


def make_camera(M, cam):
    class Camera(M.Cameras):
        def get_intrinsics_matrices(self):
            return dn_ref.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])[None]

        def rescale_output_resolution(self, s):
            assert s == 1

    c = Camera()
    c.camera_to_worlds = cam["c2w"][None].clone()
    c.shape = (1,)
    for k in ("fx", "fy", "cx", "cy"):
        setattr(c, k, torch.tensor([[float(cam[k])]]))
    c.width, c.height = torch.tensor([[cam["width"]]]), torch.tensor([[cam["height"]]])
    c.metadata = {"cam_idx": 3}
    return c


def make_model(M, params, background, **cfg_kw):
    from dn_splatter.losses import DepthLoss

    cfg = M.DNSplatterModelConfig(**cfg_kw)
    m = M.DNSplatterModel.__new__(M.DNSplatterModel)
    torch.nn.Module.__init__(m)
    m.config = cfg
    m.step = 30000
    m.crop_box = None
    m.gauss_params = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    m._fixed_background = background
    m.camera_optimizer = types.SimpleNamespace(apply_to_camera=lambda cam: cam.camera_to_worlds)
    # the wiring of populate_modules (dn_model.py:224-265), by hand: populate_modules itself needs seed points,
    # a CameraOptimizer and the parent's populate_modules
    m.depth_loss = DepthLoss(cfg.depth_loss_type)
    m.regularization_strategy = (M.DNRegularization() if cfg.regularization_strategy == "dn-splatter"
                                 else M.AGSMeshRegularization())
    if cfg.use_depth_loss:
        m.regularization_strategy.depth_loss_type = cfg.depth_loss_type
        m.regularization_strategy.depth_loss = m.depth_loss
        m.regularization_strategy.depth_lambda = cfg.depth_lambda
    else:
        m.regularization_strategy.depth_loss_type = None
        m.regularization_strategy.depth_loss = None
    if not cfg.use_normal_loss:
        m.regularization_strategy.normal_loss = None
    m.train()
    return m


def main():
    from dn_splatter_b200.synthetic import BACKGROUND, make_scene, ring_cameras  # deterministic scene generator only

    M = install()
    from dn_splatter.losses import DepthLossType

    cases = {
        "a": dict(n=400, W=64, H=48, view=1, mask=False, depth_key="mono_depth",
                  cfg=dict(use_depth_loss=True, depth_lambda=0.2, depth_loss_type=DepthLossType.EdgeAwareLogL1,
                           normal_supervision="mono")),
        "b": dict(n=300, W=56, H=40, view=3, mask=True, depth_key="sensor_depth",
                  cfg=dict(use_depth_loss=True, depth_lambda=0.35, depth_loss_type=DepthLossType.LogL1,
                           normal_supervision="depth", use_scale_regularization=True)),
        "c": dict(n=300, W=48, H=48, view=0, mask=False, depth_key=None,
                  cfg=dict(use_depth_loss=False, predict_normals=True, normal_supervision="mono")),
        "d": dict(n=300, W=48, H=40, view=2, mask=False, depth_key="sensor_depth", confidence=True,
                  cfg=dict(use_depth_loss=True, depth_lambda=0.2, depth_loss_type=DepthLossType.EdgeAwareLogL1,
                           normal_supervision="mono", regularization_strategy="ags-mesh")),
        "e": dict(n=300, W=48, H=40, view=4, mask=False, depth_key="mono_depth",
                  cfg=dict(use_depth_loss=True, depth_lambda=0.2, depth_loss_type=DepthLossType.EdgeAwareLogL1,
                           normal_supervision="mono", rasterize_mode="antialiased")),
        "f": dict(n=300, W=48, H=40, view=1, mask=False, depth_key="mono_depth", step=1500,  # SH degree 1 of 3
                  cfg=dict(use_depth_loss=True, depth_lambda=0.2, depth_loss_type=DepthLossType.L1,
                           normal_supervision="mono")),
        "g": dict(n=300, W=48, H=40, view=2, mask=False, depth_key=None, eval=True,  # eval mode, forward only
                  cfg=dict(use_depth_loss=False, normal_supervision="mono")),
    }
    for tag, c in cases.items():
        params = make_scene(c["n"], seed=7 + ord(tag))
        cam = ring_cameras(5, c["W"], c["H"])[c["view"]]
        H, W = c["H"], c["W"]
        g = torch.Generator().manual_seed(100 + ord(tag))
        batch = {"image": (torch.rand(H, W, 3, generator=g) * 255).to(torch.uint8), "normal": torch.rand(H, W, 3, generator=g)}
        if c["depth_key"]:
            d = 2 + 6 * torch.rand(H, W, 1, generator=g)
            d[torch.rand(H, W, 1, generator=g) < 0.1] = 0.0
            batch[c["depth_key"]] = d
        if c.get("confidence"):
            batch["confidence"] = (torch.rand(H, W, 1, generator=g) * 255).to(torch.uint8)
        if c["mask"]:
            batch["mask"] = (torch.rand(H, W, 1, generator=g) > 0.2).float()
        bg = torch.tensor(BACKGROUND)
        m = make_model(M, params, bg, ssim_lambda=0.0, num_downscales=0, max_gauss_ratio=5.0, **c["cfg"])
        m.step = c.get("step", 30000)
        if c.get("eval"):
            m.eval()
        camera = make_camera(M, cam)
        outputs = M.DNSplatterModel.get_outputs(m, camera)
        saved = {k: outputs[k].detach().clone() for k in ("rgb", "depth", "normal", "surface_normal", "accumulation")}
        if c.get("eval"):
            m.train()  # the loss below is only there to keep the file layout uniform
        loss_dict = M.DNSplatterModel.get_loss_dict(m, outputs, {k: v.clone() for k, v in batch.items()})
        total = loss_dict["main_loss"] + loss_dict["scale_reg"]
        total.backward()
        z = {"in_" + k: v.detach().numpy() for k, v in params.items()}
        z.update({"batch_" + k: v.numpy() for k, v in batch.items()})
        z["cam_c2w"] = cam["c2w"].numpy()
        z["cam_intr"] = np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["width"], cam["height"]], dtype=np.float64)
        z["background"] = bg.numpy()
        z.update({"out_" + k: v.numpy() for k, v in saved.items()})
        z["out_normal_after_loss"] = outputs["normal"].detach().numpy()  # quirk B11: masked in place in the dict
        z["out_gauss_normals"] = m.gauss_params["normals"].detach().numpy()
        z["out_radii"] = m.radii.numpy()
        z["out_main_loss"] = loss_dict["main_loss"].detach().numpy()
        z["out_scale_reg"] = loss_dict["scale_reg"].detach().numpy()
        for k in ("means", "quats", "scales", "opacities", "features_dc", "features_rest"):
            z["grad_" + k] = m.gauss_params[k].grad.numpy()
        cfgd = {k: (v.value if hasattr(v, "value") else v) for k, v in c["cfg"].items()}
        z["cfg_json"] = np.array(__import__("json").dumps(cfgd))
        z["step"], z["eval"] = np.array(c.get("step", 30000)), np.array(bool(c.get("eval", False)))
        np.savez_compressed(os.path.join(OUT, f"dn_model_glue_{tag}.npz"), **z)
        print(tag, "main_loss", float(loss_dict["main_loss"]), "scale_reg", float(loss_dict["scale_reg"]),
              "visible", int((m.radii > 0).sum()), "/", c["n"])


if __name__ == "__main__":
    main()
