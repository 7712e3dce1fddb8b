"""TEST INFRASTRUCTURE: lets the host-side training logic (DNSplatterModel, Trainer, densification) run on the CPU by
substituting the oracle for the CUDA entry points.  Only tests import this; the product has no CPU path."""
from contextlib import contextmanager

import torch

from oracle import dn_ref
from oracle import gsplat_ref as G


class _Out:
    pass


def oracle_rasterize(means, quats, scales, opacities, sh_dc, sh_rest, viewmat, K, width, height, *, sh_degree=3,
                     near_plane=0.01, far_plane=1e10, eps2d=0.3, antialiased=False, background=(0, 0, 0), render_normals=True,
                     c2w=None, activated=False, surface_normal=True, grad_sink=None, exact_lists=False, sync_free=False,
                     fixed_capacity=0, **_kernel_options):
    if isinstance(background, torch.Tensor):
        background = background.tolist()
    bg = torch.tensor(list(background), dtype=torch.float32)
    qn = quats / quats.norm(dim=-1, keepdim=True)
    colors = torch.cat([sh_dc[:, None, :], sh_rest], dim=1)
    render, alpha, info = G.rasterization(means, qn, torch.exp(scales), torch.sigmoid(opacities).squeeze(-1), colors,
                                          viewmat.float(), K.float(), width, height, 16, near_plane, far_plane, sh_degree,
                                          "antialiased" if antialiased else "classic")
    o = _Out()
    o.rgb = torch.clamp(render[..., :3] + (1 - alpha) * bg, 0.0, 1.0)
    d = render[..., 3:4]
    o.depth = torch.where(alpha > 0, d, d.detach().max())
    o.alpha = alpha
    m2d = info["means2d"]
    if m2d.requires_grad:
        m2d.retain_grad()
        m2d.register_hook(lambda g: setattr(m2d, "absgrad", g.abs()))  # proxy for the kernel's absgrad
    o.means2d, o.radii, o.depths, o.conics = m2d, info["radii"], info["depths"], info["conics"]
    o.tiles_per_gauss, o.info = info["tiles_per_gauss"], info
    if render_normals:
        nw, ncam = dn_ref.gaussian_normals(quats, scales, means, c2w.float())
        nim = G.rasterize_gaussians_legacy(m2d.detach(), info["conics"], ncam, torch.sigmoid(opacities).squeeze(-1), height,
                                           width, 16, info["isect_offsets"], info["flatten_ids"])
        nim = nim / nim.norm(dim=-1, keepdim=True)
        o.normal, o.normals_world = (nim + 1) / 2, nw.detach()
    else:
        o.normal, o.normals_world = torch.zeros(height, width, 3), torch.zeros(means.shape[0], 3)
    fx, fy, cx, cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    o.surface_normal = dn_ref.surface_normal_output(o.depth, fx, fy, cx, cy, width, height)
    return o


@contextmanager
def cpu_proxy():
    import dn_splatter_b200.dn_model as M
    import dn_splatter_b200.regularization_strategy as RS

    saved = (M.dn_rasterize, M.FusedL1, M.u8_to_float, RS._FusedDNLoss, RS._ScaleLoss, M.normal_from_depth_image)

    def nfd(depths, fx, fy, cx, cy, img_size, c2w, device, smooth=False):
        return dn_ref.normal_from_depth_image(depths, fx, fy, cx, cy, int(img_size[0]), int(img_size[1]))

    class L1Proxy:
        @staticmethod
        def apply(pred, gt, holder=None):
            g = gt.float() / 255.0 if gt.dtype == torch.uint8 else gt
            return (pred - g).abs().mean()

    class ScaleProxy:
        @staticmethod
        def apply(scales):
            return torch.exp(scales).min(dim=1, keepdim=True)[0].mean()

    class DNProxy:
        @staticmethod
        def apply(pd, pn, gd, gn, gi, dtype, lam, tol, use_normal, holder=None, edge_image=None):
            types = {0: None, 1: "EdgeAwareLogL1", 2: "LogL1", 3: "L1", 4: "MSE"}
            if edge_image is not None:
                gi = (edge_image.float() / 255.0).clamp(min=10 / 255.0)
            if gn is not None and gn.dtype == torch.uint8:
                gn = gn.float() / 255.0
            full = dn_ref.dn_regularization(pd if pd is not None else torch.zeros(1, 1, 1), gd, pn, gn, torch.zeros(1, 3),
                                            gi, depth_lambda=lam, depth_tolerance=tol, depth_loss_type=types[dtype],
                                            use_normal_loss=bool(use_normal))
            return full - 1.0  # dn_regularization adds mean(min(exp(0))) = 1 for the dummy scales

    M.dn_rasterize, M.FusedL1 = oracle_rasterize, L1Proxy
    M.u8_to_float = lambda img, divisor=255.0, clamp_min=0.0: (img.float() / divisor).clamp(min=clamp_min)
    RS._FusedDNLoss, RS._ScaleLoss = DNProxy, ScaleProxy
    M.normal_from_depth_image = nfd
    try:
        yield
    finally:
        M.dn_rasterize, M.FusedL1, M.u8_to_float, RS._FusedDNLoss, RS._ScaleLoss, M.normal_from_depth_image = saved
