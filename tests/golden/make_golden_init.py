"""Generates tests/golden/dn_init_*.npz by executing the REFERENCE's own `DNSplatterModel.populate_modules`
(/root/reference/dn_splatter/dn_model.py:131-265) and its helpers `rotate_vector_to_vector` / `matrix_to_quaternion`
(:1520-1600), unmodified.  Restated [EXT nerfstudio 1.1.3]: `k_nearest_sklearn` (sklearn NearestNeighbors, k+1, drop
self), `random_quat_tensor`, `RGB2SH`, `num_sh_bases`.  tests/test_init_golden.py checks
dn_splatter_b200.DNSplatterModel.populate_modules and the quaternion helpers against the files.

Run only where /root/reference exists:   python tests/golden/make_golden_init.py
"""
import math
import os
import sys
import types

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
import make_golden_model as G0  # noqa: E402
from oracle import gsplat_ref as G  # noqa: E402

NAMES = ("means", "scales", "quats", "features_dc", "features_rest", "opacities", "normals")


def random_quat_tensor(N):
    u, v, w = torch.rand(N), torch.rand(N), torch.rand(N)
    return torch.stack([torch.sqrt(1 - u) * torch.sin(2 * math.pi * v), torch.sqrt(1 - u) * torch.cos(2 * math.pi * v),
                        torch.sqrt(u) * torch.sin(2 * math.pi * w), torch.sqrt(u) * torch.cos(2 * math.pi * w)], dim=-1)


class Parent(G0.SplatfactoModel):
    def k_nearest_sklearn(self, x, k):
        from sklearn.neighbors import NearestNeighbors

        x_np = x.cpu().numpy()
        nn_model = NearestNeighbors(n_neighbors=k + 1, algorithm="auto", metric="euclidean").fit(x_np)
        distances, indices = nn_model.kneighbors(x_np)
        return distances[:, 1:].astype(np.float32), indices[:, 1:].astype(np.float32)


def main():
    G0.install()
    sf = sys.modules["nerfstudio.models.splatfacto"]
    sf.SplatfactoModel, sf.random_quat_tensor = Parent, random_quat_tensor
    sys.modules.pop("dn_splatter.dn_model", None)
    import dn_splatter.dn_model as M

    M.quat_to_rotmat, M.num_sh_bases = G.quat_to_rotmat, G.num_sh_bases
    g = torch.Generator().manual_seed(77)
    n = 150
    pts = torch.randn(n, 3, generator=g) * 2
    rgb = (torch.rand(n, 3, generator=g) * 255).floor()
    nrm = torch.randn(n, 3, generator=g) * 3
    nrm[0] = torch.tensor([0.0, 0.0, 1.0])   # parallel to the z axis the helper rotates from
    nrm[1] = torch.tensor([0.0, 0.0, -2.0])  # anti-parallel
    for tag, seeds in (("normals", (pts, rgb, nrm)), ("plain", (pts, rgb))):
        m = M.DNSplatterModel.__new__(M.DNSplatterModel)
        torch.nn.Module.__init__(m)
        m.config = M.DNSplatterModelConfig(use_depth_loss=True, depth_lambda=0.2)
        m.config.random_init = False
        m.config.camera_optimizer = types.SimpleNamespace(setup=lambda **k: None)
        m.seed_points, m.num_train_data = seeds, 10
        torch.manual_seed(4321)  # random_quat_tensor in the seed-points-without-normals branch
        M.DNSplatterModel.populate_modules(m)
        z = {"seed_" + k: v.numpy() for k, v in zip(("points", "rgb", "normals"), seeds)}
        z.update({"out_" + k: m.gauss_params[k].detach().numpy() for k in NAMES})
        z["background_color"] = m.background_color.numpy()
        z["depth_lambda"] = np.array(m.regularization_strategy.depth_lambda)
        np.savez_compressed(os.path.join(OUT, f"dn_init_{tag}.npz"), **z)
        print(tag, {k: tuple(m.gauss_params[k].shape) for k in NAMES})
    # the two helpers on their own, incl. degenerate pairs
    v1 = torch.nn.functional.normalize(torch.randn(64, 3, generator=g), dim=-1)
    v2 = torch.nn.functional.normalize(torch.randn(64, 3, generator=g), dim=-1)
    v2[0], v2[1] = v1[0], -v1[1]
    mat = M.rotate_vector_to_vector(v1, v2)
    q = M.matrix_to_quaternion(mat)
    np.savez_compressed(os.path.join(OUT, "dn_init_helpers.npz"), v1=v1.numpy(), v2=v2.numpy(), mat=mat.numpy(), quat=q.numpy())
    print("helpers", tuple(mat.shape), tuple(q.shape), "finite:", bool(torch.isfinite(mat).all()), bool(torch.isfinite(q).all()))


if __name__ == "__main__":
    main()
