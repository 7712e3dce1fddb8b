"""Generates tests/golden/dn_reference_losses2.npz from the REFERENCE's own loss classes that round 1 left as shells:
DSSIML1 (per-pixel), SensorDepthLoss, AdaptiveDepth, AdaptiveNormal, LocalPearsonDepthLoss
(/root/reference/dn_splatter/losses.py:73-152, 297-352, 386-485), imported unmodified through make_golden.py's stub
shim.  The reference hard-codes device="cuda" in two of them; this container has no GPU, so while those forward()s run
`torch.randint` / `torch.tensor` / `Tensor.cuda` are redirected to the CPU (the arithmetic is untouched).  The random
window corners LocalPearsonDepthLoss draws are recorded and replayed by the test.

    python tests/golden/make_golden_losses.py
"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, install_shim  # noqa: E402


class cpu_redirect:
    """Drops device="cuda" from torch.randint / torch.tensor and makes Tensor.cuda() the identity."""

    def __enter__(self):
        self.saved = (torch.randint, torch.tensor, torch.Tensor.cuda)
        self.draws = []
        ri, tt = torch.randint, torch.tensor

        def randint(*a, **k):
            k.pop("device", None)
            out = ri(*a, **k)
            self.draws.append(out.clone())
            return out

        def tensor(*a, **k):
            k.pop("device", None)
            return tt(*a, **k)

        torch.randint, torch.tensor = randint, tensor
        torch.Tensor.cuda = lambda self_, *a, **k: self_
        return self

    def __exit__(self, *exc):
        torch.randint, torch.tensor, torch.Tensor.cuda = self.saved


def main():
    install_shim()
    import nerfstudio.field_components.field_heads as fh

    fh.FieldHeadNames.SDF = "sdf"
    from dn_splatter.losses import AdaptiveDepth, AdaptiveNormal, DSSIML1, LocalPearsonDepthLoss, SensorDepthLoss

    g = torch.Generator().manual_seed(4321)
    H, W = 40, 56
    inp, out = {}, {}
    # ---- DSSIML1 per-pixel: [1,3,H,W] and [H,W,1] inputs (the reference's HWC branch only handles one channel: its second
    # test reads `pred.shape[-1] == 3` after pred has already been permuted)
    a3, b3 = torch.rand(1, 3, H, W, generator=g), torch.rand(1, 3, H, W, generator=g)
    a1, b1 = torch.rand(H, W, 1, generator=g), torch.rand(H, W, 1, generator=g)
    inp.update(dssim_a3=a3, dssim_b3=b3, dssim_a1=a1, dssim_b1=b1)
    out["dssim_pp_3"] = DSSIML1()(a3, b3)
    out["dssim_pp_1"] = DSSIML1(kernel_size=5, alpha=0.6)(a1, b1)
    # ---- SensorDepthLoss
    R, S = 64, 12
    depth_pred = 1 + 3 * torch.rand(R, 1, generator=g)
    sensor = 1 + 3 * torch.rand(R, generator=g)
    sensor[torch.rand(R, generator=g) < 0.2] = 0.0
    starts = torch.sort(0.2 + 5 * torch.rand(R, S, 1, generator=g), dim=1).values
    sdf = 0.5 * torch.randn(R, S, 1, generator=g)
    dnorm = 0.9 + 0.2 * torch.rand(R, 1, generator=g)
    inp.update(sd_depth_pred=depth_pred, sd_sensor=sensor, sd_starts=starts, sd_sdf=sdf, sd_dnorm=dnorm)
    rs = types.SimpleNamespace(frustums=types.SimpleNamespace(starts=starts))
    l1, fs, sd = SensorDepthLoss(truncation=0.25)({"sensor_depth": sensor},
                                                  {"depth": depth_pred, "ray_samples": rs, "field_outputs": {"sdf": sdf},
                                                   "directions_norm": dnorm})
    out["sensor_l1"], out["sensor_fs"], out["sensor_sdf"] = l1, fs, sd
    # ---- AdaptiveDepth / AdaptiveNormal
    pd = 0.5 + 4 * torch.rand(H, W, 1, generator=g)
    gd = pd + 0.3 * torch.randn(H, W, 1, generator=g)
    gd[torch.rand(H, W, 1, generator=g) < 0.15] = 0.0
    img = torch.rand(H, W, 3, generator=g).clamp(min=10 / 255.0)
    conf = (torch.rand(H, W, 1, generator=g) > 0.3).float()
    pn = torch.nn.functional.normalize(torch.randn(H, W, 3, generator=g), dim=-1)
    gn = torch.nn.functional.normalize(pn + 0.15 * torch.randn(H, W, 3, generator=g), dim=-1)
    inp.update(ad_pd=pd, ad_gd=gd, ad_img=img, ad_conf=conf, an_pn=(pn + 1) / 2, an_gn=(gn + 1) / 2)
    with cpu_redirect():
        for step in (100, 9000):
            out[f"adaptive_depth_{step}"] = AdaptiveDepth()(pd, gd, img, gd > 0.1, conf, step)
    for step in (100, 20000):
        out[f"adaptive_normal_{step}"] = AdaptiveNormal()((pn + 1) / 2, (gn + 1) / 2, step)
    # ---- LocalPearsonDepthLoss (random windows recorded)
    Hp, Wp = 72, 100
    lp, lg = 1 + torch.rand(Hp, Wp, generator=g), 1 + torch.rand(Hp, Wp, generator=g)
    inp.update(lp_pred=lp, lp_gt=lg)
    with cpu_redirect() as red:
        out["local_pearson"] = LocalPearsonDepthLoss()(lp, lg, box_p=24, p_corr=0.5)
    inp["lp_x0"], inp["lp_y0"] = red.draws[0], red.draws[1]
    arrs = {f"in_{k}": v.detach().numpy() for k, v in inp.items()}
    arrs.update({f"out_{k}": torch.as_tensor(v).detach().numpy() for k, v in out.items()})
    np.savez_compressed(os.path.join(OUT, "dn_reference_losses2.npz"), **arrs)
    print({k: tuple(v.shape) for k, v in arrs.items() if k.startswith("out_")})


if __name__ == "__main__":
    main()
