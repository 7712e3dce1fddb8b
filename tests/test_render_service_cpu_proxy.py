"""Host logic of the render-all-views service through the CPU proxy: order, keys, eval-mode handling."""
import torch

from tests.cpu_proxy import cpu_proxy


def test_view_renderer_yields_every_view_in_order_and_restores_mode():
    from dn_splatter_b200.cameras import Cameras
    from dn_splatter_b200.dn_model import DNSplatterModelConfig
    from dn_splatter_b200.render_service import ViewRenderer
    from dn_splatter_b200.synthetic import make_scene, ring_cameras

    W, H = 40, 32
    cams = [Cameras(c["c2w"][None], c["fx"], c["fy"], c["cx"], c["cy"], W, H) for c in ring_cameras(4, W, H)]
    with cpu_proxy():
        m = DNSplatterModelConfig(random_init=True, num_random=16, background_color="black").setup(device="cpu")
        m.load_gaussians(make_scene(50, seed=1))
        m.step = 5000
        m.train()
        direct = []
        m.eval()
        with torch.no_grad():
            for c in cams:
                direct.append(m.get_outputs(c)["depth"].clone())
        m.train()
        for to_host in (False, True):
            got = list(ViewRenderer(m, keys=("rgb", "depth"), to_host=to_host).render(cams))
            assert [i for i, _ in got] == [0, 1, 2, 3]
            assert all(set(d) == {"rgb", "depth"} for _, d in got)
            if not to_host:
                for (i, d), want in zip(got, direct):
                    assert torch.equal(d["depth"], want)
            assert m.training  # mode restored
