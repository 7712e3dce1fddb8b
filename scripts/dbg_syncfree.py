import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.helpers import cuda_outputs, scene_and_camera
import dn_splatter_b200.rasterize as R
params, cam = scene_and_camera(5000, 256, 192, view=1)
for i in range(5):
    try:
        t=time.perf_counter()
        _, o = cuda_outputs(params, cam, sync_free=True)
        torch.cuda.synchronize()
        print(i, "ok cap", o.info["n_isects"], "count", int(o.info["n_isects_dev"]), time.perf_counter()-t)
    except Exception as e:
        print(i, "ERR", e)
        print(torch.cuda.memory_summary()[:300])
        break
