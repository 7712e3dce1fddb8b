import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dn_splatter_b200.rasterize as R
orig = R._stream
def dbg():
    st = torch.cuda.current_stream()
    if st.cuda_stream == 0:
        print("WARNING: launch on the default stream", flush=True)
    return orig()
R._stream = dbg
import dn_splatter_b200.regularization_strategy as RS
RS._stream = dbg
import pytest
sys.exit(pytest.main(["tests/test_gpu_model.py", "-q", "-m", "gpu", "-x", "-k", "cuda_graph", "-s"]))
