"""dn_splatter_b200 — B200-native (sm_100a) depth+normal Gaussian rasterizer behind the dn-splatter
plugin surface.  The arithmetic lives in libdnr_b200.so (C ABI: include/dnr.h); see DESIGN.md."""
from .rasterize import RasterOutput, RasterSettings, dn_rasterize, get_viewmat  # noqa: F401

__version__ = "0.1.0"
