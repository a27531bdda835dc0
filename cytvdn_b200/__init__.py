"""cytvdn_b200 -- the TV-denoising hot path of cyTVDN (``tv.denoise3D`` / ``tv.denoise4D``) on NVIDIA
B200, as hand-written CUDA for sm_100a behind a C ABI (``include/cytvdn_b200.h``).

    import cytvdn_b200 as tv
    recon, b_norm, delta = tv.denoise4D(data, mu, iterations=100, FISTA=True)

The names exported here are the ones `cyTVDN/__init__.py:1` exports for this path.
"""
from .tv import *  # noqa: F401,F403
from .tv import __all__ as _tv_all
from ._lib import CytvdnError, device_count, launch_count, LIB_PATH  # noqa: F401

__all__ = list(_tv_all) + ["CytvdnError", "device_count", "launch_count"]
__version__ = "0.1.0"
