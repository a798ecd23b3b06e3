"""vln_ver_b200 -- B200-native (sm_100a) implementation of VER's 2D->3D volumetric lifting
hot path behind the reference's mmcv registry API.  Importing the package loads
libver_b200.so (raises if it is missing: there is no fallback) and registers the modules."""
from . import _lib, ops, registry                                  # noqa: F401
from ._lib import VerError, launch_count                           # noqa: F401
from .config import vocc_head_cfg                                  # noqa: F401
from .modules import *                                             # noqa: F401,F403
from .modules import set_compute_dtype                             # noqa: F401
from .registry import HEADS, build_from_cfg                        # noqa: F401

__version__ = '0.1.0'


def build_head(cfg):
    """registry build of a `pts_bbox_head` dict (vocc.py:87-195)."""
    return build_from_cfg(cfg, HEADS)
