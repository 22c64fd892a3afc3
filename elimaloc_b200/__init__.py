"""elimaloc_b200 — B200-native (sm_100a) implementation of ELiMaLoc's pcm_matching registration hot path.

The product is libelimaloc_b200.so (hand-written CUDA behind a C ABI, include/elimaloc_b200.h); this package is the
thin Python mirror of the reference's Registration / VoxelHashMap interface used by the tests and bench.py."""
from ._capi import AVGICP, GICP, P2P, VGICP, ElmError, RegConfig, lib  # noqa: F401
from .ekf import EkfAlgorithm, make_ekf_config, make_measurement  # noqa: F401
from .pipeline import Queues, ScanPipeline, build_deskew_tables  # noqa: F401
from .registration import Registration, RegistrationConfig, VoxelHashMap, read_pcd_xyz, shape_pcm_covariance  # noqa: F401


def device_count():
    return lib().elm_device_count()
