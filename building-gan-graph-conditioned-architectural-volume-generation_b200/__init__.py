"""B200-native hot path of Building-GAN (voxel-graph message passing).

Python host side (this package) mirrors the reference's ``models.py`` / ``data.py`` interface for
the path; the arithmetic lives in ``libbgb200.so`` (hand-written sm_100a CUDA, C ABI in
``include/bg_b200.h``).  Import name: ``building_gan_b200`` (the directory name carries the
reference repository's name and is not a Python identifier; ``building_gan_b200/__init__.py`` at the
repo root aliases it).
"""
from . import lib  # noqa: F401
from .config import Configuration  # noqa: F401
from .graph import Batch, Data, VoxelCSR, collate_fn  # noqa: F401

__all__ = ["lib", "Configuration", "Batch", "Data", "VoxelCSR", "collate_fn"]
