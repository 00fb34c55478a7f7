"""Synthetic "6types-like" buildings as this package's ``Data`` objects.

The generator itself lives in ``workloads/synth.py`` at the repository root (neutral code shared with the oracle-side
reference arm of bench.py, which must not import this package); this module wraps its field dicts in ``graph.Data``.
The real dataset (building_gan/data/6types-raw_data-10000.zip) is a Git-LFS pointer, see workloads/synth.py.
"""
from __future__ import annotations

import os
import sys
from typing import Tuple

try:
    from workloads import synth as _w
except ImportError:  # package imported from elsewhere than the repository root
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from workloads import synth as _w

from .config import Configuration
from .graph import Data

grid_arrays, raw_building, process_raw = _w.grid_arrays, _w.raw_building, _w.process_raw


def _pair(fields) -> Tuple[Data, Data]:
    return Data(**fields[0]), Data(**fields[1])


def building_pair(building_id: int, cfg=Configuration, **grid_kw) -> Tuple[Data, Data]:
    """(local Data, voxel Data) for one synthetic building - what ``GraphDataset[i]`` returns."""
    return _pair(_w.building_fields(building_id, cfg, **grid_kw))


def building_pair_fast(building_id: int, cfg=Configuration, **grid_kw) -> Tuple[Data, Data]:
    """Same pair, bit for bit (tests/test_collate.py), built straight from arrays (no JSON detour)."""
    return _pair(_w.building_fields_fast(building_id, cfg, **grid_kw))


def large_grid_pair(building_id: int, floors: int = 10, ny: int = 100, nx: int = 100, cfg=Configuration) -> Tuple[Data, Data]:
    """BASELINE config 4: one F x Y x X irregular grid (default 1e5 voxels, 576 000 directed edges)."""
    return _pair(_w.large_grid_fields(building_id, floors, ny, nx, cfg))
