"""Import alias: ``building_gan_b200`` -> ``building-gan-graph-conditioned-architectural-volume-generation_b200/``.

The package directory is named after the reference repository (hyphens, not importable by name);
this shim points ``__path__`` at it and runs its ``__init__`` so ``import building_gan_b200`` and
``from building_gan_b200.models import VoxelGNNGenerator`` work from the repo root.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "building-gan-graph-conditioned-architectural-volume-generation_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
