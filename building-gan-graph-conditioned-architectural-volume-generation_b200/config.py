"""Hyper-parameter object with the attribute surface of the reference's ``Configuration``
(building_gan/src/config.py:9-157): the models read ``configuration.GENERATOR_CONV_TYPE``,
``*_HIDDEN_DIM``, ``*_ENCODER_REPEAT``, ``Z_DIM``, ``NUM_CLASSES``, ``INPUT_ARGS``,
``USE_WGANGP``, ``DEVICE`` ... (models.py:22-31,35-113,166-225) and the trainer reads the
loss weights (trainer.py:314,340-383).  The reference's own object can be passed instead -
nothing here is required by the kernels; it exists because /root/reference is not
importable at run time on the GPU box.
"""
from __future__ import annotations

import random
from typing import Any, Dict

import numpy as np
import torch

# program-type ids (config.py:9-30); -1 in the raw JSON is remapped to VOID=6 (data.py:307-308)
_PROGRAMS = ("LOBBY_CORRIDOR", "RESTROOM", "STAIRS", "ELEVATOR", "OFFICE", "MECHANICAL_ROOM", "VOID")
_COLORS = ("brown", "red", "yellow", "green", "blue", "orange", "gray")

_DEFAULTS: Dict[str, Any] = dict(
    # normalisers applied at preprocessing time (config.py:41-45)
    NORMALIZATION_FACTOR_FLOOR_LEVEL=10,
    NORMALIZATION_FACTOR_DIMENSION=11,
    NORMALIZATION_FACTOR_LOCATION=11,
    NORMALIZATION_FACTOR_COORDINATE=42,
    NORMALIZATION_FACTOR_SITE=1600,
    LOCAL_DATA_SUFFIX="_local.pt",
    VOXEL_DATA_SUFFIX="_voxel.pt",
    # loop / optimiser (config.py:52-79)
    NUM_WORKERS=3,
    EPOCHS=5000,
    SEED=777,
    TRAIN_SPLIT_RATIO=0.65,
    VALIDATION_SPLIT_RATIO=0.25,
    TEST_SPLIT_RATIO=0.10,
    DATA_POINT=None,
    DATA_SLICER=int(1e10),
    BATCH_SIZE=512,
    N_CRITIC=5,
    LEARNING_RATE_GENERATOR=2e-4,
    LEARNING_RATE_DISCRIMINATOR=2e-4,
    LAMBDA_RATIO=0.1,
    LAMBDA_RATIO_VOID=0.1,
    LAMBDA_LABEL=0.0,
    LAMBDA_ADV=1.0,
    LAMBDA_FAR=0.1,
    LAMBDA_GP=10.0,
    BETAS=(0.5, 0.999),
    F1_SCORE_TRAIN_WEIGHT=0.05,
    F1_SCORE_VALIDATION_WEIGHT=1.0,
    METRICS_AVERAGE="macro",
    # architecture (config.py:89-106)
    GENERATOR_CONV_TYPE="GATCONV",
    GENERATOR_ENCODER_REPEAT=7,
    GENERATOR_HIDDEN_DIM=128,
    DISCRIMINATOR_CONV_TYPE="GATCONV",
    DISCRIMINATOR_ENCODER_REPEAT=3,
    DISCRIMINATOR_HIDDEN_DIM=64,
    Z_DIM=128,
    LOCAL_GRAPH_ENCODER_REPEAT=4,
    LOCAL_ENCODER_HIDDEN_DIM=128,
    ENCODER_DROPOUT_RATE=0.2,  # defined but unused by the reference (dropout is hard-coded 0.2)
    GENERATOR_MLP_ENCODER_REPEAT=4,
    INPUT_ARGS="x, edge_index",
    USE_WGANGP=True,
)


class Configuration:
    """Attribute-compatible stand-in for the reference ``Configuration``; override by assignment
    exactly like ``train.py:16`` / ``sanity.py:14-15`` do."""

    VOID_OLD = -1
    COLORS = {i: c for i, c in enumerate(_COLORS)}
    NUM_CLASSES = len(_PROGRAMS)
    DEVICE = "cuda" if torch.cuda.is_available() else "cpu"

    def __init__(self, sanity_checking: bool = False):
        self.SANITY_CHECKING = sanity_checking
        if sanity_checking:
            self.BATCH_SIZE, self.DATA_POINT = 1, 77

    @property
    def SPLIT_RATIOS(self):
        return [self.TRAIN_SPLIT_RATIO, self.VALIDATION_SPLIT_RATIO, self.TEST_SPLIT_RATIO]

    def to_dict(self) -> Dict[str, Any]:
        out = {k: getattr(self, k) for k in _DEFAULTS}
        out.update({name: i for i, name in enumerate(_PROGRAMS)})
        out.update(VOID_OLD=self.VOID_OLD, NUM_CLASSES=self.NUM_CLASSES, DEVICE=self.DEVICE)
        return out

    @staticmethod
    def set_seed(seed: int = _DEFAULTS["SEED"]) -> None:
        torch.manual_seed(seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(seed)
        np.random.seed(seed)
        random.seed(seed)
        Configuration.SEED = seed


for _i, _name in enumerate(_PROGRAMS):
    setattr(Configuration, _name, _i)
for _k, _v in _DEFAULTS.items():
    setattr(Configuration, _k, _v)
