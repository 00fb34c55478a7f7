"""ORACLE (test infrastructure, see oracle/__init__.py): the hyper-parameters the hot path reads, restated from the
reference's ``building_gan/src/config.py``: program map :9-30, normalisers :41-45, loop / loss weights :52-81,
architecture :89-106.  Standalone on purpose - the reference arm of bench.py and the oracle-side tests get their
configuration from here, not from the product package.  ``tests/test_oracle_golden.py`` checks every value against the
reference's own ``Configuration`` when /root/reference is present (build container) and against the product's."""
from __future__ import annotations

import torch


class Configuration:
    # config.py:9-30
    VOID_OLD = -1
    LOBBY_CORRIDOR, RESTROOM, STAIRS, ELEVATOR, OFFICE, MECHANICAL_ROOM, VOID = range(7)
    NUM_CLASSES = 7
    # config.py:41-45
    NORMALIZATION_FACTOR_FLOOR_LEVEL = 10
    NORMALIZATION_FACTOR_DIMENSION = 11
    NORMALIZATION_FACTOR_LOCATION = 11
    NORMALIZATION_FACTOR_COORDINATE = 42
    NORMALIZATION_FACTOR_SITE = 1600
    # config.py:52-81
    SEED = 777
    BATCH_SIZE = 512
    N_CRITIC = 5
    LEARNING_RATE_GENERATOR = 0.0002
    LEARNING_RATE_DISCRIMINATOR = 0.0002
    LAMBDA_RATIO = 0.1
    LAMBDA_RATIO_VOID = 0.1
    LAMBDA_LABEL = 0.0
    LAMBDA_ADV = 1.0
    LAMBDA_FAR = 0.1
    LAMBDA_GP = 10.0
    BETAS = (0.5, 0.999)
    METRICS_AVERAGE = "macro"
    DEVICE = "cuda" if torch.cuda.is_available() else "cpu"
    # config.py:89-106
    GENERATOR_CONV_TYPE = "GATCONV"
    GENERATOR_ENCODER_REPEAT = 7
    GENERATOR_HIDDEN_DIM = 128
    DISCRIMINATOR_CONV_TYPE = "GATCONV"
    DISCRIMINATOR_ENCODER_REPEAT = 3
    DISCRIMINATOR_HIDDEN_DIM = 64
    Z_DIM = 128
    LOCAL_GRAPH_ENCODER_REPEAT = 4
    LOCAL_ENCODER_HIDDEN_DIM = 128
    ENCODER_DROPOUT_RATE = 0.2
    GENERATOR_MLP_ENCODER_REPEAT = 4
    INPUT_ARGS = "x, edge_index"
    USE_WGANGP = True

    @classmethod
    def names(cls):
        return [k for k in vars(cls) if k.isupper()]
