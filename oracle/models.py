"""ORACLE (test infrastructure): restatement of the reference generator / discriminator,
``building_gan/src/models.py:14-155`` and ``:158-245``, on top of oracle/pyg.py.

PINNED: tests/test_oracle_golden.py loads weights recorded from the UNMODIFIED reference
``models.py`` (imported by oracle/make_golden.py with oracle/pyg_shim) and checks that this
file reproduces its outputs and gradients.  Sub-module names and indices equal the reference's
so ``state_dict`` keys interchange:
G: matched_features_encoder.{0..14}, mlp_encoder.{0..14}, encoder.module_{0..55}, decoder.{0..12};
D: mlp_encoder.{0..3}, encoder.module_{0..23}, decoder.{0..6}.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from . import pyg

CONVS: Dict[str, Callable[[int, int], nn.Module]] = {
    "GCNCONV": pyg.GCNConv,
    "GRAPHCONV": pyg.GraphConv,
    "GATCONV": pyg.GATConv,
    "GATV2CONV": pyg.GATv2Conv,
}


def conv_factory(kind: str) -> Callable[[int, int], nn.Module]:
    """models.py:22-31 / 166-175: unknown conv type => ValueError."""
    if kind not in CONVS:
        raise ValueError(f"Invalid conv_type: {kind}")
    return CONVS[kind]


def ln_mlp(widths: List[int], final_plain: Optional[int] = None) -> nn.Sequential:
    """[Linear, LayerNorm, LeakyReLU(0.2)] per consecutive width pair (models.py:33-66,92-113);
    ``final_plain`` appends one bare Linear (the generator's logits layer, models.py:112)."""
    mods: List[nn.Module] = []
    for a, b in zip(widths[:-1], widths[1:]):
        mods += [nn.Linear(a, b), nn.LayerNorm(b), nn.LeakyReLU(0.2)]
    if final_plain is not None:
        mods.append(nn.Linear(widths[-1], final_plain))
    return nn.Sequential(*mods)


def hourglass_widths(hidden: int, repeat: int) -> List[int]:
    """Channel schedule of the GNN encoder: halve ``repeat`` times, then double ``repeat``
    times (models.py:68-88, 187-208)."""
    down = [hidden // (2 ** k) for k in range(repeat + 1)]
    return down + down[-2::-1]


def gnn_stack(conv: Callable[[int, int], nn.Module], widths: List[int], input_args: str) -> pyg.Sequential:
    """tgnn.Sequential of [conv, GraphNorm, ReLU(inplace), Dropout(0.2)] blocks."""
    entries = []
    for a, b in zip(widths[:-1], widths[1:]):
        entries += [(conv(a, b), f"{input_args} -> x"), pyg.GraphNorm(b), nn.ReLU(True), nn.Dropout(0.2)]
    return pyg.Sequential(input_args, entries)


def type_matched_features(local_x: Tensor, local_type: Tensor, voxel_type: Tensor) -> Tensor:
    """models.py:122-129 / 230-237: every voxel of (ground-truth) type t receives the mean of the
    program-graph feature rows of type t, pooled over the WHOLE batch; zero when no such row."""
    out = torch.zeros((voxel_type.shape[0], local_x.shape[1]), device=local_x.device, dtype=local_x.dtype)
    for t in torch.unique(voxel_type):
        rows = local_type == t
        if rows.sum() > 0:
            out[voxel_type == t] = local_x[rows].mean(dim=0)
    return out


def gumbel_straight_through(logits: Tensor, noise: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """models.py:150-153.  ``noise`` (Gumbel(0,1) samples, same shape as logits) may be injected;
    otherwise it is drawn exactly as torch.nn.functional.gumbel_softmax draws it."""
    if noise is None:
        noise = -torch.empty_like(logits, memory_format=torch.legacy_contiguous_format).exponential_().log()
    soft = ((logits + noise) / 1.0).softmax(-1)
    hard = torch.zeros_like(soft)
    hard.scatter_(-1, soft.argmax(dim=1, keepdim=True), 1.0)
    return hard - soft.detach() + soft, soft


class OracleGenerator(nn.Module):
    def __init__(self, configuration, local_graph_dim: int, voxel_graph_dim: int):
        super().__init__()
        c = configuration
        self.configuration = c
        conv = conv_factory(c.GENERATOR_CONV_TYPE)
        le, gh = c.LOCAL_ENCODER_HIDDEN_DIM, c.GENERATOR_HIDDEN_DIM
        self.matched_features_encoder = ln_mlp([local_graph_dim] + [le] * (c.LOCAL_GRAPH_ENCODER_REPEAT + 1))
        self.mlp_encoder = ln_mlp([le + voxel_graph_dim + c.Z_DIM] + [gh] * (c.GENERATOR_MLP_ENCODER_REPEAT + 1))
        widths = hourglass_widths(gh, c.GENERATOR_ENCODER_REPEAT)
        self.encoder = gnn_stack(conv, widths, c.INPUT_ARGS)
        dec_in = le + voxel_graph_dim + c.Z_DIM + widths[-1] + gh
        self.decoder = ln_mlp([dec_in, gh, gh // 2, gh // 4, gh // 8], final_plain=c.NUM_CLASSES)

    def forward(self, local_graph, voxel_graph, z: Tensor, gumbel_noise: Optional[Tensor] = None):
        matched = type_matched_features(local_graph.x, local_graph.type, voxel_graph.type)
        enc = self.matched_features_encoder(matched)
        zz = z.squeeze(0)
        x = self.mlp_encoder(torch.cat([enc, voxel_graph.x, zz], dim=-1))
        encoded = self.encoder(x=x, edge_index=voxel_graph.edge_index)
        logits = self.decoder(torch.cat([encoded, x, enc, voxel_graph.x, zz], dim=-1))
        hard, soft = gumbel_straight_through(logits, gumbel_noise)
        return logits, hard, soft


class OracleDiscriminator(nn.Module):
    def __init__(self, configuration, local_graph_dim: int, voxel_graph_dim: int):
        super().__init__()
        c = configuration
        self.configuration = c
        conv = conv_factory(c.DISCRIMINATOR_CONV_TYPE)
        dh = c.DISCRIMINATOR_HIDDEN_DIM
        self.mlp_encoder = nn.Sequential(
            nn.Linear(local_graph_dim + voxel_graph_dim + c.NUM_CLASSES, dh), nn.ReLU(True),
            nn.Linear(dh, dh), nn.ReLU(True),
        )
        self.encoder = gnn_stack(conv, hourglass_widths(dh, c.DISCRIMINATOR_ENCODER_REPEAT), c.INPUT_ARGS)
        tail: List[nn.Module] = [
            nn.Linear(dh, dh // 2), nn.ReLU(True),
            nn.Linear(dh // 2, dh // 4), nn.ReLU(True),
            nn.Linear(dh // 4, dh // 8), nn.ReLU(True),
            nn.Linear(dh // 8, 1),
        ]
        if not c.USE_WGANGP:
            tail.append(nn.Sigmoid())
        self.decoder = nn.Sequential(*tail)

    def forward(self, local_graph, voxel_graph, label_hard: Tensor) -> Tensor:
        matched = type_matched_features(local_graph.x, local_graph.type, voxel_graph.type)
        x = self.mlp_encoder(torch.cat([matched, voxel_graph.x, label_hard.squeeze(0)], dim=-1))
        return self.decoder(self.encoder(x=x, edge_index=voxel_graph.edge_index))
