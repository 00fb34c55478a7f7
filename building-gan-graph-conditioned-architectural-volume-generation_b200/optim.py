"""Adam over the flat parameter / gradient buckets of a B200 model: ONE kernel launch per ``step()``.

The reference builds ``torch.optim.Adam(model.parameters(), lr=..., betas=...)`` (train.py:36-37, sanity.py:34-35) and
steps it after every backward (trainer.py:481,495).  ``torch.optim.Adam`` keeps working on the drop-in models; on the
batch-32 training step its foreach path costs ~0.5 ms (D) / ~1.1 ms (G) of host time per call (6 calls per step, ~190
parameter tensors, ~14 launches each).  This class is the same optimiser behind the same constructor
(``from building_gan_b200.optim import Adam`` is the one-line change in train.py): parameters, gradients, ``exp_avg`` and
``exp_avg_sq`` of a model live in four index-aligned flat fp32 buffers (the gradient one is the bucket the backward kernels
accumulate into and the NCCL all-reduce payload), and ``step()`` is one ``bg_adam_flat`` launch.

* Same update as torch (``_multi_tensor_adam``, amsgrad / maximize off): tests/test_optim.py compares it against
  ``torch.optim.Adam`` over many steps.
* ``state_dict()`` / ``load_state_dict()`` use torch.optim.Adam's format in both directions (``states.pt``,
  trainer.py:715-736): per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq`` entries are views of the flat buffers.
* ``param_groups[0]["lr"]`` is read at every step, so ``CosineAnnealingLR`` (train.py:38) drives it unchanged.
* One documented difference: ``zero_grad()`` zeroes the bucket instead of dropping the ``.grad`` tensors, and a parameter
  that received no gradient is updated with a zero gradient (torch skips ``grad is None`` parameters).  Every parameter of
  both models receives a gradient in every backward of the reference step, so the two never differ there.
* ``capturable=True`` keeps the step count in a device counter that ``step()`` increments with a tiny in-stream add, so
  the optimiser can sit inside a CUDA graph.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
from torch import Tensor

from . import lib

_TORCH_ADAM_DEFAULTS = dict(amsgrad=False, maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                            decoupled_weight_decay=False)


def _owner(params: List[torch.nn.Parameter]):
    from . import models

    for m in models.live_models():
        mine = models._param_list(m)
        if len(mine) == len(params) and all(a is b for a, b in zip(mine, params)):
            return m
    raise ValueError("building_gan_b200.optim.Adam takes the full ``model.parameters()`` of ONE VoxelGNNGenerator / "
                     "VoxelGNNDiscriminator of this package (use torch.optim.Adam for anything else)")


class Adam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, amsgrad: bool = False, *, maximize: bool = False, capturable: bool = False, **kw):
        if amsgrad or maximize or kw.get("decoupled_weight_decay") or kw.get("differentiable"):
            raise NotImplementedError("building_gan_b200.optim.Adam: amsgrad / maximize / decoupled_weight_decay / differentiable are not on the B200 path")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or not 0.0 <= weight_decay:
            raise ValueError(f"Invalid Adam hyper-parameters: lr={lr} betas={betas} eps={eps} weight_decay={weight_decay}")
        params = list(params)
        if params and isinstance(params[0], dict):
            raise ValueError("building_gan_b200.optim.Adam: one parameter group (the model's own parameter list) only")
        self._model = _owner(params)
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, **_TORCH_ADAM_DEFAULTS)
        defaults["capturable"] = bool(capturable)
        super().__init__(params, defaults)
        self._t = 0
        self._m: Optional[Tensor] = None
        self._v: Optional[Tensor] = None
        self._t_dev: Optional[Tensor] = None
        self._peer = None  # (BgPeers, epoch, ticket) once dist.PeerSync.attach() fused the gradient exchange into step()

    # -- flat state ---------------------------------------------------------------------------------
    def _flats(self):
        st, params = self._model._native, self.param_groups[0]["params"]
        pflat = st.flat_params(params)
        if self._m is None or self._m.device != pflat.device:
            old = [(self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"]) if p in self.state and "exp_avg" in self.state[p] else None
                   for p in params]
            self._m, self._v = torch.zeros_like(pflat), torch.zeros_like(pflat)
            self._t_dev = torch.full((1,), self._t, dtype=torch.int64, device=pflat.device)
            self._bind_state(old)
        return pflat, self._m, self._v

    def _bind_state(self, old=None) -> None:
        """``self.state[p]`` = torch.optim.Adam's per-parameter entries, as views of the flat moment buffers."""
        lay, params = self._model._native.layout, self.param_groups[0]["params"]
        step = torch.tensor(float(self._t), dtype=torch.float32)
        for i, (p, name) in enumerate(zip(params, lay.names)):
            mv, vv = lay.view(self._m, name), lay.view(self._v, name)
            if old is not None and old[i] is not None:
                mv.copy_(old[i][0])
                vv.copy_(old[i][1])
            self.state[p] = {"step": step.clone(), "exp_avg": mv, "exp_avg_sq": vv}

    # -- torch.optim.Optimizer interface ------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g = self.param_groups[0]
        params, st = g["params"], self._model._native
        if params[0].grad is None and all(p.grad is None for p in params):
            return loss  # nothing was back-propagated since zero_grad(set_to_none=True): torch.optim.Adam does nothing either
        pflat, m, v = self._flats()
        if params[0].grad is None or st.views is None or params[0].grad.data_ptr() != st.views[0].data_ptr() \
                or params[-1].grad is None or params[-1].grad.data_ptr() != st.views[-1].data_ptr():
            st.bind_grads(params)  # gradients delivered by autograd (BG_GRADS=autograd / op-by-op executor): gather them
        self._t += 1
        b1, b2 = g["betas"]
        if g["capturable"]:
            self._t_dev.add_(1)
        if self._peer is not None:  # data parallel: all ranks' buckets are averaged inside the same launch (csrc/bg_p2p.cu)
            peers, epoch, ticket = self._peer
            lib.p2p_allreduce_adam_(peers, epoch, ticket, pflat.numel(), pflat, m, v, None, float(g["lr"]), float(b1), float(b2),
                                    float(g["eps"]), float(g["weight_decay"]), self._t, self._t_dev if g["capturable"] else None)
            return loss
        lib.adam_flat_(pflat, st.bucket, m, v, float(g["lr"]), float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]),
                       self._t, self._t_dev if g["capturable"] else None)
        return loss

    def zero_grad(self, set_to_none: bool = True) -> None:
        st, params = self._model._native, self.param_groups[0]["params"]
        if st.bucket is not None and st.views is not None and params[0].grad is not None and params[-1].grad is not None \
                and params[0].grad.data_ptr() == st.views[0].data_ptr() and params[-1].grad.data_ptr() == st.views[-1].data_ptr():
            st.bucket.zero_()  # every p.grad is a view of the bucket: one memset
            return
        super().zero_grad(set_to_none)

    def state_dict(self):
        if self._m is not None:
            if self.param_groups[0]["capturable"]:
                self._t = int(self._t_dev.item())
            for p in self.param_groups[0]["params"]:
                if p in self.state:
                    self.state[p]["step"] = torch.tensor(float(self._t), dtype=torch.float32)
        return super().state_dict()

    def load_state_dict(self, state_dict) -> None:
        super().load_state_dict(state_dict)
        params = self.param_groups[0]["params"]
        loaded = [self.state.get(p) for p in params]
        have = [s for s in loaded if s and "exp_avg" in s]
        if not have:
            self._t, self._m, self._v = 0, None, None
            return
        if len(have) != len(params):
            raise ValueError("building_gan_b200.optim.Adam.load_state_dict: optimiser state covers only part of the parameters")
        self._t = int(float(have[0]["step"]))
        old = [(s["exp_avg"], s["exp_avg_sq"]) for s in loaded]
        pflat = self._model._native.flat_params(params)
        self._m, self._v = torch.zeros_like(pflat), torch.zeros_like(pflat)
        self._t_dev = torch.full((1,), self._t, dtype=torch.int64, device=pflat.device)
        self._bind_state(old)
