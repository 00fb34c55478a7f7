"""CUDA-graph replay of the repeated parts of the training step (reference trainer.py:467-495).

One step runs the SAME critic update N_CRITIC (=5) times on the same batch: sampling pass of the generator (no grad), then
D(real) / D(fake) / gradient penalty with its second-order backward, then Adam (trainer.py:468-481).  At batch 32 that is
~370 kernel launches per update and the step is bound by the host's launch rate (bench history in profiles/).  Batch shapes
differ from step to step, so a graph cannot outlive its step; ``GraphedStep`` therefore captures, once per step,

* S - one sampling pass  (z drawn in-graph, generator forward, straight-through labels), on the sampling stream, and
* C - one critic update  (zero_grad, the three critic passes on their lanes, backward, lane merge, Adam), on the main stream,

and replays each N_CRITIC times: S(k+1) runs beside C(k).  The generator update (once per step) runs eagerly.  Capturing
costs about what running the update once costs on the host, so the host side of a step drops from 5 updates to ~1.4.

What makes the replays differ (a graph freezes every kernel argument):
* z and the gradient-penalty mixing factor come from torch's device generator, which torch re-seeds per replay;
* dropout masks / Gumbel noise come from the in-kernel Philox with a per-call offset: the captured kernels add a device
  counter (bg_set_rng_base) that is set before every replay so that replay r, call j uses the ticket an eager run of the
  same call sequence would have used;
* Adam's step count lives in a device counter (optim.Adam(capturable=True) semantics, switched on for the captured step).

Not supported here (use step.train_step): rng="cpu" (host draws cannot be captured), torch.optim.Adam for the critic.
All four conv types are covered (the non-default ones through the op-by-op executor).
"""
from __future__ import annotations

import collections

import torch

from . import lib, models
from . import step as _step
from .optim import Adam


class _Captured:
    """A graph plus every tensor allocated from its private pool that must stay alive while it may still run."""
    __slots__ = ("graph", "out", "calls", "done")

    def __init__(self):
        self.graph, self.out, self.calls, self.done = torch.cuda.CUDAGraph(), None, 0, None


class _Prepared:
    """The captures and static buffers of one step (GraphedStep._prepare)."""
    __slots__ = ("lb", "vb", "sample", "losses", "S", "C")


class GraphedStep:
    def __init__(self, generator, discriminator, opt_g, opt_d, cfg, grad_sync=None):
        """``grad_sync(model)``: the data-parallel gradient all-reduce (dist.GradSync), called after each backward and before
        the optimiser step; for the critic it is captured into the graph (NCCL collectives are capturable; the communicator is
        created by the eager first step)."""
        self.grad_sync = grad_sync
        if not isinstance(opt_d, Adam):
            raise TypeError("GraphedStep: the critic optimiser must be building_gan_b200.optim.Adam (its step is captured)")
        if models.GRAD_MODE != "bucket" or models.RNG_MODE != "philox" or not cfg.USE_WGANGP:
            raise NotImplementedError("GraphedStep covers WGAN-GP with BG_GRADS=bucket and BG_RNG=philox")
        # GCNCONV / GRAPHCONV / GATV2CONV models run the op-by-op executor (models._executor_for): its launches (library kernels +
        # torch's dropout / Gumbel draws from the graph-registered generator + autograd's gradient accumulation) are all
        # stream-ordered and allocation-pool safe, so the same capture / replay scheme applies to them.
        self.G, self.D, self.opt_g, self.opt_d, self.cfg = generator, discriminator, opt_g, opt_d, cfg
        self.dev = next(discriminator.parameters()).device
        # the stream of the gradient-penalty chain (the step's critical path) outranks the lanes (-1) and the weight-gradient
        # side streams (0): 12.8 -> 12.1 ms per step (the batched weight-gradient launches fill every SM for ~40 us)
        self.main = torch.cuda.Stream(device=self.dev, priority=_step.stream_priorities()[0])
        self.lanes = _step.Lanes.get(self.dev)
        self.pool_s, self.pool_c = torch.cuda.graph_pool_handle(), torch.cuda.graph_pool_handle()
        self.base_s = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.base_c = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._alive = collections.deque()  # captured graphs of the last steps (their replays may still be queued)
        self._warm = False
        self._prepared = None  # the next step, prepared ahead of its call (see _prepare)
        self.last_losses = None  # [N_CRITIC + 1] device tensor of the last step's losses (critic updates, generator update)

    # ------------------------------------------------------------------------------------------------------------
    def __call__(self, local_graph, voxel_graph, sync_losses="step", next_batch=None):
        """One training step; same return value as step.train_step (losses as floats for sync_losses="step", as 0-dim
        device tensors for False).  ``next_batch`` = (local_graph, voxel_graph) of the FOLLOWING call, when the caller knows it (a
        prefetching data loader does): its graphs are then captured during this step, off the critical path."""
        caller = torch.cuda.current_stream()
        self.main.wait_stream(caller)
        with torch.cuda.stream(self.main), _step.lanes_pdl_scope():
            if not self._warm:
                # first call: the eager overlapped step on the same streams creates every lazily cached buffer (per-stream
                # workspaces, gradient / lane buckets, flat parameters, Adam state) outside any graph pool
                out = _step.train_step(self.G, self.D, self.opt_g, self.opt_d, local_graph, voxel_graph, self.cfg, rng="device",
                                       grad_sync=self.grad_sync, sync_losses=sync_losses if sync_losses == "step" else False,
                                       overlap=True)
                self._warm = True
            else:
                out = self._run(local_graph, voxel_graph, sync_losses, next_batch)
        caller.wait_stream(self.main)
        return out

    # ------------------------------------------------------------------------------------------------------------
    def _capture_sampling(self, lb, vb, n) -> _Captured:
        cap, gen = _Captured(), self.lanes.gen
        with torch.cuda.stream(gen):
            gen.wait_stream(self.main)
            lib.set_rng_base(self.base_s)
            c0 = models._philox_calls
            try:
                cap.graph.capture_begin(pool=self.pool_s, capture_error_mode="thread_local")
                with torch.no_grad():
                    z = torch.randn(1, n, self.cfg.Z_DIM, device=self.dev)
                    _, hard, soft = self.G(lb, vb, z)
                    cap.out = torch.stack([hard, soft])  # [2, N, K]
                cap.graph.capture_end()
            finally:
                lib.set_rng_base(None)
            cap.calls = models._philox_calls - c0
        return cap

    def _capture_critic(self, lb, vb, sample) -> _Captured:
        cap, cfg = _Captured(), self.cfg
        g = self.opt_d.param_groups[0]
        was = g["capturable"]
        if not was:  # move the step count into the device counter for the captured update
            self.opt_d._flats()
            self.opt_d._t_dev.fill_(self.opt_d._t)
            g["capturable"] = True
        lib.set_rng_base(self.base_c)
        c0 = models._philox_calls
        try:
            cap.graph.capture_begin(pool=self.pool_c, capture_error_mode="thread_local")
            self.opt_d.zero_grad()
            d_loss = _step.discriminator_loss(self.D, lb, vb, sample[0].unsqueeze(0), sample[1].unsqueeze(0), cfg, "device",
                                              self.lanes, None)
            d_loss.backward()
            self.main.wait_stream(self.lanes.real)
            self.main.wait_stream(self.lanes.fake)
            self.D.merge_lanes()
            if self.grad_sync is not None:
                self.grad_sync(self.D)
            self.opt_d.step()
            cap.out = d_loss.detach()
            del d_loss
            cap.graph.capture_end()
        finally:
            lib.set_rng_base(None)
        self.opt_d._t -= 1  # the captured step() counted on the host too; the replays count below
        cap.calls = models._philox_calls - c0
        return cap

    def _prepare(self, lb, vb) -> "_Prepared":
        """Everything of a step that can be done before its first launch: per-batch constants, the two captures, the static
        buffers.  Nothing here executes on the GPU (captures only record), so the step AFTER the current one can be prepared
        while the current one's critic updates are still running (``__call__(..., next_batch=...)``)."""
        cfg, dev = self.cfg, self.dev
        n, R = vb.num_nodes, cfg.N_CRITIC
        models.prepare_batch(lb, vb, cfg.NUM_CLASSES)
        if self.D._native.bucket is None or next(self.D.parameters()).grad is None:
            self.D._native.bind_grads(models._param_list(self.D))
        P = _Prepared()
        P.lb, P.vb = lb, vb
        P.sample = torch.empty(2, n, cfg.NUM_CLASSES, dtype=torch.float32, device=dev)  # the critic graph's static input
        P.losses = torch.empty(R + 1, dtype=torch.float32, device=dev)
        t0 = models._philox_calls
        P.S = self._capture_sampling(lb, vb, n)
        models._philox_calls = t0 + R * P.S.calls  # S's replays use tickets t0+1 .. t0+R*S.calls; C's start after them
        c0 = models._philox_calls
        P.C = self._capture_critic(lb, vb, P.sample)
        models._philox_calls = c0 + R * P.C.calls  # the tickets an eager run of the R critic updates would have consumed
        return P

    def _run(self, lb, vb, sync_losses, next_batch=None):
        cfg, main, gen, dev = self.cfg, self.main, self.lanes.gen, self.dev
        n, R = vb.num_nodes, cfg.N_CRITIC
        while len(self._alive) > 4:  # graphs older than two steps: wait until their last replay has finished, then drop them
            old = self._alive.popleft()
            old.done.synchronize()
        P, self._prepared = self._prepared, None
        if P is None or P.lb is not lb or P.vb is not vb:  # nothing prepared ahead (or for another batch): prepare now
            P = self._prepare(lb, vb)
        S, C, sample, losses = P.S, P.C, P.sample, P.losses
        taken, gen_out, side = None, None, None
        for k in range(R):
            with torch.cuda.stream(gen):
                if taken is not None:
                    gen.wait_event(taken)  # the previous sample has been copied out of S's output buffer
                self.base_s.fill_(k * S.calls * 2048)
                S.graph.replay()
                ready = gen.record_event()
                if k == R - 1:
                    # the generator update's forward pass (trainer.py:483-485) depends on nothing the critic updates change:
                    # it runs on the sampling stream beside the last critic updates (eager: once per step, needs autograd)
                    z = torch.randn(1, n, cfg.Z_DIM, device=dev)
                    gen_out = self.G(lb, vb, z)
                    # the critic-independent loss terms and their gradients: here, beside the critic updates (step.SideLoss)
                    side = _step.SideLoss(vb, gen_out[0], gen_out[1].unsqueeze(0), cfg)
                    for t in tuple(gen_out) + (side.r_main, side.ce, side.r_void, side.far, side.g_logits, side.g_hard_main,
                                        side.g_hard_void):
                        t.record_stream(main)
                    gen_ready = gen.record_event()
            main.wait_event(ready)
            sample.copy_(S.out)
            taken = main.record_event()
            self.base_c.fill_(k * C.calls * 2048)
            C.graph.replay()
            losses[k].copy_(C.out)
        self.opt_d._t += R
        S.done, C.done = gen.record_event(), main.record_event()
        self._alive.extend((S, C))
        if next_batch is not None:
            # The host is ~5 critic updates ahead of the GPU here.  Capturing the NEXT step's graphs now, before the generator
            # update's tail is enqueued, takes the two captures (~2.4 ms of host time) out of the window between this step's
            # last launch and the next step's first one, where the GPU had only the ~1.5 ms tail left to run
            # (profiles/r02b_summary.md: ~0.9 ms of idle GPU per step).
            self._prepared = self._prepare(*next_batch)
        # generator update, the part that needs the updated critic: D(fake), backward, Adam (trainer.py:486-495)
        main.wait_event(gen_ready)
        logits, hard, soft = gen_out
        hard = hard.unsqueeze(0)
        self.opt_g.zero_grad()
        g_loss = _step.generator_loss(self.D, lb, vb, logits, hard, cfg, side=side)
        with self.D.input_grads_only():  # the critic's own .grad from this backward would be zeroed unread (next zero_grad)
            g_loss.backward()  # the generator's backward pass runs on the stream of its forward (autograd's stream affinity)
        main.wait_stream(gen)
        if self.grad_sync is not None:
            self.grad_sync(self.G)
        losses[R].copy_(g_loss.detach())
        self.opt_g.step()
        self.last_losses = losses
        if sync_losses == "step":
            vals = losses.tolist()
            return vals[:-1], vals[-1], hard.detach()
        return list(losses[:-1].unbind(0)), losses[R], hard.detach()
