"""The callers either side of the Multi-ATGCN path (f3 is ``ops.output_head``), on the device.

f1  ``FusedClipAdam`` + ``fused_train_step``: the loop body of ``TrafficStateExecutor._train_epoch``
    (libcity/executor/traffic_state_executor.py:413-422) with the optimiser half -
    ``clip_grad_norm_(model.parameters(), max_grad_norm)`` (executor:420-421) and ``torch.optim.Adam.step()``
    (executor:146-147) - as two launches over ONE flat bucket (``matgcn_grad_sumsq`` + ``matgcn_adam_clip_step``)
    instead of ~60 per-parameter kernels.  Parameters, gradients and both moments are views into four flat fp32
    buffers (each parameter starts on a 256-byte boundary, which also keeps every weight tensor TMA-aligned).
f2  ``DeviceWindowBank``: ``MTHDataset._generate_input_data`` (mth_dataset.py:112-158) + the per-batch collate
    and upload (libcity/data/utils.py:68-72, batch.py:43-57) replaced by a gather from a series resident in HBM
    (``matgcn_assemble_windows``); the host sends only the B label-start indices of a batch.

No CPU fallback: both need the CUDA library and CUDA tensors and raise ``MatgcnError`` otherwise.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist

from . import _cabi
from ._cabi import MatgcnError

_ALIGN = 64  # floats: every parameter's slot starts on a 256-byte boundary


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class FusedClipAdam(torch.optim.Optimizer):
    """Adam (no amsgrad) with an optional global-norm clip, over flat buffers.  Same arithmetic as
    ``clip_grad_norm_`` followed by ``torch.optim.Adam(params, lr, betas, eps, weight_decay).step()``.

    It IS a ``torch.optim.Optimizer``: ``param_groups[0]`` carries ``lr`` / ``betas`` / ``eps`` / ``weight_decay`` and
    ``step`` reads them from there, so the executor's ``_build_lr_scheduler`` (``MultiStepLR`` with the shipped
    ``lr_decay`` recipe, MultiATGCN.json) attaches to it unchanged, and ``state_dict()`` / ``load_state_dict()`` use
    ``torch.optim.Adam``'s own per-parameter layout (``step``, ``exp_avg``, ``exp_avg_sq``), so the checkpoints the executor
    writes (executor:95, 106, 118, 136) move both ways between this class and a stock ``torch.optim.Adam``.

    After construction ``p.data`` and ``p.grad`` of every trainable parameter are views into ``self.param`` /
    ``self.grad`` (so autograd accumulates straight into the bucket and a data-parallel all-reduce is one call on
    ``self.grad``).  ``zero_grad`` is one memset.  ``step`` verifies that the views are still in place (a later
    ``model.to(...)`` / ``.float()`` would silently detach the model from the bucket) and raises otherwise.

    One difference from ``torch.optim.Adam``: a parameter that received no gradient in a step keeps a zero-filled slot instead
    of ``grad is None``, so with ``weight_decay != 0`` it is decayed where torch would skip it (with the executor's default
    ``weight_decay = 0`` the update of such a slot is exactly zero, as in torch)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-2, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, max_grad_norm: Optional[float] = None):
        plist = [p for p in params if p.requires_grad]
        if not plist:
            raise ValueError("no trainable parameters")
        dev = plist[0].device
        if dev.type != "cuda":
            raise MatgcnError("FusedClipAdam needs CUDA parameters: there is no CPU fallback")
        for p in plist:
            if p.dtype != torch.float32 or p.device != dev:
                raise MatgcnError("FusedClipAdam needs float32 parameters on one device")
        super().__init__(plist, dict(lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps),
                                     weight_decay=float(weight_decay)))
        if len(self.param_groups) != 1:
            raise ValueError("FusedClipAdam keeps one parameter group (one flat bucket)")
        self.params = plist
        self.max_grad_norm = max_grad_norm
        self.offsets = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total = off
        self.param = torch.zeros(off, device=dev)
        self.grad = torch.zeros(off, device=dev)
        self.exp_avg = torch.zeros(off, device=dev)
        self.exp_avg_sq = torch.zeros(off, device=dev)
        self._sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.grad_norm = torch.zeros(1, device=dev)      # total norm of the last step (what clip_grad_norm_ returns)
        self.step_count = 0
        self._step_dev = None   # device-resident step count / learning rate (set by GraphedTrainStep: see enable_device_state)
        self._lr_dev = None
        self._lr_dev_value = None
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                self.param[o:o + p.numel()].copy_(p.data.reshape(-1))
                p.data = self.param[o:o + p.numel()].view_as(p)
        self._pin_grads()
        self._lib = _cabi.lib()

    # hyper-parameters live in param_groups[0] (what lr schedulers and checkpoints touch); these are conveniences
    @property
    def lr(self) -> float:
        return float(self.param_groups[0]["lr"])

    @lr.setter
    def lr(self, v: float):
        self.param_groups[0]["lr"] = float(v)

    # -- gradient bucket -------------------------------------------------------------------------
    def _pin_grads(self):
        es = self.grad.element_size()
        base = self.grad.data_ptr()
        for p, o in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != base + o * es:
                p.grad = self.grad[o:o + p.numel()].view_as(p)

    def _check_param_views(self):
        es = self.param.element_size()
        base = self.param.data_ptr()
        for p, o in zip(self.params, self.offsets):
            if p.data_ptr() != base + o * es:
                raise MatgcnError("FusedClipAdam: a parameter's storage was replaced after the optimiser was built "
                                  "(model.to(...) / .float() / .half()?): the model no longer reads the flat bucket the "
                                  "update writes - rebuild the optimiser after moving the model")

    def zero_grad(self, set_to_none: bool = False):
        """Replaces optimizer.zero_grad() (executor:414): one memset; the views are re-pinned in case a caller
        dropped them."""
        self.grad.zero_()
        self._pin_grads()

    def all_reduce(self, group=None) -> float:
        """Sums the bucket over the data-parallel ranks and returns the factor (1/world) that ``step`` folds into
        the update (each rank's loss is a mean over its shard: SURVEY.md 8e)."""
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(group)
            if world > 1:
                dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
                return 1.0 / world
        return 1.0

    # -- device-resident step state (CUDA-graph replays) -------------------------------------------
    def enable_device_state(self):
        """Moves the two values that change between steps - Adam's step count and the learning rate - into device memory
        (``matgcn_adam_clip_step_dev``), so that a captured step can be replayed: ``matgcn_step_tick`` advances the count on the
        device, ``sync_lr`` pushes a scheduler's new ``param_groups[0]['lr']`` (executor:155-197) when it changed."""
        if self._step_dev is None:
            dev = self.param.device
            self._step_dev = torch.tensor([self.step_count], device=dev, dtype=torch.int64)
            self._lr_dev = torch.tensor([self.lr], device=dev, dtype=torch.float32)
            self._lr_dev_value = self.lr
        return self._step_dev

    def disable_device_state(self):
        if self._step_dev is not None:
            self.step_count = int(self._step_dev.item())
        self._step_dev = self._lr_dev = self._lr_dev_value = None

    def sync_lr(self):
        if self._lr_dev is not None and self.lr != self._lr_dev_value:
            self._lr_dev.fill_(self.lr)
            self._lr_dev_value = self.lr

    # -- update ----------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._check_param_views()
        es, base = self.grad.element_size(), self.grad.data_ptr()
        for p, o in zip(self.params, self.offsets):   # a gradient that was re-created outside the bucket (zero_grad(set_to_none)
            if p.grad is not None and p.grad.data_ptr() != base + o * es:   # by a foreign caller) is moved in, not dropped
                self.grad[o:o + p.numel()].copy_(p.grad.reshape(-1))
                p.grad = self.grad[o:o + p.numel()].view_as(p)
        g = self.param_groups[0]
        st = _stream()
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        _cabi.check(self._lib.matgcn_grad_sumsq(self.grad.data_ptr(), self.total, self._sumsq.data_ptr(), st), "grad_sumsq")
        if self._step_dev is not None:
            # the count was advanced on the device (matgcn_step_tick); the learning rate is read from device memory
            _cabi.check(self._lib.matgcn_adam_clip_step_dev(
                self.param.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.total,
                self._sumsq.data_ptr(), float(self.max_grad_norm) if clip else 0.0, float(grad_scale), self._lr_dev.data_ptr(),
                float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), self._step_dev.data_ptr(), 1,
                self.grad_norm.data_ptr(), st), "adam_clip_step_dev")
            return loss
        self.step_count += 1
        _cabi.check(self._lib.matgcn_adam_clip_step(
            self.param.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.total,
            self._sumsq.data_ptr(), float(self.max_grad_norm) if clip else 0.0, float(grad_scale), float(g["lr"]),
            float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]), self.step_count, 1,
            self.grad_norm.data_ptr(), st), "adam_clip_step")
        return loss

    # -- checkpointing (executor:95, 106, 118, 136 save/load optimizer.state_dict()) ---------------
    def state_dict(self):
        """``torch.optim.Adam``'s layout: state[i] = {step, exp_avg, exp_avg_sq} per parameter + param_groups."""
        for p, o in zip(self.params, self.offsets):
            n = p.numel()
            self.state[p] = {"step": torch.tensor(float(self.step_count)),
                             "exp_avg": self.exp_avg[o:o + n].view_as(p).clone(),
                             "exp_avg_sq": self.exp_avg_sq[o:o + n].view_as(p).clone()}
        sd = super().state_dict()
        self.state.clear()
        return sd

    def load_state_dict(self, sd):
        """Accepts a ``torch.optim.Adam`` (or own) state dict; the moments are copied into the flat buffers."""
        super().load_state_dict(sd)
        steps = set()
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                st = self.state.get(p)
                if not st:
                    continue
                n = p.numel()
                self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise MatgcnError("FusedClipAdam.load_state_dict: parameters carry different step counts (%s); one flat bucket "
                              "has one bias-correction step" % sorted(steps))
        if steps:
            self.step_count = steps.pop()
        self.state.clear()


def fused_train_step(model, batch, opt: FusedClipAdam, micro_batches: int = 1):
    """One iteration of ``TrafficStateExecutor._train_epoch`` (executor:413-422): zero_grad -> calculate_loss ->
    backward -> [data-parallel all-reduce] -> clip_grad_norm_ + Adam (fused).  Returns the device loss tensor.

    ``micro_batches > 1`` runs the forward/backward over that many consecutive slices of the batch, each weighted by its share
    of the samples, and accumulates their gradients in the flat bucket before the single update: the saved activations of only
    one slice are alive at a time (the N = 8192 shape at 64 samples per GPU needs it).  The result equals the one-shot step
    whenever the loss is a mean over samples with equal mask density per slice (the same caveat as batch sharding)."""
    if opt._step_dev is not None:
        raise MatgcnError("fused_train_step: the optimiser's step state lives on the device (a GraphedTrainStep owns it); "
                          "call that object, or its close(), instead")
    return _train_step_body(model, batch, opt, micro_batches)


def _train_step_body(model, batch, opt: FusedClipAdam, micro_batches: int = 1):
    opt.zero_grad()
    if micro_batches <= 1:
        loss = model.calculate_loss(batch)
        loss.backward()
        loss = loss.detach()
    else:
        total = next(iter(batch.values())).shape[0]
        bounds = [total * i // micro_batches for i in range(micro_batches + 1)]
        loss = None
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            if hi == lo:
                continue
            part = model.calculate_loss({k: v[lo:hi] for k, v in batch.items()}) * ((hi - lo) / total)
            part.backward()
            loss = part.detach() if loss is None else loss + part.detach()
    scale = opt.all_reduce()
    opt.step(grad_scale=scale)
    return loss


class GraphedTrainStep:
    """``fused_train_step`` captured ONCE as a CUDA graph and replayed for every step: the loop body of
    ``TrafficStateExecutor._train_epoch`` (executor:413-422) becomes one ``cudaGraphLaunch`` - ~80 launches of this library, the
    ~80 small torch kernels of the view fusion / gradient accumulation, the memsets and copies, with their launch gaps and
    2.6 - 3.2 ms of host enqueue time per step gone.

    What changes from step to step lives in device memory and is advanced by the graph's first node (``matgcn_step_tick``):
    Adam's step count (bias corrections) and the dropout key (a replay draws a new mask, as ``F.dropout`` does, MA.py:416).  The
    learning rate is a device scalar too; ``__call__`` pushes ``param_groups[0]['lr']`` when a scheduler changed it.  The batch
    is copied into static buffers (``self.batch``) before the replay; the returned loss is a static device scalar.

    Data-parallel runs capture zero_grad .. backward and issue the all-reduce and the update eagerly after the replay (three
    more launches) unless ``capture_collective=True``.

    Not capturable (raises): ``add_static`` models - the reference re-draws a randomised ``torch.pca_lowrank`` on the host in
    every forward (MA.py:405-409).  Semantics are those of ``fused_train_step``; tests/test_gpu_train.py compares the two step
    by step."""

    def __init__(self, model, opt: FusedClipAdam, example_batch, micro_batches: int = 1, warmup: int = 2,
                 capture_collective: bool = False, input_slots: int = 1):
        if getattr(model, "static", None) is not None:
            raise MatgcnError("GraphedTrainStep: add_static models draw a host-side randomised PCA in every forward and cannot "
                              "be captured; use fused_train_step")
        if not model.training:
            raise MatgcnError("GraphedTrainStep captures a TRAIN step: call model.train() first")
        self.model, self.opt, self.micro_batches = model, opt, int(micro_batches)
        dev = opt.param.device
        # input_slots > 1: that many sets of static input buffers, one captured graph each (they share one memory pool: the graphs
        # are replayed one after the other, never concurrently), used round-robin - the host batch of step i+1 can then be uploaded
        # on a copy stream straight into the buffers of the NEXT graph while step i runs, with no staging copy on the compute stream
        self.batches = [{k: torch.empty_like(v, device=dev).copy_(v) for k, v in example_batch.items() if torch.is_tensor(v)}
                        for _ in range(max(1, int(input_slots)))]
        self.batch = self.batches[0]
        self._next = 0
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._split = world > 1 and not capture_collective
        self._lib = _cabi.lib()
        # device-resident step state
        self._step_dev = opt.enable_device_state()
        self._key_dev = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).to(dev)
        model._dropout_key_dev = self._key_dev
        model._dropout_key_host = int(torch.randint(0, 2 ** 62, (1,)).item())
        # warm-up on a side stream (lazy initialisation inside torch / the library), then put the training state back
        keep = [t.clone() for t in (opt.param, opt.exp_avg, opt.exp_avg_sq, self._step_dev, self._key_dev)]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._body(eager=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        with torch.no_grad():
            for t, k in zip((opt.param, opt.exp_avg, opt.exp_avg_sq, self._step_dev, self._key_dev), keep):
                t.copy_(k)
        torch.cuda.synchronize(dev)
        n0 = self._lib.matgcn_launch_count()
        self.graphs, self.losses = [], []
        for s in range(len(self.batches)):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, **({"pool": self.graphs[0].pool()} if s else {})):
                loss = self._body(eager=False, slot=s)
            self.graphs.append(g)
            self.losses.append(loss)
        self.graph, self.loss = self.graphs[0], self.losses[0]
        self.library_kernel_nodes = int(self._lib.matgcn_launch_count() - n0) // len(self.graphs)   # this library's kernels per replay

    def _body(self, eager: bool, slot: int = 0):
        opt = self.opt
        _cabi.check(self._lib.matgcn_step_tick(self._key_dev.data_ptr(), self._step_dev.data_ptr(), _stream()), "step_tick")
        if self._split and not eager:
            # data parallel: the graph ends after the backward; all-reduce + update follow eagerly in __call__
            opt.zero_grad()
            return self._fwd_bwd(slot)
        loss = _train_step_body(self.model, self.batches[slot], opt, self.micro_batches)
        return loss

    def _fwd_bwd(self, slot: int = 0):
        model, batch, mb = self.model, self.batches[slot], self.micro_batches
        if mb <= 1:
            loss = model.calculate_loss(batch)
            loss.backward()
            return loss.detach()
        total = next(iter(batch.values())).shape[0]
        bounds = [total * i // mb for i in range(mb + 1)]
        loss = None
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            if hi == lo:
                continue
            part = model.calculate_loss({k: v[lo:hi] for k, v in batch.items()}) * ((hi - lo) / total)
            part.backward()
            loss = part.detach() if loss is None else loss + part.detach()
        return loss

    def next_slot(self) -> int:
        """The input slot the next ``__call__`` without an explicit slot will read."""
        return self._next

    def load_batch(self, batch, non_blocking: bool = True, slot: int = 0):
        """Copies a batch (host or device tensors of the captured shapes) into the static input buffers of ``slot``, on the current
        stream.  With ``input_slots > 1`` a caller may do this on a copy stream for the NEXT slot while the current replay runs, as long
        as that stream has waited for the replay that last read the slot (bench.py's e2e loop shows the event pattern)."""
        for k, dst in self.batches[slot].items():
            src = batch[k]
            if src.shape != dst.shape:
                raise MatgcnError("GraphedTrainStep: batch['%s'] has shape %s, the captured step has %s (capture another one for "
                                  "a ragged last batch)" % (k, tuple(src.shape), tuple(dst.shape)))
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=non_blocking)

    def __call__(self, batch=None, slot=None):
        """One train step from input slot ``slot`` (default: round-robin); ``batch=None`` reuses whatever the slot's buffers hold.
        Returns the (static, per slot) device loss tensor."""
        if slot is None:
            slot = self._next
        if batch is not None:
            self.load_batch(batch, slot=slot)
        self.opt.sync_lr()
        self.graphs[slot].replay()
        self._next = (slot + 1) % len(self.graphs)
        if self._split:
            scale = self.opt.all_reduce()
            self.opt.step(grad_scale=scale)
        self.opt.step_count += 1   # host mirror (checkpoints: state_dict reads it)
        return self.losses[slot]

    def close(self):
        """Hands the step state back to the host side of the optimiser (eager ``fused_train_step`` works again)."""
        self.opt.disable_device_state()
        self.model._dropout_key_dev = None
        self.graph = None
        self.graphs = []


class DeviceWindowBank:
    """The series ``[T_total, N, F]`` stays in HBM; ``assemble(label_starts)`` gathers a batch
    ``X [B, (len_c+len_p+len_t)*input_window, N, F]``, ``y [B, output_window, N, F]`` exactly as
    ``MTHDataset._generate_input_data`` (mth_dataset.py:112-158) lays samples out: closeness, period, trend
    segments, each oldest to newest (mth_dataset.py:59, 140-153), target = series[s : s + output_window] (:105).

    The reference scales windows after cutting them (an elementwise affine map with global statistics), so handing
    this class the already scaled series is equivalent."""

    def __init__(self, series: torch.Tensor, input_window: int, output_window: int, len_closeness: int, len_period: int,
                 len_trend: int, interval_period: int = 1, interval_trend: int = 7, points_per_hour: int = 1,
                 hour_each_day: int = 24):
        if series.device.type != "cuda":
            raise MatgcnError("DeviceWindowBank needs a CUDA series: there is no CPU fallback")
        if series.dim() != 3 or series.dtype != torch.float32:
            raise ValueError("series must be float32 [T_total, N, F]")
        assert len_closeness + len_period + len_trend > 0        # mth_dataset.py:16
        self.series = series.contiguous()
        self.T_total, self.N, self.F = self.series.shape
        self.input_window, self.output_window = int(input_window), int(output_window)
        offs = []
        # the reference's own expressions (mth_dataset.py:50, 82-83, 90-91, 98-99), segments oldest first
        for i in range(len_closeness, 0, -1):
            offs.append(int(points_per_hour * (input_window / points_per_hour) * i))
        for i in range(len_period, 0, -1):
            offs.append(int(points_per_hour * (interval_period * hour_each_day) * i))
        for i in range(len_trend, 0, -1):
            offs.append(int(points_per_hour * (interval_trend * hour_each_day) * i))
        self.seg_offsets_host = offs
        self.n_seg = len(offs)
        self.seg_offsets = torch.tensor(offs, device=series.device, dtype=torch.int32)
        self._bad = torch.zeros(1, device=series.device, dtype=torch.int32)
        self._lib = _cabi.lib()

    def valid_label_starts(self) -> torch.Tensor:
        """Label starts the reference would emit a sample for (mth_dataset.py:45-46, 52-58, 78-79), in order."""
        lo = max(self.seg_offsets_host)
        hi = self.T_total - max(self.input_window, self.output_window)
        return torch.arange(lo, hi + 1, dtype=torch.int64) if hi >= lo else torch.empty(0, dtype=torch.int64)

    def assemble(self, label_starts: torch.Tensor, out_x: Optional[torch.Tensor] = None, out_y: Optional[torch.Tensor] = None,
                 check: bool = False):
        """label_starts: int64 [B] (host or device; a host tensor is the only per-batch upload).  Returns {'X','y'}."""
        dev = self.series.device
        ls = label_starts.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
        B = ls.numel()
        X = out_x if out_x is not None else torch.empty(B, self.n_seg * self.input_window, self.N, self.F, device=dev)
        y = out_y if out_y is not None else torch.empty(B, self.output_window, self.N, self.F, device=dev)
        if check:
            self._bad.zero_()
        _cabi.check(self._lib.matgcn_assemble_windows(self.series.data_ptr(), self.T_total, self.N, self.F,
                                                      self.seg_offsets.data_ptr(), self.n_seg, self.input_window,
                                                      self.output_window, ls.data_ptr(), B, X.data_ptr(), y.data_ptr(),
                                                      self._bad.data_ptr(), _stream()), "assemble_windows")
        if check and int(self._bad.item()):
            raise MatgcnError("assemble_windows: a label start is not a valid sample under mth_dataset.py:45-58, 78-79")
        return {"X": X, "y": y}
