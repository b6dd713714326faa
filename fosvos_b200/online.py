"""One-shot online fine-tuning and per-sequence inference on one GPU.

Mirrors the hot loops of the reference drivers, with the same argument meaning:
  ``finetune`` / ``OnlineTrainer`` <- ``train_online._train`` loop body (``src/train_online.py:75-101``)
  ``infer_sequence``               <- ``util/experiment_helper.test`` (``src/util/experiment_helper.py:34-64``)
  ``sequences_for_rank``           <- the ``--sequence-group`` round-robin (``src/train_online.py:184-186``)

Differences that do not change results: the frame/mask stay resident on the device (the
reference re-uploads the same first frame every iteration, ``train_online.py:77``), the loss
scalar stays on the device (the reference syncs every iteration, ``:82``), the optimizer step
and ``zero_grad`` are one launch, and the micro-iteration (forward, loss, backward) and the
optimizer step (+ weight re-packing) are each replayed from a CUDA graph that is captured once per
frame size and reused for every sequence the process handles.
"""
from __future__ import annotations

import functools
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import ops
from .networks import OSVOS_VGG, _act_dtype
from .optim import FusedSGD, get_optimizer_online
from .sharding import GradBuckets, allreduce_flat


def _on_device(fn):
    """Run a trainer method with the trainer's GPU current (streams, graphs and launches belong to it)."""
    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)
    return wrapper


MAX_RESIDENT_TRAINERS = 6        # 3 train-time scales x {plain, flipped} share a size: 6 covers every shape of SURVEY section 7 step 5


def _repack_in_place(net: OSVOS_VGG, weights: bool = True) -> None:
    """Refresh the packed weight copies INTO their existing buffers (their addresses are baked into
    captured graphs) after the parameters changed.  Tensor-core mode: one launch for all layers.
    ``weights=False``: the conv weights were re-packed by the fused optimizer step; only the side-chain block is rebuilt."""
    dt = _act_dtype(net.precision)
    tc = net._impl() == "tc"
    convs = [c for st in net._stage_convs() for c in st] + list(net.side_prep)
    if not weights:
        pass
    elif tc:
        ent = []
        for conv in convs:
            pc = net._packed.get(id(conv))
            if pc is None or pc.w_fwd is None:
                continue
            b = None if conv.bias is None else conv.bias.detach()
            ent.append((conv.weight.detach(), b, pc.w_fwd, pc.w_dgrad, pc.bias))
        if ent:
            key = tuple((w.data_ptr(), 0 if b is None else b.data_ptr(), f.data_ptr(), 0 if d is None else d.data_ptr(), bo.data_ptr())
                        for w, b, f, d, bo in ent)
            cached = net.__dict__.get("_repack_table")
            if cached is None or cached[0] != key:
                cached = (key, ops.repack_table(ent, ent[0][0].device))
                net.__dict__["_repack_table"] = cached
            ops.repack_all(cached[1])
    else:
        for conv in convs:
            pc = net._packed.get(id(conv))
            if pc is None:
                continue
            b = conv.bias
            if pc.w_fwd is not None:
                ops.pack_weight(conv.weight, L.W_SIMT_FWD, dt, out=pc.w_fwd)
                ops.pad_bias(b, conv.out_channels, conv.weight.device, out=pc.bias)
            if pc.w_dgrad is not None:
                ops.pack_weight(conv.weight, L.W_SIMT_DGRAD, dt, out=pc.w_dgrad)
    if net._side_params is not None:
        ops.side_params_prepare([m.weight for m in net.upscale], [m.weight for m in net.upscale_],
                                [m.weight for m in net.score_dsn], [m.bias for m in net.score_dsn],
                                net.fuse.weight, net.fuse.bias, out=net._side_params)
    _sync_cache_keys(net)


def _sync_cache_keys(net: OSVOS_VGG) -> None:
    """Mark the packed copies as current for the parameters' present versions."""
    convs = [c for st in net._stage_convs() for c in st] + list(net.side_prep)
    for conv in convs:
        pc = net._packed.get(id(conv))
        if pc is not None:
            b = conv.bias
            pc.key = (conv.weight.data_ptr(), conv.weight._version, None if b is None else (b.data_ptr(), b._version),
                      net.precision, tuple(conv.weight.shape))
    if net._side_params is not None:
        net._side_key = net._side_cache_key()


def _keys_current(net: OSVOS_VGG) -> bool:
    convs = [c for st in net._stage_convs() for c in st] + list(net.side_prep)
    for conv in convs:
        pc = net._packed.get(id(conv))
        if pc is not None and pc.key is not None:
            b = conv.bias
            if pc.key[1] != conv.weight._version or (b is not None and pc.key[2] != (b.data_ptr(), b._version)):
                return False
    return net._side_params is None or net._side_key == net._side_cache_key()


class OnlineTrainer:
    """The fine-tune loop of ``train_online._train`` for one frame size, reusable across sequences.

    ``micro`` = fwd -> balanced loss on the fused map (size_average=False) -> / avg_grad_every_n ->
    bwd accumulating into ``p.grad``;  every ``avg_grad_every_n`` micro-iterations:
    ``optimizer.step(); optimizer.zero_grad()`` (one fused launch) and re-packing of the kernel-side
    weight copies.  ``deep_supervision=w`` adds the four side-map losses with weight ``w``
    (``train_offline.py:84-88``: ``w = 1 - epoch/n_epochs``).
    """

    def __init__(self, net: OSVOS_VGG, height: int, width: int, avg_grad_every_n: int = 5,
                 optimizer: Optional[FusedSGD] = None, use_graph: bool = True, deep_supervision: Optional[float] = None,
                 data_parallel: bool = False, world_size: int = 1, fuse_window: bool = True,
                 share_with: Optional["OnlineTrainer"] = None, overlap_allreduce: bool = False):
        self.device = next(net.parameters()).device
        with torch.cuda.device(self.device):
            self._init(net, height, width, avg_grad_every_n, optimizer, use_graph, deep_supervision, data_parallel, world_size,
                       fuse_window, share_with, overlap_allreduce)

    def _init(self, net, height, width, avg_grad_every_n, optimizer, use_graph, deep_supervision, data_parallel, world_size,
              fuse_window, share_with, overlap_allreduce):
        """``fuse_window``: the ``avg_grad_every_n`` micro-iterations between two optimizer steps all see the SAME weights
        and their gradients are summed, so they are independent of each other: run them as ONE batched
        forward/backward over the window's frames (every frame is still computed in full, each with its own loss and
        label statistics; the accumulated gradient is the same sum in a different fp32 order).  At batch 1 a third of
        every layer's time is launch, prologue, pipeline fill and tail waves -- the window amortises them.
        ``data_parallel``: offline parent training sharded over ``world_size`` ranks (one process per GPU): every
        rank runs ``avg_grad_every_n // world_size`` micro-iterations on its own frames, gradients (scaled by
        1/avg_grad_every_n as in the reference) are summed with ONE all-reduce of a flat fp32 buffer, then every
        rank applies the same optimizer step.  With ``overlap_allreduce`` (off by default: measured slower at 2 GPUs -- the NCCL
        kernels take SMs from the persistent 148-CTA conv grids) the flat buffer is laid out in per-stage buckets
        (deepest stage first, the order the backward pass finishes them) and every bucket's weight-gradient fold +
        all-reduce is issued on a side stream as soon as that stage's weight gradients are done, so the exchange hides
        behind the rest of the backward pass (needs the fused window, runs it outside CUDA graphs).
        ``share_with``: another trainer of the same network (a different frame size): gradients, weight-gradient
        accumulators, optimizer (momentum), counters and the optimizer-step graph are SHARED, so consecutive iterations may
        alternate between frame sizes (train-time Resize scales, custom_transforms.py:63-109) without re-capturing."""
        dev = self.device
        L.require_device(dev)
        self.net = net
        self.data_parallel = bool(data_parallel) and world_size > 1
        if self.data_parallel and avg_grad_every_n % world_size != 0:
            raise ValueError(f"avg_grad_every_n={avg_grad_every_n} must be a multiple of the world size {world_size}")
        self.n = int(avg_grad_every_n) // (world_size if self.data_parallel else 1)      # micro-iterations per step on this rank
        self.scale = 1.0 / float(avg_grad_every_n)
        self.deep = deep_supervision
        # the side-loss weight (1 - epoch / n_epochs, train_offline.py:88) lives on the device: captured graphs read it
        self.deep_w = None if deep_supervision is None else torch.full((), float(deep_supervision), dtype=torch.float32, device=dev)
        sh = share_with._shared if share_with is not None else None
        if sh is not None and optimizer is not None and optimizer is not sh["optimizer"]:
            raise ValueError("share_with: the shared trainer already owns the optimizer")
        self.optimizer = sh["optimizer"] if sh is not None else (optimizer if optimizer is not None else get_optimizer_online(net))
        self.fused = isinstance(self.optimizer, FusedSGD)
        self.overlap_ar = self.data_parallel and bool(overlap_allreduce) and self.fused and net._impl() == "tc" and \
            bool(fuse_window) and self.n > 1 and sh is None
        self.use_graph = bool(use_graph) and self.fused
        self.fuse = bool(fuse_window) and self.n > 1
        slots = self.n if self.fuse else 1
        # one resident slot per micro-iteration of a window; slot 0 doubles as the single-iteration buffers
        self.frames = torch.zeros((slots, net.stages[0][0].in_channels, height, width), dtype=torch.float32, device=dev)
        self.masks = torch.zeros((slots, 1, height, width), dtype=torch.float32, device=dev)
        self.frame, self.mask = self.frames[0:1], self.masks[0:1]
        self.window_losses = torch.zeros(slots, dtype=torch.float32, device=dev)
        self._part_losses = torch.zeros(slots, dtype=torch.float32, device=dev)
        params = dict(net.named_parameters())
        if sh is not None:
            self._shared = sh
        else:
            grads: Dict[str, torch.Tensor] = {}
            flat_grad, buckets = None, None
            if self.data_parallel:
                # one flat fp32 buffer, bucket-major: [stage 4 + side_prep.3][stage 3 + side_prep.2] ... [stage 0][heads]
                buckets = GradBuckets(net, params, dev)
                flat_grad = buckets.flat
            for name in net._grad_names():
                p = params[name]
                if buckets is not None:
                    p.grad = buckets.view(name)
                elif p.grad is None:
                    p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
                grads[name] = p.grad
            # tensor-core weight gradients accumulate in their own [tap][M][N] layout over the micro-iterations and are
            # folded into .grad once per optimizer step
            wgrad_ws: Optional[Dict[str, torch.Tensor]] = None
            if self.fused and net._impl() == "tc":
                wgrad_ws = {}
                for name in grads:
                    if name.endswith(".weight") and params[name].dim() == 4 and tuple(params[name].shape[2:]) == (3, 3):
                        cout, cin = params[name].shape[0], params[name].shape[1]
                        wgrad_ws[name[:-len(".weight")]] = ops.wgrad_workspace(ops.pad8(cin), ops.pad8(cout), dev)
            # single-GPU tensor-core fine-tune: fold + SGD + repack of the 3x3 conv weights are ONE fused launch per step
            # (step.cu: conv_step_kernel); the optimizer keeps the biases and the heads
            import os as _os
            conv_step = self.fused and wgrad_ws is not None and not self.data_parallel and _os.environ.get("FOSVOS_FUSED_STEP", "1") != "0"
            if conv_step:
                self.optimizer.exclude([params[name + ".weight"] for name in wgrad_ws])
            self._shared = dict(optimizer=self.optimizer, grads=grads, flat_grad=flat_grad, buckets=buckets, wgrad_ws=wgrad_ws,
                                conv_step=conv_step, conv_table=None, conv_key=None,
                                fold_table=None, step_graph=None, calls_step=0, step_gen=-1, counter=0,
                                loss_sum=torch.zeros((), dtype=torch.float32, device=dev),
                                last_loss=torch.zeros((), dtype=torch.float32, device=dev))
        self.grads = self._shared["grads"]
        self.flat_grad = self._shared["flat_grad"]
        self.wgrad_ws = self._shared["wgrad_ws"]
        # label counts of the resident mask: they only change with set_frame(), so loss and gradient are ONE pass
        self.loss_stats_all = torch.zeros((slots, L.lib().fosvos_bal_loss_stats_bytes() // 8), dtype=torch.float64, device=dev)
        self.loss_stats = self.loss_stats_all[0]
        self._label_counts()
        self.loss_sum = self._shared["loss_sum"]
        self.last_loss = self._shared["last_loss"]
        self._micro_total = torch.zeros(1, dtype=torch.float32, device=dev)
        self._micro_graph = None
        self._window_graph = None
        self._calls_micro = 0
        self._calls_window = 0
        self._comm_stream: Optional[torch.cuda.Stream] = None
        self._overlap_graph = None
        self.overlap_capture_error: Optional[str] = None

    # shared between the trainers of one network (see ``share_with``)
    @property
    def counter(self) -> int:
        return self._shared["counter"]

    @counter.setter
    def counter(self, v: int) -> None:
        self._shared["counter"] = v

    @property
    def _fold_table(self):
        return self._shared["fold_table"]

    @_fold_table.setter
    def _fold_table(self, v) -> None:
        self._shared["fold_table"] = v

    @property
    def _step_graph(self):
        return self._shared["step_graph"]

    @_step_graph.setter
    def _step_graph(self, v) -> None:
        self._shared["step_graph"] = v

    @property
    def _calls_step(self) -> int:
        return self._shared["calls_step"]

    @_calls_step.setter
    def _calls_step(self, v: int) -> None:
        self._shared["calls_step"] = v

    # ---------------------------------------------------------------- pieces
    def _label_counts(self) -> None:
        for i in range(self.masks.shape[0]):
            ops.bal_loss_fwd(self.masks[i:i + 1], self.masks[i:i + 1], False, stats=self.loss_stats_all[i])

    def _window(self) -> None:
        """All micro-iterations of one accumulation window as one batched pass (see ``fuse_window``)."""
        net = self.net
        n = self.frames.shape[0]
        outs, _, _, saved = net._run_forward(self.frames, save=True)
        douts: List[Optional[torch.Tensor]] = [None] * 5
        # the class-balanced loss normalises by the label counts of ITS frame (osvos_layers.py:26-39): one launch per
        # output map computes every frame's loss and gradient with that frame's statistics
        _, douts[4] = ops.bal_loss_fwd_bwd_frames(outs[4], self.masks, False, self.loss_stats_all, None, self.scale,
                                                  losses=self.window_losses)
        if self.deep_w is not None:
            for j in range(4):
                _, douts[j] = ops.bal_loss_fwd_bwd_frames(outs[j], self.masks, False, self.loss_stats_all, self.deep_w, self.scale,
                                                          losses=self._part_losses)
                ops.loss_accumulate(self._part_losses, self.window_losses, self.deep_w)
        ops.loss_window_finish(self.window_losses, self.loss_sum, self.last_loss)
        net._run_backward(saved, douts, self.grads, self.wgrad_ws, **self._backward_hooks())

    def _micro(self) -> None:
        net = self.net
        outs, _, _, saved = net._run_forward(self.frame, save=True)
        douts: List[Optional[torch.Tensor]] = [None] * 5
        loss, douts[4] = ops.bal_loss_fwd_bwd(outs[4], self.mask, False, self.loss_stats, None, self.scale)
        ops.loss_accumulate(loss.view(1), self._micro_total, None, init=True)
        if self.deep_w is not None:
            for i in range(4):
                li, douts[i] = ops.bal_loss_fwd_bwd(outs[i], self.mask, False, self.loss_stats, self.deep_w, self.scale)
                ops.loss_accumulate(li.view(1), self._micro_total, self.deep_w)
        ops.loss_window_finish(self._micro_total, self.loss_sum, self.last_loss)
        net._run_backward(saved, douts, self.grads, self.wgrad_ws)

    def _backward_hooks(self) -> dict:
        """Data-parallel overlap: a callback the backward pass fires when a stage's weight gradients have been issued
        (``OSVOS_VGG._run_backward(stage_done=...)``): fold that bucket's accumulators and all-reduce it on the side stream."""
        if not self.overlap_ar:
            return {}
        return {"stage_done": self._bucket_ready}

    def _bucket_ready(self, bucket: int, aux_stream: Optional[torch.cuda.Stream]) -> None:
        bk = self._shared["buckets"]
        comm = self._comm_stream
        main = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(aux_stream if aux_stream is not None else main)        # the bucket's weight gradients run there
        comm.wait_event(ev)
        if bucket == bk.n_buckets - 1:
            ev2 = torch.cuda.Event()
            ev2.record(main)                                             # the heads' gradients (side_bwd) run on the main stream
            comm.wait_event(ev2)
        with torch.cuda.stream(comm):
            tab = bk.fold_table(bucket, self.wgrad_ws, self.grads)
            if tab is not None:
                ops.wgrad_fold_all(tab)
            self._ar_works.append(bk.allreduce(bucket))

    def _ensure_conv_table(self) -> None:
        """Device table of the fused conv-weight step; rewritten IN PLACE when a learning rate, weight decay or buffer changed
        (captured step graphs read it at replay)."""
        sh = self._shared
        if sh.get("conv_static") is None:
            params = dict(self.net.named_parameters())
            modules = dict(self.net.named_modules())
            groups = {id(q): g for g in self.optimizer.param_groups for q in g["params"]}
            sh["conv_static"] = [(ws, modules[name], params[name + ".weight"], groups[id(params[name + ".weight"])])
                                 for name, ws in self.wgrad_ws.items()]
        ent = []
        for ws, conv, w, grp in sh["conv_static"]:
            pc = self.net._packed_for(conv, need_dgrad=True)
            ent.append((ws, None, w.detach(), self.optimizer.momentum_buffer(w), None if conv.bias is None else conv.bias.detach(),
                        pc.w_fwd, pc.w_dgrad, pc.bias, float(grp["lr"]), float(grp["weight_decay"])))
        key = tuple((e[0].data_ptr(), e[2].data_ptr(), e[3].data_ptr(), 0 if e[4] is None else e[4].data_ptr(), e[5].data_ptr(),
                     0 if e[6] is None else e[6].data_ptr(), e[7].data_ptr(), e[8], e[9]) for e in ent)
        if key != sh["conv_key"]:
            old = sh["conv_table"]
            sh["conv_table"] = ops.convstep_table(ent, self.device, out=old)
            if old is not None and sh["conv_table"] is not old:
                sh["step_gen"] = -1                      # a new table address: the captured step must be re-captured
            sh["conv_key"] = key

    def _fold_wgrads(self) -> None:
        if self._shared["conv_step"]:
            return                                       # the fused step consumes the accumulators directly
        if self.wgrad_ws:
            if self._fold_table is None:
                self._fold_table = ops.fold_table([(ws, self.grads[name + ".weight"]) for name, ws in self.wgrad_ws.items()],
                                                  self.frame.device)
            ops.wgrad_fold_all(self._fold_table)

    @_on_device
    def set_deep_supervision(self, weight: float) -> None:
        """New epoch: side-loss weight ``1 - epoch / n_epochs`` (train_offline.py:88); no re-capture needed."""
        self.deep_w.fill_(float(weight))

    def _step(self) -> None:
        """optimizer.step(); optimizer.zero_grad() (+ re-packing).  The caller has folded the weight-gradient
        accumulators (and all-reduced the gradients) before."""
        if self.fused and self._shared["conv_step"]:
            self.optimizer.step_and_zero()               # biases and heads (the conv weights are excluded)
            if self._shared["conv_table"] is None:
                self._ensure_conv_table()
            ops.conv_step_all(self._shared["conv_table"], self.optimizer._momentum)
            for _, _, w, _ in self._shared["conv_static"]:
                torch.autograd.graph.increment_version(w)
            _repack_in_place(self.net, weights=False)
        elif self.fused:
            self.optimizer.step_and_zero()
            _repack_in_place(self.net)
        else:
            self.optimizer.step()
            for g in self.grads.values():
                g.zero_()

    def _capture(self, window: bool) -> None:
        """Warm up once (packs weights, sizes the allocator), then capture the graph of one micro-iteration or of one
        fused window, and (first time) the optimizer-step graph."""
        import os
        assert self.counter == 0, "graphs are captured at a window boundary (the warm-up pass clears the gradient accumulators)"
        body = self._window if window else self._micro
        hooks, self.overlap_ar = self.overlap_ar, False        # (the overlapped data-parallel window is never captured)
        try:
            self._capture_body(body, window)
        finally:
            self.overlap_ar = hooks

    def _capture_body(self, body, window: bool) -> None:
        import os
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            c0 = L.CALLS[0]
            saved_sum = self.loss_sum.clone()
            body()
            calls = L.CALLS[0] - c0
            for g in self.grads.values():
                g.zero_()
            for ws in (self.wgrad_ws or {}).values():
                ws.zero_()
            self.loss_sum.copy_(saved_sum)
            self.optimizer._ensure_table()          # momentum buffers + device table exist before capture
            if self.wgrad_ws and self._fold_table is None:
                self._fold_table = ops.fold_table([(ws, self.grads[name + ".weight"]) for name, ws in self.wgrad_ws.items()],
                                                  self.frame.device)
            _repack_in_place(self.net)              # builds the multi-tensor repack table (host -> device copies)
            if self._shared["conv_step"]:
                self._ensure_conv_table()
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        pool = next((g.pool() for g in (self._micro_graph, self._window_graph, self._step_graph) if g is not None), None)   # one pool per network
        # captured on a high-priority stream: the dependency chain (forward, data gradients) wins the SMs over the
        # weight gradients / side branches that the network issues on its default-priority auxiliary stream
        hp = torch.cuda.Stream(priority=-1 if os.environ.get("FOSVOS_HP", "1") != "0" else 0)
        hp.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(graph, stream=hp, **({} if pool is None else {"pool": pool})):
            body()
        torch.cuda.current_stream().wait_stream(hp)
        if window:
            self._window_graph, self._calls_window = graph, calls
        else:
            self._micro_graph, self._calls_micro = graph, calls
        if self._step_graph is None:
            self._capture_step(graph.pool())
        # capture does not execute: gradients are still zero, parameters untouched; versions were bumped
        _sync_cache_keys(self.net)

    def _capture_step(self, pool=None) -> None:
        self.optimizer._ensure_table()
        if self._shared["conv_step"]:
            self._ensure_conv_table()
        c0 = L.CALLS[0]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, **({} if pool is None else {"pool": pool})):
            self._step()
        self._step_graph = g
        self._calls_step = L.CALLS[0] - c0
        self._shared["step_gen"] = self.optimizer._table_gen
        _sync_cache_keys(self.net)

    # ---------------------------------------------------------------- public
    @_on_device
    def prepare(self, window: Optional[bool] = None) -> None:
        """Capture this trainer's CUDA graphs now (at a window boundary) instead of on first use."""
        if not self.use_graph:
            return
        window = self.fuse if window is None else window
        if window and self._window_graph is None and not self.overlap_ar:
            self._capture(window=True)
        if not window and self._micro_graph is None:
            self._capture(window=False)

    @_on_device
    def reset(self, state_dict: Optional[dict] = None) -> None:
        """Start a new sequence: (optionally) reload the parent weights in place
        (``NetworkProvider.load_model``, network_provider.py:53-58), clear gradients and momentum."""
        if state_dict is not None:
            self.net.load_state_dict(state_dict)
        for g in self.grads.values():
            g.zero_()
        for ws in (self.wgrad_ws or {}).values():
            ws.zero_()
        for st in self.optimizer.state.values():
            buf = st.get("momentum_buffer")
            if buf is not None:
                buf.zero_()
        self.loss_sum.zero_()
        self.counter = 0
        if self.net._packed:
            _repack_in_place(self.net)
        if state_dict is not None and self.net._side_params is not None:
            self.net._side_key = None          # new up-sampling weights: re-run the structure check (one sync,
            self.net._side()                   # outside any graph) and rebuild the parameter block in place

    @_on_device
    def set_frame(self, frame: torch.Tensor, mask: torch.Tensor) -> None:
        """Copy the annotated frame (1,3,H,W) and its mask (1,1,H,W) into the resident buffers -- into every slot of
        the window: each micro-iteration reads its own copy (host tensors are uploaded; pinned ones asynchronously)."""
        self.frame.copy_(frame.reshape(self.frame.shape), non_blocking=True)
        self.mask.copy_(mask.reshape(self.mask.shape), non_blocking=True)
        for i in range(1, self.frames.shape[0]):
            self.frames[i:i + 1].copy_(self.frame)
            self.masks[i:i + 1].copy_(self.mask)
        self._label_counts()                                                      # label counts of the new mask(s)

    @_on_device
    def set_frames(self, frames: torch.Tensor, masks: torch.Tensor) -> None:
        """Distinct frames for the micro-iterations of a window (e.g. flipped copies): (n,3,H,W) and (n,1,H,W)."""
        self.frames.copy_(frames.reshape(self.frames.shape), non_blocking=True)
        self.masks.copy_(masks.reshape(self.masks.shape), non_blocking=True)
        self._label_counts()

    def _optimizer_step(self, exchanged: bool = False) -> None:
        if not exchanged:
            self._fold_wgrads()
            if self.data_parallel:
                allreduce_flat(self.flat_grad)
        if self.use_graph:
            # param_groups (lr, weight decay) or momentum buffers changed since the capture?  The device table is rewritten
            # in place; only a new geometry / momentum needs a fresh capture of the step
            self.optimizer._ensure_table()
            if self._shared["conv_step"]:
                self._ensure_conv_table()
            if self._step_graph is None or self._shared["step_gen"] != self.optimizer._table_gen:
                self._capture_step(next((g.pool() for g in (self._micro_graph, self._window_graph) if g is not None), None))
            self._step_graph.replay()
            L.CALLS[0] += self._calls_step
            for p in self.net.parameters():
                torch.autograd.graph.increment_version(p)
            _sync_cache_keys(self.net)
        else:
            self._step()

    def _window_overlapped(self) -> None:
        """Data-parallel window with the exchange overlapped: eager launches (the NCCL calls sit between them), per-stage
        buckets folded and all-reduced on a side stream while the backward pass continues."""
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        self._ar_works = []
        self._window()
        main = torch.cuda.current_stream(self.device)
        for w in self._ar_works:
            if w is not None:
                w.wait()                                # the main stream waits for the reduced bucket
        main.wait_stream(self._comm_stream)
        self._ar_works = []

    @_on_device
    def run(self, n_iters: int, losses_out: Optional[list] = None) -> torch.Tensor:
        if self.use_graph and (self._micro_graph is not None or self._window_graph is not None) and not _keys_current(self.net):
            _repack_in_place(self.net)          # parameters were changed from outside since the last replay
        done = 0
        while done < n_iters:
            if self.fuse and self.counter == 0 and n_iters - done >= self.n:
                # a whole accumulation window: one batched pass
                if self.overlap_ar:
                    if self.use_graph and self._step_graph is None:
                        self._capture_overlap_warmup()
                    if self._overlap_graph is not None:
                        self._overlap_graph.replay()            # the window with its NCCL calls inside, one graph launch
                        L.CALLS[0] += self._calls_window
                    else:
                        self._window_overlapped()
                    if losses_out is not None:
                        losses_out.extend(float(v) for v in self.window_losses.tolist())
                    done += self.n
                    self._optimizer_step(exchanged=True)
                    continue
                if self.use_graph:
                    if self._window_graph is None:
                        self._capture(window=True)
                    self._window_graph.replay()
                    L.CALLS[0] += self._calls_window
                else:
                    self._window()
                if losses_out is not None:
                    losses_out.extend(float(v) for v in self.window_losses.tolist())
                done += self.n
                self._optimizer_step()
                continue
            if self.use_graph:
                if self._micro_graph is None:
                    self._capture(window=False)
                self._micro_graph.replay()
                L.CALLS[0] += self._calls_micro
            else:
                self._micro()
            if losses_out is not None:
                losses_out.append(float(self.last_loss.item()))
            done += 1
            self.counter += 1
            if self.counter % self.n == 0:
                self._optimizer_step()
                self.counter = 0
        return self.loss_sum


def _capture_overlap_warmup(self: OnlineTrainer) -> None:
    """First overlapped window of a data-parallel trainer: warm up once (packs weights, builds the tables), clear what the
    warm-up accumulated, capture the optimizer-step graph."""
    saved_sum = self.loss_sum.clone()
    self._ar_works = []
    hooks, self.overlap_ar = self.overlap_ar, False
    try:
        self._window()
    finally:
        self.overlap_ar = hooks
    torch.cuda.current_stream(self.device).synchronize()
    for g in self.grads.values():
        g.zero_()
    for ws in (self.wgrad_ws or {}).values():
        ws.zero_()
    self.loss_sum.copy_(saved_sum)
    self.optimizer._ensure_table()
    _repack_in_place(self.net)
    self._capture_step(None)
    # One eager overlapped window (builds the bucket fold tables, warms the NCCL channels), cleared again; then the same
    # window -- forks to the auxiliary and the communication stream, per-bucket folds and NCCL all-reduces included -- is
    # captured into ONE graph, so the exchange overlaps the backward pass without ~100 eager launches per window.
    self._overlap_graph = None
    if os.environ.get("FOSVOS_DP_GRAPH", "1") != "0":
        saved_sum = self.loss_sum.clone()
        self._window_overlapped()
        torch.cuda.current_stream(self.device).synchronize()
        for g in self.grads.values():
            g.zero_()
        for ws in (self.wgrad_ws or {}).values():
            ws.zero_()
        self.loss_sum.copy_(saved_sum)
        try:
            c0 = L.CALLS[0]
            graph = torch.cuda.CUDAGraph()
            hp = torch.cuda.Stream(device=self.device, priority=-1)
            hp.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.graph(graph, stream=hp, pool=self._step_graph.pool()):
                self._window_overlapped()
            torch.cuda.current_stream(self.device).wait_stream(hp)
            self._overlap_graph, self._calls_window = graph, L.CALLS[0] - c0
        except Exception as ex:                      # a PyTorch / NCCL build that cannot capture collectives: stay eager
            self._overlap_graph = None
            self.overlap_capture_error = f"{type(ex).__name__}: {ex}"[:300]
            torch.cuda.synchronize(self.device)
    _sync_cache_keys(self.net)


OnlineTrainer._capture_overlap_warmup = _capture_overlap_warmup


def finetune(net: OSVOS_VGG, frame: torch.Tensor, mask: torch.Tensor, n_iters: int, avg_grad_every_n: int = 5,
             optimizer: Optional[FusedSGD] = None, use_graph: bool = False,
             losses_out: Optional[list] = None) -> torch.Tensor:
    """Fine-tune ``net`` on one annotated frame (``train_online.py:70-101`` with a 1-sample loader).

    frame (1,3,H,W) fp32, mask (1,1,H,W) fp32.  Returns the device scalar of the summed
    per-iteration losses (no host sync inside the loop unless ``losses_out`` is given).  The trainer
    (and its CUDA graphs) is cached on ``net`` per frame size, so later calls replay."""
    L.require_device(next(net.parameters()).device)
    key = (int(frame.shape[-2]), int(frame.shape[-1]), int(avg_grad_every_n), bool(use_graph), id(optimizer), net.precision)
    cache = net.__dict__.setdefault("_trainers", {})
    tr = cache.get(key)
    if tr is None:
        # trainers (and their captured graphs) stay resident per frame size -- the online loop alternates between the
        # three train-time scales of 480x854 (custom_transforms.py:63-109); the oldest is dropped beyond MAX_RESIDENT_TRAINERS
        while len(cache) >= MAX_RESIDENT_TRAINERS:
            cache.pop(next(iter(cache)))
        tr = OnlineTrainer(net, key[0], key[1], avg_grad_every_n, optimizer, use_graph)
        cache[key] = tr
    elif optimizer is None:
        tr.reset()                              # a fresh optimizer per call, like get_optimizer() per sequence
    else:
        tr.loss_sum.zero_()
        tr.counter = 0
    tr.set_frame(frame, mask)
    return tr.run(n_iters, losses_out)


def finetune_samples(net: OSVOS_VGG, samples: Sequence[Tuple[torch.Tensor, torch.Tensor]], avg_grad_every_n: int = 5,
                     optimizer: Optional[FusedSGD] = None, use_graph: bool = True, losses_out: Optional[list] = None) -> torch.Tensor:
    """The online loop over an explicit list of per-iteration samples ``[(frame (1,3,h,w), mask (1,1,h,w)), ...]`` whose
    sizes may differ from one iteration to the next -- what the reference's loader delivers with train-time augmentation
    on (random flip, Resize scale in {0.5, 0.8, 1}: ``io_helper.py:62-70``, ``custom_transforms.py:63-109``; 480x854 gives
    240x427, 384x683, 480x854).  One resident trainer (and CUDA graph of the micro-iteration) per frame size; gradients,
    momentum, the accumulation counter and the optimizer-step graph are shared between them, so alternating sizes costs
    a graph replay, never a re-capture.  Returns the device scalar of the summed per-iteration losses."""
    dev = next(net.parameters()).device
    L.require_device(dev)
    n = int(avg_grad_every_n)
    key0 = (n, bool(use_graph), id(optimizer), net.precision)
    cache = net.__dict__.setdefault("_sample_trainers", {})
    if cache.get("key") != key0:
        cache.clear()
        cache["key"], cache["base"], cache["by_size"] = key0, None, {}
    by_size = cache["by_size"]
    if cache["base"] is not None:
        base = cache["base"]
        if optimizer is None:
            base.reset()                                # a fresh optimizer state per call, like get_optimizer() per sequence
        else:
            base.loss_sum.zero_()
            base.counter = 0
    for frame, _ in samples:                           # every size is captured at a window boundary, before the loop
        hw = (int(frame.shape[-2]), int(frame.shape[-1]))
        if hw not in by_size:
            if len(by_size) >= MAX_RESIDENT_TRAINERS:
                raise RuntimeError(f"finetune_samples: more than {MAX_RESIDENT_TRAINERS} distinct frame sizes")
            tr = OnlineTrainer(net, hw[0], hw[1], n, optimizer if cache["base"] is None else None, use_graph, fuse_window=False,
                               share_with=cache["base"])
            if cache["base"] is None:
                cache["base"] = tr
            tr.prepare()
            by_size[hw] = tr
    base = cache["base"]
    for frame, mask in samples:
        tr = by_size[(int(frame.shape[-2]), int(frame.shape[-1]))]
        tr.set_frame(frame, mask)
        tr.run(1, losses_out)
    return base.loss_sum


@torch.no_grad()
def infer_sequence(net: OSVOS_VGG, frames: torch.Tensor, batch_size: int = 1, want_prob: bool = False):
    """Segment every frame of a sequence (``experiment_helper.test``'s loop): returns uint8 masks
    (F,1,H,W) [and fp32 probabilities] on the device; sigmoid + threshold are fused into the
    side-chain kernel instead of a host-side numpy pass (``experiment_helper.py:55-57``)."""
    dev = next(net.parameters()).device
    L.require_device(dev)
    masks, probs = [], []
    for i in range(0, frames.shape[0], batch_size):
        fb = frames[i:i + batch_size]
        if fb.device != dev:
            fb = fb.to(dev, non_blocking=True)
        _, prob, mask = net.predict(fb)
        masks.append(mask)
        if want_prob:
            probs.append(prob)
    masks = torch.cat(masks)
    return (masks, torch.cat(probs)) if want_prob else masks


def sequences_for_rank(sequences: Sequence, rank: int, world_size: int) -> list:
    """``[s for i, s in enumerate(sequences) if i % group_size == group]`` -- the reference's manual
    job sharding (``train_online.py:184-186``); one process per GPU, no collective."""
    return [s for i, s in enumerate(sequences) if i % world_size == rank]


def region_iou(pred_masks: torch.Tensor, gt_masks: torch.Tensor) -> torch.Tensor:
    """(F,2) int64 intersection/union counts per frame: the integers behind DAVIS J."""
    return ops.mask_iou(pred_masks.reshape(pred_masks.shape[0], -1), gt_masks.reshape(gt_masks.shape[0], -1))
