"""OSVOS VGG-16 on B200: drop-in for the reference ``networks/osvos_vgg.py``.

Same constructor, sub-module names, ``state_dict`` layout and ``forward`` contract as the
reference ``OSVOS_VGG`` (``/root/reference/src/networks/osvos_vgg.py:17-83``): ``forward(x)``
takes an NCHW fp32 frame batch and returns ``[side0, side1, side2, side3, fused]`` logit maps,
each ``(N,1,H,W)`` fp32.  Parameters are ordinary fp32 ``nn.Parameter`` s in the reference
layout; everything between them and the five outputs runs in hand-written sm_100a kernels
behind the C ABI (``include/fosvos_b200.h``): NHWC activations, tcgen05 implicit-GEMM 3x3
convolutions, a fused side-output chain.  There is no CPU path and no cuDNN path.

Precision (``net.precision``):
  'bf16'       bf16 activations/weights, fp32 accumulation in TMEM (tcgen05 kernels) -- default
  'fp32'       fp32 activations/weights, fp32 FMA (direct kernels): the strict-parity mode (forward and backward)
  'fp32_tc'    fp32-equivalent arithmetic on the tcgen05 kernels: every fp32 value travels as three bf16 terms, a product
               keeps its six significant term products, accumulation is fp32 in TMEM ("bf16x6"; csrc/split.cu) -- the
               strict-parity mode of INFERENCE on the same kernels the bf16 numbers are measured on
  'bf16x3'     the same with two terms / three products (16-17 significant bits per operand)
  'bf16_simt'  bf16 data through the direct kernels (debug cross-check of the tcgen05 path)
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.modules as modules

from . import _lib as L
from . import leaf
from . import ops
from .layers import interp_surgery

_LAY_LIST = [[64, 64], ['M', 128, 128], ['M', 256, 256, 256], ['M', 512, 512, 512], ['M', 512, 512, 512]]
_IN_CHANNELS = [3, 64, 128, 256, 512]


def _act_dtype(precision: str) -> torch.dtype:
    return torch.float32 if precision == "fp32" else torch.bfloat16


class _PackedConv:
    """Derived kernel-side copies of one 3x3 conv's parameters (never serialised)."""

    def __init__(self):
        self.key = None
        self.w_fwd = self.w_dgrad = self.bias = None


class OSVOS_VGG(nn.Module):
    def __init__(self, pretrained=1):
        super(OSVOS_VGG, self).__init__()
        lay_list, in_channels = _LAY_LIST, _IN_CHANNELS
        stages = modules.ModuleList()
        side_prep = modules.ModuleList()
        score_dsn = modules.ModuleList()
        upscale = modules.ModuleList()
        upscale_ = modules.ModuleList()
        for i in range(0, len(lay_list)):
            stages.append(self._make_layers_osvos(lay_list[i], in_channels[i]))
            if i > 0:
                side_prep.append(nn.Conv2d(lay_list[i][-1], 16, kernel_size=3, padding=1))
                score_dsn.append(nn.Conv2d(16, 1, kernel_size=1, padding=0))
                upscale_.append(nn.ConvTranspose2d(1, 1, kernel_size=2 ** (1 + i), stride=2 ** i, bias=False))
                upscale.append(nn.ConvTranspose2d(16, 16, kernel_size=2 ** (1 + i), stride=2 ** i, bias=False))
        # attribute order fixes the state_dict key order (reference osvos_vgg.py:50-56)
        self.upscale = upscale
        self.upscale_ = upscale_
        self.stages = stages
        self.side_prep = side_prep
        self.score_dsn = score_dsn
        self.fuse = nn.Conv2d(64, 1, kernel_size=1, padding=0)
        self._initialize_weights(pretrained)

        self.precision = os.environ.get("FOSVOS_PRECISION", "bf16")       # (property: also turns the leaves into fosvos_b200.leaf classes)
        self.introspect = False                                            # True: always run module by module (hooks fire), see leaf.py
        self.fuse_pool = os.environ.get("FOSVOS_FUSE_POOL", "1") != "0"    # max pool written by the producing conv's epilogue
        # training: the fused pool also records the window index of every maximum (what autograd keeps for MaxPool2d's
        # backward), so the pool gradient never re-reads the full-resolution activation -- and conv1_2's output, which only
        # the pool consumes (stage 0 has no side output, osvos_vgg.py:63,68), is not written at all
        self.pool_arg = os.environ.get("FOSVOS_POOL_ARG", "1") != "0"
        self._keep_activations = False      # set by consumers that read every conv's output (prune.py: Taylor ranks)
        # side_prep through the row-stacked tcgen05 kernel with the 1x1 heads fused into its epilogue (conv_side_tc.cu)
        self.side_tc = os.environ.get("FOSVOS_SIDE_TC", "1") != "0"
        # side_prep convs and weight gradients are off the critical dependency chain: issue them on a second stream so
        # their launch gaps, prologues and tail waves overlap the next layer (also inside captured CUDA graphs)
        self.overlap = os.environ.get("FOSVOS_OVERLAP", "1") != "0"
        self._side_stream: Optional[torch.cuda.Stream] = None
        self._packed: Dict[int, _PackedConv] = {}
        self._side_key = None
        self._side_params: Optional[torch.Tensor] = None
        self._side_general = False
        self._side_separable = False

    @property
    def precision(self) -> str:
        return self.__dict__.get("_precision", "bf16")

    @precision.setter
    def precision(self, value: str) -> None:
        self.__dict__["_precision"] = value
        leaf.adopt(self, value)

    # ------------------------------------------------------------------ construction
    @staticmethod
    def _make_layers_osvos(cfg, in_channels):
        layers = []
        for v in cfg:
            if v == 'M':
                layers.append(nn.MaxPool2d(kernel_size=2, stride=2, ceil_mode=True))
            else:
                layers.extend([nn.Conv2d(in_channels, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)])
                in_channels = v
        return nn.Sequential(*layers)

    def _initialize_weights(self, pretrained):
        # reference osvos_vgg.py:97-116
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0, 0.001)
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.ConvTranspose2d):
                m.weight.data.zero_()
                m.weight.data = interp_surgery(m)
        if pretrained == 1:
            self._load_from_pytorch()
        elif pretrained == 2:
            self._load_from_caffe()

    def _load_from_pytorch(self) -> None:
        """``pretrained=1``: the 13 convolutions of torchvision's ImageNet VGG-16, in order, become the 13 stage
        convolutions (what reference osvos_vgg.py:118-129 does with deepcopy; needs the torchvision weights on disk)."""
        from torchvision.models import vgg16
        donors = [m for m in vgg16(pretrained=True).features if isinstance(m, nn.Conv2d)]
        ours = [m for stage in self.stages for m in stage if isinstance(m, nn.Conv2d)]
        if len(donors) != len(ours):
            raise RuntimeError(f"torchvision VGG-16 has {len(donors)} convolutions, OSVOS stages have {len(ours)}")
        for dst, src in zip(ours, donors):
            dst.weight = nn.Parameter(src.weight.detach().clone())
            dst.bias = nn.Parameter(src.bias.detach().clone())

    def _load_from_caffe(self, models_dir: Optional[str] = None) -> None:
        """``pretrained=2``: ``vgg_hed_caffe.mat`` (reference osvos_vgg.py:139-153; its directory comes from
        ``config.mypath`` there, from ``models_dir`` / ``$FOSVOS_MODELS_DIR`` here).  The file holds one weight array and
        one bias column per convolution; Caffe stores weights (kW, kH, Cin, Cout), hence the full transpose."""
        import scipy.io
        mat = scipy.io.loadmat(os.path.join(models_dir or os.environ.get("FOSVOS_MODELS_DIR", "."), 'vgg_hed_caffe.mat'))
        convs = [m for stage in self.stages for m in stage if isinstance(m, nn.Conv2d)]
        for k, conv in enumerate(convs):
            w = torch.from_numpy(mat['weights'][0][k].transpose())
            b = torch.from_numpy(mat['biases'][0][k][:, 0])
            if tuple(w.shape) != tuple(conv.weight.shape) or tuple(b.shape) != tuple(conv.bias.shape):
                raise RuntimeError(f"vgg_hed_caffe.mat: layer {k} holds {tuple(w.shape)} / {tuple(b.shape)}, "
                                   f"expected {tuple(conv.weight.shape)} / {tuple(conv.bias.shape)}")
            conv.weight.data = w
            conv.bias.data = b

    # ------------------------------------------------------------------ derived caches
    def _stage_convs(self) -> List[List[nn.Conv2d]]:
        return [[m for m in st if isinstance(m, nn.Conv2d)] for st in self.stages]

    _SPLIT_TERMS = {"fp32_tc": 3, "bf16x3": 2}

    def _impl(self) -> str:
        if self.precision == "bf16":
            return "tc"
        if self.precision in ("fp32", "bf16_simt"):
            return "simt"
        if self.precision in self._SPLIT_TERMS:
            return "tc_split"
        raise RuntimeError(f"fosvos_b200: unknown precision mode '{self.precision}'")

    def _packed_split_for(self, conv: nn.Conv2d, terms: int) -> _PackedConv:
        """Split-operand packed weight (one GEMM-K segment per kept term product) + padded bias of one conv, cached."""
        pc = self._packed.setdefault((id(conv), terms), _PackedConv())
        b = conv.bias
        key = (conv.weight.data_ptr(), conv.weight._version, None if b is None else (b.data_ptr(), b._version), terms,
               tuple(conv.weight.shape))
        if pc.key != key:
            same = pc.key is not None and pc.key[3:] == key[3:]
            pc.w_fwd = ops.pack_weight_split(conv.weight, terms, out=pc.w_fwd if same else None)
            pc.bias = ops.pad_bias(b, conv.out_channels, conv.weight.device, out=pc.bias if same else None)
            pc.key = key
        return pc

    def _run_pipeline_split(self, x: torch.Tensor, want_prob: bool, want_mask: bool):
        """Inference with fp32-equivalent arithmetic on the tensor cores (precision 'fp32_tc' / 'bf16x3'): split frames ->
        13 split convs (+ ReLU) and 4 split pools -> side_prep as plain fp32 -> the fp32 side chain."""
        terms = self._SPLIT_TERMS[self.precision]
        with torch.cuda.device(x.device):
            H, W = int(x.shape[-2]), int(x.shape[-1])
            a = ops.split_frames(x.float(), terms)
            sps: List[torch.Tensor] = []
            for si, convs in enumerate(self._stage_convs()):
                if si > 0:
                    a = ops.maxpool2x2_split(a, terms)
                for conv in convs:
                    pc = self._packed_split_for(conv, terms)
                    a = ops.conv3x3_split(a, pc.w_fwd, pc.bias, ops.pad8(conv.out_channels), terms, L.CONV_BIAS | L.CONV_RELU)
                if si > 0:
                    pc = self._packed_split_for(self.side_prep[si - 1], terms)
                    sps.append(ops.conv3x3_split(a, pc.w_fwd, pc.bias, 16, terms, L.CONV_BIAS))
            params = self._side()
            mode = 1 if self._side_general else (2 if self._side_separable else 0)
            return ops.side_fwd(sps, params, H, W, general=mode, want_prob=want_prob, want_mask=want_mask)

    def _wgrad_impl(self, cin_p: int) -> str:
        # bf16 mode: every layer takes the tensor-core weight gradient, the 3-channel first layer (padded to 8) through
        # its one-halo-box `c8` form (conv_wgrad_tc.cu)
        return "tc" if self._impl() == "tc" else "simt"

    def _packed_for(self, conv: nn.Conv2d, need_dgrad: bool) -> _PackedConv:
        pc = self._packed.setdefault(id(conv), _PackedConv())
        b = conv.bias
        key = (conv.weight.data_ptr(), conv.weight._version, None if b is None else (b.data_ptr(), b._version),
               self.precision, tuple(conv.weight.shape))
        dt = _act_dtype(self.precision)
        tc = self._impl() == "tc"
        stale = pc.key != key
        if stale and (pc.key is None or pc.key[3:] != key[3:] or pc.key[0] != key[0]):
            pc.w_fwd = pc.w_dgrad = pc.bias = None      # different precision / shape / storage: new buffers
        # same shape: re-pack INTO the existing buffers, so captured CUDA graphs stay valid
        if pc.w_fwd is None or stale:
            pc.w_fwd = ops.pack_weight(conv.weight, L.W_TC_FWD if tc else L.W_SIMT_FWD, dt, out=pc.w_fwd)
            pc.bias = ops.pad_bias(b, conv.out_channels, conv.weight.device, out=pc.bias)
        if (need_dgrad and pc.w_dgrad is None) or (stale and pc.w_dgrad is not None):
            pc.w_dgrad = ops.pack_weight(conv.weight, L.W_TC_DGRAD if tc else L.W_SIMT_DGRAD, dt, out=pc.w_dgrad)
        pc.key = key
        return pc

    def invalidate_packed(self) -> None:
        """Drop the packed copies (call after mutating parameters without bumping ``_version``)."""
        self._packed.clear()
        self._side_key = None

    def _side_cache_key(self):
        ps = [m.weight for m in self.upscale] + [m.weight for m in self.upscale_] + \
             [m.weight for m in self.score_dsn] + [m.bias for m in self.score_dsn] + [self.fuse.weight, self.fuse.bias]
        return tuple((p.data_ptr(), p._version) for p in ps)

    def _side(self) -> torch.Tensor:
        key = self._side_cache_key()
        if key != self._side_key:
            up_key = key[:4]
            if self._side_key is None or self._side_key[:4] != up_key:
                # structure check of the up-sampling weights (one sync, only when they change)
                self._side_general = bool(ops.side_check_diagonal([m.weight for m in self.upscale]).item() != 0)
            recheck = self._side_key is None or self._side_key[:8] != key[:8]
            self._side_params = ops.side_params_prepare([m.weight for m in self.upscale], [m.weight for m in self.upscale_],
                                                        [m.weight for m in self.score_dsn], [m.bias for m in self.score_dsn],
                                                        self.fuse.weight, self.fuse.bias, out=self._side_params)
            if recheck:
                # exact separability of the two shared kernels (one more sync, only when up-sampling weights change)
                self._side_separable = (not self._side_general) and ops.side_separable(self._side_params)
            self._side_key = key
        return self._side_params

    # ------------------------------------------------------------------ the path
    #: BGR mean of the DAVIS loaders (reference dataloaders/davis_2016.py:24)
    MEANVAL = (104.00699, 116.66877, 122.67892)

    def _run_forward(self, x: torch.Tensor, save: bool, want_prob: bool = False, want_mask: bool = False):
        """NHWC pipeline.  Returns (outs, prob, mask, saved) ; saved holds what backward needs.
        ``x``: (N,3,H,W) fp32 mean-subtracted frames (the reference contract), or raw (N,H,W,3) uint8 frames as
        cv2 delivers them -- mean subtraction and layout change then happen in one ingest kernel."""
        L.require_device(x.device)
        if self._impl() == "tc_split":
            if save:
                raise RuntimeError(f"OSVOS_VGG: precision '{self.precision}' is an inference mode (split-operand tensor-core forward); "
                                   "train with precision 'fp32' (strict) or 'bf16', or call it under torch.no_grad() / net.predict()")
            if x.dtype == torch.uint8:
                if x.dim() != 4 or x.shape[3] != 3:
                    raise RuntimeError(f"OSVOS_VGG: uint8 frames must be (N,H,W,3), got {tuple(x.shape)}")
                x = ops.nhwc_to_nchw(ops.ingest_u8(x, self.MEANVAL, torch.float32), 3)
            if x.dim() != 4 or x.shape[1] != self.stages[0][0].in_channels:
                raise RuntimeError(f"OSVOS_VGG.forward expects (N,{self.stages[0][0].in_channels},H,W), got {tuple(x.shape)}")
            outs, prob, mask = self._run_pipeline_split(x, want_prob, want_mask)
            return outs, prob, mask, None
        if x.dtype == torch.uint8:
            if x.dim() != 4 or x.shape[3] != 3 or self.stages[0][0].in_channels != 3:
                raise RuntimeError(f"OSVOS_VGG: uint8 frames must be (N,H,W,3), got {tuple(x.shape)}")
            return self._run_pipeline(ops.ingest_u8(x, self.MEANVAL, _act_dtype(self.precision)), int(x.shape[1]), int(x.shape[2]),
                                      save, want_prob, want_mask)
        if x.dim() != 4 or x.shape[1] != self.stages[0][0].in_channels:
            raise RuntimeError(f"OSVOS_VGG.forward expects (N,{self.stages[0][0].in_channels},H,W), got {tuple(x.shape)}")
        H, W = int(x.shape[-2]), int(x.shape[-1])
        return self._run_pipeline(ops.nchw_to_nhwc(x.float(), _act_dtype(self.precision)), H, W, save, want_prob, want_mask)

    def _aux_stream(self, device) -> Optional[torch.cuda.Stream]:
        if not self.overlap:
            return None
        if self._side_stream is None or self._side_stream.device != device:
            self._side_stream = torch.cuda.Stream(device=device)
        return self._side_stream

    def _run_pipeline(self, a: torch.Tensor, H: int, W: int, save: bool, want_prob: bool, want_mask: bool):
        with torch.cuda.device(a.device):          # streams, events and launches all belong to the tensors' GPU
            return self._run_pipeline_on_device(a, H, W, save, want_prob, want_mask)

    def _run_pipeline_on_device(self, a: torch.Tensor, H: int, W: int, save: bool, want_prob: bool, want_mask: bool):
        impl = self._impl()
        main = torch.cuda.current_stream(a.device)
        aux = self._aux_stream(a.device)
        params = self._side()
        mode = 1 if self._side_general else (2 if self._side_separable else 0)
        n = int(a.shape[0])
        # side_prep through the row-stacked kernel: the two 1x1 heads of every stage leave its epilogue (fp32 accumulators),
        # so the side-chain needs no heads launch and -- in inference -- the 16-channel maps are never written
        stacked = impl == "tc" and self.side_tc and all(ops.side_tc_supported(ops.pad8(c.in_channels)) and c.out_channels == 16
                                                        for c in self.side_prep)
        fuse_heads = stacked and mode != 1
        zs_flat = zs_views = heads = None
        if fuse_heads:
            hs, ws, h_, w_ = [], [], H, W
            for _ in range(4):
                h_, w_ = (h_ + 1) // 2, (w_ + 1) // 2
                hs.append(h_)
                ws.append(w_)
            zs_flat, zs_views = ops.side_zs_workspace(n, hs, ws, a.device)
            heads = ops.side_heads_views(params)
        conv_in: List[torch.Tensor] = []        # input activation of every stage conv, in order
        conv_out: List[torch.Tensor] = []
        pool_in: List[Optional[torch.Tensor]] = []
        pool_arg: List[Optional[torch.Tensor]] = []        # per pool: recorded window indices (training, fused pool) or None
        arg: Optional[torch.Tensor] = None
        stage_out: List[torch.Tensor] = []
        sps: List[torch.Tensor] = []
        pooled: Optional[torch.Tensor] = None   # the next stage's input when the pool rode along in the conv epilogue
        for si, convs in enumerate(self._stage_convs()):
            if si > 0:
                pool_in.append(a)
                pool_arg.append(arg)
                a = pooled if pooled is not None else ops.maxpool2x2(a)
                pooled = arg = None
            for ci, conv in enumerate(convs):
                pc = self._packed_for(conv, need_dgrad=save)
                conv_in.append(a)
                cp = ops.pad8(conv.out_channels)
                if impl == "tc" and si < 4 and ci == len(convs) - 1 and cp >= 64 and a.shape[3] > 8 and self.fuse_pool:
                    if si == 0 and not save:
                        # inference: nothing but the pool reads conv1_2 (stage 0 has no side output, osvos_vgg.py:63,68):
                        # its 52 MB/frame full-resolution map is never written
                        pooled = ops.conv3x3_pool_only(a, pc.w_fwd, pc.bias, cp, L.CONV_BIAS | L.CONV_RELU)
                        a = None
                    elif save and getattr(self, "pool_arg", True) and cp % 32 == 0:
                        a, pooled, arg = ops.conv3x3_pool_arg(a, pc.w_fwd, pc.bias, cp, L.CONV_BIAS | L.CONV_RELU,
                                                              want_y=si > 0 or getattr(self, "_keep_activations", False))
                    else:
                        a, pooled = ops.conv3x3_pool(a, pc.w_fwd, pc.bias, cp, L.CONV_BIAS | L.CONV_RELU)
                else:
                    a = ops.conv3x3(a, pc.w_fwd, pc.bias, cp, L.CONV_BIAS | L.CONV_RELU, impl=impl)
                conv_out.append(a)                          # (None for a pool-only launch: inference keeps no activations)
            stage_out.append(a)
            if si > 0:
                sp_conv = self.side_prep[si - 1]
                pc = self._packed_for(sp_conv, need_dgrad=save)
                want_sp = save or not fuse_heads           # the 16-channel map itself: backward and the general side path read it
                # output allocated on the main stream (its pool), possibly written on the auxiliary one
                sp = torch.empty((a.shape[0], a.shape[1], a.shape[2], 16), dtype=a.dtype, device=a.device) if want_sp else None

                def run_side_prep(a=a, pc=pc, sp=sp, i=si - 1):
                    if stacked:
                        ops.conv3x3_side(a, pc.w_fwd, pc.bias, out=sp, zs=zs_views[i] if fuse_heads else None,
                                         heads=heads[i] if fuse_heads else None, want_y=sp is not None)
                    else:
                        ops.conv3x3(a, pc.w_fwd, pc.bias, 16, L.CONV_BIAS, out=sp, impl=impl)

                if aux is None:
                    run_side_prep()
                else:
                    ev = torch.cuda.Event()
                    ev.record(main)
                    aux.wait_event(ev)
                    with torch.cuda.stream(aux):
                        run_side_prep()
                sps.append(sp)
        if aux is not None:
            main.wait_stream(aux)
        if fuse_heads:
            outs, prob, mask = ops.side_fwd_heads_done(zs_flat, hs, ws, params, n, H, W, general=mode, want_prob=want_prob, want_mask=want_mask)
        else:
            outs, prob, mask = ops.side_fwd(sps, params, H, W, general=mode, want_prob=want_prob, want_mask=want_mask)
        saved = None
        if save:
            saved = dict(conv_in=conv_in, conv_out=conv_out, pool_in=pool_in, pool_arg=pool_arg, stage_out=stage_out, sps=sps, H=H, W=W,
                         params=params)
        return outs, prob, mask, saved

    def _run_backward(self, saved, douts: Sequence[Optional[torch.Tensor]], grads: Dict[str, torch.Tensor],
                      wgrad_ws: Optional[Dict[str, torch.Tensor]] = None,
                      taylor: Optional[Dict[str, torch.Tensor]] = None, stage_done=None) -> None:
        """Accumulate (+=) parameter gradients into ``grads`` (name -> fp32 tensor, reference layout).
        With ``wgrad_ws`` (conv name -> live accumulator, ``ops.wgrad_workspace``) the tensor-core weight
        gradients are left in their accumulators; the caller folds them into ``grads`` once per optimizer
        step (``ops.conv3x3_wgrad_finish``).  With ``taylor`` (stage-conv name -> (Cout,) fp32) the pruning
        criterion sum(activation * gradient) / (N H W) of every stage conv is accumulated on the way
        (reference ``prune.py:163-178``; post-ReLU activation x masked gradient == the hooked product).
        ``stage_done(bucket, aux_stream)`` is called when all weight gradients of stage ``4 - bucket`` (and of the
        ``side_prep`` conv hanging off it) have been issued -- the data-parallel trainer starts that bucket's all-reduce."""
        if self._side_general:
            raise RuntimeError("fosvos_b200: backward through non-diagonal `upscale` weights is not supported "
                               "(the reference keeps them fixed with lr=0, network_provider.py:154-155)")
        with torch.cuda.device(saved["sps"][0].device):
            self._run_backward_on_device(saved, douts, grads, wgrad_ws, taylor, stage_done)

    def _run_backward_on_device(self, saved, douts, grads, wgrad_ws, taylor, stage_done=None) -> None:
        impl = self._impl()
        H, W = saved["H"], saved["W"]
        sps = saved["sps"]
        if douts[4] is None:
            douts = list(douts)
            douts[4] = torch.zeros_like(next(d for d in douts if d is not None))
        g = lambda k: grads.get(k)  # noqa: E731
        dev = sps[0].device
        main = torch.cuda.current_stream(dev)
        aux = self._aux_stream(dev)
        keep = []          # operands of in-flight weight gradients: not released before the streams join

        def wgrad(name: str, x_in: torch.Tensor, dz: torch.Tensor) -> None:
            impl_w = self._wgrad_impl(x_in.shape[3])

            def run():
                if wgrad_ws is not None and impl_w == "tc":
                    ops.conv3x3_wgrad_accumulate(x_in, dz, wgrad_ws[name], g(name + ".bias"), grads[name + ".weight"].shape[0])
                else:
                    ops.conv3x3_wgrad(x_in, dz, grads[name + ".weight"], g(name + ".bias"), impl=impl_w)

            if aux is None:
                run()
                return
            ev = torch.cuda.Event()
            ev.record(main)            # dz (and x_in) are complete on the main stream here
            aux.wait_event(ev)
            with torch.cuda.stream(aux):
                run()
            keep.append((x_in, dz))

        dsp = ops.side_bwd(sps, saved["params"], douts, H, W, g("fuse.weight"), g("fuse.bias"),
                           [g(f"score_dsn.{i}.weight") for i in range(4)], [g(f"score_dsn.{i}.bias") for i in range(4)])
        convs = self._stage_convs()
        names = [[f"stages.{si}.{mi}" for mi, m in enumerate(st) if isinstance(m, nn.Conv2d)] for si, st in enumerate(self.stages)]
        # flat index of the first conv of each stage
        first = [0]
        for c in convs[:-1]:
            first.append(first[-1] + len(c))
        # The side_prep branch of every stage output only needs dsp: its data gradient (masked by the stage output's
        # ReLU) is issued one stage ahead, off the main chain; the pool gradient coming down the backbone is then
        # added INTO it (fan-in in the pool kernel, a coalesced elementwise read).
        dA_side: List[Optional[torch.Tensor]] = [None] * 5
        side_done: List[Optional[torch.cuda.Event]] = [None] * 5

        def side_branch(si: int, on_aux: bool) -> None:
            a_out = saved["stage_out"][si]
            pc = self._packed_for(self.side_prep[si - 1], need_dgrad=True)
            dA_side[si] = torch.empty_like(a_out)
            if aux is None or not on_aux:
                ops.conv3x3(dsp[si - 1], pc.w_dgrad, None, a_out.shape[3], L.CONV_MASK, mask=a_out, out=dA_side[si], impl=impl)
            else:
                ev = torch.cuda.Event()
                ev.record(main)
                aux.wait_event(ev)
                with torch.cuda.stream(aux):
                    ops.conv3x3(dsp[si - 1], pc.w_dgrad, None, a_out.shape[3], L.CONV_MASK, mask=a_out, out=dA_side[si], impl=impl)
                    side_done[si] = torch.cuda.Event()
                    side_done[si].record(aux)
            wgrad(f"side_prep.{si - 1}", a_out, dsp[si - 1])

        # where the two gradients of a stage output meet: "pool" (default) = the side_prep data gradient is written on its own
        # (off the main chain, on the auxiliary stream) and the pool's backward adds it while routing (coalesced 16-byte reads);
        # "epilogue" = the side_prep data gradient's epilogue accumulates into the pool gradient (a per-pixel read-modify-write:
        # 32 lines per warp instruction).  Measured on the sequence job: 0.4156 against 0.4221 s of fine-tune per sequence.
        fanin_pool = os.environ.get("FOSVOS_BWD_FANIN", "pool") == "pool"
        if fanin_pool:
            side_branch(4, on_aux=False)
        dA: Optional[torch.Tensor] = None          # gradient w.r.t. the current stage's output (pre-activation-masked)
        for si in range(4, -1, -1):
            if fanin_pool:
                if si == 4:
                    dA = dA_side[4]
                if si - 1 > 0:
                    side_branch(si - 1, on_aux=True)       # needed by the pool gradient at the END of this stage
            elif si > 0:
                # fan-in in the conv epilogue: dA (+)= mask * convT(dsp) on the main chain
                a_out = saved["stage_out"][si]
                pc = self._packed_for(self.side_prep[si - 1], need_dgrad=True)
                wgrad(f"side_prep.{si - 1}", a_out, dsp[si - 1])
                flags = L.CONV_MASK | (L.CONV_ACCUMULATE if dA is not None else 0)
                dA = ops.conv3x3(dsp[si - 1], pc.w_dgrad, None, a_out.shape[3], flags, mask=a_out, out=dA, impl=impl)
            dz = dA
            for j in range(len(convs[si]) - 1, -1, -1):
                conv = convs[si][j]
                k = first[si] + j
                x_in = saved["conv_in"][k]
                name = names[si][j]
                wgrad(name, x_in, dz)
                if taylor is not None and name in taylor:
                    if saved["conv_out"][k] is None:
                        raise RuntimeError("fosvos_b200: Taylor ranks need every conv's output: run the forward pass with net._keep_activations = True")
                    ops.taylor_rank(saved["conv_out"][k], dz, taylor[name])
                if si == 0 and j == 0:
                    break
                pc = self._packed_for(conv, need_dgrad=True)
                # dX masked by (x_in > 0): x_in is the previous layer's post-ReLU output (or its pooled copy)
                dz = ops.conv3x3(dz, pc.w_dgrad, None, x_in.shape[3], L.CONV_MASK, mask=x_in, impl=impl)
            if stage_done is not None:
                stage_done(4 - si, aux)
            if si > 0:
                if si - 1 > 0 and side_done[si - 1] is not None:
                    main.wait_event(side_done[si - 1])
                arg = saved.get("pool_arg", [None] * 4)[si - 1]
                if arg is not None:
                    hh, ww = self._pool_input_hw(saved, si - 1)
                    dA = ops.maxpool2x2_bwd_arg(arg, dz, hh, ww, add=dA_side[si - 1])
                else:
                    dA = ops.maxpool2x2_bwd(saved["pool_in"][si - 1], dz, add=dA_side[si - 1])
        if aux is not None:
            main.wait_stream(aux)
        keep.clear()

    @staticmethod
    def _pool_input_hw(saved, pool_index: int):
        """(height, width) of the input of pool ``pool_index`` (0 = after stage 0): ceil-mode halvings of the frame size."""
        h, w = saved["H"], saved["W"]
        for _ in range(pool_index):
            h, w = (h + 1) // 2, (w + 1) // 2
        return h, w

    def _grad_names(self) -> List[str]:
        return [n for n, p in self.named_parameters() if p.requires_grad and not n.startswith("upscale")]

    # ------------------------------------------------------------------ introspection (module by module)
    def _leaf_hooks_present(self) -> bool:
        for root in (self.stages, self.side_prep):
            for m in root.modules():
                if m._forward_hooks or m._forward_pre_hooks:
                    return True
        return False

    def _forward_introspect(self, x: torch.Tensor):
        """Module-by-module forward for hook-style consumers (prune.py:94-103; SURVEY 8b "hookable per-conv outputs"): every
        leaf of ``stages`` / ``side_prep`` is CALLED (so ``register_forward_hook`` fires, and tensor hooks registered on the
        outputs fire in backward), each running this repo's kernels through ``leaf.py``; the side chain stays one fused
        node.  Slower than the fused pipeline (layout changes around every module); same arithmetic."""
        L.require_device(x.device)
        if x.dtype == torch.uint8 or x.dim() != 4:
            raise RuntimeError("OSVOS_VGG introspection mode expects the reference contract: (N,C,H,W) float frames")
        leaf.adopt(self, self.precision)             # surgery may have assigned plain torch leaves
        H, W = int(x.shape[-2]), int(x.shape[-1])
        a = x.float()
        sps = []
        for si, stage in enumerate(self.stages):
            a = stage(a)                              # osvos_vgg.py:63,68
            if si > 0:
                sps.append(self.side_prep[si - 1](a))   # :69
        heads = [m.weight for m in self.score_dsn] + [m.bias for m in self.score_dsn]
        return list(_SideChainFunction.apply(self, H, W, self.fuse.weight, self.fuse.bias, *heads, *sps))

    def forward(self, x):
        if self.introspect or self._leaf_hooks_present():
            return self._forward_introspect(x)
        params = dict(self.named_parameters())
        if torch.is_grad_enabled() and any(p.requires_grad for p in params.values()):
            names = self._grad_names()
            outs = _OSVOSFunction.apply(self, x, names, *[params[n] for n in names])
            return list(outs)
        outs, _, _, _ = self._run_forward(x, save=False)
        return outs

    @torch.no_grad()
    def predict(self, x):
        """Inference with the consumer-side post-processing fused in: returns
        (outs, prob, mask): sigmoid (util/experiment_helper.py:57) and the 0.5 threshold
        (run_webcam.py:92-93) of the fused map come out of the same kernel."""
        outs, prob, mask, _ = self._run_forward(x, save=False, want_prob=True, want_mask=True)
        return outs, prob, mask


class _OSVOSFunction(torch.autograd.Function):
    """Autograd bridge: the whole network is one node whose backward runs our kernels.
    Gradients w.r.t. the input frame and the (lr=0) up-sampling weights are not produced."""

    @staticmethod
    def forward(ctx, net: OSVOS_VGG, x, names, *params):
        outs, _, _, saved = net._run_forward(x.detach(), save=True)
        ctx.net, ctx.saved, ctx.names = net, saved, names
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.device = x.device
        # outputs without a consumer hand backward None instead of a zero-filled full-resolution map: the online loop
        # uses outputs[-1] only (train_online.py:81), so the four side-map gradients are skipped, not multiplied by zero
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        if ctx.saved is None:
            raise RuntimeError("OSVOS_VGG: the activations of this forward pass were released by the first backward "
                               "(call forward again; retain_graph=True is not supported by the fused backward)")
        grads = {n: torch.zeros(s, dtype=torch.float32, device=ctx.device) for n, s in zip(ctx.names, ctx.shapes)}
        ctx.net._run_backward(ctx.saved, [None if d is None else d.contiguous() for d in douts], grads)
        ctx.saved = None
        return (None, None, None) + tuple(grads[n] for n in ctx.names)


class _SideChainFunction(torch.autograd.Function):
    """The fused side chain (osvos_vgg.py:71-81) as one autograd node of the introspection path: four side_prep maps
    (N,16,h,w) -> [side0..3, fused]."""

    @staticmethod
    def forward(ctx, net: OSVOS_VGG, H, W, fuse_w, fuse_b, sw0, sw1, sw2, sw3, sb0, sb1, sb2, sb3, *sps):
        dt = _act_dtype(net.precision)
        params = net._side()
        mode = 1 if net._side_general else (2 if net._side_separable else 0)
        sph = [ops.nchw_to_nhwc(t.detach().float().contiguous(), dt, cp=16) for t in sps]
        outs, _, _ = ops.side_fwd(sph, params, H, W, general=mode)
        ctx.net, ctx.sph, ctx.params, ctx.hw = net, sph, params, (H, W)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        net = ctx.net
        if net._side_general:
            raise RuntimeError("fosvos_b200: backward through non-diagonal `upscale` weights is not supported "
                               "(the reference keeps them fixed with lr=0, network_provider.py:154-155)")
        H, W = ctx.hw
        dev = ctx.sph[0].device
        douts = [None if d is None else d.float().contiguous() for d in douts]
        if douts[4] is None:
            douts[4] = torch.zeros((ctx.sph[0].shape[0], 1, H, W), dtype=torch.float32, device=dev)
        dfw, dfb = torch.zeros_like(net.fuse.weight), torch.zeros_like(net.fuse.bias)
        dsw = [torch.zeros_like(m.weight) for m in net.score_dsn]
        dsb = [torch.zeros_like(m.bias) for m in net.score_dsn]
        dsp = ops.side_bwd(ctx.sph, ctx.params, douts, H, W, dfw, dfb, dsw, dsb)
        dsp_nchw = [ops.nhwc_to_nchw(t, 16) for t in dsp]
        return (None, None, None, dfw, dfb, *dsw, *dsb, *dsp_nchw)
