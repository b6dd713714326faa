#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's configs[1].

Workload ("step"): ONE synthetic DAVIS-shaped sequence job on one B200 --
one-shot online fine-tuning (500 SGD iterations, batch 1, 480x854, avg_grad_every_n=5, the online
param groups) on the annotated first frame, then inference over the 80 frames of the sequence
(sigmoid + 0.5 threshold fused, masks produced on the device).  With N GPUs every rank runs its own
sequences (sharded by sequence, no collective: reference train_online.py:184-186) -> weak scaling.

    python bench.py --gpus N --steps K --warmup W            our arm
    python bench.py --impl reference ...                     the reference algorithm on the host CPU

Prints ONE JSON line.  `value` = frames segmented per second of whole-job time (fine-tune included),
inputs resident in HBM; `e2e` = the same through the public API from pinned HOST buffers with the
H2D/D2H copies inside the timed region.  Extra keys split the step into its parts
(`inference_fps`, `finetune_s_per_sequence`) and carry the roofline / CPU-baseline objects.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

H, W = 480, 854
FWD_GFLOP = 258.23          # per frame, 17 3x3 convs
ITER_GFLOP = 773.27         # per fine-tune iteration: fwd + dgrad + wgrad (no dgrad for conv1_1)
METRIC = "osvos_vgg16_480x854_sequence_frames_per_sec"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for l in self.lines:
            f = [t.strip() for t in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


def calibrated_state_dict(kind="parent"):
    """Seeded structured weights calibrated on a small frame with the device forward (data only)."""
    from fosvos_b200 import synth
    import fosvos_b200 as FB
    xs, ms = synth.make_frame(0, 0, 120, 214)
    def fwd(sd, x):
        net = FB.OSVOS_VGG(pretrained=0)
        net.load_state_dict(sd)
        net = net.cuda()
        net.precision = "fp32"
        with torch.no_grad():
            return [o.cpu() for o in net(x.cuda())]
    # 'parent' weights: trained-like activation scale, so that the reference's lr = 1e-8 fine-tune converges (synth.py)
    return synth.calibrate(synth.make_state_dict(0, kind), fwd, xs, mask=ms)


def make_sequence_gpu(seq: int, n_frames: int):
    """80 distinct frames derived on the device from a few CPU-synthesised key frames (the CPU
    generator takes ~0.1 s per 480x854 frame): frame f = key[f % K] rolled by f pixels."""
    from fosvos_b200 import synth
    K = 4
    keys = [synth.make_frame(seq, k, H, W) for k in range(K)]
    frames = torch.stack([torch.roll(keys[f % K][0][0], shifts=f, dims=2) for f in range(n_frames)])
    masks = torch.stack([torch.roll(keys[f % K][1][0], shifts=f, dims=2) for f in range(n_frames)])
    return frames.contiguous(), masks.contiguous()


def run_ours(args):
    import fosvos_b200 as FB
    from fosvos_b200 import _lib as L, ops, sharding
    rank, world = sharding.init_distributed("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L.require_device(dev)
    peaks = load_peaks()
    sd0 = calibrated_state_dict(args.weights)
    frames_h, masks_h = make_sequence_gpu(rank, args.frames)
    frames_pin, masks_pin = frames_h.pin_memory(), masks_h.pin_memory()
    frames_d, masks_d = frames_h.to(dev), masks_h.to(dev)
    # the raw frames a loader holds: uint8 HWC (cv2 order); frame = uint8 - mean exactly (davis_2016.py:127-128)
    mean = torch.tensor(FB.OSVOS_VGG.MEANVAL, dtype=torch.float32).view(1, 3, 1, 1)
    frames_u8 = (frames_h + mean).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    assert torch.equal(frames_u8.permute(0, 3, 1, 2).float() - mean, frames_h), "synthetic frames are uint8 - mean"
    frames_u8_pin = frames_u8.pin_memory()

    def new_net():
        net = FB.OSVOS_VGG(pretrained=0)
        net.load_state_dict(sd0)                 # NetworkProvider.load_model (network_provider.py:53-58)
        net = net.to(dev)
        net.precision = args.precision
        return net

    launches = L.CALLS            # ABI compute calls (>= 1 kernel each); graph replays add their node count

    # One resident worker per GPU, as a production process would run it: the network object, its
    # optimizer and the captured CUDA graphs persist; every sequence re-loads the parent weights in
    # place (NetworkProvider.load_model), resets gradients/momentum, and fine-tunes on its own frame.
    from fosvos_b200.online import OnlineTrainer
    net = new_net()
    sd_dev = {k: v.to(dev) for k, v in sd0.items()}
    trainer = OnlineTrainer(net, H, W, args.avg_grad_every_n, FB.get_optimizer_online(net), use_graph=bool(args.graph),
                            fuse_window=bool(args.fuse_window))
    masks_out_pin = torch.empty((args.frames, 1, H, W), dtype=torch.uint8).pin_memory()

    def sequence_job(host: bool):
        """fine-tune on frame 0 + its mask, then segment all frames. Returns (t_finetune_ms, t_infer_ms) device times."""
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        trainer.reset(sd_dev)
        if host:
            trainer.set_frame(frames_pin[0:1], masks_pin[0:1])         # H2D from pinned memory
        else:
            trainer.set_frame(frames_d[0:1], masks_d[0:1])
        trainer.run(args.iters)
        e[1].record()
        out_masks = []
        for i in range(0, args.frames, args.batch):
            # end to end: raw uint8 frames cross PCIe (4x fewer bytes), mean subtraction + layout happen in the ingest kernel
            fb = frames_u8_pin[i:i + args.batch].to(dev, non_blocking=True) if host else frames_d[i:i + args.batch]
            _, _, mask = net.predict(fb)
            if host:
                masks_out_pin[i:i + args.batch].copy_(mask, non_blocking=True)     # D2H of the result
            else:
                out_masks.append(mask)
        e[2].record()
        torch.cuda.synchronize()
        return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), out_masks

    for _ in range(args.warmup):
        sequence_job(False)
    # ---- timed: device-resident ------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sharding.barrier(); torch.cuda.synchronize()
    launches[0] = 0
    t_ft = t_inf = 0.0
    for _ in range(args.steps):
        a, b, _ = sequence_job(False)
        t_ft += a; t_inf += b
    torch.cuda.synchronize(); sharding.barrier()
    n_launch = launches[0]
    clocks = sampler.stop() if rank == 0 else None
    t_total = sharding.max_over_ranks(t_ft + t_inf)
    t_ft_max, t_inf_max = sharding.max_over_ranks(t_ft), sharding.max_over_ranks(t_inf)
    total_frames = world * args.steps * args.frames
    value = total_frames / (t_total / 1e3)

    # ---- timed: end to end from pinned host buffers --------------------------------------------
    sequence_job(True)
    sharding.barrier(); torch.cuda.synchronize()
    te = 0.0
    for _ in range(args.steps):
        a, b, _ = sequence_job(True)
        te += a + b
    torch.cuda.synchronize(); sharding.barrier()
    te = sharding.max_over_ranks(te)
    e2e = dict(value=total_frames / (te / 1e3), unit="frames/s",
               h2d_bytes_per_step=int(args.frames * 3 * H * W + 4 * H * W * 4),      # uint8 frames + the fp32 annotated frame and mask
               d2h_bytes_per_step=int(args.frames * H * W))

    # ---- for transparency: the strictly sequential loop (one micro-iteration per pass), one sequence, not part of `value`
    t_seq_ms = None
    if args.fuse_window and rank == 0:
        net_s = new_net()
        tr_s = OnlineTrainer(net_s, H, W, args.avg_grad_every_n, FB.get_optimizer_online(net_s), use_graph=bool(args.graph), fuse_window=False)
        tr_s.set_frame(frames_d[0:1], masks_d[0:1])
        tr_s.run(2 * args.avg_grad_every_n)
        tr_s.reset(sd_dev)
        tr_s.set_frame(frames_d[0:1], masks_d[0:1])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); tr_s.run(args.iters); e1.record(); torch.cuda.synchronize()
        t_seq_ms = e0.elapsed_time(e1)
        del tr_s, net_s

    out = None
    extras = {}
    if rank == 0 and world == 1:
        # ---- parity of the benchmarked job: one more (untimed) job with the loss trajectory read back, then the same job
        # on the on-box oracle; then the stock-PyTorch/cuDNN arms on this GPU ------------------------------------------
        job_s = (t_ft_max + t_inf_max) / 1e3 / args.steps
        if args.parity:
            losses_ours = []
            trainer.reset(sd_dev)
            trainer.set_frame(frames_d[0:1], masks_d[0:1])
            trainer.run(args.iters, losses_ours)
            extras["parity"] = parity_check(net, trainer, sd0, frames_d, masks_d, frames_h, args, losses_ours)
        if args.gpu_reference:
            extras["gpu_reference"] = gpu_reference(sd0, frames_d, masks_d, args, job_s)
    # ---- BASELINE configs[3] and configs[4]: the multi-GPU pipelines (every rank takes part) --------------------------
    # (a failing extra leg is reported in its own object; it must not take the headline line down with it)
    if args.config4:
        try:
            extras["config4"] = config4_leg(args, rank, world, dev, sd_dev, trainer, net)
        except Exception as ex:
            extras["config4"] = {"error": f"{type(ex).__name__}: {ex}"[:400]}
    if args.config5 and world > 1:
        try:
            extras["config5"] = config5_leg(args, rank, world, dev, sd0, frames_d, masks_d)
        except Exception as ex:
            extras["config5"] = {"error": f"{type(ex).__name__}: {ex}"[:400]}
    if rank == 0:
        # ---- roofline of the dominant kernel family (3x3 conv implicit GEMM), measured live -------
        roof = conv_roofline(new_net(), frames_d[:args.batch], peaks, args.precision)
        side = side_roofline(new_net(), frames_d[:args.batch], peaks)
        loss_roof = loss_roofline(args.batch, dev, peaks)
        cpu = cpu_baseline(sd0, frames_h, masks_h, args) if world == 1 else None
        if world == 1:
            extras["roofline_wgrad"], extras["roofline_dgrad"] = backward_rooflines(new_net(), frames_d, peaks, args.avg_grad_every_n)
            if args.config3:
                extras["config3"] = config3_leg(sd0, frames_d, frames_h, peaks, dev)
            if args.fp32_modes:
                try:
                    extras["fp32_modes"] = fp32_modes_leg(sd0, frames_d, frames_h, dev)
                except Exception as ex:
                    extras["fp32_modes"] = {"error": f"{type(ex).__name__}: {ex}"[:400]}
        ft_tflops = ITER_GFLOP * args.iters * args.steps / (t_ft_max / 1e3) / 1e3 if t_ft_max > 0 else None
        if ft_tflops:
            extras["roofline_step"] = dict(bound="tensor", kernel="one fine-tune step: fused 5-iteration window (forward, loss, backward) + optimizer step, as timed in `value`",
                                           achieved=ft_tflops, peak=peaks["bf16_tflops"], unit="TFLOP/s", frac=ft_tflops / peaks["bf16_tflops"],
                                           frac_of_sustained=ft_tflops / peaks["bf16_tflops_sustained"] if peaks.get("bf16_tflops_sustained") else None,
                                           traffic=None, peak_source=peaks["source"] + " burst (sustained beside it: the step runs for 0.4 s under the power cap)",
                                           gflop_per_iteration=ITER_GFLOP, ms_per_iteration=t_ft_max / args.steps / args.iters)
        out = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"configs[1]: one-shot online fine-tune ({args.iters} SGD iters, batch 1, avg_grad_every_n={args.avg_grad_every_n}) + inference over a {args.frames}-frame 480x854 sequence, per GPU",
                       "weights": f"synthetic '{args.weights}' state_dict (fosvos_b200/synth.py), heads calibrated on a 120x214 frame",
                       "frames_per_sequence": args.frames, "finetune_iters": args.iters, "inference_batch": args.batch,
                       "sharding": "by sequence, one per rank, no collective", "cuda_graph": bool(args.graph),
                       "fuse_window": bool(args.fuse_window),
                       "fuse_window_note": "the 5 micro-iterations of an accumulation window (same weights, summed gradients) run as one batched forward/backward; every iteration's frame is computed in full; --fuse-window 0 gives the strictly sequential loop",
                       "l2": "each step touches > 126 MB of distinct activations (52 MB/layer at stage 0); no explicit flush"},
            "inference_fps": world * args.steps * args.frames / (t_inf_max / 1e3),
            "finetune_s_per_sequence": t_ft_max / 1e3 / args.steps,
            "finetune_tflops": ITER_GFLOP * args.iters * args.steps / (t_ft_max / 1e3) / 1e3 if t_ft_max > 0 else None,
            "finetune_s_per_sequence_sequential_loop": None if t_seq_ms is None else t_seq_ms / 1e3,
            "e2e": e2e, "gpu_launches": n_launch, "clocks": clocks, "roofline": roof, "roofline_side_chain": side, "roofline_loss": loss_roof,
            "cpu_baseline": cpu, "peaks": peaks,
        }
        out.update(extras)
        print(json.dumps(out), flush=True)
        if "parity" in extras and not extras["parity"]["ok"]:
            print("bench.py: PARITY FAILED (mask IoU against the oracle below 0.995): the numbers above are void", file=sys.stderr, flush=True)
            sys.exit(3)
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
    return out


def _time_ms(fn, reps=5):
    """Device time of one call: `reps` calls captured into one CUDA graph (so the host's launch latency does not
    leak into the 10-100 us kernels measured here), replayed 3 times, CUDA events on the replay stream, best replay."""
    fn(); fn(); torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
        g.replay(); torch.cuda.synchronize()
        best = float("inf")
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps)
    torch.cuda.current_stream().wait_stream(st)
    return best


def _conv_flops(conv, h, w, n):
    return 2.0 * conv.in_channels * conv.out_channels * 9 * h * w * n / 1e9


def conv_roofline(net, frames, peaks, precision):
    """Achieved TFLOP/s of the 17 conv3x3 launches of an INFERENCE forward pass exactly as the network issues them (first
    layer c8; conv1_2 pool-only; last conv of stages 1-3 with the fused pool; side_prep through the row-stacked kernel
    with the heads in its epilogue): CUDA events around exactly those launches on the launching stream, vs the measured
    dense bf16 peak."""
    from fosvos_b200 import _lib as L, ops
    from fosvos_b200.networks import _act_dtype
    n = frames.shape[0]
    dt = _act_dtype(precision)
    impl = net._impl()
    tc = impl == "tc"
    RB = L.CONV_BIAS | L.CONV_RELU
    with torch.no_grad():
        params = net._side()
        heads = ops.side_heads_views(params)
        a = ops.nchw_to_nhwc(frames, dt)
        plan = []                       # (callable, conv, h, w, kind)
        stage_convs = net._stage_convs()
        for si, convs in enumerate(stage_convs):
            if si > 0:
                a = ops.maxpool2x2(a)
            for ci, conv in enumerate(convs):
                pc = net._packed_for(conv, False)
                cp = ops.pad8(conv.out_channels)
                x = a
                last = ci == len(convs) - 1
                poolable = tc and last and si < 4 and cp >= 64 and x.shape[3] > 8 and net.fuse_pool      # (the network's own rule)
                if poolable and si == 0:
                    plan.append((lambda x=x, pc=pc, cp=cp: ops.conv3x3_pool_only(x, pc.w_fwd, pc.bias, cp, RB), conv, x.shape[1], x.shape[2], "pool-only"))
                elif poolable:
                    plan.append((lambda x=x, pc=pc, cp=cp: ops.conv3x3_pool(x, pc.w_fwd, pc.bias, cp, RB), conv, x.shape[1], x.shape[2], "conv+pool"))
                else:
                    o = torch.empty((x.shape[0], x.shape[1], x.shape[2], cp), dtype=x.dtype, device=x.device)
                    plan.append((lambda x=x, pc=pc, cp=cp, o=o: ops.conv3x3(x, pc.w_fwd, pc.bias, cp, RB, out=o, impl=impl), conv, x.shape[1], x.shape[2], "conv"))
                a = ops.conv3x3(a, pc.w_fwd, pc.bias, cp, RB, impl=impl)
            if si > 0:
                sp_conv = net.side_prep[si - 1]
                pc = net._packed_for(sp_conv, False)
                x = a
                if tc and net.side_tc and ops.side_tc_supported(x.shape[3]):
                    zs = torch.empty((x.shape[0] * x.shape[1] * x.shape[2], 2), dtype=torch.float32, device=x.device)
                    plan.append((lambda x=x, pc=pc, zs=zs, hv=heads[si - 1]: ops.conv3x3_side(x, pc.w_fwd, pc.bias, zs=zs, heads=hv, want_y=False),
                                 sp_conv, x.shape[1], x.shape[2], "side_prep row-stacked + heads"))
                else:
                    o = torch.empty((x.shape[0], x.shape[1], x.shape[2], 16), dtype=x.dtype, device=x.device)
                    plan.append((lambda x=x, pc=pc, o=o: ops.conv3x3(x, pc.w_fwd, pc.bias, 16, L.CONV_BIAS, out=o, impl=impl), sp_conv, x.shape[1], x.shape[2], "side_prep"))

        def run():
            for fn, *_ in plan:
                fn()
        ms = _time_ms(run)
        per_layer = []
        for fn, conv, h, w, kind in plan:
            t = _time_ms(fn, reps=3)
            per_layer.append(dict(cin=conv.in_channels, cout=conv.out_channels, hw=[h, w], kind=kind, ms=round(t, 4),
                                  tflops=round(_conv_flops(conv, h, w, n) / t, 1)))
    gflop = sum(_conv_flops(conv, h, w, n) for _, conv, h, w, _ in plan)
    achieved = gflop / ms                             # GFLOP / ms = TFLOP/s
    peak = peaks["bf16_tflops"]
    # DRAM bytes of the same launches from the committed ncu --set full capture (only if it was taken at this batch)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r02_conv_forward_traffic.json")
    if tc and os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("batch") == n:
            traffic, traffic_src = tj["traffic_bytes"], "profiles/r02_conv_forward_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum over the 17 launches of one batch-%d forward)" % n
    return dict(bound="tensor", kernel="conv3x3_tc_kernel + conv3x3_stack_tc_kernel (conv1_2) + conv3x3_side_tc_kernel (17 launches = one inference forward pass)" if tc else "conv3x3_simt_kernel",
                achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=traffic, traffic_source=traffic_src,
                peak_source=peaks["source"] + " burst", batch=n, gflop_per_frame=gflop / n, ms_per_forward_convs=ms, per_layer=per_layer)


def backward_rooflines(net, frames, peaks, n_window):
    """Weight-gradient and data-gradient kernels at the fine-tune window's batch (CUDA events around exactly those launches):
    17 tensor-core weight gradients (258.23 GFLOP per frame) and 16 data gradients (256.81: the first conv needs none)."""
    from fosvos_b200 import _lib as L, ops
    from fosvos_b200.networks import _act_dtype
    if net._impl() != "tc":
        return None, None
    n = n_window
    g = torch.Generator(device=frames.device).manual_seed(3)
    with torch.no_grad():
        a = ops.nchw_to_nhwc(frames[:1].expand(n, -1, -1, -1).contiguous(), torch.bfloat16)
        wg, dg = [], []
        for si, convs in enumerate(net._stage_convs()):
            if si > 0:
                a = ops.maxpool2x2(a)
            for ci, conv in enumerate(convs):
                pc = net._packed_for(conv, True)
                cp = ops.pad8(conv.out_channels)
                x = a
                a = ops.conv3x3(a, pc.w_fwd, pc.bias, cp, L.CONV_BIAS | L.CONV_RELU)
                dz = (torch.randn(a.shape, device=a.device, generator=g) * 0.01).to(torch.bfloat16)
                ws = ops.wgrad_workspace(x.shape[3], cp, a.device)
                db = torch.zeros(conv.out_channels, device=a.device)
                wg.append((lambda x=x, dz=dz, ws=ws, db=db, co=conv.out_channels: ops.conv3x3_wgrad_accumulate(x, dz, ws, db, co), conv, x.shape[1], x.shape[2]))
                if not (si == 0 and ci == 0):
                    o = torch.empty_like(x)
                    dg.append((lambda x=x, dz=dz, pc=pc, o=o: ops.conv3x3(dz, pc.w_dgrad, None, x.shape[3], L.CONV_MASK, mask=x, out=o), conv, x.shape[1], x.shape[2]))
            if si > 0:
                sp_conv = net.side_prep[si - 1]
                pc = net._packed_for(sp_conv, True)
                x = a
                dz = (torch.randn((a.shape[0], a.shape[1], a.shape[2], 16), device=a.device, generator=g) * 0.01).to(torch.bfloat16)
                ws = ops.wgrad_workspace(x.shape[3], 16, a.device)
                db = torch.zeros(16, device=a.device)
                wg.append((lambda x=x, dz=dz, ws=ws, db=db: ops.conv3x3_wgrad_accumulate(x, dz, ws, db, 16), sp_conv, x.shape[1], x.shape[2]))
                o = torch.empty_like(x)
                dg.append((lambda x=x, dz=dz, pc=pc, o=o: ops.conv3x3(dz, pc.w_dgrad, None, x.shape[3], L.CONV_MASK, mask=x, out=o), sp_conv, x.shape[1], x.shape[2]))
        out = []
        for name, plan in (("conv3x3_wgrad_tc_kernel + conv3x3_wgrad_tc_pair_kernel (17 launches = the weight gradients of one window)", wg),
                           ("conv3x3_tc_kernel / conv3x3_stack_tc_kernel, data-gradient use (16 launches = the data gradients of one window)", dg)):
            ms = _time_ms(lambda: [fn() for fn, *_ in plan], reps=3)
            gflop = sum(_conv_flops(c, h, w, n) for _, c, h, w in plan)
            per_layer = [dict(cin=c.in_channels, cout=c.out_channels, hw=[h, w], ms=round(_time_ms(fn, reps=3), 4)) for fn, c, h, w in plan]
            for d, (_, c, h, w) in zip(per_layer, plan):
                d["tflops"] = round(_conv_flops(c, h, w, n) / d["ms"], 1)
            out.append(dict(bound="tensor", kernel=name, achieved=gflop / ms, peak=peaks["bf16_tflops"], unit="TFLOP/s", frac=gflop / ms / peaks["bf16_tflops"],
                            traffic=None, peak_source=peaks["source"] + " burst", batch=n, gflop_per_frame=gflop / n, ms=ms, per_layer=per_layer))
    return out[0], out[1]


def config3_leg(sd0, frames_d, frames_h, peaks, dev, batch=32, reps=3):
    """BASELINE configs[2]: channel-pruned VGG (50 % of the filters of every stage conv: prune.l2_prune_half, the reference's
    literal bias-free surgery prune.py:490-514), batched bf16 inference, batch 32 at 480x854: frames/s, the conv roofline
    with FLOPs recomputed from the actual channel table, parity of one frame of the batch against the CPU oracle (and the
    reference's own bf16 loss on the same weights next to it)."""
    import fosvos_b200 as FB
    from fosvos_b200 import prune as P
    from oracle import osvos_oracle as O
    net = FB.OSVOS_VGG(pretrained=0)
    net.load_state_dict(sd0)
    net = net.to(dev)
    net.precision = "bf16"
    P.l2_prune_half(net, 0.5)
    psd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    fb = frames_d[:batch]
    with torch.no_grad():
        for _ in range(3):
            net.predict(fb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            _, prob, mask = net.predict(fb)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    roof = conv_roofline(net, fb, peaks, "bf16")
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = O.vgg_forward(psd, frames_h[0:1])
        ctx, prep = _oracle_on_gpu("bf16_autocast")
        with ctx():
            o16 = O.vgg_forward({k: v.to(dev) for k, v in psd.items()}, frames_d[0:1])
    pr = O.probabilities(ref[4])
    i_, u_ = O.mask_iou_counts(mask[0:1].cpu(), O.binarise(pr))
    channels = [m.out_channels for st in net._stage_convs() for m in st]
    return dict(workload=f"configs[2]: l2_prune_half(0.5) of the parent weights, batch {batch} x 480x854, bf16 (predict: ingest -> 5 maps + prob + mask)",
                channels=channels, params=int(sum(p.numel() for p in net.parameters())), frames_per_s=batch / (ms / 1e3), ms_per_batch=ms,
                roofline={k: v for k, v in roof.items() if k != "per_layer"}, per_layer=roof["per_layer"],
                parity=dict(frame=0, max_dprob=float((prob[0:1].cpu() - pr).abs().max()), iou=1.0 if u_ == 0 else i_ / u_,
                            reference_under_autocast_max_dprob=float((torch.sigmoid(o16[4].float().cpu()) - pr).abs().max())))


def side_roofline(net, frames, peaks):
    """HBM roofline of the fused side-output chain as the network runs it.  Since round 2 the two 1x1 heads leave the
    side_prep convolutions' epilogue (conv_side_tc.cu), so in inference the chain is ONE kernel -- transposed-conv
    up-sampling + crop + fuse + sigmoid + threshold -- that reads 8 B per low-res pixel (the head maps) and writes the five
    logit maps, the probabilities and the mask: 11.34 MB per 480x854 frame (SURVEY 8d's 14.61 MB counted the 16-channel
    side_prep maps, 4.36 MB, which are no longer written or read)."""
    from fosvos_b200 import ops
    n = frames.shape[0]
    fused = net._impl() == "tc" and net.side_tc
    with torch.no_grad():
        _, _, _, saved = net._run_forward(frames, save=True)
        sps, params = saved["sps"], saved["params"]
        mode = 1 if net._side_general else (2 if net._side_separable else 0)      # the path the network itself takes
        hs, ws = [int(t.shape[1]) for t in sps], [int(t.shape[2]) for t in sps]
        low = sum(h * w for h, w in zip(hs, ws))
        if fused and mode != 1:
            zs_flat, zs_views = ops.side_zs_workspace(n, hs, ws, frames.device)
            heads = ops.side_heads_views(params)
            for i in range(4):
                pc = net._packed_for(net.side_prep[i], False)
                ops.conv3x3_side(saved["stage_out"][i + 1], pc.w_fwd, pc.bias, zs=zs_views[i], heads=heads[i], want_y=False)
            ms = _time_ms(lambda: ops.side_fwd_heads_done(zs_flat, hs, ws, params, n, H, W, general=mode, want_prob=True, want_mask=True), reps=10)
            bytes_per_frame = 8 * low + 5 * H * W * 4 + H * W * 4 + H * W            # read head maps; write 5 maps + prob + mask
            kernel = "side_upsample_sep2_kernel" if mode == 2 and W % 2 == 0 else "side_upsample_sep_kernel" if mode == 2 else "side_upsample_kernel"
        else:
            ms = _time_ms(lambda: ops.side_fwd(sps, params, H, W, general=mode, want_prob=True, want_mask=True), reps=10)
            bytes_per_frame = 16 * low * sps[0].element_size() + 5 * H * W * 4 + H * W * 4 + H * W       # read sp; write 5 maps + prob + mask
            kernel = "side_heads2_kernel + side_upsample kernel"
    achieved = bytes_per_frame * n / (ms / 1e3) / 1e9
    # DRAM bytes of the launch from the committed ncu --set full capture (only if it was taken at this batch)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r02_side_traffic.json")
    if fused and os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("batch") == n:
            traffic = tj["traffic_bytes"]
    return dict(bound="hbm", kernel=kernel, achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s", frac=achieved / peaks["hbm_gbs"], traffic=traffic,
                batch=n, ms=ms, bytes_per_frame=bytes_per_frame, survey_bytes_per_frame=16 * low * 2 + 5 * H * W * 4 + H * W * 4 + H * W)


def loss_roofline(n, dev, peaks):
    """HBM roofline of the class-balanced loss as the fine-tune loop runs it: loss + gradient in one pass over
    logits and label (label counts cached with the label): read 8 B/px, write 4 B/px."""
    from fosvos_b200 import ops
    x = torch.randn((n, 1, H, W), device=dev)
    lab = (torch.rand((n, 1, H, W), device=dev) > 0.8).float()
    dx = torch.empty_like(x)
    _, stats = ops.bal_loss_fwd(x, lab, False)
    ms = _time_ms(lambda: ops.bal_loss_fwd_bwd(x, lab, False, stats, None, 1.0, out=dx), reps=10)
    b = 12 * x.numel()
    achieved = b / (ms / 1e3) / 1e9
    return dict(bound="hbm", kernel="bal_loss_fused_kernel (forward + backward, one pass)", achieved=achieved, peak=peaks["hbm_gbs"],
                unit="GB/s", frac=achieved / peaks["hbm_gbs"], traffic=None, batch=n, ms=ms, bytes_per_frame=12 * H * W)


def _oracle_gpu_variant(name):
    """(allow_tf32, autocast dtype or None, channels_last) of a stock-PyTorch variant of the reference arithmetic."""
    return {"fp32_strict": (False, None, False), "tf32": (True, None, False), "bf16_autocast": (True, torch.bfloat16, False),
            "bf16_autocast_channels_last": (True, torch.bfloat16, True)}[name]


def _oracle_on_gpu(name):
    """Context + tensor preparation for running oracle/osvos_oracle.py (the reference's own torch calls) on CUDA tensors
    under stock PyTorch / cuDNN: returns (ctx factory, prepare(tensor))."""
    import contextlib
    tf32, ac, cl = _oracle_gpu_variant(name)

    @contextlib.contextmanager
    def ctx():
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        try:
            if ac is None:
                yield
            else:
                with torch.autocast("cuda", dtype=ac):
                    yield
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old

    def prep(t):
        return t.contiguous(memory_format=torch.channels_last) if (cl and t.dim() == 4) else t
    return ctx, prep


def _stock_finetune(sd_d, x, m, n_iters, n_avg, prep, losses_out=None):
    """train_online.py:75-101 under stock PyTorch on the tensors' device: the oracle's forward and loss, autograd, and
    torch.optim.SGD with the online param groups (network_provider.py:144-159)."""
    from oracle import osvos_oracle as O
    params = {k: prep(v.clone()).requires_grad_(True) for k, v in sd_d.items()}
    groups = [dict(params=[params[k] for k in g["keys"]], lr=g["lr"], weight_decay=g["weight_decay"])
              for g in O.optimizer_groups(list(params.keys()), "online") if g["keys"]]
    opt = torch.optim.SGD(groups, lr=1e-8, momentum=0.9)
    x = prep(x)
    for it in range(n_iters):
        outs = O.vgg_forward(params, x)
        loss = O.class_balanced_cross_entropy_loss(outs[-1].float(), m, size_average=False)
        if losses_out is not None:
            losses_out.append(loss.detach())
        (loss / n_avg).backward()
        if (it + 1) % n_avg == 0:
            opt.step()
            opt.zero_grad(set_to_none=True)
    return {k: v.detach() for k, v in params.items()}


def gpu_reference(sd, frames_d, masks_d, args, ours_job_s):
    """The GPU reference to beat: the unmodified reference arithmetic -- oracle/osvos_oracle.py is the
    reference's own torch.nn.functional calls -- under stock PyTorch / cuDNN on THIS B200, in four settings.  Bounded
    sample of the same job, CUDA events, synchronize on both sides (experiment_helper.py:29-53 protocol: batch 1,
    first minibatch discarded); batch `args.batch` inference is timed too.  None of the repo's kernels run here."""
    from oracle import osvos_oracle as O
    dev = frames_d.device
    sd_d = {k: v.to(dev) for k, v in sd.items()}
    out = {}
    best = None
    for name in ("fp32_strict", "tf32", "bf16_autocast", "bf16_autocast_channels_last"):
        ctx, prep = _oracle_on_gpu(name)
        try:
            with ctx():
                w = {k: prep(v) for k, v in sd_d.items()}
                res = {}
                for bs, n_batches in ((1, 8), (args.batch, 3)):
                    xb = [prep(frames_d[(i * bs) % args.frames:(i * bs) % args.frames + bs]) for i in range(n_batches + 2)]
                    with torch.no_grad():
                        for i in range(2):
                            O.vgg_forward(w, xb[i])                 # discarded (cuDNN autotune, first minibatch)
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for i in range(n_batches):
                            o = O.vgg_forward(w, xb[2 + i])
                            torch.sigmoid(o[-1].float())            # the consumer's sigmoid (experiment_helper.py:57), on the device
                        e1.record(); torch.cuda.synchronize()
                    res[f"inference_fps_batch{bs}"] = n_batches * bs / (e0.elapsed_time(e1) / 1e3)
                x, m = frames_d[0:1], masks_d[0:1]
                _stock_finetune(sd_d, x, m, 10, 5, prep)            # warm-up: autotune, allocator
                torch.cuda.synchronize()
                n_it, best_ms = 10, float("inf")
                for _ in range(3):                                  # best of three: cuDNN's autotuned picks settle after a few calls
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); _stock_finetune(sd_d, x, m, n_it, 5, prep); e1.record(); torch.cuda.synchronize()
                    best_ms = min(best_ms, e0.elapsed_time(e1) / n_it)
                res["finetune_ms_per_iter"] = best_ms
            job = args.iters * res["finetune_ms_per_iter"] / 1e3 + args.frames / res[f"inference_fps_batch{args.batch}"]
            res["job_frames_per_s_extrapolated"] = args.frames / job
            res["job_s_extrapolated"] = job
            out[name] = res
            if best is None or job < out[best]["job_s_extrapolated"]:
                best = name
        except Exception as ex:  # a variant cuDNN cannot run is reported, not fatal
            out[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        torch.cuda.empty_cache()
    out["best"] = best
    out["sample"] = (f"per variant: 8 frames at batch 1 + 3 batches of {args.batch} (2 discarded warm-up batches each), best of 3 x 10 fine-tune "
                     f"iterations (2 optimizer steps each) after 10 warm-up iterations; job extrapolated linearly to {args.iters} iterations + {args.frames} frames")
    if best is not None and ours_job_s:
        out["ours_over_best"] = out[best]["job_s_extrapolated"] / ours_job_s
    out["what"] = "oracle/osvos_oracle.py (the reference's torch.nn.functional calls) + autograd + torch.optim.SGD on CUDA tensors; stock PyTorch %s / cuDNN %s" % (torch.__version__, torch.backends.cudnn.version())
    return out


def parity_check(net, trainer, sd0, frames_d, masks_d, frames_h, args, losses_ours):
    """Parity of the BENCHMARKED job itself (bf16 + fused window + CUDA graphs at 480x854, all `args.iters` iterations):
    the same job is run by the on-box oracle (oracle/osvos_oracle.py on CUDA, strict fp32: TF32 off -- SURVEY 8c) and the
    two results are compared: loss trajectory, weight updates of five named tensors, fused probabilities and binarised
    masks of every frame, J counts.  Frames 0 / mid / last are additionally pushed through the CPU oracle with the
    GPU-fine-tuned weights (forward parity of the inference path in isolation).  iou_min < 0.995 fails the run."""
    from oracle import osvos_oracle as O
    from fosvos_b200 import ops
    dev = frames_d.device
    F_ = args.frames
    # ours: masks / probabilities of every frame with the weights the timed job left behind
    probs, masks = [], []
    with torch.no_grad():
        for i in range(0, F_, args.batch):
            _, pr, mk = net.predict(frames_d[i:i + args.batch])
            probs.append(pr); masks.append(mk)
    prob_o, mask_o = torch.cat(probs), torch.cat(masks)
    sd_ours = {k: v.detach().clone() for k, v in net.state_dict().items()}
    # (a) CPU oracle forward with OUR fine-tuned weights on three frames
    torch.set_num_threads(os.cpu_count() or 1)
    sd_cpu = {k: v.cpu() for k, v in sd_ours.items()}
    idx = sorted({0, F_ // 2, F_ - 1})
    fwd_dprob, fwd_iou = 0.0, 1.0
    with torch.no_grad():
        for f in idx:
            ref = O.vgg_forward(sd_cpu, frames_h[f:f + 1])
            pr = O.probabilities(ref[4])
            fwd_dprob = max(fwd_dprob, float((prob_o[f:f + 1].cpu() - pr).abs().max()))
            i_, u_ = O.mask_iou_counts(mask_o[f:f + 1].cpu(), O.binarise(pr))
            fwd_iou = min(fwd_iou, 1.0 if u_ == 0 else i_ / u_)
    # (b) the whole job on the on-box oracle (strict fp32 under stock PyTorch)
    ctx, prep = _oracle_on_gpu("fp32_strict")
    sd_d = {k: v.to(dev) for k, v in sd0.items()}
    losses_ref = []
    with ctx():
        t0 = time.perf_counter()
        sd_ref = _stock_finetune(sd_d, frames_d[0:1], masks_d[0:1], args.iters, args.avg_grad_every_n, prep, losses_ref)
        torch.cuda.synchronize()
        t_ref_ft = time.perf_counter() - t0
        ref_prob = []
        with torch.no_grad():
            for i in range(0, F_, 8):
                ref_prob.append(O.probabilities(O.vgg_forward(sd_ref, frames_d[i:i + 8])[4]))
        ref_prob = torch.cat(ref_prob)
    ref_mask = (ref_prob >= 0.5).to(torch.uint8)
    counts = ops.mask_iou(mask_o.reshape(F_, -1), ref_mask.reshape(F_, -1)).cpu().double()
    ious = torch.where(counts[:, 1] > 0, counts[:, 0] / counts[:, 1].clamp(min=1), torch.ones(F_, dtype=torch.float64))
    # J of both against the synthetic ground truth (what the DAVIS scorer would report)
    gt = (masks_d >= 0.5).to(torch.uint8)
    cj_o = ops.mask_iou(mask_o.reshape(F_, -1), gt.reshape(F_, -1)).cpu().double()
    cj_r = ops.mask_iou(ref_mask.reshape(F_, -1), gt.reshape(F_, -1)).cpu().double()
    j_o = float((cj_o[:, 0] / cj_o[:, 1].clamp(min=1)).mean()); j_r = float((cj_r[:, 0] / cj_r[:, 1].clamp(min=1)).mean())
    losses_ref = [float(v) for v in torch.stack(losses_ref).cpu()]
    lo = losses_ours[:len(losses_ref)]
    loss_rel = max(abs(a - b) / max(abs(b), 1e-30) for a, b in zip(lo, losses_ref)) if lo else None
    deltas = {}
    for k in ["stages.0.0.weight", "stages.2.3.weight", "stages.4.5.weight", "side_prep.3.weight", "fuse.weight"]:
        d_o, d_r = (sd_ours[k] - sd_d[k]).double(), (sd_ref[k] - sd_d[k]).double()
        deltas[k] = dict(rel_l2=float((d_o - d_r).norm() / (d_r.norm() + 1e-300)), ref_delta_max=float(d_r.abs().max()),
                         cos=float((d_o * d_r).sum() / (d_o.norm() * d_r.norm() + 1e-300)))
    res = dict(what="whole benchmarked job (bf16, fused window, CUDA graphs, %d iterations + %d frames at 480x854) vs the on-box oracle (oracle/osvos_oracle.py, CUDA strict fp32, TF32 off) run on the same inputs" % (args.iters, F_),
               max_dprob=float((prob_o - ref_prob).abs().max()), mean_dprob=float((prob_o - ref_prob).abs().mean()),
               iou_min=float(ious.min()), iou_mean=float(ious.mean()), frames_checked=F_,
               J_mean_ours=j_o, J_mean_oracle=j_r,
               loss_first=[lo[0], losses_ref[0]] if lo else None, loss_last=[lo[-1], losses_ref[-1]] if lo else None,
               loss_max_rel_err=loss_rel, weight_update=deltas,
               forward_only_cpu_oracle=dict(frames=idx, max_dprob=fwd_dprob, iou_min=fwd_iou,
                                            what="CPU oracle forward with the GPU-fine-tuned weights vs net.predict"),
               oracle_finetune_s=t_ref_ft)
    res["mask_fg_fraction"] = [float(mask_o.float().mean()), float(ref_mask.float().mean())]
    res["loss_decreased"] = bool(losses_ref[-1] < losses_ref[0])
    # a fine-tune that collapses to empty masks would make the IoU check vacuous: require a non-degenerate result
    res["ok"] = bool(res["iou_min"] >= 0.995 and fwd_iou >= 0.995 and j_r > 0.3 and res["loss_decreased"])
    return res


def fp32_modes_leg(sd0, frames_d, frames_h, dev, batch=8):
    """The strict-parity modes next to each other (inference, batch 8 at 480x854): 'fp32' = direct fp32-FMA kernels,
    'fp32_tc' = fp32-equivalent arithmetic on the tcgen05 kernels (three bf16 terms per operand, six products, fp32
    accumulation), 'bf16x3' = two terms / three products; each with its max|dprob| against the CPU fp32 oracle on frame 0."""
    import fosvos_b200 as FB
    from oracle import osvos_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = O.vgg_forward(sd0, frames_h[0:1])
    out = {}
    for prec in ("fp32", "fp32_tc", "bf16x3", "bf16"):
        net = FB.OSVOS_VGG(pretrained=0)
        net.load_state_dict(sd0)
        net = net.to(dev)
        net.precision = prec
        fb = frames_d[:batch]
        with torch.no_grad():
            outs, _, _ = net.predict(fb)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 2 if prec == "fp32" else 4
            e0.record()
            for _ in range(reps):
                net.predict(fb)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        err = max(float((torch.sigmoid(o[0:1].cpu()) - torch.sigmoid(r)).abs().max()) for o, r in zip(outs, ref))
        out[prec] = dict(frames_per_s=batch / (ms / 1e3), ms_per_batch=ms, max_dprob_vs_cpu_fp32_oracle=err)
        del net
        torch.cuda.empty_cache()
    out["workload"] = f"net.predict on {batch} x 480x854 frames (ingest -> 5 maps + prob + mask), parent weights"
    return out


def config4_leg(args, rank, world, dev, sd_dev, trainer, net, n_sequences=20):
    """BASELINE configs[3]: the DAVIS-2016-val-shaped pipeline -- 20 synthetic sequences x `args.frames` frames, per-sequence
    fine-tune + inference, sequence i on rank i % world (train_online.py:184-186), no collective.  Every rank works through
    its sequences back to back on its resident trainer; wall = max over ranks of the device time."""
    from fosvos_b200 import sharding
    from fosvos_b200.online import sequences_for_rank
    mine = sequences_for_rank(list(range(n_sequences)), rank, world)
    data = [make_sequence_gpu(q, args.frames) for q in mine]
    data = [(f.to(dev), m.to(dev)) for f, m in data]
    sharding.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for frames, masks in data:
        trainer.reset(sd_dev)
        trainer.set_frame(frames[0:1], masks[0:1])
        trainer.run(args.iters)
        for i in range(0, args.frames, args.batch):
            net.predict(frames[i:i + args.batch])
    e1.record(); torch.cuda.synchronize(); sharding.barrier()
    t = e0.elapsed_time(e1) / 1e3
    t_max, t_sum = sharding.max_over_ranks(t), sharding.sum_over_ranks(t)
    t_min = -sharding.max_over_ranks(-t)
    per_rank = [len(sequences_for_rank(list(range(n_sequences)), r, world)) for r in range(world)]
    return dict(workload=f"configs[3]: {n_sequences} sequences x {args.frames} frames, fine-tune ({args.iters} it) + inference per sequence, sequence i -> rank i % {world}",
                wall_s=t_max, frames_per_s=n_sequences * args.frames / t_max, s_per_sequence=t_sum / n_sequences,
                sequences_per_rank=per_rank, ideal_speedup=n_sequences / max(per_rank), rank_time_s_min_max=[t_min, t_max],
                tail_imbalance=t_max / max(t_min, 1e-9), collective="none",
                note="the optimizer steps of a sequence are sequential, so a tail sequence cannot be spread over the ranks that finished; splitting "
                     "only its inference (4 % of a sequence job) over idle ranks would need a 61 MB weight broadcast for a < 1 % gain")


def config5_leg(args, rank, world, dev, sd0, frames_d, masks_d, windows=6):
    """BASELINE configs[4]: offline parent training, data parallel -- every rank runs 5 micro-iterations (one fused window) on
    its own frames with the 5-map deep-supervision loss (train_offline.py:84-88) and the offline param groups
    (network_provider.py:98-125); gradients are summed over the ranks (61 MB fp32 per optimizer step) before the fused SGD
    step.  Three schedules are timed on the device (max over ranks): no exchange (the floor), ONE flat all-reduce after the
    backward pass, and per-stage buckets all-reduced on a side stream while the backward pass continues."""
    import fosvos_b200 as FB
    from fosvos_b200 import sharding
    from fosvos_b200.online import OnlineTrainer
    n_local = args.avg_grad_every_n
    res = {}
    # same optimizer-step kernels in all three schedules (the single-GPU trainer would otherwise fuse fold + SGD + repack, which
    # the data-parallel ones cannot: the all-reduce sits between the fold and the SGD step)
    fused_step_env = os.environ.get("FOSVOS_FUSED_STEP")
    os.environ["FOSVOS_FUSED_STEP"] = "0"
    for name, dp, overlap in (("no_exchange", False, False), ("flat_allreduce", True, False), ("bucketed_overlapped", True, True)):
        net = FB.OSVOS_VGG(pretrained=0)
        net.load_state_dict(sd0)
        net = net.to(dev)
        net.precision = args.precision
        tr = OnlineTrainer(net, H, W, n_local * (world if dp else 1), FB.get_optimizer_offline(net), use_graph=True, deep_supervision=0.5,
                           data_parallel=dp, world_size=world, fuse_window=True, overlap_allreduce=overlap)
        idx = [(rank * n_local + j) % frames_d.shape[0] for j in range(n_local)]
        tr.set_frames(frames_d[idx], masks_d[idx])
        tr.run(2 * n_local)                                   # warm-up: capture, NCCL channels
        sharding.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); tr.run(windows * n_local); e1.record(); torch.cuda.synchronize(); sharding.barrier()
        ms = sharding.max_over_ranks(e0.elapsed_time(e1)) / windows
        res[name] = dict(ms_per_optimizer_step=ms, micro_iterations_per_s=world * n_local / (ms / 1e3))
        if overlap:
            res[name]["window_captured_in_cuda_graph"] = tr._overlap_graph is not None
            if tr.overlap_capture_error:
                res[name]["capture_error"] = tr.overlap_capture_error
        del tr, net
        torch.cuda.empty_cache()
    if fused_step_env is None:
        os.environ.pop("FOSVOS_FUSED_STEP", None)
    else:
        os.environ["FOSVOS_FUSED_STEP"] = fused_step_env
    base = res["no_exchange"]["ms_per_optimizer_step"]
    for k in ("flat_allreduce", "bucketed_overlapped"):
        res[k]["exposed_allreduce_ms"] = res[k]["ms_per_optimizer_step"] - base
        res[k]["exposed_allreduce_frac_of_step"] = (res[k]["ms_per_optimizer_step"] - base) / res[k]["ms_per_optimizer_step"]
    res["workload"] = (f"configs[4]: data-parallel offline parent training at 480x854, deep supervision (5 losses), offline param groups, "
                       f"{n_local} micro-iterations per rank and optimizer step (global accumulation {n_local * world}), NCCL all-reduce of the fp32 gradients")
    res["allreduce_bytes_per_step"] = 4 * sum(p.numel() for n_, p in FB.OSVOS_VGG(pretrained=0).named_parameters() if not n_.startswith("upscale"))
    res["scaling"] = "weak (per-rank work fixed)"
    return res


def cpu_baseline(sd, frames, masks, args, forward_frames=2, ft_iters=1):
    """The oracle port of the reference path on the host cores, bounded sample, extrapolated."""
    from oracle import osvos_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x, m = frames[0:1], masks[0:1]
    with torch.no_grad():
        O.vgg_forward(sd, x)                                       # warm-up
        t0 = time.perf_counter()
        for f in range(forward_frames):
            O.vgg_forward(sd, frames[f:f + 1])
        t_fwd = (time.perf_counter() - t0) / forward_frames
    t0 = time.perf_counter()
    O.finetune(sd, x, m, ft_iters, 1)
    t_it = (time.perf_counter() - t0) / ft_iters
    job = args.iters * t_it + args.frames * t_fwd
    return dict(value=args.frames / job, unit="frames/s", cores=cores, kind="port",
                sample=f"{forward_frames} forward frames ({t_fwd:.3f} s each) + {ft_iters} fine-tune iteration(s) ({t_it:.3f} s each) at 1x3x480x854 fp32, "
                       f"extrapolated linearly to {args.iters} iterations + {args.frames} frames",
                forward_s_per_frame=t_fwd, finetune_s_per_iter=t_it)


def run_reference(args):
    """--impl reference: the reference algorithm (oracle port; the Python reference itself cannot
    travel to the GPU box) on the host cores.  Each step is a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fosvos_b200 import synth
    from oracle import osvos_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    xs, ms = synth.make_frame(0, 0, 120, 214)
    sd = synth.calibrate(synth.make_state_dict(0, "parent"), O.vgg_forward, xs, mask=ms)
    x, m = synth.make_frame(0, 0, H, W)
    def step():
        t0 = time.perf_counter()
        O.finetune(sd, x, m, 1, 1)
        t_it = time.perf_counter() - t0
        t0 = time.perf_counter()
        with torch.no_grad():
            O.vgg_forward(sd, x)
        return t_it, time.perf_counter() - t0
    for _ in range(min(args.warmup, 1)):
        step()
    tis, tfs = zip(*[step() for _ in range(args.steps)])
    t_it, t_fwd = sum(tis) / len(tis), sum(tfs) / len(tfs)
    job = args.iters * t_it + args.frames * t_fwd
    v = args.frames / job
    sample = f"per step: 1 fine-tune iteration + 1 forward frame at 1x3x480x854 fp32 on {cores} threads, extrapolated to {args.iters} iterations + {args.frames} frames"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": job * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1]: one-shot online fine-tune ({args.iters} SGD iters) + inference over a {args.frames}-frame 480x854 sequence (bounded sample, extrapolated)"},
        "cpu_baseline": dict(value=v, unit="frames/s", cores=cores, kind="port", sample=sample, forward_s_per_frame=t_fwd, finetune_s_per_iter=t_it),
        "e2e": dict(value=v, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--iters", type=int, default=500)
    ap.add_argument("--avg-grad-every-n", type=int, default=5)
    ap.add_argument("--frames", type=int, default=80)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--fuse-window", type=int, default=1,
                    help="run the avg_grad_every_n micro-iterations between two optimizer steps as one batched pass (same gradients)")
    ap.add_argument("--weights", default="parent", choices=["parent", "structured"],
                    help="synthetic parent network (synth.make_state_dict kind); 'structured' is the round-1 set on which the lr=1e-8 fine-tune diverges")
    ap.add_argument("--fp32-modes", type=int, default=1, help="N=1: inference throughput / error of the strict modes (fp32 direct, fp32_tc, bf16x3)")
    ap.add_argument("--config3", type=int, default=1, help="N=1: BASELINE configs[2] line (pruned 50 %, batch 32, bf16)")
    ap.add_argument("--config4", type=int, default=1, help="BASELINE configs[3] leg: 20 sequences sharded by sequence over the ranks")
    ap.add_argument("--config5", type=int, default=1, help="N>1: BASELINE configs[4] leg: data-parallel offline step with the gradient all-reduce")
    ap.add_argument("--parity", type=int, default=1, help="N=1: run the benchmarked job on the on-box oracle too and compare (adds ~1 min)")
    ap.add_argument("--gpu-reference", type=int, default=1, help="N=1: time the reference arithmetic under stock PyTorch/cuDNN on this GPU")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
