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


def calibrated_state_dict():
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
    return synth.calibrate(synth.make_state_dict(0, "structured"), fwd, xs, mask=ms)


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
    sd0 = calibrated_state_dict()
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
    if rank == 0:
        # ---- roofline of the dominant kernel family (3x3 conv implicit GEMM), measured live -------
        roof = conv_roofline(new_net(), frames_d[:args.batch], peaks, args.precision)
        side = side_roofline(new_net(), frames_d[:args.batch], peaks)
        loss_roof = loss_roofline(args.batch, dev, peaks)
        cpu = cpu_baseline(sd0, frames_h, masks_h, args)
        out = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"configs[1]: one-shot online fine-tune ({args.iters} SGD iters, batch 1, avg_grad_every_n={args.avg_grad_every_n}) + inference over a {args.frames}-frame 480x854 sequence, per GPU",
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
        print(json.dumps(out), flush=True)
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


def conv_roofline(net, frames, peaks, precision):
    """Achieved TFLOP/s of the 17 conv3x3 launches of a forward pass (CUDA events around exactly
    those launches, current stream), vs the measured dense bf16 peak."""
    from fosvos_b200 import _lib as L, ops
    from fosvos_b200.networks import _act_dtype
    n = frames.shape[0]
    dt = _act_dtype(precision)
    impl = net._impl()
    with torch.no_grad():
        a = ops.nchw_to_nhwc(frames, dt)
        plan = []
        for si, convs in enumerate(net._stage_convs()):
            if si > 0:
                a = ops.maxpool2x2(a)
            for conv in convs:
                pc = net._packed_for(conv, False)
                plan.append((a, pc, ops.pad8(conv.out_channels), L.CONV_BIAS | L.CONV_RELU, conv))
                a = ops.conv3x3(a, pc.w_fwd, pc.bias, ops.pad8(conv.out_channels), L.CONV_BIAS | L.CONV_RELU, impl=impl)
            if si > 0:
                pc = net._packed_for(net.side_prep[si - 1], False)
                plan.append((a, pc, 16, L.CONV_BIAS, net.side_prep[si - 1]))
        outs = [torch.empty((x.shape[0], x.shape[1], x.shape[2], cp), dtype=x.dtype, device=x.device) for x, _, cp, _, _ in plan]
        def run():
            for (x, pc, cp, fl, _), o in zip(plan, outs):
                ops.conv3x3(x, pc.w_fwd, pc.bias, cp, fl, out=o, impl=impl)
        ms = _time_ms(run)
        per_layer = []
        for (x, pc, cp, fl, conv), o in zip(plan, outs):
            t = _time_ms(lambda: ops.conv3x3(x, pc.w_fwd, pc.bias, cp, fl, out=o, impl=impl), reps=3)
            gf = 2.0 * conv.in_channels * conv.out_channels * 9 * x.shape[1] * x.shape[2] * n / 1e9
            per_layer.append(dict(cin=conv.in_channels, cout=conv.out_channels, hw=[x.shape[1], x.shape[2]], ms=round(t, 4), tflops=round(gf / t, 1)))
    achieved = FWD_GFLOP * n / ms                     # GFLOP / ms = TFLOP/s
    peak = peaks["bf16_tflops"]
    # DRAM bytes of the same 17 launches from the committed ncu --set full capture (only if it was taken at this batch)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r01h_conv_forward_traffic.json")
    if impl == "tc" and os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("batch") == n:
            traffic, traffic_src = tj["traffic_bytes"], "profiles/r01h_conv_forward_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum over the 17 launches of one batch-%d forward)" % n
    return dict(bound="tensor", kernel="conv3x3_tc_kernel (17 launches = one forward pass)" if impl == "tc" else "conv3x3_simt_kernel",
                achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=traffic, traffic_source=traffic_src,
                peak_source=peaks["source"] + " burst", batch=n, ms_per_forward_convs=ms, per_layer=per_layer)


def side_roofline(net, frames, peaks):
    """HBM roofline of the fused side-output chain (heads + upsample/fuse/sigmoid/threshold)."""
    from fosvos_b200 import ops
    n = frames.shape[0]
    with torch.no_grad():
        _, _, _, saved = net._run_forward(frames, save=True)
        sps, params = saved["sps"], saved["params"]
        mode = 1 if net._side_general else (2 if net._side_separable else 0)      # the path the network itself takes
        ms = _time_ms(lambda: ops.side_fwd(sps, params, H, W, general=mode, want_prob=True, want_mask=True), reps=10)
    esz = sps[0].element_size()
    low = sum(t.shape[1] * t.shape[2] for t in sps)
    bytes_per_frame = 16 * low * esz + 5 * H * W * 4 + H * W * 4 + H * W       # read sp; write 5 maps + prob + mask
    achieved = bytes_per_frame * n / (ms / 1e3) / 1e9
    # DRAM bytes of the two launches from the committed ncu --set full capture (only if it was taken at this batch)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01i_side_traffic.json")
    if mode == 2 and W % 2 == 0 and esz == 2 and os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("batch") == n:
            traffic = tj["traffic_bytes"]
    return dict(bound="hbm", kernel="side_heads2_kernel + " + ("side_upsample_sep2_kernel" if mode == 2 and W % 2 == 0 else "side_upsample_sep_kernel" if mode == 2 else "side_upsample_kernel"), achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s",
                frac=achieved / peaks["hbm_gbs"], traffic=traffic, batch=n, ms=ms, bytes_per_frame=bytes_per_frame)


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
    sd = synth.calibrate(synth.make_state_dict(0, "structured"), O.vgg_forward, xs, mask=ms)
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
