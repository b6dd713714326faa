#!/usr/bin/env python
"""What does bf16 cost the REFERENCE itself?  

For every parity case of tests/test_gpu_network.py (and the two 480x854 cases) this runs, on the B200:
  ref32      oracle/osvos_oracle.py on CUDA, strict fp32 (TF32 off)          -- the on-box oracle
  ref_bf16   the same code under torch.autocast(bfloat16): stock PyTorch's own bf16 (cuDNN bf16 convs, fp32 accumulation,
             every layer output rounded to bf16 -- the same number format our tcgen05 path computes in)
  ref_bf16_backbone   autocast only around the 3x3 convs / pools, side chain in fp32 (our path keeps the side chain in fp32)
  ours_bf16  fosvos_b200 precision='bf16'
  ours_fp32  fosvos_b200 precision='fp32'
and reports max|dprob| of the five maps against the CPU fp32 oracle, mask IoU, and for the backward the norm-wise
relative error of the gradients.  Output: one JSON document (stdout and --out).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import fosvos_b200 as FB  # noqa: E402
from fosvos_b200 import synth  # noqa: E402
from oracle import osvos_oracle as O  # noqa: E402

DEV = "cuda"


def strict():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def forward_backbone_autocast(sd, x):
    """O.vgg_forward with autocast(bf16) around the backbone and side_prep convs only; the side chain runs in fp32."""
    idxs = O.stage_conv_indices()
    H, W = x.shape[-2:]
    side, side_out = [], []
    for si in range(5):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if si > 0:
                x = F.max_pool2d(x, 2, 2, ceil_mode=True)
            for mi in idxs[si]:
                x = F.relu(F.conv2d(x, sd[f"stages.{si}.{mi}.weight"], sd.get(f"stages.{si}.{mi}.bias"), padding=1))
            if si > 0:
                sp = F.conv2d(x, sd[f"side_prep.{si - 1}.weight"], sd.get(f"side_prep.{si - 1}.bias"), padding=1)
        if si > 0:
            i = si - 1
            sp = sp.float()
            side.append(O.center_crop(F.conv_transpose2d(sp, sd[f"upscale.{i}.weight"], stride=2 ** si), H, W))
            sc = F.conv2d(sp, sd[f"score_dsn.{i}.weight"], sd[f"score_dsn.{i}.bias"])
            side_out.append(O.center_crop(F.conv_transpose2d(sc, sd[f"upscale_.{i}.weight"], stride=2 ** si), H, W))
    side_out.append(F.conv2d(torch.cat(side, 1), sd["fuse.weight"], sd["fuse.bias"]))
    return side_out


def dprob(outs, ref):
    return [float((torch.sigmoid(o.float().cpu()) - torch.sigmoid(r)).abs().max()) for o, r in zip(outs, ref)]


def iou(a, b):
    i, u = O.mask_iou_counts(a, b)
    return 1.0 if u == 0 else i / u


def ours(sd, x, precision):
    net = FB.OSVOS_VGG(pretrained=0)
    # pruned state dicts: rebuild the narrower convs first (prune.py:490-514 surgery)
    for si, idx in enumerate(O.stage_conv_indices()):
        for mi in idx:
            w = sd[f"stages.{si}.{mi}.weight"]
            if tuple(w.shape) != tuple(net.stages[si][mi].weight.shape) or f"stages.{si}.{mi}.bias" not in sd:
                net.stages[si][mi] = torch.nn.Conv2d(w.shape[1], w.shape[0], 3, padding=1, bias=f"stages.{si}.{mi}.bias" in sd)
        if si > 0 and (tuple(sd[f"side_prep.{si - 1}.weight"].shape) != tuple(net.side_prep[si - 1].weight.shape) or f"side_prep.{si - 1}.bias" not in sd):
            net.side_prep[si - 1] = torch.nn.Conv2d(sd[f"side_prep.{si - 1}.weight"].shape[1], 16, 3, padding=1, bias=f"side_prep.{si - 1}.bias" in sd)
    net.load_state_dict(sd)
    net = net.to(DEV)
    net.precision = precision
    return net


def forward_case(name, sd, x):
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = O.vgg_forward(sd, x)                      # CPU fp32: the strict oracle
    sd_d = {k: v.to(DEV) for k, v in sd.items()}
    xd = x.to(DEV)
    strict()
    res = {"case": name, "shape": list(x.shape), "logit_std": float(ref[4].std()), "frac_within_0.1_of_threshold": float((torch.sigmoid(ref[4]) - 0.5).abs().lt(0.1).float().mean())}
    ref_mask = O.binarise(O.probabilities(ref[4]))
    with torch.no_grad():
        o = O.vgg_forward(sd_d, xd)
        res["ref32_cuda_strict"] = dict(max_dprob=dprob(o, ref))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            o = O.vgg_forward(sd_d, xd)
        res["ref_bf16_autocast"] = dict(max_dprob=dprob(o, ref), iou=iou(O.binarise(O.probabilities(o[4].float().cpu())), ref_mask))
        xc = xd.contiguous(memory_format=torch.channels_last)
        sdc = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd_d.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16):
            o = O.vgg_forward(sdc, xc)
        res["ref_bf16_autocast_channels_last"] = dict(max_dprob=dprob(o, ref), iou=iou(O.binarise(O.probabilities(o[4].float().cpu())), ref_mask))
        o = forward_backbone_autocast(sd_d, xd)
        res["ref_bf16_backbone_only"] = dict(max_dprob=dprob(o, ref), iou=iou(O.binarise(O.probabilities(o[4].float().cpu())), ref_mask))
        torch.backends.cudnn.allow_tf32 = True
        o = O.vgg_forward(sd_d, xd)
        res["ref_tf32"] = dict(max_dprob=dprob(o, ref))
        strict()
        for prec in ("bf16", "fp32"):
            net = ours(sd, x, prec)
            outs, prob, mask = net.predict(xd)
            res[f"ours_{prec}"] = dict(max_dprob=dprob(outs, ref), iou=iou(mask.cpu(), ref_mask))
    return res


def backward_case(name, sd, x, m):
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    O.class_balanced_cross_entropy_loss(O.vgg_forward(params, x)[-1], m, size_average=False).backward()
    gref = {k: p.grad for k, p in params.items() if p.grad is not None}
    keys = [k for k in gref if not k.startswith("upscale") and not k.startswith("score_dsn")]
    strict()
    out = {"case": name}

    def rel(g):
        r = {k: float((g[k].float().cpu() - gref[k]).norm() / (gref[k].norm() + 1e-12)) for k in keys}
        return dict(worst=max(r.values()), worst_key=max(r, key=r.get), backbone_weights_median=sorted(v for k, v in r.items() if k.startswith("stages") and k.endswith("weight"))[6],
                    heads=max(v for k, v in r.items() if k.startswith(("fuse", "side_prep"))), per_tensor=r)

    pd = {k: v.to(DEV).requires_grad_(True) for k, v in sd.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o = O.vgg_forward(pd, x.to(DEV))
    O.class_balanced_cross_entropy_loss(o[-1].float(), m.to(DEV), size_average=False).backward()
    out["ref_bf16_autocast"] = rel({k: pd[k].grad for k in keys})
    pd = {k: v.to(DEV).requires_grad_(True) for k, v in sd.items()}
    o = forward_backbone_autocast(pd, x.to(DEV))
    O.class_balanced_cross_entropy_loss(o[-1].float(), m.to(DEV), size_average=False).backward()
    out["ref_bf16_backbone_only"] = rel({k: pd[k].grad for k in keys})
    for prec in ("bf16", "fp32"):
        net = ours(sd, x, prec)
        outs = net.forward(x.to(DEV))
        FB.class_balanced_cross_entropy_loss(outs[-1], m.to(DEV), size_average=False).backward()
        out[f"ours_{prec}"] = rel({k: p.grad for k, p in net.named_parameters() if k in keys})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bf16_yardstick.json"))
    ap.add_argument("--full", type=int, default=1, help="include the 480x854 cases")
    args = ap.parse_args()
    doc = {"forward": [], "backward": [], "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "gpu": torch.cuda.get_device_name(0)}
    small = [("fwd_48x72_random", 1, 0, 48, 72, "random", True), ("fwd_45x70_random", 2, 0, 45, 70, "random", False),
             ("fwd_64x96_structured", 3, 0, 64, 96, "structured", False)]
    gold = os.path.join(ROOT, "tests", "golden")
    for name, _, _, _, _, kind, _ in small:
        fix = torch.load(os.path.join(gold, name + ".pt"), weights_only=False)
        x, m = synth.make_frame(fix["seq"], fix["frame"], fix["H"], fix["W"], noise=fix.get("noise", False))
        sd = synth.calibrate(synth.make_state_dict(0, kind), O.vgg_forward, x, mask=m if kind == "structured" else None)
        doc["forward"].append(forward_case(name, sd, x))
        if kind == "random":
            doc["backward"].append(backward_case(name, sd, x, m))
        print(json.dumps(doc["forward"][-1]), flush=True)
    if args.full:
        xs, ms = synth.make_frame(0, 0, 120, 214)
        x, m = synth.make_frame(0, 0, 480, 854)
        sd = synth.calibrate(synth.make_state_dict(0, "structured"), O.vgg_forward, xs, mask=ms)
        doc["forward"].append(forward_case("480x854_structured", sd, x))
        print(json.dumps(doc["forward"][-1]), flush=True)
        sdp = synth.calibrate(synth.make_state_dict(0, "parent"), O.vgg_forward, xs, mask=ms)
        doc["forward"].append(forward_case("480x854_parent (bench weights)", sdp, x))
        print(json.dumps(doc["forward"][-1]), flush=True)
        sdr = synth.calibrate(synth.make_state_dict(0, "random"), O.vgg_forward, xs)
        doc["forward"].append(forward_case("480x854_random", sdr, x))
        print(json.dumps(doc["forward"][-1]), flush=True)
        psd = synth.calibrate(synth.prune_state_dict(synth.make_state_dict(0, "structured"), 0.5), O.vgg_forward, xs, mask=ms)
        doc["forward"].append(forward_case("480x854_pruned50_structured_recalibrated", psd, x))
        print(json.dumps(doc["forward"][-1]), flush=True)
        # config 3 as the test builds it: prune the CALIBRATED dense net (no re-calibration)
        from fosvos_b200 import prune as P
        net = ours(sd, x, "bf16")
        P.l2_prune_half(net, 0.5)
        psd2 = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        doc["forward"].append(forward_case("480x854_pruned50_of_calibrated_dense (test_config2 weights)", psd2, x))
        print(json.dumps(doc["forward"][-1]), flush=True)
    for b in doc["backward"]:
        print(json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != "per_tensor"}) for k, v in b.items()}), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(doc, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
