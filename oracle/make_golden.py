"""Pin the oracle against the LIVE reference and mint tests/golden/*.pt.

Run in the build container (needs /root/reference):
    python -m oracle.make_golden

For every case it (1) runs the reference's own ``OSVOS_VGG`` /
``class_balanced_cross_entropy_loss`` / ``VGGOnlineProvider.get_optimizer``
(imported unmodified through oracle/_refshim.py), (2) asserts that
oracle/osvos_oracle.py reproduces the result, and (3) stores the reference's
outputs as a small fixture.  Inputs and weights are regenerated from seeds by
``fosvos_b200.synth`` (checksums are stored to detect RNG drift).
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from fosvos_b200 import synth
from oracle import _refshim, osvos_oracle as O, osvos_numpy as ON

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (name, H, W, weight kind, noise input)
CASES = [
    ("fwd_48x72_random", 48, 72, "random", True),
    ("fwd_45x70_random", 45, 70, "random", False),   # odd H -> odd crop, ceil pooling on both axes
    ("fwd_64x96_structured", 64, 96, "structured", False),
]


def checksum(sd):
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in sd.items()}


def calibrated_sd(kind, x, mask):
    sd = synth.make_state_dict(0, kind)
    return synth.calibrate(sd, O.vgg_forward, x, mask=mask if kind == "structured" else None)


def ref_net(ns, sd):
    net = ns.OSVOS_VGG(pretrained=0)
    net.load_state_dict(sd)
    return net


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ns = _refshim.load()
    os.makedirs(GOLD, exist_ok=True)
    L = ns.layers

    # ---- known-answer tests of the small layer functions -------------------
    kat = {}
    for k in (4, 8, 16, 32, 3, 5):
        ref = L.upsample_filt(k)
        assert np.array_equal(ref, O.upsample_filt(k)), k
        kat[f"upsample_filt_{k}"] = torch.from_numpy(ref.copy())
    assert np.allclose(L.upsample_filt(4), np.outer([.25, .75, .75, .25], [.25, .75, .75, .25]))
    crops = {}
    for (ih, iw, h, w) in [(483, 857, 480, 854), (482, 856, 480, 854), (496, 880, 480, 854), (48, 74, 45, 70), (10, 10, 10, 10), (13, 9, 8, 4)]:
        x = torch.arange(ih * iw, dtype=torch.float32).view(1, 1, ih, iw)
        ref = L.center_crop(x, h, w)
        mine = O.center_crop(x, h, w)
        assert torch.equal(ref, mine)
        crops[f"{ih}x{iw}->{h}x{w}"] = (int(ref[0, 0, 0, 0].item()) // iw, int(ref[0, 0, 0, 0].item()) % iw)
    kat["center_crop_origin"] = crops
    lay = torch.nn.ConvTranspose2d(16, 16, 8, stride=4, bias=False)
    lay.weight.data.zero_()
    assert torch.equal(L.interp_surgery(lay), O.interp_surgery_weight(16, 8))
    # loss on a 4x4 example + a bigger random one, both reductions
    g = torch.Generator().manual_seed(7)
    for name, shape in (("4x4", (1, 1, 4, 4)), ("2x1x37x53", (2, 1, 37, 53))):
        out = (torch.randn(shape, generator=g) * 4).requires_grad_(True)
        lab = (torch.rand(shape, generator=g) > 0.7).float()
        for sa in (True, False):
            ref = L.class_balanced_cross_entropy_loss(out, lab, size_average=sa)
            mine = O.class_balanced_cross_entropy_loss(out, lab, size_average=sa)
            assert torch.equal(ref, mine)
            (gref,) = torch.autograd.grad(ref, out)
            gmine = O.class_balanced_cross_entropy_grad(out.detach(), lab, sa)
            assert torch.allclose(gref, gmine, rtol=1e-5, atol=1e-7), (gref - gmine).abs().max()
            kat[f"loss_{name}_sa{int(sa)}"] = dict(output=out.detach().clone(), label=lab, loss=ref.detach().clone(), grad=gref.clone())
    # teacher-logit labels (mimic.py:213 passes logits as `label`)
    out = torch.randn(1, 1, 9, 11, generator=g).requires_grad_(True)
    lab = torch.randn(1, 1, 9, 11, generator=g) * 2
    ref = L.class_balanced_cross_entropy_loss(out, lab)
    assert torch.equal(ref, O.class_balanced_cross_entropy_loss(out, lab))
    kat["loss_teacher_logits"] = dict(output=out.detach().clone(), label=lab, loss=ref.detach().clone(),
                                      grad=torch.autograd.grad(ref, out)[0])
    # optimizer group tables from the live providers
    for mode, prov_cls, sett_cls, extra in (("online", ns.VGGOnlineProvider, ns.OnlineSettings, (None, None)),
                                            ("offline", ns.VGGOfflineProvider, ns.OfflineSettings, (False,))):
        sett = sett_cls(*([True, True, 0, 10, 5, 10, False, 10, 1, 1, False, False, None, False] + list(extra)))
        prov = prov_cls("vgg16", (None, None), sett)
        prov.init_network(pretrained=0)
        opt = prov.get_optimizer()
        id2k = {id(p): k for k, p in prov.network.named_parameters()}
        table = [dict(keys=[id2k[id(p)] for p in grp["params"]], lr=grp["lr"], weight_decay=grp["weight_decay"],
                      momentum=grp["momentum"]) for grp in opt.param_groups]
        mine = O.optimizer_groups(list(id2k.values()), mode)
        assert len(mine) == len(table)
        for a, b in zip(mine, table):
            assert a["keys"] == b["keys"] and abs(a["lr"] - b["lr"]) <= 1e-20 and a["weight_decay"] == b["weight_decay"], (a, b)
        kat[f"optimizer_groups_{mode}"] = table
    kat["state_dict_spec"] = [(k, tuple(v.shape)) for k, v in ns.OSVOS_VGG(pretrained=0).state_dict().items()]
    assert kat["state_dict_spec"] == O.state_dict_spec()
    torch.save(kat, os.path.join(GOLD, "kat.pt"))
    print("kat ok")

    # ---- end-to-end forward / loss / backward / fine-tune fixtures ----------
    for name, H, W, kind, noise in CASES:
        x, m = synth.make_frame(3, 0, H, W, noise=noise)
        sd = calibrated_sd(kind, x, m)
        net = ref_net(ns, sd)
        outs_ref = net.forward(x)
        outs_mine = O.vgg_forward(sd, x)
        for a, b in zip(outs_ref, outs_mine):
            assert torch.equal(a, b), (a - b).abs().max()
        # numpy second opinion (float64)
        outs_np = ON.vgg_forward({k: v.numpy() for k, v in sd.items()}, x.numpy())
        err_np = max(float(np.abs(a.detach().numpy() - b).max()) for a, b in zip(outs_ref, outs_np))
        assert err_np < 5e-4, err_np
        loss = L.class_balanced_cross_entropy_loss(outs_ref[-1], m, size_average=False)
        net.zero_grad()
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
        keep = ["fuse.weight", "fuse.bias", "side_prep.0.bias", "side_prep.3.weight", "stages.4.5.bias",
                "stages.0.0.weight", "stages.0.0.bias", "stages.2.3.bias", "stages.1.1.weight"]
        fix = dict(H=H, W=W, kind=kind, noise=noise, seq=3, frame=0,
                   sd_checksum=checksum(sd), x_checksum=float(x.double().sum()),
                   outs=[o.detach().clone() for o in outs_ref], loss=loss.detach().clone(),
                   grads={k: grads[k] for k in keep}, numpy_fp64_max_abs=err_np,
                   grad_norms={k: float(v.norm()) for k, v in grads.items()})
        # 3 fine-tune iterations, step every iteration, with the reference's own SGD groups
        if name == "fwd_48x72_random":
            sett = ns.OnlineSettings(*([True, True, 0, 10, 1, 10, False, 10, 1, 1, False, False, None, False, None, None]))
            prov = ns.VGGOnlineProvider("vgg16", (None, None), sett)
            prov.init_network(pretrained=0)
            prov.network.load_state_dict(sd)
            lr_scale = 1.0  # the reference default lr (1e-8): the loss falls monotonically on this case
            opt = prov.get_optimizer(learning_rate=1e-8 * lr_scale)
            ref_losses = []
            n_it, n_avg = 6, 2
            for it in range(n_it):                                   # train_online.py:75-101
                o = prov.network.forward(x)
                ls = L.class_balanced_cross_entropy_loss(o[-1], m, size_average=False)
                ref_losses.append(float(ls.item()))
                ls = ls / n_avg
                ls.backward()
                if (it + 1) % n_avg == 0:
                    opt.step()
                    opt.zero_grad()
            new_sd = {k: v.detach().clone() for k, v in prov.network.state_dict().items()}
            mine_sd, mine_losses = O.finetune(sd, x, m, n_it, n_avg, learning_rate=1e-8 * lr_scale)
            for k in new_sd:
                assert torch.allclose(new_sd[k], mine_sd[k], rtol=0, atol=1e-7 * max(1.0, float(new_sd[k].abs().max()))), k
            assert np.allclose(ref_losses, mine_losses, rtol=1e-6)
            fix["finetune"] = dict(n_iters=n_it, avg_grad_every_n=n_avg, learning_rate=1e-8 * lr_scale,
                                   losses=ref_losses,
                                   deltas={k: (new_sd[k] - sd[k]).clone() for k in keep},
                                   fused_after=prov.network.forward(x)[-1].detach().clone())
        torch.save(fix, os.path.join(GOLD, name + ".pt"))
        print(name, "ok; numpy fp64 max-abs", err_np, "fused std", float(outs_ref[-1].std()),
              "pos frac", float((outs_ref[-1] > 0).float().mean()))

    # ---- pruned variant (reference class with narrower stage convs) ---------
    H, W = 48, 72
    x, m = synth.make_frame(5, 0, H, W, noise=True)
    sd = synth.calibrate(synth.prune_state_dict(synth.make_state_dict(0, "random"), 0.5), O.vgg_forward, x)
    net = ns.OSVOS_VGG(pretrained=0)
    idxs = O.stage_conv_indices()
    for si in range(5):
        for mi in idxs[si]:
            w = sd[f"stages.{si}.{mi}.weight"]
            net.stages[si][mi] = torch.nn.Conv2d(w.shape[1], w.shape[0], 3, padding=1, bias=False)   # prune.py:493-500
        if si > 0:
            w = sd[f"side_prep.{si - 1}.weight"]
            net.side_prep[si - 1] = torch.nn.Conv2d(w.shape[1], 16, 3, padding=1)
    net.load_state_dict(sd)
    outs_ref = net.forward(x)
    outs_mine = O.vgg_forward(sd, x)
    for a, b in zip(outs_ref, outs_mine):
        assert torch.equal(a, b)
    torch.save(dict(H=H, W=W, seq=5, frame=0, keep=0.5, sd_checksum=checksum(sd),
                    outs=[o.detach().clone() for o in outs_ref],
                    n_params=sum(v.numel() for v in sd.values())),
               os.path.join(GOLD, "fwd_48x72_pruned50.pt"))
    print("pruned ok", sum(v.numel() for v in sd.values()))


if __name__ == "__main__":
    sys.exit(main())
