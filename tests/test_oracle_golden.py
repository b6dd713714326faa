"""CPU: the oracle restatement reproduces every fixture minted from the live
reference (oracle/make_golden.py), and the numpy second opinion agrees."""
import os

import numpy as np
import pytest
import torch

from fosvos_b200 import synth
from oracle import osvos_oracle as O, osvos_numpy as ON
from conftest import GOLDEN


def _load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def _case_inputs(fix):
    x, m = synth.make_frame(fix["seq"], fix["frame"], fix["H"], fix["W"], noise=fix.get("noise", False))
    assert abs(float(x.double().sum()) - fix["x_checksum"]) < 1e-6 * max(1.0, abs(fix["x_checksum"]))
    sd = synth.make_state_dict(0, fix["kind"])
    sd = synth.calibrate(sd, O.vgg_forward, x, mask=m if fix["kind"] == "structured" else None)
    for k, (s, a) in fix["sd_checksum"].items():
        assert abs(float(sd[k].double().sum()) - s) <= 1e-9 * max(1.0, a), k
    return x, m, sd


def test_kat_layers():
    kat = _load("kat.pt")
    for k in (4, 8, 16, 32, 3, 5):
        assert np.array_equal(kat[f"upsample_filt_{k}"].numpy(), O.upsample_filt(k))
    assert np.allclose(O.upsample_filt(4), np.outer([.25, .75, .75, .25], [.25, .75, .75, .25]))
    for key, (oy, ox) in kat["center_crop_origin"].items():
        src, dst = key.split("->")
        ih, iw = map(int, src.split("x"))
        h, w = map(int, dst.split("x"))
        l, r, t, b = O.center_crop_pads(ih, iw, h, w)
        assert (-t, -l) == (oy, ox)
        assert ih + t + b == h and iw + l + r == w
    assert kat["center_crop_origin"]["483x857->480x854"] == (1, 1)
    assert kat["state_dict_spec"] == O.state_dict_spec()
    for mode in ("online", "offline"):
        table = kat[f"optimizer_groups_{mode}"]
        mine = O.optimizer_groups([k for k, _ in O.state_dict_spec()], mode)
        assert [g["keys"] for g in mine] == [g["keys"] for g in table]
        assert np.allclose([g["lr"] for g in mine], [g["lr"] for g in table], rtol=1e-12, atol=0)
        assert [g["weight_decay"] for g in mine] == [g["weight_decay"] for g in table]
    online_keys = sum((g["keys"] for g in kat["optimizer_groups_online"]), [])
    assert not any(k.startswith("score_dsn") for k in online_keys)


@pytest.mark.parametrize("name", ["loss_4x4_sa1", "loss_4x4_sa0", "loss_2x1x37x53_sa1", "loss_2x1x37x53_sa0", "loss_teacher_logits"])
def test_kat_loss(name):
    c = _load("kat.pt")[name]
    sa = not name.endswith("sa0")
    out = c["output"].clone().requires_grad_(True)
    loss = O.class_balanced_cross_entropy_loss(out, c["label"], size_average=sa)
    assert torch.equal(loss.detach(), c["loss"])
    (g,) = torch.autograd.grad(loss, out)
    assert torch.equal(g, c["grad"])
    g2 = O.class_balanced_cross_entropy_grad(c["output"], c["label"], sa)
    assert torch.allclose(g2, c["grad"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["fwd_48x72_random", "fwd_45x70_random", "fwd_64x96_structured"])
def test_forward_and_backward_fixture(name):
    fix = _load(name + ".pt")
    x, m, sd = _case_inputs(fix)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    outs = O.vgg_forward(params, x)
    for a, b in zip(outs, fix["outs"]):
        assert torch.equal(a.detach(), b)
    loss = O.class_balanced_cross_entropy_loss(outs[-1], m, size_average=False)
    assert torch.equal(loss.detach(), fix["loss"])
    loss.backward()
    for k, g in fix["grads"].items():
        assert torch.allclose(params[k].grad, g, rtol=1e-5, atol=1e-6 * float(g.abs().max())), k
    # independent numpy/fp64 restatement of the third-party primitives
    outs_np = ON.vgg_forward({k: v.numpy() for k, v in sd.items()}, x.numpy())
    for a, b in zip(fix["outs"], outs_np):
        assert float(np.abs(a.numpy() - b).max()) < 5e-4


def test_finetune_fixture():
    fix = _load("fwd_48x72_random.pt")
    x, m, sd = _case_inputs(fix)
    ft = fix["finetune"]
    new_sd, losses = O.finetune(sd, x, m, ft["n_iters"], ft["avg_grad_every_n"], learning_rate=ft["learning_rate"])
    assert np.allclose(losses, ft["losses"], rtol=1e-6)
    for k, d in ft["deltas"].items():
        mine = new_sd[k] - sd[k]
        assert torch.allclose(mine, d, rtol=1e-4, atol=1e-7 * max(1.0, float(sd[k].abs().max()))), k
    assert not any((new_sd[k] != sd[k]).any() for k in sd if k.startswith("score_dsn") or k.startswith("upscale"))
    fused = O.vgg_forward(new_sd, x)[-1]
    assert torch.allclose(fused, ft["fused_after"], rtol=0, atol=2e-5)


def test_pruned_fixture():
    fix = _load("fwd_48x72_pruned50.pt")
    x, _ = synth.make_frame(fix["seq"], fix["frame"], fix["H"], fix["W"], noise=True)
    sd = synth.calibrate(synth.prune_state_dict(synth.make_state_dict(0, "random"), fix["keep"]), O.vgg_forward, x)
    assert sum(v.numel() for v in sd.values()) == fix["n_params"] == 4129141
    assert [tuple(sd[k].shape) for k, _ in O.state_dict_spec(synth.pruned_channels(0.5), stage_bias=False)] == \
           [s for _, s in O.state_dict_spec(synth.pruned_channels(0.5), stage_bias=False)]
    for a, b in zip(O.vgg_forward(sd, x), fix["outs"]):
        assert torch.equal(a, b)


def test_numpy_primitives_small():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 7, 9, generator=g)
    w = torch.randn(5, 3, 3, 3, generator=g)
    b = torch.randn(5, generator=g)
    import torch.nn.functional as F
    assert np.allclose(ON.conv2d(x.numpy(), w.numpy(), b.numpy(), 1), F.conv2d(x, w, b, padding=1).numpy(), atol=1e-5)
    assert np.array_equal(ON.max_pool2x2_ceil(x.double().numpy()), F.max_pool2d(x.double(), 2, 2, ceil_mode=True).numpy())
    wt = torch.randn(3, 3, 8, 8, generator=g)
    assert np.allclose(ON.conv_transpose2d(x.numpy(), wt.numpy(), 4), F.conv_transpose2d(x, wt, stride=4).numpy(), atol=1e-5)


def test_mask_iou_counts():
    a = torch.tensor([[1, 1, 0, 0]], dtype=torch.uint8)
    b = torch.tensor([[1, 0, 1, 0]], dtype=torch.uint8)
    assert O.mask_iou_counts(a, b) == (1, 3)
    assert torch.equal(O.binarise(O.probabilities(torch.tensor([-1.0, 0.0, 2.0]))), torch.tensor([0, 1, 1], dtype=torch.uint8))
