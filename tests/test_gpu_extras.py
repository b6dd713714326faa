"""GPU: the callers either side of the hot path -- Adam, distillation criteria, the Taylor
pruning criterion and VGG channel surgery, raw uint8 frame ingest, the offline deep-supervision trainer --
against torch / the oracle on seeded inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import fosvos_b200 as FB  # noqa: E402
from fosvos_b200 import ops, synth  # noqa: E402
from fosvos_b200 import prune as P  # noqa: E402
from fosvos_b200.distill import MimicTrainer  # noqa: E402
from fosvos_b200.online import OnlineTrainer  # noqa: E402
from oracle import osvos_oracle as O  # noqa: E402

DEV = "cuda"


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _net(sd, precision="fp32"):
    net = FB.OSVOS_VGG(pretrained=0)
    net.load_state_dict(sd)
    net = net.to(DEV)
    net.precision = precision
    return net


def _small_case(H=45, W=70, kind="random"):
    x, m = synth.make_frame(3, 0, H, W)
    sd = synth.calibrate(synth.make_state_dict(0, kind), O.vgg_forward, x, mask=m if kind == "structured" else None)
    return x, m, sd


def test_fused_adam_matches_torch_adam():
    g = _gen(41)
    shapes = [(64, 3, 3, 3), (64,), (5000,), (16, 128, 3, 3), (1,)]
    ref = [torch.randn(s, generator=g).requires_grad_(True) for s in shapes]
    mine = [r.detach().clone().to(DEV).requires_grad_(True) for r in ref]
    o_ref = torch.optim.Adam(ref, lr=1e-3, weight_decay=2e-4)
    o_mine = FB.FusedAdam(mine, lr=1e-3, weight_decay=2e-4)
    for step in range(4):
        for r, m in zip(ref, mine):
            gr = torch.randn(r.shape, generator=g)
            r.grad = gr.clone()
            if m.grad is None:
                m.grad = gr.to(DEV)
            else:
                m.grad.copy_(gr)                      # same buffer: the device table stays valid
        o_ref.step()
        o_mine.step_and_zero()
        for r, m in zip(ref, mine):
            assert torch.allclose(m.detach().cpu(), r.detach(), rtol=2e-6, atol=2e-7), (step, float((m.detach().cpu() - r.detach()).abs().max()))
            assert float(m.grad.abs().max()) == 0.0
    st = o_mine.state[mine[0]]
    assert torch.allclose(st["exp_avg"].cpu(), o_ref.state[ref[0]]["exp_avg"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("kind", ["mse", "l1"])
@pytest.mark.parametrize("size_average", [False, True])
def test_pixel_losses(kind, size_average):
    g = _gen(43)
    x = torch.randn(2, 1, 37, 53, generator=g).requires_grad_(True)
    t = torch.randn(2, 1, 37, 53, generator=g)
    red = "mean" if size_average else "sum"
    ref = F.mse_loss(x, t, reduction=red) if kind == "mse" else F.l1_loss(x, t, reduction=red)
    ref.backward()
    xd = x.detach().to(DEV).requires_grad_(True)
    fn = FB.mse_loss if kind == "mse" else FB.l1_loss
    loss = fn(xd, t.to(DEV), size_average=size_average)
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    (3.0 * loss).backward()
    assert torch.allclose(xd.grad.cpu(), 3.0 * x.grad, rtol=1e-5, atol=1e-8)


def test_ingest_u8_matches_loader_arithmetic():
    """uint8 BGR frame -> float32 - meanval -> CHW (davis_2016.py:127-128 + ToTensor) == one ingest kernel."""
    g = _gen(45)
    img = torch.randint(0, 256, (2, 19, 23, 3), generator=g, dtype=torch.uint8)
    mean = torch.tensor(FB.OSVOS_VGG.MEANVAL, dtype=torch.float32)
    ref = (img.float() - mean).permute(0, 3, 1, 2).contiguous()
    y = ops.ingest_u8(img.to(DEV), FB.OSVOS_VGG.MEANVAL, torch.float32)
    assert torch.equal(ops.nhwc_to_nchw(y, 3).cpu(), ref)
    assert float(y[..., 3:].abs().max()) == 0.0
    # and through the network: same logits as the fp32 NCHW contract
    x, m, sd = _small_case()
    net = _net(sd, "fp32")
    img = torch.randint(0, 256, (1, 45, 70, 3), generator=g, dtype=torch.uint8)
    a = net.predict(img.to(DEV))[0][4]
    b = net.predict((img.float() - mean).permute(0, 3, 1, 2).contiguous().to(DEV))[0][4]
    assert torch.equal(a, b)


@pytest.mark.parametrize("is_offline", [False, True])
def test_taylor_ranks_match_hooked_reference(is_offline):
    """FilterPruner.compute_rank (prune.py:163-178): sum(activation * grad) / (N H W) per output filter."""
    x, m, sd = _small_case()
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    outs, inter = O.vgg_forward(params, x, return_intermediates=True)
    acts = {k: v for k, v in inter.items() if k.startswith("stages.")}
    for v in acts.values():
        v.retain_grad()
    ls = [O.class_balanced_cross_entropy_loss(o, m, size_average=False) for o in outs]
    (sum(ls[:-1]) + ls[-1] if is_offline else ls[-1]).backward()
    net = _net(sd, "fp32")
    pr = P.FilterPruner(net)
    pr.forward_backward(x.to(DEV), m.to(DEV), is_offline)
    names = [f"stages.{si}.{mi}" for si, mi in P.stage_conv_index(net)]
    for k, name in enumerate(names):
        a = acts[name]
        ref = (a * a.grad).sum(dim=(0, 2, 3)) / (a.shape[0] * a.shape[2] * a.shape[3])
        got = pr.filter_ranks[k].cpu()
        assert torch.allclose(got, ref.detach(), rtol=2e-3, atol=1e-6 + 2e-4 * float(ref.abs().max())), (name, float((got - ref).abs().max()))


def test_prune_plan_and_surgery_keep_the_function_of_survivors():
    """Plan bookkeeping (prune.py:203-223) and surgery (prune.py:490-514): after removing filters, the pruned
    network equals the oracle evaluated on the pruned state_dict, and shapes follow the topology rules."""
    x, m, sd = _small_case(48, 72)
    net = _net(sd, "fp32")
    plan = P.prune_step(net, [x.to(DEV)], [m.to(DEV)], n_filters=40, is_offline=False)
    assert len(plan) == 40
    convs = [net.stages[si][mi] for si, mi in P.stage_conv_index(net)]
    removed = {}
    for l, _ in plan:
        removed[l] = removed.get(l, 0) + 1
    full = [64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512]
    for k, c in enumerate(convs):
        assert c.out_channels == full[k] - removed.get(k, 0)
        if k + 1 < len(convs):
            assert convs[k + 1].in_channels == c.out_channels
    last = {1: 3, 2: 6, 3: 9, 4: 12}
    for si, k in last.items():
        assert net.side_prep[si - 1].in_channels == convs[k].out_channels
    new_sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    with torch.no_grad():
        ref = O.vgg_forward(new_sd, x)
    outs = net.forward(x.to(DEV))
    for o, r in zip(outs, ref):
        assert float((torch.sigmoid(o.cpu()) - torch.sigmoid(r)).abs().max()) <= 1e-4
    # bf16 tensor-core path on the ragged channel counts
    net.precision = "bf16"
    outs_b = net.forward(x.to(DEV))
    from conftest import bf16_budget
    assert float((torch.sigmoid(outs_b[4].cpu()) - torch.sigmoid(ref[4])).abs().max()) <= bf16_budget(new_sd, x, ref)
    # the VGG surgery keeps the trained biases (sliced on out-channel pruning); bias-free convs are the explicit opt-in
    assert all(c.bias is not None and c.bias.shape[0] == c.out_channels for c in convs) and net.side_prep[0].bias is not None


def test_l2_prune_half_builds_the_config3_network():
    x0, m0, sd = _small_case(48, 72, "structured")
    net = _net(sd, "bf16")
    P.l2_prune_half(net, 0.5)
    convs = [net.stages[si][mi] for si, mi in P.stage_conv_index(net)]
    assert [c.out_channels for c in convs] == [32, 32, 64, 64, 128, 128, 128, 256, 256, 256, 256, 256, 256]
    x = torch.randn(2, 3, 40, 56, generator=_gen(5)) * 50
    outs = net.forward(x.to(DEV))
    new_sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    with torch.no_grad():
        ref = O.vgg_forward(new_sd, x)
    from conftest import bf16_budget
    # noise frames through a pruned structured net: held to the reference's own bf16 loss on the same weights and input
    assert float((torch.sigmoid(outs[4].detach().cpu()) - torch.sigmoid(ref[4])).abs().max()) <= bf16_budget(new_sd, x, ref)


@pytest.mark.parametrize("criterion", ["MSE", "L1", "CBCEL"])
def test_mimic_step_matches_autograd_reference(criterion):
    """mimic.py:144-218 with learn_from='teacher': 2 minibatches, avg_grad_every_n=2, one Adam step."""
    x, m, sd_t = _small_case(45, 70, "structured")
    sd_s = synth.make_state_dict(7, "random")
    x2 = torch.roll(x, 5, 3)
    epoch, n_epochs, n = 3, 10, 2
    # reference: torch autograd through the oracle forward, torch.optim.Adam
    params = {k: v.clone().requires_grad_(not k.startswith("upscale")) for k, v in sd_s.items()}
    opt = torch.optim.Adam([p for p in params.values() if p.requires_grad], lr=1e-3, weight_decay=0.0002)
    crit = {"MSE": lambda a, b: F.mse_loss(a, b, reduction="sum"), "L1": lambda a, b: F.l1_loss(a, b, reduction="sum"),
            "CBCEL": lambda a, b: O.class_balanced_cross_entropy_loss(a, b)}[criterion]
    ref_losses = []
    for xb in (x, x2):
        with torch.no_grad():
            t_out = O.vgg_forward(sd_t, xb)
        s_out = O.vgg_forward(params, xb)
        ls = [crit(a, b) for a, b in zip(s_out, t_out)]
        loss = (1 - epoch / n_epochs) * sum(ls[:-1]) + ls[-1]
        ref_losses.append(float(loss))
        (loss / n).backward()
    opt.step()
    student, teacher = _net(sd_s, "fp32"), _net(sd_t, "fp32")
    tr = MimicTrainer(student, teacher, criterion, learning_rate=1e-3, weight_decay=0.0002, avg_grad_every_n=n)
    losses = [float(tr.step(xb.to(DEV), None, epoch, n_epochs)) for xb in (x, x2)]
    assert np.allclose(losses, ref_losses, rtol=1e-4)
    mine = dict(student.named_parameters())
    for k in ["stages.0.0.weight", "stages.2.3.bias", "stages.4.5.weight", "side_prep.1.weight", "score_dsn.2.weight", "fuse.weight", "fuse.bias"]:
        d = (mine[k].detach().cpu() - sd_s[k])
        dr = (params[k].detach() - sd_s[k])
        # Adam's first step moves every weight by ~lr * g / (|g| + eps): compare the updates where the gradient is
        # above the fp32 noise of the two summation orders (elsewhere its SIGN, hence the whole step, is noise)
        gr = params[k].grad
        sig = gr.abs() > 1e-3 * float(gr.abs().max())
        assert int(sig.sum()) > 0
        assert float((d - dr)[sig].abs().max()) <= 0.05 * 1e-3 + 1e-7, (k, float((d - dr)[sig].abs().max()))
    assert float(tr.flat_grad.abs().max()) == 0.0


def test_offline_trainer_deep_supervision_weight_on_device():
    """OnlineTrainer(deep_supervision=w) == oracle offline loop (train_offline.py:84-88,102-110); the weight is a
    device scalar, so the captured graphs follow set_deep_supervision() without re-capture."""
    x, m, sd = _small_case()
    net = _net(sd, "fp32")
    opt = FB.get_optimizer_offline(net)
    tr = OnlineTrainer(net, 45, 70, avg_grad_every_n=2, optimizer=opt, use_graph=True, deep_supervision=0.75)
    tr.set_frame(x.to(DEV), m.to(DEV))
    losses = []
    tr.run(2, losses)
    ref_sd, ref_losses = O.finetune(sd, x, m, 2, 2, mode="offline", epoch_frac=0.25)
    assert np.allclose(losses, ref_losses, rtol=2e-4)
    tr.set_deep_supervision(0.5)
    losses2 = []
    tr.run(2, losses2)
    ref_sd2, ref_losses2 = O.finetune(ref_sd, x, m, 2, 2, mode="offline", epoch_frac=0.5)
    # (momentum restarts in the oracle call: compare the losses of the first micro-iteration only, then the weights loosely)
    assert abs(losses2[0] - ref_losses2[0]) <= 2e-4 * abs(ref_losses2[0])
