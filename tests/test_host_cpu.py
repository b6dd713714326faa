"""CPU: host logic, the C-ABI surface, loud failure without a GPU, sharding (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

import fosvos_b200 as FB
from fosvos_b200 import _lib as L, synth
from oracle import osvos_oracle as O
from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "fosvos_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fosvos_[a-z0-9_]+)\s*\(", src)) - {"fosvos_sgd_entry"})


def test_library_exports_every_declared_symbol():
    from fosvos_b200.build import build
    build()
    lib = ctypes.CDLL(L.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fosvos_b200.h but not exported"
    assert sorted(L.SIGNATURES) == names, "ctypes binding and header disagree"
    assert L.lib().fosvos_abi_version() >= 1
    # scalars + k x k tables (g1, gs, G[16]) per stage, then the two 340-phase float4 tap tables of the fast path, the 2 x 30 float4 separable tables and the flag
    assert L.lib().fosvos_side_params_bytes() == 4 * (4 + 4 * 36 + 18 * (16 + 64 + 256 + 1024)) + 2 * 16 * (4 + 16 + 64 + 256) + 2 * 16 * 30 + 16


@pytest.mark.parametrize("HW", [(480, 854), (240, 426), (384, 682), (48, 72), (45, 70), (70, 1100), (33, 18), (6, 2), (17, 2050)])
def test_side_upsample_plan_keeps_every_tap_inside_its_staging_window(HW):
    """The two-pixel separable up-sampling kernel (side.cu: side_upsample_sep2_kernel) reads the low-res taps of a
    work item from a shared-memory window staged per stage.  The launch plan is host arithmetic exported through the
    C ABI; the kernel's index formulas are restated here and checked exhaustively: every tap a pixel pair touches lies
    inside the staged window, the window fits the plan's shared memory, the column blocks cover the frame."""
    H, W = HW
    hs, ws, h, w = [], [], H, W
    for _ in range(4):
        h, w = (h + 1) // 2, (w + 1) // 2            # ceil-mode 2x2 pools (osvos_vgg.py:90)
        hs.append(h)
        ws.append(w)
    top = [((2 << i) * hs[i] + (2 << i) - H) // 2 for i in range(4)]      # center_crop offsets (osvos_layers.py:47-54)
    left = [((2 << i) * ws[i] + (2 << i) - W) // 2 for i in range(4)]
    for N, sms in ((1, 148), (5, 148), (16, 148), (3, 4)):
        rows, ppi, smem = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        assert L.lib().fosvos_side_upsample_plan(N, H, W, sms, ctypes.byref(rows), ctypes.byref(ppi), ctypes.byref(smem)) == 0
        R, P = rows.value, ppi.value
        pairs = W // 2
        assert 1 <= R <= H and 1 <= P <= 256
        xblocks = -(-pairs // P)
        assert xblocks * P >= pairs
        cw = [((2 * P - 1) >> (i + 1)) + 4 for i in range(4)]
        rh = [((R - 1) >> (i + 1)) + 3 for i in range(4)]
        assert smem.value == 2 * 8 * sum(c * r for c, r in zip(cw, rh)) <= 100 * 1024
        for strip in range(-(-H // R)):
            y0, y1 = strip * R, min(H, strip * R + R)
            for i in range(4):
                ro = ((y0 + top[i]) >> (i + 1)) - 1
                lrs = [((y + top[i]) >> (i + 1)) - ro for y in range(y0, y1)]
                assert min(lrs) - 1 >= 0 and max(lrs) <= rh[i] - 1, (N, strip, i)
                assert ro + max(lrs) <= hs[i]            # row h is the zero row below the map, never further
        for xb in range(xblocks):
            x0 = 2 * xb * P
            for i in range(4):
                co = ((x0 + left[i]) >> (i + 1)) - 1
                for t in range(P):
                    if xb * P + t >= pairs:
                        break
                    c = ((x0 + 2 * t + left[i]) >> (i + 1)) - co
                    assert c - 1 >= 0 and c + 1 <= cw[i] - 1, (N, xb, i, t)
    bad = ctypes.c_int()
    assert L.lib().fosvos_side_upsample_plan(1, 33, 17, 148, ctypes.byref(bad), ctypes.byref(bad), ctypes.byref(bad)) != 0     # odd width


def test_no_cpu_fallback():
    net = FB.OSVOS_VGG(pretrained=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 3, 16, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FB.class_balanced_cross_entropy_loss(torch.zeros(1, 1, 4, 4), torch.ones(1, 1, 4, 4))
    if not torch.cuda.is_available():
        assert L.lib().fosvos_device_check(0) == -3
        assert "no CPU fallback" in L.last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fosvos_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f


def test_module_surface_and_state_dict():
    net = FB.OSVOS_VGG(pretrained=0)
    assert [(k, tuple(v.shape)) for k, v in net.state_dict().items()] == O.state_dict_spec()
    sd = synth.make_state_dict(0)
    net.load_state_dict(sd)                              # strict
    assert all(torch.equal(net.state_dict()[k], sd[k]) for k in sd)
    # default init statistics (osvos_vgg.py:97-111)
    fresh = FB.OSVOS_VGG(pretrained=0)
    w = fresh.stages[2][3].weight
    assert abs(float(w.std()) - 1e-3) < 1e-4 and float(fresh.stages[2][3].bias.abs().max()) == 0.0
    assert torch.equal(fresh.upscale_[1].weight.data[0, 0], torch.from_numpy(FB.upsample_filt(8)).float())
    with pytest.raises(Exception, match="channels need to be the same"):
        FB.interp_surgery(torch.nn.ConvTranspose2d(2, 3, 4))
    # whole-module pickling as NetworkProvider.save_model does (network_provider.py:60-63)
    import io
    buf = io.BytesIO()
    torch.save(net, buf)
    buf.seek(0)
    net2 = torch.load(buf, weights_only=False)
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))


def test_layer_helpers_match_oracle():
    import numpy as np
    for k in (3, 4, 5, 8, 16, 32):
        assert np.array_equal(FB.upsample_filt(k), O.upsample_filt(k))
    x = torch.arange(13 * 9, dtype=torch.float32).view(1, 1, 13, 9)
    assert torch.equal(FB.center_crop(x, 8, 4), O.center_crop(x, 8, 4))
    x = torch.arange(483 * 857, dtype=torch.float32).view(1, 1, 483, 857)
    assert FB.center_crop(x, 480, 854)[0, 0, 0, 0] == 857 + 1


def test_optimizer_groups_and_fused_sgd_config():
    net = FB.OSVOS_VGG(pretrained=0)
    id2k = {id(p): k for k, p in net.named_parameters()}
    for mode, fn in (("online", FB.get_optimizer_online), ("offline", FB.get_optimizer_offline)):
        for fused in (True, False):
            opt = fn(net, fused=fused)
            mine = [dict(keys=[id2k[id(p)] for p in g["params"]], lr=g["lr"], weight_decay=g["weight_decay"]) for g in opt.param_groups]
            ref = O.optimizer_groups(list(id2k.values()), mode)
            assert [m["keys"] for m in mine] == [r["keys"] for r in ref]
            assert all(abs(m["lr"] - r["lr"]) < 1e-22 and m["weight_decay"] == r["weight_decay"] for m, r in zip(mine, ref))
    with pytest.raises(ValueError):
        FB.FusedSGD(net.parameters(), lr=0.1, nesterov=True, momentum=0.9)


def test_sequence_sharding_rule():
    seqs = list(range(20))
    parts = [FB.sequences_for_rank(seqs, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == seqs
    assert [len(p) for p in parts] == [3, 3, 3, 3, 2, 2, 2, 2]
    assert parts[1] == [1, 9, 17]                        # i % group_size == group (train_online.py:184-186)


def test_synth_is_deterministic():
    a, ma = synth.make_frame(2, 5, 48, 72)
    b, mb = synth.make_frame(2, 5, 48, 72)
    assert torch.equal(a, b) and torch.equal(ma, mb)
    assert 0.05 < float(ma.mean()) < 0.35
    c, _ = synth.make_frame(2, 6, 48, 72)
    assert not torch.equal(a, c)
    sd1, sd2 = synth.make_state_dict(0), synth.make_state_dict(0)
    assert all(torch.equal(sd1[k], sd2[k]) for k in sd1)
    p = synth.prune_state_dict(sd1, 0.5)
    assert p["stages.4.5.weight"].shape == (256, 256, 3, 3) and "stages.0.0.bias" not in p


def test_bench_sharding_two_ranks_gloo(tmp_path):
    """world_size-2 run of the multi-GPU plumbing (sequence sharding + max-over-ranks reduce) on CPU/gloo."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from fosvos_b200.sharding import init_distributed, max_over_ranks, sum_over_ranks, allreduce_flat\n"
        "from fosvos_b200 import sequences_for_rank\n"
        "rank, world = init_distributed(backend='gloo')\n"
        "mine = sequences_for_rank(list(range(5)), rank, world)\n"
        "t = max_over_ranks(float(10 + rank), device='cpu')\n"
        "n = sum_over_ranks(float(len(mine)), device='cpu')\n"
        "assert t == 11.0 and n == 5.0, (t, n)\n"
        "flat = torch.arange(6, dtype=torch.float32) * (rank + 1)\n"
        "allreduce_flat(flat)          # the data-parallel gradient exchange of offline training / distillation\n"
        "assert torch.equal(flat, torch.arange(6, dtype=torch.float32) * 3), flat\n"
        "from fosvos_b200.sharding import allreduce_gradients\n"
        "ps = [torch.nn.Parameter(torch.zeros(n)) for n in (5, 300, 7, 2)]\n"
        "for i, p in enumerate(ps[:3]): p.grad = torch.full_like(p, float((i + 1) * (rank + 1)))\n"
        "allreduce_gradients(ps, bucket_bytes=1024)       # three gradients in two buckets, one parameter without a gradient\n"
        "assert all(torch.equal(p.grad, torch.full_like(p, 3.0 * (i + 1))) for i, p in enumerate(ps[:3])) and ps[3].grad is None\n"
        "# bucketed exchange of the data-parallel trainer: per-stage buckets, p.grad are views, async all-reduce per bucket\n"
        "import fosvos_b200 as FB\n"
        "from fosvos_b200.sharding import GradBuckets\n"
        "net = FB.OSVOS_VGG(pretrained=0)\n"
        "params = dict(net.named_parameters())\n"
        "bk = GradBuckets(net, params, 'cpu')\n"
        "for k, name in enumerate(net._grad_names()): bk.view(name).fill_(float((k % 7 + 1) * (rank + 1)))\n"
        "works = [bk.allreduce(b) for b in range(bk.n_buckets)]\n"
        "for w in works: w.wait()\n"
        "for k, name in enumerate(net._grad_names()): assert torch.equal(bk.view(name), torch.full_like(params[name], 3.0 * (k % 7 + 1))), name\n"
        "open(os.path.join(os.path.dirname(__file__), f'rank{rank}.txt'), 'w').write(repr(mine))\n"
        "dist.destroy_process_group()\n")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    # each rank writes its own file: the two stdout streams interleave under torchrun
    assert (tmp_path / "rank0.txt").read_text() == "[0, 2, 4]" and (tmp_path / "rank1.txt").read_text() == "[1, 3]"


def test_grad_buckets_layout():
    """The data-parallel gradient buffer: bucket k = stage 4 - k (+ its side_prep conv), heads in the last bucket; every
    gradient-carrying parameter has exactly one view, buckets are contiguous and cover the buffer (59.7 MB for the full VGG:
    the lr = 0 up-sampling weights are left out)."""
    from fosvos_b200.sharding import GradBuckets
    net = FB.OSVOS_VGG(pretrained=0)
    params = dict(net.named_parameters())
    bk = GradBuckets(net, params, "cpu")
    names = net._grad_names()
    assert sorted(sum(bk.bucket_names, [])) == sorted(names) and not any(n.startswith("upscale") for n in names)
    assert bk.flat.numel() == sum(params[n].numel() for n in names) == 15267157 - 2 * (16 * 16 + 1) * (16 + 64 + 256 + 1024) // 2
    assert bk.ranges[0][0] == 0 and bk.ranges[-1][1] == bk.flat.numel()
    assert all(bk.ranges[i][1] == bk.ranges[i + 1][0] for i in range(4))
    assert all(n.startswith(("stages.4.", "side_prep.3.")) for n in bk.bucket_names[0])
    assert [round((hi - lo) * 4 / 1e6, 1) for lo, hi in bk.ranges] == [28.6, 23.9, 6.0, 1.0, 0.2]
    for n in names:
        v = bk.view(n)
        assert v.shape == params[n].shape and v.untyped_storage().data_ptr() == bk.flat.untyped_storage().data_ptr()


def test_leaf_modules_are_the_reference_types_and_never_fall_back():
    """Leaves stay nn.Conv2d / nn.ReLU / nn.MaxPool2d / nn.ConvTranspose2d instances (prune.py:49-50 isinstance checks), plain
    torch leaves assigned by surgery are adopted, and a leaf called on a CPU tensor raises instead of running a library kernel."""
    from fosvos_b200 import leaf
    net = FB.OSVOS_VGG(pretrained=0)
    assert all(isinstance(m, leaf.Conv2d) for m in net.modules() if isinstance(m, torch.nn.Conv2d))
    assert isinstance(net.stages[1][0], torch.nn.MaxPool2d) and isinstance(net.stages[0][1], torch.nn.ReLU)
    net.stages[0][0] = torch.nn.Conv2d(3, 48, 3, padding=1, bias=False)           # prune-style surgery (prune.py:490-514)
    net.precision = "fp32"
    assert type(net.stages[0][0]) is leaf.Conv2d and net.stages[0][0]._fosvos_precision == "fp32"
    for m, x in ((net.stages[0][0], torch.zeros(1, 3, 8, 8)), (net.stages[0][1], torch.zeros(1, 4, 8, 8)), (net.stages[1][0], torch.zeros(1, 4, 8, 8))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m(x)
    with pytest.raises(RuntimeError, match="only fused"):
        net.upscale[0](torch.zeros(1, 16, 4, 4))
    h = net.stages[2][3].register_forward_hook(lambda *a: None)
    assert net._leaf_hooks_present()
    h.remove()
    assert not net._leaf_hooks_present()
