"""CPU: host-side logic of the drop-in package — module construction / state_dict layout, error behaviour,
flat parameter groups, tile heuristics, and the data-parallel gradient bucketing over gloo (world size 2)."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

from oracle import nets as onets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_module_layout_is_the_reference_checkpoint_layout():
    from gan_inpainting_b200.lib.models import networks

    g = networks.get_network("generator", "unet")
    ref = onets.init_unet(0)
    assert list(g.state_dict().keys()) == list(ref.keys())
    assert all(g.state_dict()[k].shape == ref[k].shape for k in ref)
    g.load_state_dict(ref, strict=True)
    assert [n for n, _ in g.named_parameters() if "bias" in n] != []  # the `"bias" in n` filter of wgan_l1.py:103-108 is meaningful
    d = networks.PatchGANDiscriminator(sigmoid=False, imagedim=256)
    refd = onets.init_patchgan(0, res=256)
    assert list(d.state_dict().keys()) == list(refd.keys())
    assert d.state_dict()["model.13.weight"].shape == (1, 169)  # (256/16 - 3)^2, SURVEY.md D4
    assert networks.PatchGANDiscriminator().state_dict()["model.13.weight"].shape == (1, 25)
    # dropout placement: get_network passes the truthy string 'False'  -> levels 5 and 6
    levels = [i + 1 for i, p in enumerate(g._levels()) if p["dropout"]]
    assert levels == [5, 6]
    g2 = networks.UnetGenerator(1, 1, 7, ngf=64, norm_layer=networks.get_norm_layer("batch"), use_dropout=False)
    assert not any(p["dropout"] for p in g2._levels())


@pytest.mark.skipif(not os.path.exists("/root/reference/lib/models/networks.py"), reason="reference tree not mounted (GPU box)")
def test_checkpoint_interchange_with_the_reference_modules(tmp_path):
    """an ``epochN_G.pt`` written by either side (``torch.save(net_G.state_dict())``, wgan_l1.py:269-271)
    loads strictly into the other side's module (eval.py:77-79), for the generator and the discriminator."""
    from gan_inpainting_b200.lib.models import networks

    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference")
    try:
        for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
            del sys.modules[k]
        from lib.models import networks as ref_networks
    finally:
        sys.path.pop(0)
        for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
            del sys.modules[k]
    torch.manual_seed(3)
    ours_g, ref_g = networks.get_network("generator", "unet"), ref_networks.get_network("generator", "unet")
    ours_d, ref_d = networks.get_network("discriminator", "patchgan"), ref_networks.get_network("discriminator", "patchgan")
    for ours, ref, name in ((ours_g, ref_g, "G"), (ours_d, ref_d, "D")):
        path = tmp_path / f"epoch1_{name}.pt"
        torch.save(ours.state_dict(), path)  # ours -> reference
        ref.load_state_dict(torch.load(path), strict=True)
        for (k, a), (k2, b) in zip(ours.state_dict().items(), ref.state_dict().items()):
            assert k == k2 and torch.equal(a, b), k
        with torch.no_grad():
            for p_ in ref.parameters():
                p_.mul_(0.5).add_(0.01)
        torch.save(ref.state_dict(), path)  # reference -> ours
        ours.load_state_dict(torch.load(path), strict=True)
        assert all(torch.equal(a, b) for a, b in zip(ours.state_dict().values(), ref.state_dict().values()))
        assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]


def test_default_initialisation_is_pytorch_default():
    from gan_inpainting_b200.lib.models import networks

    torch.manual_seed(0)
    g = networks.get_network("generator", "unet")
    w = g.state_dict()["model.model.1.model.1.weight"]  # Conv2d(64,128): U(-1/sqrt(64*16), +)
    assert float(w.abs().max()) <= 1 / 32 + 1e-6 and float(w.abs().max()) > 0.9 / 32
    assert float(g.state_dict()["model.model.1.model.2.weight"].min()) == 1.0  # BN gamma = 1


def test_no_cpu_fallback_and_error_conventions():
    from gan_inpainting_b200._lib import B200Error
    from gan_inpainting_b200.lib.models import loss, networks

    g = networks.get_network("generator", "unet")
    with pytest.raises(B200Error, match="CUDA"):
        g(torch.zeros(1, 1, 128, 128))
    with pytest.raises(B200Error, match="CUDA"):
        networks.PatchGANDiscriminator()(torch.zeros(1, 1, 128, 128))
    with pytest.raises(B200Error):
        loss.RMSELoss()(torch.zeros(4, 4), torch.zeros(4, 4))
    with pytest.raises(Exception, match="Invalid generator network name"):
        networks.get_network("generator", "resnet")
    with pytest.raises(NotImplementedError):
        networks.get_norm_layer("layer")
    with pytest.raises(TypeError):
        loss.LocalLoss(loss.RMSELoss)  # unconstructible in the reference too (loss.py:28)
    assert isinstance(loss.LocalLoss(torch.nn.L1Loss), torch.nn.Module)


def test_set_requires_grad_and_average_meter():
    from gan_inpainting_b200.lib.models import networks, util

    d = networks.PatchGANDiscriminator()
    util.set_requires_grad([d], False)
    assert util.count_parameters(d) == 0
    util.set_requires_grad([d], True)
    assert util.count_parameters(d) == 2_763_546
    m = util.AverageMeter()
    m.update(2.0, 3)
    m.update(4.0, 1)
    assert m.avg == 2.5 and m.val == 4.0


def test_flat_groups():
    from gan_inpainting_b200 import optim as bo

    net = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.BatchNorm1d(3), torch.nn.Linear(3, 1))
    before = [p.detach().clone() for p in net.parameters()]
    fg = bo.flatten_parameters(net)
    assert fg.numel % 4 == 0 and all(o % 4 == 0 for o in fg.offsets)
    for p, b, o in zip(net.parameters(), before, fg.offsets):
        assert torch.equal(p.detach(), b)
        assert p.data_ptr() == fg.params.data_ptr() + 4 * o and p.grad.data_ptr() == fg.grads.data_ptr() + 4 * o
    assert bo.flatten_parameters(net) is fg
    net(torch.randn(4, 5)).sum().backward()  # autograd accumulates into the flat views in place
    assert float(fg.grads.abs().sum()) > 0
    groups, loose = bo._plan(list(net.parameters()))
    assert groups == [fg] and loose == []
    groups, loose = bo._plan(list(net.parameters())[:2])
    assert groups == [] and len(loose) == 2
    names, spans = fg.logging_offsets()
    assert all("bias" not in n for n in names) and len(names) == len(spans) == 3


def test_tile_heuristics():
    from gan_inpainting_b200 import engine

    assert engine._m_tiles(64, 64, 64) == 64 * 64 * 64 // 128
    assert engine._m_tiles(64, 1, 1) == 1 and engine._m_tiles(129, 1, 1) == 2
    assert engine.pick_bn(2048, 128) == 128          # plenty of pixel tiles: widest tile
    assert engine.pick_bn(8, 512) == 32               # d6 at batch 64: narrow tiles to fill 148 SMs
    assert engine.pick_bn(2, 512, phases=4) == 32


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _bucket_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from gan_inpainting_b200 import optim as bo
    from gan_inpainting_b200.trainer import GradBucketer

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(*[torch.nn.Linear(64, 64) for _ in range(6)])
    fg = bo.flatten_parameters(net)
    b = GradBucketer(fg, bucket_mb=64 * 64 * 4 * 2.5 / (1 << 20))  # ~2.5 layers per bucket
    # backward order == descending addresses; every rank contributes rank+1
    params = list(net.parameters())
    fg.grads.fill_(float(rank + 1))
    for p in reversed(params):
        b.ready(p)
    launched = b.finish()
    expect = sum(r + 1 for r in range(world)) / world
    ok = bool(torch.allclose(fg.grads, torch.full_like(fg.grads, expect))) and launched >= 2
    # a second step reuses the bucketer
    fg.grads.fill_(float(2 * rank))
    for p in reversed(params):
        b.ready(p)
    b.finish()
    ok = ok and bool(torch.allclose(fg.grads, torch.full_like(fg.grads, sum(2 * r for r in range(world)) / world)))
    ret[rank] = (ok, launched)
    dist.destroy_process_group()


def test_gradient_bucketing_world2_gloo():
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_bucket_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r][0] for r in range(world)), dict(ret)
    assert ret[0][1] == ret[1][1] >= 2  # same bucket schedule on every rank


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the B200 arm): one JSON line with the contract's keys, timed on
    the host cores with the oracle port of the step; a bounded sample (here 1 step of 2 images at 128x128)."""
    import json
    import subprocess

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--res", "128",
                        "--ref-batch", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == os.cpu_count() and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_bench_clock_sample_parsing(tmp_path):
    """ClockSampler.stop: samples are filtered to the timed region by their nvidia-smi timestamps; throttle reasons are collected."""
    import importlib.util
    import time

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class FakeProc:
        def terminate(self):
            pass

        def wait(self):
            pass

    cs = bench.ClockSampler(0)
    cs.p = FakeProc()
    now = time.time()
    fmt = lambda t: time.strftime("%Y/%m/%d %H:%M:%S", time.localtime(t)) + ".%03d" % int((t % 1) * 1000)
    rows = [(now - 5.0, 1200, "Not Active"), (now - 1.0, 1950, "Active"), (now - 0.5, 1965, "Not Active"), (now + 5.0, 900, "Not Active")]
    with open(cs.f.name, "w") as f:
        for t, mhz, cap in rows:
            f.write(f"{fmt(t)}, 0, {mhz}, 1965, 400.0, 0x0, Not Active, Not Active, Not Active, {cap}\n")
    out = cs.stop(now - 2.0, now)
    assert out["samples"] == 2 and out["samples_total"] == 4 and out["window"] == "timed region"
    assert out["sm_mhz"] in (1950.0, 1965.0) and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]


def test_bench_inputs_and_flop_counts_are_independent_of_the_oracle_but_equal_to_it():
    """The B200 arm of bench.py imports nothing from oracle/ (the oracle is the checker, not part of what is measured): its synthetic
    inputs (data.synthetic_batch) and FLOP counts (bench.conv_flops) are its own — and identical to the oracle's."""
    import importlib.util

    from gan_inpainting_b200.data import synthetic_batch
    from oracle import synth

    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for res in (128, 256, 512):
        g, d = bench.conv_flops(res)
        assert g == onets.conv_flops_unet(res) and d == onets.conv_flops_patchgan(res)
    assert abs(bench.flops_per_image(256, "wgan_l1")["total"] / 1e9 - 61.59) < 0.01  # 3 G + 8 D
    a, b = synthetic_batch(3, 128, seed=9, rank=1), synth.make_batch(3, 128, seed=9, rank=1)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    src = open(os.path.join(ROOT, "bench.py")).read()
    b200_arm = src[src.index("def run_b200("):src.index("def main(")]
    assert "oracle" not in b200_arm.replace("oracle port", "").replace("time_cpu", "")


def test_gpu_load_transform_tap_tables_match_the_oracle():
    """data.GpuLoadTransform computes Pillow's resample taps on the host (the kernel only does the integer MACs): same first index,
    tap count and 22-bit weights as the oracle's restatement, same output size rule as transforms.Resize(int) (train.py:69-72)."""
    import numpy as np
    from gan_inpainting_b200 import data
    from oracle import preprocess

    for n_in, n_out in [(256, 128), (260, 128), (96, 128), (128, 128), (513, 256), (1024, 256), (7, 3)]:
        ksize, first, count, weights = data._triangle_taps(n_in, n_out)
        k2, xmin, xcnt, kk = preprocess.bilinear_coeffs(n_in, n_out)
        assert ksize == k2 and list(xmin) == first and list(xcnt) == count and np.array_equal(kk, np.array(weights, dtype=np.int32))
    t = data.GpuLoadTransform(128, "cpu")
    for h, w in [(300, 260), (260, 300), (128, 128), (64, 200)]:
        assert t.output_size(h, w) == preprocess.resize_output_size(h, w, 128)


def test_flat_group_detects_and_repairs_a_detached_parameter():
    """optim.FlatGroup re-homes parameters into one flat buffer; if something replaces ``p.data`` afterwards (``load_state_dict(assign=True)``,
    a manual ``p.data = ...``) the fused optimizer would update a buffer the module no longer reads.  ``check_alias`` (run by every
    optimizer step via ``rebind_grads`` and by ``flatten_parameters``) copies the new value back into the flat view and re-binds it."""
    import gc

    from gan_inpainting_b200 import optim

    lin = torch.nn.Linear(8, 4)
    fg = optim.flatten_parameters(lin)
    assert optim.flatten_parameters(lin) is fg and lin.weight.data_ptr() == fg.params.data_ptr()
    new_w = torch.full((4, 8), 3.0)
    lin.weight.data = new_w  # detaches the parameter from the flat buffer
    assert lin.weight.data_ptr() != fg.params.data_ptr()
    fg.check_alias()
    assert lin.weight.data_ptr() == fg.params.data_ptr() and float(fg.params[:32].min()) == 3.0
    fg.params.mul_(2.0)  # what a fused optimizer kernel does: the module must see it
    assert float(lin.weight.min()) == 6.0
    # the registry holds the group weakly: dropping the module frees it
    import weakref
    r = weakref.ref(fg)
    del fg, lin
    gc.collect()
    assert r() is None


def test_nccl_group_options_and_trainer_argument_checks(monkeypatch):
    """trainer.nccl_group_options: None unless a cap is asked for (argument or B200_NCCL_CTAS); the extra generator loss terms of
    config 4 are rejected for the global_local recipe before anything touches a device."""
    import pytest
    from gan_inpainting_b200 import trainer

    monkeypatch.delenv("B200_NCCL_CTAS", raising=False)
    assert trainer.nccl_group_options() is None and trainer.nccl_group_options(0) is None
    o = trainer.nccl_group_options(8)
    assert o.config.max_ctas == 8 and o.config.min_ctas == 1
    monkeypatch.setenv("B200_NCCL_CTAS", "4")
    assert trainer.nccl_group_options().config.max_ctas == 4
    with pytest.raises(ValueError):
        trainer.InpaintingTrainer("global_local", None, None, net_D_local=object(), tv_weight=1.0)
    with pytest.raises(ValueError):
        trainer.InpaintingTrainer("nonsense", None, None)


def test_vgg_feature_losses_host_side():
    """Host logic of the VGG19 feature losses (loss.py:49-136): the engine's plan over vgg19.features[0..29] (13 convolutions, 4
    max-pools, ReLUs fused), the first layer's channel-summed filters, and the loud failure on CPU tensors (no fallback)."""
    import pytest
    import torch
    from gan_inpainting_b200 import _lib
    from gan_inpainting_b200.engine import VGG_FEATURE_LAYERS, VggFeatureEngine
    from gan_inpainting_b200.lib.models import loss, networks

    w = networks.VGG19Wrapper()
    eng = w._get_engine()
    assert isinstance(eng, VggFeatureEngine) and eng is w._get_engine()
    kinds = [k for k, _ in eng.plan]
    assert kinds.count("conv") == 13 and kinds.count("pool") == 4 and kinds.count("relu") == 13 and max(i for _, i in eng.plan) == max(VGG_FEATURE_LAYERS)
    assert all(eng.plan[n + 1][0] == "relu" for n, (k, _) in enumerate(eng.plan) if k == "conv")
    with pytest.raises(_lib.B200Error):
        eng.features_of(torch.rand(1, 1, 256, 256))
    with pytest.raises(_lib.B200Error):
        networks.VGG19Wrapper(pretrained=True)  # a download: trained weights come in through state_dict=
    with pytest.raises(NotImplementedError):
        w(torch.rand(1, 3, 32, 32))
    loss.set_vgg(w)
    try:
        with pytest.raises(_lib.B200Error):
            loss.perceptual_and_style_loss(torch.rand(1, 1, 256, 256), torch.rand(1, 1, 256, 256))
        with pytest.raises(_lib.B200Error):
            loss.gram_matrix(torch.rand(1, 64, 8, 8))
    finally:
        loss.set_vgg(None)


def test_every_reference_experiment_script_has_a_drop_in():
    """One module per script of the reference's experiment_list/ (names as of abeytheo/gan-inpainting), each with
    ``begin(state, loaders)``; without CUDA they fail loudly instead of falling back to the CPU."""
    import importlib

    import pytest
    import torch

    names = ["wgan_l1", "wgan_rmse", "minimaxgan_rmse", "minimaxgan_l1", "wgan_perceptual_style_faceparsing", "experiment1_global_local_D",
             "experiment3_global_D"]
    for n in names:
        m = importlib.import_module(f"gan_inpainting_b200.experiment_list.{n}")
        assert callable(m.begin), n
    if not torch.cuda.is_available():
        import tempfile

        from gan_inpainting_b200.experiment_list import wgan_l1

        with tempfile.TemporaryDirectory() as d, pytest.raises(RuntimeError):
            wgan_l1.begin(dict(numepoch=0, saveevery=1, evalevery=1, experiment_dir=d, log_dir=d), {"train": [], "test": []})
