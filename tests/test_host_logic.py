"""Host-side logic without a GPU: weight packing, parity / phase / window views and the whole engine graph,
executed by the CPU emulator of the kernel's addressing (tests/emulate.py) and compared with the oracle."""

import pytest
import torch

import conv_cases
import emulate
from oracle import unet_oracle
from robosat_b200 import synth
from robosat_b200.engine import UNetEngine, choose_block_n, choose_tile, pack_upsample_phases


@pytest.mark.parametrize("i", range(len(conv_cases.default_cases("cpu"))))
def test_single_conv_descriptor_semantics(i):
    case = conv_cases.default_cases("cpu")[i]()
    emulate.run_desc(case.desc)
    got, ref = case.result(), case.ref()
    err = (got - ref).abs().max().item()
    # fp16 output rounding (half an ulp at |x| < 16 is 4e-3) + fp16 rounding of pre-summed phase weights
    assert err <= 6e-3 * max(1.0, ref.abs().max().item() / 8), (case.name, err)


@pytest.mark.parametrize("i", range(len(conv_cases.split_cases("cpu"))))
def test_split_conv_descriptor_semantics(i):
    """strict precision: hi/lo planes in, hi/lo planes out; the emulated descriptor matches a float64 reference to fp32 round-off"""
    case = conv_cases.split_cases("cpu")[i]()
    emulate.run_desc(case.desc)
    got, ref = case.result(), case.ref()
    assert (got - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item()), case.name


@pytest.mark.parametrize("i", range(len(conv_cases.row_cases("cpu"))))
def test_row_conv_descriptor_semantics(i):
    case = conv_cases.row_cases("cpu")[i]()
    emulate.run_rowdesc(case.desc)
    got, ref = case.result(), case.ref()
    assert (got - ref).abs().max().item() <= 6e-3 * max(1.0, ref.abs().max().item() / 8), case.name


@pytest.mark.parametrize("i", range(len(conv_cases.row_split_cases("cpu"))))
def test_row_split_descriptor_semantics(i):
    case = conv_cases.row_split_cases("cpu")[i]()
    emulate.run_rowdesc(case.desc)
    got, ref = case.result(), case.ref()
    assert (got - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item()), case.name


def test_strict_plan_uses_the_line_buffer_for_the_head_only():
    sd = synth.make_state_dict(2, seed=0)
    eng = UNetEngine(sd, 2, 1, 64, 512, device="cpu", plan_only=True, precision="strict")
    kinds = [(op[1].name, type(op[1]).__name__) for op in eng.ops if op[0] == "conv"]
    assert [n for n, k in kinds if k == "RowConvOp"] == ["dec5+final"]
    x = synth.normalize_tiles(synth.make_tiles_u8(1, 512, seed=1)[:, :64].contiguous())
    got = emulate.run_engine(eng, x)
    with torch.no_grad():
        ref = unet_oracle.unet_forward(sd, x)
    assert ((got - ref).pow(2).sum().sqrt() / ref.pow(2).sum().sqrt()).item() < 2e-5


def test_engine_graph_with_line_buffer_layers_matches_oracle():
    """512-wide input: layer1 3x3, dec4 and dec5 + final take the line-buffer plan; same result as the oracle"""
    sd = synth.make_state_dict(2, seed=0)
    x = synth.normalize_tiles(synth.make_tiles_u8(1, 512, seed=1)[:, :64].contiguous())  # 64 x 512 strip keeps the CPU run short
    eng = UNetEngine(sd, 2, 1, 64, 512, device="cpu", plan_only=True, precision="fast")
    kinds = [type(op[1]).__name__ for op in eng.ops if op[0] == "conv"]
    assert kinds.count("RowConvOp") == 5  # layer1.{0,1,2}.conv2, dec4, dec5+final
    got = emulate.run_engine(eng, x)
    with torch.no_grad():
        ref = unet_oracle.unet_forward(sd, x)
    assert ((got - ref).pow(2).sum().sqrt() / ref.pow(2).sum().sqrt()).item() < 5e-3


def test_upsample_phase_decomposition_is_exact_in_fp64():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(8, 5, 3, 3, generator=g, dtype=torch.float64)
    x = torch.randn(2, 5, 6, 7, generator=g, dtype=torch.float64)
    ref = torch.nn.functional.conv2d(torch.nn.functional.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)
    wp = pack_upsample_phases(w).reshape(4, 8, 2, 2, 5)
    out = torch.zeros_like(ref)
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1))
    for a in range(2):
        for b in range(2):
            acc = 0
            for th in range(2):
                for tw in range(2):
                    patch = xp[:, :, a + th:a + th + 6, b + tw:b + tw + 7]
                    acc = acc + torch.einsum("nchw,oc->nohw", patch, wp[2 * a + b, :, th, tw, :])
            out[:, :, a::2, b::2] = acc
    assert (out - ref).abs().max().item() < 1e-12


def test_tile_and_block_choice():
    for dims in [(256, 256, 32), (16, 16, 32), (8, 8, 32), (4, 4, 2), (72, 72, 1), (9, 9, 2), (5, 5, 3)]:
        tw, th, tn = choose_tile(*dims)
        assert tw * th * tn == 128
    assert choose_block_n(2048, 64, 1) >= 128
    assert choose_block_n(32, 10, 1) == 32
    assert choose_block_n(64, 100000, 1) == 64
    # small layers: one round of wide tiles beats several rounds of narrow ones (N <= 64 MMAs run at half rate)
    assert choose_block_n(512, 64, 1, kblocks=72) >= 128    # layer4 3x3 at batch 32 x 512^2
    assert choose_block_n(256, 16, 4, kblocks=128) >= 128   # center
    # short-K, store-bound layers keep enough tiles to fill the machine
    assert choose_block_n(256, 4096, 1, kblocks=1) in (64, 128, 256)


@pytest.mark.parametrize("precision", ["fast", "strict"])
@pytest.mark.parametrize("size,batch,classes", [(64, 2, 2), (128, 1, 6)])
def test_engine_graph_matches_oracle(size, batch, classes, precision):
    """fast: fp16 operands (~2e-3); strict: hi/lo operand pairs -- the plan's host logic (weight split + scaling, plane
    strides, pair max-pool) reproduces the fp32 oracle to fp32 round-off"""
    sd = synth.make_state_dict(classes, seed=0)
    x = synth.normalize_tiles(synth.make_tiles_u8(batch, size, seed=1))
    eng = UNetEngine(sd, classes, batch, size, size, device="cpu", plan_only=True, precision=precision)
    got = emulate.run_engine(eng, x)
    with torch.no_grad():
        ref, feats = unet_oracle.unet_forward(sd, x, return_features=True)
    rel_l2 = ((got - ref).pow(2).sum().sqrt() / ref.pow(2).sum().sqrt()).item()
    tol_logits, tol_feat = (5e-3, 3e-3) if precision == "fast" else (2e-5, 1e-5)
    assert rel_l2 < tol_logits, rel_l2
    if precision == "strict":
        assert (got - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
        assert int((got.argmax(1) != ref.argmax(1)).sum()) <= 2
    for name in ("stem", "enc0", "enc1", "enc2", "enc3", "enc4", "center", "dec0", "dec1", "dec2", "dec3", "dec4"):
        a, b = eng.feature_nchw(name), feats[name]
        r = ((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt()).item()
        assert r < tol_feat, (name, r)
    assert len(eng.ops) == 59  # 1 pre-pass + 2 max pools + 56 conv launches (downsample fused into conv3, final into dec5)


def test_train_engine_plan_matches_autograd_on_cpu():
    """The whole training plan (forward in train mode + backward op lists) replayed by the CPU emulator:
    forward vs the oracle's train-mode forward, parameter gradients vs autograd through the mask-frozen fp32 network."""
    import linearized
    from robosat_b200.train_engine import UNetTrainEngine

    C, B, S = 2, 2, 64
    sd0 = {k[7:]: v.clone() for k, v in synth.make_state_dict(C, seed=0).items()}
    x = synth.normalize_tiles(synth.make_tiles_u8(B, S, seed=1))
    dlogits = torch.randn((B, C, S, S), generator=torch.Generator().manual_seed(3)) * 1e-3
    params = {k: v.clone() for k, v in sd0.items()}
    eng = UNetTrainEngine(params, C, B, S, S, device="cpu", plan_only=True, loss_scale=1024.0)
    emulate.run_train_ops(eng, eng.fwd_ops, x=x)
    sd = {k: v.clone() for k, v in sd0.items()}
    with torch.no_grad():
        ref = unet_oracle.unet_forward_train(sd, x)
    assert ((eng.logits - ref).pow(2).sum().sqrt() / ref.pow(2).sum().sqrt()).item() < 5e-3
    assert torch.allclose(params["resnet.bn1.running_mean"], sd["resnet.bn1.running_mean"], atol=1e-3)
    assert int(params["resnet.bn1.num_batches_tracked"]) == int(sd0["resnet.bn1.num_batches_tracked"]) + 1
    emulate.run_train_ops(eng, eng.bwd_ops, dlogits=dlogits)
    sdg = {k: v.clone() for k, v in sd0.items()}
    for k, v in sdg.items():
        if v.dtype == torch.float32 and "running" not in k:
            v.requires_grad_(True)
    (linearized.forward(eng, sdg, x) * dlogits).sum().backward()
    rels = [((eng.grads[k] - v.grad).pow(2).sum().sqrt() / v.grad.pow(2).sum().sqrt().clamp_min(1e-30)).item()
            for k, v in sdg.items() if v.requires_grad and not k.startswith("resnet.fc")]
    assert len(rels) == 168 and max(rels) < 6e-2 and sorted(rels)[len(rels) // 2] < 3e-2, (max(rels), sorted(rels)[len(rels) // 2])


def test_cta_pair_policy_respects_kernel_constraints():
    """Every descriptor the predict plan marks for CTA pairs satisfies what rsb_conv_plan_create demands (mode 0, block_n >= 128,
    at least two spatial tiles) and the policy is on by default for the long-K wide layers (dec0/dec1/dec3, 3x3 of layer2-4)."""
    from robosat_b200 import synth
    from robosat_b200.engine import UNetEngine

    eng = UNetEngine(synth.make_state_dict(2, seed=0), 2, 32, 512, 512, device="cpu", plan_only=True, precision="fast", plan_overrides={})
    paired = {}
    for op in eng.ops:
        if op[0] != "conv" or not hasattr(op[1].desc, "nseg"):
            continue
        d = op[1].desc
        kblocks = sum(d.segs[i].cblocks for i in range(d.nseg))
        tiles = -(-d.Wt // d.TW) * -(-d.Ht // d.TH) * -(-d.Nt // d.TN)
        if d.cta_pair:
            assert d.mode == 0 and d.block_n >= 128 and tiles >= 2 and d.Cout % d.block_n == 0, op[1].name
            assert (d.block_n == 256 and kblocks >= 6) or (d.block_n == 128 and kblocks >= 18), op[1].name
            paired[op[1].name] = d.block_n
    for name in ("dec0", "dec1", "dec3", "resnet.layer2.1.conv2", "resnet.layer3.2.conv2", "resnet.layer4.1.conv2"):
        assert name in paired, name
    assert "dec2" not in paired and "stem" not in paired  # Cout = 64: narrow tiles stay on one CTA


def test_usable_cores_honours_the_cgroup_quota(monkeypatch, tmp_path):
    """A container's CPU quota (cgroup cpu.max), not the hardware-thread count, sizes the codec pools"""
    import builtins

    from robosat_b200 import hostinfo

    real_open = builtins.open

    def fake(content):
        def _open(path, *a, **k):
            if path == "/sys/fs/cgroup/cpu.max":
                f = tmp_path / "cpu.max"
                f.write_text(content)
                return real_open(f, *a, **k)
            return real_open(path, *a, **k)
        return _open

    monkeypatch.setattr(hostinfo.os, "sched_getaffinity", lambda pid: set(range(128)), raising=False)
    monkeypatch.setattr(builtins, "open", fake("1600000 100000\n"))
    assert hostinfo.usable_cores() == 16
    monkeypatch.setattr(builtins, "open", fake("max 100000\n"))
    assert hostinfo.usable_cores() == 128
    monkeypatch.setattr(builtins, "open", fake("50000 100000\n"))
    assert hostinfo.usable_cores() == 1


def test_plan_overrides_change_the_tiles_not_the_result():
    """engine.plan_table() / plan_overrides: a measured (block_n, CTA pair) choice replaces the modelled one where it is valid for
    the layer, is ignored where it is not, and the emulated plan computes the same logits either way"""
    from robosat_b200 import synth
    from robosat_b200.engine import UNetEngine

    sd = synth.make_state_dict(2, seed=0)
    x = synth.normalize_tiles(synth.make_tiles_u8(1, 64, seed=1))
    ov = {"resnet.layer1.0.conv1": {"block_n": 32, "cta_pair": 0},        # Cout 64: narrower tile
          "resnet.layer2.0.conv2": {"block_n": 64, "cta_pair": 0},
          "resnet.layer3.0.conv1": {"block_n": 96, "cta_pair": 0},        # not a tile width: ignored
          "resnet.layer4.1.conv2": {"block_n": 256, "cta_pair": 1},       # 1 x 2 x 2 pixels = one spatial tile: no pair possible, ignored
          "dec2": {"block_n": 128, "cta_pair": 0}}                        # Cout 64 is not a multiple of 128: ignored
    base = UNetEngine(sd, 2, 1, 64, 64, device="cpu", plan_only=True, precision="fast", plan_overrides={})
    tuned = UNetEngine(sd, 2, 1, 64, 64, device="cpu", plan_only=True, precision="fast", plan_overrides=ov)
    db = {op[1].name: op[1].desc for op in base.ops if op[0] == "conv"}
    dt = {op[1].name: op[1].desc for op in tuned.ops if op[0] == "conv"}
    assert dt["resnet.layer1.0.conv1"].block_n == 32 and dt["resnet.layer2.0.conv2"].block_n == 64
    for name in ("resnet.layer3.0.conv1", "resnet.layer4.1.conv2", "dec2"):
        assert (dt[name].block_n, dt[name].cta_pair) == (db[name].block_n, db[name].cta_pair), name
    assert torch.equal(emulate.run_engine(base, x), emulate.run_engine(tuned, x))


def test_multi_rank_validation_covers_the_reference_tile_set():
    """rs train on N ranks (TOML batch split over the ranks, DistributedSampler + loader, both drop_last) evaluates exactly the
    validation tiles the reference's one DataLoader(batch_size=B, shuffle=False, drop_last=True) does, each once"""
    from torch.utils.data import DataLoader
    from torch.utils.data.distributed import DistributedSampler

    for n, world, B in ((103, 2, 16), (37, 4, 8), (64, 8, 16), (15, 2, 16), (1000, 8, 32)):
        data = list(range(n))
        ref = [int(i) for batch in DataLoader(data, batch_size=B, shuffle=False, drop_last=True) for i in batch]
        seen = []
        for rank in range(world):
            sampler = DistributedSampler(data, num_replicas=world, rank=rank, shuffle=False, drop_last=True)
            seen += [int(i) for batch in DataLoader(data, batch_size=B // world, sampler=sampler, drop_last=True) for i in batch]
        assert sorted(seen) == ref, (n, world, B)
