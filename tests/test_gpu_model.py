"""Module-level and end-to-end parity of the CUDA path against the CPU oracle and the golden
reference runs (`-m gpu`).

Bar (BASELINE.json north_star): depth outputs within 1e-3 relative L1 of the reference path on
identical inputs, weights and noise.  The fp32 kernels land around 1e-6..1e-5; the tolerances below
leave one order of magnitude of head-room and are far inside the bar.
"""
import pytest
import torch

from diffmvs_b200 import synth
from diffmvs_b200.models import CasDiffMVS, ContextNet, FeatureNet
from oracle import diffmvs_ref as O
from oracle import spec
from tests.helpers import GOLDEN_CASES, case_setup, load_golden, rel_l1, replay_noise

pytestmark = pytest.mark.gpu
DEV = "cuda"
DEPTH_TOL = 1e-4          # rel-L1 on depth maps (bar: 1e-3)
REPORT = []


def _to_dev(imgs, proj, dv):
    return [i.to(DEV) for i in imgs], {k: v.to(DEV) for k, v in proj.items()}, dv.to(DEV)


def _build(args, sd):
    model = CasDiffMVS(args, test=True)
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    assert all(k.rsplit(".", 1)[-1] in spec.SCHEDULE_BUFFERS for k in missing.missing_keys)
    return model.to(DEV).eval()


def _patch_noise(monkeypatch, draws_fn):
    monkeypatch.setattr(torch, "randn_like", lambda like, **kw: draws_fn(like).to(like.device).view(like.shape))


def test_feature_and_context_nets_match_oracle():
    args = synth.workload_args("cas_tiny")
    sd = synth.synth_state_dict(spec.state_dict_shapes(args), 123)
    imgs, _, _ = synth.workload_inputs("cas_tiny")
    fsd = {k[len("feature."):]: v for k, v in sd.items() if k.startswith("feature.")}
    net = FeatureNet(8, [48, 32, 16])
    net.load_state_dict(fsd)
    out = net.to(DEV).eval()(imgs[1].to(DEV))
    ref = O.feature_net(sd, "feature", imgs[1], True)
    from diffmvs_b200 import ops
    # FFMA (default) agrees to ~1e-7; the tcgen05 modes accumulate in TMEM (truncating adder): ~2e-6
    tol = 5e-6 if ops.get_precision() == "fp32" else 3e-5
    for k in ref:
        assert out[k].shape == ref[k].shape
        assert rel_l1(out[k], ref[k]) < tol, k
    csd = {k[len("context."):]: v for k, v in sd.items() if k.startswith("context.")}
    cnet = ContextNet([32, 64, 36])
    cnet.load_state_dict(csd)
    cout = cnet.to(DEV).eval()(imgs[0].to(DEV))
    cref = O.context_net(sd, "context", imgs[0], True)
    for k in cref:
        assert rel_l1(cout[k], cref[k]) < tol, k


def test_training_mode_and_cpu_are_refused():
    args = synth.workload_args("cfg1")
    model = CasDiffMVS(args, test=True)
    imgs, proj, dv = synth.workload_inputs("cfg1")
    with pytest.raises(RuntimeError):
        model.eval()(imgs, proj, dv)                       # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        model.to(DEV).train()(*_to_dev(imgs, proj, dv))    # training mode


@pytest.mark.parametrize("case", list(GOLDEN_CASES))
def test_end_to_end_matches_golden_reference(case, monkeypatch):
    """CUDA path vs the recorded outputs of the real reference (tests/golden)."""
    g = load_golden(case)
    args, sd, imgs, proj, dv = case_setup(case)
    model = _build(args, sd)
    _patch_noise(monkeypatch, replay_noise(g))
    out = model(*_to_dev(imgs, proj, dv))
    n_depth = len([k for k in g.files if k.startswith("depth_")])
    assert len(out["depth"]) == n_depth and out["conf"] == []
    for i, d in enumerate(out["depth"]):
        ref = torch.from_numpy(g[f"depth_{i}"])
        assert tuple(d.shape) == tuple(ref.shape)
        r = rel_l1(d, ref)
        REPORT.append((case, f"depth[{i}]", r))
        assert r < DEPTH_TOL, (case, i, r)
    for i, c in enumerate(out["photometric_confidence"]):
        ref = torch.from_numpy(g[f"photo_conf_{i}"])
        assert tuple(c.shape) == tuple(ref.shape)
        assert (c.cpu() - ref).abs().mean().item() < 1e-4, (case, i)


@pytest.mark.parametrize("workload,batch", [("cas_small", 1), ("cfg2", 1), ("cas_tiny", 2)])
def test_end_to_end_matches_oracle(workload, batch, monkeypatch):
    """CUDA path vs the CPU oracle, with operator-boundary taps, on cases without golden files."""
    args = synth.workload_args(workload)
    sd = synth.synth_state_dict(spec.state_dict_shapes(args), 123)
    imgs, proj, dv = synth.workload_inputs(workload, seed=1, batch=batch)

    def mk():
        gen = torch.Generator().manual_seed(11)
        return lambda like: torch.randn(like.shape, generator=gen, dtype=torch.float32)

    taps_ref = {}
    with torch.no_grad():
        ref = O.casdiffmvs_forward(sd, args, imgs, proj, dv, randn=mk(), taps=taps_ref)
    model = _build(args, sd)
    _patch_noise(monkeypatch, mk())
    taps = {}
    with torch.no_grad():
        out = model.plan(DEV).forward(*_to_dev(imgs, proj, dv), taps=taps)
    # stage-1 operator taps
    vol = taps["stage1_volume"].permute(0, 4, 1, 2, 3)
    assert rel_l1(vol, taps_ref["stage1_volume"]) < 5e-4
    assert rel_l1(taps["view_weights"], taps_ref["view_weights"]) < 1e-4
    fl_mismatch = (taps["stage1_floor"].cpu().long() != taps_ref["stage1_floor"][:, 0]).float().mean().item()
    REPORT.append((workload, "stage1 floor-index mismatch rate", fl_mismatch))
    assert fl_mismatch < 2e-3
    for key in [k for k in taps_ref if k.endswith("_cost")]:
        got = taps[key].permute(0, 3, 1, 2)
        assert rel_l1(got, taps_ref[key]) < 1e-3, key
    for key in [k for k in taps_ref if k.endswith("_update")]:
        r = rel_l1(taps[key].unsqueeze(1), taps_ref[key])
        REPORT.append((workload, key, r))
        assert r < 2e-3, (key, r)
    for i, (d, r) in enumerate(zip(out["depth"], ref["depth"])):
        e = rel_l1(d, r)
        REPORT.append((workload, f"depth[{i}]", e))
        assert e < DEPTH_TOL, (workload, i, e)
    for i, (c, r) in enumerate(zip(out["photometric_confidence"], ref["photometric_confidence"])):
        assert (c.cpu() - r).abs().mean().item() < 1e-4, (workload, "conf", i)


def test_noise_comes_from_default_cuda_generator():
    """Two forwards with the same CUDA seed agree BIT FOR BIT (GroupNorm statistics are accumulated as fixed-point
    integers, so no result depends on the order in which thread blocks arrive - like the reference); different seeds
    differ (update.py:472)."""
    args = synth.workload_args("cfg1")
    sd = synth.synth_state_dict(spec.state_dict_shapes(args), 123)
    model = _build(args, sd)
    inp = _to_dev(*synth.workload_inputs("cfg1"))
    torch.manual_seed(5)
    a = model(*inp)["depth"][-1].clone()
    torch.manual_seed(5)
    b = model(*inp)["depth"][-1].clone()
    torch.manual_seed(6)
    c = model(*inp)["depth"][-1].clone()
    assert torch.equal(a, b)
    assert rel_l1(a, c) > 1e-3


@pytest.mark.parametrize("workload", ["cfg3"])
def test_full_size_against_oracle_on_device(workload, monkeypatch):
    """BASELINE.json headline size.  The CPU oracle takes ~12 s/run here, so the restatement is run with
    stock torch CUDA ops in strict fp32 (TF32 off) as the reference; plus size-independent
    properties: finite, inside the depth range, deterministic stage-1 output."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    args = synth.workload_args(workload)
    sd = synth.synth_state_dict(spec.state_dict_shapes(args), 123)
    imgs, proj, dv = synth.workload_inputs(workload)
    dimgs, dproj, ddv = _to_dev(imgs, proj, dv)

    def mk():
        gen = torch.Generator().manual_seed(3)
        return lambda like: torch.randn(like.shape, generator=gen, dtype=torch.float32)

    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    with torch.no_grad():
        r = mk()
        ref = O.casdiffmvs_forward(sd_dev, args, dimgs, dproj, ddv, randn=lambda like: r(like).to(DEV))
    model = _build(args, sd)
    _patch_noise(monkeypatch, mk())
    out = model(dimgs, dproj, ddv)
    for i, (d, rr) in enumerate(zip(out["depth"], ref["depth"])):
        assert torch.isfinite(d).all()
        assert d.min().item() >= synth.DEPTH_MIN - 1e-2 and d.max().item() <= synth.DEPTH_MAX + 1e-2
        e = rel_l1(d, rr)
        REPORT.append((workload, f"depth[{i}] vs torch-CUDA fp32 oracle", e))
        assert e < 1e-3, (i, e)
    assert tuple(out["depth"][-1].shape) == (1, imgs[0].shape[2], imgs[0].shape[3])

@pytest.mark.parametrize("workload", ["cfg3", "cfg4"])
def test_full_size_against_cpu_fp32_oracle(workload, monkeypatch):
    """BASELINE.json configs[2] (cfg3: DTU 1600x1152, 7 views, D_init 48) and configs[3] (cfg4: Tanks & Temples shaped
    1920x1024, 11 views, D_init 96, `scale 0 .125 .025`, `/root/reference/scripts/test/test_tank_casdiffmvs.sh:11-17`)
    end to end against the CPU fp32 oracle (the restatement pinned bit-exact to the reference), same inputs, weights
    and noise.  Records depth rel-L1 per output and the stage-1 floor-index mismatch rate (north_star: <= 1e-3 rel-L1,
    hypothesis indices exact)."""
    args = synth.workload_args(workload)
    if workload == "cfg4":
        assert args.numdepth_initial == 96 and args.scale == [0.0, 0.125, 0.025] and args.ddim_eta == [0, 1, 1]
    sd = synth.synth_state_dict(spec.state_dict_shapes(args), 123)
    imgs, proj, dv = synth.workload_inputs(workload)

    def mk():
        gen = torch.Generator().manual_seed(3)
        return lambda like: torch.randn(like.shape, generator=gen, dtype=torch.float32)

    taps_ref = {}
    with torch.no_grad():
        ref = O.casdiffmvs_forward(sd, args, imgs, proj, dv, randn=mk(), taps=taps_ref)
    keep = {k: taps_ref[k] for k in ("stage1_floor", "view_weights")}
    taps_ref.clear()
    model = _build(args, sd)
    _patch_noise(monkeypatch, mk())
    taps = {}
    with torch.no_grad():
        out = model.plan(DEV).forward(*_to_dev(imgs, proj, dv), taps=taps)
    mism = (taps["stage1_floor"].cpu().long() != keep["stage1_floor"][:, 0]).float().mean().item()
    REPORT.append((workload, "stage1 floor-index mismatch rate vs CPU fp32 oracle", mism))
    assert mism < 2e-3, mism
    assert rel_l1(taps["view_weights"], keep["view_weights"]) < 1e-4
    assert len(out["depth"]) == len(ref["depth"]) == 6
    for i, (d, r) in enumerate(zip(out["depth"], ref["depth"])):
        assert tuple(d.shape) == tuple(r.shape)
        assert torch.isfinite(d).all()
        e = rel_l1(d, r)
        REPORT.append((workload, f"depth[{i}] vs CPU fp32 oracle", e))
        assert e < DEPTH_TOL, (workload, i, e)
    for i, (c, r) in enumerate(zip(out["photometric_confidence"], ref["photometric_confidence"])):
        assert (c.cpu() - r).abs().mean().item() < 1e-4, (workload, "conf", i)


@pytest.mark.parametrize("mode,tol", [("ws_tf32x3", DEPTH_TOL), ("ws2_tf32x3", DEPTH_TOL), ("ws2_f16c", DEPTH_TOL),
                                      ("fp32", DEPTH_TOL), ("ws_tf32", 1e-2)])
@pytest.mark.parametrize("workload", ["cas_small", "cfg2"])
def test_tensor_core_modes_meet_the_parity_bar(workload, mode, tol, monkeypatch):
    """The convolution modes against the CPU oracle.  The 3xTF32 modes (operand split) and the FFMA mode must stay in the
    fp32 class.  Plain "ws_tf32" (torch/cuDNN default numerics) does NOT meet the north_star bar of 1e-3 on these
    weights (measured 1e-3..3e-3 rel-L1, 2-6 % stage-1 index flips, profiles/r1_parity_report.txt), which is why it
    is an opt-in mode and never the default; the test only guards against gross breakage (1e-2) and records
    the numbers."""
    from diffmvs_b200 import ops
    args = synth.workload_args(workload)
    sd = synth.synth_state_dict(spec.state_dict_shapes(args), 123)
    imgs, proj, dv = synth.workload_inputs(workload, seed=2)

    def mk():
        gen = torch.Generator().manual_seed(13)
        return lambda like: torch.randn(like.shape, generator=gen, dtype=torch.float32)

    taps_ref = {}
    with torch.no_grad():
        ref = O.casdiffmvs_forward(sd, args, imgs, proj, dv, randn=mk(), taps=taps_ref)
    model = _build(args, sd)
    _patch_noise(monkeypatch, mk())
    old = ops.get_precision()
    ops.set_precision(mode)
    try:
        taps = {}
        with torch.no_grad():
            out = model.plan(DEV).forward(*_to_dev(imgs, proj, dv), taps=taps)
    finally:
        ops.set_precision(old)
    mism = (taps["stage1_floor"].cpu().long() != taps_ref["stage1_floor"][:, 0]).float().mean().item()
    REPORT.append((workload, f"[{mode}] stage1 floor-index mismatch rate", mism))
    for i, (d, r) in enumerate(zip(out["depth"], ref["depth"])):
        e = rel_l1(d, r)
        REPORT.append((workload, f"[{mode}] depth[{i}]", e))
        assert e < tol, (workload, mode, i, e)



def test_cuda_graph_replay_matches_eager():
    """`use_cuda_graph()`: the captured forward returns what the eager forward returns for the same generator
    state, keeps consuming the default CUDA generator (two replays differ), and follows new inputs."""
    args = synth.workload_args("cas_tiny")
    sd = synth.synth_state_dict(spec.state_dict_shapes(args), 123)
    imgs, proj, dv = _to_dev(*synth.workload_inputs("cas_tiny"))
    model = _build(args, sd)
    torch.manual_seed(11)
    eager = model(imgs, proj, dv)
    eager2 = model(imgs, proj, dv)
    model.use_cuda_graph(True)
    torch.manual_seed(11)
    g1 = model(imgs, proj, dv)          # captures (warm-up restores the generator), then replays
    g2 = model(imgs, proj, dv)
    for a, b in zip(eager["depth"], g1["depth"]):
        assert rel_l1(a, b) < 1e-5
    for a, b in zip(eager2["depth"], g2["depth"]):
        assert rel_l1(a, b) < 1e-5
    assert rel_l1(g1["depth"][-1], g2["depth"][-1]) > 1e-6      # fresh noise per replay
    imgs2, proj2, dv2 = _to_dev(*synth.workload_inputs("cas_tiny", seed=5))
    model.use_cuda_graph(False)
    torch.manual_seed(12)
    e3 = model(imgs2, proj2, dv2)
    model.use_cuda_graph(True)
    torch.manual_seed(12)
    g3 = model(imgs2, proj2, dv2)
    assert rel_l1(e3["depth"][-1], g3["depth"][-1]) < 1e-5
    assert g3["depth"][-1].data_ptr() != g1["depth"][-1].data_ptr()   # results are the caller's own tensors


def test_feature_cache_against_the_oracle(monkeypatch):
    """Cross-ref-view feature cache against the CPU ORACLE, not against the uncached CUDA path:
    reference view B = source view 1 of call A, its pyramids come from call A's `return_features` (computed there in
    other batch positions); every cache pattern (full / partial / mixed / none) is checked against the oracle in eager
    mode and against the eager call under CUDA-graph replay."""
    args = synth.workload_args("cas_tiny")
    sd = synth.synth_state_dict(spec.state_dict_shapes(args), 123)
    imgs, proj, dv = synth.workload_inputs("cas_tiny")
    order = [1, 0, 2]                                           # call B: view 1 becomes the reference
    imgs_b = [imgs[v] for v in order]
    proj_b = {k: p[:, order].contiguous() for k, p in proj.items()}

    def mk():
        gen = torch.Generator().manual_seed(17)
        return lambda like: torch.randn(like.shape, generator=gen, dtype=torch.float32)

    with torch.no_grad():
        ref = O.casdiffmvs_forward(sd, args, imgs_b, proj_b, dv, randn=mk())
    model = _build(args, sd)
    d_imgs, d_proj, d_dv = _to_dev(imgs, proj, dv)
    d_imgs_b, d_proj_b, _ = _to_dev(imgs_b, proj_b, dv)
    with torch.no_grad():
        a = model(d_imgs, d_proj, d_dv, return_features=True)
    feats_a = a["features"]
    assert len(feats_a) == 3 and set(feats_a[0]) == {"stage1", "stage2", "stage3"}
    assert feats_a[0]["stage1"].data_ptr() != feats_a[1]["stage1"].data_ptr()       # the caller's own copies
    feats_b = [feats_a[v] for v in order]
    patterns = (("full", feats_b), ("partial", [None] + feats_b[1:]), ("mixed", [feats_b[0], None, feats_b[2]]), ("none", None))
    # eager, noise replayed from the oracle run: every cache pattern against the oracle
    for name, feats in patterns:
        _patch_noise(monkeypatch, mk())
        with torch.no_grad():
            out = model(d_imgs_b, d_proj_b, d_dv, features=feats)
        for i, (d, r) in enumerate(zip(out["depth"], ref["depth"])):
            e = rel_l1(d, r)
            REPORT.append(("cas_tiny", f"feature cache [{name}] depth[{i}] vs oracle", e))
            assert e < DEPTH_TOL, (name, i, e)
    monkeypatch.undo()
    # CUDA-graph replay (noise from the CUDA generator, which a graph can consume): equal to the eager call with the same
    # seed for every pattern (a different batch size may pick another back end: last-ulp differences allowed)
    for name, feats in patterns:
        model.use_cuda_graph(False)
        torch.manual_seed(31)
        eager = model(d_imgs_b, d_proj_b, d_dv, features=feats)
        model.use_cuda_graph(True)
        torch.manual_seed(31)
        graphed = model(d_imgs_b, d_proj_b, d_dv, features=feats)      # captures, then replays
        torch.manual_seed(31)
        again = model(d_imgs_b, d_proj_b, d_dv, features=feats)
        for x, y, z in zip(eager["depth"], graphed["depth"], again["depth"]):
            assert rel_l1(y, x) < 1e-5 and torch.equal(y, z), name
    model.use_cuda_graph(False)
    with pytest.raises(ValueError):
        bad = [dict(f) for f in feats_b]
        bad[1]["stage1"] = bad[1]["stage1"][:, :-1]
        model(d_imgs_b, d_proj_b, d_dv, features=bad)


def test_scan_runner_reuses_pyramids():
    """`scan.ScanRunner`: a sliding window of reference views encodes each image once; outputs equal the uncached call."""
    from diffmvs_b200.scan import ScanRunner
    args = synth.workload_args("cas_tiny")
    sd = synth.synth_state_dict(spec.state_dict_shapes(args), 123)
    model = _build(args, sd)
    imgs, proj, dv = _to_dev(*synth.workload_inputs("cas_tiny"))
    pool = imgs + [i.flip(-1).contiguous() for i in imgs]                 # six distinct images
    runner = ScanRunner(model, capacity=8)
    for start in range(4):
        ids = [start, start + 1, start + 2]
        batch = [pool[i] for i in ids]
        torch.manual_seed(100 + start)
        got = runner(ids, batch, proj, dv)
        torch.manual_seed(100 + start)
        want = model(batch, proj, dv)
        assert all(rel_l1(x, y) < 1e-5 for x, y in zip(got["depth"], want["depth"]))
    assert runner.cache.misses == 6 and runner.cache.hits == 6       # 3 new images in the first call, then one per call


def test_zz_report():
    """Prints the collected parity numbers (kept in gpurun_out/ when run on the GPU box)."""
    import os
    lines = [f"{w:12s} {k:45s} {v:.3e}" for w, k, v in REPORT]
    text = "\n".join(lines)
    print("\n" + text)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_report.txt", "w") as f:
        f.write(text + "\n")
