"""Host-side logic of the width-stacked tcgen05 convolution that runs without a GPU: the tile planner
(`dmvs_conv_ws_plan`), the back-end query and the save / load round trip of the autotuning table."""
import ctypes as C
import json
import os

import pytest

from diffmvs_b200 import _cabi, ops


def _desc(N, cin, cout, k, stride, dims, gn=False):
    d = _cabi.ConvDesc()
    D, H, W = dims
    kd, kh, kw = k
    pd, ph, pw = kd // 2, kh // 2, kw // 2
    d.x = d.w = d.w_ws = d.y = d.bias = 0x1000           # only alignment is inspected
    d.N, d.D, d.H, d.W, d.C1, d.C2 = N, D, H, W, cin, 0
    d.x_ps, d.y_ps = cin, cout
    d.KD, d.KH, d.KW, d.stride, d.pad_d, d.pad_h, d.pad_w = kd, kh, kw, stride, pd, ph, pw
    d.Do = (D + 2 * pd - kd) // stride + 1
    d.Ho = (H + 2 * ph - kh) // stride + 1
    d.Wo = (W + 2 * pw - kw) // stride + 1
    d.Cout = cout
    d.precision = _cabi.PREC_WS_TF32X3
    if gn:
        d.in_stats = d.in_g1 = d.in_g0 = 0x1000
    return d


LAYERS = [
    (7, 8, 8, (1, 3, 3), 1, (1, 1152, 1600)),
    (7, 8, 16, (1, 5, 5), 2, (1, 1152, 1600)),
    (7, 64, 16, (1, 3, 3), 1, (1, 576, 800)),
    (7, 64, 64, (1, 3, 3), 1, (1, 144, 200)),
    (1, 64, 16, (1, 7, 7), 1, (1, 288, 400)),
    (1, 64, 64, (1, 1, 5), 1, (1, 144, 200)),
    (1, 52, 40, (1, 5, 1), 1, (1, 144, 200)),
    (6, 4, 8, (3, 3, 3), 1, (48, 144, 200)),
    (1, 8, 16, (3, 3, 3), 2, (48, 144, 200)),
    (1, 32, 32, (1, 3, 3), 1, (1, 16, 20)),
    (1, 64, 144, (1, 3, 3), 1, (1, 36, 50)),
]


@pytest.mark.parametrize("N,cin,cout,k,stride,dims", LAYERS)
def test_tile_plan_respects_the_hardware_budgets(N, cin, cout, k, stride, dims):
    lib = _cabi.lib()
    d = _desc(N, cin, cout, k, stride, dims)
    assert lib.dmvs_conv_backends(C.byref(d)) & 8, "the width-stacked back end should accept this layer"
    out = (C.c_int32 * 64)()
    n = lib.dmvs_conv_ws_plan(C.byref(d), out, 8)
    assert n >= 1
    covered = 0
    for i in range(n):
        CC, Nn, TH, TW, n_blk, R, ctas, smem = out[8 * i:8 * i + 8]
        assert CC % 8 == 0 and 8 <= CC <= 64 and Nn % 16 == 0 and Nn <= 256
        assert n_blk * Nn <= (256 if ctas >= 2 else 512), "accumulators must fit the TMEM columns of a CTA"
        assert smem <= (112 if ctas >= 2 else 216) * 1024, "shared memory budget of co-resident CTAs"
        assert R in (2, 3) and TH >= 1 and TW >= 1 and ctas in (1, 2)
        covered += CC
    assert covered >= cout


def test_unsupported_layers_are_refused():
    lib = _cabi.lib()
    d = _desc(1, 3, 8, (1, 3, 3), 1, (1, 64, 64))          # 3 input channels: no 16-byte channel groups
    assert not lib.dmvs_conv_backends(C.byref(d)) & 8
    out = (C.c_int32 * 8)()
    assert lib.dmvs_conv_ws_plan(C.byref(d), out, 1) == -3
    d = _desc(1, 16, 20, (1, 3, 3), 1, (1, 64, 64))
    d.out_stats = 0x1000                                    # GroupNorm groups of 5 channels: pairs would straddle
    assert not lib.dmvs_conv_backends(C.byref(d)) & 8


def test_autotune_table_round_trip(tmp_path):
    ops._TUNED.clear()
    sig = (7, 1, 576, 800, 64, 0, 16, 1, 3, 3, 1, 0, 1, 1, 0, False, 0, 0, 0, 64, 0, 16, 1, True)
    assert len(sig) == ops._SIG_LEN
    ops._TUNED[(3,) + sig] = (ops.PREC_WS_TF32X3, {ops.PREC_FP32: 1.39, ops.PREC_WS_TF32X3: 0.61})   # tuned on cuda:3
    path = tmp_path / "tuned.json"
    ops.save_tuned(str(path))
    rows = json.load(open(path))
    assert rows[0]["choice"] == "ws_tf32x3"
    ops._TUNED.clear()
    assert ops.load_tuned(str(path)) == 1
    assert ops._tuned_lookup(0, sig)[0] == ops.PREC_WS_TF32X3      # a loaded table applies to every device
    assert ops._tuned_lookup(5, sig)[0] == ops.PREC_WS_TF32X3
    ops._TUNED.clear()
    json.dump([{"sig": list(sig[:22]), "choice": "ws_tf32x3", "ms": {}}], open(path, "w"))   # round-1 format: skipped
    assert ops.load_tuned(str(path)) == 0


def test_span_overlap_detection():
    import torch
    buf = torch.zeros(4, 6, 8)
    a, b = buf[..., :4], buf[..., 4:]                     # channel slices of one buffer: spans overlap, bytes do not
    assert ops._disjoint(a, b) and ops._disjoint(b, a)
    assert not ops._disjoint(buf[..., :5], buf[..., 4:])
    assert not ops._disjoint(buf, a)
    assert ops._disjoint(buf[:2], buf[2:])
    assert ops._disjoint(buf, None, torch.zeros(3))


def test_shipped_tuned_table_loads():
    """The per-layer back-end table shipped for B200 matches the current signature format and names known modes."""
    ops._TUNED.clear()
    assert os.path.exists(ops.DEFAULT_TUNED_TABLE)
    rows = json.load(open(ops.DEFAULT_TUNED_TABLE))
    assert len(rows) > 100 and all(len(r["sig"]) == ops._SIG_LEN and r["choice"] in ops.PRECISIONS for r in rows)
    assert ops.load_tuned(ops.DEFAULT_TUNED_TABLE) == len(rows)
    # the headline layer 16 -> 16 3x3 at 576 x 800 x 7 views (FeatureNet conv1.1) has an entry, on a tcgen05 back end
    hits = [v for k, v in ops._TUNED.items() if k[1:8] == (7, 1, 576, 800, 16, 0, 16) and k[8:11] == (1, 3, 3)]
    assert hits and all(c in (ops.PREC_WS2_TF32X3, ops.PREC_WS2_TF32_F16C, ops.PREC_WS_TF32X3) for c, _ in hits)
    ops._TUNED.clear()
