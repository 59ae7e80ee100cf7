"""Torch-tensor front end of the C-ABI kernels.

Tensors handled here are channels-last in *shape*: maps are ``[N,H,W,C]``, volumes ``[N,D,H,W,C]``,
possibly channel slices of a wider buffer (``t[..., a:b]``), which the kernels address through a
pixel stride.  PyTorch only provides device memory, the caching allocator and the current stream.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _cabi
import os

from ._cabi import (ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SILU, ACT_TANH, EPI_GRU_Q, EPI_GRU_ZR, EPI_STD,  # noqa: F401
                    PREC_AUTO, PREC_FP32, PREC_TC_TF32, PREC_TC_TF32X3, PREC_TF32, PREC_TF32X3, PREC_WS_TF32, PREC_WS_TF32X3,
                    PREC_WS2_TF32X3, PREC_WS2_TF32_F16C,
                    RES_NONE, RES_POST_ACT, RES_PRE_ACT, ConvDesc, check)

# Arithmetic of the convolutions (storage is always fp32).  fp32-class modes (parity-safe, <= 5e-6 depth rel-L1):
#   "auto"  (default)  per layer the fastest of the back ends below, measured on first use (like cudnn.benchmark)
#   "fp32"             CUDA-core kernel everywhere (packed fp32x2 FMAs)
#   "ws2_tf32x3"       TMA-fed width-stacked tcgen05/TMEM kernel (conv_ws2.cu) wherever it applies, 3xTF32 operand split
#   "ws2_f16c"         as "ws2_tf32x3" with the two correction products from one fp16 MMA (2 MMAs per kernel row instead
#                      of 3, the same 11 significant bits per factor)
#   "ws_tf32x3"        the first-generation width-stacked tcgen05 kernel (conv_ws.cu; still used for upsampled inputs)
# Plain-TF32 mode (NOT parity-safe on the synthetic weights: 1e-3..3e-3 depth rel-L1, 2-6 % index flips):
#   "ws_tf32"          operands rounded to TF32 - the numerics cuDNN uses under torch defaults
# "tf32x3", "tf32", "tc_tf32x3", "tc_tf32": the round-1 mma.sync / tap-offset tcgen05 back ends (csrc/legacy/), only in a
# library built with DMVS_BUILD_LEGACY=1 (`legacy_backends()`); no shipped configuration uses them.
PRECISIONS = {"fp32": PREC_FP32, "tf32x3": PREC_TF32X3, "tf32": PREC_TF32, "tc_tf32x3": PREC_TC_TF32X3,
              "tc_tf32": PREC_TC_TF32, "auto": PREC_AUTO, "ws_tf32x3": PREC_WS_TF32X3, "ws_tf32": PREC_WS_TF32,
              "ws2_tf32x3": PREC_WS2_TF32X3, "ws2_f16c": PREC_WS2_TF32_F16C}
_precision = PRECISIONS[os.environ.get("DMVS_PRECISION", "auto")]


LEGACY_MODES = ("tf32x3", "tf32", "tc_tf32x3", "tc_tf32")


def legacy_backends() -> bool:
    """True when the loaded library was built with the round-1 back ends (DMVS_BUILD_LEGACY=1)."""
    return b"legacy back ends" in _cabi.lib().dmvs_build_info()


def set_precision(name: str) -> None:
    global _precision
    if name in LEGACY_MODES and not legacy_backends():
        raise ValueError(f"precision mode {name!r} needs a library built with DMVS_BUILD_LEGACY=1 (csrc/legacy/)")
    _precision = PRECISIONS[name]


# ------------------------------------------------------------------------------------------------
# Per-layer back-end autotuning ("auto" precision only).  The reference runs with `cudnn.benchmark = True`
# (test.py:18): cuDNN times its algorithms the first time it sees a layer shape and keeps the fastest.  Same
# here: the first call of a new convolution signature times every fp32-class back end that can run it (FFMA2,
# tcgen05 tap-offset, tcgen05 width-stacked) on the live buffers and caches the winner for the process.  All
# candidates meet the same parity class, so the choice only affects speed.  DMVS_AUTOTUNE=0 falls back to the
# static rule inside dmvs_conv_f32.
# ------------------------------------------------------------------------------------------------
_AUTOTUNE = os.environ.get("DMVS_AUTOTUNE", "1") != "0"
_AUTOTUNE_WS = os.environ.get("DMVS_AUTO_WS", "1") != "0"
_TUNED: dict = {}
_AUTOTUNE_F16C = os.environ.get("DMVS_AUTO_F16C", "1") != "0"     # let `auto` consider the fp16-correction variant
_BACKEND_BITS = ((1, PREC_FP32), (4, PREC_TC_TF32X3), (8, PREC_WS_TF32X3), (16, PREC_WS2_TF32X3), (32, PREC_WS2_TF32_F16C))


def set_autotune(enabled: bool, use_ws: Optional[bool] = None) -> None:
    global _AUTOTUNE, _AUTOTUNE_WS
    _AUTOTUNE = bool(enabled)
    if use_ws is not None:
        _AUTOTUNE_WS = bool(use_ws)
    _TUNED.clear()


# The table is keyed by (device index, signature); entries loaded from a file apply to every device (index -1).  A
# signature is the 24-tuple built in `conv`: shapes, kernel, stride, padding, prologue / epilogue kinds, pixel strides,
# activation and bias presence.
_SIG_LEN = 24
DEFAULT_TUNED_TABLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tuned", "b200_default.json")
_default_table_state = {"checked": False}


def _tuned_lookup(dev: int, sig: tuple):
    hit = _TUNED.get((dev,) + sig)
    return hit if hit is not None else _TUNED.get((-1,) + sig)


def save_tuned(path: str) -> None:
    """Write the autotuning table as JSON (signature -> chosen back end and the measured times), device independent."""
    import json
    names = {v: k for k, v in PRECISIONS.items()}
    rows, seen = [], set()
    for k, (c, ts) in _TUNED.items():
        sig = k[1:]
        if sig in seen:
            continue
        seen.add(sig)
        rows.append({"sig": [int(x) for x in sig], "choice": names[c], "ms": {names[m]: t for m, t in ts.items()}})
    with open(path, "w") as f:
        json.dump(rows, f, indent=0)


def load_tuned(path: str) -> int:
    """Preload an autotuning table written by `save_tuned`: its signatures use the recorded back end on every device
    instead of being timed (reproducible back-end choice across processes and ranks; also what a run under a profiler,
    whose timings are distorted, needs).  Entries of another signature format are skipped.  Returns the number loaded."""
    import json
    with open(path) as f:
        rows = json.load(f)
    n = 0
    for r in rows:
        sig = r["sig"]
        if len(sig) != _SIG_LEN or r["choice"] not in PRECISIONS:
            continue
        key = (-1,) + tuple(sig[:15]) + (bool(sig[15]),) + tuple(sig[16:22]) + (int(sig[22]), bool(sig[23]))
        _TUNED[key] = (PRECISIONS[r["choice"]], {PRECISIONS[m]: t for m, t in r.get("ms", {}).items() if m in PRECISIONS})
        n += 1
    return n


def _load_default_table_once(dev: int) -> None:
    """First tuned convolution of the process: preload the table shipped for this GPU model (B200) so that the layers of
    the known workloads run on recorded back ends - no timing noise decides them, every process and rank agrees, and
    results are bit-reproducible across runs.  DMVS_TUNED_TABLE=<path> selects another table, DMVS_TUNED_TABLE=0 none."""
    if _default_table_state["checked"]:
        return
    _default_table_state["checked"] = True
    path = os.environ.get("DMVS_TUNED_TABLE", DEFAULT_TUNED_TABLE)
    if path in ("0", "", "none") or not os.path.exists(path):
        return
    if path == DEFAULT_TUNED_TABLE and "B200" not in torch.cuda.get_device_name(dev):
        return
    load_tuned(path)


def tuned_table() -> dict:
    """signature -> (chosen precision code, {code: ms}) of every convolution tuned so far."""
    return dict(_TUNED)


_capture_misses: set = set()


def _span(t: Optional[Tensor]):
    """[first byte, last byte + 1) of the memory a strided view can touch."""
    if t is None or t.numel() == 0:
        return None
    last = sum((n - 1) * st for n, st in zip(t.shape, t.stride()))
    return t.data_ptr(), t.data_ptr() + (last + 1) * t.element_size()


def _disjoint(out: Tensor, *others: Optional[Tensor]) -> bool:
    """True when `out` shares no byte with any of `others`.  Spans are compared, except for the layout this package uses
    all the time: two channel slices of one channels-last buffer (same strides, offsets within one pixel record)."""
    a = _span(out)
    for o in others:
        b = _span(o)
        if a is None or b is None or not (a[0] < b[1] and b[0] < a[1]):
            continue
        if (out.dim() == o.dim() and out.stride() == o.stride() and tuple(out.shape[:-1]) == tuple(o.shape[:-1])
                and out.stride(-1) == 1 and out.dim() >= 2):
            delta = (o.data_ptr() - out.data_ptr()) // out.element_size()       # channel offset of o relative to out
            ps = out.stride(-2)
            if abs(delta) < ps and (delta >= out.shape[-1] or -delta >= o.shape[-1]):
                continue
        return False
    return True


def _tune(d: "ConvDesc", key, safe: bool = True) -> int:
    lib = _cabi.lib()
    mask = lib.dmvs_conv_backends(C.byref(d))
    cands = [code for bit, code in _BACKEND_BITS if mask & bit and (code != PREC_WS_TF32X3 or _AUTOTUNE_WS)
             and (code != PREC_WS2_TF32_F16C or _AUTOTUNE_F16C)]
    if len(cands) < 2:
        return cands[0]
    if torch.cuda.is_current_stream_capturing():
        if key not in _capture_misses:
            _capture_misses.add(key)
            import warnings
            warnings.warn("diffmvs_b200: a convolution signature was first seen during CUDA-graph capture; it runs on the "
                          "static back-end rule and is not tuned (run one eager forward first)", RuntimeWarning)
        return PREC_AUTO
    if not safe:      # the output overlaps an input: trial launches would change what later launches read
        return PREC_AUTO
    stream = _stream()
    stats, d.out_stats = d.out_stats, None        # accumulating statistics must not see the trial runs
    times = {}
    try:
        for code in cands:
            d.precision = code
            if lib.dmvs_conv_f32(C.byref(d), stream) != 0:
                continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = float("inf")
            for _ in range(3):
                e0.record()
                lib.dmvs_conv_f32(C.byref(d), stream)
                e1.record()
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            times[code] = best
    finally:
        d.out_stats = stats
    choice = min(times, key=times.get) if times else PREC_FP32
    _TUNED[key] = (choice, times)
    return choice



def get_precision() -> str:
    return {v: k for k, v in PRECISIONS.items()}[_precision]

Tensor = torch.Tensor

_replayed_launches = 0


def count_replayed_launches(n: int) -> None:
    """Kernels executed through CUDA-graph replays (the library's own counter only sees direct launches)."""
    global _replayed_launches
    _replayed_launches += n


def launch_count() -> int:
    """Kernels of libdiffmvs_b200.so run so far in this process: direct launches + graph-replayed ones."""
    return int(_cabi.lib().dmvs_launch_count()) + _replayed_launches


class Profiler:
    """Optional per-call CUDA-event timing of the kernel entry points (used by bench.py / tools; off by default).

    Each record is (name, tag, algorithmic_bytes, start_event, end_event); `summary()` synchronises and
    aggregates by (name, tag).  Bytes are the logical sizes of every tensor argument and result - the
    unique traffic an ideally fused launch would move."""

    def __init__(self):
        self.records = []

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, tag, nbytes, e0, e1 in self.records:
            k = (name, tag)
            a = agg.setdefault(k, [0, 0.0, 0])
            a[0] += 1
            a[1] += e0.elapsed_time(e1)
            a[2] += nbytes
        return {k: {"calls": v[0], "ms": v[1], "bytes": v[2]} for k, v in agg.items()}


_PROFILER: Optional[Profiler] = None
_LAST_CONV_BACKEND = PREC_FP32
# C kernel behind each back-end code (what `cuobjdump` / ncu list for that launch)
KERNEL_OF_BACKEND = {PREC_FP32: "conv_kernel", PREC_TF32X3: "conv_mma_kernel", PREC_TF32: "conv_mma_kernel",
                     PREC_TC_TF32X3: "conv_tc_kernel", PREC_TC_TF32: "conv_tc_kernel", PREC_WS_TF32X3: "conv_ws_kernel",
                     PREC_WS_TF32: "conv_ws_kernel", PREC_WS2_TF32X3: "conv_ws2_kernel",
                     PREC_WS2_TF32_F16C: "conv_ws2_kernel", PREC_AUTO: "conv_kernel|conv_tc_kernel"}


def set_profiler(p: Optional[Profiler]) -> None:
    global _PROFILER
    _PROFILER = p


def _tensor_bytes(obj) -> int:
    if isinstance(obj, torch.Tensor):
        return obj.numel() * obj.element_size()
    if isinstance(obj, (tuple, list)):
        return sum(_tensor_bytes(o) for o in obj)
    if isinstance(obj, PackedConv):
        return _tensor_bytes(obj.w)
    if isinstance(obj, GroupNormIn):
        return 0
    return 0


def _first_cuda_device(a, k) -> Optional[int]:
    for t in list(a) + list(k.values()):
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device.index
    return None


def _profiled(name: str):
    """Decorator of every kernel entry point: runs the call with the tensors' device current (launches, streams and
    the per-device shared-memory opt-in all follow the CUDA current device, so a model on cuda:1 must not launch
    on cuda:0's stream), and records CUDA-event timings when a `Profiler` is installed."""
    def deco(fn):
        def wrapper(*a, **k):
            dev = _first_cuda_device(a, k)
            if dev is not None and dev != torch.cuda.current_device():
                with torch.cuda.device(dev):
                    return wrapper(*a, **k)
            prof = _PROFILER
            if prof is None:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            nbytes = _tensor_bytes(a) + _tensor_bytes(list(k.values()))
            if k.get("out") is None and k.get("cost_out") is None:
                nbytes += _tensor_bytes(out)
            tag = ""
            if name == "conv":
                pc = a[1]
                tag = (f"{KERNEL_OF_BACKEND.get(_LAST_CONV_BACKEND, '?')}|{pc.cin}->{pc.cout} k{'x'.join(map(str, pc.k))} "
                       f"s{k.get('stride', 1)} {tuple(out.shape[1:-1])}")
            prof.records.append((name, tag, nbytes, e0, e1))
            return out
        wrapper.__name__, wrapper.__doc__ = fn.__name__, fn.__doc__
        return wrapper
    return deco


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req_cuda_f32(t: Tensor, name: str) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name}: the diffmvs_b200 kernels run on CUDA only (got {t.device}); there is no CPU path")
    if t.dtype != torch.float32:
        raise ValueError(f"{name}: expected float32, got {t.dtype}")


def pixel_stride(t: Tensor, name: str = "tensor") -> int:
    """Pixel stride of a channels-last map/volume view; validates the layout."""
    _req_cuda_f32(t, name)
    if t.dim() not in (4, 5) or t.stride(-1) != 1:
        raise ValueError(f"{name}: expected a channels-last [N,(D,)H,W,C] view, got shape {tuple(t.shape)} "
                         f"strides {t.stride()}")
    ps = t.stride(-2)
    expect = ps
    for dim in range(t.dim() - 2, 0, -1):  # W, H, (D)
        if t.shape[dim] > 1 and t.stride(dim) != expect:
            raise ValueError(f"{name}: not a dense channels-last view (strides {t.stride()})")
        expect *= t.shape[dim]
    if t.shape[0] > 1 and t.stride(0) != expect:
        raise ValueError(f"{name}: batch stride {t.stride(0)} != {expect}")
    if ps < t.shape[-1]:
        raise ValueError(f"{name}: pixel stride {ps} < channels {t.shape[-1]}")
    return ps


@dataclass
class PackedConv:
    """Host-prepared convolution: weights [KD,KH,KW,cin_pad,cout_pad] (BN folded), optional bias."""
    w: Tensor
    bias: Optional[Tensor]
    cin: int
    cout: int
    k: Tuple[int, int, int]
    w_t: Optional[Tensor] = None   # tensor-core layout [KD,KH,KW,cout_pad8,cin_pad8]
    w_tc: Optional[Tensor] = None  # tcgen05 layout [2(hi,lo),KD,KH*KW,cin_pad8/4,cout_pad16,4]
    w_ws: Optional[Tensor] = None  # width-stacked tcgen05 slabs for stride 1 (packing.pack_ws)
    ws_strided: Optional[dict] = None   # (stride, pad_h, pad_w) -> slabs for strided use, built on first use
    w_ws_pair: Optional[Tensor] = None  # <= 4 input channels: slabs with kernel rows paired along K (packing.pack_ws_pair)
    w_ws16: Optional[Tensor] = None     # `w_ws` with fp16 correction planes (packing.pack_ws(..., corr16=True)), stride 1
    w_host: Optional[Tensor] = None     # 8 -> 1 3x3x3 layers: [kd][kh][kw][ci] on the HOST (ops.conv3d_to1 launch parameters)
    bias_host: float = 0.0

    def ws_slabs(self, stride: int, pad_h: int, pad_w: int, corr16: bool = False) -> Optional[Tensor]:
        if stride == 1 or self.w_ws is None:
            return self.w_ws16 if corr16 else self.w_ws
        if self.ws_strided is None:
            self.ws_strided = {}
        key = (stride, pad_h, pad_w, corr16)
        if key not in self.ws_strided:
            from . import packing
            self.ws_strided[key] = packing.pack_ws_from_packed(self.w, self.cout, stride, (pad_h, pad_w),
                                                               corr16=corr16).to(self.w.device)
        return self.ws_strided[key]

    def to(self, device) -> "PackedConv":
        mv = lambda t: None if t is None else t.to(device)
        return PackedConv(self.w.to(device), mv(self.bias), self.cin, self.cout, self.k, mv(self.w_t), mv(self.w_tc),
                          mv(self.w_ws), None, mv(self.w_ws_pair), mv(self.w_ws16), self.w_host, self.bias_host)


def _row_strided(t: Tensor, name: str) -> Tuple[int, int]:
    """(pixel stride, row stride) of a channels-last [N,H,W,C] view whose rows may be strided (phase launches)."""
    _req_cuda_f32(t, name)
    if t.dim() != 4 or t.stride(-1) != 1 or (t.shape[0] > 1 and t.stride(0) != t.shape[1] * t.stride(1)):
        raise ValueError(f"{name}: expected a [N,H,W,C] view with uniformly strided rows, got strides {t.stride()}")
    if t.stride(2) < t.shape[3] or t.stride(1) < t.shape[2] * t.stride(2):
        raise ValueError(f"{name}: overlapping pixels / rows (strides {t.stride()})")
    return t.stride(2), t.stride(1)


@dataclass
class GroupNormIn:
    """GroupNorm(4)+affine+SiLU of the producer, applied while the consumer stages its input."""
    stats: Tensor   # [N,4,2] int64 fixed-point accumulators (2^-20 units) filled through `conv(..., out_stats=)`
    g1: Tensor      # [C]
    g0: Tensor      # [C]


@_profiled("conv")
def conv(x: Tensor, pc: PackedConv, *, x2: Optional[Tensor] = None, stride: int = 1,
         pad: Optional[Tuple[int, int, int]] = None, act: int = ACT_NONE, act_c0: int = 0,
         res: Optional[Tensor] = None, res_mode: int = RES_NONE, res_up2: bool = False, in_up2: bool = False,
         in_gn: Optional[GroupNormIn] = None, out: Optional[Tensor] = None, out_stats: Optional[Tensor] = None,
         epi: int = EPI_STD, aux1: Optional[Tensor] = None, aux2: Optional[Tensor] = None,
         gru_hidden: int = 0, explicit_out: Optional[Tuple[int, int]] = None) -> Tensor:
    """2-D (x: [N,H,W,C]) or 3-D (x: [N,D,H,W,C]) convolution with fused prologue/epilogue.
    `explicit_out` = (Ho, Wo) marks a phase launch (see `conv_up2`): one-sided padding, `out` / `res` may be views whose
    rows are strided (every other row of a larger tensor); TMA-fed tcgen05 back end only."""
    three_d = x.dim() == 5
    x_ps = pixel_stride(x, "conv input")
    if three_d:
        N, D, H, W, C1 = x.shape
    else:
        N, H, W, C1 = x.shape
        D = 1
    if in_up2:
        H, W = 2 * H, 2 * W
    C2 = 0
    x2_ps = 0
    if x2 is not None:
        x2_ps = pixel_stride(x2, "conv input 2")
        if tuple(x2.shape[:-1]) != tuple(x.shape[:-1]):
            raise ValueError(f"conv: concat inputs disagree: {tuple(x.shape)} vs {tuple(x2.shape)}")
        C2 = x2.shape[-1]
    if C1 + C2 != pc.cin:
        raise ValueError(f"conv: input has {C1}+{C2} channels, weights expect {pc.cin}")
    KD, KH, KW = pc.k
    if pad is None:
        pad = (KD // 2, KH // 2, KW // 2)
    pd, ph, pw = pad
    Do = (D + 2 * pd - KD) // stride + 1
    Ho = (H + 2 * ph - KH) // stride + 1
    Wo = (W + 2 * pw - KW) // stride + 1
    if explicit_out is not None:
        if three_d or out is None or stride != 1 or res_up2 or in_up2:
            raise ValueError("conv: a phase launch is a 2-D stride-1 convolution into a given output view")
        Ho, Wo = explicit_out
    oshape = (N, Do, Ho, Wo, pc.cout) if three_d else (N, Ho, Wo, pc.cout)
    if out is None:
        out = torch.empty(oshape, device=x.device, dtype=torch.float32)
    elif tuple(out.shape) != oshape:
        raise ValueError(f"conv: out has shape {tuple(out.shape)}, expected {oshape}")
    y_rs = res_rs = 0
    if explicit_out is not None:
        y_ps, y_rs = _row_strided(out, "conv output")
    else:
        y_ps = pixel_stride(out, "conv output")

    d = ConvDesc()
    d.x, d.x2 = _ptr(x), _ptr(x2)
    d.N, d.D, d.H, d.W = N, D, H, W
    d.C1, d.C2, d.x_ps, d.x2_ps, d.in_up2 = C1, C2, x_ps, x2_ps, int(in_up2)
    if in_gn is not None:
        d.in_stats, d.in_g1, d.in_g0 = _ptr(in_gn.stats), _ptr(in_gn.g1), _ptr(in_gn.g0)
        d.in_inv_count = 1.0 / float(D * H * W * (C1 // 4))
    d.w, d.bias = _ptr(pc.w), _ptr(pc.bias)
    d.w_t, d.w_tc, d.w_ws = _ptr(pc.w_t), _ptr(pc.w_tc), _ptr(pc.ws_slabs(stride, ph, pw))
    d.w_ws_pair = _ptr(pc.w_ws_pair) if stride == 1 else None
    d.w_ws16 = _ptr(pc.ws_slabs(stride, ph, pw, corr16=True)) if pc.w_ws16 is not None else None
    d.precision = _precision if pc.w_t is not None else PREC_FP32
    if d.precision in (PREC_TC_TF32X3, PREC_TC_TF32, PREC_WS_TF32X3, PREC_WS_TF32, PREC_WS2_TF32X3,
                       PREC_WS2_TF32_F16C) and pc.w_tc is None:
        d.precision = PREC_TF32X3 if d.precision in (PREC_TC_TF32X3, PREC_WS_TF32X3, PREC_WS2_TF32X3,
                                                     PREC_WS2_TF32_F16C) else PREC_TF32
    d.KD, d.KH, d.KW, d.stride = KD, KH, KW, stride
    d.pad_d, d.pad_h, d.pad_w = pd, ph, pw
    d.y, d.Do, d.Ho, d.Wo, d.Cout, d.y_ps = _ptr(out), Do, Ho, Wo, pc.cout, y_ps
    d.act, d.act_c0, d.res_mode = act, act_c0, res_mode
    if res is not None and explicit_out is not None:
        d.res_ps, res_rs = _row_strided(res, "conv residual")
        d.res = _ptr(res)
        if tuple(res.shape) != tuple(out.shape):
            raise ValueError("conv: phase-launch residual must have the output view's shape")
    elif res is not None:
        d.res, d.res_ps, d.res_up2 = _ptr(res), pixel_stride(res, "conv residual"), int(res_up2)
        want = (N, Ho // 2, Wo // 2) if res_up2 else tuple(oshape[:-1])
        if tuple(res.shape[:-1]) != want or res.shape[-1] < pc.cout:
            raise ValueError(f"conv: residual shape {tuple(res.shape)} incompatible with output {oshape}")
    d.epi, d.gru_hidden = epi, gru_hidden
    if explicit_out is not None:
        d.explicit_extent, d.y_row_stride, d.res_row_stride = 1, y_rs, res_rs
        if d.precision not in (PREC_AUTO, PREC_WS2_TF32X3, PREC_WS2_TF32_F16C):
            d.precision = PREC_WS2_TF32X3        # the only back end with strided rows; same fp32-class arithmetic
    if aux1 is not None:
        d.aux1, d.aux1_ps = _ptr(aux1), pixel_stride(aux1, "conv aux1")
    if aux2 is not None:
        d.aux2, d.aux2_ps = _ptr(aux2), pixel_stride(aux2, "conv aux2")
    if out_stats is not None and out_stats.dtype != torch.int64:
        raise ValueError("conv: out_stats must be a zero-initialised int64 [N,4,2] tensor (fixed-point accumulators)")
    d.out_stats = _ptr(out_stats)
    if d.precision == PREC_AUTO and explicit_out is not None:
        d.precision = PREC_WS2_TF32_F16C if (_AUTOTUNE_F16C and pc.w_ws16 is not None) else PREC_WS2_TF32X3
    if d.precision == PREC_AUTO and _AUTOTUNE:
        dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
        _load_default_table_once(dev)
        sig = (N, D, H, W, C1, C2, pc.cout, KD, KH, KW, stride, pd, ph, pw, int(in_up2), in_gn is not None, epi,
               res_mode, int(res_up2), x_ps, x2_ps, y_ps, int(act), pc.bias is not None)
        hit = _tuned_lookup(dev, sig)
        d.precision = hit[0] if hit is not None else _tune(d, (dev,) + sig, _disjoint(out, x, x2, res, aux1, aux2))
    global _LAST_CONV_BACKEND
    _LAST_CONV_BACKEND = PREC_WS_TF32X3 if (d.precision in (PREC_WS2_TF32X3, PREC_WS2_TF32_F16C) and in_up2) else d.precision
    check(_cabi.lib().dmvs_conv_f32(C.byref(d), _stream()), "dmvs_conv_f32")
    return out


@_profiled("deconv3d")
def deconv3d(x: Tensor, w: Tensor, bias: Tensor, skip: Tensor) -> Tensor:
    N, D, H, W, Cin = x.shape
    Cout = skip.shape[-1]
    for t, n in ((x, "x"), (skip, "skip")):
        _req_cuda_f32(t, n)
        if not t.is_contiguous():
            raise ValueError("deconv3d: dense channels-last volumes required")
    y = torch.empty_like(skip)
    check(_cabi.lib().dmvs_deconv3d_f32(_ptr(x), _ptr(w), _ptr(bias), _ptr(skip), _ptr(y), N, D, H, W, Cin, Cout,
                                        _stream()), "dmvs_deconv3d_f32")
    return y


def conv_up2(x: Tensor, pcu: Tuple[PackedConv, PackedConv], *, out: Optional[Tensor] = None,
             accumulate: bool = False) -> Tensor:
    """conv3x3(pad 1)(nearest_x2(x)) for x [N,H,W,C] -> [N,2H,2W,Cout] WITHOUT forming the upsampled map: output row
    parity py and column parity px select which of the 3x3 taps fall on the same low-resolution pixel, so each parity
    class is a 2x2 convolution of x with summed taps (packing.pack_up2_phases).  Two launches (py = 0, 1), each a
    KH=2 x KW=3 convolution producing both column parities as 2*Cout channels = two adjacent output pixels: 4/9 of the
    multiply-adds of the direct form and a quarter of its input traffic.  `accumulate` adds to `out` in place."""
    N, H, W, C = x.shape
    cout = pcu[0].cout // 2
    if out is None:
        if accumulate:
            raise ValueError("conv_up2: accumulate needs `out`")
        out = torch.empty((N, 2 * H, 2 * W, cout), device=x.device, dtype=torch.float32)
    if tuple(out.shape) != (N, 2 * H, 2 * W, cout) or not out.is_contiguous():
        raise ValueError(f"conv_up2: out must be a dense [N,2H,2W,{cout}] tensor")
    rows = out.view(N, H, 2, W, 2 * cout)
    for py in (0, 1):
        view = rows[:, :, py]                                  # [N,H,W,2*cout]: every other output row, pixel pairs
        conv(x, pcu[py], pad=(0, 1 - py, 1), out=view, explicit_out=(H, W),
             res=view if accumulate else None, res_mode=RES_PRE_ACT if accumulate else RES_NONE)
    return out


@_profiled("border_bias_add")
def border_bias_add(y: Tensor, table: Tensor) -> Tensor:
    """y [N,H,W,C] += table[3,3,C] on the one-pixel frame of every image (in place)."""
    ps = pixel_stride(y, "border_bias_add")
    N, H, W, Cc = y.shape
    if tuple(table.shape) != (3, 3, Cc) or not table.is_contiguous():
        raise ValueError(f"border_bias_add: table must be a dense [3,3,{Cc}] tensor")
    _req_cuda_f32(table, "table")
    check(_cabi.lib().dmvs_border_bias_add(_ptr(y), ps, _ptr(table), N, H, W, Cc, _stream()), "dmvs_border_bias_add")
    return y


@_profiled("conv3d_to1")
def conv3d_to1(x: Tensor, pc: "PackedConv", sigmoid_max: bool = False) -> Tensor:
    """Conv3d(8 -> 1, k=3, p=1) of a channels-last volume x [N,D,H,W,8] (a channel slice of a wider buffer is fine).
    Returns the logits [N,D,H,W] or, with `sigmoid_max`, max_d sigmoid(logit) [N,H,W] (PixelViewWeight.forward,
    module.py:459-463).  fp32 FFMA arithmetic whatever the precision mode."""
    _req_cuda_f32(x, "x")
    if x.dim() != 5 or pc.cin != 8 or pc.cout != 1 or tuple(pc.k) != (3, 3, 3) or pc.w_host is None:
        raise ValueError("conv3d_to1: needs [N,D,H,W,8] input and a packed 8 -> 1 3x3x3 layer")
    N, D, H, W, C = x.shape
    ps = pixel_stride(x, "conv3d_to1 input")
    y = torch.empty((N, H, W) if sigmoid_max else (N, D, H, W), device=x.device, dtype=torch.float32)
    check(_cabi.lib().dmvs_conv3d_to1_f32(_ptr(x), ps, pc.w_host.data_ptr(), float(pc.bias_host), _ptr(y), N, D, H, W,
                                          1 if sigmoid_max else 0, _stream()), "dmvs_conv3d_to1_f32")
    return y


@_profiled("compose_homographies")
def compose_homographies(proj: Tensor) -> Tensor:
    """proj [B,V,2,4,4] -> [B,V-1,12]."""
    _req_cuda_f32(proj, "proj_matrices")
    proj = proj.contiguous()
    B, V = proj.shape[:2]
    hom = torch.empty((B, V - 1, 12), device=proj.device, dtype=torch.float32)
    check(_cabi.lib().dmvs_compose_homographies(_ptr(proj), _ptr(hom), B, V, _stream()), "dmvs_compose_homographies")
    return hom


@_profiled("warp_volume")
def warp_volume(src: Tensor, hom: Tensor, depth: Tensor) -> Tensor:
    """src [B,Hs,Ws,C], hom [B,12], depth [B,D,H,W] -> [B,D,H,W,C]."""
    ps = pixel_stride(src, "warp source")
    B, Hs, Ws, Cc = src.shape
    _, D, H, W = depth.shape
    out = torch.empty((B, D, H, W, Cc), device=src.device, dtype=torch.float32)
    check(_cabi.lib().dmvs_warp_volume(_ptr(src), ps, _ptr(hom.contiguous()), _ptr(depth.contiguous()), _ptr(out), B, Cc,
                                       Hs, Ws, D, H, W, _stream()), "dmvs_warp_volume")
    return out


@_profiled("plane_sweep_corr")
def plane_sweep_corr(feats: Tensor, hom: Tensor, plane_depth: Tensor, G: int) -> Tensor:
    """feats [V,B,H,W,C] dense, hom [B,V-1,12], plane_depth [B,D] -> cor [B*(V-1),D,H,W,G]."""
    _req_cuda_f32(feats, "features")
    if not feats.is_contiguous():
        raise ValueError("plane_sweep_corr: dense [V,B,H,W,C] features required")
    V, B, H, W, Cc = feats.shape
    D = plane_depth.shape[1]
    cor = torch.empty((B * (V - 1), D, H, W, G), device=feats.device, dtype=torch.float32)
    check(_cabi.lib().dmvs_plane_sweep_corr(_ptr(feats), _ptr(hom), _ptr(plane_depth.contiguous()), _ptr(cor), B, V, Cc, G,
                                            D, H, W, _stream()), "dmvs_plane_sweep_corr")
    return cor


@_profiled("view_weight_max")
def view_weight_max(logit: Tensor) -> Tensor:
    """logit [N,D,H,W] (dense) -> [N,H,W]."""
    N, D, H, W = logit.shape
    w = torch.empty((N, H, W), device=logit.device, dtype=torch.float32)
    check(_cabi.lib().dmvs_view_weight_max(_ptr(logit), _ptr(w), N, D, H * W, _stream()), "dmvs_view_weight_max")
    return w


@_profiled("aggregate_views")
def aggregate_views(cor: Tensor, w: Tensor, B: int) -> Tensor:
    """cor [B*V1,D,H,W,G], w [B*V1,H,W] -> [B,D,H,W,G]."""
    NV, D, H, W, G = cor.shape
    V1 = NV // B
    vol = torch.empty((B, D, H, W, G), device=cor.device, dtype=torch.float32)
    check(_cabi.lib().dmvs_aggregate_views(_ptr(cor), _ptr(w), _ptr(vol), B, V1, D, H * W, G, _stream()),
          "dmvs_aggregate_views")
    return vol


@_profiled("depth_regression")
def depth_regression(logits: Tensor, depth_min: Tensor, depth_max: Tensor, want_floor: bool = False):
    """logits [B,D,H,W] dense -> (norm_inv, depth, conf [B,H,W], floor_idx or None)."""
    B, D, H, W = logits.shape
    mk = lambda: torch.empty((B, H, W), device=logits.device, dtype=torch.float32)
    n, dep, conf = mk(), mk(), mk()
    fl = torch.empty((B, H, W), device=logits.device, dtype=torch.int32) if want_floor else None
    check(_cabi.lib().dmvs_depth_regression(_ptr(logits), _ptr(depth_min), _ptr(depth_max), _ptr(n), _ptr(dep), _ptr(conf),
                                            _ptr(fl), B, D, H * W, _stream()), "dmvs_depth_regression")
    return n, dep, conf, fl


@_profiled("get_cost")
def get_cost(feats: Tensor, hom: Tensor, inv_depth: Tensor, conf: Optional[Tensor], view_w: Tensor, depth_min: Tensor,
             depth_max: Tensor, G: int, D: int, wshift: int, interval: float, min_radius: float, max_radius: float,
             cost_out: Optional[Tensor] = None, samples_out: Optional[Tensor] = None):
    """feats [V,B,H,W,C]; inv_depth/conf [B,H,W]; view_w [B,V-1,H>>s,W>>s] -> cost [B,H,W,G*D], samples [B,H,W,D]."""
    if not feats.is_contiguous():
        raise ValueError("get_cost: dense [V,B,H,W,C] features required")
    V, B, H, W, Cc = feats.shape
    if cost_out is None:
        cost_out = torch.empty((B, H, W, G * D), device=feats.device, dtype=torch.float32)
    if samples_out is None:
        samples_out = torch.empty((B, H, W, D), device=feats.device, dtype=torch.float32)
    if tuple(view_w.shape) != (B, V - 1, H >> wshift, W >> wshift) or not view_w.is_contiguous():
        raise ValueError(f"get_cost: view weights {tuple(view_w.shape)} do not match {(B, V - 1, H >> wshift, W >> wshift)}")
    conf_ps = 1
    if conf is not None:
        if conf.dim() == 4:      # [B,H,W,k] channel slice of the U-Net head
            conf_ps = pixel_stride(conf, "conf")
        elif not conf.is_contiguous():
            raise ValueError("get_cost: conf must be dense [B,H,W] or a channel slice [B,H,W,1]")
    check(_cabi.lib().dmvs_get_cost(_ptr(feats), _ptr(hom), _ptr(inv_depth), _ptr(conf), conf_ps, _ptr(view_w), _ptr(depth_min),
                                    _ptr(depth_max), _ptr(cost_out), pixel_stride(cost_out, "cost"), _ptr(samples_out),
                                    pixel_stride(samples_out, "samples"), B, V, Cc, G, D, H, W, wshift, interval,
                                    min_radius, max_radius, _stream()), "dmvs_get_cost")
    return cost_out, samples_out


@_profiled("groupnorm_silu_add")
def groupnorm_silu_add(x: Tensor, gn: GroupNormIn, res: Optional[Tensor], out: Optional[Tensor] = None) -> Tensor:
    """x [N,H,W,C] dense raw conv output -> silu(GN(x)) + res."""
    N, H, W, Cc = x.shape
    if not x.is_contiguous():
        raise ValueError("groupnorm_silu_add: dense input required")
    if out is None:
        out = torch.empty_like(x)
    check(_cabi.lib().dmvs_groupnorm_silu_add(_ptr(x), _ptr(gn.stats), _ptr(gn.g1), _ptr(gn.g0), _ptr(res),
                                              0 if res is None else pixel_stride(res, "residual"), _ptr(out),
                                              pixel_stride(out, "out"), N, H * W, Cc, _stream()),
          "dmvs_groupnorm_silu_add")
    return out


@_profiled("upsample_depth")
def upsample_depth(n: Tensor, mask: Tensor, depth_min: Optional[Tensor], depth_max: Optional[Tensor], ratio: int,
                   want: str = "depth+norm"):
    """n [B,H,W], mask [B,H,W,9r^2] -> tuple of the requested [B,rH,rW] maps ("raw", "depth", "norm")."""
    B, H, W = n.shape
    mk = lambda: torch.empty((B, H * ratio, W * ratio), device=n.device, dtype=torch.float32)
    keys = want.split("+")
    outs = {k: mk() for k in keys}
    check(_cabi.lib().dmvs_upsample_depth(_ptr(n.contiguous()), _ptr(mask), pixel_stride(mask, "mask"), _ptr(depth_min),
                                          _ptr(depth_max), _ptr(outs.get("raw")), _ptr(outs.get("depth")),
                                          _ptr(outs.get("norm")), B, H, W, ratio, _stream()), "dmvs_upsample_depth")
    return tuple(outs[k] for k in keys)


@_profiled("refine_update")
def refine_update(mode: int, inv0: Tensor, src: Optional[Tensor], src_ps: int, scale: float, delta: Tensor, inv: Tensor,
                  inv_slot: Optional[Tensor], slot_ps: int, depth_min: Optional[Tensor] = None,
                  depth_max: Optional[Tensor] = None, depth: Optional[Tensor] = None) -> None:
    B = inv0.shape[0]
    HW = inv0.numel() // B
    check(_cabi.lib().dmvs_refine_update(mode, _ptr(inv0), _ptr(src), src_ps, scale, _ptr(delta), _ptr(inv), _ptr(inv_slot),
                                         slot_ps, _ptr(depth_min), _ptr(depth_max), _ptr(depth), B, HW, _stream()),
          "dmvs_refine_update")


@_profiled("ddim_step")
def ddim_step(img: Tensor, delta: Tensor, noise: Tensor, k_recip: float, k_recipm1: float, sqrt_a_next: float, c: float,
              sigma: float, scale: float) -> None:
    check(_cabi.lib().dmvs_ddim_step(_ptr(img), _ptr(delta), _ptr(noise), k_recip, k_recipm1, sqrt_a_next, c, sigma, scale,
                                     img.numel(), _stream()), "dmvs_ddim_step")


@_profiled("upsample_nearest")
def upsample_nearest(x: Tensor, factor: int) -> Tensor:
    """x [B,H,W] dense or a [B,H,W,1] channel slice -> [B,fH,fW]."""
    x_ps = 1
    if x.dim() == 4:
        x_ps = pixel_stride(x, "upsample_nearest input")
    elif not x.is_contiguous():
        x = x.contiguous()
    B, H, W = x.shape[:3]
    y = torch.empty((B, H * factor, W * factor), device=x.device, dtype=torch.float32)
    check(_cabi.lib().dmvs_upsample_nearest(_ptr(x), x_ps, _ptr(y), B, H, W, factor, _stream()), "dmvs_upsample_nearest")
    return y


@_profiled("to_nhwc")
def to_nhwc(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """Logical NCHW tensor -> [N,H,W,C] (zero-copy when x is already channels_last)."""
    _req_cuda_f32(x, "input")
    N, Cc, H, W = x.shape
    v = x.permute(0, 2, 3, 1)
    if out is None and v.is_contiguous():
        return v
    if out is None:
        out = torch.empty((N, H, W, Cc), device=x.device, dtype=torch.float32)
    if v.is_contiguous():
        out.copy_(v)
        return out
    xc = x.contiguous()
    check(_cabi.lib().dmvs_nchw_to_nhwc(_ptr(xc), _ptr(out), pixel_stride(out, "out"), N, Cc, H * W, _stream()),
          "dmvs_nchw_to_nhwc")
    return out


@_profiled("to_nhwc")
def image_to_nhwc4(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """RGB image batch [N,3,H,W] -> [N,H,W,4] (fourth channel zero).  float32 images in [0,1] (what the reference's loader
    yields) or uint8 images as decoded from disk (extension: scaled by 1/255 on the device with the loader's exact fp32
    division, so results are bit-identical while the host->device copy is 4x smaller); a uint8 input may be a permuted
    view of an interleaved [N,H,W,3] array."""
    if x.dtype == torch.uint8:
        return _image_u8_to_nhwc4(x, out)
    _req_cuda_f32(x, "image")
    N, Cc, H, W = x.shape
    if Cc != 3:
        raise ValueError(f"image_to_nhwc4: expected 3 channels, got {Cc}")
    if out is None:
        out = torch.empty((N, H, W, 4), device=x.device, dtype=torch.float32)
    elif tuple(out.shape) != (N, H, W, 4) or not out.is_contiguous():
        raise ValueError("image_to_nhwc4: out must be a dense [N,H,W,4] tensor")
    check(_cabi.lib().dmvs_image_to_nhwc4(_ptr(x.contiguous()), _ptr(out), N, H * W, _stream()), "dmvs_image_to_nhwc4")
    return out


def _image_u8_to_nhwc4(x: Tensor, out: Optional[Tensor]) -> Tensor:
    if not x.is_cuda:
        raise ValueError("image: CUDA tensor required (there is no CPU path)")
    N, Cc, H, W = x.shape
    if Cc != 3:
        raise ValueError(f"image_to_nhwc4: expected 3 channels, got {Cc}")
    sn, sc, sh, sw = x.stride()
    if sh != W * sw or (N > 1 and sn < 3 * H * W) or not ((sc == H * W and sw == 1) or (sc == 1 and sw == 3)):
        x = x.contiguous()
        sn, sc, sh, sw = x.stride()
    if out is None:
        out = torch.empty((N, H, W, 4), device=x.device, dtype=torch.float32)
    elif tuple(out.shape) != (N, H, W, 4) or not out.is_contiguous():
        raise ValueError("image_to_nhwc4: out must be a dense [N,H,W,4] tensor")
    check(_cabi.lib().dmvs_image_u8_to_nhwc4(_ptr(x), sn, sc, sw, _ptr(out), N, H * W, _stream()), "dmvs_image_u8_to_nhwc4")
    return out


def to_nchw_view(y: Tensor) -> Tensor:
    """[N,H,W,C] -> logical NCHW view (channels_last memory format, no copy)."""
    return y.permute(0, 3, 1, 2)


@_profiled("to_nchw_dense")
def to_nchw_dense(y: Tensor) -> Tensor:
    """[N,H,W,C] (possibly a channel slice) -> contiguous NCHW copy."""
    N, H, W, Cc = y.shape
    out = torch.empty((N, Cc, H, W), device=y.device, dtype=torch.float32)
    check(_cabi.lib().dmvs_nhwc_to_nchw(_ptr(y), pixel_stride(y, "y"), _ptr(out), N, Cc, H * W, _stream()),
          "dmvs_nhwc_to_nchw")
    return out
