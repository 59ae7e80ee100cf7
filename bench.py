#!/usr/bin/env python
"""Benchmark of the DiffMVS / CasDiffMVS inference hot path (ref-views/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl ours|reference]

A *step* is one forward of `CasDiffMVS` for one reference view (the region the reference times in
`/root/reference/test.py:122-127`) on synthetic DTU-shaped inputs (`diffmvs_b200/synth.py`).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ref-views/sec at DTU 1600x1152xD384x7-view; depth L1 vs ref"
UNIT = "ref-views/s"
# algorithmic bytes per ref-view at cfg3 with ideal per-operator fusion (fp32)
ALGO_BYTES_CFG3 = 2.75e9


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            p = json.load(open(path))
            return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def _best_cpu_threads(cores: int) -> int:
    """torch's intra-op pools oversubscribe badly on big hosts (all 128 threads of the B200 box: 221 s/ref-view,
    16 threads: 7.5 s); pick the fastest of {32, 16, 8} threads on a small forward (cfg2) for the baseline."""
    import torch
    cands = sorted({c for c in (min(cores, 32), 16, 8) if c <= cores}, reverse=True)
    if len(cands) == 1:
        return cands[0]
    best, best_t = cands[0], float("inf")
    for c in cands:                       # ~1 s per forward at cfg2: the whole sweep stays under ~15 s
        run = _oracle_forward_cpu("cfg2", c)
        run()
        t = run()
        if t < best_t:
            best, best_t = c, t
    return best


def _reference_kind() -> str:
    """"reference" when the reference tree is mounted (build container: its own `CasDiffMVS` runs, unmodified),
    "port" otherwise (GPU box: the oracle restatement, bit-identical to it on CPU - tests/test_oracle_golden.py)."""
    from oracle import refimport
    return "reference" if refimport.reference_available() else "port"


def _oracle_forward_cpu(workload: str, threads: int):
    """One CPU forward of the reference algorithm on `threads` host threads; returns a callable giving seconds.
    Runs `/root/reference`'s own module when that tree exists, the oracle port otherwise."""
    import torch
    from diffmvs_b200 import synth
    from oracle import diffmvs_ref as O
    from oracle import refimport, spec
    torch.set_num_threads(threads)
    args = synth.workload_args(workload)
    sd = synth.synth_state_dict(spec.state_dict_shapes(args), 123)
    imgs, proj, dv = synth.workload_inputs(workload)
    gen = torch.Generator().manual_seed(1)
    randn = lambda like: torch.randn(like.shape, generator=gen, dtype=torch.float32)
    if refimport.reference_available():
        ref_models = refimport.import_reference_models()
        model = ref_models.CasDiffMVS(args, test=True).eval()
        full = dict(model.state_dict())
        full.update(sd)
        model.load_state_dict(full, strict=True)

        def run():
            t0 = time.perf_counter()
            with torch.no_grad():
                model(imgs, proj, dv)
            return time.perf_counter() - t0
        return run

    def run():
        t0 = time.perf_counter()
        with torch.no_grad():
            O.casdiffmvs_forward(sd, args, imgs, proj, dv, randn=randn)
        return time.perf_counter() - t0
    return run


def _gpu_baseline(workload: str, dev, steps: int = 5):
    """The reference algorithm on stock torch CUDA ops (cuDNN / ATen) on THIS GPU - what a user of the reference gets on
    a B200 ("the real bar to beat"): the oracle restatement with device tensors, cudnn.benchmark on as in
    test.py:18, once with torch's default TF32 convolutions and once in strict fp32."""
    import torch
    from diffmvs_b200 import synth
    from oracle import diffmvs_ref as O
    from oracle import spec
    args = synth.workload_args(workload)
    sd = {k: v.to(dev) for k, v in synth.synth_state_dict(spec.state_dict_shapes(args), 123).items()}
    imgs, proj, dv = synth.workload_inputs(workload)
    imgs, proj, dv = [i.to(dev) for i in imgs], {k: v.to(dev) for k, v in proj.items()}, dv.to(dev)
    out = {"kind": "oracle restatement on stock torch CUDA ops (cuDNN/ATen), cudnn.benchmark=True, same GPU", "unit": UNIT}
    old = (torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cudnn.benchmark = True
    try:
        for name, tf32 in (("tf32_default", True), ("fp32_strict", False)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            sampler = ClockSampler(dev.index or 0)
            with torch.no_grad():
                for _ in range(3):
                    O.casdiffmvs_forward(sd, args, imgs, proj, dv)
                torch.cuda.synchronize(dev)
                sampler.start()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    O.casdiffmvs_forward(sd, args, imgs, proj, dv)
                e1.record()
                torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": 1e3 / ms, "ms_per_step": ms, "steps": steps, "clocks": sampler.stop()}
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    del sd, imgs, proj, dv
    torch.cuda.empty_cache()
    return out


def run_reference(a):
    """`--impl reference`: the reference algorithm on the host cores - the reference's own module where its tree is
    mounted, the oracle port on the GPU box (the Python reference cannot travel there)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = _best_cpu_threads(os.cpu_count() or 1)
    run = _oracle_forward_cpu(a.workload, cores)
    first = run()                                   # warm-up (thread pools, allocator)
    budget = 200.0
    n_warm = max(0, min(a.warmup - 1, int(20.0 / max(first, 1e-3))))
    for _ in range(n_warm):
        run()
    n = max(1, min(a.steps, int(budget / max(first, 1e-3))))
    times = [run() for _ in range(n)]
    sec = sum(times) / len(times)
    value = 1.0 / sec
    sample = f"{n} timed forward(s) of 1 ref-view at {a.workload} (of --steps {a.steps}), {n_warm + 1} warm-up"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": _config(a.workload),
        "parallelism": "reference algorithm on the host cores (rank 0 only)",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": _reference_kind(), "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _config(workload):
    """The `config` object of the JSON line: names the workload only, identical for both arms (how the work is spread
    over GPUs is reported in the separate `parallelism` key)."""
    from diffmvs_b200 import synth
    variant, H, W, V, D0 = synth.WORKLOADS[workload]
    return {"workload": workload, "variant": variant, "image": [W, H], "views": V, "numdepth_initial": D0,
            "numdepth": 384, "batch": 1,
            "l2": (f"inputs ({V * 3 * H * W * 4 / 1e6:.0f} MB) and per-step working set exceed the 126 MB L2"
                   if V * 3 * H * W * 4 > 126e6 else
                   f"inputs ({V * 3 * H * W * 4 / 1e6:.1f} MB) fit in L2 and L2 is NOT flushed: not a headline configuration")}


def _fusion_leg(a, dev, world, rank, H, W, V, timed):
    """Depth filter + fusion of the scan the ranks just reconstructed (north_star: "a single
    NCCL gather for the fused point cloud only").  Every rank owns the depth / confidence maps of its reference views;
    a step = one NCCL all_gather of the depth maps (a reference view's source views may live on other ranks), one fused
    filter launch per owned reference view against its V-1 neighbours, and the single variable-length NCCL gather of
    the fused points to rank 0.  Synthetic scan: a tilted plane seen by laterally shifted cameras (tests/helpers)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from diffmvs_b200 import fusion, sharding
    from tests.helpers import plane_scene
    n_views = max(V, 2 * world)          # every reference view has V-1 source views, as in the depth-estimation step
    sc = plane_scene(H, W, n_views, 7)
    mine = list(sharding.shard_views(n_views, rank, world))
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    local_depth = torch.stack([t(sc["depth"][v]) for v in mine], 0)
    confs = [t(c) for c in sc["conf"]]
    img = t(sc["img"])
    counts = [len(sharding.shard_views(n_views, r, world)) for r in range(world)]
    n_src = min(V - 1, n_views - 1)
    neighbours = {v: [u for u in sorted(range(n_views), key=lambda u: (abs(u - v), u)) if u != v][:n_src] for v in mine}
    mats = {v: torch.from_numpy(np.stack([fusion.pair_matrices(sc["K"], sc["E"][v], sc["K"], sc["E"][u]) for u in neighbours[v]])).to(dev)
            for v in mine}
    state = {}

    def step():
        all_depth = sharding.all_gather_maps(local_depth, counts) if world > 1 else local_depth
        pts, cols = [], []
        for i, v in enumerate(mine):
            src = [(all_depth[u], sc["K"], sc["E"][u]) for u in neighbours[v]]
            out = fusion.fuse_view(local_depth[i], sc["K"], sc["E"][v], sc["depth_max"], sc["depth_min"], confs, [0.3, 0.5, 0.5], src,
                                   ref_img=img, geo_mask_thres=min(3, n_src), mats=mats[v])
            pts.append(out["points"])
            cols.append(out["colors"])
        p, c = torch.cat(pts, 0), torch.cat(cols, 0)
        if world > 1:
            p, c = sharding.gather_points(p, c, dst=0)
        state["n"] = 0 if p is None else int(p.shape[0])

    for _ in range(3):
        step()
    steps = max(3, a.steps // 2)
    ms, _ = timed(step, steps)
    return {"value": n_views * steps / (ms / 1e3), "unit": "fused ref-views/s", "ms_per_step": ms / steps,
            "ref_views_per_step": n_views, "source_views": n_src, "points_on_rank0": state.get("n"),
            "all_gather_bytes_per_step": (n_views * H * W * 4) if world > 1 else 0,
            "exchange": "NCCL all_gather of depth maps + one variable-length NCCL gather of 15-byte point records" if world > 1
                        else "single GPU: no exchange",
            "kernel": "fuse_view_kernel (one launch per reference view: all source views, masks, averaged depth, points)"}


def run_ours(a):
    import torch
    import torch.distributed as dist
    from diffmvs_b200 import _cabi, ops, synth
    from diffmvs_b200.models import CasDiffMVS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.lib()
    if a.load_tuned:
        ops.load_tuned(a.load_tuned)

    variant, H, W, V, D0 = synth.WORKLOADS[a.workload]
    args = synth.workload_args(a.workload)
    model = CasDiffMVS(args, test=True)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synth.synth_state_dict(shapes, 123), strict=False)
    model.to(dev).eval()
    if not a.no_graph:
        model.use_cuda_graph(True)   # public switch: the forward is captured once and replayed as one CUDA graph

    # each rank owns its own reference views (weak scaling: one ref-view per step per GPU)
    imgs, proj, dv = synth.workload_inputs(a.workload, seed=rank)
    h_imgs = [i.pin_memory() for i in imgs]
    # the same images as 8-bit data (what a loader decodes from disk); the model divides by 255 on the device
    h_imgs_u8 = [(i * 255).round().clamp(0, 255).to(torch.uint8).pin_memory() for i in imgs]
    host = {"imgs": h_imgs}
    h_proj = {k: v.pin_memory() for k, v in proj.items()}
    h_dv = dv.pin_memory()
    d_imgs = [i.to(dev) for i in imgs]
    d_proj = {k: v.to(dev) for k, v in proj.items()}
    d_dv = dv.to(dev)
    h2d = sum(t.numel() * 4 for t in h_imgs) + sum(t.numel() * 4 for t in h_proj.values()) + h_dv.numel() * 4
    n_maps = 1 + (3 if variant == "casdiffmvs" else 2)           # final depth + the full-resolution confidence maps
    gather_buf = [torch.empty((n_maps, H, W), device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None

    def step_resident():
        out = model(d_imgs, d_proj, d_dv)
        if world > 1:   # single gather of the per-view results (depth + confidences) to rank 0
            maps = torch.cat([out["depth"][-1]] + list(out["photometric_confidence"]), 0)
            dist.gather(maps, gather_buf, dst=0)              # same exchange as sharding.gather_maps
        return out

    h_out = {}

    # End-to-end: every step copies its inputs from pinned host memory and returns its results to pinned host
    # memory.  Like a DataLoader with pin_memory, the H2D copy of step i+1 runs on a copy stream while step i
    # computes (two device input slots), and like a writer thread the D2H copy of step i's maps runs on a second
    # copy stream while step i+1 computes (two pinned result slots; the host picks up result i-1 after launching
    # step i).  Every byte still moves inside the timed region, which ends with a full device synchronisation.
    copy_stream = torch.cuda.Stream(device=dev)
    d2h_stream = torch.cuda.Stream(device=dev)
    d2h_done = [None, None]
    slots = [None, None]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"i": 0, "d2h": 0}

    def prefetch(slot):
        if slots[slot] is None:                          # persistent device input slots (no allocator traffic per step)
            slots[slot] = ([torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host["imgs"]],
                           {k: torch.empty(t.shape, dtype=t.dtype, device=dev) for k, t in h_proj.items()},
                           torch.empty(h_dv.shape, dtype=h_dv.dtype, device=dev))
            torch.cuda.current_stream().synchronize()
        di, dp, dd = slots[slot]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])      # the forward that read this slot has finished
            for dst, src in zip(di, host["imgs"]):
                dst.copy_(src, non_blocking=True)
            for k, src in h_proj.items():
                dp[k].copy_(src, non_blocking=True)
            dd.copy_(h_dv, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e():
        i = e2e_state["i"]
        cur = i & 1
        if slots[cur] is None:
            prefetch(cur)                                # first step: its own copy, not overlapped
        main = torch.cuda.current_stream()
        main.wait_event(ready[cur])
        di, dp, dd = slots[cur]
        out = model(di, dp, dd)
        consumed[cur].record(main)
        prefetch(cur ^ 1)                                # next step's inputs, overlapped with this forward
        res = [out["depth"][-1]] + list(out["photometric_confidence"])
        produced = torch.cuda.Event()
        produced.record(main)
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(produced)
            for k, t in enumerate(res):
                if (cur, k) not in h_out:
                    h_out[(cur, k)] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                t.record_stream(d2h_stream)
                h_out[(cur, k)].copy_(t, non_blocking=True)
            d2h_done[cur] = torch.cuda.Event()
            d2h_done[cur].record(d2h_stream)
        if d2h_done[cur ^ 1] is not None:
            d2h_done[cur ^ 1].synchronize()              # the caller now holds the previous view's maps (test.py:130)
        e2e_state["i"] = i + 1
        e2e_state["d2h"] = sum(t.numel() * 4 for t in res)
        return e2e_state["d2h"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms[0].item(), ms[1].item()

    with torch.no_grad():
        for _ in range(max(a.warmup, 3)):
            step_resident()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        l0 = ops.launch_count()
        ms_dev, ms_wall = timed(step_resident, a.steps)
        launches = ops.launch_count() - l0   # direct launches + kernels inside replayed graphs
        clocks = sampler.stop() if rank == 0 else None
        # end to end through the public API with host buffers
        d2h = step_e2e()
        step_e2e()
        _, e2e_wall = timed(step_e2e, a.steps)
        # the same loop fed with uint8 host images (extension of the public call: 4x fewer H2D bytes, ...)
        host["imgs"] = h_imgs_u8
        slots[0] = slots[1] = None
        torch.cuda.synchronize()
        for _ in range(3):               # new input signature: captures a second graph, then steady state
            step_e2e()
        _, e2e_u8_wall = timed(step_e2e, a.steps)
        host["imgs"] = h_imgs
        slots[0] = slots[1] = None
        # scan mode: consecutive reference views of a scan share V-1 of their V images; with the
        # cross-ref-view feature cache FeatureNet runs on ONE new image per reference view.  Different unit of work
        # (the parity-pinned headline re-encodes all V images like test.py does), so it is reported beside it.
        scan = None
        if world == 1 and not a.no_scan_mode:
            from diffmvs_b200.scan import ScanRunner
            runner = ScanRunner(model, capacity=2 * V)
            pool = [d_imgs[i % V].roll(shifts=7 * (i // V), dims=-1) for i in range(V + a.steps + 6)]
            state = {"s": 0}

            def step_scan():
                s0 = state["s"]
                ids = [s0 + V - 1 - k for k in range(V)]          # newest image = reference view, the rest seen before
                out = runner(ids, [pool[i] for i in ids], d_proj, d_dv)
                state["s"] = s0 + 1
                return out
            for _ in range(5):
                step_scan()
            h0, m0 = runner.cache.hits, runner.cache.misses
            ms_scan, _ = timed(step_scan, a.steps)
            scan = {"value": a.steps / (ms_scan / 1e3), "unit": UNIT, "ms_per_step": ms_scan / a.steps,
                    "new_images_per_step": (runner.cache.misses - m0) / a.steps,
                    "cached_pyramids_per_step": (runner.cache.hits - h0) / a.steps,
                    "note": "FeatureNet pyramids of already-seen images are reused across reference views (scan.ScanRunner, "
                            "LRU on the device); depth maps equal the uncached call (tests/test_gpu_model.py)"}
        # several reference views per call (the model's batch dimension): the per-view work is unchanged, the ~300 small
        # launches of the refinement stages are shared.  Reported beside the headline, which keeps test.py's batch of 1.
        batched = None
        if world == 1 and not a.no_batched:
            Bb = 4
            bi, bp, bd = synth.workload_inputs(a.workload, seed=rank, batch=Bb)
            bi = [t.to(dev) for t in bi]
            bp = {k: v.to(dev) for k, v in bp.items()}
            bd = bd.to(dev)
            for _ in range(4):               # new input signature: tunes unseen layers, captures its graph
                model(bi, bp, bd)
            nb = max(3, a.steps // Bb)
            ms_b, _ = timed(lambda: model(bi, bp, bd), nb)
            batched = {"value": Bb * nb / (ms_b / 1e3), "unit": UNIT, "ref_views_per_call": Bb, "ms_per_call": ms_b / nb,
                       "note": "model(imgs, proj, depth_values) with a batch of reference views, inputs resident"}
            del bi, bp, bd
        fusion_leg = None
        if not a.no_fusion:
            try:
                fusion_leg = _fusion_leg(a, dev, world, rank, H, W, V, timed)
            except Exception as e:      # cv2 / test helpers missing: report, do not fail the bench
                fusion_leg = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
        # per-kernel-family device time over one more step (CUDA events on the launch stream)
        model.use_cuda_graph(False)          # per-call events need the eager path
        step_resident()                      # untimed: lets the caching allocator serve this stream without cudaMalloc
        torch.cuda.synchronize()
        prof = ops.Profiler()
        # the host needs longer to enqueue an eager step than the device needs to run it: park the stream behind a
        # device-side delay first, so that the event pairs bracket kernels that run back to back (otherwise every short
        # launch is charged the host's enqueue latency)
        from diffmvs_b200 import pipeline as _pl
        branches_were = _pl.Branch.enabled
        _pl.Branch.enabled = False           # one stream: concurrent branches would each be charged the overlapped time
        torch.cuda._sleep(int(4e8))
        ops.set_profiler(prof)
        step_resident()
        ops.set_profiler(None)
        _pl.Branch.enabled = branches_were
        model.use_cuda_graph(not a.no_graph)
        summ = prof.summary() if rank == 0 else {}
        # the other arithmetic modes of the convolutions, device-resident timing only (same storage: fp32)
        alt = {}
        if world == 1 and not a.no_alt_modes:
            base = ops.get_precision()
            for mode in ("fp32", "ws_tf32x3", "ws_tf32"):
                if mode == base:
                    continue
                ops.set_precision(mode)
                for _ in range(2):
                    step_resident()
                ms_alt, _ = timed(step_resident, max(3, a.steps // 2))
                alt[mode] = {"value": max(3, a.steps // 2) / (ms_alt / 1e3), "unit": UNIT}
            ops.set_precision(base)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = ms_dev / a.steps
    value = world * a.steps / (ms_dev / 1e3)
    e2e_value = world * a.steps / (e2e_wall / 1e3)
    e2e_u8_value = world * a.steps / (e2e_u8_wall / 1e3)
    h2d_u8 = h2d - sum(t.numel() * 3 for t in h_imgs_u8)
    peak, peak_src = _peaks()
    # Per-kernel device time of one eager step (CUDA events on the launch stream around every C-ABI call).  Convolution
    # calls are attributed to the CUDA kernel behind the back end that ran them (conv_ws2_kernel, conv_ws_kernel,
    # conv_kernel, ...); everything else to its own kernel.
    kern = {}
    for (name, tag), r in summ.items():
        kname = tag.split("|", 1)[0] if name == "conv" and "|" in tag else name
        k = kern.setdefault(kname, {"ms": 0.0, "bytes": 0, "calls": 0})
        k["ms"] += r["ms"]; k["bytes"] += r["bytes"]; k["calls"] += r["calls"]
    tot_ms = sum(k["ms"] for k in kern.values()) or 1.0
    top_name = max(kern, key=lambda k: kern[k]["ms"]) if kern else "n/a"
    top = kern.get(top_name, {"ms": 1.0, "bytes": 0, "calls": 1})
    conv_names = [k for k in kern if k.startswith("conv")]
    conv_ms = sum(kern[k]["ms"] for k in conv_names)
    conv_bytes = sum(kern[k]["bytes"] for k in conv_names)
    scale = (H * W * V) / (1152 * 1600 * 7)
    contract_bytes = ALGO_BYTES_CFG3 * scale                  # per ref-view, ideal per-operator fusion
    contract_gbs = contract_bytes / (ms_step / 1e3) / 1e9
    # DRAM traffic of the convolution kernels over one step, from the committed ncu launch list (profiles/)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "conv_dram_traffic.json")
    if a.workload == "cfg3" and os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic, traffic_src = float(tj["conv_dram_bytes_per_step"]), tj.get("source")
        except Exception:
            pass
    top_gbs = top["bytes"] / (top["ms"] / 1e3) / 1e9 if top["ms"] > 0 else 0.0
    roofline = {
        # headline: the whole step against SURVEY.md 8(d)'s contract figure (2.75 GB per ref-view at cfg3, scaled by
        # H*W*V for other workloads) - what an ideally fused implementation would have to move
        "bound": "hbm", "achieved": contract_gbs, "peak": peak, "unit": "GB/s", "frac": contract_gbs / peak,
        "algorithmic_bytes": contract_bytes, "basis": "SURVEY 8(d) operator-boundary bytes of one ref-view / device time of one step",
        "peak_source": peak_src,
        "traffic": traffic, "traffic_source": traffic_src, "traffic_scope": "all convolution kernels, one step",
        # the dominant kernel on its own: layer-wise algorithmic bytes (inputs + outputs + weights of each launch)
        "kernel": {"name": top_name, "launches_per_step": top["calls"], "avg_us": 1e3 * top["ms"] / max(top["calls"], 1),
                   "algorithmic_bytes_per_launch": top["bytes"] / max(top["calls"], 1), "achieved": top_gbs,
                   "frac": top_gbs / peak, "share_of_step": top["ms"] / tot_ms,
                   "basis": "layer-wise bytes (every launch's own inputs + outputs), CUDA events per launch, eager step "
                            "on one stream, enqueued behind a device-side delay (kernels run back to back)"},
        "conv_family": {"ms": conv_ms, "algorithmic_bytes": conv_bytes, "achieved": conv_bytes / (conv_ms / 1e3) / 1e9 if conv_ms else 0.0,
                        "frac": conv_bytes / (conv_ms / 1e3) / 1e9 / peak if conv_ms else 0.0, "share_of_step": conv_ms / tot_ms},
        "kernels_ms": {k: round(v["ms"], 3) for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])},
    }
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        cores = _best_cpu_threads(os.cpu_count() or 1)
        run = _oracle_forward_cpu(a.workload, cores)
        t = run()
        if t < 15.0:
            t = min(t, run())
        kind = _reference_kind()
        what = "the reference's own CasDiffMVS module" if kind == "reference" else "oracle port"
        cpu = {"value": 1.0 / t, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"1 ref-view at {a.workload} (one full step), {what} (torch CPU fp32) on {cores} of "
                         f"{os.cpu_count()} host threads (fastest of a thread-count sweep)"}
    gpu_base = None
    if world == 1 and not a.no_gpu_baseline:
        try:
            gpu_base = _gpu_baseline(a.workload, dev)
            gpu_base["speedup_vs_tf32_default"] = value / gpu_base["tf32_default"]["value"]
            gpu_base["speedup_vs_fp32_strict"] = value / gpu_base["fp32_strict"]["value"]
        except Exception as e:   # e.g. out of memory on a smaller part: report, do not fail the bench
            gpu_base = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "precision": ops.get_precision(), "cuda_graph": not a.no_graph, "alt_modes": alt,
        "config": _config(a.workload),
        "parallelism": f"ref-views sharded one per GPU per step over {world} GPU(s); no data-path collective, one NCCL gather "
                       f"of depth + confidence maps to rank 0 per step",
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "e2e_u8": {"value": e2e_u8_value, "unit": UNIT, "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": d2h,
                   "note": "same call with uint8 host images (as decoded from disk), /255 on the device: depth maps "
                           "bit-identical to the float32-image call (tests/test_gpu_kernels.py)"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "gpu_baseline": gpu_base,
        "scan_mode": scan,
        "batched": batched,
        "fusion": fusion_leg,
        "wall_ms_per_step": ms_wall / a.steps,
    }
    print(json.dumps(line), flush=True)
    if a.dump_tuned:
        ops.save_tuned(a.dump_tuned)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt-modes", action="store_true")
    ap.add_argument("--no-fusion", action="store_true", help="skip the depth-filter / point-cloud fusion leg")
    ap.add_argument("--no-scan-mode", action="store_true", help="skip the feature-cache (scan mode) timing")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock-torch CUDA timing of the reference algorithm")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of one CUDA graph")
    ap.add_argument("--no-batched", action="store_true", help="skip the several-reference-views-per-call leg")
    ap.add_argument("--dump-tuned", default=None, help="write the per-layer autotuning table (JSON) to this path")
    ap.add_argument("--load-tuned", default=None, help="preload an autotuning table (use under a profiler, whose "
                    "timings would mislead the tuner)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
