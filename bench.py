#!/usr/bin/env python
"""Benchmark of the kt_service imaging hot path (BASELINE.json metric: CT slices/s, series -> labels).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[2], the largest single-GPU configuration): synthetic 320-slice
512x512 int16 series, dicom_sequences_auto in throughput mode -- coronal rib scan + slice pick for
the series AND every slice through body mask -> HU window/NCHW -> YOLO11s-seg (PyTorch/cuDNN,
random-init, class bias shifted) -> NMS -> mask decode/overlay -> label clean-up.  With N GPUs the
batch is N such series, each cut into N contiguous z-ranges (weak scaling: 320 slices per GPU per
step); the only exchange is the all-gather of coronal rows/min-max and of the selected indices.

One "step" = one pass over the batch.  ``value`` counts slices/s with the int16 pixels already in
HBM; ``e2e`` is the same pass through ``ImagingPipeline`` from pinned HOST memory with the
host->device copy of the pixels and the device->host copy of the label maps inside the timed
region.  Inputs (168 MB per GPU) exceed the 126 MB L2, so no explicit L2 flush is needed.
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
sys.path.insert(0, ROOT)

N_SLICES = 320
SIZE = 512
METRIC = "ct_slices_per_sec_series_to_labels"
UNIT = "slices/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chunk", type=int, default=160, help="slices per CNN batch / CUDA graph")
    ap.add_argument("--first-chunk", type=int, default=0, help="size of a smaller first chunk (0: all chunks equal)")
    ap.add_argument("--chunks", default="", help="explicit comma-separated chunk sizes (must add up to the local slice count)")
    ap.add_argument("--slices", type=int, default=N_SLICES)
    ap.add_argument("--series", type=int, default=0, help="series per step (default: one per GPU = weak scaling; "
                    "64 with --gpus 8 is BASELINE configs[3])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mesh", action="store_true", help="skip the configs[4] mesh element classification measurement")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    return ap.parse_args()


def config(world, series=0):
    return {"workload": "configs[2]: synthetic 320-slice 512x512 int16 series, dicom_sequences_auto "
                        "(rib-slice selection + every slice segmented and labelled)" if not series or series == world else
                        f"configs[3]: batch of {series} synthetic 320-slice series sharded by slice across {world} GPUs",
            "series_per_step": series or world, "slices_per_series": N_SLICES, "slice": [SIZE, SIZE],
            "sharding": "z-range per GPU, all-gather of coronal rows" if world > 1 else "single GPU",
            "l2": "inputs (168 MB/GPU) larger than L2, no flush", "weights": "random-init YOLO11s-seg x3, class bias shifted"}


# =============================================================================== clocks
class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# =============================================================================== CPU arm
def cpu_path_rate(budget_s: float, threads: int, steps: int = 1, warmup: int = 0):
    """Oracle port of the per-series path on the host cores.  Returns (slices/s, sample text, ms/step)."""
    import numpy as np
    import torch
    from eitsynthai_b200 import synth
    from eitsynthai_b200.yolo_seg import YOLO11sSeg
    from oracle import cpu_path

    torch.set_num_threads(threads)
    try:
        import cv2
        cv2.setNumThreads(threads)
    except Exception:
        pass
    torch.manual_seed(1)
    model = YOLO11sSeg(4).eval()
    # same class-bias idea as the GPU arm so NMS / mask decode see candidates
    px0 = synth.phantom_slice(0)
    with torch.no_grad():
        from oracle import imaging as O, yolo_post as Y
        x = Y.preprocess(O.apply_mask(O.classic_norm(px0), O.body_mask(px0, -1024, 1)), SIZE)
        model.shift_class_bias(x)
    t0 = time.perf_counter()
    cpu_path.segment_slice_cpu(px0, model)                       # warm-up + cost probe
    per_slice = time.perf_counter() - t0
    n = max(1, min(32, int(budget_s / max(per_slice, 1e-3) / max(steps + warmup, 1))))
    slices = [synth.phantom_slice(100 + i) for i in range(n)]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for px in slices:
            cpu_path.segment_slice_cpu(px, model)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return n / dt, f"{n} phantom slices per step through the oracle port (norm, body mask, CNN fp32 on CPU, NMS, " \
                   f"mask decode, label clean-up); rib scan excluded (1 CNN call per 320 slices)", dt * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    budget = float(os.environ.get("EITB_REF_BUDGET_S", "120"))    # whole run, all steps; bounded so the arm ends in minutes
    rate, sample, ms = cpu_path_rate(budget, threads, max(args.steps, 1), min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config(args.gpus),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# =============================================================================== GPU arm
class StageTimer:
    """CUDA-event pairs around each stage on the launching stream."""

    def __init__(self, torch):
        self.torch, self.pairs, self.on = torch, {}, False

    def __call__(self, name):
        return _Span(self, name)

    def totals(self):
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in self.pairs.items()}


class _Span:
    def __init__(self, t, name):
        self.t, self.name = t, name

    def __enter__(self):
        if self.t.on:
            self.a = self.t.torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if self.t.on:
            b = self.t.torch.cuda.Event(enable_timing=True)
            b.record()
            self.t.pairs.setdefault(self.name, []).append((self.a, b))


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from eitsynthai_b200 import cabi, host, ops, sharded, synth
    from eitsynthai_b200.pipeline import CONF, IOU, MAX_DET, ImagingPipeline, SeriesBatchRunner, SeriesMeta

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True

    pipe = ImagingPipeline(dev, torch.float16, seed=0)
    S = args.series or world                                      # series per step (default: weak scaling, one per GPU)
    nslices = args.slices
    z0, z1 = sharded.shard_range(nslices, world, rank)
    nl = z1 - z0
    # this rank's shard of every series, file order shuffled inside the shard
    vols, insts = [], []
    distinct = min(S, max(world, 2))                              # more series than that reuse the generated pixels
    for s in range(S):
        if s < distinct:
            v, i = synth.phantom_series(nslices, seed=s, shuffle_seed=17 + s, z_range=(z0, z1))
        else:
            v, i = vols[s % distinct], insts[s % distinct]
        vols.append(v); insts.append(i)
    px_host = torch.from_numpy(np.stack(vols)).pin_memory()        # [S, nl, H, W]
    labels_host = torch.empty((S, nl, SIZE, SIZE), dtype=torch.uint8).pin_memory()
    metas = [SeriesMeta(insts[s]) for s in range(S)]
    timer = StageTimer(torch)
    profiling = {"on": False}
    # the public throughput engine of the package; bench.py only times it
    runner = SeriesBatchRunner(pipe, metas, nslices, SIZE, args.chunk, use_graphs=not args.no_graphs, timer=timer,
                               first_chunk=args.first_chunk,
                               chunk_sizes=[int(c) for c in args.chunks.split(",")] if args.chunks else None)
    runner.load(px_host)
    runner.capture()
    graphs, outs = runner.graphs, runner.outs
    step_eager, step_device = runner.step_eager, runner.step_device

    def step_e2e():
        """Same pass through the public API from pinned host memory: H2D of the pixels, D2H of the label
        maps and of the selected indices inside the step."""
        return runner.step_host(px_host, labels_host)

    rows_bytes = runner.rows_dev.numel() * 2

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        cabi.profile_enable(profiling["on"])                       # drops what the warm-up recorded
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        timer.on = True
        timer.pairs = {}
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.profiler.start()                                # no-op unless run under ncu --profile-from-start off
        a.record()
        for _ in range(steps):
            out = fn()
        b.record()
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        timer.on = False
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, out, None, timer.totals()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, sel, _, _ = timed(step_device, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    # ---- eager pass with per-stage and per-kernel CUDA events (same work, no graphs)
    profiling["on"] = True
    torch.cuda.synchronize(dev)
    ms_eager, _, _, stages = timed(step_eager, args.steps, 0)
    kernels = cabi.profile_report()
    profiling["on"] = False
    cabi.profile_enable(False)
    n_launch = sum(c for c, _ in kernels.values())                 # own kernels per timed region (graphs replay the same)
    with torch.no_grad():
        ndet_mean = float(torch.cat([o[1] for o in outs]).float().mean()) if outs else float("nan")
    total_slices = S * nslices                                     # all ranks together
    value = total_slices / (ms_dev / 1e3)

    e2e = None
    if not args.no_e2e:
        ms_e2e, _, _, _ = timed(step_e2e, args.steps, max(args.warmup, 3))
        e2e = {"value": total_slices / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(px_host.numel() * 2 + rows_bytes), "d2h_bytes_per_step": int(labels_host.numel() + S * 16)}

    # ---------------------------------------------------------------- roofline of the dominant own kernel
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    px = SIZE * SIZE
    # algorithmic bytes per 512x512 slice and launch; CC passes count source + label traffic
    per_slice_bytes = {
        "hu_window_kernel": px * (2 + 1 + 6), "thr_bits_kernel": px * 2 + px // 8, "morph5_bits_kernel": px // 4,
        "cc_local_kernel": px * 5, "cc_merge_kernel": 15 * SIZE * 8, "cc_flatten_kernel": px * 8,
        "area_kernel": px * 4, "best_kernel": px * 8, "write_mask_kernel": px * 5,
        "nms_kernel": 40 * 5376 * 2 + MAX_DET * 38 * 4,
        "mask_decode_kernel": 32 * 128 * 128 * 2 + MAX_DET * 38 * 4 + px,
        "fill_body_kernel": px * 3, "small_first_kernel": px, "small_repaint_kernel": px // 8,
        "contour_cand_kernel": px * 5, "contour_repaint_kernel": px // 8,
    }
    per_slice_bytes["mask_decode_tc_kernel"] = per_slice_bytes["mask_decode_kernel"]
    kernels = dict(kernels)
    if "conv_epilogue_kernel" in kernels:                          # K9 has two entry points; account them together
        a, b_ = kernels.pop("conv_epilogue_kernel"), kernels.get("bias_act_kernel", (0, 0.0))
        kernels["bias_act_kernel"] = (a[0] + b_[0], a[1] + b_[1])
    # K9 (conv epilogue): one read + one write of every Conv output of the network, per slice
    act_elems = {}
    def _count(m, i, o):                                         # per slice: C_out x H_out x W_out of every Conv
        st = m.conv.stride[0]
        act_elems["n"] = act_elems.get("n", 0) + m.conv.out_channels * (-(-i[0].shape[2] // st)) * (-(-i[0].shape[3] // st))
    from eitsynthai_b200.yolo_seg import Conv
    hooks = [m.register_forward_hook(_count) for m in pipe.axial_512_torch.modules() if isinstance(m, Conv)]
    with torch.no_grad():
        pipe.axial_512_torch(torch.zeros((1, 3, SIZE, SIZE), dtype=torch.float16, device=dev).contiguous(memory_format=torch.channels_last))
    for h in hooks:
        h.remove()
    # exact bytes of the timed region for K9 are filled in below (the rib network adds its own launches)
    axial_act_bytes = act_elems["n"] * 2 * 2                                    # fp16, read + write, per slice
    per_slice_bytes["bias_act_kernel"] = axial_act_bytes
    own = {k: v for k, v in kernels.items() if k in per_slice_bytes}
    roof = None
    if own:
        top = max(own, key=lambda k: own[k][1])
        cnt, tot_ms = own[top]
        ms_call = tot_ms / cnt
        ach = per_slice_bytes[top] * min(args.chunk, S * nl) / (ms_call / 1e3) / 1e9
        if top == "bias_act_kernel":                                # ~90 launches of different sizes per network call
            n_rib_px = 416 * 640 if nslices == N_SLICES else 0
            step_bytes = axial_act_bytes * S * nl + axial_act_bytes * n_rib_px / (SIZE * SIZE) * len([s for s in range(S) if s % world == rank])
            ach = step_bytes * args.steps / (tot_ms / 1e3) / 1e9
        traffic = None
        try:                                                        # dram bytes per launch from the committed ncu --set full capture
            with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
                traffic = json.load(f).get(top, {}).get("dram_bytes_per_launch")
        except OSError:
            pass
        roof = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                "traffic": traffic, "ms_per_launch": ms_call, "launches": cnt, "slices_per_launch": min(args.chunk, S * nl),
                "algorithmic_bytes_per_slice": per_slice_bytes[top],
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650",
                "how": "CUDA events around every launch of the kernel (libeitb200 launch profiler) over an eager pass of the same steps"}
    # ---------------------------------------------------------------- configs[4]: mesh element classification (K8)
    mesh = None
    if not args.no_mesh:                                          # every rank labels its contiguous block of the elements
        from eitsynthai_b200.kt_service.ai_tools import utils as kt_utils
        hd, pr = synth.teacher_heads(seed=0)
        dets_m, _, n_m = ops.nms(torch.from_numpy(hd[None]).to(dev), 4, want_idx=False)
        code_m, _, _ = ops.mask_decode(dets_m, n_m, torch.from_numpy(pr[None]).to(dev))
        body_m = ops.body_mask(torch.from_numpy(synth.phantom_slice(0)[None]).to(dev), 1, -1024, True)
        ops.label_cleanup(code_m, body_m)
        polys = kt_utils.codes_to_polygons(code_m[0].cpu().numpy(), [0.753906, 0.753906], body_m[0].cpu().numpy())[2:]
        xy, off, pcls = host.prepare_polygons(host.parse_contours(polys, host.find_outer_index(polys)))
        nodes, tris_all = synth.delaunay_mesh((20, 40, 490, 470), 1.43, seed=0)
        t0_, t1_ = sharded.shard_range(len(tris_all), world, rank)   # polygon table replicated, no collective on the data path
        tris = tris_all[t0_:t1_]
        dm = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (nodes, tris, xy, off, pcls)]
        if world > 1:
            dist.barrier()
        for _ in range(3):
            cls_gpu = ops.tri_label(*dm)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            cls_gpu = ops.tri_label(*dm)
        b.record()
        torch.cuda.synchronize(dev)
        ms_mesh = a.elapsed_time(b) / 10
        if world > 1:
            tm = torch.tensor([ms_mesh], device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms_mesh = float(tm)
        mesh = {"triangles": int(len(tris_all)), "triangles_per_gpu": int(len(tris)), "polygons": int(len(pcls)),
                "polygon_vertices": int(len(xy)), "ms": ms_mesh, "elements_per_sec": len(tris_all) / (ms_mesh / 1e3),
                "class_histogram": torch.bincount(cls_gpu, minlength=5).tolist()}
        # the raster look-up the north star also names (centroid pixel of the label map), against the
        # reference's polygon semantics on the same mesh
        cls_raster = ops.tri_label_raster(dm[0], dm[1], code_m[0].contiguous())
        a.record()
        for _ in range(10):
            ops.tri_label_raster(dm[0], dm[1], code_m[0].contiguous())
        b.record()
        torch.cuda.synchronize(dev)
        mesh["raster_mode"] = {"ms": a.elapsed_time(b) / 10, "elements_per_sec_per_gpu": len(tris) / (a.elapsed_time(b) / 10 / 1e3),
                               "disagreement_vs_polygon_mode": float((cls_raster != cls_gpu).float().mean())}
        if not args.no_cpu_baseline and rank == 0:
            from oracle import tri_label as TL                      # CPU baseline leg: the C restatement, one core
            nb = min(20000, len(tris))
            t0 = time.perf_counter()
            ref = TL.label_triangles(nodes, tris[:nb], xy, off, pcls)
            dt = time.perf_counter() - t0
            mesh["cpu_elements_per_sec"] = nb / dt
            mesh["cpu_sample"] = f"first {nb} triangles, oracle/tri_label.c, 1 core"
            mesh["labels_match_cpu"] = bool(np.array_equal(ref, cls_gpu[:nb].cpu().numpy()))

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            rate, sample, _ = cpu_path_rate(20.0, threads)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f16 (CNN) / int16,u8,f32 (kernels)", "data": "synthetic",
                "config": dict(config(world, S), chunk=args.chunk, class_bias_shift=pipe.bias_shift,
                               mean_detections_per_slice=ndet_mean, cuda_graphs=bool(graphs)),
                "clocks": clocks, "e2e": e2e, "gpu_launches": n_launch, "roofline": roof, "cpu_baseline": cpu,
                "eager_profiled_ms_per_step": ms_eager,
                "stage_ms_per_step": {k: v / args.steps for k, v in sorted(stages.items())},
                "kernel_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in sorted(kernels.items(), key=lambda kv: -kv[1][1])},
                "kernel_gbs": {k: round(per_slice_bytes[k] * min(args.chunk, S * nl) * v[0] / (v[1] / 1e3) / 1e9, 1)
                               for k, v in kernels.items() if k in per_slice_bytes and v[1] > 0 and k != "bias_act_kernel"},
                "selected_slices": sel.cpu().tolist(), "mesh_labelling": mesh}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
