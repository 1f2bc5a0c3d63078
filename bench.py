#!/usr/bin/env python
"""Benchmark of the kt_service imaging hot path (BASELINE.json metric: CT slices/s, series -> labels).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline workload (BASELINE.json configs[2], the largest single-GPU configuration): synthetic 320-slice
512x512 int16 series, dicom_sequences_auto in throughput mode -- coronal rib scan + slice pick for the
series AND every slice through body mask -> HU window/NCHW -> YOLO11s-seg (own tcgen05 convolutions,
random-init, class bias shifted) -> NMS -> mask decode/overlay -> label clean-up.  With N GPUs the batch is
N such series, each cut into N contiguous z-ranges (weak scaling: 320 slices per GPU per step); the only
exchange is one all-gather of coronal rows + min/max and one all-reduce of the selected indices.

One "step" = one pass over the batch.  ``value`` counts slices/s with the int16 pixels already in HBM;
``e2e`` is the same pass through ``SeriesBatchRunner.step_host`` from pinned HOST memory with the
host->device copy of the pixels and the device->host copy of the label maps inside the timed region.
Inputs (168 MB per GPU) exceed the 126 MB L2, so no explicit L2 flush is needed.

The same JSON line also carries (each measured in this run, none inside the headline's timed region):
``roofline`` (dominant own kernel) and ``roofline_kernels`` (every own kernel of the step),
``kernels_isolated`` (the SURVEY §8 kernels alone on 160 slices), ``single_slice`` (configs[0] dicom_frame and
configs[1] PNG latency through the mirror entry points), ``teacher_heads`` (the chunk path with realistic
detections), ``mesh_labelling`` (configs[4]: synthetic 200 k-triangle mesh and the reference's 16 k-vertex
polygon set), ``strong_scaling`` (one series over the N GPUs, N > 1) and ``config3`` (64 series over 8 GPUs).
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
    ap.add_argument("--chunk", type=int, default=320, help="slices per CNN batch / CUDA graph")
    ap.add_argument("--first-chunk", type=int, default=0, help="size of a smaller first chunk (0: all chunks equal)")
    ap.add_argument("--chunks", default="", help="explicit comma-separated chunk sizes (must add up to the local slice count)")
    ap.add_argument("--slices", type=int, default=N_SLICES)
    ap.add_argument("--series", type=int, default=0, help="series per step (default: one per GPU = weak scaling; "
                    "64 with --gpus 8 is BASELINE configs[3]; 1 with --gpus N is strong scaling)")
    ap.add_argument("--engine", default="eitb", choices=["eitb", "cudnn"], help="network engine (own K11/K12 kernels or cuDNN + K9)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-mesh", action="store_true", help="skip the configs[4] mesh element classification measurement")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip isolated kernels, latency, teacher heads, strong scaling, config3")
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--label-fan", type=int, default=1, help="parallel graph branches of the K2 / K5-K7 chains per chunk")
    ap.add_argument("--no-overlap", action="store_true", help="one graph per chunk on one stream, steps joined one by one "
                    "(default: K2 / K5-K7 graphs on side streams beside the CNN graphs, host passes two deep)")
    return ap.parse_args()


def config(world, series=0, slices=N_SLICES):
    s = series or world
    if s == world:
        w = f"configs[2]: synthetic {slices}-slice 512x512 int16 series, dicom_sequences_auto (rib-slice selection + every slice " \
            f"segmented and labelled)" + (f"; {world} series, each cut by slice over the {world} GPUs" if world > 1 else "")
    elif s == 1:
        w = f"configs[2] strong scaling: ONE synthetic {slices}-slice series sharded by slice across {world} GPUs"
    else:
        w = f"configs[3]: batch of {s} synthetic {slices}-slice series sharded by slice across {world} GPUs"
    return {"workload": w, "series_per_step": s, "slices_per_series": slices, "slice": [SIZE, SIZE],
            "sharding": "z-range per GPU, one all-gather of coronal rows + min/max, one all-reduce of indices" if world > 1 else "single GPU",
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
class _Results:
    """Duck-typed ultralytics Results for the reference's create_segmentations_masks (utils.py:476-478)."""

    class _B:
        pass

    def __init__(self, masks_u8, cls, size):
        import torch
        self.masks = _Results._B()
        self.masks.data = masks_u8 if hasattr(masks_u8, "numpy") else torch.from_numpy(masks_u8)
        self.boxes = _Results._B()
        self.boxes.cls = cls if hasattr(cls, "numpy") else torch.from_numpy(cls)
        self.orig_shape = (size, size)


def cpu_path_rate(budget_s: float, threads: int, steps: int = 1, warmup: int = 0, n_slices: int = N_SLICES):
    """The reference's CPU path on the host cores: per series the coronal rib scan (reformat, rib network, NMS, slice
    pick), per slice norm -> body mask -> CNN -> NMS -> mask decode -> label image.  The functions the reference owns
    (classic_norm, get_axial_slice_body_mask, create_segmentations_masks, create_color_output) run as the
    reference's OWN code from oracle/_ref when it is built (kind "reference"); the ultralytics-owned stages run as the
    oracle's restatement and the CNN as the same PyTorch module in fp32.  A step = the rib scan + a bounded sample of
    slices; slices/s is quoted for the whole series: n_slices / (t_rib + n_slices * t_slice).
    Returns (slices/s, kind, sample text, ms per step)."""
    import numpy as np
    import torch
    from eitsynthai_b200 import synth
    from eitsynthai_b200.yolo_seg import YOLO11sSeg
    from oracle import cpu_path, imaging as O, ref_import, yolo_post as Y

    torch.set_num_threads(threads)
    try:
        import cv2
        cv2.setNumThreads(threads)
    except Exception:
        cv2 = None
    ref = ref_import.load_reference_utils() if (cv2 is not None and ref_import.reference_available()) else None
    torch.manual_seed(1)
    model = YOLO11sSeg(4).eval()
    ribs = YOLO11sSeg(1).eval()
    px0 = synth.phantom_slice(0)
    with torch.no_grad():      # same class-bias idea as the GPU arm so NMS / mask decode see candidates
        model.shift_class_bias(Y.preprocess(O.apply_mask(O.classic_norm(px0), O.body_mask(px0, -1024, 1)), SIZE))

    def one_slice(px):
        if ref is None:
            return cpu_path.segment_slice_cpu(px, model)
        norm = ref.classic_norm(px)
        body = ref.get_axial_slice_body_mask(ref_import.DuckDataset(px))
        x = Y.preprocess(cv2.bitwise_and(norm, norm, mask=body), SIZE, torch.float32)
        with torch.no_grad():
            head, protos = model(x)
        r = Y.postprocess(head[0].float(), protos[0].float(), 4, (SIZE, SIZE), (SIZE, SIZE))
        d = ref.create_segmentations_masks(_Results(r["masks"], r["cls"], SIZE))
        return ref.create_color_output(d, body)

    # coronal image of a series: only the mid rows matter, so build them without generating 320 full slices
    rows = np.stack([synth.phantom_slice(1000 + z)[SIZE // 2] for z in range(0, n_slices, 8)]).repeat(8, 0)[:n_slices]
    vol_rows = np.zeros((n_slices, SIZE, SIZE), np.int16)
    vol_rows[:, SIZE // 2] = rows

    def rib_scan():
        return cpu_path.rib_select_cpu(vol_rows, ribs)[0]

    t0 = time.perf_counter()
    one_slice(px0)                                                  # warm-up + cost probe
    per_slice = time.perf_counter() - t0
    t0 = time.perf_counter()
    rib_scan()
    per_rib = time.perf_counter() - t0
    n = max(1, min(32, int((budget_s - per_rib) / max(per_slice, 1e-3) / max(steps + warmup, 1))))
    slices = [synth.phantom_slice(100 + i) for i in range(n)]
    t_slice, t_rib = [], []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        rib_scan()
        t1 = time.perf_counter()
        for px in slices:
            one_slice(px)
        t2 = time.perf_counter()
        if it >= warmup:
            t_rib.append(t1 - t0); t_slice.append((t2 - t1) / n)
    tr, ts = sum(t_rib) / len(t_rib), sum(t_slice) / len(t_slice)
    rate = n_slices / (tr + n_slices * ts)
    kind = "reference" if ref is not None else "port"
    own = "the reference's own utils.py functions (oracle/_ref byte code) for norm, body mask and label image" if ref is not None \
        else "the oracle port for every stage"
    sample = f"per step: 1 rib scan ({n_slices}x512 coronal image, rib CNN fp32 on CPU, NMS, slice pick: {tr * 1e3:.0f} ms) + {n} phantom " \
             f"slices ({ts * 1e3:.0f} ms each) through {own}; ultralytics stages restated (oracle/yolo_post.py), CNN = same module fp32; " \
             f"slices/s = {n_slices} / (t_rib + {n_slices} * t_slice)"
    return rate, kind, sample, (tr + n * ts) * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    budget = float(os.environ.get("EITB_REF_BUDGET_S", "120"))    # whole run, all steps; bounded so the arm ends in minutes
    rate, kind, sample, ms = cpu_path_rate(budget, threads, max(args.steps, 1), min(args.warmup, 1), args.slices)
    cfg = config(args.gpus, args.series, args.slices)             # the same object as the b200 arm's; the sample is in cpu_baseline
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
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


def _events_ms(torch, fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


_ALL_CPUS = None


def _bind_near_gpu(index: int):
    """Pin this process to the CPUs NVML names as closest to its GPU, so that the pinned host buffers it allocates afterwards
    (series in, label maps out) live on the GPU's NUMA node: with 8 ranks the host-side copies otherwise cross the
    socket interconnect and share one memory controller.  Returns the CPU list, or None when NVML cannot tell."""
    global _ALL_CPUS
    try:
        _ALL_CPUS = os.sched_getaffinity(0)
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
        cpus = sorted(os.sched_getaffinity(0))
        return [cpus[0], cpus[-1], len(cpus)] if cpus else None
    except Exception:
        return None


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from eitsynthai_b200 import cabi, convnet, host, ops, sharded, synth
    from eitsynthai_b200.pipeline import CONF, IOU, MAX_DET, ImagingPipeline, SeriesBatchRunner, SeriesMeta

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpu_affinity = _bind_near_gpu(local)                           # before any pinned allocation: first touch decides the NUMA node
    if world > 1:
        # the exchange is 0.3 MB of coronal rows per step: one NCCL CTA moves it, and more would take SMs away from the
        # persistent one-CTA-per-SM convolution kernels that run concurrently on the main stream
        opts = None
        try:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = 1
            opts.config.min_ctas = 1
        except Exception:
            opts = None
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    torch.backends.cudnn.benchmark = True

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    tf_sustained = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json (measured copy GB/s, sustained bf16 cuBLAS TFLOP/s)" if peaks else "fallback 6650 GB/s / 1400 TFLOP/s"

    pipe = ImagingPipeline(dev, torch.float16, seed=0, engine=args.engine)
    timer = StageTimer(torch)
    profiling = {"on": False}

    def make_runner(S, nslices, chunk):
        """This rank's shard of S series (file order shuffled inside the shard), pinned, + the public throughput engine."""
        z0, z1 = sharded.shard_range(nslices, world, rank)
        vols, insts = [], []
        distinct = min(S, max(world, 2))                          # more series than that reuse the generated pixels
        for s in range(S):
            if s < distinct:
                v, i = synth.phantom_series(nslices, seed=s, shuffle_seed=17 + s, z_range=(z0, z1))
            else:
                v, i = vols[s % distinct], insts[s % distinct]
            vols.append(v); insts.append(i)
        px_host = torch.from_numpy(np.stack(vols)).pin_memory()    # [S, nl, H, W]
        # two label buffers: the host path keeps two passes in flight (submit_host / wait_host)
        labels_host = [torch.empty((S, z1 - z0, SIZE, SIZE), dtype=torch.uint8).pin_memory() for _ in range(2)]
        total = S * (z1 - z0)
        headline = S == (args.series or world) and nslices == args.slices
        runner = SeriesBatchRunner(pipe, [SeriesMeta(i) for i in insts], nslices, SIZE, min(chunk, total), use_graphs=not args.no_graphs,
                                   timer=timer, first_chunk=args.first_chunk if headline else 0, overlap=not args.no_overlap, label_fan=args.label_fan,
                                   chunk_sizes=[int(c) for c in args.chunks.split(",")] if args.chunks and headline else None)
        runner.load(px_host)
        runner.capture()
        return runner, px_host, labels_host

    by_rank = {}

    def timed(fn, steps, warmup, finish=None):
        """``finish`` closes what ``fn`` leaves in flight (side streams, the last host pass); it runs inside the timed region."""
        for _ in range(warmup):
            fn()
        if finish is not None:
            finish()
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
        if finish is not None:
            last = finish()
            out = out if last is None else last
        b.record()
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        timer.on = False
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            every = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(every, ms)
            by_rank["last"] = [float(t) / steps for t in every]   # the reported time is the max; this shows the spread
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, out, timer.totals()

    # ---------------------------------------------------------------- headline: configs[2] (or --series S)
    S = args.series or world
    nslices = args.slices
    runner, px_host, labels_host = make_runner(S, nslices, args.chunk)
    nl = runner.nl
    runner_overlap = bool(runner.overlap)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    def dev_pass(r):
        """Device-resident passes back to back; the runner's side streams are joined once, inside the timed region."""
        return (lambda: r.step_device(join=False)), r.join

    def host_pass(r, ph, lh):
        """Passes from pinned host memory, two in flight: pass i+1 is submitted before the host waits for pass i, so its
        host->device copies run under pass i's kernels; every pass still copies its own pixels in and its own label maps
        and selected-slice table out, and the last pass is waited for inside the timed region."""
        state = {"pending": None, "k": 0}

        def fn():
            h = r.submit_host(ph, lh[state["k"] & 1])
            state["k"] += 1
            prev, state["pending"] = state["pending"], h
            return r.wait_host(prev) if prev is not None else None

        def finish():
            prev, state["pending"] = state["pending"], None
            out = r.wait_host(prev) if prev is not None else None
            r.join()
            return out
        return fn, finish

    fn_d, fin_d = dev_pass(runner)
    ms_dev, sel, _ = timed(fn_d, args.steps, max(args.warmup, 3), fin_d)
    ms_dev_by_rank = by_rank.get("last")
    clocks = sampler.stop() if rank == 0 else None
    # eager pass with per-stage and per-kernel CUDA events (same work, no graphs)
    profiling["on"] = True
    ms_eager, _, stages = timed(runner.step_eager, args.steps, 0)
    kernels = cabi.profile_report()
    profiling["on"] = False
    cabi.profile_enable(False)
    n_launch = sum(c for c, _ in kernels.values()) // max(args.steps, 1)   # own kernels per step (graphs replay the same)
    convnet.STATS = {}                                              # algorithmic work of the convolutions of one step
    runner.step_eager()
    torch.cuda.synchronize(dev)
    conv_stats, convnet.STATS = convnet.STATS, None
    with torch.no_grad():
        ndet_mean = float(torch.cat([o[1] for o in runner.outs]).float().mean()) if runner.outs else float("nan")
    total_slices = S * nslices                                     # all ranks together
    value = total_slices / (ms_dev / 1e3)
    e2e = None
    if not args.no_e2e:
        fn_h, fin_h = host_pass(runner, px_host, labels_host)
        ms_e2e, _, _ = timed(fn_h, args.steps, max(args.warmup, 3), fin_h)
        ms_sync, _, _ = timed(lambda: runner.step_host(px_host, labels_host[0]), args.steps, 3)
        h2d = int(px_host.numel() * 2 + runner.rows_dev.numel() * 2)
        e2e = {"value": total_slices / (ms_e2e / 1e3), "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(labels_host[0].numel() + S * 16), "h2d_gbs_per_rank": h2d / (ms_e2e / 1e3) / 1e9,
               "passes_in_flight": 2, "one_pass_at_a_time": {"value": total_slices / (ms_sync / 1e3), "ms_per_step": ms_sync},
               "api": "SeriesBatchRunner.submit_host / wait_host (step_host = both, one pass at a time)"}

    # ---------------------------------------------------------------- roofline: every own kernel of the step
    px = SIZE * SIZE
    local_slices = S * nl
    # algorithmic bytes per 512x512 slice and launch; CC passes count source + label traffic
    per_slice_bytes = {
        "hu_window_kernel": px * (2 + 1 + (1 if pipe.fused_input else 6)), "thr_bits_kernel": px * 2 + px // 8, "morph5_bits_kernel": px // 4,
        "frame_flood_kernel": px // 4,
        # cc_local / cc_merge / cc_flatten / area / best / write_mask: the general body-mask path; its kernels return at once for
        # the images the dominant-component fast path has answered (all of them here), so bytes per slice would be fiction
        "nms_kernel": 40 * 5376 * 2 + MAX_DET * 38 * 4,
        "mask_decode_kernel": 32 * 128 * 128 * 2 + MAX_DET * 38 * 4 + px,
        "fill_body_kernel": px * 3, "small_first_kernel": px, "small_repaint_kernel": px // 8,
        "contour_cand_kernel": px + 3 * px // 8, "contour_repaint_kernel": 3 * px // 8,
        "head_decode_kernel": 5376 * (64 + 8 + 32 + 40) * 2, "sppf_kernel": 256 * 256 * 5 * 2,
        "stem_conv_kernel": px * (1 if pipe.fused_input else 6) + (px // 4) * 64, "upsample2x_concat_kernel": 2 * (1024 * 768 + 4096 * 512) * 2,
    }
    per_slice_bytes["mask_decode_tc_kernel"] = per_slice_bytes["mask_decode_kernel"]
    kroof = {}
    n_chunks = max(len(runner.bounds), 1)
    for name, (cnt, tot_ms) in kernels.items():
        per_step = tot_ms / args.steps
        ent = {"ms_per_step": round(per_step, 4), "launches_per_step": cnt // max(args.steps, 1)}
        if name == "conv_tc_kernel" and conv_stats.get("by_kind", {}).get("gemm"):
            fl = conv_stats["by_kind"]["gemm"][0]
            ent.update({"bound": "tensor", "achieved": fl / (per_step / 1e3) / 1e12, "peak": tf_sustained, "unit": "TFLOP/s",
                        "flops_per_step": fl})
            ent["frac"] = ent["achieved"] / tf_sustained
        elif per_slice_bytes.get(name):
            # per launch a kernel sees one chunk; K2/K7 sub-kernels run several times per chunk: bytes x launches
            # (the rib network launches the stem / SPPF / head decode / NMS kernels once more per owned series, on one
            # image: those launches are not another chunk's worth of bytes)
            rib_launches = len(runner.mine) if name in ("stem_conv_kernel", "sppf_kernel", "head_decode_kernel", "nms_kernel") else 0
            launches_per_chunk = max(1, round((cnt / args.steps - rib_launches) / n_chunks))
            ach = per_slice_bytes[name] * local_slices * launches_per_chunk / (per_step / 1e3) / 1e9
            ent.update({"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                        "algorithmic_bytes_per_slice": per_slice_bytes[name]})
        kroof[name] = ent
    roof = None
    judged = {k: v for k, v in kroof.items() if "frac" in v}
    if judged:
        top = max(judged, key=lambda k: judged[k]["ms_per_step"])
        e = judged[top]
        roof = {"kernel": top, "bound": e["bound"], "achieved": e["achieved"], "peak": e["peak"], "unit": e["unit"], "frac": e["frac"],
                "traffic": None, "ms_per_step": e["ms_per_step"], "launches_per_step": e["launches_per_step"], "peak_source": peak_src,
                "how": "CUDA events around every launch of the kernel (libeitb200 launch profiler) over an eager pass of the same steps; "
                       "achieved = algorithmic work of the step (convolutions: 2*MACs of every layer, counted by convnet.STATS; streaming "
                       "kernels: SURVEY 8(d) bytes per slice x slices) / summed launch time; traffic (dram bytes) cannot be read inside "
                       "bench.py: the ncu figures are under profiles/"}
        if top == "conv_tc_kernel":
            roof["activation_and_weight_gbs"] = conv_stats.get("bytes", 0) / (e["ms_per_step"] / 1e3) / 1e9
            roof["note"] = "the convolutions of YOLO11s-seg mix bandwidth-bound layers (C <= 64 at 128^2 / 256^2) and tensor-bound ones; " \
                           "both fractions are given: frac = of sustained bf16 tensor peak, activation_and_weight_gbs / %.0f = of HBM copy peak" % hbm

    extras = {}
    if not args.no_extras:
        # -------------------------------------------------------- §8 kernels in isolation, bench-sized inputs (160 slices)
        try:
            nb = min(160, local_slices)
            pxd = runner.flat[:nb]
            body = ops.body_mask(pxd, 1, -1024, True)
            code0 = runner.outs[0][0][:nb].clone() if runner.outs else None
            iso = {}
            t = _events_ms(torch, lambda: ops.hu_window(pxd, body_mask=body, want_u8=False, nchw_dtype=torch.float16, channels_last=False))
            iso["K1_hu_window_nchw"] = {"ms": t, "bytes_per_slice": px * (2 + 1 + 6)}
            t = _events_ms(torch, lambda: ops.hu_window(pxd, body_mask=body, want_u8=False, nchw_dtype=torch.float16, channels_last=True))
            iso["K1_hu_window_nhwc"] = {"ms": t, "bytes_per_slice": px * (2 + 1 + 6)}
            t = _events_ms(torch, lambda: ops.hu_window(pxd, body_mask=body, want_u8=True, nchw_dtype=None))
            iso["K1_hu_window_u8_for_fused_stem"] = {"ms": t, "bytes_per_slice": px * (2 + 1 + 1)}
            t = _events_ms(torch, lambda: ops.body_mask(pxd, 1, -1024, True))
            iso["K2_body_mask"] = {"ms": t, "bytes_per_slice": 786432}
            hd, pr = synth.random_heads(8, 48, seed=3)
            reps = max(nb // 8, 1)
            hd = torch.from_numpy(np.tile(hd, (reps, 1, 1))).to(dev).half()
            pr = torch.from_numpy(np.tile(pr, (reps, 1, 1, 1))).to(dev).half().contiguous(memory_format=torch.channels_last)
            t = _events_ms(torch, lambda: ops.nms(hd, 4, want_idx=False))
            iso["K5_nms"] = {"ms": t, "bytes_per_slice": 430080 + 45600}
            dets, _, nd = ops.nms(hd, 4, want_idx=False)
            for nm_, var in (("K6_mask_decode_tcgen05", 0x20), ("K6_mask_decode_scalar", 0x10), ("K6_mask_decode_warp_mma_default", 0)):
                t = _events_ms(torch, lambda: ops.mask_decode(dets, nd, pr, var))
                iso[nm_] = {"ms": t, "bytes_per_slice": 1048576 + 45600 + 262144, "flops_per_slice": 1048576 * float(nd.float().mean())}
            if code0 is not None:
                t = _events_ms(torch, lambda: ops.label_cleanup(code0.clone(), body))
                t0 = _events_ms(torch, lambda: code0.clone())
                iso["K7_label_cleanup"] = {"ms": t - t0, "bytes_per_slice": 3 * 262144}
            for v in iso.values():
                n_ = hd.shape[0] if "flops_per_slice" in v or v["bytes_per_slice"] == 430080 + 45600 else nb
                v["slices"] = n_
                v["gbs"] = v["bytes_per_slice"] * n_ / (v["ms"] / 1e3) / 1e9
                v["frac_of_hbm_peak"] = v["gbs"] / hbm
            extras["kernels_isolated"] = iso
        except Exception as e:                                      # an optional section must not lose the headline
            extras["kernels_isolated"] = {"error": repr(e)}

        # -------------------------------------------------------- teacher-forced heads: the chunk path with realistic detections
        try:
            th = [synth.teacher_heads(seed=s_) for s_ in range(8)]
            nb = min(args.chunk, local_slices)
            thd = torch.from_numpy(np.stack([th[i % 8][0] for i in range(nb)])).to(dev).half()
            tpr = torch.from_numpy(np.stack([th[i % 8][1] for i in range(nb)])).to(dev).half().contiguous(memory_format=torch.channels_last)
            pxd = runner.flat[:nb]

            def teacher_chunk():
                body_ = ops.body_mask(pxd, 1, -1024, True)
                pipe._net(pipe.axial_model_512, pipe.window_input(pxd, body_))                # the CNN runs; its head is replaced by the teacher's
                d_, _, n_ = ops.nms(thd, 4, CONF, IOU, MAX_DET, want_idx=False)
                c_, _, _ = ops.mask_decode(d_, n_, tpr, pipe.mask_variant)
                ops.label_cleanup(c_, body_)
                return n_
            t = _events_ms(torch, teacher_chunk, n=5, warm=2)
            parts = {}
            body_ = ops.body_mask(pxd, 1, -1024, True)
            d_, _, n_ = ops.nms(thd, 4, CONF, IOU, MAX_DET, want_idx=False)
            parts["K5_nms"] = _events_ms(torch, lambda: ops.nms(thd, 4, CONF, IOU, MAX_DET, want_idx=False), n=5, warm=2)
            parts["K6_mask_decode"] = _events_ms(torch, lambda: ops.mask_decode(d_, n_, tpr, pipe.mask_variant), n=5, warm=2)
            c_, _, _ = ops.mask_decode(d_, n_, tpr, pipe.mask_variant)
            t0 = _events_ms(torch, lambda: c_.clone(), n=5, warm=2)
            parts["K7_label_cleanup"] = _events_ms(torch, lambda: ops.label_cleanup(c_.clone(), body_), n=5, warm=2) - t0
            extras["teacher_heads"] = {"slices_per_chunk": nb, "ms_per_chunk_eager": t, "slices_per_sec_eager": nb / (t / 1e3),
                                       "mean_detections_per_slice": float(n_.float().mean()), "stage_ms_per_chunk": parts,
                                       "note": "same chunk path, eager (no CUDA graph), CNN executed, heads/prototypes replaced by "
                                               "synth.teacher_heads (phantom tissues, 8-32 jittered boxes per structure)"}
        except Exception as e:
            extras["teacher_heads"] = {"error": repr(e)}

        # -------------------------------------------------------- configs[0] / configs[1]: one slice through the mirror entry points
        if rank == 0:
            try:
                import io
                import zipfile
                from eitsynthai_b200.kt_service import kt_service_config as kc
                from eitsynthai_b200.kt_service.ai_tools import ai_tools as A
                from eitsynthai_b200.kt_service.ai_tools import dicom_io
                from oracle import imaging as O
                A._PIPELINES[(str(torch.device(A.config.device())), kc.ribs_segm_model, kc.axial_slice_segm_model_256,
                              kc.axial_slice_segm_model_512)] = pipe
                buf = io.BytesIO()
                with zipfile.ZipFile(buf, "w", zipfile.ZIP_STORED) as zf:
                    zf.writestr("slice.dcm", dicom_io.write_dicom(synth.phantom_slice(0), 1))
                zbytes = buf.getvalue()
                frame, img2 = A.DICOMToMask(), A.ImageToMask()
                png = O.apply_mask(O.classic_norm(synth.phantom_slice(0)), O.body_mask(synth.phantom_slice(0), -1024, 1))

                def lat(fn, n=12):
                    ts, ans = [], None
                    for i in range(n + 2):
                        t0 = time.perf_counter()
                        ans = fn()
                        torch.cuda.synchronize(dev)
                        if i >= 2:
                            ts.append((time.perf_counter() - t0) * 1e3)
                    ts.sort()
                    return ts[len(ts) // 2], ts[0], ans
                m0, b0, a0 = lat(lambda: frame.get_coordinate_slice_from_dicom_frame(io.BytesIO(zbytes)))
                m1, b1, a1 = lat(lambda: img2.get_coordinate_slice_from_image(png))
                extras["single_slice"] = {
                    "config0_dicom_frame": {"median_ms": m0, "best_ms": b0, "ok": bool(a0) and a0.get("status") == "success",
                                            "path": "zip bytes -> DICOM parse -> K2/K1 -> CNN -> K5 -> K6 -> K7 -> polygons -> Delaunay stand-in "
                                                    "mesh -> K8 -> answer dict (DICOMToMask.get_coordinate_slice_from_dicom_frame)",
                                            "mesh_triangles": len(a0["mesh_data"]["CLASS"]) if a0 and a0.get("mesh_data") else 0},
                    "config1_png": {"median_ms": m1, "best_ms": b1, "ok": bool(a1) and a1.get("status") == "success",
                                    "path": "u8 image -> NCHW -> CNN -> K5 -> K6 -> K7 (no body mask) -> polygons -> answer dict "
                                            "(ImageToMask.get_coordinate_slice_from_image)"},
                    "unit": "ms per request, wall clock around the call + device synchronise"}
            except Exception as e:
                extras["single_slice"] = {"error": repr(e)}

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

        def label_bench(nodes_, tris_all_, xy_, off_, pcls_, reps=10):
            t0_, t1_ = sharded.shard_range(len(tris_all_), world, rank)   # polygon table replicated, no collective on the data path
            tris_ = tris_all_[t0_:t1_]
            dm_ = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (nodes_, tris_, xy_, off_, pcls_)]
            if world > 1:
                dist.barrier()
            ms_ = _events_ms(torch, lambda: ops.tri_label(*dm_), n=reps, warm=3)
            if world > 1:
                tm = torch.tensor([ms_], device=dev)
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                ms_ = float(tm)
            return ms_, ops.tri_label(*dm_), dm_, tris_
        ms_mesh, cls_gpu, dm, tris = label_bench(nodes, tris_all, xy, off, pcls)
        mesh = {"triangles": int(len(tris_all)), "triangles_per_gpu": int(len(tris)), "polygons": int(len(pcls)),
                "polygon_vertices": int(len(xy)), "ms": ms_mesh, "elements_per_sec": len(tris_all) / (ms_mesh / 1e3),
                "class_histogram": torch.bincount(cls_gpu, minlength=5).tolist()}
        # the raster look-up the north star also names (centroid pixel of the label map), against the
        # reference's polygon semantics on the same mesh
        cls_raster = ops.tri_label_raster(dm[0], dm[1], code_m[0].contiguous())
        ms_r = _events_ms(torch, lambda: ops.tri_label_raster(dm[0], dm[1], code_m[0].contiguous()))
        mesh["raster_mode"] = {"ms": ms_r, "elements_per_sec_per_gpu": len(tris) / (ms_r / 1e3),
                               "disagreement_vs_polygon_mode": float((cls_raster != cls_gpu).float().mean())}
        # the reference's largest real polygon set (mesh_service_trials.py: 111 polygons, 16 k vertices, mm coordinates)
        try:
            z6 = np.load(os.path.join(ROOT, "tests", "golden", "reference_polygon_sets.npz"))
            rx, ro, rc = z6["set6_xy"], z6["set6_off"], z6["set6_cls"]
            cont = [[float(rc[p_])] + rx[ro[p_]:ro[p_ + 1]].reshape(-1).tolist() for p_ in range(len(rc))]
            outer6 = next((i_ for i_, c_ in enumerate(cont) if int(c_[0]) == 4), None)       # create_mesh drops the outer contour
            xy6, off6, cls6 = host.prepare_polygons([c_ for i_, c_ in enumerate(cont) if i_ != outer6])
            lo, hi = xy6.min(0), xy6.max(0)
            pitch = float(np.sqrt((hi[0] - lo[0]) * (hi[1] - lo[1]) * 2 / 200000.0))
            nodes6, tris6 = synth.delaunay_mesh((lo[0], lo[1], hi[0], hi[1]), pitch, seed=0)
            ms6, cls6_gpu, _, _ = label_bench(nodes6, tris6, xy6, off6, cls6, reps=5)
            mesh["reference_set6"] = {"triangles": int(len(tris6)), "polygons": int(len(cls6)), "polygon_vertices": int(len(xy6)),
                                      "ms": ms6, "elements_per_sec": len(tris6) / (ms6 / 1e3),
                                      "class_histogram": torch.bincount(cls6_gpu, minlength=5).tolist()}
        except Exception as e:
            mesh["reference_set6"] = {"error": repr(e)}
        if not args.no_cpu_baseline and rank == 0:
            from oracle import tri_label as TL                      # CPU baseline leg: the C restatement, one core
            nb = min(20000, len(tris))
            t0 = time.perf_counter()
            ref = TL.label_triangles(nodes, tris[:nb], xy, off, pcls)
            dt = time.perf_counter() - t0
            mesh["cpu_elements_per_sec"] = nb / dt
            mesh["cpu_sample"] = f"first {nb} triangles, oracle/tri_label.c, 1 core"
            mesh["labels_match_cpu"] = bool(np.array_equal(ref, cls_gpu[:nb].cpu().numpy()))

    # ---------------------------------------------------------------- N > 1: strong scaling and configs[3]
    if world > 1 and not args.no_extras and not args.series:
        try:
            del runner
            torch.cuda.empty_cache()
            r1, ph1, lh1 = make_runner(1, nslices, args.chunk)
            fn1, fin1 = dev_pass(r1)
            ms1, _, _ = timed(fn1, max(args.steps, 10), 3, fin1)
            fn1, fin1 = host_pass(r1, ph1, lh1)
            ms1h, _, _ = timed(fn1, max(args.steps, 10), 3, fin1)
            extras["strong_scaling"] = {"workload": config(world, 1, nslices)["workload"], "slices_per_gpu": r1.nl, "ms_per_step": ms1,
                                        "slices_per_sec": nslices / (ms1 / 1e3), "e2e_ms_per_step": ms1h,
                                        "e2e_slices_per_sec": nslices / (ms1h / 1e3)}
            del r1, ph1, lh1
            torch.cuda.empty_cache()
        except Exception as e:
            extras["strong_scaling"] = {"error": repr(e)}
        if world == 8:
            try:
                r3, ph3, lh3 = make_runner(64, nslices, args.chunk)
                fn3, fin3 = dev_pass(r3)
                ms3, _, _ = timed(fn3, 5, 2, fin3)
                fn3, fin3 = host_pass(r3, ph3, lh3)
                ms3h, _, _ = timed(fn3, 5, 2, fin3)
                extras["config3"] = {"workload": config(world, 64, nslices)["workload"], "slices_per_gpu_per_step": 64 * r3.nl,
                                     "ms_per_step": ms3, "slices_per_sec": 64 * nslices / (ms3 / 1e3), "e2e_ms_per_step": ms3h,
                                     "e2e_slices_per_sec": 64 * nslices / (ms3h / 1e3)}
                del r3, ph3, lh3
            except Exception as e:
                extras["config3"] = {"error": repr(e)}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            if _ALL_CPUS:
                os.sched_setaffinity(0, _ALL_CPUS)                  # the CPU baseline may use every host core again
            threads = os.cpu_count() or 1
            rate, kind, sample, _ = cpu_path_rate(25.0, threads, 1, 0, nslices)
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True,
                "scaling": "strong" if (args.series and args.series < world) else "weak",
                "vs_baseline": None, "dtype": "f16 (CNN, fp32 accumulate) / int16,u8,f32,f64 (kernels)", "data": "synthetic",
                # `config` names the workload and is the same object in both arms (--impl reference); how THIS arm ran it
                # is under `run`
                "config": config(world, S, nslices),
                "run": dict(chunk=args.chunk, engine=args.engine, class_bias_shift=pipe.bias_shift,
                            mean_detections_per_slice=ndet_mean, cuda_graphs=not args.no_graphs,
                            overlap=runner_overlap, label_fan=args.label_fan, cpu_affinity_first_last_count=cpu_affinity),
                "clocks": clocks, "e2e": e2e, "gpu_launches": n_launch, "roofline": roof, "cpu_baseline": cpu,
                "roofline_kernels": kroof, "conv_work_per_step": {k: v for k, v in conv_stats.items() if k != "by_kind"},
                "eager_profiled_ms_per_step": ms_eager,
                "stage_ms_per_step": {k: v / args.steps for k, v in sorted(stages.items())},
                "kernel_ms_per_step": {k: round(v[1] / args.steps, 4) for k, v in sorted(kernels.items(), key=lambda kv: -kv[1][1])},
                "selected_slices": sel.cpu().tolist(), "mesh_labelling": mesh}
        if ms_dev_by_rank:
            line["ms_per_step_by_rank"] = ms_dev_by_rank
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
