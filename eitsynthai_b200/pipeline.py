"""Device-resident imaging pipeline: series -> coronal rib scan -> slice pick -> per-slice
body mask / window / CNN / NMS / mask decode / label clean-up, and mesh labelling.

Mirrors the order of ``DICOMSequencesToMask.get_coordinate_slice_from_dicom``
(kt_service/ai_tools/ai_tools.py:188-231) and of the other four entry points; every stage except
the CNN is one libeitb200 call (``ops``).  Host buffers come in pinned, go to HBM on a copy
stream and the label maps come back the same way, chunk by chunk, so copies overlap compute.
"""
from __future__ import annotations

import logging
import os
from dataclasses import dataclass, field

import numpy as np
import torch

from . import host, ops
from .yolo_seg import build_model

logger = logging.getLogger(__name__)

CONF = 0.3          # ai_tools.py:121,153
IOU = 0.7           # ultralytics default
MAX_DET = 300


@dataclass
class SeriesMeta:
    """The DICOM tags the hot path reads (utils.py:96-105, 621-656; ai_tools.py:384)."""
    instance_numbers: np.ndarray                       # per file, any order
    patient_position: str = "HFS"
    image_orientation: tuple = (1, 0, 0, 0, 1, 0)
    patient_orientation: tuple | None = None
    rescale_slope: int = 1
    rescale_intercept: int = -1024
    pixel_spacing: tuple = (0.753906, 0.753906)


@dataclass
class SeriesResult:
    labels: torch.Tensor | None = None                 # [n, H, W] u8 code image per processed slice (device)
    body: torch.Tensor | None = None
    selected: list = field(default_factory=list)       # [y6, y7, mid(+custom)] or [] (reference sentinel)
    front_u8: torch.Tensor | None = None               # [N, W] coronal image
    rib_boxes: torch.Tensor | None = None              # [k, 4] in coronal px
    n_det: torch.Tensor | None = None


class ImagingPipeline:
    """Owns the three networks (rib, axial 256, axial 512 -- ai_tools.py:52,66-67) and scratch memory."""

    def __init__(self, device="cuda:0", dtype=torch.float16, seed: int = 0, calibrate: bool = True,
                 mask_variant: int = 0, weights: dict | None = None, engine: str = "eitb"):
        """``weights``: {"ribs" | "axial256" | "axial512": checkpoint path} (ultralytics ``.pt`` or a state_dict
        in its key layout, see ``weights.py``).  A network whose file exists is loaded with its real BatchNorm
        statistics; the others are seeded random-init (the reference's files are not distributed) with the
        class bias calibrated so that the post-process sees candidates.  ``engine``: "eitb" runs the networks
        on libeitb200's own convolution kernels (K11/K12), "cudnn" through PyTorch/cuDNN + the K9 epilogue."""
        self.device = torch.device(device)
        self.dtype = dtype
        self.mask_variant = mask_variant
        self.engine = engine
        self.loaded = {}
        weights = weights or {}

        def make(key, nc, sd):
            path = weights.get(key)
            if path and os.path.exists(path):
                from .weights import load_model
                self.loaded[key] = path
                return load_model(path, self.device, dtype, nc)
            if path:
                logger.warning("weights for %s not found at %s: seeded random-init network instead", key, path)
            return build_model(nc, self.device, dtype, sd)

        self.ribs_torch = make("ribs", 1, seed)
        self.axial_512_torch = make("axial512", 4, seed + 1)
        self.axial_256_torch = make("axial256", 4, seed + 2)
        self._bind_engine()
        self.bias_shift = {}
        if calibrate:
            self._calibrate()

    def _net(self, model, x, out=None):
        """Run a network on a replicated gray image (every input of the service is one): lets the own-kernel
        engine read a single input channel in the stem.  ``out`` = (head, NHWC prototypes) buffers to write into."""
        if self.engine == "eitb" and self.dtype == torch.float16:
            return model(x, gray=True) if out is None else model(x, gray=True, out=out)
        head, protos = model(x)
        if out is not None:
            out[0].copy_(head)
            out[1].copy_(protos.permute(0, 2, 3, 1))
            head, protos = out[0], out[1].permute(0, 3, 1, 2)
        return head, protos

    @property
    def fused_input(self) -> bool:
        """Own-kernel engine: the stem takes the u8 window image (K1's u8 output) and applies the ultralytics
        preprocess itself, so the normalised [B,3,H,W] tensor is never written."""
        return self.engine == "eitb" and self.dtype == torch.float16

    def window_input(self, px, body, rot180: bool = True, out=None):
        """K1 for the axial networks: the u8 window image on the fused path, the normalised NCHW tensor otherwise
        (written into ``out`` when given)."""
        if self.fused_input:
            u8, _ = ops.hu_window(px, body_mask=body, want_u8=True, nchw_dtype=None, rot180=rot180, u8_out=out)
            return u8
        _, x = ops.hu_window(px, body_mask=body, want_u8=False, nchw_dtype=self.dtype, rot180=rot180, channels_last=True,
                             nchw_out=out)
        return x

    def _bind_engine(self):
        if self.engine == "eitb" and self.dtype == torch.float16:
            from .convnet import ConvNet
            self.ribs_model, self.axial_model_512, self.axial_model_256 = (
                ConvNet(self.ribs_torch), ConvNet(self.axial_512_torch), ConvNet(self.axial_256_torch))
        else:
            self.ribs_model, self.axial_model_512, self.axial_model_256 = self.ribs_torch, self.axial_512_torch, self.axial_256_torch

    # ------------------------------------------------------------------ random-init calibration
    @torch.no_grad()
    def _calibrate(self):
        """Random-init class heads score ~0: shift the class bias so ~1 % of the anchors
        of a phantom batch pass conf=0.3.  The shifts are reported by bench.py."""
        from . import synth
        if len(self.loaded) == 3:
            return
        px = torch.from_numpy(np.stack([synth.phantom_slice(s) for s in range(4)])).to(self.device)
        body = ops.body_mask(px, 1, -1024, True)
        _, x = ops.hu_window(px, body_mask=body, want_u8=False, nchw_dtype=self.dtype)
        x = x.contiguous(memory_format=torch.channels_last)
        if "axial512" not in self.loaded:
            self.bias_shift["axial512"] = self.axial_512_torch.shift_class_bias(x)
        x256 = torch.nn.functional.interpolate(x, size=(256, 256)).contiguous(memory_format=torch.channels_last)
        if "axial256" not in self.loaded:
            self.bias_shift["axial256"] = self.axial_256_torch.shift_class_bias(x256)
        if "ribs" in self.loaded:
            return
        vol, inst = synth.phantom_series(64, seed=0)
        front = self.coronal(torch.from_numpy(vol).to(self.device), SeriesMeta(inst))
        self.bias_shift["ribs"] = self.ribs_torch.shift_class_bias(self._rib_input(front[None])[0], frac=0.004)

    # ------------------------------------------------------------------ a2/a3: coronal image
    def coronal(self, px: torch.Tensor, meta: SeriesMeta, minmax: torch.Tensor | None = None,
                return_rows: bool = False):
        """[N,H,W] int16 in file order -> coronal image [N,W] u8 (one row per slice, MINMAX-normalised)."""
        n, H, W = px.shape
        order = torch.from_numpy(host.instance_order(meta.instance_numbers)).to(px.device, non_blocking=True)
        row, fx, fz = host.front_geometry(H, meta.patient_position, meta.image_orientation, meta.patient_orientation)
        rows, mm = ops.front_rows(px, order, n, row, fx, fz, minmax)
        if return_rows:
            return rows, mm
        return ops.minmax_u8(rows, mm)

    def _rib_input(self, front: torch.Tensor):
        """[S,N,W] u8 -> letterboxed network input (imgsz 640, auto) + the scale_boxes parameters."""
        S, N, W = front.shape
        nh, nw, top, bottom, left, right = host.letterbox_geometry(N, W, 640)
        x = ops.letterbox_nchw(front, nh, nw, top, left, nh + top + bottom, nw + left + right, self.dtype)
        x = x.contiguous(memory_format=torch.channels_last)
        gain, pad_x, pad_y = host.scale_boxes_params((nh + top + bottom, nw + left + right), (N, W))
        return x, (gain, pad_x, pad_y, W, N)

    # ------------------------------------------------------------------ a4/a5: rib detection + slice pick
    @torch.no_grad()
    def rib_select(self, front: torch.Tensor, custom: torch.Tensor | None = None):
        """[S,N,W] coronal images -> ([S,4] int32 (y6, y7, mid+custom, ok), boxes [S,300,4], k [S])."""
        x, (gain, pad_x, pad_y, w0, h0) = self._rib_input(front)
        head, _ = self._net(self.ribs_model, x)
        dets, _, k = ops.nms(head.contiguous(), 1, CONF, IOU, MAX_DET, want_idx=False)
        boxes = ops.scale_boxes(dets, k, gain, pad_x, pad_y, w0, h0)
        # image_width is hard-coded to 512 in the reference (utils.py:166)
        return ops.rib_select(boxes, k, 512.0, custom), boxes, k

    # ------------------------------------------------------------------ a6-a19: per-slice path
    @torch.no_grad()
    def segment(self, px: torch.Tensor, slope: int = 1, intercept: int = -1024, use_body: bool = True,
                rot180: bool = True):
        """[B,S,S] int16 stored pixels -> (labels [B,S,S] u8 codes, body [B,S,S] u8, n_det [B])."""
        B, H, W = px.shape
        body = ops.body_mask(px, slope, intercept, True) if use_body else None
        return self._segment_nchw(self.window_input(px, body, rot180), body)

    @torch.no_grad()
    def segment_u8(self, gray: torch.Tensor):
        """jpg_png route (ai_tools.py:365-400): [B,S,S] u8, no windowing, no body mask."""
        return self._segment_nchw(gray.contiguous() if self.fused_input else ops.u8_to_nchw(gray, self.dtype), None)

    @torch.no_grad()
    def segment_bgr_u8(self, bgr: torch.Tensor):
        """jpg_png route with a genuinely coloured upload: [B,S,S,3] u8 in OpenCV's BGR order.  The reference converts
        BGR -> RGB and hands all three channels to the model (ai_tools.py:134,153); the three-channel network input
        takes the generic 27-tap stem instead of the gray one."""
        x = (bgr.flip(-1).to(torch.float32) / 255).to(self.dtype).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        S = x.shape[-1]
        model = self.axial_model_256 if S == 256 else self.axial_model_512
        head, protos = model(x, gray=False) if self.engine == "eitb" and self.dtype == torch.float16 else model(x)
        dets, _, n = ops.nms(head.contiguous(), 4, CONF, IOU, MAX_DET, want_idx=False)
        code, _, _ = ops.mask_decode(dets, n, protos, self.mask_variant)
        ops.label_cleanup(code, None)
        return code, None, n

    @torch.no_grad()
    def predict_instances(self, x: torch.Tensor, drop_empty: bool = True):
        """What ``model(img, conf=0.3, imgsz=S)[0]`` holds in the reference (ai_tools.py:153), per image: the kept boxes
        (xyxy, conf, cls) and the per-instance binary masks [n, S, S] u8 -- K5 -> K6 with the per-instance bit masks
        kept.  ``x``: network input as for ``_segment_nchw``.  Returns a list of (dets [n, 6], masks [n, S, S])."""
        S = x.shape[-1]
        model = self.axial_model_256 if S == 256 else self.axial_model_512
        head, protos = self._net(model, x if x.dtype == torch.uint8 else x.contiguous(memory_format=torch.channels_last))
        dets, _, n = ops.nms(head.contiguous(), 4, CONF, IOU, MAX_DET, want_idx=False)
        _, area, bits = ops.mask_decode(dets, n, protos, self.mask_variant, want_area=True, want_bits=True)
        shifts = torch.arange(8, device=bits.device, dtype=torch.uint8)
        out = []
        for b in range(x.shape[0]):
            k = int(n[b])
            m = ((bits[b, :k, :, :, None] >> shifts) & 1).reshape(k, bits.shape[2], bits.shape[3] * 8)
            keep = area[b, :k] > 0 if drop_empty else torch.ones(k, dtype=torch.bool, device=bits.device)
            out.append((dets[b, :k, :6][keep], m[keep]))
        return out

    def _segment_nchw(self, x: torch.Tensor, body):
        """``x``: the normalised [B,3,S,S] network input, or (fused path) the u8 image [B,S,S] it is made from."""
        S = x.shape[-1]
        # get_axial_slice_size / model choice, ai_tools.py:138-146: 256 -> the 256 model, else the 512 one
        model = self.axial_model_256 if S == 256 else self.axial_model_512
        head, protos = self._net(model, x if x.dtype == torch.uint8 else x.contiguous(memory_format=torch.channels_last))
        dets, _, n = ops.nms(head.contiguous(), 4, CONF, IOU, MAX_DET, want_idx=False)
        code, _, _ = ops.mask_decode(dets, n, protos, self.mask_variant)
        ops.label_cleanup(code, body)
        return code, body, n

    # ------------------------------------------------------------------ whole series from host memory
    @torch.no_grad()
    def run_series(self, px_host: torch.Tensor, meta: SeriesMeta, labels_host: torch.Tensor | None = None,
                   chunk: int = 64, custom: int = 0, all_slices: bool = True) -> SeriesResult:
        """px_host [N,H,W] int16 (pinned) in file order.  ``all_slices`` pushes every slice through the
        per-slice path (throughput mode); otherwise only the three selected slices (service mode).
        ``labels_host`` (pinned, [N,H,W] u8, file order) receives the label maps."""
        dev = self.device
        N, H, W = px_host.shape
        res = SeriesResult()
        copy_in, copy_out = self._streams()
        main = torch.cuda.current_stream(dev)
        px = torch.empty((N, H, W), dtype=torch.int16, device=dev)
        labels = torch.empty((N, H, W), dtype=torch.uint8, device=dev)
        ready = []
        with torch.cuda.stream(copy_in):
            for c0 in range(0, N, chunk):
                px[c0:c0 + chunk].copy_(px_host[c0:c0 + chunk], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_in)
                ready.append(ev)
        ndet = torch.empty((N,), dtype=torch.int32, device=dev)
        done = []
        for ci, c0 in enumerate(range(0, N, chunk)):
            main.wait_event(ready[ci])
            if all_slices:
                code, _, n = self.segment(px[c0:c0 + chunk], meta.rescale_slope, meta.rescale_intercept)
                labels[c0:c0 + chunk] = code
                ndet[c0:c0 + chunk] = n
                if labels_host is not None:
                    ev = torch.cuda.Event()
                    ev.record(main)
                    copy_out.wait_event(ev)
                    with torch.cuda.stream(copy_out):
                        labels_host[c0:c0 + chunk].copy_(labels[c0:c0 + chunk], non_blocking=True)
        # coronal scan needs every slice resident
        front = self.coronal(px, meta)
        cus = torch.tensor([custom], dtype=torch.int32, device=dev)
        sel, boxes, k = self.rib_select(front[None], cus)
        res.front_u8, res.rib_boxes, res.n_det = front, boxes[0], ndet
        sel_host = sel.cpu()[0].tolist()                # the one host sync of the series
        res.selected = sel_host[:3] if sel_host[3] else []
        if not all_slices and res.selected and not all(-N <= i < N for i in res.selected):
            # the reference indexes the sorted slice list with these numbers (ai_tools.py:177-178): a number past the end
            # raises there and the request ends with the failure sentinel; a negative one wraps like a Python index
            logger.error(f"selected slice numbers {res.selected} outside a series of {N} slices")
            res.selected = []
        if not all_slices and res.selected:
            order = host.instance_order(meta.instance_numbers)
            idx = [int(order[i % N]) for i in res.selected]
            code, body, n = self.segment(px[idx].contiguous(), meta.rescale_slope, meta.rescale_intercept)
            labels, res.body = code, body
            if labels_host is not None:
                labels_host[:len(idx)].copy_(code)
        main.wait_stream(copy_out)
        res.labels = labels
        return res

    def _streams(self):
        if not hasattr(self, "_copy_in"):
            self._copy_in = torch.cuda.Stream(self.device)
            self._copy_out = torch.cuda.Stream(self.device)
        return self._copy_in, self._copy_out

    # ------------------------------------------------------------------ a22-a24: mesh labelling
    def label_mesh(self, nodes_xy: np.ndarray, triangles: np.ndarray, polygon_strings, outer_class: int = 4):
        """Per-element classes for a triangle mesh (the CLASS vector of export_mesh_for_femm)."""
        strs = list(polygon_strings)
        contours = host.parse_contours(strs, host.find_outer_index(strs))
        xy, off, cls = host.prepare_polygons(contours)
        dev = self.device
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        return ops.tri_label(t(np.asarray(nodes_xy, np.float64)), t(np.asarray(triangles, np.int64)), t(xy), t(off),
                             t(cls), outer_class)


class _NullTimer:
    def __call__(self, name):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class SeriesBatchRunner:
    """Throughput engine for a batch of series, slice-sharded over the ranks of ``torch.distributed``
    (SURVEY §8(e)): this rank owns a contiguous z-range of every series.

    * ``step_device()``  -- one pass with the pixels already in HBM;
    * ``step_host(px_host, labels_host)`` -- the same pass from pinned host memory: the 1 KiB coronal
      rows travel first (the per-series decision runs while the first chunk of slices is on the
      wire), slices go host->device and label maps device->host chunk by chunk on two copy streams.

    The per-slice path of every chunk (K2, K1, CNN, K5, K6, K7; ~1,000 launches) and the per-series
    decision (K3 normalise, letterbox, rib CNN, K5, K4) are captured once as CUDA graphs over fixed
    buffers and replayed; the NCCL exchange between them stays eager.
    """

    def __init__(self, pipe: ImagingPipeline, metas, n_slices: int, size: int = 512, chunk: int = 160,
                 use_graphs: bool = True, timer=None, first_chunk: int = 0, chunk_sizes=None, overlap: bool = True, label_fan: int = 1):
        from . import sharded
        self.pipe, self.sharded = pipe, sharded
        self.dev = pipe.device
        self.rank, self.world = sharded.world()
        self.S, self.n_slices, self.size, self.chunk = len(metas), n_slices, size, chunk
        z0, z1 = sharded.shard_range(n_slices, self.world, self.rank)
        self.nl = z1 - z0
        self.metas = metas
        dev = self.dev
        self.orders = torch.from_numpy(np.stack([host.instance_order(m.instance_numbers) for m in metas])).to(dev)   # [S, nl]
        # per-series geometry and rescale tags (a batch may mix orientations and scanners)
        geo = [host.front_geometry(size, m.patient_position, m.image_orientation, m.patient_orientation) for m in metas]
        # z is reversed AFTER the gather (a per-shard reversal would come out as [rev(shard0), rev(shard1), ...])
        self.flip_z = [s for s, g in enumerate(geo) if g[2]]
        self.geom = torch.tensor([[g[0], int(g[1]), 0] for g in geo], dtype=torch.int32, device=dev)
        self.geom_row0 = self.geom.clone()
        self.geom_row0[:, 0] = 0                                # host path ships only the coronal row of every slice
        self.rows_of = [g[0] for g in geo]
        self.rescale = [(m.rescale_slope, m.rescale_intercept) for m in metas]
        self.uniform_rescale = len(set(self.rescale)) == 1
        self.timer = timer or _NullTimer()
        self.px = torch.empty((self.S, self.nl, size, size), dtype=torch.int16, device=dev)      # the resident batch
        self.flat = self.px.view(self.S * self.nl, size, size)
        total = self.S * self.nl
        # (start, stop) of every chunk; a smaller first chunk lets compute start sooner on the host path
        starts = list(range(0, total, chunk)) if not first_chunk else [0] + list(range(first_chunk, total, chunk))
        self.bounds = [(a, min(b, total)) for a, b in zip(starts, starts[1:] + [total])]
        if chunk_sizes:                                         # explicit list, e.g. small head and tail chunks for the host path
            assert sum(chunk_sizes) == total, "chunk_sizes must add up to the local slice count"
            edges = [0]
            for c in chunk_sizes:
                edges.append(edges[-1] + c)
            self.bounds = list(zip(edges[:-1], edges[1:]))
        self.chunks = [a for a, _ in self.bounds]
        self.mine = [s for s in range(self.S) if sharded.owner_of_series(s, self.world) == self.rank]
        self.mine_idx = torch.tensor(self.mine, dtype=torch.int64, device=dev)
        self.rows_static = torch.zeros((self.S, n_slices, size), dtype=torch.int16, device=dev)
        self.mm_static = torch.zeros((self.S, 2), dtype=torch.int32, device=dev)
        self.rows_dev = torch.empty((self.S, self.nl, 1, size), dtype=torch.int16, device=dev)
        self.copy_in, self.copy_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.side = torch.cuda.Stream(dev)                      # the per-series decision runs beside the slice chunks
        self.graphs, self._outs, self.rib_graph, self.sel_static = [], [], None, None
        self.use_graphs = use_graphs
        # Overlapped replay (graphs only): a chunk is three graphs -- K2 on ``pre_s``, K1 + CNN on the caller's stream,
        # K5/K6/K7 on ``post_s`` -- so the latency-bound label kernels of chunk c (a handful of busy warps per image)
        # run in the tails and launch gaps of chunk c+1's persistent convolution kernels instead of after them, also
        # across steps.  Buffers are guarded by events (previous reader -> next writer), never by stream joins.
        self.overlap = bool(overlap and use_graphs)
        self.pre_s, self.post_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        # The network's graphs are captured on and replayed in a HIGH-priority stream: when one of its persistent
        # convolution kernels starts, its CTAs are placed before pending CTAs of the label kernels and of the rib network
        # (which then fill the tails), instead of waiting behind them with a static tile split
        self.cnn_prio = os.environ.get("EITB_CNN_PRIORITY", "1") == "1" and self.overlap
        self.cnn_s = torch.cuda.Stream(dev, priority=-1) if self.cnn_prio else None
        # K2 and K7 are chains of latency-bound kernels (one CTA per image, a few busy warps): inside their graphs a chunk is
        # cut into ``label_fan`` groups of images whose chains run on parallel branches, so that several stages are
        # resident at once (and the instruction-bound K6 of one group runs beside the K7 of another)
        self.label_fan = max(1, int(label_fan)) if self.overlap else 1
        self.fan_s = [torch.cuda.Stream(dev) for _ in range(self.label_fan - 1)]
        # two sets of hand-over buffers and graphs, used by alternate passes: the K2 / K1 graph of pass k+1 does not have
        # to wait for the label kernels of pass k (which still read pass k's body mask), so with one chunk per pass the
        # three stages of consecutive passes still run side by side
        self.n_sets = 2 if self.overlap else 1
        self.pre_graphs, self.cnn_graphs, self.post_graphs = [[], []], [[], []], [[], []]
        nch = len(self.bounds)
        self.ev_px = [None] * nch                                 # K2 / K1 have read the chunk's pixels (guards the next copy-in)
        self.ev_post = [[None] * nch for _ in range(2)]           # per set: the label kernels have read body / head / prototypes
        self.ev_d2h = [[None] * nch for _ in range(2)]            # per set: the copy-out has read the label image
        self._passes, self._last_set = 0, 0
        self._sel_ring, self._submitted = [], 0

    # ---------------------------------------------------------------- stages
    def _fan(self, n: int, fn, enabled: bool):
        """``fn(lo, hi)`` over ``label_fan`` contiguous groups of ``n`` images, group 0 on the current stream and the others
        on the fan streams (forked from and joined to it, so a stream capture records parallel branches)."""
        k = min(self.label_fan, n) if enabled else 1
        if k <= 1:
            fn(0, n)
            return
        cur = torch.cuda.current_stream(self.dev)
        edges = [n * i // k for i in range(k + 1)]
        for i in range(1, k):
            st = self.fan_s[i - 1]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                fn(edges[i], edges[i + 1])
        fn(edges[0], edges[1])
        for i in range(1, k):
            cur.wait_stream(self.fan_s[i - 1])

    @torch.no_grad()
    def pre_stage(self, px_chunk, a: int = 0, b: int | None = None, out=None):
        """K2 + K1 for the slices [a, b) of the flattened [S * n_local] batch: (body mask, network input) -- the two stages
        that read the pixels.  ``out`` (here and in the next two stages): the hand-over buffers to write the results into."""
        with self.timer("K2_body_mask"):
            if self.uniform_rescale and out is not None:
                body = out[0]
                self._fan(px_chunk.shape[0], lambda lo, hi: ops.body_mask(px_chunk[lo:hi], self.rescale[0][0], self.rescale[0][1],
                                                                           True, out=body[lo:hi]), True)
            elif self.uniform_rescale:
                body = ops.body_mask(px_chunk, self.rescale[0][0], self.rescale[0][1], True)
            else:                                               # per-series RescaleSlope / RescaleIntercept
                body = torch.empty(px_chunk.shape, dtype=torch.uint8, device=self.dev) if out is None else out[0]
                b = a + px_chunk.shape[0] if b is None else b
                s0 = a // self.nl
                while s0 * self.nl < b:
                    lo, hi = max(a, s0 * self.nl) - a, min(b, (s0 + 1) * self.nl) - a
                    body[lo:hi] = ops.body_mask(px_chunk[lo:hi], self.rescale[s0][0], self.rescale[s0][1], True)
                    s0 += 1
        with self.timer("K1_hu_window_nchw"):
            x = self.pipe.window_input(px_chunk, body, out=None if out is None else out[1])
        return body, x

    @torch.no_grad()
    def cnn_stage(self, x, out=None):
        """The axial network on K1's output."""
        pipe = self.pipe
        with self.timer("CNN_axial"):
            head, protos = pipe._net(pipe.axial_model_256 if self.size == 256 else pipe.axial_model_512, x, out=out)
            head = head.contiguous()
        return head, protos

    @torch.no_grad()
    def post_stage(self, head, protos, body, out=None):
        """K5 -> K6 -> K7."""
        t, pipe = self.timer, self.pipe
        if out is not None:                                       # graph capture: groups of images on parallel branches

            def group(lo, hi):
                dets, _, n = ops.nms(head[lo:hi], 4, CONF, IOU, MAX_DET, want_idx=False)
                out[1][lo:hi].copy_(n)
                ops.mask_decode(dets, n, protos[lo:hi], pipe.mask_variant, code_out=out[0][lo:hi])
                ops.label_cleanup(out[0][lo:hi], body[lo:hi])
            self._fan(head.shape[0], group, True)
            return out[0], out[1]
        with t("K5_nms"):
            dets, _, n = ops.nms(head, 4, CONF, IOU, MAX_DET, want_idx=False)
        with t("K6_mask_decode"):
            code, _, _ = ops.mask_decode(dets, n, protos, pipe.mask_variant)
        with t("K7_label_cleanup"):
            ops.label_cleanup(code, body)
        return code, n

    def slice_stage(self, px_chunk, a: int = 0, b: int | None = None):
        """K2 .. K7 for the slices [a, b) of the flattened [S * n_local] batch."""
        body, x = self.pre_stage(px_chunk, a, b)
        head, protos = self.cnn_stage(x)
        return self.post_stage(head, protos, body)

    def rib_rows(self, px, row0: bool = False):
        """One launch for the whole batch: the coronal row of every local slice of every series + per-series
        (min, max).  ``row0``: ``px`` holds only that row of each slice ([S, n_local, 1, W], host path)."""
        with self.timer("K3_front_rows"):
            rows, mm = ops.front_rows_batch(px, self.orders, self.geom_row0 if row0 else self.geom)
        return rows, mm

    @torch.no_grad()
    def rib_decide(self, rows_all, mm_all):
        t, pipe = self.timer, self.pipe
        sel = torch.zeros((self.S, 4), dtype=torch.int32, device=self.dev)
        if self.mine:
            with t("K3_minmax_letterbox"):
                front = torch.stack([ops.minmax_u8(rows_all[s], mm_all[s]) for s in self.mine])
                x, (gain, pad_x, pad_y, w0, h0) = pipe._rib_input(front)
            with t("CNN_ribs"):
                head, _ = pipe._net(pipe.ribs_model, x)
                head = head.contiguous()
            with t("K5_nms_ribs"):
                dets, _, k = ops.nms(head, 1, CONF, IOU, MAX_DET, want_idx=False)
            with t("K4_rib_select"):
                boxes = ops.scale_boxes(dets, k, gain, pad_x, pad_y, w0, h0)
                sel.index_copy_(0, self.mine_idx, ops.rib_select(boxes, k, 512.0))
        return sel

    def rib_stage(self, px, graphed=False, row0=False):
        rows, mm = self.rib_rows(px, row0)
        with self.timer("C1_exchange"):
            rows_all, mm_all = self.sharded.gather_rows(rows, mm, self.n_slices, self.flip_z)
        if graphed and self.rib_graph is not None:
            self.rows_static.copy_(rows_all)
            self.mm_static.copy_(mm_all)
            self.rib_graph.replay()
            sel = self.sel_static.clone()
        else:
            sel = self.rib_decide(rows_all, mm_all)
        with self.timer("C1_exchange"):
            sel = self.sharded.share_selected(sel)
        return sel

    # ---------------------------------------------------------------- steps
    def load(self, px_host: torch.Tensor):
        """Place a batch [S, n_local, H, W] in the resident buffer (outside any timed region)."""
        self.px.copy_(px_host)
        torch.cuda.synchronize(self.dev)

    def step_eager(self):
        sel = self.rib_stage(self.px)
        for a, b in self.bounds:
            self.slice_stage(self.flat[a:b], a, b)
        return sel

    def capture(self, warm: int = 2):
        """cuDNN autotune / lazy loading eagerly, then the CUDA graphs of every chunk and one for the rib decision."""
        for _ in range(warm):
            self.step_eager()
        torch.cuda.synchronize(self.dev)
        if not self.use_graphs:
            return
        if self.overlap:
            # Hand-over buffers (body mask, head, prototypes, label image, counts) live OUTSIDE the graph pools: inside a
            # shared pool the outputs of chunk 1's graph may sit where chunk 0's graph keeps its intermediates, which is
            # only safe while nothing reads them beside a later replay of chunk 0 -- exactly what the overlap does.
            shapes = []
            with torch.no_grad():
                for a, b in self.bounds:
                    body, x = self.pre_stage(self.flat[a:b], a, b)
                    head, protos = self.cnn_stage(x)
                    code, n = self.post_stage(head, protos, body)
                    shapes.append((body, x, head, protos, code, n))
            hand = [[dict(body=torch.empty_like(body), x=torch.empty_like(x), head=torch.empty_like(head),
                          protos=torch.empty((protos.shape[0], protos.shape[2], protos.shape[3], protos.shape[1]),
                                             dtype=protos.dtype, device=self.dev),      # NHWC
                          code=torch.empty_like(code), n=torch.empty_like(n))
                     for body, x, head, protos, code, n in shapes] for _ in range(self.n_sets)]
            del shapes, body, x, head, protos, code, n
            torch.cuda.synchronize(self.dev)
            # three memory pools: graphs of one kind replay in order on one stream, graphs of different kinds side by side
            pools = [torch.cuda.graph_pool_handle() for _ in range(3)]
            self._outs = [[], []]
            for st in range(self.n_sets):
                for ci, (a, b) in enumerate(self.bounds):
                    h = hand[st][ci]
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=pools[0]):
                        _, x = self.pre_stage(self.flat[a:b], a, b, out=(h["body"], h["x"]))
                    assert x.data_ptr() == h["x"].data_ptr()
                    self.pre_graphs[st].append(g)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=pools[1], **({"stream": self.cnn_s} if self.cnn_prio else {})):
                        head, protos = self.cnn_stage(h["x"], out=(h["head"], h["protos"]))
                    assert head.data_ptr() == h["head"].data_ptr() and protos.data_ptr() == h["protos"].data_ptr()
                    self.cnn_graphs[st].append(g)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=pools[2]):
                        code, n = self.post_stage(h["head"], h["protos"].permute(0, 3, 1, 2), h["body"], out=(h["code"], h["n"]))
                    assert code.data_ptr() == h["code"].data_ptr()
                    self.post_graphs[st].append(g)
                    self._outs[st].append((h["code"], h["n"]))
            self._static = hand
        else:
            pool = torch.cuda.graph_pool_handle()
            for a, b in self.bounds:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    o = self.slice_stage(self.flat[a:b], a, b)
                self.graphs.append(g)
                self._outs.append(o)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):                                 # own memory pool: it replays beside the chunk graphs
            self.sel_static = self.rib_decide(self.rows_static, self.mm_static)
        self.rib_graph = g
        torch.cuda.synchronize(self.dev)

    @property
    def outs(self):
        """(label image, detection count) of every chunk of the latest pass (after ``join()`` in overlapped mode)."""
        return self._outs[self._last_set] if self.overlap and self._outs else self._outs

    def run_chunk(self, ci):
        if self.graphs:
            self.graphs[ci].replay()
            return self._outs[ci]
        a, b = self.bounds[ci]
        return self.slice_stage(self.flat[a:b], a, b)

    def _event(self, stream):
        e = torch.cuda.Event()
        e.record(stream)
        return e

    def _begin_pass(self):
        self._last_set = self._passes % self.n_sets
        self._passes += 1
        return self._last_set

    def _chunk_overlapped(self, st, ci, main, ready=None):
        """pre (K2, K1) -> cnn (network) -> post (K5, K6, K7) of chunk ``ci`` with buffer set ``st`` on three streams.
        Waits: the chunk's pixels (``ready``), and whoever still reads a buffer this replay overwrites -- the post graph of
        the pass that last used the set reads body / head / prototypes, its device->host copy reads the label image."""
        with torch.cuda.stream(self.pre_s):
            if ready is not None:
                self.pre_s.wait_event(ready)
            if self.ev_post[st][ci] is not None:
                self.pre_s.wait_event(self.ev_post[st][ci])
            self.pre_graphs[st][ci].replay()
            ev_pre = self._event(self.pre_s)
        self.ev_px[ci] = ev_pre
        cnn_s = self.cnn_s if self.cnn_prio else main
        with torch.cuda.stream(cnn_s):
            cnn_s.wait_event(ev_pre)
            self.cnn_graphs[st][ci].replay()
            ev_cnn = self._event(cnn_s)
        with torch.cuda.stream(self.post_s):
            self.post_s.wait_event(ev_cnn)
            if self.ev_d2h[st][ci] is not None:
                self.post_s.wait_event(self.ev_d2h[st][ci])
            self.post_graphs[st][ci].replay()
            self.ev_post[st][ci] = self._event(self.post_s)
        return self._outs[st][ci]

    def join(self):
        """Make the caller's stream wait for everything the runner has in flight on its own streams."""
        main = torch.cuda.current_stream(self.dev)
        for st in (self.pre_s, self.post_s, self.copy_out, self.side) + ((self.cnn_s,) if self.cnn_prio else ()):
            main.wait_stream(st)

    def _rib_on_side(self, px, row0=False, after=None, free_running=False):
        """The coronal decision is independent of the per-slice path and tiny (one image per series): run it
        on a side stream so its ~300 small launches hide under the chunk graphs.  ``free_running``: the side stream
        neither waits for the caller's stream nor is waited for (resident pixels; the caller joins later)."""
        main = torch.cuda.current_stream(self.dev)
        if after is not None:
            self.side.wait_event(after)                           # host path: the coronal rows have landed
        elif not free_running:
            self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            sel = self.rib_stage(px, graphed=True, row0=row0)
        sel.record_stream(main)
        return sel

    def step_device(self, join: bool = True):
        """One pass over the resident batch (``load`` is its only writer and synchronises).  ``join=False`` leaves the
        per-series decision and the label kernels of the last chunk running beside the next pass; call ``join()``
        before reading ``outs`` or the returned table."""
        main = torch.cuda.current_stream(self.dev)
        free = self.overlap and not join
        if os.environ.get("EITB_EXPERIMENT_NO_RIB") == "1":        # measurement only: what the per-series decision costs the pass
            sel = torch.zeros((self.S, 4), dtype=torch.int32, device=self.dev)
        else:
            sel = self._rib_on_side(self.px, free_running=free)
        st = self._begin_pass()
        for ci in range(len(self.chunks)):
            if self.overlap:
                self._chunk_overlapped(st, ci, main)
            else:
                self.run_chunk(ci)
        if not free:
            main.wait_stream(self.side)
        if join and self.overlap:
            self.join()
        return sel

    def submit_host(self, px_host: torch.Tensor, labels_host: torch.Tensor):
        """Enqueue one pass from pinned host memory and return a handle for ``wait_host``; nothing blocks the host, so
        a caller that keeps two passes in flight (two ``labels_host`` buffers) overlaps the copies of one pass with
        the kernels of the other.  px_host [S, n_local, H, W] int16 pinned, labels_host [S, n_local, H, W] u8 pinned
        (file order).  At most eight passes may be between ``submit_host`` and ``wait_host`` (the selected-slice tables
        are eight pinned buffers used in turn).  ``wait_host`` (or ``join()`` + a device synchronise) before switching to
        ``step_device``: the resident-pixel path does not wait for copies that are still in flight."""
        main = torch.cuda.current_stream(self.dev)
        flat_host = px_host.view(self.S * self.nl, self.size, self.size)
        flat_out = labels_host.view(self.S * self.nl, self.size, self.size)
        if self.overlap:
            self.copy_in.wait_stream(self.side)                   # the previous pass's rib stage has read rows_dev
        else:
            self.copy_in.wait_stream(main)
        evs = []
        with torch.cuda.stream(self.copy_in):
            # strided DMA of the coronal rows straight from the pinned series: no CPU gather, no staging buffer
            if len(set(self.rows_of)) == 1:
                ops.rows_h2d(flat_host, self.rows_of[0], self.rows_dev.view(self.S * self.nl, self.size))
            else:
                for s_ in range(self.S):
                    ops.rows_h2d(px_host[s_], self.rows_of[s_], self.rows_dev[s_].view(self.nl, self.size))
            ev_rows = self._event(self.copy_in)
            for ci, (a, b) in enumerate(self.bounds):
                if self.overlap and self.ev_px[ci] is not None:
                    self.copy_in.wait_event(self.ev_px[ci])       # K2 / K1 of the previous pass have read these pixels
                self.flat[a:b].copy_(flat_host[a:b], non_blocking=True)
                evs.append(self._event(self.copy_in))
        if self.overlap:
            sel = self._rib_on_side(self.rows_dev, row0=True, after=ev_rows)
        else:
            main.wait_event(ev_rows)
            sel = self._rib_on_side(self.rows_dev, row0=True)
        st = self._begin_pass()
        for ci, (a, b) in enumerate(self.bounds):
            if self.overlap:
                code, _ = self._chunk_overlapped(st, ci, main, evs[ci])
                self.copy_out.wait_event(self.ev_post[st][ci])
            else:
                main.wait_event(evs[ci])
                code, _ = self.run_chunk(ci)
                self.copy_out.wait_event(self._event(main))
            with torch.cuda.stream(self.copy_out):
                flat_out[a:b].copy_(code, non_blocking=True)
                self.ev_d2h[st][ci] = self._event(self.copy_out)
            code.record_stream(self.copy_out)
        # the selected-slice table travels last on the copy-out stream: its event closes the pass
        # one of eight pinned tables in turn: a caller may have up to eight passes between submit_host and wait_host
        if not self._sel_ring:
            self._sel_ring = [torch.empty((self.S, 4), dtype=torch.int32, pin_memory=True) for _ in range(8)]
        sel_host = self._sel_ring[self._submitted % len(self._sel_ring)]
        self._submitted += 1
        self.copy_out.wait_stream(self.side)
        with torch.cuda.stream(self.copy_out):
            sel_host.copy_(sel, non_blocking=True)
            done = self._event(self.copy_out)
        sel.record_stream(self.copy_out)
        if not self.overlap:
            main.wait_stream(self.copy_out)
            main.wait_stream(self.side)
        return sel_host, done

    @staticmethod
    def wait_host(handle):
        """Block until the pass behind ``handle`` has delivered its label maps; returns the selected-slice table [S, 4]."""
        sel_host, done = handle
        done.synchronize()
        return sel_host.clone()

    def step_host(self, px_host: torch.Tensor, labels_host: torch.Tensor):
        """One pass from pinned host memory, synchronously: returns the selected-slice table [S, 4] on the host once
        the label maps are in ``labels_host``."""
        return self.wait_host(self.submit_host(px_host, labels_host))
