/*
 * eitb200.h -- C ABI of libeitb200.so: the B200 (sm_100a) kernels behind the kt_service
 * imaging hot path of AndreyKatsupeev/EITSynthAI.
 *
 * The reference is pure Python and has no FFI; its seam for this path is the set of
 * Python functions that kt_service/ai_tools/ai_tools.py imports from utils.py
 * (ai_tools.py:12-15), the ultralytics post-process reached through model(...)
 * (ai_tools.py:121-122,153) and the triangle labeller of
 * mesh_tools/femm_generator.py:12-184.  Every entry point below names the reference
 * interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain C, no C++/torch types; all tensors are row-major, batch dimension first;
 *   - unless an argument is documented as "host", every pointer is a DEVICE pointer
 *     owned by the caller; the library never allocates, frees or keeps pointers;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); nothing syncs;
 *   - return value: EITB_OK (0) or a negative EITB_ERR_* code; no exceptions, no exit();
 *   - scratch memory is passed in (`ws`, `ws_bytes`); sizes come from *_workspace_bytes().
 */
#ifndef EITB200_H
#define EITB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EITB_OK 0
#define EITB_ERR_BAD_ARG (-1)      /* null pointer, non-positive size, unsupported shape */
#define EITB_ERR_WORKSPACE (-2)    /* ws_bytes smaller than *_workspace_bytes()           */
#define EITB_ERR_LAUNCH (-3)       /* cudaGetLastError() != cudaSuccess after a launch    */
#define EITB_ERR_UNSUPPORTED (-4)  /* valid request this build does not implement         */

/* element types for "void*" image / tensor arguments */
#define EITB_F32 0
#define EITB_F16 1
#define EITB_BF16 2

/* label-image codes: code = B<<2 | G<<1 | R of the reference's BGR colours
 * (utils.py:468-473): 0 black, 1 muscle (0,0,255), 3 adipose (0,255,255),
 * 6 lung (255,255,0), 7 bone (255,255,255). */
#define EITB_CODE_BLACK 0
#define EITB_CODE_MUSCLE 1
#define EITB_CODE_ADIPOSE 3
#define EITB_CODE_LUNG 6
#define EITB_CODE_BONE 7

typedef void* eitb_stream_t;

const char* eitb_strerror(int code);
int eitb_version(void);

/* Launch profiler (the reference only wall-clocks the segmentation call, ai_tools.py:152-155).
 * enable(1) clears the records and starts bracketing every kernel launch of the library with CUDA
 * events on its stream; enable(0) stops.  report() synchronises on the recorded events and writes
 * one "kernel_name launches total_ms" line per kernel into buf (NUL-terminated, truncated to cap);
 * returns the size needed.  Host pointers; do not enable while capturing a CUDA graph. */
int eitb_profile_enable(int on);
long long eitb_profile_report(char* buf, size_t cap);

/* ---- K1: HU window + normalise + rot180 + body-mask AND + NCHW ------------------------
 * Replaces classic_norm (utils.py:272-313), cv2.bitwise_and(norm, norm, mask=body)
 * (ai_tools.py:212-213,287-288,339-340,433-434) and, for the square 256/512 inputs the
 * service feeds, the ultralytics preprocess (gray -> 3 equal channels, HWC->NCHW,
 * /255).
 *   px        [B,H,W] int16 stored pixel values
 *   lo, hi    window bounds (reference: level 40, width 400 -> -160, 240)
 *   rot180    1: rotate each slice by 180 degrees (utils.py:309)
 *   body_mask [B,H,W] u8 or NULL; applied in OUTPUT orientation (nonzero keeps)
 *   out_u8    [B,H,W] u8 or NULL      -- classic_norm (+mask) result
 *   out_nchw  [B,3,H,W] of out_dtype or NULL -- u8/255 rounded once to out_dtype
 *   channels_last  0: NCHW memory order; 1: the same logical tensor stored NHWC ([B,H,W,3]),
 *             the layout a channels-last cuDNN network reads (fp16 / bf16 outputs only)
 * W must be a multiple of 8. */
int eitb_hu_window_nchw(const int16_t* px, int B, int H, int W, int lo, int hi, int rot180,
                        const uint8_t* body_mask, uint8_t* out_u8, void* out_nchw,
                        int out_dtype, int channels_last, eitb_stream_t stream);

/* u8 gray image(s) -> 3-channel NCHW /255 (the jpg_png route, ai_tools.py:365-400). */
int eitb_u8_to_nchw(const uint8_t* gray, int B, int H, int W, void* out_nchw, int out_dtype,
                    eitb_stream_t stream);

/* ---- K2: body mask ----------------------------------------------------------------------
 * Replaces get_axial_slice_body_mask (utils.py:526-585) and, with flipud=0, slope=1,
 * intercept=0, get_axial_slice_body_mask_nii (utils.py:588-618): HU = slope*px+intercept
 * wrapped to int16, -500 < HU < 1000, 5x5 opening, the external contour with the largest
 * cv2.contourArea filled with 255 (ties: the contour starting last in raster order).
 *   mask [B,H,W] u8 out, values {0,255} (all 0 when nothing survives the opening). */
size_t eitb_body_mask_workspace_bytes(int B, int H, int W);
int eitb_body_mask(const int16_t* px, int B, int H, int W, int slope, int intercept, int flipud,
                   uint8_t* mask, void* ws, size_t ws_bytes, eitb_stream_t stream);

/* Connected-component labelling, the building block of K2 and K7 (what scipy.ndimage.label /
 * cv2.findContours supply to utils.py:572, 721, 792).  mask [B,H,W] u8 (nonzero = in the set);
 * connectivity 4 or 8; labels [B,H,W] int32 out: the smallest pixel index (y*W+x) of the pixel's
 * component, -2 for pixels outside the set, and -1 for components touching the image frame when
 * link_outside != 0 (flood fill from the frame). */
int eitb_cc_label(const uint8_t* mask, int B, int H, int W, int connectivity, int link_outside,
                  int32_t* labels, eitb_stream_t stream);

/* ---- K3: coronal mid-row gather + min/max, MINMAX normalise ------------------------------
 * Replaces convert_to_3d + axial_to_sagittal + mid-plane (utils.py:73-163,
 * ai_tools.py:98-99) without materialising the (H,W,N) volume: one row per slice.
 *   px     [n_total,H,W] slices in file order
 *   order  [n] int32 indices into px giving InstanceNumber order (NULL: identity)
 *   row    source row (H/2, or H-1-H/2 when ImageOrientationPatient[4] == -1)
 *   flip_x reverse each row; flip_z reverse the slice order (see host mirror for the rules)
 *   rows   [n,W] int16 out
 *   minmax [2] int32 in/out: atomically min/max-combined, caller initialises to
 *          {INT32_MAX, INT32_MIN} (lets several shards or calls accumulate). */
int eitb_front_rows(const int16_t* px, const int32_t* order, int n, int H, int W, int row,
                    int flip_x, int flip_z, int16_t* rows, int32_t* minmax, eitb_stream_t stream);

/* The same for S series in one launch (a batch of series, BASELINE configs[3]): series s starts at
 * px + s*series_stride elements and has its own order row [S,n], geometry geom [S,3] int32 device
 * (row, flip_x, flip_z), output rows [S,n,W] and minmax [S,2]. */
int eitb_front_rows_batch(const int16_t* px, long long series_stride, const int32_t* order, const int32_t* geom,
                          int S, int n, int H, int W, int16_t* rows, int32_t* minmax, eitb_stream_t stream);

/* Host path: row `row` of each of n_slices [H,W] int16 slices in PINNED HOST memory -> dev_rows [n_slices,W],
 * one strided DMA on `stream` (host_px is a host pointer). */
int eitb_rows_h2d(const int16_t* host_px, long long n_slices, int H, int W, int row, int16_t* dev_rows,
                  eitb_stream_t stream);

/* cv2.normalize(front, None, 0, 255, NORM_MINMAX, CV_8U) (ai_tools.py:101) with OpenCV's
 * arithmetic: scale/shift in double, cast to float, one float FMA per pixel, round-half-even.
 *   minmax [2] int32 device pointer (from eitb_front_rows / an all-reduce). */
int eitb_minmax_u8(const int16_t* rows, int64_t count, const int32_t* minmax, uint8_t* out,
                   eitb_stream_t stream);

/* ---- a11: letterbox for non-square inputs (rib model on the (N,512) coronal image) -------
 * Restates ultralytics LetterBox(auto=True)+preprocess (call site
 * ai_tools.py:120-122): cv2.resize(INTER_LINEAR) fixed-point bilinear to (nh,nw), constant
 * 114 border, 3 equal channels, /255.  Geometry is computed by the host mirror.
 *   gray [B,H,W] u8 -> out [B,3,outH,outW]. */
int eitb_letterbox_nchw(const uint8_t* gray, int B, int H, int W, int nh, int nw, int top,
                        int left, int outH, int outW, void* out, int out_dtype,
                        eitb_stream_t stream);

/* ---- K4: rib arg-select -------------------------------------------------------------------
 * Replaces search_number_axial_slice (utils.py:166-269) for S series at once.
 *   xyxy   [S,max_k,4] float32 boxes in coronal-image pixels, k[S] valid counts
 *   out    [S,4] int32: int(y1[5]), int(y1[6]), int(|y1[5]+y1[6]|/2)+custom, ok(1/0)
 *          (ok=0 <=> fewer than 7 boxes right of image_width/2: the reference returns []). */
int eitb_rib_select(const float* xyxy, const int32_t* k, int S, int max_k, float image_width,
                    const int32_t* custom /* [S] or NULL */, int32_t* out, eitb_stream_t stream);

/* ---- K5: confidence filter + class-aware NMS ------------------------------------------------
 * Restates ultralytics non_max_suppression + torchvision.ops.nms (SURVEY Appendix A.3;
 * call sites ai_tools.py:121-122,153: conf=0.3, iou=0.7, max_det=300, max_wh=7680).
 *   head     [B,4+nc+nm,A] (xywh, class scores, mask coefficients) of head_dtype
 *   dets     [B,max_det,6+nm] float32 out: x1,y1,x2,y2,conf,cls,coef.. in descending score
 *   keep_idx [B,max_det] int32 out or NULL: anchor index of every kept detection
 *   n_out    [B] int32 out */
size_t eitb_nms_workspace_bytes(int B, int A);
int eitb_nms(const void* head, int head_dtype, int B, int nc, int nm, int A, float conf,
             float iou, int max_det, float max_wh, float* dets, int32_t* keep_idx,
             int32_t* n_out, void* ws, size_t ws_bytes, eitb_stream_t stream);

/* ultralytics scale_boxes, the step between the NMS output and
 * sv.Detections.xyxy (ai_tools.py:123): x = clamp((x - pad) / gain, 0, orig) in float32.
 *   dets [B,max_det,row_floats] as written by eitb_nms (xyxy first), n [B];
 *   xyxy [B,max_det,4] float32 out (rows >= n[b] are zeroed) -- the input of eitb_rib_select. */
int eitb_scale_boxes(const float* dets, const int32_t* n, int B, int max_det, int row_floats, float gain,
                     float pad_x, float pad_y, float orig_w, float orig_h, float* xyxy,
                     eitb_stream_t stream);

/* ---- K6: mask decode fused with the label-image overlay ---------------------------------------
 * Restates ultralytics process_mask(upsample=True) -- coef x proto
 * contraction, crop to box/4, bilinear x(H/mh) upsample, threshold -- fused with
 * create_segmentations_masks + overlay_segmentation_masks (utils.py:437-523,395-434): per
 * pixel OR of the colour codes of every instance covering it.
 *   dets/n_det as written by eitb_nms; protos [B,nm,mh,mw] of proto_dtype, stored NCHW
 *   (proto_channels_last 0) or NHWC, i.e. [B,mh,mw,nm] (1: what a channels-last network emits)
 *   variant   bit 0 -- 0: logits, interpolate, > 0 (8.3.x)   1: sigmoid, interpolate, > 0.5 (8.0-8.2)
 *             bit 2 -- the CPU crop of late-2025 ultralytics: images with fewer than 50 detections
 *             crop with boxes.round().int() used as Python slice bounds instead of the float comparison
 *             bit 5 (0x20) -- fp16 prototypes with nm == 32: contraction as tcgen05.mma with the accumulator in
 *             tensor memory (same semantics).  Not the default: the contraction is K = 32 per pixel and the kernel
 *             is bound by the crop / upsample / threshold epilogue, where the CUDA-core kernel is 1.3-1.6x faster
 *             (bench.py kernels_isolated).
 *             bit 4 (0x10) -- the CUDA-core kernel with scalar fp32 FMAs for the contraction.  Without bits 4 and 5, fp16
 *             NHWC prototypes with nm == 32 (the network's own output) take the same kernel with the contraction as
 *             warp-level mma.sync m16n8k16 (fp32 accumulate, coefficients as fp16 hi + lo): same semantics, another
 *             summation order (<= 1e-5 of the mask pixels differ from the scalar kernel).
 *   code      [B,H,W] u8 out: overlay codes
 *   inst_area [B,max_det] int32 out or NULL: mask pixel count (the empty-mask filter)
 *   inst_bits [B,max_det,H,W/8] u8 out or NULL: per-instance bit masks (LSB = lowest x)
 * H == 4*mh, W == 4*mw. */
size_t eitb_mask_decode_workspace_bytes(int B, int max_det, int nm, int mh, int mw);
int eitb_mask_decode(const float* dets, const int32_t* n_det, int max_det, const void* protos,
                     int proto_dtype, int proto_channels_last, int B, int nm, int mh, int mw, int H, int W, int variant,
                     uint8_t* code, int32_t* inst_area, uint8_t* inst_bits, void* ws,
                     size_t ws_bytes, eitb_stream_t stream);

/* ---- K7: label-image clean-up -------------------------------------------------------------------
 * Replaces clear_color_output (utils.py:691-755; skipped when body == NULL, utils.py:1005)
 * followed by highlight_small_masks (utils.py:758-843), on code images, in place. */
size_t eitb_label_cleanup_workspace_bytes(int B, int H, int W);
int eitb_label_cleanup(uint8_t* code, const uint8_t* body, int B, int H, int W, void* ws,
                       size_t ws_bytes, eitb_stream_t stream);

/* code image -> the reference's BGR colour image, [n] -> [n,3]. */
int eitb_codes_to_bgr(const uint8_t* code, uint8_t* bgr, int64_t n, eitb_stream_t stream);

/* Per-function entry points of the mirror (the fused kernels cover these rules on the batched path):
 *   eitb_apply_mask_u8  cv2.bitwise_and(img, img, mask=m), ai_tools.py:212: out[p, c] = mask[p] ? img[p, c] : 0
 *   eitb_class_images   create_segmentations_masks, utils.py:437-523: masks [n, n_px] fp32 (> 0 = set), cls [n] int32
 *                       -> bgr4 [4, n_px, 3]: per class (bone, muscles, lung, adipose) the union painted in its colour
 *   eitb_bgr_or_code    overlay_segmentation_masks, utils.py:395-434, one class image at a time: code[p] |= value where
 *                       any channel of bgr[p] is non-zero (code must be initialised by the caller) */
int eitb_apply_mask_u8(const uint8_t* img, const uint8_t* mask, int64_t n_px, int channels, uint8_t* out, eitb_stream_t stream);
int eitb_class_images(const float* masks, const int32_t* cls, int n, int64_t n_px, uint8_t* bgr4, eitb_stream_t stream);
int eitb_bgr_or_code(const uint8_t* bgr, int64_t n_px, int value, uint8_t* code, eitb_stream_t stream);

/* ---- K9: conv epilogue of the CNN ------------------------------------------------------------------
 * In-place per-channel bias (BatchNorm folded into the convolution, as ultralytics' fuse() does for
 * the models loaded at ai_tools.py:69-71) + SiLU on a channels-last activation tensor.
 *   x [n_pixels, C] of dtype (EITB_F16 / EITB_BF16), C % 8 == 0; bias [C] float32 or NULL;
 *   act 1: SiLU, 0: bias only. */
int eitb_bias_act_nhwc(void* x, int dtype, long long n_pixels, int C, const float* bias, int act,
                       eitb_stream_t stream);

/* General form: y = act(src + bias [+ residual]); y is written to `out` (may alias src; NULL to skip)
 * and/or into channels [out2_off, out2_off + C) of a wider channels-last tensor out2 [n_pixels, out2_C]
 * (the concat buffer of a C3k2 / C3k block: no separate torch.cat, no separate residual add). */
int eitb_conv_epilogue_nhwc(const void* src, int dtype, long long n_pixels, int C, const float* bias, int act,
                            const void* residual, void* out, void* out2, int out2_C, int out2_off,
                            eitb_stream_t stream);

/* Upsample(x2, nearest) + Concat of the YOLO11 neck in one pass: a [B,h,w,Ca], b [B,2h,2w,Cb]
 * -> out [B,2h,2w,Ca+Cb], all channels-last 16-bit. */
int eitb_upsample2x_concat_nhwc(const void* a, const void* b, void* out, int dtype, int B, int h, int w, int Ca,
                                int Cb, eitb_stream_t stream);

/* ---- K10: non-convolution tails of the network, fp16 channels-last ------------------------------------
 * Detect/Segment inference tail: box[l] [B,h_l,w_l,64] DFL logits, cls[l]
 * [B,h_l,w_l,nc] class logits, mc[l] [B,h_l,w_l,nm] for the 3 levels (host arrays of device pointers,
 * sizes and strides) -> head [B,4+nc+nm,A] fp16: xywh in input pixels, sigmoid scores, coefficients.
 * box_bias / cls_bias / mc_bias: per-level device float32 bias vectors of the last convolution of each
 * branch, added on the fly (host arrays of 3 pointers, or NULL when the convolutions carry their bias).
 * cls_cstride: halves per pixel of the cls buffers (0 = nc; K11 pads nc to a multiple of 8). */
int eitb_yolo_head_decode(const void* const* box, const void* const* cls, const void* const* mc,
                          const float* const* box_bias, const float* const* cls_bias, const float* const* mc_bias,
                          const int* hs, const int* ws, const int* strides, int B, int nc, int nm, int cls_cstride,
                          void* head, eitb_stream_t stream);

/* SPPF (yaml layer 9): x [B,h,w,C] -> out [B,h,w,4C] = x | maxpool5 | maxpool9 | maxpool13 (-inf padding). */
int eitb_sppf_pool_concat(const void* x, int B, int h, int w, int C, void* out, eitb_stream_t stream);

/* ---- K11 / K12: the convolutions of the networks -----------------------------------------------------
 * Replace cuDNN + the K9 pass for the YOLO11s-seg models loaded at ai_tools.py:69-71 (called at
 * ai_tools.py:121-122,153): implicit GEMM on tcgen05 tensor cores, operands moved by TMA, the Conv
 * epilogue fused.  All activations are NHWC fp16; a tensor is addressed as (pointer, channels per pixel of
 * the buffer `*_ctot`, first channel `*_coff`), so producers write and consumers read channel slices of
 * concat buffers in place (ctot and coff multiples of 8).
 *   y = act(conv(x, w) + bias) + res          (res added after the activation: Bottleneck shortcut)
 *   x [N,H,W,x_ctot], channels [x_coff, x_coff+Cin), Cin % 16 == 0
 *   w_packed [ksize*ksize][ceil16(Cout)][Cin] fp16 (tap-major, K contiguous; padded output rows zero)
 *   bias [Cout] float32 or NULL; act 1 = SiLU, 0 = none; res NULL or [N,Ho,Wo,res_ctot] slice (res_mode 1)
 *   res_mode 2 (ksize 1, stride 1, even H and W): y = act(conv(x, w) + bias + up2(res)), res [N,Ho/2,Wo/2,res_ctot]
 *   read with nearest-neighbour x2 upsampling.  A 1x1 convolution commutes with Upsample, so
 *   Conv1x1(Concat(Upsample(a), b)) = act(up2(Wa a) + Wb b + bias): the neck's Upsample + Concat tensors
 *   (yaml layers 11-12, 14-15) are never built
 *   ksize 1 or 3 (padding ksize/2), stride 1 or 2;  Ho = (H + 2*pad - ksize)/stride + 1
 *   y_up 1: y [N,Ho,Wo,y_ctot].  y_up 2: y [N,2Ho,2Wo,y_ctot] and the result lands on pixels
 *   (2*oy + y_dy, 2*ox + y_dx) -- one of the four phases of ConvTranspose2d(k=2, s=2). */
int eitb_conv2d_nhwc(const void* x, int N, int H, int W, int x_ctot, int x_coff, int Cin,
                     const void* w_packed, const float* bias, int Cout, int ksize, int stride, int act,
                     const void* res, int res_ctot, int res_coff, int res_mode,
                     void* y, int y_ctot, int y_coff, int y_up, int y_dy, int y_dx, eitb_stream_t stream);
/* tile-shape experiments (host, process-wide): largest MMA N, deepest shared-memory ring, grid cap; <= 0 keeps a value */
int eitb_conv2d_tuning(int ntile_max, int stage_cap, int grid_cap);
/* kernel experiments (host, process-wide; 0 = production): 1 skip the TMA store, 2 skip the staging writes,
 * 8 halo-mode descriptors carry base_offset, 16 halo mode off (nine shifted TMA loads per K chunk instead) */
int eitb_conv2d_debug(int flags);

/* Stem Conv(3 -> 32, k3, s2, pad 1): x [N,H,W,3] fp16, w27 [27][32] float32 ((r*3+s)*3+ci major).
 * gray 1: the caller guarantees three equal input channels (a gray slice replicated by K1 / the letterbox,
 * ai_tools.py:120,135): channel 0 is read and the weights are summed over the input channel.
 * gray 2: x is the u8 image itself, [N,H,W] (K1's out_u8): the kernel applies the ultralytics preprocess
 * (u8 -> fp16 / 255, three equal channels) while it stages its input tile, so the normalised NCHW tensor is
 * never written to HBM. */
int eitb_stem_conv3x3s2_nhwc(const void* x, int N, int H, int W, const float* w27, const float* bias, int Cout, int act,
                             int gray, void* y, int y_ctot, int y_coff, eitb_stream_t stream);
/* gray 2 runs its 9 x 32 contraction as warp-level mma.sync (fp16 inputs, weight sums as fp16 hi + lo, fp32 accumulate);
 * eitb_stem_debug(1) selects the CUDA-core kernel instead (host, process-wide; A/B runs and tests). */
int eitb_stem_debug(int scalar);
/* Depthwise Conv(C -> C, k3, s1, pad 1, groups C): w9 [9][C] fp16, C % 8 == 0. */
int eitb_dwconv3x3_nhwc(const void* x, int N, int H, int W, int x_ctot, int x_coff, int C, const void* w9, const float* bias,
                        int act, void* y, int y_ctot, int y_coff, eitb_stream_t stream);

/* ---- K8: per-triangle tissue labelling ----------------------------------------------------------
 * Replaces divide_triangles_into_groups / process_triangle / the CLASS vector of
 * export_mesh_for_femm (mesh_tools/femm_generator.py:12-85,118-184,187-265) with the
 * reference's polygon semantics in fp64: polygons in ascending area order, skipping class
 * == outer_cls; centroid strictly inside -> class, stop; else intersection area / triangle
 * area > 0.5 -> class, stop; else arg-max positive intersection; default outer_cls.
 *   nodes_xy [n_nodes,2] f64; tri [T,3] int64 (0-based node indices)
 *   poly_xy  [V,2] f64 closed rings back to back; poly_off [P+1] int32 vertex offsets;
 *   poly_cls [P] int32; polygons already sorted by ascending area (host mirror does it)
 *   cls_out  [T] int32
 *   V        number of polygon vertices (poly_off[P])
 *   ws       eitb_tri_label_workspace_bytes(P, V) bytes (per-polygon bounding boxes, orientation, and the
 *            y-slab edge lists that let a triangle visit only the edges near it)
 * nodes_xy and poly_xy must be 16-byte aligned. */
size_t eitb_tri_label_workspace_bytes(int P, int V);
int eitb_tri_label(const double* nodes_xy, int64_t n_nodes, const int64_t* tri, int64_t T,
                   const double* poly_xy, const int32_t* poly_off, const int32_t* poly_cls, int P, int V,
                   int outer_cls, int32_t* cls_out, void* ws, size_t ws_bytes, eitb_stream_t stream);

/* Raster mode named by the north star: class of the label-map pixel under the centroid
 * (code -> class: 7->0, 1->1, 6->2, 3->3, else outer_cls).  Not the reference semantics;
 * offered for comparison only. */
int eitb_tri_label_raster(const double* nodes_xy, int64_t n_nodes, const int64_t* tri, int64_t T,
                          const uint8_t* code, int H, int W, int outer_cls, int32_t* cls_out,
                          eitb_stream_t stream);

/* ---- K13: polygons of the cleaned label image ---------------------------------------------------
 * Replaces create_list_crd_from_color_output (utils.py:1191-1279) and get_only_body_mask_contours
 * (utils.py:1157-1188): per tissue colour in the reference's dict order (classes "3","0","1","2"),
 * cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) -> approxPolyDP(0.001 * arcLength, closed)
 * -> closing point when the polygon has more than two points; then, when body != NULL, the body
 * outline (class 4: every border pixel of the last external contour with >= 5 points).
 *   code   [B,H,W] u8 code image (K6/K7 output); body [B,H,W] u8 or NULL; W % 32 == 0, W <= 1024,
 *          both 16-byte aligned
 *   n_polys [B]; poly_cls [B,max_polys]; poly_off [B,max_polys+1] point offsets; points_xy
 *          [B,max_points,2] int32 (x, y); polygons in the order the reference lists them
 *   status [B] bit 0: more than max_polys polygons, bit 1: more than max_points points, bit 2:
 *          contour scratch (3*H*W raw points) exhausted -- the lists are truncated; bit 3: a body mask
 *          was given but has no outline with >= 5 points (the reference appends [] then) */
size_t eitb_label_polygons_workspace_bytes(int B, int H, int W, int max_polys);
int eitb_label_polygons(const uint8_t* code, const uint8_t* body, int B, int H, int W, int max_polys,
                        int max_points, int32_t* n_polys, int32_t* poly_cls, int32_t* poly_off,
                        int32_t* points_xy, int32_t* status, void* ws, size_t ws_bytes,
                        eitb_stream_t stream);

/* K13 -> K8 without leaving the device: what create_mesh / divide_triangles_into_groups /
 * build_polygons_with_area (mesh_tools/femm_generator.py:454-459, 49-60, 88-115) do to that list:
 * the class-4 outline is the outer contour and leaves the list, polygons with fewer than four points
 * are dropped, rings are closed, and the rest is sorted (stable) by ascending shoelace area.
 *   out_xy [B,max_points+max_polys,2] f64; out_off [B,max_polys+1]; out_cls [B,max_polys]; out_n [B]
 * -- per image exactly eitb_tri_label's poly_xy / poly_off / poly_cls / P. */
size_t eitb_polygons_for_mesh_workspace_bytes(int B, int max_polys);
int eitb_polygons_for_mesh(const int32_t* n_polys, const int32_t* poly_cls, const int32_t* poly_off,
                           const int32_t* points_xy, int B, int max_polys, int max_points,
                           double* out_xy, int32_t* out_off, int32_t* out_cls, int32_t* out_n,
                           void* ws, size_t ws_bytes, eitb_stream_t stream);

/* ---- host codecs for compressed DICOM pixel data (ingest) -------------------------
 * Replace what pydicom + pylibjpeg do behind pydicom.dcmread(...).pixel_array (utils.py:52-60, 98;
 * requirements.txt:9-13) for the two lossless syntaxes of CT exports.  Host pointers, no CUDA.
 *   eitb_rle_decode_frame: one RLE Lossless frame (PS3.5 Annex G) -> rows*cols samples, little endian
 *   eitb_jpeg_lossless_decode: a JPEG lossless (SOF3, one component, 2-16 bits, any predictor, restart
 *   intervals) stream -> out[rows*cols] uint16; out == NULL only reports rows / cols / precision */
int eitb_rle_decode_frame(const uint8_t* frag, size_t frag_len, int rows, int cols, int bytes_per_sample,
                          uint8_t* out);
int eitb_jpeg_lossless_decode(const uint8_t* data, size_t len, int* rows, int* cols, int* precision,
                              uint16_t* out, size_t out_capacity);

#ifdef __cplusplus
}
#endif
#endif /* EITB200_H */
