// K3: coronal mid-row gather + min/max, MINMAX->u8 normalise; a11 letterbox; K4 rib arg-select.
//
// Reference: convert_to_3d + axial_to_sagittal + mid-plane (utils.py:73-163,
// ai_tools.py:98-101) reduce to "one row of every slice" (SURVEY §8 a3); the rib model's
// preprocess is ultralytics LetterBox + /255 (Appendix A.2); the slice pick is
// search_number_axial_slice (utils.py:166-269).
#include "common.cuh"
#include <limits.h>

namespace {

// ---------------------------------------------------------------- front rows
__global__ void __launch_bounds__(256)
front_rows_kernel(const int16_t* __restrict__ px, const int32_t* __restrict__ order, int n, int H, int W,
                  int row, int flip_x, int flip_z, int16_t* __restrict__ rows, int32_t* __restrict__ minmax,
                  const int32_t* __restrict__ geom, long long series_stride) {
    if (geom) {                                          // batched form: blockIdx.y = series, per-series geometry
        const int s = blockIdx.y;
        row = geom[3 * s]; flip_x = geom[3 * s + 1]; flip_z = geom[3 * s + 2];
        px += (long long)s * series_stride;
        if (order) order += (long long)s * n;
        rows += (long long)s * n * W;
        minmax += 2 * s;
    }
    const int upr = W / 8;                               // 16-byte units per row
    const long long n_units = (long long)n * upr;
    int mn = INT_MAX, mx = INT_MIN;
    for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < n_units;
         u += (long long)gridDim.x * blockDim.x) {
        const int z = (int)(u / upr);
        const int c = (int)(u - (long long)z * upr);
        const int zs = flip_z ? n - 1 - z : z;
        const long long sl = order ? order[zs] : zs;
        const int sc = flip_x ? upr - 1 - c : c;
        const int4 raw = *reinterpret_cast<const int4*>(px + (sl * H + row) * (long long)W + sc * 8);
        int w[4] = {raw.x, raw.y, raw.z, raw.w};
        if (flip_x) {                                    // reverse the 8 halves
            int r0 = __byte_perm(w[3], 0, 0x1032), r1 = __byte_perm(w[2], 0, 0x1032);
            int r2 = __byte_perm(w[1], 0, 0x1032), r3 = __byte_perm(w[0], 0, 0x1032);
            w[0] = r0; w[1] = r1; w[2] = r2; w[3] = r3;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int a = (int)(short)(w[k] & 0xffff), b = w[k] >> 16;
            mn = min(mn, min(a, b));
            mx = max(mx, max(a, b));
        }
        *reinterpret_cast<int4*>(rows + ((long long)z * W + c * 8)) = make_int4(w[0], w[1], w[2], w[3]);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    __shared__ int smn[8], smx[8];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { smn[wid] = mn; smx[wid] = mx; }
    __syncthreads();
    if (wid == 0) {
        mn = lane < (int)(blockDim.x >> 5) ? smn[lane] : INT_MAX;
        mx = lane < (int)(blockDim.x >> 5) ? smx[lane] : INT_MIN;
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (lane == 0 && mn <= mx) { atomicMin(minmax, mn); atomicMax(minmax + 1, mx); }
    }
}

// ---------------------------------------------------------------- MINMAX -> u8
// OpenCV: scale = 255 * (1/(max-min)) (0 if max-min <= DBL_EPSILON), shift = -min*scale in
// double; convertTo(CV_8U) evaluates fmaf((float)v, (float)scale, (float)shift), rounds
// half-to-even and saturates (checked against cv2 4.13 in tests/test_oracle_imaging.py).
__global__ void __launch_bounds__(256)
minmax_u8_kernel(const int16_t* __restrict__ rows, long long count, const int32_t* __restrict__ minmax,
                 uint8_t* __restrict__ out) {
    const double mn = (double)minmax[0], mx = (double)minmax[1];
    const double d = mx - mn;
    const double scale = 255.0 * (d > 2.220446049250313e-16 ? 1.0 / d : 0.0);
    const float a = (float)scale, b = (float)(0.0 - mn * scale);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (long long)gridDim.x * blockDim.x) {
        const float v = fmaf((float)rows[i], a, b);
        int r = __float2int_rn(v);
        out[i] = (uint8_t)min(max(r, 0), 255);
    }
}

// ---------------------------------------------------------------- letterbox
// cv2.resize(INTER_LINEAR) for 8-bit images, restated: 11-bit fixed-point coefficients,
// horizontal pass in int32, vertical pass ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2.
__device__ __forceinline__ void lin_coef(int dpos, double scale, int ssize, bool clamp_edges, int& s0, int& c0, int& c1) {
    float f = (float)(((double)dpos + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (clamp_edges) {
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
    }
    s0 = s;
    // saturate_cast<short>(x * 2048): round-half-even
    c0 = __float2int_rn((1.f - f) * 2048.f);
    c1 = __float2int_rn(f * 2048.f);
}

template <typename T>
__global__ void __launch_bounds__(256)
letterbox_kernel(const uint8_t* __restrict__ gray, int B, int H, int W, int nh, int nw, int top, int left,
                 int outH, int outW, T* __restrict__ out) {
    const double scale_x = 1.0 / ((double)nw / (double)W);
    const double scale_y = 1.0 / ((double)nh / (double)H);
    const long long total = (long long)B * outH * outW;
    const long long plane = (long long)outH * outW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / plane);
        const int rem = (int)(i - (long long)b * plane);
        const int oy = rem / outW, ox = rem - oy * outW;
        const int dy = oy - top, dx = ox - left;
        int u8 = 114;
        if (dy >= 0 && dy < nh && dx >= 0 && dx < nw) {
            const uint8_t* src = gray + (long long)b * H * W;
            if (nh == H && nw == W) {
                u8 = src[dy * W + dx];
            } else {
                int sx, a0, a1, sy, b0, b1;
                lin_coef(dx, scale_x, W, true, sx, a0, a1);
                lin_coef(dy, scale_y, H, false, sy, b0, b1);
                const int y0 = min(max(sy, 0), H - 1), y1 = min(max(sy + 1, 0), H - 1);
                int S0, S1;
                if (sx >= W - 1) {
                    S0 = src[y0 * W + W - 1] * 2048;
                    S1 = src[y1 * W + W - 1] * 2048;
                } else {
                    S0 = src[y0 * W + sx] * a0 + src[y0 * W + sx + 1] * a1;
                    S1 = src[y1 * W + sx] * a0 + src[y1 * W + sx + 1] * a1;
                }
                u8 = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
                u8 = min(max(u8, 0), 255);
            }
        }
        const T v = unit_from_u8<T>(u8);
        T* o = out + (long long)b * 3 * plane + rem;
        o[0] = v; o[plane] = v; o[2 * plane] = v;
    }
}

// ---------------------------------------------------------------- rib select
// One warp per series.  Right-side boxes (x1 > image_width/2), stable ascending order of
// y1: rank(i) = #{j : y1[j] < y1[i] or (y1[j] == y1[i] and j < i)} over right-side boxes.
__global__ void rib_select_kernel(const float* __restrict__ xyxy, const int32_t* __restrict__ k, int S, int max_k,
                                  float image_width, const int32_t* __restrict__ custom, int32_t* __restrict__ out) {
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= S) return;
    const float mid = image_width / 2;
    const float* bx = xyxy + (long long)s * max_k * 4;
    const int n = min(k[s], max_k);
    float y6 = 0.f, y7 = 0.f;
    int have = 0, n_right = 0;
    for (int i = lane; i < ((n + 31) & ~31); i += 32) {
        const bool right = i < n && bx[i * 4] > mid;
        n_right += __popc(__ballot_sync(0xffffffffu, right));
        int rank = -1;
        if (right) {
            const float yi = bx[i * 4 + 1];
            rank = 0;
            for (int j = 0; j < n; ++j) {
                if (bx[j * 4] > mid) {
                    const float yj = bx[j * 4 + 1];
                    rank += (yj < yi) || (yj == yi && j < i);
                }
            }
        }
        const unsigned m5 = __ballot_sync(0xffffffffu, rank == 5);
        const unsigned m6 = __ballot_sync(0xffffffffu, rank == 6);
        const float yv = right ? bx[i * 4 + 1] : 0.f;
        if (m5) { y6 = __shfl_sync(0xffffffffu, yv, __ffs(m5) - 1); have |= 1; }
        if (m6) { y7 = __shfl_sync(0xffffffffu, yv, __ffs(m6) - 1); have |= 2; }
    }
    if (lane == 0) {
        int32_t* o = out + s * 4;
        if (have == 3 && n_right >= 7) {
            const float sum = __fadd_rn(y6, y7);          // float32 add, utils.py:260
            const float half = fabsf(sum) / 2;             // exact
            o[0] = (int)y6; o[1] = (int)y7;
            o[2] = (int)half + (custom ? custom[s] : 0);
            o[3] = 1;
        } else {
            o[0] = o[1] = o[2] = 0; o[3] = 0;
        }
    }
}

// ultralytics scale_boxes: network-input px -> original-image px.
__global__ void scale_boxes_kernel(const float* __restrict__ dets, const int32_t* __restrict__ n, int B, int max_det,
                                   int D, float gain, float pad_x, float pad_y, float w0, float h0,
                                   float* __restrict__ xyxy) {
    const int total = B * max_det;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int b = i / max_det, r = i - b * max_det;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < n[b]) {
            const float* d = dets + (long long)i * D;
            o.x = fminf(fmaxf(__fdiv_rn(__fsub_rn(d[0], pad_x), gain), 0.f), w0);
            o.y = fminf(fmaxf(__fdiv_rn(__fsub_rn(d[1], pad_y), gain), 0.f), h0);
            o.z = fminf(fmaxf(__fdiv_rn(__fsub_rn(d[2], pad_x), gain), 0.f), w0);
            o.w = fminf(fmaxf(__fdiv_rn(__fsub_rn(d[3], pad_y), gain), 0.f), h0);
        }
        reinterpret_cast<float4*>(xyxy)[i] = o;
    }
}

}  // namespace

extern "C" int eitb_scale_boxes(const float* dets, const int32_t* n, int B, int max_det, int row_floats, float gain,
                                float pad_x, float pad_y, float orig_w, float orig_h, float* xyxy,
                                eitb_stream_t stream) {
    if (!dets || !n || !xyxy || B < 0 || max_det <= 0 || row_floats < 4 || !(gain > 0.f)) return EITB_ERR_BAD_ARG;
    if (B == 0) return EITB_OK;
    eitb_prof_begin("scale_boxes_kernel", (cudaStream_t)stream);
    scale_boxes_kernel<<<eitb_grid((long long)B * max_det, 256, 4), 256, 0, (cudaStream_t)stream>>>(
        dets, n, B, max_det, row_floats, gain, pad_x, pad_y, orig_w, orig_h, xyxy);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

extern "C" int eitb_front_rows(const int16_t* px, const int32_t* order, int n, int H, int W, int row,
                               int flip_x, int flip_z, int16_t* rows, int32_t* minmax, eitb_stream_t stream) {
    if (!px || !rows || !minmax || n < 0 || H <= 0 || W <= 0 || row < 0 || row >= H) return EITB_ERR_BAD_ARG;
    if (W % 8) return EITB_ERR_UNSUPPORTED;
    if (n == 0) return EITB_OK;
    const long long units = (long long)n * (W / 8);
    eitb_prof_begin("front_rows_kernel", (cudaStream_t)stream);
    front_rows_kernel<<<eitb_grid(units, 256, 4), 256, 0, (cudaStream_t)stream>>>(px, order, n, H, W, row, flip_x, flip_z, rows, minmax,
                                                                                  nullptr, 0);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

extern "C" int eitb_front_rows_batch(const int16_t* px, long long series_stride, const int32_t* order, const int32_t* geom,
                                     int S, int n, int H, int W, int16_t* rows, int32_t* minmax, eitb_stream_t stream) {
    if (!px || !geom || !rows || !minmax || S < 0 || n < 0 || H <= 0 || W <= 0) return EITB_ERR_BAD_ARG;
    if (W % 8 || S > 65535) return EITB_ERR_UNSUPPORTED;
    if (n == 0 || S == 0) return EITB_OK;
    const long long units = (long long)n * (W / 8);
    int gx = eitb_grid(units, 256, 4);
    if (gx > 8 && S >= 8) gx = 8;                        // many series: a few CTAs each already fill the chip
    eitb_prof_begin("front_rows_kernel", (cudaStream_t)stream);
    front_rows_kernel<<<dim3(gx, S), 256, 0, (cudaStream_t)stream>>>(px, order, n, H, W, 0, 0, 0, rows, minmax, geom, series_stride);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

extern "C" int eitb_minmax_u8(const int16_t* rows, int64_t count, const int32_t* minmax, uint8_t* out,
                              eitb_stream_t stream) {
    if (!rows || !minmax || !out || count < 0) return EITB_ERR_BAD_ARG;
    if (count == 0) return EITB_OK;
    eitb_prof_begin("minmax_u8_kernel", (cudaStream_t)stream);
    minmax_u8_kernel<<<eitb_grid(count, 256, 4), 256, 0, (cudaStream_t)stream>>>(rows, count, minmax, out);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

extern "C" int eitb_letterbox_nchw(const uint8_t* gray, int B, int H, int W, int nh, int nw, int top, int left,
                                   int outH, int outW, void* out, int out_dtype, eitb_stream_t stream) {
    if (!gray || !out || B < 0 || H <= 0 || W <= 0 || nh <= 0 || nw <= 0 || top < 0 || left < 0 ||
        top + nh > outH || left + nw > outW)
        return EITB_ERR_BAD_ARG;
    if (B == 0) return EITB_OK;
    const long long total = (long long)B * outH * outW;
    const int grid = eitb_grid(total, 256, 8);
    cudaStream_t s = (cudaStream_t)stream;
    eitb_prof_begin("letterbox_kernel", s);
    switch (out_dtype) {
        case EITB_F32: letterbox_kernel<float><<<grid, 256, 0, s>>>(gray, B, H, W, nh, nw, top, left, outH, outW, (float*)out); break;
        case EITB_F16: letterbox_kernel<__half><<<grid, 256, 0, s>>>(gray, B, H, W, nh, nw, top, left, outH, outW, (__half*)out); break;
        case EITB_BF16: letterbox_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(gray, B, H, W, nh, nw, top, left, outH, outW, (__nv_bfloat16*)out); break;
        default: return EITB_ERR_BAD_ARG;
    }
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

extern "C" int eitb_rib_select(const float* xyxy, const int32_t* k, int S, int max_k, float image_width,
                               const int32_t* custom, int32_t* out, eitb_stream_t stream) {
    if (!xyxy || !k || !out || S < 0 || max_k <= 0) return EITB_ERR_BAD_ARG;
    if (S == 0) return EITB_OK;
    const int warps = 4;
    eitb_prof_begin("rib_select_kernel", (cudaStream_t)stream);
    rib_select_kernel<<<eitb_div_up(S, warps), warps * 32, 0, (cudaStream_t)stream>>>(xyxy, k, S, max_k, image_width, custom, out);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}
