// K1: HU window + normalise + rot180 + body-mask AND + 3-channel NCHW write.
//
// Reference semantics: classic_norm (kt_service/ai_tools/utils.py:272-313) -- clip to
// [lo,hi], ((c-lo)/(hi-lo)*255).astype(uint8) (== ((c-lo)*255)//(hi-lo), exact integer
// form), rotate 180 -- then cv2.bitwise_and(norm, norm, mask=body) (ai_tools.py:212) and
// the ultralytics preprocess for square inputs (3 equal channels, NCHW, /255).
//
// HBM-bound streaming kernel: one 16-byte load (8 int16) per thread-unit, reversed in
// registers for the rotation, optional 8-byte mask load, one 8-byte u8 store and three
// 16/32-byte channel stores.  The window->u8->T mapping is a per-CTA shared-memory LUT
// (<= 2049 entries) so the per-pixel work is clamp + one LDS.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxLut = 2048;   // window widths up to this use the smem LUT

template <typename T> struct Pack8;   // 8 outputs of type T
template <> struct Pack8<__half> { int4 v; };
template <> struct Pack8<__nv_bfloat16> { int4 v; };
template <> struct Pack8<float> { int4 v[2]; };

template <typename T>
__device__ __forceinline__ void store8(T* dst, const T (&x)[8]) {
    if constexpr (sizeof(T) == 2) {
        int4 v;
        const uint16_t* h = reinterpret_cast<const uint16_t*>(x);
        v.x = h[0] | (h[1] << 16); v.y = h[2] | (h[3] << 16);
        v.z = h[4] | (h[5] << 16); v.w = h[6] | (h[7] << 16);
        st_stream_int4(reinterpret_cast<int4*>(dst), v);
    } else {
        const int* f = reinterpret_cast<const int*>(x);
        st_stream_int4(reinterpret_cast<int4*>(dst), make_int4(f[0], f[1], f[2], f[3]));
        st_stream_int4(reinterpret_cast<int4*>(dst) + 1, make_int4(f[4], f[5], f[6], f[7]));
    }
}

template <typename T, bool USE_LUT>
__global__ void __launch_bounds__(kThreads)
hu_window_kernel(const int16_t* __restrict__ px, long long n_units, int units_per_slice, int lo,
                 int hi, int rot180, const uint8_t* __restrict__ mask, uint8_t* __restrict__ out_u8,
                 T* __restrict__ out_nchw) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int d = hi - lo;
    T* lutT = reinterpret_cast<T*>(smem);
    uint8_t* lut8 = smem + (size_t)(d + 1) * sizeof(T);
    if constexpr (USE_LUT) {
        for (int v = threadIdx.x; v <= d; v += kThreads) {
            int u = (v * 255) / d;
            lut8[v] = (uint8_t)u;
            lutT[v] = unit_from_u8<T>(u);
        }
        __syncthreads();
    }
    const long long hw = (long long)units_per_slice * 8;
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long u = (long long)blockIdx.x * kThreads + threadIdx.x; u < n_units; u += stride) {
        const long long b = u / units_per_slice;
        const int o = (int)(u - b * units_per_slice);
        const long long in_off = rot180 ? (b * hw + hw - 8 - (long long)o * 8) : (b * hw + (long long)o * 8);
        const int4 raw = ld_stream_int4(reinterpret_cast<const int4*>(px + in_off));
        uint2 mk = make_uint2(0xffffffffu, 0xffffffffu);
        if (mask) mk = ld_stream_uint2(reinterpret_cast<const uint2*>(mask + b * hw + (long long)o * 8));
        const int w[4] = {raw.x, raw.y, raw.z, raw.w};
        T vals[8];
        uint32_t u8lo = 0, u8hi = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int src = rot180 ? 7 - j : j;
            int x = (int)(short)((w[src >> 1] >> ((src & 1) * 16)) & 0xffff);
            x = min(max(x, lo), hi) - lo;
            const uint32_t m = ((j < 4 ? mk.x : mk.y) >> ((j & 3) * 8)) & 0xffu;
            int u8;
            T t;
            if constexpr (USE_LUT) {
                u8 = lut8[x];
                t = lutT[x];
            } else {
                u8 = (int)(((unsigned)x * 255u) / (unsigned)d);
                t = unit_from_u8<T>(u8);
            }
            if (m == 0) { u8 = 0; t = unit_from_u8<T>(0); }
            vals[j] = t;
            if (j < 4) u8lo |= (uint32_t)u8 << (8 * j); else u8hi |= (uint32_t)u8 << (8 * (j - 4));
        }
        const long long out_pix = b * hw + (long long)o * 8;
        if (out_u8) st_stream_uint2(reinterpret_cast<uint2*>(out_u8 + out_pix), make_uint2(u8lo, u8hi));
        if (out_nchw) {
            T* base = out_nchw + b * 3 * hw + (long long)o * 8;
            store8(base, vals);
            store8(base + hw, vals);
            store8(base + 2 * hw, vals);
        }
    }
}

template <typename T>
int launch_hu(const int16_t* px, int B, int H, int W, int lo, int hi, int rot180, const uint8_t* mask,
              uint8_t* out_u8, void* out_nchw, cudaStream_t s) {
    const long long n_units = (long long)B * H * W / 8;
    const int ups = H * W / 8;
    const int d = hi - lo;
    const int grid = eitb_grid(n_units, kThreads, 8);
    if (d <= kMaxLut) {
        size_t smem = (size_t)(d + 1) * (sizeof(T) + 1);
        eitb_prof_begin("hu_window_kernel", s);
        hu_window_kernel<T, true><<<grid, kThreads, smem, s>>>(px, n_units, ups, lo, hi, rot180, mask, out_u8, (T*)out_nchw);
    } else {
        eitb_prof_begin("hu_window_kernel", s);
        hu_window_kernel<T, false><<<grid, kThreads, 0, s>>>(px, n_units, ups, lo, hi, rot180, mask, out_u8, (T*)out_nchw);
    }
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
u8_to_nchw_kernel(const uint8_t* __restrict__ gray, long long n_units, int units_per_slice, T* __restrict__ out) {
    __shared__ T lut[256];
    for (int v = threadIdx.x; v < 256; v += kThreads) lut[v] = unit_from_u8<T>(v);
    __syncthreads();
    const long long hw = (long long)units_per_slice * 8;
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long u = (long long)blockIdx.x * kThreads + threadIdx.x; u < n_units; u += stride) {
        const long long b = u / units_per_slice;
        const int o = (int)(u - b * units_per_slice);
        const uint2 g = ld_stream_uint2(reinterpret_cast<const uint2*>(gray + u * 8));
        T vals[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) vals[j] = lut[((j < 4 ? g.x : g.y) >> ((j & 3) * 8)) & 0xffu];
        T* base = out + b * 3 * hw + (long long)o * 8;
        store8(base, vals);
        store8(base + hw, vals);
        store8(base + 2 * hw, vals);
    }
}

}  // namespace

extern "C" int eitb_hu_window_nchw(const int16_t* px, int B, int H, int W, int lo, int hi, int rot180,
                                   const uint8_t* body_mask, uint8_t* out_u8, void* out_nchw,
                                   int out_dtype, eitb_stream_t stream) {
    if (!px || B < 0 || H <= 0 || W <= 0 || hi <= lo || lo < -32768 || hi > 32767) return EITB_ERR_BAD_ARG;
    if (!out_u8 && !out_nchw) return EITB_ERR_BAD_ARG;
    if ((W % 8) != 0) return EITB_ERR_UNSUPPORTED;
    if (B == 0) return EITB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    switch (out_dtype) {
        case EITB_F32: return launch_hu<float>(px, B, H, W, lo, hi, rot180, body_mask, out_u8, out_nchw, s);
        case EITB_F16: return launch_hu<__half>(px, B, H, W, lo, hi, rot180, body_mask, out_u8, out_nchw, s);
        case EITB_BF16: return launch_hu<__nv_bfloat16>(px, B, H, W, lo, hi, rot180, body_mask, out_u8, out_nchw, s);
        default: return EITB_ERR_BAD_ARG;
    }
}

extern "C" int eitb_u8_to_nchw(const uint8_t* gray, int B, int H, int W, void* out_nchw, int out_dtype,
                               eitb_stream_t stream) {
    if (!gray || !out_nchw || B < 0 || H <= 0 || W <= 0) return EITB_ERR_BAD_ARG;
    if ((W % 8) != 0) return EITB_ERR_UNSUPPORTED;
    if (B == 0) return EITB_OK;
    const long long n_units = (long long)B * H * W / 8;
    const int ups = H * W / 8;
    const int grid = eitb_grid(n_units, kThreads, 8);
    cudaStream_t s = (cudaStream_t)stream;
    eitb_prof_begin("u8_to_nchw_kernel", s);
    switch (out_dtype) {
        case EITB_F32: u8_to_nchw_kernel<float><<<grid, kThreads, 0, s>>>(gray, n_units, ups, (float*)out_nchw); break;
        case EITB_F16: u8_to_nchw_kernel<__half><<<grid, kThreads, 0, s>>>(gray, n_units, ups, (__half*)out_nchw); break;
        case EITB_BF16: u8_to_nchw_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(gray, n_units, ups, (__nv_bfloat16*)out_nchw); break;
        default: return EITB_ERR_BAD_ARG;
    }
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}
