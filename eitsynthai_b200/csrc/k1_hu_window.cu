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
hu_window_kernel(const int16_t* __restrict__ px_all, long long n_units, int units_per_slice, int lo,
                 int hi, int rot180, const uint8_t* __restrict__ mask_all, uint8_t* __restrict__ out_u8_all,
                 T* __restrict__ out_nchw_all) {
    // grid (x: unit blocks, y: slice): no per-unit division
    (void)n_units;
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
    const int hw = units_per_slice * 8;
    const long long b = blockIdx.y;
    const int16_t* px = px_all + b * hw;
    const uint8_t* mask = mask_all ? mask_all + b * hw : nullptr;
    uint8_t* out_u8 = out_u8_all ? out_u8_all + b * hw : nullptr;
    T* out_nchw = out_nchw_all ? out_nchw_all + b * 3 * hw : nullptr;
    for (int o = blockIdx.x * kThreads + threadIdx.x; o < units_per_slice; o += gridDim.x * kThreads) {
        const int in_off = rot180 ? hw - 8 - o * 8 : o * 8;
        const int4 raw = ld_stream_int4(reinterpret_cast<const int4*>(px + in_off));
        uint2 mk = make_uint2(0xffffffffu, 0xffffffffu);
        if (mask) mk = ld_stream_uint2(reinterpret_cast<const uint2*>(mask + o * 8));
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
        const int out_pix = o * 8;
        if (out_u8) st_stream_uint2(reinterpret_cast<uint2*>(out_u8 + out_pix), make_uint2(u8lo, u8hi));
        if (out_nchw) {
            T* base = out_nchw + o * 8;
            store8(base, vals);
            store8(base + hw, vals);
            store8(base + 2 * hw, vals);
        }
    }
}

// ---- 16-bit outputs (fp16 / bf16), window <= 2048: the production path ---------------------
// Two pixels per 32-bit lane all the way: SIMD clamp, one LDS per pixel for the 16-bit pattern of
// u8/255, byte-mask expanded with PRMT, 16-byte stores.  NHWC writes the three equal channels
// interleaved (48 contiguous bytes per 8 pixels), which is what a channels-last cuDNN network reads.
__device__ __forceinline__ uint16_t unit16_bits(int u, int bf16) {
    const float f = __fdiv_rn((float)u, 255.0f);
    return bf16 ? __bfloat16_as_ushort(__float2bfloat16_rn(f)) : __half_as_ushort(__float2half_rn(f));
}

template <bool NHWC, bool WANT_U8>
__global__ void __launch_bounds__(kThreads)
hu_window16_kernel(const int16_t* __restrict__ px_all, int units_per_slice, long long total_units, int lo, int hi, int rot180,
                   const uint8_t* __restrict__ mask_all, uint8_t* __restrict__ out_u8_all,
                   uint16_t* __restrict__ out_all, int bf16) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int d = hi - lo;
    int4* stage = reinterpret_cast<int4*>(smem);                  // [8 warps][96] store staging (NHWC)
    uint16_t* lutT = reinterpret_cast<uint16_t*>(smem + (kThreads / 32) * 96 * sizeof(int4));
    uint8_t* lut8 = reinterpret_cast<uint8_t*>(lutT) + (size_t)(d + 1) * 2;
    for (int v = threadIdx.x; v <= d; v += kThreads) {
        const int u = (v * 255) / d;
        lutT[v] = unit16_bits(u, bf16);
        if (WANT_U8) lut8[v] = (uint8_t)u;
    }
    __syncthreads();
    const int hw = units_per_slice * 8;
    const unsigned lo2 = ((unsigned)lo & 0xffffu) | ((unsigned)lo << 16), hi2 = ((unsigned)hi & 0xffffu) | ((unsigned)hi << 16);
    // One flat grid over the 8-pixel units of ALL slices (exactly one resident wave: a (CTAs per slice, slice) grid left
    // 960 CTAs on 1,184 slots, 7 on some SMs and 6 on others), four units per trip: all four loads (and mask loads) are
    // issued before any arithmetic -- 96 bytes in flight per thread.
    constexpr int kUnr = 4;
    const long long step = (long long)gridDim.x * kThreads;
    for (long long g0 = (long long)blockIdx.x * kThreads + threadIdx.x; g0 < total_units; g0 += kUnr * step) {
        int4 raws[kUnr];
        uint2 mks[kUnr];
        long long bs[kUnr];
        int os[kUnr];
#pragma unroll
        for (int u = 0; u < kUnr; ++u) {
            const long long g = g0 + u * step;
            if (g < total_units) {
                const long long bb = g / units_per_slice;
                const int oo = (int)(g - bb * units_per_slice);
                bs[u] = bb; os[u] = oo;
                raws[u] = ld_stream_int4(reinterpret_cast<const int4*>(px_all + bb * hw + (rot180 ? hw - 8 - oo * 8 : oo * 8)));
                if (mask_all) mks[u] = ld_stream_uint2(reinterpret_cast<const uint2*>(mask_all + bb * hw + oo * 8));
            }
        }
#pragma unroll
        for (int u = 0; u < kUnr; ++u) {
            if (g0 + u * step >= total_units) break;
            const long long b = bs[u];
            const int o = os[u];
            const bool mask = mask_all != nullptr;
            const int4 raw = raws[u];
        unsigned w[4];
        if (rot180) {
            w[0] = __byte_perm(raw.w, 0, 0x1032); w[1] = __byte_perm(raw.z, 0, 0x1032);
            w[2] = __byte_perm(raw.y, 0, 0x1032); w[3] = __byte_perm(raw.x, 0, 0x1032);
        } else {
            w[0] = raw.x; w[1] = raw.y; w[2] = raw.z; w[3] = raw.w;
        }
        uint2 mb = make_uint2(0xffffffffu, 0xffffffffu);
        if (mask) { mb.x = __vcmpne4(mks[u].x, 0u); mb.y = __vcmpne4(mks[u].y, 0u); }
        unsigned v[4], u8lo = 0, u8hi = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned idx = __vsub2(__vmins2(__vmaxs2(w[k], lo2), hi2), lo2);
            const unsigned i0 = idx & 0xffffu, i1 = idx >> 16;
            const unsigned pair = (unsigned)lutT[i0] | ((unsigned)lutT[i1] << 16);
            const unsigned mw = k < 2 ? mb.x : mb.y;
            v[k] = pair & __byte_perm(mw, 0, (k & 1) ? 0x3322 : 0x1100);
            if (WANT_U8) {
                const unsigned two = ((unsigned)lut8[i0] | ((unsigned)lut8[i1] << 8)) << (16 * (k & 1));
                if (k < 2) u8lo |= two; else u8hi |= two;
            }
        }
        if (WANT_U8)
            st_stream_uint2(reinterpret_cast<uint2*>(out_u8_all + b * hw + o * 8), make_uint2(u8lo & mb.x, u8hi & mb.y));
        if (out_all) {
            if (NHWC) {
                unsigned t[12];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    t[3 * k] = __byte_perm(v[k], 0, 0x1010); t[3 * k + 1] = v[k]; t[3 * k + 2] = __byte_perm(v[k], 0, 0x3232);
                }
                if ((units_per_slice & 31) == 0) {
                    // the warp's 32 x 48 B are one contiguous 1536 B block: exchange through shared memory so
                    // that every store instruction writes 512 contiguous bytes (whole sectors, no partial writes)
                    const int lane = threadIdx.x & 31;
                    int4* stg = stage + (threadIdx.x >> 5) * 96;
                    stg[lane * 3] = make_int4(t[0], t[1], t[2], t[3]);
                    stg[lane * 3 + 1] = make_int4(t[4], t[5], t[6], t[7]);
                    stg[lane * 3 + 2] = make_int4(t[8], t[9], t[10], t[11]);
                    __syncwarp();
                    int4* dst = reinterpret_cast<int4*>(out_all + (b * hw + (long long)(o - lane) * 8) * 3);
                    st_stream_int4(dst + lane, stg[lane]);
                    st_stream_int4(dst + 32 + lane, stg[32 + lane]);
                    st_stream_int4(dst + 64 + lane, stg[64 + lane]);
                    __syncwarp();
                } else {                                            // ragged rows: plain 48-byte-strided stores
                    int4* dst = reinterpret_cast<int4*>(out_all + (b * hw + (long long)o * 8) * 3);
                    st_stream_int4(dst, make_int4(t[0], t[1], t[2], t[3]));
                    st_stream_int4(dst + 1, make_int4(t[4], t[5], t[6], t[7]));
                    st_stream_int4(dst + 2, make_int4(t[8], t[9], t[10], t[11]));
                }
            } else {
                uint16_t* base = out_all + b * 3 * hw + o * 8;
                const int4 q = make_int4(v[0], v[1], v[2], v[3]);
                st_stream_int4(reinterpret_cast<int4*>(base), q);
                st_stream_int4(reinterpret_cast<int4*>(base + hw), q);
                st_stream_int4(reinterpret_cast<int4*>(base + 2 * hw), q);
            }
        }
        }
    }
}

int launch_hu16(const int16_t* px, int B, int H, int W, int lo, int hi, int rot180, const uint8_t* mask, uint8_t* out_u8,
                void* out, int bf16, int nhwc, cudaStream_t s) {
    const int ups = H * W / 8;
    const int d = hi - lo;
    const size_t smem = (size_t)(d + 1) * 3 + 16 + (kThreads / 32) * 96 * sizeof(int4);
    const long long total = (long long)B * ups;
    eitb_prof_begin("hu_window_kernel", s);
    auto go = [&](auto kernel) {
        // one resident wave, and every CTA the same number of four-unit trips (a grid of exactly `cap` CTAs leaves a ragged
        // last trip: 17.3 units per thread = 4.3 trips on 160 slices)
        const long long need = (total + kThreads - 1) / kThreads;
        const long long cap = eitb_resident_ctas(kernel, kThreads, smem);
        const long long trips = (need + cap * 4 - 1) / (cap * 4);
        const long long grid = (need + trips * 4 - 1) / (trips * 4);
        kernel<<<(int)grid, kThreads, smem, s>>>(px, ups, total, lo, hi, rot180, mask, out_u8, (uint16_t*)out, bf16);
    };
    if (nhwc) {
        if (out_u8) go(hu_window16_kernel<true, true>);
        else go(hu_window16_kernel<true, false>);
    } else {
        if (out_u8) go(hu_window16_kernel<false, true>);
        else go(hu_window16_kernel<false, false>);
    }
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

template <typename T>
int launch_hu(const int16_t* px, int B, int H, int W, int lo, int hi, int rot180, const uint8_t* mask,
              uint8_t* out_u8, void* out_nchw, cudaStream_t s) {
    const long long n_units = (long long)B * H * W / 8;
    const int ups = H * W / 8;
    const int d = hi - lo;
    if (B > 65535) return EITB_ERR_UNSUPPORTED;
    const dim3 grid(eitb_grid_per_image(ups, kThreads, B), B);
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
                                   int out_dtype, int channels_last, eitb_stream_t stream) {
    if (!px || B < 0 || H <= 0 || W <= 0 || hi <= lo || lo < -32768 || hi > 32767) return EITB_ERR_BAD_ARG;
    if (!out_u8 && !out_nchw) return EITB_ERR_BAD_ARG;
    if ((W % 8) != 0) return EITB_ERR_UNSUPPORTED;
    if (B == 0) return EITB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if ((out_dtype == EITB_F16 || out_dtype == EITB_BF16 || !out_nchw) && hi - lo <= kMaxLut &&
        !(reinterpret_cast<uintptr_t>(px) & 15) && !(reinterpret_cast<uintptr_t>(out_nchw) & 15))
        return launch_hu16(px, B, H, W, lo, hi, rot180, body_mask, out_u8, out_nchw, out_dtype == EITB_BF16, channels_last, s);
    if (channels_last) return EITB_ERR_UNSUPPORTED;
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
