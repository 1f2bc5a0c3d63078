// K9: in-place per-channel bias + SiLU on channels-last activations -- the epilogue of every
// Conv(+folded BatchNorm)+SiLU block of the YOLO11s-seg network (ultralytics fuses BN into the
// convolution at predict time, exactly like the host mirror does; SiLU stays a separate pass in
// plain PyTorch).  One read and one write per activation instead of three of each.
// HBM-bound: 16-byte loads/stores, bias staged in shared memory as fp32.
#include "common.cuh"

namespace {

template <typename T> struct Vec8;
template <> struct Vec8<__half> {
    static __device__ __forceinline__ void unpack(const int4& v, float (&f)[8]) {
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
    static __device__ __forceinline__ int4 pack(const float (&f)[8]) {
        int4 v;
        __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        return v;
    }
};
template <> struct Vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void unpack(const int4& v, float (&f)[8]) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
    }
    static __device__ __forceinline__ int4 pack(const float (&f)[8]) {
        int4 v;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        return v;
    }
};

template <typename T, typename I>
__global__ void __launch_bounds__(256)
bias_act_kernel(T* __restrict__ x, const float* __restrict__ bias, I n_vec, int C, int act) {
    extern __shared__ float sb[];
    for (int c = threadIdx.x; c < C; c += blockDim.x) sb[c] = bias ? bias[c] : 0.f;
    __syncthreads();
    const unsigned vpc = (unsigned)C >> 3;                        // 8-element vectors per pixel
    const bool pow2 = (vpc & (vpc - 1)) == 0;
    int4* xv = reinterpret_cast<int4*>(x);
    constexpr int U = 4;                                           // 64 bytes in flight per thread
    const I stride = (I)gridDim.x * blockDim.x;
    for (I i0 = (I)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_vec; i0 += stride * U) {
        int4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (i0 + u * stride < n_vec) v[u] = xv[i0 + u * stride];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const I i = i0 + u * stride;
            if (i >= n_vec) break;
            const int c0 = (int)(pow2 ? ((unsigned)i & (vpc - 1)) : (unsigned)(i % vpc)) << 3;
            float f[8];
            Vec8<T>::unpack(v[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float a = f[j] + sb[c0 + j];
                f[j] = act ? __fdividef(a, 1.f + __expf(-a)) : a;   // SiLU
            }
            xv[i] = Vec8<T>::pack(f);
        }
    }
}

}  // namespace

extern "C" int eitb_bias_act_nhwc(void* x, int dtype, long long n_pixels, int C, const float* bias, int act,
                                  eitb_stream_t stream) {
    if (!x || n_pixels < 0 || C <= 0 || (act != 0 && act != 1)) return EITB_ERR_BAD_ARG;
    if ((C & 7) || (reinterpret_cast<uintptr_t>(x) & 15) || C > 8192) return EITB_ERR_UNSUPPORTED;
    if (n_pixels == 0) return EITB_OK;
    const long long n_vec = n_pixels * (C >> 3);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)C * sizeof(float);
    const bool small = n_vec < (1LL << 30);                        // 32-bit indexing whenever it fits
    const long long need = (n_vec + 256 * 4 - 1) / (256 * 4);      // 4 vectors per thread per trip
#define EITB_BIAS_ACT(T, I)                                                                               \
    do {                                                                                                  \
        const long long cap = eitb_resident_ctas(bias_act_kernel<T, I>, 256, smem);                       \
        bias_act_kernel<T, I><<<(int)(need < cap ? need : cap), 256, smem, s>>>((T*)x, bias, (I)n_vec, C, act); \
    } while (0)
    eitb_prof_begin("bias_act_kernel", s);
    switch (dtype) {
        case EITB_F16:
            if (small) EITB_BIAS_ACT(__half, unsigned); else EITB_BIAS_ACT(__half, long long);
            break;
        case EITB_BF16:
            if (small) EITB_BIAS_ACT(__nv_bfloat16, unsigned); else EITB_BIAS_ACT(__nv_bfloat16, long long);
            break;
        default: return EITB_ERR_UNSUPPORTED;
    }
#undef EITB_BIAS_ACT
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

// ---------------------------------------------------------------------------------------------
// General conv epilogue: y = act(src + bias) [+ residual] (the Bottleneck shortcut is added AFTER the activation); y goes to `out` (may alias src: in
// place) and/or into a channel slice of a wider channels-last tensor `out2` -- the concat buffer
// of a C3k2 / C3k block, so that torch.cat and the residual add never run as separate passes.
namespace {

template <typename T>
__global__ void __launch_bounds__(256)
conv_epilogue_kernel(const T* __restrict__ src, const float* __restrict__ bias, const T* __restrict__ residual,
                     T* __restrict__ out, T* __restrict__ out2, unsigned n_vec, int C, int act, int out2_C, int out2_off) {
    extern __shared__ float sb[];
    for (int c = threadIdx.x; c < C; c += blockDim.x) sb[c] = bias ? bias[c] : 0.f;
    __syncthreads();
    const unsigned vpc = (unsigned)C >> 3;
    const bool pow2 = (vpc & (vpc - 1)) == 0;
    const int sh = 31 - __clz(vpc);
    const int4* sv = reinterpret_cast<const int4*>(src);
    const int4* rv = reinterpret_cast<const int4*>(residual);
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
        const unsigned pix = pow2 ? i >> sh : i / vpc;
        const unsigned cv = i - pix * vpc;
        float f[8];
        Vec8<T>::unpack(sv[i], f);
        if (residual) {
            float r[8];
            Vec8<T>::unpack(rv[i], r);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float a = f[j] + sb[cv * 8 + j];
                f[j] = (act ? __fdividef(a, 1.f + __expf(-a)) : a) + r[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float a = f[j] + sb[cv * 8 + j];
                f[j] = act ? __fdividef(a, 1.f + __expf(-a)) : a;
            }
        }
        const int4 o = Vec8<T>::pack(f);
        if (out) reinterpret_cast<int4*>(out)[i] = o;
        if (out2) *reinterpret_cast<int4*>(out2 + (size_t)pix * out2_C + out2_off + cv * 8) = o;
    }
}

// nearest x2 upsample of a [B,h,w,Ca] + copy of b [B,2h,2w,Cb] -> out [B,2h,2w,Ca+Cb] (channels-last)
template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_concat_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int B, int h, int w, int Ca, int Cb) {
    const int C = Ca + Cb, vpc = C >> 3, va = Ca >> 3;
    const long long n_vec = (long long)B * 4 * h * w * vpc;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % vpc);
        const long long pix = i / vpc;
        const int x = (int)(pix % (2 * w));
        const long long t = pix / (2 * w);
        const int y = (int)(t % (2 * h)), bi = (int)(t / (2 * h));
        int4 v;
        if (cv < va) v = reinterpret_cast<const int4*>(a + (((long long)bi * h + (y >> 1)) * w + (x >> 1)) * Ca)[cv];
        else v = reinterpret_cast<const int4*>(b + pix * Cb)[cv - va];
        reinterpret_cast<int4*>(out)[i] = v;
    }
}

}  // namespace

extern "C" int eitb_conv_epilogue_nhwc(const void* src, int dtype, long long n_pixels, int C, const float* bias, int act,
                                       const void* residual, void* out, void* out2, int out2_C, int out2_off,
                                       eitb_stream_t stream) {
    if (!src || n_pixels < 0 || C <= 0 || (act != 0 && act != 1) || (!out && !out2)) return EITB_ERR_BAD_ARG;
    if ((C & 7) || C > 8192 || (out2 && ((out2_C & 7) || (out2_off & 7) || out2_off + C > out2_C))) return EITB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(out2) |
         reinterpret_cast<uintptr_t>(residual)) & 15)
        return EITB_ERR_UNSUPPORTED;
    const long long n_vec = n_pixels * (C >> 3);
    if (n_vec >= (1LL << 31)) return EITB_ERR_UNSUPPORTED;
    if (n_pixels == 0) return EITB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)C * sizeof(float);
    const long long need = (n_vec + 255) / 256;
    eitb_prof_begin("conv_epilogue_kernel", s);
    if (dtype == EITB_F16) {
        const long long cap = eitb_resident_ctas(conv_epilogue_kernel<__half>, 256, smem);
        conv_epilogue_kernel<__half><<<(int)(need < cap ? need : cap), 256, smem, s>>>(
            (const __half*)src, bias, (const __half*)residual, (__half*)out, (__half*)out2, (unsigned)n_vec, C, act, out2_C, out2_off);
    } else if (dtype == EITB_BF16) {
        const long long cap = eitb_resident_ctas(conv_epilogue_kernel<__nv_bfloat16>, 256, smem);
        conv_epilogue_kernel<__nv_bfloat16><<<(int)(need < cap ? need : cap), 256, smem, s>>>(
            (const __nv_bfloat16*)src, bias, (const __nv_bfloat16*)residual, (__nv_bfloat16*)out, (__nv_bfloat16*)out2,
            (unsigned)n_vec, C, act, out2_C, out2_off);
    } else {
        return EITB_ERR_UNSUPPORTED;
    }
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

extern "C" int eitb_upsample2x_concat_nhwc(const void* a, const void* b, void* out, int dtype, int B, int h, int w, int Ca,
                                           int Cb, eitb_stream_t stream) {
    if (!a || !b || !out || B < 0 || h <= 0 || w <= 0 || Ca <= 0 || Cb <= 0) return EITB_ERR_BAD_ARG;
    if ((Ca & 7) || (Cb & 7) || (dtype != EITB_F16 && dtype != EITB_BF16)) return EITB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15) return EITB_ERR_UNSUPPORTED;
    if (B == 0) return EITB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const long long n_vec = (long long)B * 4 * h * w * ((Ca + Cb) >> 3);
    eitb_prof_begin("upsample2x_concat_kernel", s);
    // both 16-bit types move as raw 16-byte vectors
    upsample2x_concat_kernel<__half><<<eitb_grid(n_vec, 256, 8), 256, 0, s>>>((const __half*)a, (const __half*)b, (__half*)out, B, h, w, Ca, Cb);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}
