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
