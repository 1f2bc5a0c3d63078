#include "common.cuh"

extern "C" const char* eitb_strerror(int code) {
    switch (code) {
        case EITB_OK: return "ok";
        case EITB_ERR_BAD_ARG: return "bad argument";
        case EITB_ERR_WORKSPACE: return "workspace too small";
        case EITB_ERR_LAUNCH: return "CUDA launch failed";
        case EITB_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown error";
    }
}
extern "C" int eitb_version(void) { return 100; }

__global__ void codes_to_bgr_kernel(const uint8_t* __restrict__ code, uint8_t* __restrict__ bgr, int64_t n) {
    // 4 codes -> 12 bytes per thread
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (; i < n; i += stride) {
        if (i + 4 <= n) {
            uint32_t c = *reinterpret_cast<const uint32_t*>(code + i);
            uint32_t w[3] = {0, 0, 0};
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                int px = k / 3, ch = k % 3;                 // ch 0:B (bit2) 1:G (bit1) 2:R (bit0)
                uint32_t v = ((c >> (8 * px)) >> (2 - ch)) & 1u ? 255u : 0u;
                w[k / 4] |= v << (8 * (k % 4));
            }
            uint32_t* o = reinterpret_cast<uint32_t*>(bgr + i * 3);
            o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
        } else {
            for (int64_t j = i; j < n; ++j) {
                uint8_t c = code[j];
                bgr[j * 3 + 0] = (c & 4) ? 255 : 0;
                bgr[j * 3 + 1] = (c & 2) ? 255 : 0;
                bgr[j * 3 + 2] = (c & 1) ? 255 : 0;
            }
        }
    }
}

extern "C" int eitb_codes_to_bgr(const uint8_t* code, uint8_t* bgr, int64_t n, eitb_stream_t stream) {
    if (!code || !bgr || n < 0) return EITB_ERR_BAD_ARG;
    if (n == 0) return EITB_OK;
    if ((reinterpret_cast<uintptr_t>(code) & 3) || (reinterpret_cast<uintptr_t>(bgr) & 3)) return EITB_ERR_BAD_ARG;
    codes_to_bgr_kernel<<<eitb_grid((n + 3) / 4, 256, 8), 256, 0, (cudaStream_t)stream>>>(code, bgr, n);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}
