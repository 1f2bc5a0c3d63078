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

// ---------------------------------------------------------------------------------------------
// Launch profiler: when enabled, every kernel launch of the library is bracketed by two CUDA
// events on its own stream; eitb_profile_report() aggregates them per kernel name.  Meant for
// bench.py's roofline line (timing "live, with CUDA events on the launching stream"); not
// usable while a CUDA graph is being captured.
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <cstdio>
#include <cstring>

namespace {
struct ProfRec { const char* name; cudaEvent_t a, b; };
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_prof_recs;
thread_local ProfRec t_pending = {nullptr, nullptr, nullptr};
thread_local cudaStream_t t_stream = nullptr;
}  // namespace

void eitb_prof_begin(const char* kernel_name, cudaStream_t stream) {
    if (!g_prof_on) return;
    ProfRec r{kernel_name, nullptr, nullptr};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, stream);
    t_pending = r;
    t_stream = stream;
}

void eitb_prof_end() {
    if (!t_pending.name) return;
    cudaEventRecord(t_pending.b, t_stream);
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        g_prof_recs.push_back(t_pending);
    }
    t_pending.name = nullptr;
}

extern "C" int eitb_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof_recs.clear();
    g_prof_on = on != 0;
    return EITB_OK;
}

extern "C" long long eitb_profile_report(char* buf, size_t cap) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<long long, double>> agg;
    for (auto& r : g_prof_recs) {
        if (cudaEventSynchronize(r.b) != cudaSuccess) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
        auto& e = agg[r.name];
        e.first += 1;
        e.second += ms;
    }
    std::string out;
    char line[256];
    for (auto& kv : agg) {
        snprintf(line, sizeof line, "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (buf && cap > 0) {
        const size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return (long long)out.size() + 1;
}

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
    eitb_prof_begin("codes_to_bgr_kernel", (cudaStream_t)stream);
    codes_to_bgr_kernel<<<eitb_grid((n + 3) / 4, 256, 8), 256, 0, (cudaStream_t)stream>>>(code, bgr, n);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

// ---------------------------------------------------------------------------------------------
// The per-function entry points of the mirror (callers that use the reference's helpers one by one, with numpy images
// in between): the same pixel rules as the fused kernels, as library calls instead of host-framework arithmetic.

// cv2.bitwise_and(img, img, mask=m) of ai_tools.py:212: out = mask != 0 ? img : 0   (K1 fuses this on the batched path)
__global__ void apply_mask_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask, int64_t n, int ch,
                                  uint8_t* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * ch; i += stride) out[i] = mask[i / ch] ? img[i] : 0;
}

extern "C" int eitb_apply_mask_u8(const uint8_t* img, const uint8_t* mask, int64_t n_px, int channels, uint8_t* out,
                                  eitb_stream_t stream) {
    if (!img || !mask || !out || n_px < 0 || channels < 1 || channels > 4) return EITB_ERR_BAD_ARG;
    if (n_px == 0) return EITB_OK;
    eitb_prof_begin("apply_mask_kernel", (cudaStream_t)stream);
    apply_mask_kernel<<<eitb_grid(n_px * channels, 256, 8), 256, 0, (cudaStream_t)stream>>>(img, mask, n_px, channels, out);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

// create_segmentations_masks (utils.py:437-523): per class the union of its instances' masks (> 0) painted in the class
// colour.  masks [n, S*S] fp32, cls [n] int32 (0 bone, 1 muscles, 2 lung, 3 adipose; others ignored) -> bgr4 [4, S*S, 3] u8.
__global__ void class_images_kernel(const float* __restrict__ masks, const int32_t* __restrict__ cls, int n, int64_t px,
                                    uint8_t* __restrict__ bgr4) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < px; p += stride) {
        unsigned hit = 0;
        for (int i = 0; i < n; ++i) {
            const int c = cls[i];
            if (c >= 0 && c < 4 && masks[(int64_t)i * px + p] > 0.f) hit |= 1u << c;
        }
        // colours of utils.py:468-473 (BGR): bone white, muscles red, lung cyan-ish (255,255,0), adipose yellow (0,255,255)
        const uint8_t col[4][3] = {{255, 255, 255}, {0, 0, 255}, {255, 255, 0}, {0, 255, 255}};
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int k = 0; k < 3; ++k) bgr4[((int64_t)c * px + p) * 3 + k] = (hit >> c) & 1u ? col[c][k] : 0;
    }
}

extern "C" int eitb_class_images(const float* masks, const int32_t* cls, int n, int64_t n_px, uint8_t* bgr4, eitb_stream_t stream) {
    if (!bgr4 || n < 0 || n_px < 0 || (n > 0 && (!masks || !cls))) return EITB_ERR_BAD_ARG;
    if (n_px == 0) return EITB_OK;
    eitb_prof_begin("class_images_kernel", (cudaStream_t)stream);
    class_images_kernel<<<eitb_grid(n_px, 256, 8), 256, 0, (cudaStream_t)stream>>>(masks, cls, n, n_px, bgr4);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

// overlay_segmentation_masks (utils.py:395-434) of one class image: code |= value wherever any BGR channel is non-zero
// (the saturating colour adds of the four class images are the OR of their 3-bit colour codes).
__global__ void bgr_or_code_kernel(const uint8_t* __restrict__ bgr, int64_t px, int value, uint8_t* __restrict__ code) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < px; p += stride)
        if (bgr[p * 3] | bgr[p * 3 + 1] | bgr[p * 3 + 2]) code[p] |= (uint8_t)value;
}

extern "C" int eitb_bgr_or_code(const uint8_t* bgr, int64_t n_px, int value, uint8_t* code, eitb_stream_t stream) {
    if (!bgr || !code || n_px < 0 || value < 0 || value > 7) return EITB_ERR_BAD_ARG;
    if (n_px == 0) return EITB_OK;
    eitb_prof_begin("bgr_or_code_kernel", (cudaStream_t)stream);
    bgr_or_code_kernel<<<eitb_grid(n_px, 256, 8), 256, 0, (cudaStream_t)stream>>>(bgr, n_px, value, code);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

// ---------------------------------------------------------------------------------------------
// Host path of K3: the coronal image needs one row of every slice.  Ship exactly those rows
// from the pinned host series with one strided DMA (no CPU gather, no staging copy): row `row` of each of
// n_slices [H,W] int16 slices -> dev_rows [n_slices, W].
extern "C" int eitb_rows_h2d(const int16_t* host_px, long long n_slices, int H, int W, int row, int16_t* dev_rows,
                             eitb_stream_t stream) {
    if (!host_px || !dev_rows || n_slices < 0 || H <= 0 || W <= 0 || row < 0 || row >= H) return EITB_ERR_BAD_ARG;
    if (n_slices == 0) return EITB_OK;
    const cudaError_t e = cudaMemcpy2DAsync(dev_rows, (size_t)W * 2, host_px + (size_t)row * W, (size_t)H * W * 2, (size_t)W * 2,
                                            (size_t)n_slices, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    return e == cudaSuccess ? EITB_OK : EITB_ERR_LAUNCH;
}
