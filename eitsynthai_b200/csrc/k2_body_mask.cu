// K2: body mask.
//
// Reference: get_axial_slice_body_mask (kt_service/ai_tools/utils.py:526-585) and its NIfTI twin
// (:588-618): flipud (:551), HU = slope*px + intercept cast to int16 (:554-559), -500 < HU < 1000
// (:565), 5x5 MORPH_OPEN (:569), external contours (:572), the one with the largest
// cv2.contourArea (:577) filled with 255 (:581-582).
//
// Restated without contours (oracle/imaging.py largest_contour_fill_np, pinned to the reference
// on the golden vectors):
//   * pixels outside every external contour = 4-connected background reachable from the frame;
//   * every 8-connected component of the rest is one filled external contour;
//   * its contourArea is N4 + N3/2 over the 2x2 pixel blocks with 4 / exactly 3 filled pixels;
//   * ties go to the component whose first pixel comes last in raster order.
// Stages: threshold+erode -> dilate -> CC(background, 4, frame-linked) -> CC(filled, 8) ->
// per-component 2*area accumulation -> arg-max -> write mask.
#include "cc.cuh"
#include "bitflood.cuh"

namespace {

using namespace eitb_cc;

__device__ __forceinline__ bool hu_in_range(int px, int slope, int intercept) {
    const int hu = (int)(short)(slope * px + intercept);          // .astype(int16) wraps (utils.py:559)
    return hu > -500 && hu < 1000;
}

// erosion of the thresholded, vertically flipped slice with a 5x5 box; cv2's default border
// for erosion never removes pixels (outside counts as set).
__global__ void __launch_bounds__(256)
thr_erode_kernel(const int16_t* __restrict__ px, int B, int H, int W, int slope, int intercept, int flipud,
                 uint8_t* __restrict__ er) {
    const long long n = (long long)B * H * W;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(t / ((long long)H * W));
        const int r = (int)(t - (long long)b * H * W);
        const int y = r / W, x = r - y * W;
        const int16_t* img = px + (long long)b * H * W;
        bool all = true;
        for (int dy = -2; dy <= 2 && all; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
            const int sy = flipud ? H - 1 - yy : yy;
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                all = all && hu_in_range(img[(long long)sy * W + xx], slope, intercept);
            }
        }
        er[t] = all ? 1 : 0;
    }
}

// dilation with a 5x5 box (outside counts as unset) -> opened mask (1/0)
__global__ void __launch_bounds__(256)
dilate_kernel(const uint8_t* __restrict__ er, int B, int H, int W, uint8_t* __restrict__ opened) {
    const long long n = (long long)B * H * W;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(t / ((long long)H * W));
        const int r = (int)(t - (long long)b * H * W);
        const int y = r / W, x = r - y * W;
        const uint8_t* img = er + (long long)b * H * W;
        bool any = false;
        for (int dy = -2; dy <= 2 && !any; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= W) continue;
                any = any || img[(long long)yy * W + xx];
            }
        }
        opened[t] = any ? 1 : 0;
    }
}

// ---- bit-image path (W % 32 == 0): one bit per pixel, 32 pixels per word, LSB = lowest x
__global__ void __launch_bounds__(256)
thr_bits_kernel(const int16_t* __restrict__ px, long long n_units, int H, int W, int slope, int intercept, int flipud,
                uint32_t* __restrict__ bits) {
    const int upr = W >> 3;                                       // 8-pixel units per row
    const long long nround = (n_units + 31) & ~31LL;              // keep whole warps in the shuffles below
    for (long long uu = (long long)blockIdx.x * blockDim.x + threadIdx.x; uu < nround; uu += (long long)gridDim.x * blockDim.x) {
        const bool active = uu < n_units;
        const long long u = active ? uu : n_units - 1;
        const long long rowid = u / upr;                          // b * H + y (output orientation)
        const int c = (int)(u - rowid * upr);
        const long long b = rowid / H;
        const int y = (int)(rowid - b * H);
        const int sy = flipud ? H - 1 - y : y;
        const int4 raw = ld_stream_int4(reinterpret_cast<const int4*>(px + ((b * H + sy) * (long long)W + c * 8)));
        const int w[4] = {raw.x, raw.y, raw.z, raw.w};
        unsigned v = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int p = (int)(short)((w[j >> 1] >> ((j & 1) * 16)) & 0xffff);
            v |= (hu_in_range(p, slope, intercept) ? 1u : 0u) << j;
        }
        v <<= 8 * (threadIdx.x & 3);                               // 4 neighbouring threads make one word
        v |= __shfl_xor_sync(0xffffffffu, v, 1);
        v |= __shfl_xor_sync(0xffffffffu, v, 2);
        if (active && (threadIdx.x & 3) == 0) bits[u >> 2] = v;
    }
}

// 5x5 box erosion (ERODE: outside the image counts as set) or dilation (outside counts as unset)
template <bool ERODE>
__global__ void __launch_bounds__(256)
morph5_bits_kernel(const uint32_t* __restrict__ in, long long n_words, int H, int wpr, uint32_t* __restrict__ out) {
    const uint32_t fill = ERODE ? 0xffffffffu : 0u;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_words; t += (long long)gridDim.x * blockDim.x) {
        const long long rowid = t / wpr;
        const int wx = (int)(t - rowid * wpr);
        const int y = (int)(rowid % H);
        uint32_t acc = fill;
#pragma unroll
        for (int dy = -2; dy <= 2; ++dy) {
            if (y + dy < 0 || y + dy >= H) continue;
            const uint32_t* r = in + (rowid + dy) * wpr;
            const uint32_t C = r[wx], L = wx > 0 ? r[wx - 1] : fill, R = wx + 1 < wpr ? r[wx + 1] : fill;
            const uint32_t m1 = (C << 1) | (L >> 31), m2 = (C << 2) | (L >> 30);
            const uint32_t p1 = (C >> 1) | (R << 31), p2 = (C >> 2) | (R << 30);
            if (ERODE) acc &= C & m1 & m2 & p1 & p2; else acc |= C | m1 | m2 | p1 | p2;
        }
        out[t] = acc;
    }
}

// 2*contourArea per component: +2 for every full 2x2 block, +1 for every block with 3 pixels.
// Blocks are anchored at (y, x) = top-left pixel, y in [-1, H-1], x in [-1, W-1].
__global__ void __launch_bounds__(256)
area_kernel(const int32_t* __restrict__ lab, int B, int H, int W, int32_t* __restrict__ area2) {
    // grid (x: block-row chunks, y: block row + 1, z: image); one thread per 2x2 block anchor
    const int b = blockIdx.z, y = (int)blockIdx.y - 1;
    const int bw = W + 1;
    const int nround = (bw + 31) & ~31;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nround; t += gridDim.x * blockDim.x) {
        int root = -1, add = 0;
        if (t < bw) {
            const int x = t - 1;
            const int32_t* L = lab + (long long)b * H * W;
            int cnt = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int yy = y + (k >> 1), xx = x + (k & 1);
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                    const int l = L[yy * W + xx];
                    if (l >= 0) { ++cnt; root = l; }
                }
            }
            add = cnt == 4 ? 2 : cnt == 3 ? 1 : 0;
            if (!add) root = -1;
        }
        // warp-aggregate by (image, root)
        const unsigned grp = __match_any_sync(0xffffffffu, root);
        const int sum = __reduce_add_sync(grp, add);
        if (root >= 0 && (int)(__ffs(grp) - 1) == (int)(threadIdx.x & 31)) atomicAdd(area2 + (long long)b * H * W + root, sum);
    }
}

// best[b] = max over roots of (area2 << 32 | root)
__global__ void __launch_bounds__(256)
best_kernel(const int32_t* __restrict__ lab, const int32_t* __restrict__ area2, int hw,
            long long* __restrict__ best) {                  // grid (x: pixel blocks, y: image)
    const long long off = (long long)blockIdx.y * hw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x)
        if (lab[off + i] == i) atomicMax(best + blockIdx.y, ((long long)area2[off + i] << 32) | (long long)i);
}

__global__ void __launch_bounds__(256)
write_mask_kernel(const int32_t* __restrict__ lab, const long long* __restrict__ best, int hw,
                  uint8_t* __restrict__ mask) {              // grid (x: 4-pixel blocks, y: image); hw % 4 == 0
    const long long off = (long long)blockIdx.y * hw;
    const long long k = best[blockIdx.y];
    const int root = k >= 0 ? (int)(k & 0xffffffffLL) : -3;     // -3 matches no label
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < hw; i += gridDim.x * blockDim.x * 4) {
        const int4 l = *reinterpret_cast<const int4*>(lab + off + i);
        const uint32_t v = (l.x == root ? 0xffu : 0u) | (l.y == root ? 0xff00u : 0u) | (l.z == root ? 0xff0000u : 0u) |
                           (l.w == root ? 0xff000000u : 0u);
        *reinterpret_cast<uint32_t*>(mask + off + i) = v;
    }
}

__global__ void __launch_bounds__(256)
write_mask_scalar_kernel(const int32_t* __restrict__ lab, const long long* __restrict__ best, int hw,
                         uint8_t* __restrict__ mask) {
    const long long off = (long long)blockIdx.y * hw;
    const long long k = best[blockIdx.y];
    const int root = k >= 0 ? (int)(k & 0xffffffffLL) : -3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x)
        mask[off + i] = lab[off + i] == root ? 255 : 0;
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t eitb_body_mask_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    const size_t n = (size_t)B * H * W;
    // er u8, opened u8, labels A int32, labels B int32, area2 int32, best int64[B]
    return align256(n) * 2 + align256(n * 4) * 3 + align256((size_t)B * 8);
}

extern "C" int eitb_body_mask(const int16_t* px, int B, int H, int W, int slope, int intercept, int flipud,
                              uint8_t* mask, void* ws, size_t ws_bytes, eitb_stream_t stream) {
    if (!px || !mask || B < 0 || H <= 0 || W <= 0) return EITB_ERR_BAD_ARG;
    if (B == 0) return EITB_OK;
    if ((long long)H * W >= (1LL << 31)) return EITB_ERR_UNSUPPORTED;
    if (!ws || ws_bytes < eitb_body_mask_workspace_bytes(B, H, W)) return EITB_ERR_WORKSPACE;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)B * H * W;
    char* p = reinterpret_cast<char*>(ws);
    uint8_t* er = reinterpret_cast<uint8_t*>(p); p += align256(n);
    uint8_t* opened = reinterpret_cast<uint8_t*>(p); p += align256(n);
    int32_t* labA = reinterpret_cast<int32_t*>(p); p += align256(n * 4);
    int32_t* labB = reinterpret_cast<int32_t*>(p); p += align256(n * 4);
    int32_t* area2 = reinterpret_cast<int32_t*>(p); p += align256(n * 4);
    long long* best = reinterpret_cast<long long*>(p);
    const int grid = eitb_grid((long long)n, 256, 8);

    int rc;
    if ((W & 31) == 0 && !(reinterpret_cast<uintptr_t>(px) & 15)) {
        // bit images: threshold -> erode -> dilate touch 1/16 of the bytes of the u8 path
        uint32_t* b0 = reinterpret_cast<uint32_t*>(er);
        uint32_t* b1 = reinterpret_cast<uint32_t*>(opened);
        const long long n_words = (long long)n / 32;
        eitb_prof_begin("thr_bits_kernel", s);
        thr_bits_kernel<<<eitb_grid((long long)n / 8, 256, 8), 256, 0, s>>>(px, (long long)n / 8, H, W, slope, intercept, flipud, b0);
        EITB_CHECK_LAUNCH();
        eitb_prof_begin("morph5_bits_kernel", s);
        morph5_bits_kernel<true><<<eitb_grid(n_words, 256, 8), 256, 0, s>>>(b0, n_words, H, W / 32, b1);
        EITB_CHECK_LAUNCH();
        eitb_prof_begin("morph5_bits_kernel", s);
        morph5_bits_kernel<false><<<eitb_grid(n_words, 256, 8), 256, 0, s>>>(b1, n_words, H, W / 32, b0);
        EITB_CHECK_LAUNCH();
        if (eitb_flood::flood_supported(H, W)) {
            // background reachable from the frame = outside every external contour: bit-parallel flood, then the
            // 8-connected components of everything else straight from the bit image
            uint32_t* outside = b1;
            rc = eitb_flood::frame_flood<eitb_flood::SRC_BITS_ZERO>(b0, B, H, W, 1, 0, 0, 0, nullptr, outside, s);
            if (rc != EITB_OK) return rc;
            rc = cc_label<PRED_BIT_ZERO, 8>(outside, (size_t)H * W / 8, 0, B, H, W, 0, labB, s);
            goto labelled;
        }
        rc = cc_label<PRED_BIT_ZERO, 4>(b0, (size_t)H * W / 8, 0, B, H, W, 1, labA, s);         // background, frame-linked
    } else {
        eitb_prof_begin("thr_erode_kernel", s);
        thr_erode_kernel<<<grid, 256, 0, s>>>(px, B, H, W, slope, intercept, flipud, er);
        EITB_CHECK_LAUNCH();
        eitb_prof_begin("dilate_kernel", s);
        dilate_kernel<<<grid, 256, 0, s>>>(er, B, H, W, opened);
        EITB_CHECK_LAUNCH();
        rc = cc_label<PRED_U8_ZERO, 4>(opened, (size_t)H * W, 0, B, H, W, 1, labA, s);
    }
    if (rc != EITB_OK) return rc;
    rc = cc_label<PRED_LABEL_NOT_OUT, 8>(labA, (size_t)H * W * 4, 0, B, H, W, 0, labB, s);   // filled regions
labelled:
    if (rc != EITB_OK) return rc;
    if (cudaMemsetAsync(area2, 0, n * 4, s) != cudaSuccess) return EITB_ERR_LAUNCH;
    if (cudaMemsetAsync(best, 0xff, (size_t)B * 8, s) != cudaSuccess) return EITB_ERR_LAUNCH;  // -1
    eitb_prof_begin("area_kernel", s);
    if (H + 1 > 65535 || B > 65535) return EITB_ERR_UNSUPPORTED;
    area_kernel<<<dim3(eitb_div_up(W + 1, 256), H + 1, B), 256, 0, s>>>(labB, B, H, W, area2);
    EITB_CHECK_LAUNCH();
    eitb_prof_begin("best_kernel", s);
    best_kernel<<<dim3(eitb_grid_per_image((long long)H * W, 256, B), B), 256, 0, s>>>(labB, area2, H * W, best);
    EITB_CHECK_LAUNCH();
    eitb_prof_begin("write_mask_kernel", s);
    if (((H * W) & 3) == 0 && !(reinterpret_cast<uintptr_t>(mask) & 3))
        write_mask_kernel<<<dim3(eitb_grid_per_image((long long)H * W / 4, 256, B), B), 256, 0, s>>>(labB, best, H * W, mask);
    else
        write_mask_scalar_kernel<<<dim3(eitb_grid_per_image((long long)H * W, 256, B), B), 256, 0, s>>>(labB, best, H * W, mask);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

extern "C" int eitb_cc_label(const uint8_t* mask, int B, int H, int W, int connectivity, int link_outside,
                             int32_t* labels, eitb_stream_t stream) {
    if (!mask || !labels || B < 0 || H <= 0 || W <= 0 || (connectivity != 4 && connectivity != 8)) return EITB_ERR_BAD_ARG;
    if (B == 0) return EITB_OK;
    if ((long long)H * W >= (1LL << 31)) return EITB_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    return connectivity == 4 ? cc_label<PRED_U8_NONZERO, 4>(mask, (size_t)H * W, 0, B, H, W, link_outside, labels, s)
                             : cc_label<PRED_U8_NONZERO, 8>(mask, (size_t)H * W, 0, B, H, W, link_outside, labels, s);
}
