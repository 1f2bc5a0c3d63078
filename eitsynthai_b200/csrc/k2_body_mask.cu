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
area_kernel(const int32_t* __restrict__ lab, int B, int H, int W, int32_t* __restrict__ area2, const int* __restrict__ skip) {
    // grid (x: block-row chunks, y: block rows (strided), z: image); one thread per 2x2 block anchor
    const int b = blockIdx.z;
    if (skip && skip[b]) return;
    const int bw = W + 1;
    const int nround = (bw + 31) & ~31;
    for (int y = (int)blockIdx.y - 1; y < H; y += gridDim.y)
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
            long long* __restrict__ best, const int* __restrict__ skip) {   // grid (x: pixel blocks, y: image)
    if (skip && skip[blockIdx.y]) return;
    const long long off = (long long)blockIdx.y * hw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x)
        if (lab[off + i] == i) atomicMax(best + blockIdx.y, ((long long)area2[off + i] << 32) | (long long)i);
}

__global__ void __launch_bounds__(256)
write_mask_kernel(const int32_t* __restrict__ lab, const long long* __restrict__ best, int hw,
                  uint8_t* __restrict__ mask, const int* __restrict__ skip) {   // grid (x: 4-pixel blocks, y: image); hw % 4 == 0
    if (skip && skip[blockIdx.y]) return;
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
                         uint8_t* __restrict__ mask, const int* __restrict__ skip) {
    if (skip && skip[blockIdx.y]) return;
    const long long off = (long long)blockIdx.y * hw;
    const long long k = best[blockIdx.y];
    const int root = k >= 0 ? (int)(k & 0xffffffffLL) : -3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x)
        mask[off + i] = lab[off + i] == root ? 255 : 0;
}

// ---- fast path: the dominant component, entirely on bit images.
// The filled regions (everything that is not frame-connected background) of a CT slice are one big body plus a few
// small islands (table, tubes).  A component whose 2 x contourArea exceeds half of the sum over ALL components is the
// strict arg-max, whatever the others are -- so: total = block count over the whole filled image, flood the component
// under the pixel nearest the image centre (8-connected, bit-parallel, same row machinery as bitflood.cuh), count its
// blocks, and if 2 * component > total write the mask from the component's bits and mark the image done.  Images that
// fail the test (no dominant body) fall through to the general connected-component path below, which skips done ones.
// 2 x contourArea = 2 * (full 2x2 blocks) + (blocks with exactly three pixels), anchors x in [0, W-2], y in [0, H-2].
__device__ __forceinline__ int block_area2(uint32_t A, uint32_t B, uint32_t An, uint32_t Bn) {
    const uint32_t a1 = (A >> 1) | (An << 31), b1 = (B >> 1) | (Bn << 31);           // pixel x + 1 at bit x
    const uint32_t full = A & a1 & B & b1;
    const uint32_t three = (A & a1 & B & ~b1) | (A & a1 & ~B & b1) | (A & ~a1 & B & b1) | (~A & a1 & B & b1);
    return 2 * __popc(full) + __popc(three);
}

__global__ void __launch_bounds__(eitb_flood::kWarps * 32)
body_dominant_kernel(const uint32_t* __restrict__ outside, int H, int W, uint8_t* __restrict__ mask, int* __restrict__ done) {
    using namespace eitb_flood;
    extern __shared__ uint32_t dsm[];
    const int wpr = W >> 5, b = blockIdx.x;
    uint32_t* filled = dsm;
    uint32_t* comp = dsm + (size_t)H * wpr;
    __shared__ int s_total, s_comp, s_seed;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lane_mask = wpr >= 32 ? 0xffffffffu : ((1u << wpr) - 1u);
    const uint32_t* src = outside + (size_t)b * H * wpr;
    for (int i = threadIdx.x; i < H * wpr; i += kWarps * 32) { filled[i] = ~src[i]; comp[i] = 0u; }
    if (threadIdx.x == 0) { s_total = 0; s_comp = 0; s_seed = -1; }
    __syncthreads();
    auto area2_of = [&](const uint32_t* X) -> int {
        int sum = 0;
        for (int y = warp; y + 1 < H; y += kWarps) {
            const uint32_t A = lane < wpr ? X[y * wpr + lane] : 0u, Bv = lane < wpr ? X[(y + 1) * wpr + lane] : 0u;
            uint32_t An = __shfl_down_sync(0xffffffffu, A, 1), Bn = __shfl_down_sync(0xffffffffu, Bv, 1);
            if (lane + 1 >= wpr) { An = 0u; Bn = 0u; }
            sum += block_area2(A, Bv, An & 1u, Bn & 1u);
        }
        return warp_sum(sum);
    };
    {
        const int t = area2_of(filled);
        if (lane == 0) atomicAdd(&s_total, t);
    }
    // seed: the filled pixel of the centre row nearest the centre column
    if (threadIdx.x == 0) {
        const int yc = H >> 1, xc = W >> 1;
        for (int d = 0; d < W && s_seed < 0; ++d) {
            const int xa = xc + d, xb = xc - d;
            if (xa < W && ((filled[yc * wpr + (xa >> 5)] >> (xa & 31)) & 1u)) s_seed = xa;
            else if (xb >= 0 && ((filled[yc * wpr + (xb >> 5)] >> (xb & 31)) & 1u)) s_seed = xb;
        }
    }
    __syncthreads();
    if (s_seed < 0) { if (threadIdx.x == 0) done[b] = 0; return; }
    if (warp == 0) {
        const int yc = H >> 1;
        const uint32_t a = lane < wpr ? filled[yc * wpr + lane] : 0u;
        const uint32_t sd = lane == (s_seed >> 5) ? 1u << (s_seed & 31) : 0u;
        const uint32_t r = hfill(a, sd, lane, wpr, lane_mask);
        if (lane < wpr) comp[yc * wpr + lane] = r;
    }
    __syncthreads();
    // 8-connected flood: a row receives its neighbours' bits and their left / right shifts, then extends along its runs
    const int band = (H + kWarps - 1) / kWarps;
    const int y_lo = warp * band, y_hi = min(y_lo + band, H);
    auto spread = [&](uint32_t p) -> uint32_t {
        uint32_t up = __shfl_up_sync(0xffffffffu, p, 1), dn = __shfl_down_sync(0xffffffffu, p, 1);
        if (lane == 0) up = 0u;
        if (lane + 1 >= wpr) dn = 0u;
        return p | (p << 1) | (up >> 31) | (p >> 1) | (dn << 31);
    };
    for (;;) {
        bool changed = false;
        uint32_t prev = (y_lo > 0 && lane < wpr) ? comp[(y_lo - 1) * wpr + lane] : 0u;
        for (int y = y_lo; y < y_hi; ++y) {
            const uint32_t a = lane < wpr ? filled[y * wpr + lane] : 0u;
            const uint32_t cur = lane < wpr ? comp[y * wpr + lane] : 0u;
            const uint32_t sd = cur | (spread(prev) & a);
            uint32_t r = cur;
            if (__any_sync(0xffffffffu, sd != cur)) {
                r = hfill(a, sd, lane, wpr, lane_mask);
                if (lane < wpr) comp[y * wpr + lane] = r;
                changed = true;
            }
            prev = r;
        }
        prev = (y_hi < H && lane < wpr) ? comp[y_hi * wpr + lane] : 0u;
        for (int y = y_hi - 1; y >= y_lo; --y) {
            const uint32_t a = lane < wpr ? filled[y * wpr + lane] : 0u;
            const uint32_t cur = lane < wpr ? comp[y * wpr + lane] : 0u;
            const uint32_t sd = cur | (spread(prev) & a);
            uint32_t r = cur;
            if (__any_sync(0xffffffffu, sd != cur)) {
                r = hfill(a, sd, lane, wpr, lane_mask);
                if (lane < wpr) comp[y * wpr + lane] = r;
                changed = true;
            }
            prev = r;
        }
        if (!__syncthreads_or(changed ? 1 : 0)) break;
    }
    {
        const int t = area2_of(comp);
        if (lane == 0) atomicAdd(&s_comp, t);
    }
    __syncthreads();
    const bool win = 2LL * s_comp > (long long)s_total;
    if (threadIdx.x == 0) done[b] = win ? 1 : 0;
    if (!win) return;
    uint8_t* out = mask + (size_t)b * H * W;
    for (int i = threadIdx.x; i < H * wpr; i += kWarps * 32) {
        const uint32_t w = comp[i];
        uint32_t v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = ((((w >> (4 * k)) & 0xfu) * 0x00204081u) & 0x01010101u) * 0xffu;
        int4* o = reinterpret_cast<int4*>(out + (size_t)i * 32);
        o[0] = make_int4((int)v[0], (int)v[1], (int)v[2], (int)v[3]);
        o[1] = make_int4((int)v[4], (int)v[5], (int)v[6], (int)v[7]);
    }
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

extern "C" size_t eitb_body_mask_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    const size_t n = (size_t)B * H * W;
    // er u8, opened u8, labels A int32, labels B int32, area2 int32, best int64[B], done int32[B]
    return align256(n) * 2 + align256(n * 4) * 3 + align256((size_t)B * 8) + align256((size_t)B * 4);
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
    long long* best = reinterpret_cast<long long*>(p); p += align256((size_t)B * 8);
    int* done = reinterpret_cast<int*>(p);
    const int* skip = nullptr;                                    // images the fast path answered
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
            if (!(reinterpret_cast<uintptr_t>(mask) & 15)) {
                const size_t dsm_bytes = (size_t)H * (W >> 5) * 8;
                if (cudaFuncSetAttribute(body_dominant_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm_bytes) != cudaSuccess)
                    return EITB_ERR_LAUNCH;
                eitb_prof_begin("body_dominant_kernel", s);
                body_dominant_kernel<<<B, eitb_flood::kWarps * 32, dsm_bytes, s>>>(outside, H, W, mask, done);
                EITB_CHECK_LAUNCH();
                skip = done;
            }
            rc = cc_label<PRED_BIT_ZERO, 8>(outside, (size_t)H * W / 8, 0, B, H, W, 0, labB, s, 1, skip);
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
    area_kernel<<<dim3(eitb_div_up(W + 1, 256), H + 1 < 64 ? H + 1 : 64, B), 256, 0, s>>>(labB, B, H, W, area2, skip);
    EITB_CHECK_LAUNCH();
    eitb_prof_begin("best_kernel", s);
    best_kernel<<<dim3(eitb_grid_per_image((long long)H * W, 256, B), B), 256, 0, s>>>(labB, area2, H * W, best, skip);
    EITB_CHECK_LAUNCH();
    eitb_prof_begin("write_mask_kernel", s);
    if (((H * W) & 3) == 0 && !(reinterpret_cast<uintptr_t>(mask) & 3))
        write_mask_kernel<<<dim3(eitb_grid_per_image((long long)H * W / 4, 256, B), B), 256, 0, s>>>(labB, best, H * W, mask, skip);
    else
        write_mask_scalar_kernel<<<dim3(eitb_grid_per_image((long long)H * W, 256, B), B), 256, 0, s>>>(labB, best, H * W, mask, skip);
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
