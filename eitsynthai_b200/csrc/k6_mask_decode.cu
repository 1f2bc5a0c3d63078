// K6: mask decode (coef x proto contraction, crop, x4 bilinear upsample, threshold) fused with
// the label-image overlay.
//
// Restates ultralytics process_mask(upsample=True) (call site
// ai_tools.py:153) and the reference's create_segmentations_masks +
// overlay_segmentation_masks (utils.py:437-523, 395-434), whose saturating colour adds
// reduce to a per-pixel OR of 3-bit colour codes (bone 7, muscle 1, lung 6, adipose 3).
//
// One CTA owns a 16x16 tile of prototype pixels (+1 halo, replicated at the image frame like
// the clamped source index of F.interpolate) = 64x64 output pixels.  The prototype tile is
// staged once in shared memory as fp32; instances whose crop box touches the tile are
// processed 8 at a time: logits for the 18x18 halo tile (fp32 FMA, crop applied), then every
// thread interpolates its own 16 output pixels (fixed weights .125/.375/.625/.875) and ORs the
// class code into registers.  Per-instance fp32 masks never exist in HBM: traffic is the
// prototype read + one u8 code per pixel.
//
// MMA = true (fp16 NHWC prototypes with 32 channels, what the network emits): the halo tile is staged as fp16 rows
// (straight 16-byte copies) and the logits of a chunk are one warp-level tensor-core contraction per 16 halo pixels
// (mma.sync m16n8k16, fp32 accumulate; coefficients split into fp16 hi + lo so fp32 coefficients lose nothing) instead
// of 8 x 32 scalar FMAs per pixel; the crop / upsample / threshold / overlay part is the same code.
#include "common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int PT = 16;          // prototype pixels per tile side
constexpr int HT = PT + 2;      // with halo
constexpr int HP = HT * HT;     // 324
constexpr int G = 8;            // instances per chunk
constexpr int kThreads = 256;
constexpr int kMaxDet = 1024;
constexpr int kNm = 32;                      // MMA path: prototype channels
constexpr int kKS = 40;                      // MMA path: halfs per staged row (32 + 8 of padding: conflict-free fragment loads)
constexpr int kMTiles = (HP + 15) / 16;      // 21 tiles of 16 halo pixels
constexpr int kMRows = kMTiles * 16;         // 336 staged rows (the last 12 are zero)

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <typename T> __device__ __forceinline__ float ld_f32(const T* p);
template <> __device__ __forceinline__ float ld_f32<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_f32<__half>(const __half* p) { return __half2float(__ldg(p)); }
template <> __device__ __forceinline__ float ld_f32<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }

__device__ __forceinline__ int class_code(float cls) {
    // class id -> colour code (utils.py:468-473, 498-507); other ids are skipped
    const int c = (int)cls;
    return c == 0 ? EITB_CODE_BONE : c == 1 ? EITB_CODE_MUSCLE : c == 2 ? EITB_CODE_LUNG : c == 3 ? EITB_CODE_ADIPOSE : 0;
}

// Crop box of one detection in prototype pixels.  Float form (all ultralytics versions on the GPU, CPU with >= 50
// masks): keep x1 <= col < x2.  Integer form (late-2025 crop_mask on the CPU with fewer than 50 masks):
// boxes.round().int() used as Python slice bounds -- masks[:, :x1] = 0, masks[:, x2:] = 0 -- including Python's
// meaning of a negative bound (counted from the end).  Both reduce to "lo <= col < hi" on float bounds.
__device__ __forceinline__ float py_slice_bound(float v, int size) {
    const int i = __float2int_rn(v);                                  // torch.round: half to even, then .int()
    return (float)(i >= 0 ? min(i, size) : max(0, size + i));
}
__device__ __forceinline__ float4 eitb_crop_box(const float* d, float rx, float ry, int mw, int mh, bool int_crop) {
    float4 b = make_float4(d[0] * rx, d[1] * ry, d[2] * rx, d[3] * ry);
    if (int_crop) {
        b.x = py_slice_bound(b.x, mw); b.z = py_slice_bound(b.z, mw);
        b.y = py_slice_bound(b.y, mh); b.w = py_slice_bound(b.w, mh);
    }
    return b;
}

template <typename T, bool MMA>
__global__ void __launch_bounds__(kThreads)
mask_decode_kernel(const float* __restrict__ dets, const int32_t* __restrict__ n_det, int max_det,
                   const T* __restrict__ protos, int proto_nhwc, int nm, int mh, int mw, int H, int W, int tiles_x,
                   int tiles_per_img, int variant, uint8_t* __restrict__ code, int32_t* __restrict__ inst_area,
                   uint8_t* __restrict__ inst_bits) {
    extern __shared__ __align__(16) float smem[];
    float* P = smem;                         // [nm][HP] fp32                       | MMA: [kMRows][kKS] fp16
    float* Ls = MMA ? smem + kMRows * kKS / 2 : P + nm * HP;                 // [G][HP]
    float* sc = Ls + G * HP;                 // [nm][G]: the G coefficients of one prototype channel are two 16-byte words
                                             // MMA: [2][G][nm] fp16 (hi, lo), the same number of bytes
    float* sbox = sc + G * nm;               // [G][4]  crop box in prototype pixels
    int* sinfo = reinterpret_cast<int*>(sbox + G * 4);   // [G][2]  code, instance index
    int* act = sinfo + G * 2;                // [max_det]
    __shared__ int s_nact;

    const int tid = threadIdx.x;
    const int b = blockIdx.x / tiles_per_img;
    const int tile = blockIdx.x - b * tiles_per_img;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int D = 6 + nm;
    const float* dimg = dets + (long long)b * max_det * D;
    const int n = min(n_det[b], max_det);
    const float rx = (float)((double)mw / (double)W), ry = (float)((double)mh / (double)H);
    const bool int_crop = (variant & 4) && n_det[b] < 50;
    variant &= 1;
    if (tid == 0) s_nact = 0;

    // ---- stage the prototype tile (halo replicated at the frame)
    {
        const T* pimg = protos + (long long)b * nm * mh * mw;
        if (MMA) {                                         // a pixel's 32 halfs = four 16-byte words, copied as they are
            __half* P16 = reinterpret_cast<__half*>(P);
            for (int idx = tid; idx < kMRows * 4; idx += kThreads) {
                const int h = idx >> 2, q = idx & 3;
                int4 v = make_int4(0, 0, 0, 0);
                if (h < HP) {
                    const int hy = h / HT, hx = h - hy * HT;
                    const int py = min(max(ty * PT - 1 + hy, 0), mh - 1);
                    const int px = min(max(tx * PT - 1 + hx, 0), mw - 1);
                    v = __ldg(reinterpret_cast<const int4*>(pimg + ((long long)py * mw + px) * kNm) + q);
                }
                *reinterpret_cast<int4*>(P16 + h * kKS + q * 8) = v;
            }
        } else if (proto_nhwc) {                                  // [mh][mw][nm]: the nm values of a pixel are contiguous
            for (int idx = tid; idx < nm * HP; idx += kThreads) {
                const int h = idx / nm, k = idx - h * nm;
                const int hy = h / HT, hx = h - hy * HT;
                const int py = min(max(ty * PT - 1 + hy, 0), mh - 1);
                const int px = min(max(tx * PT - 1 + hx, 0), mw - 1);
                P[k * HP + h] = ld_f32(pimg + ((long long)py * mw + px) * nm + k);
            }
        } else {
            for (int idx = tid; idx < nm * HP; idx += kThreads) {
                const int k = idx / HP, h = idx - k * HP;
                const int hy = h / HT, hx = h - hy * HT;
                const int py = min(max(ty * PT - 1 + hy, 0), mh - 1);
                const int px = min(max(tx * PT - 1 + hx, 0), mw - 1);
                P[idx] = ld_f32(pimg + ((long long)k * mh + py) * mw + px);
            }
        }
    }
    __syncthreads();

    // ---- instances whose crop box reaches this tile's halo
    {
        const float cx_lo = (float)max(tx * PT - 1, 0), cx_hi = (float)min(tx * PT + PT, mw - 1);
        const float cy_lo = (float)max(ty * PT - 1, 0), cy_hi = (float)min(ty * PT + PT, mh - 1);
        for (int i = tid; i < n; i += kThreads) {
            const float* d = dimg + (long long)i * D;
            const float4 bx = eitb_crop_box(d, rx, ry, mw, mh, int_crop);
            if (bx.z > cx_lo && bx.x <= cx_hi && bx.w > cy_lo && bx.y <= cy_hi && class_code(d[5]) != 0)
                act[atomicAdd(&s_nact, 1)] = i;
        }
    }
    __syncthreads();
    const int nact = s_nact;

    // this thread's 16 output pixels: row r of the tile, columns 16*cg .. 16*cg+15
    const int r = tid >> 2, cg = tid & 3;
    const int jy = r & 3;
    const int hy0 = (r >> 2) + (jy >= 2);
    const float ly = jy == 0 ? 0.625f : jy == 1 ? 0.875f : jy == 2 ? 0.125f : 0.375f;
    const int oy = ty * (PT * 4) + r;
    const int ox0 = tx * (PT * 4) + cg * 16;
    const bool in_img = oy < H && ox0 < W;
    uint32_t codes[4] = {0, 0, 0, 0};        // 16 x u8
    const float thr = variant == 1 ? 0.5f : 0.0f;
    // prototype rows read by the 8 output rows of this warp (clamped like the halo)
    const float wrow_lo = (float)min(max(ty * PT - 1 + ((tid >> 5) << 1), 0), mh - 1);
    const float wrow_hi = (float)min(max(ty * PT - 1 + ((tid >> 5) << 1) + 3, 0), mh - 1);

    for (int c0 = 0; c0 < nact; c0 += G) {
        const int ng = min(G, nact - c0);
        if (MMA) {
            __half* sch = reinterpret_cast<__half*>(sc);   // [G][nm] hi, then [G][nm] lo; unused slots are zero
            for (int i = tid; i < G * kNm; i += kThreads) {
                const int g = i >> 5, k = i & 31;
                const float c = g < ng ? dimg[(long long)act[c0 + g] * D + 6 + k] : 0.f;
                const __half hi = __float2half_rn(c);
                sch[i] = hi;
                sch[G * kNm + i] = __float2half_rn(c - __half2float(hi));
            }
        } else {
        for (int i = tid; i < ng * nm; i += kThreads) {
            const int g = i / nm, k = i - g * nm;
            sc[k * G + g] = dimg[(long long)act[c0 + g] * D + 6 + k];
        }
        }
        if (tid < ng) {
            const float* d = dimg + (long long)act[c0 + tid] * D;
            const float4 bx = eitb_crop_box(d, rx, ry, mw, mh, int_crop);
            sbox[tid * 4 + 0] = bx.x; sbox[tid * 4 + 1] = bx.y; sbox[tid * 4 + 2] = bx.z; sbox[tid * 4 + 3] = bx.w;
            sinfo[tid * 2] = class_code(d[5]);
            sinfo[tid * 2 + 1] = act[c0 + tid];
        }
        __syncthreads();
        // logits of the halo tile for the chunk, cropped to each instance's box
        if (MMA) {
            const int warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
            const __half* P16 = reinterpret_cast<const __half*>(P);
            const __half* sch = reinterpret_cast<const __half*>(sc);
            uint32_t bh[2][2], bl[2][2];                   // B fragments: instance gid, channels 16 s + 2 tig (+ 8)
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                bh[s2][0] = *reinterpret_cast<const uint32_t*>(sch + gid * kNm + 16 * s2 + 2 * tig);
                bh[s2][1] = *reinterpret_cast<const uint32_t*>(sch + gid * kNm + 16 * s2 + 2 * tig + 8);
                bl[s2][0] = *reinterpret_cast<const uint32_t*>(sch + G * kNm + gid * kNm + 16 * s2 + 2 * tig);
                bl[s2][1] = *reinterpret_cast<const uint32_t*>(sch + G * kNm + gid * kNm + 16 * s2 + 2 * tig + 8);
            }
            for (int mt = warp; mt < kMTiles; mt += kThreads / 32) {
                float c[4] = {0.f, 0.f, 0.f, 0.f};         // rows gid / gid + 8 of the tile x instances 2 tig / 2 tig + 1
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    const __half* ab = P16 + (mt * 16 + gid) * kKS + 16 * s2 + 2 * tig;
                    const uint32_t a0 = *reinterpret_cast<const uint32_t*>(ab);
                    const uint32_t a1 = *reinterpret_cast<const uint32_t*>(ab + 8 * kKS);
                    const uint32_t a2 = *reinterpret_cast<const uint32_t*>(ab + 8);
                    const uint32_t a3 = *reinterpret_cast<const uint32_t*>(ab + 8 * kKS + 8);
                    mma_16816(c, a0, a1, a2, a3, bl[s2][0], bl[s2][1]);
                    mma_16816(c, a0, a1, a2, a3, bh[s2][0], bh[s2][1]);
                }
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int h = mt * 16 + gid + 8 * rr;
                    if (h >= HP) continue;
                    const int hy = h / HT, hx = h - hy * HT;
                    const float fy = (float)min(max(ty * PT - 1 + hy, 0), mh - 1);
                    const float fx = (float)min(max(tx * PT - 1 + hx, 0), mw - 1);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int g = 2 * tig + j;
                        if (g < ng) {
                            const bool in = fx >= sbox[g * 4] && fx < sbox[g * 4 + 2] && fy >= sbox[g * 4 + 1] && fy < sbox[g * 4 + 3];
                            float v = c[rr * 2 + j];
                            if (variant == 1) v = 1.f / (1.f + expf(-v));
                            Ls[g * HP + h] = in ? v : 0.f;
                        }
                    }
                }
            }
        } else
        for (int h = tid; h < HP; h += kThreads) {
            float acc[G];
#pragma unroll
            for (int g = 0; g < G; ++g) acc[g] = 0.f;
            // per channel: one prototype value + two broadcast 16-byte coefficient loads feed 8 FMAs (coefficients of
            // slots >= ng are stale, their sums are never used)
            const float4* sc4 = reinterpret_cast<const float4*>(sc);
#pragma unroll 4
            for (int k = 0; k < nm; ++k) {
                const float p = P[k * HP + h];
                const float4 c0 = sc4[2 * k], c1 = sc4[2 * k + 1];
                acc[0] = fmaf(c0.x, p, acc[0]); acc[1] = fmaf(c0.y, p, acc[1]); acc[2] = fmaf(c0.z, p, acc[2]); acc[3] = fmaf(c0.w, p, acc[3]);
                acc[4] = fmaf(c1.x, p, acc[4]); acc[5] = fmaf(c1.y, p, acc[5]); acc[6] = fmaf(c1.z, p, acc[6]); acc[7] = fmaf(c1.w, p, acc[7]);
            }
            const int hy = h / HT, hx = h - hy * HT;
            const float fy = (float)min(max(ty * PT - 1 + hy, 0), mh - 1);
            const float fx = (float)min(max(tx * PT - 1 + hx, 0), mw - 1);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                if (g < ng) {
                    const bool in = fx >= sbox[g * 4] && fx < sbox[g * 4 + 2] && fy >= sbox[g * 4 + 1] && fy < sbox[g * 4 + 3];
                    float v = acc[g];
                    if (variant == 1) v = 1.f / (1.f + expf(-v));
                    Ls[g * HP + h] = in ? v : 0.f;
                }
            }
        }
        __syncthreads();
        // x4 bilinear (align_corners=False) + threshold + OR
        for (int g = 0; g < ng; ++g) {
            if (sbox[g * 4 + 3] <= wrow_lo || sbox[g * 4 + 1] > wrow_hi) continue;   // box misses this warp's rows (warp-uniform)
            const float* L0 = Ls + g * HP + hy0 * HT + cg * 4;
            const float* L1 = L0 + HT;
            float a0[6], a1[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) { a0[i] = L0[i]; a1[i] = L1[i]; }
            float vmin = fminf(a0[0], a1[0]), vmax = fmaxf(a0[0], a1[0]);
#pragma unroll
            for (int i = 1; i < 6; ++i) { vmin = fminf(vmin, fminf(a0[i], a1[i])); vmax = fmaxf(vmax, fmaxf(a0[i], a1[i])); }
            // every output is a convex combination of these 12 logits: all above / none above the threshold
            // decides the 16 pixels without interpolating (only mask borders take the slow path)
            uint32_t bits = vmin > thr ? 0xffffu : 0u;
            if (vmax > thr && !(vmin > thr)) {
            float V[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) V[i] = (1.f - ly) * a0[i] + ly * a1[i];
            bits = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int mm = j >> 2, jj = j & 3;
                const int x0 = mm + (jj >= 2);
                const float lx = jj == 0 ? 0.625f : jj == 1 ? 0.875f : jj == 2 ? 0.125f : 0.375f;
                const float v = (1.f - lx) * V[x0] + lx * V[x0 + 1];
                bits |= (v > thr ? 1u : 0u) << j;
            }
            }
            if (!in_img) bits = 0;
            const uint32_t cc = (uint32_t)sinfo[g * 2];
#pragma unroll
            for (int k = 0; k < 4; ++k)                            // 4 mask bits -> 4 bytes of the class code
                codes[k] |= ((((bits >> (4 * k)) & 0xfu) * 0x00204081u) & 0x01010101u) * cc;
            if (inst_bits && in_img) {
                const long long o = (((long long)b * max_det + sinfo[g * 2 + 1]) * H + oy) * (W >> 3) + (ox0 >> 3);
                *reinterpret_cast<uint16_t*>(inst_bits + o) = (uint16_t)bits;
            }
            if (inst_area) {
                const int cnt = warp_sum(__popc(bits));
                if ((tid & 31) == 0 && cnt) atomicAdd(inst_area + (long long)b * max_det + sinfo[g * 2 + 1], cnt);
            }
        }
        __syncthreads();
    }
    if (in_img)
        st_stream_int4(reinterpret_cast<int4*>(code + ((long long)b * H + oy) * W + ox0),
                       make_int4((int)codes[0], (int)codes[1], (int)codes[2], (int)codes[3]));
}

size_t decode_smem(int nm, int max_det, bool mma) {
    const size_t tile = mma ? (size_t)kMRows * kKS * 2 : (size_t)nm * HP * 4;
    return tile + (size_t)(G * HP + G * nm + G * 4) * 4 + G * 2 * 4 + (size_t)max_det * 4;
}

template <typename T, bool MMA = false>
int launch_decode(const float* dets, const int32_t* n_det, int max_det, const void* protos, int nhwc, int B, int nm, int mh,
                  int mw, int H, int W, int variant, uint8_t* code, int32_t* inst_area, uint8_t* inst_bits,
                  cudaStream_t s) {
    const int tiles_x = eitb_div_up(mw, PT), tiles_y = eitb_div_up(mh, PT);
    const size_t smem = decode_smem(nm, max_det, MMA);
    if (smem > 200 * 1024) return EITB_ERR_UNSUPPORTED;
    if (cudaFuncSetAttribute(mask_decode_kernel<T, MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return EITB_ERR_LAUNCH;
    const long long grid = (long long)B * tiles_x * tiles_y;
    if (grid > 0x7fffffffLL) return EITB_ERR_UNSUPPORTED;
    eitb_prof_begin("mask_decode_kernel", s);
    mask_decode_kernel<T, MMA><<<(unsigned)grid, kThreads, smem, s>>>(dets, n_det, max_det, (const T*)protos, nhwc, nm, mh, mw, H, W,
                                                                      tiles_x, tiles_x * tiles_y, variant, code, inst_area,
                                                                      inst_bits);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}

}  // namespace

int eitb_mask_decode_tc(const float* dets, const int32_t* n_det, int max_det, const void* protos, int nhwc, int B, int mh,
                        int mw, int H, int W, int variant, uint8_t* code, int32_t* inst_area, uint8_t* inst_bits,
                        cudaStream_t s);                     // k6_mask_decode_tc.cu

extern "C" size_t eitb_mask_decode_workspace_bytes(int B, int max_det, int nm, int mh, int mw) {
    (void)B; (void)max_det; (void)nm; (void)mh; (void)mw;
    return 0;
}

extern "C" int eitb_mask_decode(const float* dets, const int32_t* n_det, int max_det, const void* protos,
                                int proto_dtype, int proto_channels_last, int B, int nm, int mh, int mw, int H, int W,
                                int variant, uint8_t* code, int32_t* inst_area, uint8_t* inst_bits, void* ws, size_t ws_bytes,
                                eitb_stream_t stream) {
    (void)ws; (void)ws_bytes;
    if (!dets || !n_det || !protos || !code || B < 0 || nm <= 0 || mh <= 0 || mw <= 0 || max_det <= 0 ||
        max_det > kMaxDet || (variant & ~0x35))
        return EITB_ERR_BAD_ARG;
    if (H != 4 * mh || W != 4 * mw || (mw % 4) != 0) return EITB_ERR_UNSUPPORTED;
    if (B == 0) return EITB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (inst_area && cudaMemsetAsync(inst_area, 0, (size_t)B * max_det * sizeof(int32_t), s) != cudaSuccess)
        return EITB_ERR_LAUNCH;
    if (inst_bits && cudaMemsetAsync(inst_bits, 0, (size_t)B * max_det * H * (W / 8), s) != cudaSuccess)
        return EITB_ERR_LAUNCH;
    // fp16 prototypes with 32 channels (what the network emits): contraction on the tensor cores
    if (proto_dtype == EITB_F16 && nm == 32 && (variant & 0x20) && !(variant & 0x10) && !(reinterpret_cast<uintptr_t>(protos) & 15))
        return eitb_mask_decode_tc(dets, n_det, max_det, protos, proto_channels_last, B, mh, mw, H, W, variant & 5, code,
                                   inst_area, inst_bits, s);
    // the same prototypes stored NHWC, unless bit 4 asks for the scalar kernel: warp-level MMA for the logits
    if (proto_dtype == EITB_F16 && nm == kNm && proto_channels_last && !(variant & 0x10) && !(reinterpret_cast<uintptr_t>(protos) & 15))
        return launch_decode<__half, true>(dets, n_det, max_det, protos, 1, B, nm, mh, mw, H, W, variant & 5, code, inst_area, inst_bits, s);
    variant &= 5;
    switch (proto_dtype) {
        case EITB_F32: return launch_decode<float>(dets, n_det, max_det, protos, proto_channels_last, B, nm, mh, mw, H, W, variant, code, inst_area, inst_bits, s);
        case EITB_F16: return launch_decode<__half>(dets, n_det, max_det, protos, proto_channels_last, B, nm, mh, mw, H, W, variant, code, inst_area, inst_bits, s);
        case EITB_BF16: return launch_decode<__nv_bfloat16>(dets, n_det, max_det, protos, proto_channels_last, B, nm, mh, mw, H, W, variant, code, inst_area, inst_bits, s);
        default: return EITB_ERR_BAD_ARG;
    }
}
