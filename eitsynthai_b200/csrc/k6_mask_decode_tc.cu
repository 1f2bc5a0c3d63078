// K6 on the 5th-generation tensor cores: the coefficient x prototype contraction of the mask
// decode as tcgen05.mma with the accumulator in tensor memory, fused with crop, x4 bilinear
// upsample, threshold and the label-image overlay (same semantics as k6_mask_decode.cu, which
// remains the path for fp32 / bf16 prototypes).
//
// Per CTA: a 16 x 12 tile of prototype pixels (+1 halo = 18 x 14 = 252 columns, padded to N = 256)
// -> 64 x 48 output pixels.
//   B operand  [N = 256 pixels][K = 64] fp16, K-major, 128-byte swizzle: the 32 prototype values of a
//              pixel, twice (for the high and the low half of the coefficients);
//   A operand  [M = 128 rows][K = 64] fp16: coefficient split hi | lo (c = hi + lo exactly to
//              22 bits), so fp32 coefficients lose nothing and every product is exact in fp32;
//              up to 56 live instances per pass, 14 in each 32-lane TMEM quadrant, so that all eight
//              warps can pull logits out of tensor memory at the same time;
//   D          [128 lanes][256 columns] fp32 in TMEM: 4 x tcgen05.mma (M128 N256 K16), one commit.
// Epilogue (round 2): every warp reads its quadrant's half of the columns with tcgen05.ld.x32, applies
// the crop and parks the logits of ALL live instances in shared memory in one sweep (conflict-free: a lane
// is an instance, rows are 253 words apart); one barrier; then the fixed-weight upsample + threshold + OR
// runs over the instances exactly as in the CUDA-core kernel.  Round 1 parked and upsampled one quadrant
// at a time with two warps reading TMEM and six waiting (33 % of the stall samples on that barrier).
// 108 KB of shared memory and 256 TMEM columns per CTA: two CTAs per SM.  fp32 masks never leave the SM.
#include "common.cuh"

namespace {

constexpr int PTX_ = 16, PTY_ = 12;          // prototype pixels per tile
constexpr int HTX = PTX_ + 2, HTY = PTY_ + 2;
constexpr int HPT = HTX * HTY;               // 252 halo pixels
constexpr int NCOL = 256;                    // MMA N, TMEM columns
constexpr int MROW = 128;                    // MMA M
constexpr int LPQ = 14;                      // live instances per TMEM lane quadrant
constexpr int LIVE = 4 * LPQ;                // live instances per pass
constexpr int KDIM = 64;                     // 32 coefficients, hi | lo
constexpr int LSTRIDE = 253;                 // odd row stride of the parked logits: conflict-free
constexpr int kThreads = 256;

constexpr int OFF_B = 0;                                   // 256 x 128 B
constexpr int OFF_A = OFF_B + NCOL * 128;                  // 128 x 128 B
constexpr int OFF_LS = OFF_A + MROW * 128;                 // LIVE x 253 x 4
constexpr int OFF_BOX = OFF_LS + LIVE * LSTRIDE * 4;       // LIVE x 4 f32
constexpr int OFF_INFO = OFF_BOX + LIVE * 16;              // LIVE x 2 int
constexpr int OFF_ACT = OFF_INFO + LIVE * 8;               // max_det ints
// after act: mbarrier (8 B), tmem slot (4 B), counters

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ int class_code(float cls) {
    const int c = (int)cls;
    return c == 0 ? EITB_CODE_BONE : c == 1 ? EITB_CODE_MUSCLE : c == 2 ? EITB_CODE_LUNG : c == 3 ? EITB_CODE_ADIPOSE : 0;
}

__device__ __forceinline__ float py_slice_bound(float v, int size) {       // see k6_mask_decode.cu
    const int i = __float2int_rn(v);
    return (float)(i >= 0 ? min(i, size) : max(0, size + i));
}
__device__ __forceinline__ float4 crop_box(const float* d, float rx, float ry, int mw, int mh, bool int_crop) {
    float4 b = make_float4(d[0] * rx, d[1] * ry, d[2] * rx, d[3] * ry);
    if (int_crop) {
        b.x = py_slice_bound(b.x, mw); b.z = py_slice_bound(b.z, mw);
        b.y = py_slice_bound(b.y, mh); b.w = py_slice_bound(b.w, mh);
    }
    return b;
}

// shared-memory matrix descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);               // start address
    d |= (uint64_t)1 << 16;                                  // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                        // stride byte offset
    d |= (uint64_t)1 << 46;                                  // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                  // SWIZZLE_128B
    return d;
}

// byte offset of 16-byte chunk `c` (0..7) of row `r` in the swizzled operand tile
__device__ __forceinline__ int swz(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }

__global__ void __launch_bounds__(kThreads, 2)
mask_decode_tc_kernel(const float* __restrict__ dets, const int32_t* __restrict__ n_det, int max_det,
                      const __half* __restrict__ protos, int proto_nhwc, int mh, int mw, int H, int W, int tiles_x,
                      int tiles_per_img, int variant, uint8_t* __restrict__ code, int32_t* __restrict__ inst_area,
                      uint8_t* __restrict__ inst_bits) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem + OFF_B;
    unsigned char* sA = smem + OFF_A;
    float* Ls = reinterpret_cast<float*>(smem + OFF_LS);
    float* sbox = reinterpret_cast<float*>(smem + OFF_BOX);
    int* sinfo = reinterpret_cast<int*>(smem + OFF_INFO);
    int* act = reinterpret_cast<int*>(smem + OFF_ACT);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + OFF_ACT + max_det * 4 + ((8 - (max_det * 4) % 8) % 8));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
    int* s_nact = reinterpret_cast<int*>(tmem_slot + 1);

    constexpr int nm = 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x / tiles_per_img;
    const int tile = blockIdx.x - b * tiles_per_img;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int D = 6 + nm;
    const float* dimg = dets + (long long)b * max_det * D;
    const int n = min(n_det[b], max_det);
    const float rx = (float)((double)mw / (double)W), ry = (float)((double)mh / (double)H);
    const bool int_crop = (variant & 4) && n_det[b] < 50;

    if (smem_u32(sB) & 1023u) __trap();                                        // swizzled operand tiles need 1 KiB alignment
    if (tid == 0) {
        *s_nact = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(NCOL));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }

    // ---- B operand: prototype halo tile (replicated at the frame), 32 values per pixel, twice
    {
        const __half* pimg = protos + (long long)b * nm * mh * mw;
        if (proto_nhwc) {
            for (int idx = tid; idx < NCOL * 4; idx += kThreads) {            // (pixel row, 16-byte chunk of 8 values)
                const int r = idx >> 2, j = idx & 3;
                int4 v = make_int4(0, 0, 0, 0);
                if (r < HPT) {
                    const int hy = r / HTX, hx = r - hy * HTX;
                    const int py = min(max(ty * PTY_ - 1 + hy, 0), mh - 1), px = min(max(tx * PTX_ - 1 + hx, 0), mw - 1);
                    v = __ldg(reinterpret_cast<const int4*>(pimg + ((long long)py * mw + px) * nm) + j);
                }
                *reinterpret_cast<int4*>(sB + swz(r, j)) = v;
                *reinterpret_cast<int4*>(sB + swz(r, j + 4)) = v;
            }
        } else {
            for (int idx = tid; idx < NCOL * nm; idx += kThreads) {
                const int k = idx / NCOL, r = idx - k * NCOL;                  // consecutive threads: consecutive pixels
                __half v = __float2half(0.f);
                if (r < HPT) {
                    const int hy = r / HTX, hx = r - hy * HTX;
                    const int py = min(max(ty * PTY_ - 1 + hy, 0), mh - 1), px = min(max(tx * PTX_ - 1 + hx, 0), mw - 1);
                    v = __ldg(pimg + ((long long)k * mh + py) * mw + px);
                }
                *reinterpret_cast<__half*>(sB + swz(r, k >> 3) + (k & 7) * 2) = v;
                *reinterpret_cast<__half*>(sB + swz(r, (k >> 3) + 4) + (k & 7) * 2) = v;
            }
        }
    }
    __syncthreads();

    // ---- instances whose crop box reaches this tile's halo
    {
        const float cx_lo = (float)max(tx * PTX_ - 1, 0), cx_hi = (float)min(tx * PTX_ + PTX_, mw - 1);
        const float cy_lo = (float)max(ty * PTY_ - 1, 0), cy_hi = (float)min(ty * PTY_ + PTY_, mh - 1);
        for (int i = tid; i < n; i += kThreads) {
            const float* d = dimg + (long long)i * D;
            const float4 bx = crop_box(d, rx, ry, mw, mh, int_crop);
            if (bx.z > cx_lo && bx.x <= cx_hi && bx.w > cy_lo && bx.y <= cy_hi && class_code(d[5]) != 0)
                act[atomicAdd(s_nact, 1)] = i;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const int nact = *s_nact;
    const uint32_t tmem = *tmem_slot;

    // this thread's 16 output pixels: row r of the tile (48 rows used), columns 16*cg .. 16*cg+15
    const int r = tid >> 2, cg = tid & 3;
    const int jy = r & 3;
    const int hy0 = (r >> 2) + (jy >= 2);
    const float ly = jy == 0 ? 0.625f : jy == 1 ? 0.875f : jy == 2 ? 0.125f : 0.375f;
    const int oy = ty * (PTY_ * 4) + r;
    const int ox0 = tx * (PTX_ * 4) + cg * 16;
    const bool row_ok = r < PTY_ * 4;
    const bool in_img = row_ok && oy < H && ox0 < W;
    uint32_t codes[4] = {0, 0, 0, 0};
    const float thr = (variant & 1) ? 0.5f : 0.0f;
    const float wrow_lo = (float)min(max(ty * PTY_ - 1 + (warp << 1), 0), mh - 1);      // prototype rows this warp reads
    const float wrow_hi = (float)min(max(ty * PTY_ - 1 + (warp << 1) + 3, 0), mh - 1);

    // instruction descriptor: D fp32, A/B fp16, both K-major, N = 256, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(NCOL >> 3) << 17) | ((uint32_t)(MROW >> 4) << 24);
    uint32_t phase = 0;
    const int pq = warp & 3;                                                  // the TMEM lane quadrant this warp may read
    const int pslot = pq * LPQ + lane;                                        // live slot of this lane in that quadrant

    for (int p0 = 0; p0 < nact; p0 += LIVE) {
        const int np = min(LIVE, nact - p0);
        // ---- A operand: row 32*q + l holds live slot q*LPQ + l (l < LPQ), every other row is zero
        for (int idx = tid; idx < MROW * 4; idx += kThreads) {                // (row, chunk of 8 coefficients)
            const int row = idx >> 2, j = idx & 3;
            const int slot = (row >> 5) * LPQ + (row & 31);
            const bool live = (row & 31) < LPQ && slot < np;
            __align__(16) __half hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float c = live ? dimg[(long long)act[p0 + slot] * D + 6 + j * 8 + e] : 0.f;
                hi[e] = __float2half_rn(c);
                lo[e] = __float2half_rn(c - __half2float(hi[e]));
            }
            *reinterpret_cast<int4*>(sA + swz(row, j)) = *reinterpret_cast<const int4*>(hi);
            *reinterpret_cast<int4*>(sA + swz(row, j + 4)) = *reinterpret_cast<const int4*>(lo);
        }
        if (tid < np) {
            const float* d = dimg + (long long)act[p0 + tid] * D;
            const float4 bx = crop_box(d, rx, ry, mw, mh, int_crop);
            sbox[tid * 4 + 0] = bx.x; sbox[tid * 4 + 1] = bx.y; sbox[tid * 4 + 2] = bx.z; sbox[tid * 4 + 3] = bx.w;
            sinfo[tid * 2] = class_code(d[5]);
            sinfo[tid * 2 + 1] = act[p0 + tid];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> tensor-core reads
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        if (tid == 0) {
            const uint64_t da = make_desc(smem_u32(sA)), db = make_desc(smem_u32(sB));
#pragma unroll
            for (int k = 0; k < KDIM / 16; ++k) {
                const uint32_t accum = k > 0;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tmem), "l"(da + (uint64_t)(k * 2)), "l"(db + (uint64_t)(k * 2)), "r"(idesc), "r"(accum));
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)));
        }
        // ---- wait for the accumulator
        {
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done) : "r"(smem_u32(mbar)), "r"(phase) : "memory");
            }
            phase ^= 1;
        }
        asm volatile("tcgen05.fence::after_thread_sync;");

        // ---- all eight warps park the cropped logits of every live instance (lane = instance, 128 columns per warp)
        {
            const bool live = lane < LPQ && pslot < np;
            float bx1 = 0.f, by1 = 0.f, bx2 = 0.f, by2 = 0.f;
            if (live) { bx1 = sbox[pslot * 4]; by1 = sbox[pslot * 4 + 1]; bx2 = sbox[pslot * 4 + 2]; by2 = sbox[pslot * 4 + 3]; }
            const int c_begin = (warp >> 2) * (NCOL / 2);
            const int col_x0 = tx * PTX_ - 1, col_y0 = ty * PTY_ - 1;
#pragma unroll 1
            for (int c0 = c_begin; c0 < c_begin + NCOL / 2; c0 += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem + ((uint32_t)(pq * 32) << 16) + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (live) {
                    int hy = c0 / HTX, hx = c0 - hy * HTX;                     // halo position of column c0 (warp-uniform)
                    float* dst = Ls + pslot * LSTRIDE + c0;
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        if (c0 + e < HPT) {
                            const float fx = (float)min(max(col_x0 + hx, 0), mw - 1), fy = (float)min(max(col_y0 + hy, 0), mh - 1);
                            float val = __uint_as_float(v[e]);
                            if (variant & 1) val = 1.f / (1.f + expf(-val));
                            dst[e] = (fx >= bx1 && fx < bx2 && fy >= by1 && fy < by2) ? val : 0.f;
                        }
                        if (++hx == HTX) { hx = 0; ++hy; }
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        // ---- x4 bilinear (align_corners=False) + threshold + OR over the live instances
        if (row_ok) {
            for (int g = 0; g < np; ++g) {
                if (sbox[g * 4 + 3] <= wrow_lo || sbox[g * 4 + 1] > wrow_hi) continue;   // warp-uniform
                const float* L0 = Ls + g * LSTRIDE + hy0 * HTX + cg * 4;
                const float* L1 = L0 + HTX;
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
                        const float vv = (1.f - lx) * V[x0] + lx * V[x0 + 1];
                        bits |= (vv > thr ? 1u : 0u) << j;
                    }
                }
                if (!in_img) bits = 0;
                const uint32_t cc = (uint32_t)sinfo[g * 2];
#pragma unroll
                for (int k = 0; k < 4; ++k)                    // 4 mask bits -> 4 bytes of the class code
                    codes[k] |= ((((bits >> (4 * k)) & 0xfu) * 0x00204081u) & 0x01010101u) * cc;
                const int inst = sinfo[g * 2 + 1];
                if (inst_bits && in_img) {
                    const long long o = (((long long)b * max_det + inst) * H + oy) * (W >> 3) + (ox0 >> 3);
                    *reinterpret_cast<uint16_t*>(inst_bits + o) = (uint16_t)bits;
                }
                if (inst_area && bits) atomicAdd(inst_area + (long long)b * max_det + inst, __popc(bits));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();                                                      // TMEM, A and the parked logits are free for the next pass
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    if (in_img)
        st_stream_int4(reinterpret_cast<int4*>(code + ((long long)b * H + oy) * W + ox0),
                       make_int4((int)codes[0], (int)codes[1], (int)codes[2], (int)codes[3]));
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(NCOL));
}

}  // namespace

// called from eitb_mask_decode when the prototypes are fp16 with 32 channels
int eitb_mask_decode_tc(const float* dets, const int32_t* n_det, int max_det, const void* protos, int nhwc, int B, int mh,
                        int mw, int H, int W, int variant, uint8_t* code, int32_t* inst_area, uint8_t* inst_bits,
                        cudaStream_t s) {
    const int tiles_x = eitb_div_up(mw, PTX_), tiles_y = eitb_div_up(mh, PTY_);
    const size_t smem = (size_t)OFF_ACT + (size_t)max_det * 4 + 64;
    if (cudaFuncSetAttribute(mask_decode_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return EITB_ERR_LAUNCH;
    const long long grid = (long long)B * tiles_x * tiles_y;
    if (grid > 0x7fffffffLL) return EITB_ERR_UNSUPPORTED;
    eitb_prof_begin("mask_decode_tc_kernel", s);
    mask_decode_tc_kernel<<<(unsigned)grid, kThreads, smem, s>>>(dets, n_det, max_det, (const __half*)protos, nhwc, mh, mw, H, W,
                                                                 tiles_x, tiles_x * tiles_y, variant, code, inst_area, inst_bits);
    EITB_CHECK_LAUNCH();
    return EITB_OK;
}
