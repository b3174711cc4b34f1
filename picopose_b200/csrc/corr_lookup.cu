// Stage-3 correlation-window lookup and the generic bilinear sampler (sm_100a).
//
// pp_corr_lookup replaces CorrLookup.forward (reference utils/corr_lookup.py:100-134): for every
// query pixel it samples a (2r+1)^2 bilinear window of its own slice of each pyramid level.
//
// Kernel shape (HBM-bound gather):
//   * one warp owns 32 consecutive queries (flat h*W+w) of one detection, so every output channel
//     store is one coalesced 128-byte line in the reference's (B, L*D*D, H, W) layout;
//   * per level the warp first stages, with 16-byte cp.async (zero-filled outside the map), the
//     <= (D+2) x (D+2) footprint of each of its 32 windows into shared memory -- consecutive lanes
//     fetch consecutive 16-byte pieces of a row, only pieces the window really touches;
//   * then lane = query: each window sample gathers its 4 taps from shared memory.
// Coordinates follow the reference's float arithmetic op by op (x*2/(W-1)-1 and grid_sample's
// inverse, utils/corr_lookup.py:61-65 + ATen grid_sampler_unnormalize), so floor/weights agree.
#include "pp_common.cuh"

namespace pp {

constexpr int LOOKUP_MAX_LEVELS = 8;
constexpr int LOOKUP_MAX_RADIUS = 16;

struct LookupParams {
    const float* vol[LOOKUP_MAX_LEVELS];
    int vh[LOOKUP_MAX_LEVELS];
    int vw[LOOKUP_MAX_LEVELS];
    int vec_ok[LOOKUP_MAX_LEVELS];  // 16-byte path usable (width % 4 == 0, base aligned)
    int L;
    const float* flow;
    float* out;
    int B, H, W, HW;
    int radius;
    int groups_per_b;  // ceil(HW / 32)
    int total_groups;
    int nr_max, nv_max;  // staged rows / 16-byte pieces per row
    int qstride;         // words between two queries' staging areas (odd -> spreads banks)
};

// pixel coordinate -> (floor index, weight of the upper tap), replicating
//   g = p*2/max(size-1,1) - 1 ; i = ((g+1)/2)*(size-1)
__device__ __forceinline__ void axis_tap(float p, int size, int& i0, float& w1) {
    float den = (float)(size > 1 ? size - 1 : 1);
    float g = __fsub_rn(__fdiv_rn(__fmul_rn(p, 2.0f), den), 1.0f);
    float i = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
    float f = floorf(i);
    w1 = __fsub_rn(i, f);
    // clamp so that far-away (or non-finite) windows stay inside the staged footprint; every tap of
    // a clamped index is out of bounds and contributes zero, exactly as zero padding does.
    f = fminf(fmaxf(f, -2.0f), (float)size);
    i0 = (f == f) ? (int)f : -2;
    if (!(w1 >= 0.0f && w1 <= 1.0f)) w1 = 0.0f;  // non-finite coordinates: everything is padding
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <int RT>  // RT > 0: compile-time radius, tap tables live in registers; RT == 0: runtime radius
__global__ void __launch_bounds__(128) corr_lookup_kernel(const LookupParams p) {
    extern __shared__ __align__(16) float smem[];
    const int warps_per_block = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int r = RT > 0 ? RT : p.radius;
    const int D = 2 * r + 1;
    constexpr int TAB = RT > 0 ? 2 * RT + 1 : 2 * LOOKUP_MAX_RADIUS + 1;
    const int ND = RT > 0 ? TAB : D;  // compile-time trip count (full unroll) when the radius is templated

    // per-warp staging: 32 windows + a small header (origin of each window's footprint)
    float* stage = smem + (size_t)warp * (32 * p.qstride + 4 * 32);
    int* hdr = reinterpret_cast<int*>(stage + 32 * p.qstride);  // [0..31]=xs [32..63]=ymin [64..95]=xmax [96..127]=ymax

    for (int g = blockIdx.x * warps_per_block + warp; g < p.total_groups; g += gridDim.x * warps_per_block) {
        const int b = g / p.groups_per_b;
        const int hw = (g - b * p.groups_per_b) * 32 + lane;
        const bool live = hw < p.HW;
        const int hwc = live ? hw : p.HW - 1;
        const int qh = hwc / p.W, qw = hwc - qh * p.W;
        const float fx = __ldg(p.flow + ((size_t)b * 2 + 0) * p.HW + hwc);
        const float fy = __ldg(p.flow + ((size_t)b * 2 + 1) * p.HW + hwc);
        const float cx = __fadd_rn((float)qw, fx);  // coords_grid + flow, utils/corr_lookup.py:113
        const float cy = __fadd_rn((float)qh, fy);
        float* out_q = p.out + (size_t)b * p.L * D * D * p.HW + hwc;

        for (int l = 0; l < p.L; ++l) {
            const int Hl = p.vh[l], Wl = p.vw[l];
            const float inv = 1.0f / (float)(1 << l);  // exact: centroid / 2**l, utils/corr_lookup.py:125
            const float lx = __fmul_rn(cx, inv), ly = __fmul_rn(cy, inv);
            int xo[TAB], yo[TAB];
            float xw[TAB], yw[TAB];
#pragma unroll
            for (int a = 0; a < ND; ++a) {
                axis_tap(__fadd_rn(lx, (float)(a - r)), Wl, xo[a], xw[a]);
                axis_tap(__fadd_rn(ly, (float)(a - r)), Hl, yo[a], yw[a]);
            }
            const int xmin = xo[0], xmax = xo[D - 1] + 1;
            const int ymin = yo[0], ymax = yo[D - 1] + 1;
            const bool vec = p.vec_ok[l] != 0;
            const int xs = vec ? (xmin & ~3) : xmin;  // two's complement: rounds towards -inf
            __syncwarp();                              // previous level's readers are done
            hdr[lane] = xs;
            hdr[32 + lane] = ymin;
            hdr[64 + lane] = xmax;
            hdr[96 + lane] = ymax;
            __syncwarp();

            // ---- cooperative staging -------------------------------------------------------
            if (vec) {
                const int per_q = p.nr_max * p.nv_max;
                const int total = 32 * per_q;
                for (int i = lane; i < total; i += 32) {
                    const int ql = i / per_q;
                    const int rem = i - ql * per_q;
                    const int row = rem / p.nv_max;
                    const int v = rem - row * p.nv_max;
                    const int y = hdr[32 + ql] + row;
                    const int x = hdr[ql] + 4 * v;
                    if (y > hdr[96 + ql] || x > hdr[64 + ql]) continue;  // not touched by any tap
                    float* dst = stage + ql * p.qstride + (row * p.nv_max + v) * 4;
                    const int qhw = (g - b * p.groups_per_b) * 32 + ql;
                    if (y >= 0 && y < Hl && x >= 0 && x < Wl && qhw < p.HW) {
                        const size_t qq = (size_t)b * p.HW + qhw;
                        cp_async16(dst, p.vol[l] + (qq * Hl + y) * Wl + x);
                    } else {
                        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            } else {
                const int cols = p.nv_max * 4;
                const int per_q = p.nr_max * cols;
                const int total = 32 * per_q;
                for (int i = lane; i < total; i += 32) {
                    const int ql = i / per_q;
                    const int rem = i - ql * per_q;
                    const int row = rem / cols;
                    const int c = rem - row * cols;
                    const int y = hdr[32 + ql] + row;
                    const int x = hdr[ql] + c;
                    if (y > hdr[96 + ql] || x > hdr[64 + ql]) continue;
                    float* dst = stage + ql * p.qstride + row * cols + c;
                    const int qhw = (g - b * p.groups_per_b) * 32 + ql;
                    if (y >= 0 && y < Hl && x >= 0 && x < Wl && qhw < p.HW) {
                        const size_t qq = (size_t)b * p.HW + qhw;
                        cp_async4(dst, p.vol[l] + (qq * Hl + y) * Wl + x);
                    } else {
                        *dst = 0.f;
                    }
                }
            }
            cp_async_wait_all();
            __syncwarp();

            // ---- lane = query: gather + 4-tap blend ------------------------------------------
            const float* win = stage + lane * p.qstride;
            const int rowpitch = p.nv_max * 4;
            float* out_l = out_q + (size_t)l * D * D * p.HW;
            if (live) {
#pragma unroll
                for (int a = 0; a < ND; ++a) {
                    const int cxo = xo[a] - xs;
                    const float wx1 = xw[a];
                    const float wx0 = __fsub_rn(1.0f, wx1);
#pragma unroll
                    for (int bb = 0; bb < ND; ++bb) {
                        const float* t0 = win + (yo[bb] - ymin) * rowpitch + cxo;
                        const float wy1 = yw[bb];
                        const float wy0 = __fsub_rn(1.0f, wy1);
                        const float v00 = t0[0], v01 = t0[1];
                        const float v10 = t0[rowpitch], v11 = t0[rowpitch + 1];
                        // same association as ATen's grid_sampler: sum of value * (wx*wy)
                        float acc = v00 * (wx0 * wy0);
                        acc = fmaf(v01, wx1 * wy0, acc);
                        acc = fmaf(v10, wx0 * wy1, acc);
                        acc = fmaf(v11, wx1 * wy1, acc);
                        __stcs(out_l + (size_t)(a * D + bb) * p.HW, acc);
                    }
                }
            }
        }
    }
}

// ---- generic bilinear sampler (F.grid_sample bilinear / zeros) -----------------------------
__global__ void bilinear_sample_kernel(const float* __restrict__ feat, const float* __restrict__ grid,
                                       int N, int C, int Hf, int Wf, int Ho, int Wo, int grid_chw,
                                       int align_corners, int scale, float* __restrict__ out) {
    // one thread per output pixel, loops over channels: reads of the 4 taps are coalesced across
    // neighbouring pixels when the sampling field is smooth (feature warping by a flow field).
    const long long total = (long long)N * Ho * Wo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / ((long long)Ho * Wo));
        const int pix = (int)(i - (long long)n * Ho * Wo);
        float gx, gy;
        if (grid_chw) {
            gx = grid[((size_t)n * 2 + 0) * Ho * Wo + pix];
            gy = grid[((size_t)n * 2 + 1) * Ho * Wo + pix];
        } else {
            gx = grid[((size_t)n * Ho * Wo + pix) * 2 + 0];
            gy = grid[((size_t)n * Ho * Wo + pix) * 2 + 1];
        }
        if (scale) {
            gx = __fsub_rn(__fdiv_rn(__fmul_rn(gx, 2.0f), (float)(Wf > 1 ? Wf - 1 : 1)), 1.0f);
            gy = __fsub_rn(__fdiv_rn(__fmul_rn(gy, 2.0f), (float)(Hf > 1 ? Hf - 1 : 1)), 1.0f);
        }
        float ix, iy;
        if (align_corners) {
            ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.0f), 0.5f), (float)(Wf - 1));
            iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.0f), 0.5f), (float)(Hf - 1));
        } else {
            ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), (float)Wf), 1.0f), 0.5f);
            iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), (float)Hf), 1.0f), 0.5f);
        }
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const float wx1 = ix - fx0, wy1 = iy - fy0;
        const float wx0 = (fx0 + 1.0f) - ix, wy0 = (fy0 + 1.0f) - iy;
        const bool finite = (ix == ix) && (iy == iy) && fabsf(ix) < 1e9f && fabsf(iy) < 1e9f;
        const int x0 = finite ? (int)fx0 : -2, y0 = finite ? (int)fy0 : -2;
        const bool okx0 = x0 >= 0 && x0 < Wf, okx1 = x0 + 1 >= 0 && x0 + 1 < Wf;
        const bool oky0 = y0 >= 0 && y0 < Hf, oky1 = y0 + 1 >= 0 && y0 + 1 < Hf;
        const float w00 = (okx0 && oky0) ? wx0 * wy0 : 0.f, w01 = (okx1 && oky0) ? wx1 * wy0 : 0.f;
        const float w10 = (okx0 && oky1) ? wx0 * wy1 : 0.f, w11 = (okx1 && oky1) ? wx1 * wy1 : 0.f;
        const int xa = min(max(x0, 0), Wf - 1), xb = min(max(x0 + 1, 0), Wf - 1);
        const int ya = min(max(y0, 0), Hf - 1), yb = min(max(y0 + 1, 0), Hf - 1);
        const float* f = feat + (size_t)n * C * Hf * Wf;
        float* o = out + (size_t)n * C * Ho * Wo + pix;
        for (int c = 0; c < C; ++c) {
            const float* fc = f + (size_t)c * Hf * Wf;
            float acc = fc[ya * Wf + xa] * w00;
            acc = fmaf(fc[ya * Wf + xb], w01, acc);
            acc = fmaf(fc[yb * Wf + xa], w10, acc);
            acc = fmaf(fc[yb * Wf + xb], w11, acc);
            o[(size_t)c * Ho * Wo] = acc;
        }
    }
}

template <int RT>
static int launch_lookup(const LookupParams& p, int warps_per_block, size_t smem, int grid, cudaStream_t st) {
    PP_CUDA(cudaFuncSetAttribute(corr_lookup_kernel<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    corr_lookup_kernel<RT><<<grid, warps_per_block * 32, smem, st>>>(p);
    PP_LAUNCHED();
    return PP_OK;
}

}  // namespace pp

extern "C" int pp_corr_lookup(const void* const* pyr_ptrs, const int* pyr_h, const int* pyr_w, int L,
                              const float* flow, int B, int H, int W, int radius, float* out, void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (B == 0) return PP_OK;  // empty batch: nothing to do (pointers of empty tensors may be null)
    PP_CHECK_ARG(pyr_ptrs && pyr_h && pyr_w && flow && out, "pp_corr_lookup: null pointer");
    PP_CHECK_ARG(L >= 1 && L <= LOOKUP_MAX_LEVELS, "pp_corr_lookup: 1 <= levels <= %d (got %d)", LOOKUP_MAX_LEVELS, L);
    PP_CHECK_ARG(radius >= 0 && radius <= LOOKUP_MAX_RADIUS, "pp_corr_lookup: 0 <= radius <= %d (got %d)",
                 LOOKUP_MAX_RADIUS, radius);
    PP_CHECK_ARG(B >= 0 && H > 0 && W > 0, "pp_corr_lookup: bad flow shape (%d,2,%d,%d)", B, H, W);
    if (B == 0) return PP_OK;
    LookupParams p{};
    p.L = L;
    for (int l = 0; l < L; ++l) {
        PP_CHECK_ARG(pyr_ptrs[l] && pyr_h[l] > 0 && pyr_w[l] > 0, "pp_corr_lookup: bad pyramid level %d", l);
        p.vol[l] = static_cast<const float*>(pyr_ptrs[l]);
        p.vh[l] = pyr_h[l];
        p.vw[l] = pyr_w[l];
        p.vec_ok[l] = (pyr_w[l] % 4 == 0) && ((reinterpret_cast<uintptr_t>(pyr_ptrs[l]) & 15) == 0);
    }
    const int D = 2 * radius + 1;
    p.flow = flow;
    p.out = out;
    p.B = B;
    p.H = H;
    p.W = W;
    p.HW = H * W;
    p.radius = radius;
    p.groups_per_b = (p.HW + 31) / 32;
    p.total_groups = B * p.groups_per_b;
    p.nr_max = D + 2;
    p.nv_max = (D + 2 + 3 + 3) / 4;
    // every window's base must stay 16-byte aligned for cp.async; an odd number of 16-byte units
    // per window makes consecutive lanes walk through all eight 4-bank groups.
    p.qstride = ((p.nr_max * p.nv_max) | 1) * 4;
    const size_t per_warp = (size_t)(32 * p.qstride + 4 * 32) * sizeof(float);
    // as many warps per block as fit in ~110 KB (two blocks per SM), at most 4
    int wpb = (int)((110 * 1024) / per_warp);
    wpb = wpb < 1 ? 1 : (wpb > 4 ? 4 : wpb);
    const size_t smem = per_warp * wpb;
    PP_CHECK_ARG(smem <= 227 * 1024, "pp_corr_lookup: radius %d needs %zu B of shared memory", radius, smem);
    int blocks_per_sm = (int)((227 * 1024) / (smem + 1024));
    blocks_per_sm = blocks_per_sm < 1 ? 1 : (blocks_per_sm > 8 ? 8 : blocks_per_sm);
    long long want = ((long long)p.total_groups + wpb - 1) / wpb;
    long long cap = (long long)sm_count() * blocks_per_sm * 4;  // grid-stride beyond 4 waves
    int grid = (int)(want < cap ? want : cap);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (radius) {
        case 1: return launch_lookup<1>(p, wpb, smem, grid, st);
        case 2: return launch_lookup<2>(p, wpb, smem, grid, st);
        case 3: return launch_lookup<3>(p, wpb, smem, grid, st);
        case 4: return launch_lookup<4>(p, wpb, smem, grid, st);
        case 5: return launch_lookup<5>(p, wpb, smem, grid, st);
        case 6: return launch_lookup<6>(p, wpb, smem, grid, st);
        case 7: return launch_lookup<7>(p, wpb, smem, grid, st);
        case 8: return launch_lookup<8>(p, wpb, smem, grid, st);
        default: return launch_lookup<0>(p, wpb, smem, grid, st);
    }
}

extern "C" int pp_bilinear_sample(const float* feat, const float* grid, int N, int C, int Hf, int Wf, int Ho,
                                  int Wo, int grid_chw, int align_corners, int scale, float* out, void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (N == 0) return PP_OK;
    PP_CHECK_ARG(feat && grid && out, "pp_bilinear_sample: null pointer");
    PP_CHECK_ARG(N >= 0 && C > 0 && Hf > 0 && Wf > 0 && Ho > 0 && Wo > 0, "pp_bilinear_sample: bad shape");
    if (N == 0) return PP_OK;
    const long long total = (long long)N * Ho * Wo;
    int grid_dim = (int)((total + 127) / 128);
    const int cap = sm_count() * 16;
    if (grid_dim > cap) grid_dim = cap;
    bilinear_sample_kernel<<<grid_dim, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        feat, grid, N, C, Hf, Wf, Ho, Wo, grid_chw, align_corners, scale, out);
    PP_LAUNCHED();
    return PP_OK;
}
