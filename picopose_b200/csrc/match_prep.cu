// Stage-1 operand preparation ("prologue"): L2-normalise over C, cast to bf16 (optionally split into
// 2 or 3 bf16 terms for the fp32-accurate modes) and transpose to K-major for tcgen05.
// Replaces the F.normalize + rearrange lines of the reference, utils/matching.py:13-14,18-19,40-44.
//
//   in  : feats (G, C, P) fp32, P contiguous (the reference's "b c (h w)" view)
//   out : prep  (G, P, Kp) bf16, Kp = roundup(nseg*C, 64); segment j of a row holds split term
//         SEG_Q[j] (query side) or SEG_B[j] (bank side) so that  sum_j q_seg[j] . b_seg[j]
//         = q1.b1 + q1.b2 + q2.b1 (+ q1.b3 + q2.b2 + q3.b1)   -- the classic bf16xN emulation.
//
// HBM-bound: one block = (group g, 32 patches); its (C x 32) slab is read once into registers
// (coalesced 128-byte rows), the norm is reduced across warps, then 64-channel chunks are scaled, converted and transposed
// through a swizzled shared tile so every global store is a 128-byte line.
#include "pp_common.cuh"

namespace pp {

__constant__ int SEG_Q[6] = {0, 0, 1, 0, 1, 2};
__constant__ int SEG_B[6] = {0, 1, 0, 2, 1, 0};

static inline int mode_segments(int mode) { return mode == PP_MODE_BF16 ? 1 : (mode == PP_MODE_BF16X3 ? 3 : 6); }
static inline int mode_parts(int mode) { return mode == PP_MODE_BF16 ? 1 : (mode == PP_MODE_BF16X3 ? 2 : 3); }

constexpr int PREP_THREADS = 256;
constexpr int PREP_WARPS = PREP_THREADS / 32;
constexpr int PREP_PT = 32;  // patches per block
constexpr int PREP_CT = 64;  // channels per transposed chunk

// NCHUNK > 0: the block's whole (C x 32) slab lives in registers (8*NCHUNK floats per thread, C <= 64*NCHUNK),
// so HBM is read exactly once.  NCHUNK == 0: generic two-pass variant for larger C (second pass hits L2).
template <int NCHUNK>
__global__ void __launch_bounds__(PREP_THREADS)
match_prepare_kernel(const float* __restrict__ feats, int C, int P, int Kp, int nseg, int nparts, int is_query,
                     __nv_bfloat16* __restrict__ prep) {
    __shared__ float s_part[PREP_WARPS][PREP_PT];
    __shared__ float s_den[PREP_PT];
    // [part][patch][64 channels] bf16, 16-byte units XOR-swizzled by (patch & 7)
    __shared__ __align__(16) __nv_bfloat16 s_tile[3][PREP_PT][PREP_CT];

    const int g = blockIdx.y;
    const int p0 = blockIdx.x * PREP_PT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = p0 + lane;
    const bool live = p < P;
    const float* x = feats + (size_t)g * C * P + (live ? p : P - 1);

    // thread owns patch `lane` and channels {64*j + 8*warp + i}: exactly what it transposes in pass 2
    constexpr int NV = NCHUNK > 0 ? NCHUNK : 1;
    float vals[NV][8];
    float ss = 0.f;
    if (NCHUNK > 0) {
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = j * PREP_CT + warp * 8 + i;
                vals[j][i] = c < C ? __ldg(x + (size_t)c * P) : 0.f;
            }
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) ss = fmaf(vals[j][i], vals[j][i], ss);
    } else {
        for (int c = warp; c < C; c += PREP_WARPS) {
            const float v = __ldg(x + (size_t)c * P);
            ss = fmaf(v, v, ss);
        }
    }
    s_part[warp][lane] = ss;
    __syncthreads();
    if (warp == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < PREP_WARPS; ++w) t += s_part[w][lane];
        s_den[lane] = fmaxf(sqrtf(t), 1e-12f);  // F.normalize: x / max(||x||, eps)
    }
    __syncthreads();
    const float den = s_den[lane];

    __nv_bfloat16* out_g = prep + (size_t)g * P * Kp;
    const int* seg_tab = is_query ? SEG_Q : SEG_B;
    const int nchunks = (C + PREP_CT - 1) / PREP_CT;

#pragma unroll
    for (int j = 0; j < (NCHUNK > 0 ? NCHUNK : 1 << 20); ++j) {
        if (j >= nchunks) break;
        const int c0 = j * PREP_CT;
        __align__(16) __nv_bfloat16 part[3][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = c0 + warp * 8 + i;
            float v;
            if (NCHUNK > 0) v = vals[NCHUNK > 0 ? j : 0][i];
            else v = c < C ? __ldg(x + (size_t)c * P) : 0.f;
            v = __fdiv_rn(v, den);
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v);
            part[0][i] = h0;
            float r = v - __bfloat162float(h0);
            const __nv_bfloat16 h1 = __float2bfloat16_rn(r);
            part[1][i] = h1;
            r = r - __bfloat162float(h1);
            part[2][i] = __float2bfloat16_rn(r);
        }
        const int unit = warp ^ (lane & 7);
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k < nparts)
                *reinterpret_cast<uint4*>(&s_tile[k][lane][unit * 8]) = *reinterpret_cast<const uint4*>(part[k]);
        __syncthreads();
        // write-out: a warp stores 4 patch rows x 128 bytes per instruction
        const int n_units = min(PREP_CT, C - c0) / 8;  // valid 16-byte units in this chunk (C % 8 == 0)
        for (int s = 0; s < nseg; ++s) {
            const int k = seg_tab[s];
            for (int row = warp * 4 + (lane >> 3); row < PREP_PT; row += PREP_WARPS * 4) {
                const int u = lane & 7;
                const int pp = p0 + row;
                if (pp < P && u < n_units) {
                    const uint4 val = *reinterpret_cast<const uint4*>(&s_tile[k][row][(u ^ (row & 7)) * 8]);
                    *reinterpret_cast<uint4*>(out_g + (size_t)pp * Kp + (size_t)s * C + c0 + u * 8) = val;
                }
            }
        }
        __syncthreads();
    }
    // ---- zero the K padding [nseg*C, Kp) ----
    const int pad0 = nseg * C, npad = Kp - pad0;
    if (npad > 0) {
        for (int i = threadIdx.x; i < PREP_PT * npad; i += PREP_THREADS) {
            const int row = i / npad, k = i - row * npad;
            if (p0 + row < P) out_g[(size_t)(p0 + row) * Kp + pad0 + k] = __float2bfloat16_rn(0.f);
        }
    }
}

template <int NCHUNK>
static int launch_prepare(dim3 grid, cudaStream_t st, const float* feats, int C, int P, int Kp, int nseg, int nparts,
                          int is_query, __nv_bfloat16* prep) {
    match_prepare_kernel<NCHUNK><<<grid, PREP_THREADS, 0, st>>>(feats, C, P, Kp, nseg, nparts, is_query, prep);
    PP_LAUNCHED();
    return PP_OK;
}

}  // namespace pp

extern "C" int pp_match_kp(int C, int mode) {
    if (C <= 0 || mode < 0 || mode > 2) return PP_ERR_ARG;
    const int k = pp::mode_segments(mode) * C;
    return (k + 63) / 64 * 64;
}

extern "C" int pp_match_prepare(const float* feats, int64_t G, int C, int P, int mode, int is_query, void* prepared,
                                void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (G == 0) return PP_OK;
    PP_CHECK_ARG(feats && prepared, "pp_match_prepare: null pointer");
    PP_CHECK_ARG(mode >= 0 && mode <= 2, "pp_match_prepare: unknown mode %d", mode);
    PP_CHECK_ARG(C > 0 && C % 8 == 0, "pp_match_prepare: feature dim must be a positive multiple of 8 (got %d)", C);
    PP_CHECK_ARG(P > 0 && G >= 0 && G <= 65535LL * 64, "pp_match_prepare: bad shape (G=%lld, P=%d)", (long long)G, P);
    PP_CHECK_ARG((reinterpret_cast<uintptr_t>(prepared) & 15) == 0, "pp_match_prepare: output must be 16-byte aligned");
    if (G == 0) return PP_OK;
    const int Kp = pp_match_kp(C, mode);
    const int nseg = mode_segments(mode), nparts = mode_parts(mode);
    // grid.y is limited to 65535: slice the groups
    const int64_t GY = 65535;
    for (int64_t g0 = 0; g0 < G; g0 += GY) {
        const int gy = (int)((G - g0) < GY ? (G - g0) : GY);
        dim3 grid((P + PREP_PT - 1) / PREP_PT, gy);
        const float* f = feats + (size_t)g0 * C * P;
        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(prepared) + (size_t)g0 * P * Kp;
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        const int nch = (C + PREP_CT - 1) / PREP_CT;
        int rc;
        if (nch <= 1) rc = launch_prepare<1>(grid, st, f, C, P, Kp, nseg, nparts, is_query, o);
        else if (nch <= 2) rc = launch_prepare<2>(grid, st, f, C, P, Kp, nseg, nparts, is_query, o);
        else if (nch <= 4) rc = launch_prepare<4>(grid, st, f, C, P, Kp, nseg, nparts, is_query, o);
        else if (nch <= 6) rc = launch_prepare<6>(grid, st, f, C, P, Kp, nseg, nparts, is_query, o);
        else if (nch <= 8) rc = launch_prepare<8>(grid, st, f, C, P, Kp, nseg, nparts, is_query, o);
        else if (nch <= 12) rc = launch_prepare<12>(grid, st, f, C, P, Kp, nseg, nparts, is_query, o);
        else if (nch <= 16) rc = launch_prepare<16>(grid, st, f, C, P, Kp, nseg, nparts, is_query, o);
        else rc = launch_prepare<0>(grid, st, f, C, P, Kp, nseg, nparts, is_query, o);
        if (rc) return rc;
    }
    return PP_OK;
}
