// Stage-1 operand preparation ("prologue"): cast to bf16 (optionally split into 2 or 3 bf16 terms for
// the fp32-accurate modes), transpose to K-major for tcgen05, and compute 1/max(||x||_2, eps) per
// patch.  Together with the GEMM epilogue (which multiplies by the two inverse norms) this replaces
// the F.normalize + rearrange lines of the reference, utils/matching.py:13-14,18-19,40-44:
//      sim[t,s] = <a_t, b_s> / (max(|a_t|,eps) * max(|b_s|,eps))
//
//   in  : feats (G, C, P) fp32, P contiguous (the reference's "b c (h w)" view)
//   out : prep  (G, P, Kp) bf16, Kp = roundup(nseg*C, 64); segment j of a row holds split term
//         SEG_Q[j] (query side) or SEG_B[j] (bank side) so that  sum_j q_seg[j] . b_seg[j]
//         = (q3.b1 + q2.b2 + q1.b3 +) q2.b1 + q1.b2 + q1.b1   -- the classic bf16xN emulation, SMALLEST terms first:
//         the tensor core adds into its fp32 accumulator with truncation, so every MMA step costs up to one ulp of the
//         running sum; with the dominant q1.b1 segment last only its C/16 steps pay that at full magnitude (measured
//         at C = 1024: worst error 1.0e-5 with the dominant segment first);
//         rnorm (G, P) fp32.
//
// HBM-bound streaming kernel: one block = (group g, 32 patches) walks the channels in chunks of 64.
// Loads are coalesced 128-byte rows (thread = patch, 8 channels per thread per chunk, next chunk
// prefetched into registers), the 64 x 32 chunk is transposed through an XOR-swizzled shared tile and
// every global store is a full 128-byte line of the K-major output.  4 bytes read + 2*nseg written
// per feature element, each exactly once.
#include "pp_common.cuh"

namespace pp {

// the LAST nseg entries are used (3 for bf16x3, 6 for fp32): the table ends with the dominant term
__constant__ int SEG_Q[6] = {2, 1, 0, 1, 0, 0};
__constant__ int SEG_B[6] = {0, 1, 2, 0, 1, 0};

static inline int mode_segments(int mode) { return mode == PP_MODE_BF16 ? 1 : (mode == PP_MODE_BF16X3 ? 3 : 6); }
static inline int mode_parts(int mode) { return mode == PP_MODE_BF16 ? 1 : (mode == PP_MODE_BF16X3 ? 2 : 3); }

constexpr int PREP_THREADS = 256;
constexpr int PREP_WARPS = PREP_THREADS / 32;
constexpr int PREP_PT = 32;  // patches per block
constexpr int PREP_CT = 64;  // channels per transposed chunk

// Query-mask bookkeeping (QM instances only).  mask (B,Hm,Wm) is nearest-resized to H x W (F.interpolate,
// utils/matching.py:38-39) and the unmasked patches are ranked in order:
//   mrow[b,t]   resized mask value            rank[b,t]  compact row of patch t, -1 if masked
//   rowmap[b,r] patch of compact row r         tv[b]      number of unmasked patches
//   fm[b]       first masked patch, -1 if none
struct QueryMaskArgs {
    const float* mask;
    int Hm, Wm, H, W;
    float* mrow;
    int* rank;
    int* rowmap;
    int* tv;
    int* fm;
    uint4* clear;                   // optional scratch to zero in the same launch (the contraction's key arrays)
    unsigned long long clear_n16;   // its size in 16-byte units
};

// feats2 / g_split: groups g >= g_split read feats2[g - g_split] and are written as the BANK side of the split pairing --
// one launch prepares both operands of a product (query rows first, template rows after them in the same output).
template <int NPARTS, bool QM>
__global__ void __launch_bounds__(PREP_THREADS)
match_prepare_kernel(const float* __restrict__ feats, int C, int P, int Kp, int nseg, int is_query,
                     __nv_bfloat16* __restrict__ prep, float* __restrict__ rnorm, const QueryMaskArgs qm,
                     const float* __restrict__ feats2 = nullptr, int g_split = 0x7fffffff) {
    // QM: row compaction -- patch p of detection g is written to row rank(p) = number of unmasked patches before
    // it (skipped if masked), so masked query patches never reach the tensor cores.  Every block recounts the
    // detection's mask itself (P values), so the prologue stays a single launch with no inter-block dependency.
    // [buffer][part][patch][64 channels] bf16, 16-byte units XOR-swizzled by (patch & 7)
    __shared__ __align__(16) __nv_bfloat16 s_tile[2][NPARTS][PREP_PT][PREP_CT];
    __shared__ float s_part[PREP_WARPS][PREP_PT];
    __shared__ int s_rank[PREP_PT];
    __shared__ int s_cnt[PREP_WARPS][2];
    __shared__ int s_fm;

    const int g = blockIdx.y;
    const int p0 = blockIdx.x * PREP_PT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // programmatic dependent launch (bank instance): nothing is read or written before the previous kernel on the stream --
    // which may have produced the features, or may still be reading this call's output buffers -- has completed
    grid_dependency_wait();
    int tv_total = 0;
    if (QM) {
        if (qm.clear) {
            // zero the caller's scratch while we are here: saves a memset node in front of the contraction
            const unsigned long long nth = (unsigned long long)gridDim.x * gridDim.y * PREP_THREADS;
            for (unsigned long long i = ((unsigned long long)blockIdx.y * gridDim.x + blockIdx.x) * PREP_THREADS + threadIdx.x;
                 i < qm.clear_n16; i += nth)
                qm.clear[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        auto mask_at = [&](int t) {
            const int y = t / qm.W, xx = t - y * qm.W;
            return __ldg(qm.mask + ((size_t)g * qm.Hm + nearest_src(y, qm.Hm, qm.H)) * qm.Wm + nearest_src(xx, qm.Wm, qm.W));
        };
        if (threadIdx.x == 0) s_fm = 0x7fffffff;
        __syncthreads();
        int before = 0, total = 0, first_masked = 0x7fffffff;   // unmasked patches in [0, p0) and in [0, P)
        for (int t = threadIdx.x; t < P; t += PREP_THREADS) {
            const bool on = mask_at(t) != 0.f;
            total += on;
            before += on && t < p0;
            if (!on) first_masked = min(first_masked, t);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            before += __shfl_xor_sync(0xffffffffu, before, o);
            total += __shfl_xor_sync(0xffffffffu, total, o);
            first_masked = min(first_masked, __shfl_xor_sync(0xffffffffu, first_masked, o));
        }
        if (lane == 0) {
            s_cnt[warp][0] = before;
            s_cnt[warp][1] = total;
            atomicMin(&s_fm, first_masked);
        }
        __syncthreads();
        before = 0;
#pragma unroll
        for (int w = 0; w < PREP_WARPS; ++w) {
            before += s_cnt[w][0];
            tv_total += s_cnt[w][1];
        }
        if (warp == 0) {
            const int t = p0 + lane;
            const float m = t < P ? mask_at(t) : 0.f;
            const bool on = t < P && m != 0.f;
            const unsigned ball = __ballot_sync(0xffffffffu, on);
            const int r = before + __popc(ball & ((1u << lane) - 1u));
            s_rank[lane] = on ? r : -1;
            if (t < P) {
                qm.mrow[(size_t)g * P + t] = m;
                qm.rank[(size_t)g * P + t] = on ? r : -1;
                if (on) qm.rowmap[(size_t)g * P + r] = t;
            }
            if (blockIdx.x == 0 && lane == 0) {
                qm.tv[g] = tv_total;
                qm.fm[g] = s_fm == 0x7fffffff ? -1 : s_fm;
            }
        }
        __syncthreads();
        // compact rows past the last unmasked patch, up to the next 256-row tile boundary, are read by the
        // contraction's TMA but stand for no patch: keep them zero (blocks share the <= 255 rows round-robin)
        const int tv_pad = min(P, (tv_total + 255) / 256 * 256);
        __nv_bfloat16* og = prep + (size_t)g * P * Kp;
        for (int r = tv_total + blockIdx.x; r < tv_pad; r += gridDim.x) {
            for (int i = threadIdx.x * 8; i < Kp; i += PREP_THREADS * 8)
                *reinterpret_cast<uint4*>(og + (size_t)r * Kp + i) = make_uint4(0u, 0u, 0u, 0u);
            if (threadIdx.x == 0) rnorm[(size_t)g * P + r] = 0.f;
        }
    }
    const bool live = p0 + lane < P;
    if (!QM && g >= g_split) is_query = 0;
    const float* x = ((!QM && g >= g_split) ? feats2 + (size_t)(g - g_split) * C * P : feats + (size_t)g * C * P) + (live ? p0 + lane : P - 1);
    __nv_bfloat16* out_g = prep + (size_t)g * P * Kp;
    const int* seg_tab = (is_query ? SEG_Q : SEG_B) + (6 - nseg);
    const int nchunks = (C + PREP_CT - 1) / PREP_CT;

    // thread owns patch `lane` and channels 64*j + 8*warp + i of chunk j; loads run two chunks ahead of the
    // transposition (three register buffers, rotated by a 3x unrolled loop) to cover the global latency
    const size_t strideP = (size_t)P;
    auto load_chunk = [&](float (&dst)[8], int j) {
        if (j >= nchunks) return;                                    // block-uniform
        const float* xc = x + (size_t)(j * PREP_CT + warp * 8) * strideP;
        if ((j + 1) * PREP_CT <= C) {                                // full chunk: no per-channel bounds checks
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = __ldg(xc + i * strideP);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = (j * PREP_CT + warp * 8 + i < C) ? __ldg(xc + i * strideP) : 0.f;
        }
    };
    // QM: the query operand is written NORMALISED and multiplied by its mask value, x * m / max(||x||, 1e-12)
    // (utils/matching.py:40-41,48), so the contraction's epilogue has no per-row factor to apply.  That needs the norm
    // before the first store: a first pass over the patch's channels (the second one hits L2; query features are a few
    // MB per detection, the kernel is latency-bound and runs beside the bank prologue).
    float qscale = 1.0f;
    if (QM) {
        float s2 = 0.f;
        float pa[8], pb[8];
        load_chunk(pa, 0);
        for (int j = 0; j < nchunks; j += 2) {
            load_chunk(pb, j + 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) s2 = fmaf(pa[i], pa[i], s2);
            if (j + 1 < nchunks) {
                load_chunk(pa, j + 2);
#pragma unroll
                for (int i = 0; i < 8; ++i) s2 = fmaf(pb[i], pb[i], s2);
            }
        }
        s_part[warp][lane] = s2;
        __syncthreads();
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < PREP_WARPS; ++w) t += s_part[w][lane];
        const float m = live ? qm.mrow[(size_t)g * P + p0 + lane] : 0.f;   // written by warp 0 above (before a __syncthreads)
        qscale = m / fmaxf(sqrtf(t), 1e-12f);
        __syncthreads();  // s_part is reused for the epilogue of this kernel
    }
    float ss = 0.f;
    auto process = [&](const float (&cur)[8], int j) {
        const int c0 = j * PREP_CT;
        __align__(16) __nv_bfloat16 part[NPARTS][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = QM ? cur[i] * qscale : cur[i];
            ss = fmaf(v, v, ss);
#pragma unroll
            for (int k = 0; k < NPARTS; ++k) {
                const __nv_bfloat16 h = __float2bfloat16_rn(v);
                part[k][i] = h;
                v -= __bfloat162float(h);  // exact: the residual of a bf16 rounding is representable
            }
        }
        const int unit = warp ^ (lane & 7);
#pragma unroll
        for (int k = 0; k < NPARTS; ++k)
            *reinterpret_cast<uint4*>(&s_tile[j & 1][k][lane][unit * 8]) = *reinterpret_cast<const uint4*>(part[k]);
        __syncthreads();  // also orders buffer reuse: writes of chunk j+2 come after everyone's reads of chunk j
        // write-out: a warp stores 4 patch rows x 128 bytes per instruction
        const int n_units = min(PREP_CT, C - c0) / 8;  // valid 16-byte units in this chunk (C % 8 == 0)
        const int row = warp * 4 + (lane >> 3);
        const int u = lane & 7;
        if (p0 + row < P && u < n_units) {
            const int orow = QM ? s_rank[row] : p0 + row;
            if (orow >= 0) {
                __nv_bfloat16* dst = out_g + (size_t)orow * Kp + c0 + u * 8;
                if (NPARTS == 1) {  // bf16 mode: one segment
                    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(&s_tile[j & 1][0][row][(u ^ (row & 7)) * 8]);
                } else {
                    for (int s = 0; s < nseg; ++s) {
                        const uint4 val = *reinterpret_cast<const uint4*>(&s_tile[j & 1][seg_tab[s]][row][(u ^ (row & 7)) * 8]);
                        *reinterpret_cast<uint4*>(dst + (size_t)s * C) = val;
                    }
                }
            }
        }
    };
    float b0[8], b1[8], b2[8];
    load_chunk(b0, 0);
    load_chunk(b1, 1);
    for (int j = 0; j < nchunks; j += 3) {
        load_chunk(b2, j + 2);
        process(b0, j);
        if (j + 1 < nchunks) {
            load_chunk(b0, j + 3);
            process(b1, j + 1);
        }
        if (j + 2 < nchunks) {
            load_chunk(b1, j + 4);
            process(b2, j + 2);
        }
    }
    // ---- inverse norm per patch: 1 / max(||x||, 1e-12)  (F.normalize's clamp_min(eps)) ----
    s_part[warp][lane] = ss;
    __syncthreads();
    if (warp == 0 && live) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < PREP_WARPS; ++w) t += s_part[w][lane];
        const int orow = QM ? s_rank[lane] : p0 + lane;
        if (orow >= 0) rnorm[(size_t)g * P + orow] = QM ? 1.0f : 1.0f / fmaxf(sqrtf(t), 1e-12f);   // QM: already applied
    }
    // ---- zero the K padding [nseg*C, Kp) ----
    const int pad0 = nseg * C, npad = Kp - pad0;
    if (npad > 0) {
        for (int i = threadIdx.x; i < PREP_PT * npad; i += PREP_THREADS) {
            const int row = i / npad, k = i - row * npad;
            if (p0 + row < P) {
                const int orow = QM ? s_rank[row] : p0 + row;
                if (orow >= 0) out_g[(size_t)orow * Kp + pad0 + k] = __float2bfloat16_rn(0.f);
            }
        }
    }
}

}  // namespace pp

extern "C" int pp_match_kp(int C, int mode) {
    if (C <= 0 || mode < 0 || mode > 2) return PP_ERR_ARG;
    const int k = pp::mode_segments(mode) * C;
    return (k + 63) / 64 * 64;
}

extern "C" int pp_match_prepare(const float* feats, int64_t G, int C, int P, int mode, int is_query, void* prepared,
                                float* rnorm, void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (G == 0) return PP_OK;
    PP_CHECK_ARG(feats && prepared && rnorm, "pp_match_prepare: null pointer");
    PP_CHECK_ARG(mode >= 0 && mode <= 2, "pp_match_prepare: unknown mode %d", mode);
    PP_CHECK_ARG(C > 0 && C % 8 == 0, "pp_match_prepare: feature dim must be a positive multiple of 8 (got %d)", C);
    PP_CHECK_ARG(P > 0 && G > 0, "pp_match_prepare: bad shape (G=%lld, P=%d)", (long long)G, P);
    PP_CHECK_ARG((reinterpret_cast<uintptr_t>(prepared) & 15) == 0, "pp_match_prepare: output must be 16-byte aligned");
    const int Kp = pp_match_kp(C, mode);
    const int nseg = mode_segments(mode), nparts = mode_parts(mode);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // grid.y is limited to 65535: slice the groups
    const int64_t GY = 65535;
    for (int64_t g0 = 0; g0 < G; g0 += GY) {
        const int gy = (int)((G - g0) < GY ? (G - g0) : GY);
        dim3 grid((P + PREP_PT - 1) / PREP_PT, gy);
        const float* f = feats + (size_t)g0 * C * P;
        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(prepared) + (size_t)g0 * P * Kp;
        float* rn = rnorm + (size_t)g0 * P;
        const QueryMaskArgs none{};
        const float* no_feats2 = nullptr;
        const int no_split = 0x7fffffff;
        if (nparts == 1)
            PP_CUDA(launch_dependent(match_prepare_kernel<1, false>, grid, dim3(PREP_THREADS), 0, st, f, C, P, Kp, nseg, is_query, o, rn,
                                     none, no_feats2, no_split));
        else if (nparts == 2)
            PP_CUDA(launch_dependent(match_prepare_kernel<2, false>, grid, dim3(PREP_THREADS), 0, st, f, C, P, Kp, nseg, is_query, o, rn,
                                     none, no_feats2, no_split));
        else
            PP_CUDA(launch_dependent(match_prepare_kernel<3, false>, grid, dim3(PREP_THREADS), 0, st, f, C, P, Kp, nseg, is_query, o, rn,
                                     none, no_feats2, no_split));
        PP_LAUNCHED();
    }
    return PP_OK;
}

namespace pp {
// query features (G, C, P) and template features (G, C, P) -> prepared (2G, P, Kp) [queries | templates], rnorm (2G, P): one launch
int prepare_pair_impl(const float* q_feats, const float* s_feats, int64_t G, int C, int P, int mode, void* prepared, float* rnorm,
                      void* stream) {
    if (int rc = require_sm100()) return rc;
    PP_CHECK_ARG(q_feats && s_feats && prepared && rnorm, "prepare_pair: null pointer");
    PP_CHECK_ARG(mode >= 0 && mode <= 2, "prepare_pair: unknown mode %d", mode);
    PP_CHECK_ARG(C > 0 && C % 8 == 0 && P > 0 && G > 0 && 2 * G <= 65535, "prepare_pair: bad shape (G=%lld, C=%d, P=%d)", (long long)G, C, P);
    const int Kp = pp_match_kp(C, mode);
    const int nseg = mode_segments(mode), nparts = mode_parts(mode);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((P + PREP_PT - 1) / PREP_PT, (unsigned)(2 * G));
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(prepared);
    const QueryMaskArgs none{};
    if (nparts == 1) match_prepare_kernel<1, false><<<grid, PREP_THREADS, 0, st>>>(q_feats, C, P, Kp, nseg, 1, o, rnorm, none, s_feats, (int)G);
    else if (nparts == 2) match_prepare_kernel<2, false><<<grid, PREP_THREADS, 0, st>>>(q_feats, C, P, Kp, nseg, 1, o, rnorm, none, s_feats, (int)G);
    else match_prepare_kernel<3, false><<<grid, PREP_THREADS, 0, st>>>(q_feats, C, P, Kp, nseg, 1, o, rnorm, none, s_feats, (int)G);
    PP_LAUNCHED();
    return PP_OK;
}

QueryMeta split_query_meta(void* q_meta, int B, int T) {
    QueryMeta m;
    char* p = static_cast<char*>(q_meta);
    const size_t bt = (size_t)B * T * 4;
    m.mrow = reinterpret_cast<float*>(p);
    m.rank = reinterpret_cast<int*>(p + bt);
    m.rowmap = reinterpret_cast<int*>(p + 2 * bt);
    m.tv = reinterpret_cast<int*>(p + 3 * bt);
    m.fm = reinterpret_cast<int*>(p + 3 * bt + (size_t)B * 4);
    return m;
}
}  // namespace pp

extern "C" size_t pp_match_query_meta_bytes(int B, int T) {
    if (B < 0 || T < 0) return 0;
    return ((size_t)3 * B * T + 2 * (size_t)B) * 4;
}

namespace pp {
int prepare_query_impl(const float* tar_feat, const float* tar_mask, int B, int C, int H, int W, int Hm, int Wm, int mode,
                       void* q_prep, float* q_rnorm, void* q_meta, void* clear, size_t clear_bytes, void* stream);
}

extern "C" int pp_match_prepare_query(const float* tar_feat, const float* tar_mask, int B, int C, int H, int W, int Hm,
                                      int Wm, int mode, void* q_prep, float* q_rnorm, void* q_meta, void* stream) {
    return pp::prepare_query_impl(tar_feat, tar_mask, B, C, H, W, Hm, Wm, mode, q_prep, q_rnorm, q_meta, nullptr, 0, stream);
}

// `clear` (16-byte aligned, clear_bytes a multiple of 16) is zeroed by the same launch when given
int pp::prepare_query_impl(const float* tar_feat, const float* tar_mask, int B, int C, int H, int W, int Hm, int Wm, int mode,
                           void* q_prep, float* q_rnorm, void* q_meta, void* clear, size_t clear_bytes, void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (B == 0) return PP_OK;
    PP_CHECK_ARG(tar_feat && tar_mask && q_prep && q_rnorm && q_meta, "pp_match_prepare_query: null pointer");
    PP_CHECK_ARG(mode >= 0 && mode <= 2, "pp_match_prepare_query: unknown mode %d", mode);
    PP_CHECK_ARG(C > 0 && C % 8 == 0, "pp_match_prepare_query: feature dim must be a positive multiple of 8 (got %d)", C);
    PP_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && Hm > 0 && Wm > 0, "pp_match_prepare_query: bad shape");
    PP_CHECK_ARG((reinterpret_cast<uintptr_t>(q_prep) & 15) == 0 && (reinterpret_cast<uintptr_t>(q_meta) & 3) == 0,
                 "pp_match_prepare_query: misaligned output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int T = H * W;
    const int Kp = pp_match_kp(C, mode);
    const QueryMeta m = split_query_meta(q_meta, B, T);
    // one launch: mask resize + compaction bookkeeping + cast/transposition of the unmasked patches + inverse norms
    PP_CHECK_ARG(!clear || ((reinterpret_cast<uintptr_t>(clear) & 15) == 0 && clear_bytes % 16 == 0),
                 "pp_match_prepare_query: scratch to clear must be 16-byte aligned and sized");
    const QueryMaskArgs qa{tar_mask, Hm, Wm, H, W, m.mrow, m.rank, m.rowmap, m.tv, m.fm, static_cast<uint4*>(clear),
                           (unsigned long long)(clear ? clear_bytes / 16 : 0)};
    const int nseg = mode_segments(mode), nparts = mode_parts(mode);
    dim3 grid((T + PREP_PT - 1) / PREP_PT, B);
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(q_prep);
    if (nparts == 1) match_prepare_kernel<1, true><<<grid, PREP_THREADS, 0, st>>>(tar_feat, C, T, Kp, nseg, 1, o, q_rnorm, qa);
    else if (nparts == 2) match_prepare_kernel<2, true><<<grid, PREP_THREADS, 0, st>>>(tar_feat, C, T, Kp, nseg, 1, o, q_rnorm, qa);
    else match_prepare_kernel<3, true><<<grid, PREP_THREADS, 0, st>>>(tar_feat, C, T, Kp, nseg, 1, o, q_rnorm, qa);
    PP_LAUNCHED();
    return PP_OK;
}
