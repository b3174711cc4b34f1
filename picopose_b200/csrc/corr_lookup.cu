// Stage-3 correlation-window lookup and the generic bilinear sampler (sm_100a).
//
// pp_corr_lookup replaces CorrLookup.forward (reference utils/corr_lookup.py:100-134): for every
// query pixel it samples a (2r+1)^2 bilinear window of its own slice of each pyramid level.
//
// Kernel shape (HBM-bound gather):
//   * one warp owns 32 consecutive queries (flat h*W+w) of one detection, so every output channel
//     store is one coalesced 128-byte line in the reference's (B, L*D*D, H, W) layout;
//   * the window is processed in bands of JB y-taps.  Per band the warp stages, with 16-byte
//     cp.async (zero-filled outside the map), the <= (JB+2) x (D+2) footprint of each of its 32
//     windows into shared memory -- consecutive lanes fetch consecutive 16-byte pieces of a row and
//     only pieces some tap really touches are requested;
//   * then lane = query: each window sample gathers its 4 taps from shared memory.
// Coordinates follow the reference's float arithmetic op by op (x*2/(W-1)-1 and grid_sample's
// inverse, utils/corr_lookup.py:61-65 + ATen grid_sampler_unnormalize), so floor/weights agree.
#include "pp_common.cuh"
#include "corr_lookup.cuh"

#include <cstdlib>
#include <cstring>

namespace pp {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
// 16-byte copy that reads only `src_bytes` (0 or 16) from global memory and zero-fills the rest
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, uint32_t src_bytes) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// Fast path: compile-time radius R, window processed in bands of JB y-taps, all levels 16-byte loadable.
// ------------------------------------------------------------------------------------------------
template <int R, int JB>
struct LookupCfg {
    static constexpr int D = 2 * R + 1;
    static constexpr int NB = (D + JB - 1) / JB;      // bands
    static constexpr int NRB = JB + 2;                // staged rows per band (floor jitter of +-1 included)
    static constexpr int NV = (D + 5 + 3) / 4;        // 16-byte pieces per row: D+2 columns + 3 alignment slack
    static constexpr int PITCH = NV * 4;              // words per staged row
    static constexpr int QS = ((NRB * NV) | 1) * 4;   // words per query: odd number of 16-byte units (bank spread)
    static constexpr int WARP_WORDS = 32 * QS;
    // per query a table of NRB row offsets + NV piece offsets (odd pitch: the owners' stores spread over banks)
    static constexpr int TW = (NRB + NV) | 1;
    static constexpr int WARP_WORDS_TAB = WARP_WORDS + 32 * TW;
};

// table entries: an element offset (>= 0), "outside the map: zero-fill", or "no tap touches it: skip"
constexpr int TAB_OOB = (int)0x80000000u;
constexpr int TAB_SKIP = (int)0xC0000000u;

// TILED: every (Hl x Wl) slice is stored as 4-row x 8-column tiles, one 128-byte line per tile (our own
// CorrelationPyramid writes it that way on request).  DRAM moves whole 128-byte lines on this GPU
// (profiles/r1w_dram_granularity.md), and a (2r+2)^2 window crosses ((2r+1)/8 + 1) * ((2r+1)/4 + 1) tiles on average --
// 6.9 lines at r = 4 -- instead of (2r+2) * 1.17 row segments (11.7 lines) in the reference's row-major slices.  Only the
// global address of a 16-byte piece changes: pieces are 4-aligned in x, tile rows are 8 floats, so a piece never
// straddles two tiles; the staged footprint in shared memory and everything after it are the same.
//
// How the staging loop finds its addresses.  The element offset of (y, x) separates into a row part and a column part
// in both layouts (row-major: y * W + x; tiled: ((y >> 2) * (W >> 3) << 5) + ((y & 3) << 3)  +  ((x >> 3) << 5) + (x & 7)),
// so every lane -- as the owner of its query -- writes the NRB row parts of the band and the NV column parts of its
// footprint into a small table (with "zero-fill" / "skip" encoded as negative values); the staging loop, where lane =
// (row, piece) and the query index is the unrolled loop variable, then needs two shared loads at compile-time offsets, an
// OR, an ADD and two predicates per copy instead of two shuffles, the field unpacking, four range checks and the tile
// arithmetic (the first form of this kernel): ~13 instead of ~30 instructions for each of the 32 x PASSES x NB copies of a
// work item; r = 4 on tiled volumes 0.258 -> 0.224 ms, r = 8 0.885 -> 0.703 ms (profiles/r2z_*).
template <int R, int JB, bool TILED>
__global__ void __launch_bounds__(128) corr_lookup_banded_kernel(const LookupParams p) {
    using Cfg = LookupCfg<R, JB>;
    constexpr int D = Cfg::D, NB = Cfg::NB, NRB = Cfg::NRB, NV = Cfg::NV, PITCH = Cfg::PITCH, QS = Cfg::QS, TW = Cfg::TW;
    extern __shared__ __align__(16) float smem[];
    const int warps_per_block = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    float* stage = smem + (size_t)warp * Cfg::WARP_WORDS_TAB;
    int* tab = reinterpret_cast<int*>(stage + Cfg::WARP_WORDS);  // [32 queries][TW] offset tables
    grid_dependency_wait();  // flow and volumes may come from the kernel launched just before (programmatic dependent launch)

    // work item = (query group, pyramid level): levels of one group run on different warps, which keeps the
    // dependent chain per warp short when there are few queries (the native 16^2..64^2 ladder)
    const int total_items = p.total_groups * p.L;
    for (int item = blockIdx.x * warps_per_block + warp; item < total_items; item += gridDim.x * warps_per_block) {
        const int g = item / p.L;
        const int l = item - g * p.L;
        const int b = g / p.groups_per_b;
        const int hw0 = (g - b * p.groups_per_b) * 32;
        const int hw = hw0 + lane;
        const bool live = hw < p.HW;
        const int hwc = live ? hw : p.HW - 1;
        const int qh = hwc / p.W, qw = hwc - qh * p.W;
        const float fx = __ldg(p.flow + ((size_t)b * 2 + 0) * p.HW + hwc);
        const float fy = __ldg(p.flow + ((size_t)b * 2 + 1) * p.HW + hwc);
        const float cx = __fadd_rn((float)qw, fx);  // coords_grid + flow, utils/corr_lookup.py:113
        const float cy = __fadd_rn((float)qh, fy);
        float* out_b = p.out + (size_t)b * p.L * D * D * p.HW;  // warp-uniform base, 32-bit offsets below

        {
            const int Hl = p.vh[l], Wl = p.vw[l];
            const float inv = 1.0f / (float)(1 << l);  // exact: centroid / 2**l, utils/corr_lookup.py:125
            const float lx = __fmul_rn(cx, inv), ly = __fmul_rn(cy, inv);
            int xo[D], yo[D], yraw[D];
            float xw[D], yw[D];
            // Regular in x: tap a sits exactly a columns right of tap 0 (before clamping; true for all but ~1e-5 of the
            // windows).  The footprint is then staged from the UNCLAMPED origin -- columns outside the map are zero-filled,
            // which is what the clamped taps read anyway -- and a staged row is D + 1 consecutive values from column c0.
            bool regx = true;
            int xraw0 = 0;
#pragma unroll
            for (int a = 0; a < D; ++a) {
                int xraw;
                axis_tap_raw(__fadd_rn(lx, (float)(a - R)), Wl, xo[a], xw[a], xraw);
                if (a == 0) xraw0 = xraw;
                regx = regx && (xraw == xraw0 + a);
                axis_tap_raw(__fadd_rn(ly, (float)(a - R)), Hl, yo[a], yw[a], yraw[a]);
            }
            const int xb = regx ? min(max(xraw0, -(D + 2)), Wl) : xo[0];
            const int xs = xb & ~3;                           // two's complement: rounds towards -inf
            const int c0 = xb - xs;                           // regular: column of tap 0 in the staged row (0..3)
            // 16-byte pieces per row some tap touches
            const int pieces = regx ? ((c0 + D) >> 2) + 1 : ((xo[D - 1] + 1 - xs) >> 2) + 1;
            // slice base of the group's first query; query ql of the group is ql slices further
            const float* vol_g = p.vol[l] + ((size_t)b * p.HW + hw0) * ((size_t)Hl * Wl);
            const uint32_t slice = (uint32_t)Hl * (uint32_t)Wl;
            {
                __syncwarp();  // the previous item's staging loop is done with the table
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const int x = xs + 4 * v;
                    const int xoff = TILED ? ((x >> 3) << 5) + (x & 7) : x;
                    tab[lane * TW + NRB + v] = (!live || v >= pieces) ? TAB_SKIP : ((unsigned)x < (unsigned)Wl ? xoff : TAB_OOB);
                }
            }

#pragma unroll
            for (int band = 0; band < NB; ++band) {
                const int j0 = band * JB;
                const int j1 = (j0 + JB < D) ? j0 + JB : D;  // compile-time after unrolling
                // Regular band: tap j0 + bb sits exactly bb rows below tap j0 (true for all but ~1e-5 of the taps, where the
                // float round trip moves a floor).  Its rows are then staged from the UNCLAMPED origin -- rows outside
                // the map are zero-filled, which is what the clamped taps read anyway -- so that tap bb always uses
                // staged rows bb and bb + 1 and the blend below can walk the rows once.
                bool regular = true;
#pragma unroll
                for (int bb = 1; bb < j1 - j0; ++bb) regular = regular && (yraw[j0 + bb] == yraw[j0] + bb);
                const int yb = regular ? min(max(yraw[j0], -(JB + 1)), Hl) : yo[j0];
                const int rows = regular ? (j1 - j0) + 1 : yo[j1 - 1] + 1 - yb + 1;
                const bool warp_regular = __all_sync(0xffffffffu, (regular && regx) || !live);
                __syncwarp();  // previous band's readers are done with the staging area
                {
#pragma unroll
                    for (int row = 0; row < NRB; ++row) {
                        const int y = yb + row;
                        const int yoff = TILED ? (((y >> 2) * (Wl >> 3)) << 5) + ((y & 3) << 3) : y * Wl;
                        // the row part carries the query's slice too: the staging loop adds nothing but the column part
                        tab[lane * TW + row] = (!live || row >= rows) ? TAB_SKIP
                                               : ((unsigned)y < (unsigned)Hl ? yoff + lane * (int)slice : TAB_OOB);
                    }
                    __syncwarp();
                }

                // ---- cooperative staging: for query ql (the unrolled loop variable) lane -> (row, piece) of its footprint, so
                // consecutive lanes fetch consecutive 16-byte pieces of a row; where the piece lives comes out of query
                // ql's table.  Pieces outside the map are zero-filled (src-size 0), pieces no tap touches are skipped.
                constexpr int PER_Q = NRB * NV;
                constexpr int PASSES = (PER_Q + 31) / 32;
#pragma unroll
                for (int ps = 0; ps < PASSES; ++ps) {
                    const int slot = ps * 32 + lane;
                    const int row = slot / NV;  // constant divisor
                    const int v = slot - row * NV;
                    const bool slot_ok = slot < PER_Q;
                    float* dst0 = stage + slot * 4;
                    {
                        // one opaque 64-bit base: keeps the compiler from carrying (base, 64-bit index) pairs through every copy
                        const float* vol_q = vol_g;
                        asm volatile("" : "+l"(vol_q));
                        // (lanes without a slot must still read inside the table)
                        const int* ty = tab + (slot_ok ? row : 0);
                        const int* tx = tab + NRB + (slot_ok ? v : 0);
                        // fully unrolled where the band loop is short; many bands x 32 copies would outgrow the instruction cache
                        // eight queries at a time: their 16 table entries are loaded first, so the shared-memory latency is paid
                        // once per batch and not once per copy
                        constexpr int UNR = (NB * PASSES <= 2) ? 4 : 1;
#pragma unroll UNR
                        for (int q0 = 0; q0 < 32; q0 += 8) {
                            int t_y[8], t_x[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                t_y[u] = ty[(q0 + u) * TW];
                                t_x[u] = tx[(q0 + u) * TW];
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const int t = t_y[u] | t_x[u];
                                const bool ok = t >= 0;
                                const float* src = vol_q + (ok ? (uint32_t)(t_y[u] + t_x[u]) : 0u);
                                if (slot_ok && !(t & 0x40000000)) cp_async16_zfill(dst0 + (q0 + u) * QS, src, ok ? 16u : 0u);
                            }
                        }
                    }
                }
                cp_async_wait_all();
                __syncwarp();

                // ---- lane = query: gather + separable blend ----
                if (live && warp_regular) {
                    // rows once: h[rr][a] = lerp_x(row yb + rr) is shared by the two taps that touch row rr; tap bb blends
                    // rows bb and bb + 1.  Same expressions as the general form below, so the results are bit-identical;
                    // D + 1 shared loads per row (neighbouring taps share a column) instead of 4 per (tap, a), no per-sample
                    // address arithmetic.
                    const float* win = stage + lane * QS + c0;
                    // one 64-bit pointer per lane walks the channels k = a * D + j of tap row j (D * HW floats apart): a 64-bit
                    // add per store.  (Opaque to the compiler, which otherwise either rebuilds the address from a uniform base
                    // and a 64-bit index for every store or precomputes all D * D addresses into registers.)
                    char* out_q = reinterpret_cast<char*>(out_b + ((size_t)(l * D * D) * (size_t)p.HW + (size_t)hw));
                    const unsigned long long ostep = (unsigned long long)p.HW * (unsigned)(4 * D);
                    float hp[D], hc[D], wx0[D], v[D + 1];
#pragma unroll
                    for (int k = 0; k <= D; ++k) v[k] = win[k];
#pragma unroll
                    for (int a = 0; a < D; ++a) {
                        wx0[a] = __fsub_rn(1.0f, xw[a]);
                        hp[a] = fmaf(v[a + 1], xw[a], v[a] * wx0[a]);
                    }
#pragma unroll
                    for (int bb = 0; bb < JB; ++bb) {
                        const int j = j0 + bb;
                        if (j < j1) {
                            const float wy1 = yw[j], wy0 = __fsub_rn(1.0f, wy1);
                            char* op = out_q + (unsigned long long)p.HW * (unsigned)(4 * j);
#pragma unroll
                            for (int k = 0; k <= D; ++k) v[k] = win[(bb + 1) * PITCH + k];
#pragma unroll
                            for (int a = 0; a < D; ++a) {
                                hc[a] = fmaf(v[a + 1], xw[a], v[a] * wx0[a]);
                                asm volatile("" : "+l"(op));
                                __stcs(reinterpret_cast<float*>(op), fmaf(hc[a], wy1, hp[a] * wy0));
                                op += ostep;
                                hp[a] = hc[a];
                            }
                        }
                    }
                } else if (live) {
                    const float* win = stage + lane * QS;
                    // one 64-bit pointer per lane walks the channels k = a * D + j of tap row j (D * HW floats apart): a 64-bit
                    // add per store.  (Opaque to the compiler, which otherwise either rebuilds the address from a uniform base
                    // and a 64-bit index for every store or precomputes all D * D addresses into registers.)
                    char* out_q = reinterpret_cast<char*>(out_b + ((size_t)(l * D * D) * (size_t)p.HW + (size_t)hw));
                    const unsigned long long ostep = (unsigned long long)p.HW * (unsigned)(4 * D);
#pragma unroll
                    for (int bb = 0; bb < JB; ++bb) {
                        const int j = j0 + bb;
                        if (j < j1) {
                            const float wy1 = yw[j], wy0 = __fsub_rn(1.0f, wy1);
                            const float* rowp = win + (yo[j] - yb) * PITCH - xs;
                            char* op = out_q + (unsigned long long)p.HW * (unsigned)(4 * j);
#pragma unroll
                            for (int a = 0; a < D; ++a) {
                                const float* t0 = rowp + xo[a];
                                const float wx1 = xw[a], wx0 = __fsub_rn(1.0f, wx1);
                                const float h0 = fmaf(t0[1], wx1, t0[0] * wx0);
                                const float h1 = fmaf(t0[PITCH + 1], wx1, t0[PITCH] * wx0);
                                asm volatile("" : "+l"(op));
                                __stcs(reinterpret_cast<float*>(op), fmaf(h1, wy1, h0 * wy0));
                                op += ostep;
                            }
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Generic path: runtime radius (up to 16) and/or levels whose width is not a multiple of 4.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) corr_lookup_generic_kernel(const LookupParams p) {
    extern __shared__ __align__(16) float smem[];
    const int warps_per_block = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int r = p.radius;
    const int D = 2 * r + 1;
    constexpr int TAB = 2 * LOOKUP_MAX_RADIUS + 1;

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
        const float cx = __fadd_rn((float)qw, fx);
        const float cy = __fadd_rn((float)qh, fy);
        float* out_q = p.out + (size_t)b * p.L * D * D * p.HW + hwc;

        for (int l = 0; l < p.L; ++l) {
            const int Hl = p.vh[l], Wl = p.vw[l];
            const float inv = 1.0f / (float)(1 << l);
            const float lx = __fmul_rn(cx, inv), ly = __fmul_rn(cy, inv);
            int xo[TAB], yo[TAB];
            float xw[TAB], yw[TAB];
            for (int a = 0; a < D; ++a) {
                axis_tap(__fadd_rn(lx, (float)(a - r)), Wl, xo[a], xw[a]);
                axis_tap(__fadd_rn(ly, (float)(a - r)), Hl, yo[a], yw[a]);
            }
            const int xmin = xo[0], xmax = xo[D - 1] + 1;
            const int ymin = yo[0], ymax = yo[D - 1] + 1;
            const bool vec = p.vec_ok[l] != 0;
            const int xs = vec ? (xmin & ~3) : xmin;
            __syncwarp();
            hdr[lane] = xs;
            hdr[32 + lane] = ymin;
            hdr[64 + lane] = xmax;
            hdr[96 + lane] = ymax;
            __syncwarp();

            const int cols = p.nv_max * 4;
            if (vec) {
                const int per_q = p.nr_max * p.nv_max;
                for (int i = lane; i < 32 * per_q; i += 32) {
                    const int ql = i / per_q;
                    const int rem = i - ql * per_q;
                    const int row = rem / p.nv_max;
                    const int v = rem - row * p.nv_max;
                    const int y = hdr[32 + ql] + row;
                    const int x = hdr[ql] + 4 * v;
                    if (y > hdr[96 + ql] || x > hdr[64 + ql]) continue;
                    float* dst = stage + ql * p.qstride + (row * p.nv_max + v) * 4;
                    const int qhw = (g - b * p.groups_per_b) * 32 + ql;
                    if (y >= 0 && y < Hl && x >= 0 && x < Wl && qhw < p.HW) {
                        cp_async16(dst, p.vol[l] + (((size_t)b * p.HW + qhw) * Hl + y) * Wl + x);
                    } else {
                        *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            } else {
                const int per_q = p.nr_max * cols;
                for (int i = lane; i < 32 * per_q; i += 32) {
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
                        cp_async4(dst, p.vol[l] + (((size_t)b * p.HW + qhw) * Hl + y) * Wl + x);
                    } else {
                        *dst = 0.f;
                    }
                }
            }
            cp_async_wait_all();
            __syncwarp();

            const float* win = stage + lane * p.qstride;
            float* out_l = out_q + (size_t)l * D * D * p.HW;
            if (live) {
                for (int a = 0; a < D; ++a) {
                    const int cxo = xo[a] - xs;
                    const float wx1 = xw[a], wx0 = __fsub_rn(1.0f, wx1);
                    for (int bb = 0; bb < D; ++bb) {
                        const float* t0 = win + (yo[bb] - ymin) * cols + cxo;
                        const float wy1 = yw[bb], wy0 = __fsub_rn(1.0f, wy1);
                        const float h0 = fmaf(t0[1], wx1, t0[0] * wx0);
                        const float h1 = fmaf(t0[cols + 1], wx1, t0[cols] * wx0);
                        __stcs(out_l + (size_t)(a * D + bb) * p.HW, fmaf(h1, wy1, h0 * wy0));
                    }
                }
            }
        }
    }
}

// ---- generic bilinear sampler (F.grid_sample bilinear / zeros) -----------------------------
// A thread takes one output pixel and BS_CPT channels (blockIdx.y picks the channel chunk): the tap arithmetic is
// done once per BS_CPT channels, the 4 * BS_CPT tap loads are independent, and for every channel the reads of
// neighbouring pixels are coalesced when the sampling field is smooth (feature warping by a flow field) and
// the stores always are.
constexpr int BS_CPT = 8;
__global__ void __launch_bounds__(128) bilinear_sample_kernel(const float* __restrict__ feat, const float* __restrict__ grid,
                                                              int N, int C, int Hf, int Wf, int Ho, int Wo, int grid_chw,
                                                              int align_corners, int scale, float* __restrict__ out) {
    const long long total = (long long)N * Ho * Wo;
    const int c0 = blockIdx.y * BS_CPT;
    grid_dependency_wait();  // programmatic dependent launch
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / ((long long)Ho * Wo));
        const int pix = (int)(i - (long long)n * Ho * Wo);
        float gx, gy;
        if (grid_chw) {
            gx = __ldg(grid + ((size_t)n * 2 + 0) * Ho * Wo + pix);
            gy = __ldg(grid + ((size_t)n * 2 + 1) * Ho * Wo + pix);
        } else {
            const float2 g2 = __ldg(reinterpret_cast<const float2*>(grid) + (size_t)n * Ho * Wo + pix);
            gx = g2.x;
            gy = g2.y;
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
        // a NaN / infinite coordinate makes every tap weight NaN in F.grid_sample: the sample is NaN, not padding
        const bool poisoned = !(fabsf(ix) <= 3.402823466e38f) || !(fabsf(iy) <= 3.402823466e38f);
        const int x0 = finite ? (int)fx0 : -2, y0 = finite ? (int)fy0 : -2;
        const bool okx0 = x0 >= 0 && x0 < Wf, okx1 = x0 + 1 >= 0 && x0 + 1 < Wf;
        const bool oky0 = y0 >= 0 && y0 < Hf, oky1 = y0 + 1 >= 0 && y0 + 1 < Hf;
        const float w00 = (okx0 && oky0) ? wx0 * wy0 : 0.f, w01 = (okx1 && oky0) ? wx1 * wy0 : 0.f;
        const float w10 = (okx0 && oky1) ? wx0 * wy1 : 0.f, w11 = (okx1 && oky1) ? wx1 * wy1 : 0.f;
        const int xa = min(max(x0, 0), Wf - 1), xb = min(max(x0 + 1, 0), Wf - 1);
        const int ya = min(max(y0, 0), Hf - 1), yb = min(max(y0 + 1, 0), Hf - 1);
        const size_t plane = (size_t)Hf * Wf, oplane = (size_t)Ho * Wo;
        const float* f = feat + ((size_t)n * C + c0) * plane;
        float* o = out + ((size_t)n * C + c0) * oplane + pix;
        const int o00 = ya * Wf + xa, o01 = ya * Wf + xb, o10 = yb * Wf + xa, o11 = yb * Wf + xb;
        float v00[BS_CPT], v01[BS_CPT], v10[BS_CPT], v11[BS_CPT];
#pragma unroll
        for (int c = 0; c < BS_CPT; ++c) {
            const float* fc = f + (size_t)(c0 + c < C ? c : 0) * plane;   // clamp: the tail chunk re-reads its first channel
            v00[c] = __ldg(fc + o00);
            v01[c] = __ldg(fc + o01);
            v10[c] = __ldg(fc + o10);
            v11[c] = __ldg(fc + o11);
        }
#pragma unroll
        for (int c = 0; c < BS_CPT; ++c) {
            float acc = v00[c] * w00;
            acc = fmaf(v01[c], w01, acc);
            acc = fmaf(v10[c], w10, acc);
            acc = fmaf(v11[c], w11, acc);
            if (c0 + c < C) __stcs(o + (size_t)c * oplane, poisoned ? __int_as_float(0x7fc00000) : acc);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// F.grid_sample for the argument combinations PicoPose never uses (utils/corr_lookup.py:29-65 passes mode / padding_mode
// through): nearest and bicubic interpolation, border and reflection padding.  A compatibility path, one thread per
// output pixel, written after ATen's GridSampler rules:
//   unnormalise -> (border: clip | reflection: reflect, then clip) -> taps; bicubic applies the padding rule per tap and
//   uses the A = -0.75 convolution coefficients; nearest rounds half to even.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gs_reflect(float in, int twice_low, int twice_high) {
    if (twice_low == twice_high) return 0.0f;
    const float mn = (float)twice_low * 0.5f, span = (float)(twice_high - twice_low) * 0.5f;
    in = fabsf(in - mn);
    const float extra = fmodf(in, span);
    const int flips = (int)floorf(in / span);
    return (flips & 1) == 0 ? extra + mn : span - extra + mn;
}
__device__ __forceinline__ float gs_pad(float c, int size, int padding, int align_corners) {
    if (padding == 1) return fminf((float)(size - 1), fmaxf(c, 0.0f));
    if (padding == 2) {
        c = align_corners ? gs_reflect(c, 0, 2 * (size - 1)) : gs_reflect(c, -1, 2 * size - 1);
        return fminf((float)(size - 1), fmaxf(c, 0.0f));
    }
    return c;
}
__device__ __forceinline__ float gs_fetch(const float* fc, float x, float y, int Wf, int Hf) {
    // a non-finite or far-away coordinate is out of bounds
    if (!(x > -1.0e9f && x < 1.0e9f && y > -1.0e9f && y < 1.0e9f)) return 0.0f;
    const int xi = (int)x, yi = (int)y;
    return (xi >= 0 && xi < Wf && yi >= 0 && yi < Hf) ? __ldg(fc + (size_t)yi * Wf + xi) : 0.0f;
}
__device__ __forceinline__ void gs_cubic(float t, float (&w)[4]) {
    const float A = -0.75f;
    float x = t + 1.0f;
    w[0] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
    x = t;
    w[1] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
    x = 1.0f - t;
    w[2] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
    x = 2.0f - t;
    w[3] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
}

__global__ void __launch_bounds__(128) grid_sample_general_kernel(const float* __restrict__ feat, const float* __restrict__ grid,
                                                                  int N, int C, int Hf, int Wf, int Ho, int Wo, int grid_chw,
                                                                  int align_corners, int scale, int mode, int padding,
                                                                  float* __restrict__ out) {
    const long long total = (long long)N * Ho * Wo;
    const size_t plane = (size_t)Hf * Wf, oplane = (size_t)Ho * Wo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(i / ((long long)Ho * Wo));
        const int pix = (int)(i - (long long)n * Ho * Wo);
        float gx = grid_chw ? __ldg(grid + ((size_t)n * 2 + 0) * oplane + pix) : __ldg(grid + ((size_t)n * oplane + pix) * 2);
        float gy = grid_chw ? __ldg(grid + ((size_t)n * 2 + 1) * oplane + pix) : __ldg(grid + ((size_t)n * oplane + pix) * 2 + 1);
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
        const float* f = feat + (size_t)n * C * plane;
        float* o = out + (size_t)n * C * oplane + pix;
        if (mode == 1) {  // nearest
            const float x = nearbyintf(gs_pad(ix, Wf, padding, align_corners)), y = nearbyintf(gs_pad(iy, Hf, padding, align_corners));
            for (int c = 0; c < C; ++c) o[(size_t)c * oplane] = gs_fetch(f + (size_t)c * plane, x, y, Wf, Hf);
        } else if (mode == 2) {  // bicubic: the padding rule applies to each of the 4 x 4 taps
            const float fx = floorf(ix), fy = floorf(iy);
            float wx[4], wy[4];
            gs_cubic(ix - fx, wx);
            gs_cubic(iy - fy, wy);
            float xs[4], ys[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                xs[k] = gs_pad(fx - 1.0f + (float)k, Wf, padding, align_corners);
                ys[k] = gs_pad(fy - 1.0f + (float)k, Hf, padding, align_corners);
            }
            for (int c = 0; c < C; ++c) {
                const float* fc = f + (size_t)c * plane;
                float acc = 0.0f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float row = 0.0f;
#pragma unroll
                    for (int k = 0; k < 4; ++k) row += gs_fetch(fc, xs[k], ys[j], Wf, Hf) * wx[k];
                    acc += row * wy[j];
                }
                o[(size_t)c * oplane] = acc;
            }
        } else {  // bilinear with border / reflection padding
            ix = gs_pad(ix, Wf, padding, align_corners);
            iy = gs_pad(iy, Hf, padding, align_corners);
            const float fx = floorf(ix), fy = floorf(iy);
            const float wx1 = ix - fx, wy1 = iy - fy, wx0 = (fx + 1.0f) - ix, wy0 = (fy + 1.0f) - iy;
            for (int c = 0; c < C; ++c) {
                const float* fc = f + (size_t)c * plane;
                float acc = gs_fetch(fc, fx, fy, Wf, Hf) * (wx0 * wy0);
                acc += gs_fetch(fc, fx + 1.0f, fy, Wf, Hf) * (wx1 * wy0);
                acc += gs_fetch(fc, fx, fy + 1.0f, Wf, Hf) * (wx0 * wy1);
                acc += gs_fetch(fc, fx + 1.0f, fy + 1.0f, Wf, Hf) * (wx1 * wy1);
                o[(size_t)c * oplane] = acc;
            }
        }
    }
}

static int grid_for(long long items, int wpb, size_t smem) {
    int blocks_per_sm = (int)((227 * 1024) / (smem + 1024));
    blocks_per_sm = blocks_per_sm < 1 ? 1 : (blocks_per_sm > 16 ? 16 : blocks_per_sm);
    const long long want = (items + wpb - 1) / wpb;
    const long long cap = (long long)sm_count() * blocks_per_sm * 4;  // grid-stride beyond 4 waves
    return (int)(want < cap ? want : cap);
}

template <int R, int JB, bool TILED>
static int launch_banded_l(const LookupParams& p, cudaStream_t st) {
    using Cfg = LookupCfg<R, JB>;
    const size_t per_warp = (size_t)Cfg::WARP_WORDS_TAB * sizeof(float);
    // blocks of up to 4 warps, sized so that at least two blocks share an SM
    int wpb = (int)((110 * 1024) / per_warp);
    wpb = wpb < 1 ? 1 : (wpb > 4 ? 4 : wpb);
    const size_t smem = per_warp * wpb;
    PP_CUDA(cudaFuncSetAttribute(corr_lookup_banded_kernel<R, JB, TILED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PP_CUDA(launch_dependent(corr_lookup_banded_kernel<R, JB, TILED>, dim3(grid_for((long long)p.total_groups * p.L, wpb, smem)),
                             dim3(wpb * 32), smem, st, p));
    PP_LAUNCHED();
    return PP_OK;
}
template <int R, int JB>
static int launch_banded(const LookupParams& p, cudaStream_t st) {
    return p.tiled ? launch_banded_l<R, JB, true>(p, st) : launch_banded_l<R, JB, false>(p, st);
}

}  // namespace pp

namespace pp {
static int corr_lookup_impl(const void* const* pyr_ptrs, const int* pyr_h, const int* pyr_w, int L,
                            const float* flow, int B, int H, int W, int radius, float* out, void* stream, bool tiled) {
    if (int rc = require_sm100()) return rc;
    if (B == 0) return PP_OK;  // empty batch: nothing to do (pointers of empty tensors may be null)
    PP_CHECK_ARG(pyr_ptrs && pyr_h && pyr_w && flow && out, "pp_corr_lookup: null pointer");
    PP_CHECK_ARG(L >= 1 && L <= LOOKUP_MAX_LEVELS, "pp_corr_lookup: 1 <= levels <= %d (got %d)", LOOKUP_MAX_LEVELS, L);
    PP_CHECK_ARG(radius >= 0 && radius <= LOOKUP_MAX_RADIUS, "pp_corr_lookup: 0 <= radius <= %d (got %d)",
                 LOOKUP_MAX_RADIUS, radius);
    PP_CHECK_ARG(B > 0 && H > 0 && W > 0, "pp_corr_lookup: bad flow shape (%d,2,%d,%d)", B, H, W);
    LookupParams p{};
    p.L = L;
    bool all_vec = true;
    for (int l = 0; l < L; ++l) {
        PP_CHECK_ARG(pyr_ptrs[l] && pyr_h[l] > 0 && pyr_w[l] > 0, "pp_corr_lookup: bad pyramid level %d", l);
        PP_CHECK_ARG((long long)pyr_h[l] * pyr_w[l] < (1LL << 31), "pp_corr_lookup: level %d too large", l);
        p.vol[l] = static_cast<const float*>(pyr_ptrs[l]);
        p.vh[l] = pyr_h[l];
        p.vw[l] = pyr_w[l];
        p.vec_ok[l] = (pyr_w[l] % 4 == 0) && ((reinterpret_cast<uintptr_t>(pyr_ptrs[l]) & 15) == 0);
        // the banded kernels keep element offsets within a group's 32 slices in 30 bits (bits 31 / 30 of a table entry are
        // flags): larger slices take the generic kernel
        all_vec = all_vec && p.vec_ok[l] && (long long)pyr_h[l] * pyr_w[l] < (1LL << 25);
        if (tiled)
            PP_CHECK_ARG(pyr_w[l] % 8 == 0 && pyr_h[l] % 4 == 0 && p.vec_ok[l],
                         "pp_corr_lookup_tiled: level %d is %dx%d; tiled slices need H %% 4 == 0, W %% 8 == 0, 16-byte alignment", l,
                         pyr_h[l], pyr_w[l]);
    }
    p.tiled = tiled ? 1 : 0;
    PP_CHECK_ARG(!tiled || (radius >= 1 && radius <= 8), "pp_corr_lookup_tiled: radius 1..8 (got %d)", radius);
    const int D = 2 * radius + 1;
    PP_CHECK_ARG((long long)L * D * D * H * W < (1LL << 31), "pp_corr_lookup: output per detection exceeds 2^31 elements");
    p.flow = flow;
    p.out = out;
    p.B = B;
    p.H = H;
    p.W = W;
    p.HW = H * W;
    p.radius = radius;
    p.groups_per_b = (p.HW + 31) / 32;
    p.total_groups = B * p.groups_per_b;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (tiled) {
        // PICOPOSE_LOOKUP_KERNEL=tiles: the whole-tile fetch kernel (corr_lookup_tma.cu).  Measured slower than the banded
        // kernel on the same tiled volume (0.99 vs 0.28 ms at r = 4: two warps per SM cannot fill the issue slots), so it
        // is opt-in; see DESIGN.md section 3.4 and profiles/r2j_lookup_tiles.*.
        const char* e = getenv("PICOPOSE_LOOKUP_KERNEL");
        if (e && strcmp(e, "tiles") == 0) {
            bool handled = false;
            if (int rc = launch_lookup_tma(p, st, &handled)) return rc;
            if (handled) return PP_OK;
        }
    }
    if (all_vec) {
        // Band height per radius and layout, from the sweep in profiles/r2z_lookup_band_sweep.txt: taller bands amortise the
        // x-interpolated row two taps share and cut fewer tiles twice, until the staged footprint costs occupancy.
        const bool t = p.tiled != 0;
        switch (radius) {
            case 1: return launch_banded<1, 3>(p, st);
            case 2: return launch_banded<2, 5>(p, st);
            case 3: return launch_banded<3, 7>(p, st);
            case 4: return t ? launch_banded<4, 6>(p, st) : launch_banded<4, 5>(p, st);
            case 5: return launch_banded<5, 6>(p, st);
            case 6: return t ? launch_banded<6, 7>(p, st) : launch_banded<6, 5>(p, st);
            case 7: return t ? launch_banded<7, 4>(p, st) : launch_banded<7, 5>(p, st);
            case 8: return launch_banded<8, 6>(p, st);
            default: break;
        }
    }
    PP_CHECK_ARG(!tiled, "pp_corr_lookup_tiled: slices of 2^25 elements or more are not supported in the tiled layout");
    // generic kernel: whole window staged at once
    p.nr_max = D + 2;
    p.nv_max = (D + 2 + 3 + 3) / 4;
    p.qstride = ((p.nr_max * p.nv_max) | 1) * 4;
    const size_t per_warp = (size_t)(32 * p.qstride + 4 * 32) * sizeof(float);
    int wpb = (int)((110 * 1024) / per_warp);
    wpb = wpb < 1 ? 1 : (wpb > 4 ? 4 : wpb);
    const size_t smem = per_warp * wpb;
    PP_CHECK_ARG(smem <= 227 * 1024, "pp_corr_lookup: radius %d needs %zu B of shared memory", radius, smem);
    PP_CUDA(cudaFuncSetAttribute(corr_lookup_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    corr_lookup_generic_kernel<<<grid_for(p.total_groups, wpb, smem), wpb * 32, smem, st>>>(p);
    PP_LAUNCHED();
    return PP_OK;
}
}  // namespace pp

extern "C" int pp_corr_lookup(const void* const* pyr_ptrs, const int* pyr_h, const int* pyr_w, int L,
                              const float* flow, int B, int H, int W, int radius, float* out, void* stream) {
    return pp::corr_lookup_impl(pyr_ptrs, pyr_h, pyr_w, L, flow, B, H, W, radius, out, stream, false);
}

extern "C" int pp_corr_lookup_tiled(const void* const* pyr_ptrs, const int* pyr_h, const int* pyr_w, int L,
                                    const float* flow, int B, int H, int W, int radius, float* out, void* stream) {
    return pp::corr_lookup_impl(pyr_ptrs, pyr_h, pyr_w, L, flow, B, H, W, radius, out, stream, true);
}

extern "C" int pp_bilinear_sample(const float* feat, const float* grid, int N, int C, int Hf, int Wf, int Ho,
                                  int Wo, int grid_chw, int align_corners, int scale, float* out, void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (N == 0) return PP_OK;
    PP_CHECK_ARG(feat && grid && out, "pp_bilinear_sample: null pointer");
    PP_CHECK_ARG(N > 0 && C > 0 && Hf > 0 && Wf > 0 && Ho > 0 && Wo > 0, "pp_bilinear_sample: bad shape");
    const long long total = (long long)N * Ho * Wo;
    PP_CHECK_ARG((C + BS_CPT - 1) / BS_CPT <= 65535, "pp_bilinear_sample: too many channels");
    PP_CHECK_ARG(grid_chw || (reinterpret_cast<uintptr_t>(grid) & 7) == 0, "pp_bilinear_sample: grid must be 8-byte aligned");
    int grid_dim = (int)((total + 127) / 128);
    const int cap = sm_count() * 64;
    if (grid_dim > cap) grid_dim = cap;
    PP_CUDA(launch_dependent(bilinear_sample_kernel, dim3(grid_dim, (C + BS_CPT - 1) / BS_CPT), dim3(128), 0,
                             static_cast<cudaStream_t>(stream), feat, grid, N, C, Hf, Wf, Ho, Wo, grid_chw, align_corners, scale, out));
    PP_LAUNCHED();
    return PP_OK;
}

extern "C" int pp_grid_sample(const float* feat, const float* grid, int N, int C, int Hf, int Wf, int Ho, int Wo,
                              int grid_chw, int align_corners, int scale, int mode, int padding_mode, float* out,
                              void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (mode == PP_SAMPLE_BILINEAR && padding_mode == PP_PAD_ZEROS)
        return pp_bilinear_sample(feat, grid, N, C, Hf, Wf, Ho, Wo, grid_chw, align_corners, scale, out, stream);
    if (N == 0) return PP_OK;
    PP_CHECK_ARG(feat && grid && out, "pp_grid_sample: null pointer");
    PP_CHECK_ARG(N > 0 && C > 0 && Hf > 0 && Wf > 0 && Ho > 0 && Wo > 0, "pp_grid_sample: bad shape");
    PP_CHECK_ARG(mode >= 0 && mode <= 2 && padding_mode >= 0 && padding_mode <= 2, "pp_grid_sample: unknown mode %d / padding %d",
                 mode, padding_mode);
    const long long total = (long long)N * Ho * Wo;
    int grid_dim = (int)((total + 127) / 128);
    const int cap = sm_count() * 64;
    if (grid_dim > cap) grid_dim = cap;
    grid_sample_general_kernel<<<grid_dim, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        feat, grid, N, C, Hf, Wf, Ho, Wo, grid_chw, align_corners, scale, mode, padding_mode, out);
    PP_LAUNCHED();
    return PP_OK;
}
