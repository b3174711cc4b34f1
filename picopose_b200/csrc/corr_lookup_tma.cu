// Stage-3 correlation-window lookup on TILED volumes, fetched tile by tile (whole 128-byte lines, cp.async).
//
// Same result as corr_lookup.cu::corr_lookup_banded_kernel (bit for bit: same tap arithmetic, same blend), other data
// movement.  In the tiled layout every 4 x 8 patch of a slice is one 128-byte line, and DRAM moves whole lines, so the
// cost of a window is the number of tiles it crosses.  The banded kernel asks for its footprint in 16-byte pieces of
// ROWS -- on tiled volumes four scattered L2 requests per line, ~20 instructions of staging loop per piece row -- and is
// latency / issue bound there (ncu: DRAM at 60 % of peak although the traffic fell by a third).  Here every lane
// (= query) lists the tiles its window crosses, the warp then fetches the packed list cooperatively: 8 lanes per tile,
// one full-line request each, four tiles per instruction, tiles outside the map zero-filled (src-size 0); the warp's
// next item is in flight (cp.async groups) while it samples the current one (two buffers per warp).
// First version of this file used one cp.async.bulk per tile row and lane: correct, 2.7x SLOWER than the banded kernel
// (0.75 ms at r = 4) -- the copy engine takes ~60 cycles per small bulk request per SM, see profiles/r2i_*.
//
//   item   = (32 consecutive queries of one detection, pyramid level, y band); one warp (= one block) per item at a time
//   band   = JB consecutive y taps (all of them for r <= 4, two bands for r >= 5 so that the buffers stay small)
//   buffer = the lanes' tile boxes packed back to back (prefix sum of bx * by) + the list of their sources; capacity 75 % of the worst case --
//            a lane that does not fit (or whose window was stretched beyond the expected footprint by the float round
//            trip) samples straight from global memory instead, so every flow field is handled
#include "corr_lookup.cuh"
#include "pp_ptx.cuh"

namespace pp {
namespace {


template <int R>
struct TCfg {
    static constexpr int D = 2 * R + 1;
    static constexpr int NB = D > 9 ? 2 : 1;              // y bands
    static constexpr int JB = (D + NB - 1) / NB;          // y taps per band
    static constexpr int BXM = (D + 8) / 8 + 1;           // most tiles a (D+2)-wide footprint crosses along x
    static constexpr int BYM = (JB + 4) / 4 + 1;          // most tile rows a (JB+2)-high band crosses
    static constexpr int CAPQ = (BXM * BYM * 3 + 3) / 4;  // tiles budgeted per query: 75 % of the worst case
    static constexpr int CAP = 32 * CAPQ;                 // tiles per buffer
    static constexpr int BUF_BYTES = CAP * 128;
};

// 16-byte copy global -> shared that reads only `src_bytes` (0 or 16) and zero-fills the rest
__device__ __forceinline__ void cp16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct ItemGeom {
    int b, hw0, hw, l, band;
    bool live;
    float lx, ly;  // window centre at level l
};

__device__ __forceinline__ ItemGeom item_geom(const LookupParams& p, int item, int NB, int lane) {
    ItemGeom it;
    it.band = item % NB;
    const int gl = item / NB;
    const int g = gl / p.L;
    it.l = gl - g * p.L;
    it.b = g / p.groups_per_b;
    it.hw0 = (g - it.b * p.groups_per_b) * 32;
    it.hw = it.hw0 + lane;
    it.live = it.hw < p.HW;
    const int hwc = it.live ? it.hw : p.HW - 1;
    const int qh = hwc / p.W, qw = hwc - qh * p.W;
    const float cx = __fadd_rn((float)qw, __ldg(p.flow + ((size_t)it.b * 2 + 0) * p.HW + hwc));  // utils/corr_lookup.py:113
    const float cy = __fadd_rn((float)qh, __ldg(p.flow + ((size_t)it.b * 2 + 1) * p.HW + hwc));
    const float inv = 1.0f / (float)(1 << it.l);  // exact: centroid / 2**l, utils/corr_lookup.py:125
    it.lx = __fmul_rn(cx, inv);
    it.ly = __fmul_rn(cy, inv);
    return it;
}

template <int R>
__global__ void __launch_bounds__(32) corr_lookup_tma_kernel(const LookupParams p, int total_items) {
    using Cfg = TCfg<R>;
    constexpr int D = Cfg::D, NB = Cfg::NB, JB = Cfg::JB;
    extern __shared__ __align__(128) float smem[];  // [2][CAP tiles][32 floats] | tile sources [2][CAP]
    __shared__ int4 s_meta[2][32];                  // per buffer and lane: {base tile | spilled << 30, tx0, ty0, bx | by << 8}
    const int lane = threadIdx.x;
    const uint32_t buf_u32 = ptx::smem_u32(smem);
    const float** s_list = reinterpret_cast<const float**>(smem + 2 * (Cfg::BUF_BYTES / 4));

    // ---- issue: every lane lists the tiles its window (band) crosses, the warp fetches the packed list ----
    auto issue = [&](int item, int bsel) {
        const ItemGeom it = item_geom(p, item, NB, lane);
        const int Hl = p.vh[it.l], Wl = p.vw[it.l];
        const int j0 = it.band * JB, j1 = (j0 + JB < D) ? j0 + JB : D;
        int xf, xl, yf, yl;
        float w;
        axis_tap(__fadd_rn(it.lx, (float)(-R)), Wl, xf, w);
        axis_tap(__fadd_rn(it.lx, (float)R), Wl, xl, w);
        axis_tap(__fadd_rn(it.ly, (float)(j0 - R)), Hl, yf, w);
        axis_tap(__fadd_rn(it.ly, (float)(j1 - 1 - R)), Hl, yl, w);
        const int tx0 = xf >> 3, ty0 = yf >> 2;  // arithmetic shifts: floor for negative indices
        const int bx = ((xl + 1) >> 3) - tx0 + 1, by = ((yl + 1) >> 2) - ty0 + 1;
        const bool over = bx < 1 || by < 1 || bx > Cfg::BXM || by > Cfg::BYM;  // stretched window: sampled from global memory
        const int n = (it.live && !over) ? bx * by : 0;
        int end = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, end, o);
            if (lane >= o) end += up;
        }
        const bool fits = end <= Cfg::CAP;
        const int base = end - n;
        int total = fits ? end : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total = max(total, __shfl_xor_sync(0xffffffffu, total, o));
        const bool spilled = it.live && (over || !fits);
        s_meta[bsel][lane] = make_int4(base | (spilled ? (1 << 30) : 0), tx0, ty0, bx | (by << 8));
        const float** list = s_list + bsel * Cfg::CAP;
        if (n > 0 && fits) {
            const int tpr = Wl >> 3, tpc = Hl >> 2;
            const float* slice = p.vol[it.l] + ((size_t)it.b * p.HW + it.hw) * ((size_t)Hl * Wl);
            for (int r = 0; r < by; ++r) {
                const int ty = ty0 + r;
                for (int c = 0; c < bx; ++c) {
                    const int tx = tx0 + c;
                    const bool inb = (unsigned)ty < (unsigned)tpc && (unsigned)tx < (unsigned)tpr;
                    list[base + r * bx + c] = inb ? slice + (size_t)(ty * tpr + tx) * 32 : nullptr;  // null: a tile of zeros
                }
            }
        }
        __syncwarp();
        // 8 lanes per tile (one 128-byte line), 4 tiles per instruction
        const uint32_t dst0 = buf_u32 + (uint32_t)bsel * Cfg::BUF_BYTES + (uint32_t)(lane & 7) * 16u;
        for (int t = lane >> 3; t < total; t += 4) {
            const float* src = list[t];
            cp16_zfill(dst0 + (uint32_t)t * 128u, src ? src + (lane & 7) * 4 : p.flow, src ? 16u : 0u);
        }
        cp_commit();
    };

    // ---- sample: lane = query, 4 taps per window sample out of the lane's tile box (or global memory when spilled) ----
    auto sample = [&](int item, int bsel) {
        const ItemGeom it = item_geom(p, item, NB, lane);
        const int Hl = p.vh[it.l], Wl = p.vw[it.l];
        const int j0 = it.band * JB, j1 = (j0 + JB < D) ? j0 + JB : D;
        if (!it.live) return;
        const int4 meta = s_meta[bsel][lane];
        const bool spilled = (meta.x >> 30) & 1;
        const int tx0 = meta.y, ty0 = meta.z, bx = meta.w & 0xFF;
        int ox0[D], ox1[D];
        float xw[D];
        int xo[D];
#pragma unroll
        for (int a = 0; a < D; ++a) {
            axis_tap(__fadd_rn(it.lx, (float)(a - R)), Wl, xo[a], xw[a]);
            ox0[a] = (((xo[a] >> 3) - tx0) << 5) + (xo[a] & 7);
            ox1[a] = ((((xo[a] + 1) >> 3) - tx0) << 5) + ((xo[a] + 1) & 7);
        }
        float* out_b = p.out + (size_t)it.b * p.L * D * D * p.HW;
        const uint32_t obase = (uint32_t)(it.l * D * D) * (uint32_t)p.HW + (uint32_t)it.hw;
        const float* win = smem + (size_t)bsel * (Cfg::BUF_BYTES / 4) + (size_t)(meta.x & 0xFFFFFF) * 32;
        const float* slice = p.vol[it.l] + ((size_t)it.b * p.HW + it.hw) * ((size_t)Hl * Wl);
        const int tpr = Wl >> 3;
        for (int j = j0; j < j1; ++j) {
            int yo;
            float yw;
            axis_tap(__fadd_rn(it.ly, (float)(j - R)), Hl, yo, yw);
            const float wy1 = yw, wy0 = __fsub_rn(1.0f, wy1);
            if (!spilled) {
                const int oy0 = (((yo >> 2) - ty0) * bx << 5) + ((yo & 3) << 3);
                const int oy1 = ((((yo + 1) >> 2) - ty0) * bx << 5) + (((yo + 1) & 3) << 3);
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    const float wx1 = xw[a], wx0 = __fsub_rn(1.0f, wx1);
                    const float h0 = fmaf(win[oy0 + ox1[a]], wx1, win[oy0 + ox0[a]] * wx0);
                    const float h1 = fmaf(win[oy1 + ox1[a]], wx1, win[oy1 + ox0[a]] * wx0);
                    __stcs(out_b + (obase + (uint32_t)(a * D + j) * (uint32_t)p.HW), fmaf(h1, wy1, h0 * wy0));
                }
            } else {
                auto at = [&](int y, int x) {
                    const bool inb = (unsigned)y < (unsigned)Hl && (unsigned)x < (unsigned)Wl;
                    return inb ? __ldg(slice + ((((y >> 2) * tpr + (x >> 3)) << 5) + ((y & 3) << 3) + (x & 7))) : 0.f;
                };
#pragma unroll
                for (int a = 0; a < D; ++a) {
                    const float wx1 = xw[a], wx0 = __fsub_rn(1.0f, wx1);
                    const float h0 = fmaf(at(yo, xo[a] + 1), wx1, at(yo, xo[a]) * wx0);
                    const float h1 = fmaf(at(yo + 1, xo[a] + 1), wx1, at(yo + 1, xo[a]) * wx0);
                    __stcs(out_b + (obase + (uint32_t)(a * D + j) * (uint32_t)p.HW), fmaf(h1, wy1, h0 * wy0));
                }
            }
        }
    };

    int item = blockIdx.x;
    int bsel = 0;
    if (item < total_items) issue(item, 0);
    while (item < total_items) {
        const int next = item + gridDim.x;
        if (next < total_items) {
            issue(next, bsel ^ 1);
            cp_wait_group<1>();  // everything but the group just committed has landed (this lane's copies) ...
        } else {
            cp_wait_group<0>();
        }
        __syncwarp();            // ... and every other lane's
        sample(item, bsel);
        __syncwarp();            // every lane is done reading this buffer before the item after next is copied into it
        item = next;
        bsel ^= 1;
    }
}

template <int R>
int launch(const LookupParams& p, cudaStream_t st) {
    using Cfg = TCfg<R>;
    const size_t smem = 2 * (size_t)Cfg::BUF_BYTES + 2 * (size_t)Cfg::CAP * sizeof(const float*);
    auto kern = corr_lookup_tma_kernel<R>;
    PP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long items = (long long)p.total_groups * p.L * Cfg::NB;
    int per_sm = (int)((227 * 1024) / (smem + 2048));
    per_sm = per_sm < 1 ? 1 : (per_sm > 16 ? 16 : per_sm);
    long long grid = (long long)sm_count() * per_sm;
    if (grid > items) grid = items;
    kern<<<(unsigned)grid, 32, smem, st>>>(p, (int)items);
    PP_LAUNCHED();
    return PP_OK;
}

}  // namespace

int launch_lookup_tma(const LookupParams& p, cudaStream_t st, bool* handled) {
    *handled = false;
    if (!p.tiled || p.radius < 1 || p.radius > 8) return PP_OK;
    if ((long long)p.total_groups * p.L * 2 >= (1LL << 31)) return PP_OK;
    for (int l = 0; l < p.L; ++l)
        if (p.vw[l] % 8 != 0 || p.vh[l] % 4 != 0 || (reinterpret_cast<uintptr_t>(p.vol[l]) & 127) != 0) return PP_OK;
    *handled = true;
    switch (p.radius) {
        case 1: return launch<1>(p, st);
        case 2: return launch<2>(p, st);
        case 3: return launch<3>(p, st);
        case 4: return launch<4>(p, st);
        case 5: return launch<5>(p, st);
        case 6: return launch<6>(p, st);
        case 7: return launch<7>(p, st);
        default: return launch<8>(p, st);
    }
}

}  // namespace pp
