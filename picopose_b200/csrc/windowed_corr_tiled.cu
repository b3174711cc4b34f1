// Fused stage-3 correlation, tiled variant: same result as windowed_corr.cu::windowed_corr_kernel (see that
// file for the maths and the reference lines), different data movement.
//
// The per-query kernel pulls every neighbour vector through L1 once per query: 36 KB per query and level at
// r = 2, C = 256, and a 4-line LDG costs the L1 wavefront queue ~8 cycles, which is what bounds it.  Here a
// block owns an 8 x 8 tile of queries (a warp a 2 x 2 sub-block).  Per level it asks TMA for ONE square region of
// the position-major key map, 32 channels at a time (4-D box {32, RS, RS, 1}): RS = 16 placed around all the
// tile's windows when they fit (smooth flow), else RS = 24 centred on the tile's mean window.  Coordinates may
// lie outside the map: TMA zero-fills, which is exactly the lookup's zero padding.  Chunks are double buffered
// behind mbarriers and carry the tile's own 32-channel slice of f1.  A query whose window lies inside the
// region takes all its dot products from shared memory (conflict-free 128-byte rows, no bounds checks); one
// that does not (an outlier flow, or a window stretched by the float round trip) is deferred to a list and
// served from global memory after the pipeline, one per warp, so any flow field works and a straggler cannot
// stall the block.  Accumulators stay in registers across the C/32 chunks: lane group g of a warp owns
// neighbours g, g+4, ... of each of the warp's 4 queries; at level end one 8-lane transpose-reduce finishes all
// 36 dot products at once.  Taps and region origins of every level are computed up front so that the chunk
// pipeline runs across level boundaries without draining.
#include "pp_common.cuh"
#include "pp_ptx.cuh"

#include <cudaTypedefs.h>

namespace pp {

// called from windowed_corr.cu
int launch_wcorr_tiled(int radius, const float* f1t, const void* const* f2t_levels, int L, const float* flow, int N, int C,
                       int H, int W, float* out, cudaStream_t st, bool* handled, const WConv* conv = nullptr);

namespace {

constexpr int WT_WARPS = 16;
constexpr int WT_THREADS = WT_WARPS * 32;
constexpr int WT_QPW = 4;                  // queries per warp
constexpr int WT_TQ = WT_WARPS * WT_QPW;   // 64 queries: 8 x 8
constexpr int WT_CH = 32;                  // channels per chunk: one 128-byte row per key
constexpr int WT_MAX_LEVELS = 4;
constexpr long long WT_TIMEOUT_CYCLES = 4000000000LL;

struct WTileMaps {
    CUtensorMap f1;                  // (N, H, W, C) position-major query features
    CUtensorMap f2s[WT_MAX_LEVELS];  // level l: (N, H>>l, W>>l, C), small region box
    CUtensorMap f2b[WT_MAX_LEVELS];  // same tensor, big region box
};

struct WTileParams {
    const float* f1t;
    const float* f2t[WT_MAX_LEVELS];
    int hl[WT_MAX_LEVELS], wl[WT_MAX_LEVELS];
    const float* flow;  // (N, 2, H, W)
    float* out;         // (N, L*D*D, H, W)
    int N, C, H, W, HW, L;
    float scale;
    int tiles_x, tiles_y;
    WConv conv;         // conv.weight != null: fused 1x1 convolution epilogue; `out` may then be null (lookup not stored)
};

// same arithmetic as corr_lookup.cu::axis_tap / windowed_corr.cu::wc_axis_tap
__device__ __forceinline__ void wt_axis_tap(float p, int size, int& i0, float& w1) {
    float den = (float)(size > 1 ? size - 1 : 1);
    float g = __fsub_rn(__fdiv_rn(__fmul_rn(p, 2.0f), den), 1.0f);
    float i = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
    float f = floorf(i);
    w1 = __fsub_rn(i, f);
    f = fminf(fmaxf(f, -2.0f), (float)size);
    i0 = (f == f) ? (int)f : -2;
    // a non-finite coordinate leaves w1 = NaN: every sample that uses this tap is NaN, as in F.grid_sample
}

__device__ __forceinline__ void tma_load_4d(const void* desc, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            dst),
        "l"(desc), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void wt_wait(uint32_t bar, uint32_t parity) {
    if (ptx::mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!ptx::mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > WT_TIMEOUT_CYCLES) __trap();  // a lost TMA transaction must not hang the GPU
    }
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float fma4(const float4& a, const float4& v, float acc) {
    return fmaf(a.x, v.x, fmaf(a.y, v.y, fmaf(a.z, v.z, fmaf(a.w, v.w, acc))));
}

// Sums v[i] over the 8 lanes of a lane group for all NV values at once (NV % 8 == 0): three exchange stages,
// each lane sends the half it does not keep.  Afterwards lane gl holds, in v[0 .. NV/8), the totals of indices
// (gl & 4 ? NV/2 : 0) + (gl & 2 ? NV/4 : 0) + (gl & 1 ? NV/8 : 0) + j.  7/8 NV shuffles instead of 3 NV.
template <int NV>
__device__ __forceinline__ void group8_transpose_reduce(float (&v)[NV], int gl) {
    static_assert(NV % 8 == 0, "NV must be a multiple of 8");
    const bool h4 = gl & 4, h2 = gl & 2, h1 = gl & 1;
#pragma unroll
    for (int i = 0; i < NV / 2; ++i) {
        const float send = h4 ? v[i] : v[i + NV / 2], keep = h4 ? v[i + NV / 2] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) {
        const float send = h2 ? v[i] : v[i + NV / 4], keep = h2 ? v[i + NV / 4] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
#pragma unroll
    for (int i = 0; i < NV / 8; ++i) {
        const float send = h1 ? v[i] : v[i + NV / 8], keep = h1 ? v[i + NV / 8] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
}

template <int R>
struct WT {
    static constexpr int D = 2 * R + 1, G = D + 2, DD = D * D, GG = G * G;
    static constexpr int W1 = D + 1;            // side of a regular window's neighbour grid
    static constexpr int TRIPS = W1 * W1 / 4;   // 4 neighbours per trip (W1 is even)
    static constexpr int PER = (W1 % 4 == 0) ? 1 : (W1 % 2 == 0 ? W1 / 2 : W1);   // trips after which (row advance, column) repeats
    static constexpr int ROWS_PER = PER * 4 / W1;
    static constexpr int RSS = 16;              // small region: fits the windows of a tile whose flow is smooth
    static constexpr int RSB = 24;              // big region: tile + window + slack for a scattered flow
    static constexpr int KEY_WORDS = RSB * RSB * WT_CH;
    static constexpr int F1_WORDS = WT_TQ * WT_CH;
    static constexpr uint32_t STAGE_BYTES = (KEY_WORDS + F1_WORDS) * 4;
    static size_t smem_bytes(int L) {
        size_t w = 2 * (size_t)(KEY_WORDS + F1_WORDS);          // two stages, 128-byte rows
        w += ((size_t)L * DD * (WT_TQ + 1) + 3) & ~(size_t)3;   // out tile
        w += (size_t)WT_TQ * GG;                                // per-query correlation grid
        w += (size_t)L * WT_TQ * 4 * D;                         // taps of every level
        w += 8 * WT_MAX_LEVELS + 4;                             // region statistics / origins, barriers
        w += (size_t)WT_MAX_LEVELS * WT_TQ + 4;                 // deferred (level, query) list
        return w * 4 + 128;                                     // + alignment slack
    }
};

// query qi of warp `warp`: the warps tile the 8 x 8 block with 2 x 2 sub-blocks
__device__ __forceinline__ int wt_query(int warp, int qi) { return ((warp >> 2) * 2 + (qi >> 1)) * 8 + (warp & 3) * 2 + (qi & 1); }

// One 32-channel chunk for a warp: every regular in-region query walks its own W1 x W1 grid, a neighbour costs one
// 16-byte shared load and four FMAs.  Shared addresses are per-lane bases (set at level start) plus compile-time offsets.
template <int R, int RS, int STAGE>
__device__ __forceinline__ void wt_chunk(const uint32_t (&base)[WT_QPW][WT<R>::PER], const bool (&fast)[WT_QPW],
                                         float (&acc)[WT_QPW][WT<R>::TRIPS], uint32_t f1_addr) {
    using T = WT<R>;
    constexpr uint32_t SB = STAGE * T::STAGE_BYTES;
#pragma unroll
    for (int qi = 0; qi < WT_QPW; ++qi) {
        if (fast[qi]) {  // warp-uniform
            const float4 a = lds128(f1_addr + SB + ((qi >> 1) * 8 + (qi & 1)) * WT_CH * 4);
#pragma unroll
            for (int t = 0; t < T::TRIPS; ++t) {
                const float4 v = lds128(base[qi][t % T::PER] + SB + (t / T::PER) * T::ROWS_PER * RS * WT_CH * 4);
                acc[qi][t] = fma4(a, v, acc[qi][t]);
            }
        }
    }
}

template <int R>
__global__ void __launch_bounds__(WT_THREADS, 1)
windowed_corr_tiled_kernel(const __grid_constant__ WTileMaps maps, const WTileParams p) {
    using T = WT<R>;
    constexpr int D = T::D, DD = T::DD, GG = T::GG, W1 = T::W1, TRIPS = T::TRIPS, PER = T::PER, TQ = WT_TQ, QPW = WT_QPW;
    constexpr int RSS = T::RSS, RSB = T::RSB;
    extern __shared__ __align__(128) float smem[];  // TMA destinations need 128-byte alignment (checked below)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = lane >> 3, gl = lane & 7;
    const int L = p.L, rows = L * DD;
    float* stage0 = smem;                                      // [2][keys RS x RS x 32 | f1 TQ x 32]
    float* out_tile = smem + 2 * (T::KEY_WORDS + T::F1_WORDS);  // [rows][TQ + 1]
    float* vals = out_tile + (((size_t)rows * (TQ + 1) + 3) & ~(size_t)3);   // [TQ][GG]
    int* s_xo = reinterpret_cast<int*>(vals + (size_t)TQ * GG);              // [L][TQ][D]
    int* s_yo = s_xo + L * TQ * D;
    float* s_xw = reinterpret_cast<float*>(s_yo + L * TQ * D);
    float* s_yw = s_xw + L * TQ * D;
    int* s_red = reinterpret_cast<int*>(s_yw + L * TQ * D);   // [MAX_LEVELS][8]: sum x, sum y, count, min x, max x, min y, max y
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_red + 8 * WT_MAX_LEVELS);
    int* s_nslow = reinterpret_cast<int*>(s_bar + 2);
    int* s_slow = s_nslow + 4;                                               // [MAX_LEVELS * TQ]
    const uint32_t bar0 = ptx::smem_u32(s_bar), bar1 = bar0 + 8;
    const uint32_t stage_u32 = ptx::smem_u32(stage0);

    int b = blockIdx.x;
    const int tx = b % p.tiles_x;
    b /= p.tiles_x;
    const int ty = b % p.tiles_y;
    const int n = b / p.tiles_y;

    if (tid == 0) {
        if (stage_u32 & 127u) __trap();
        ptx::mbar_init(bar0, 1);
        ptx::mbar_init(bar1, 1);
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&maps.f1);
    }
    if (tid < 8 * WT_MAX_LEVELS) {
        const int k = tid & 7;
        s_red[tid] = (k == 3 || k == 5) ? 0x3fffffff : (k == 4 || k == 6) ? -0x3fffffff : 0;
    }
    if (tid == 32) *s_nslow = 0;
    __syncthreads();
    // programmatic dependent launch: the barrier set-up above overlaps the tail of the layout pass, whose key maps (and the
    // caller's flow) are only read from here on
    grid_dependency_wait();

    // ---- taps and window extents of every level; per level the tile's mean window and the box around all windows ----
    int q_hw[QPW];
#pragma unroll
    for (int qi = 0; qi < QPW; ++qi) {
        const int ql = wt_query(warp, qi);
        const int qh = ty * 8 + (ql >> 3), qw = tx * 8 + (ql & 7);
        const bool ok = qh < p.H && qw < p.W;
        q_hw[qi] = ok ? qh * p.W + qw : -1;
        float cx = 0.f, cy = 0.f;
        if (ok) {
            cx = __fadd_rn((float)qw, __ldg(p.flow + ((size_t)n * 2 + 0) * p.HW + q_hw[qi]));
            cy = __fadd_rn((float)qh, __ldg(p.flow + ((size_t)n * 2 + 1) * p.HW + q_hw[qi]));
        }
        for (int l = 0; l < L; ++l) {
            const float inv = 1.0f / (float)(1 << l);
            const int tb = (l * TQ + ql) * D;
            if (lane < D) {
                int i0;
                float w1;
                wt_axis_tap(__fadd_rn(__fmul_rn(cx, inv), (float)(lane - R)), p.wl[l], i0, w1);
                s_xo[tb + lane] = i0;
                s_xw[tb + lane] = w1;
                wt_axis_tap(__fadd_rn(__fmul_rn(cy, inv), (float)(lane - R)), p.hl[l], i0, w1);
                s_yo[tb + lane] = i0;
                s_yw[tb + lane] = w1;
            }
            __syncwarp();
            if (lane == 0 && ok) {
                const int x0 = s_xo[tb], y0 = s_yo[tb];
                atomicAdd(&s_red[l * 8 + 0], x0 + s_xo[tb + D - 1] + 2);  // 2 * window centre
                atomicAdd(&s_red[l * 8 + 1], y0 + s_yo[tb + D - 1] + 2);
                atomicAdd(&s_red[l * 8 + 2], 1);
                atomicMin(&s_red[l * 8 + 3], x0);
                atomicMax(&s_red[l * 8 + 4], x0);
                atomicMin(&s_red[l * 8 + 5], y0);
                atomicMax(&s_red[l * 8 + 6], y0);
            }
        }
    }
    __syncthreads();
    if (tid < L) {
        // small region when one RSS x RSS box holds every window of the tile (centred on them), else the big one
        // around the mean window (outliers then go to the deferred list)
        int* r = s_red + tid * 8;
        const int bw = r[4] - r[3] + W1, bh = r[6] - r[5] + W1;
        int X0, Y0, small = 0;
        if (r[2] > 0 && bw <= RSS && bh <= RSS) {
            small = 1;
            X0 = r[3] - (RSS - bw) / 2;
            Y0 = r[5] - (RSS - bh) / 2;
        } else {
            const float cnt = (float)max(r[2], 1);
            X0 = (int)floorf((float)r[0] / (2.0f * cnt) - 0.5f * (float)RSB + 0.5f);
            Y0 = (int)floorf((float)r[1] / (2.0f * cnt) - 0.5f * (float)RSB + 0.5f);
        }
        r[0] = X0;
        r[1] = Y0;
        r[2] = small;
    }
    __syncthreads();

    const int nch = p.C / WT_CH;
    const int total = L * nch;
    auto issue = [&](int it) {  // thread 0: TMA of chunk `it` into stage it & 1
        const int l = it / nch, c = it - l * nch;
        const int small = s_red[l * 8 + 2];
        const int rs = small ? RSS : RSB;
        const uint32_t bar = (it & 1) ? bar1 : bar0;
        const uint32_t st = stage_u32 + (uint32_t)(it & 1) * T::STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(bar, (uint32_t)(rs * rs + TQ) * WT_CH * 4);
        tma_load_4d(small ? &maps.f2s[l] : &maps.f2b[l], bar, st, c * WT_CH, s_red[l * 8 + 0], s_red[l * 8 + 1], n);
        tma_load_4d(&maps.f1, bar, st + T::KEY_WORDS * 4, c * WT_CH, tx * 8, ty * 8, n);
    };
    if (tid == 0) {
        issue(0);
        if (total > 1) issue(1);
    }

    // D*D bilinear samples of query ql at level bl from its neighbour grid vq (row stride gs), as corr_lookup.cu does
    auto blend = [&](int bl, int ql, const float* vq, int x0, int y0, int gs) {
        const int tb = (bl * TQ + ql) * D;
        for (int k = lane; k < DD; k += 32) {
            const int ai = k / D, bi = k - ai * D;
            const float* t0 = vq + (s_yo[tb + bi] - y0) * gs + (s_xo[tb + ai] - x0);
            const float wx1 = s_xw[tb + ai], wx0 = __fsub_rn(1.0f, wx1);
            const float wy1 = s_yw[tb + bi], wy0 = __fsub_rn(1.0f, wy1);
            const float h0 = fmaf(t0[1], wx1, t0[0] * wx0);
            const float h1 = fmaf(t0[gs + 1], wx1, t0[gs] * wx0);
            out_tile[(size_t)(bl * DD + k) * (TQ + 1) + ql] = fmaf(h1, wy1, h0 * wy0);
        }
    };

    int xmin[QPW], ymin[QPW];
    bool fast[QPW];
    bool small = false;
    uint32_t base[QPW][PER];
    float acc[QPW][T::TRIPS];
    const uint32_t f1_addr = stage_u32 + T::KEY_WORDS * 4 + (uint32_t)wt_query(warp, 0) * WT_CH * 4 + gl * 16;
    int l = 0, c = 0;
    for (int it = 0; it < total; ++it) {
        if (c == 0) {  // level start: which of this warp's queries have a regular window inside the staged region
            small = s_red[l * 8 + 2] != 0;
            const int rs = small ? RSS : RSB;
            const int X0 = s_red[l * 8 + 0], Y0 = s_red[l * 8 + 1];
#pragma unroll
            for (int qi = 0; qi < QPW; ++qi) {
                const int tb = (l * TQ + wt_query(warp, qi)) * D;
                xmin[qi] = s_xo[tb];
                ymin[qi] = s_yo[tb];
                const int gx = s_xo[tb + D - 1] + 2 - xmin[qi], gy = s_yo[tb + D - 1] + 2 - ymin[qi];
                // regular (or border-clamped, hence smaller) window whose W1 x W1 grid lies inside the staged region
                fast[qi] = q_hw[qi] >= 0 && gx <= W1 && gy <= W1 && xmin[qi] >= X0 && xmin[qi] + W1 <= X0 + rs &&
                           ymin[qi] >= Y0 && ymin[qi] + W1 <= Y0 + rs;
#pragma unroll
                for (int j = 0; j < PER; ++j) {
                    const int id = 4 * j + grp, gyi = id / W1, gxi = id % W1;
                    base[qi][j] = stage_u32 + (uint32_t)(((ymin[qi] - Y0 + gyi) * rs + (xmin[qi] - X0 + gxi)) * WT_CH * 4 + gl * 16);
                }
            }
#pragma unroll
            for (int qi = 0; qi < QPW; ++qi)
#pragma unroll
                for (int t = 0; t < T::TRIPS; ++t) acc[qi][t] = 0.f;
        }
        wt_wait((it & 1) ? bar1 : bar0, (uint32_t)(it >> 1) & 1u);
        if (small) {
            if (it & 1) wt_chunk<R, RSS, 1>(base, fast, acc, f1_addr);
            else wt_chunk<R, RSS, 0>(base, fast, acc, f1_addr);
        } else {
            if (it & 1) wt_chunk<R, RSB, 1>(base, fast, acc, f1_addr);
            else wt_chunk<R, RSB, 0>(base, fast, acc, f1_addr);
        }
        __syncthreads();  // everyone is done with this stage
        if (tid == 0 && it + 2 < total) issue(it + 2);

        if (c == nch - 1) {  // level end: finish the dot products, blend the D*D samples of each query
            {
                // 4 x 9 partial sums, padded to 4 x 10: lane gl ends up with query gl >> 1, trips (gl & 1) * 5 + j
                constexpr int TP = (TRIPS + 1) & ~1;
                float v[QPW * TP];
#pragma unroll
                for (int qi = 0; qi < QPW; ++qi)
#pragma unroll
                    for (int t = 0; t < TP; ++t) v[qi * TP + t] = t < TRIPS ? acc[qi][t] : 0.f;
                group8_transpose_reduce<QPW * TP>(v, gl);
                float* vq = vals + (size_t)wt_query(warp, gl >> 1) * GG;
#pragma unroll
                for (int j = 0; j < TP / 2; ++j) {
                    const int t = (gl & 1) * (TP / 2) + j;
                    if (t < TRIPS) vq[t * 4 + grp] = v[j] * p.scale;
                }
            }
            __syncwarp();
#pragma unroll
            for (int qi = 0; qi < QPW; ++qi) {
                if (q_hw[qi] < 0) continue;  // warp-uniform
                const int ql = wt_query(warp, qi);
                if (!fast[qi]) {
                    // stretched window (float round trip) or an outlier flow outside the staged region: deferred, so
                    // that its global-memory latency does not hold up the block's pipeline
                    if (lane == 0) s_slow[atomicAdd(s_nslow, 1)] = (l << 16) | ql;
                    continue;
                }
                blend(l, ql, vals + (size_t)ql * GG, xmin[qi], ymin[qi], W1);
            }
            __syncwarp();
            c = 0;
            ++l;
        } else {
            ++c;
        }
    }
    __syncthreads();
    // deferred queries, one per warp at a time, the per-query kernel's way (straight from global memory)
    const int nslow = *s_nslow;
    for (int i = warp; i < nslow; i += WT_WARPS) {
        const int sl = s_slow[i] >> 16, ql = s_slow[i] & 0xFFFF;
        const int qh = ty * 8 + (ql >> 3), qw = tx * 8 + (ql & 7);
        const int Hl = p.hl[sl], Wl = p.wl[sl];
        const int tb = (sl * TQ + ql) * D;
        const int x0 = s_xo[tb], y0 = s_yo[tb];
        const int sgx = s_xo[tb + D - 1] + 2 - x0, sgy = s_yo[tb + D - 1] + 2 - y0;
        const float* f1q = p.f1t + ((size_t)n * p.HW + qh * p.W + qw) * p.C + gl * 4;
        const float* f2n = p.f2t[sl] + (size_t)n * Hl * Wl * p.C + gl * 4;
        float* vq = vals + (size_t)warp * GG;  // the fast path is done with `vals`: one scratch grid per warp
        const int npts = sgx * sgy;
        const float rgx = 1.0f / (float)sgx;
        for (int idx = 0; idx < npts; idx += 4) {
            const int id = idx + grp;
            const int gyi = (int)(((float)id + 0.5f) * rgx), gxi = id - gyi * sgx;  // id / sgx, exact for these small integers
            const int x = x0 + gxi, y = y0 + gyi;
            float s = 0.f;
            if (id < npts && (unsigned)x < (unsigned)Wl && (unsigned)y < (unsigned)Hl) {
                const float* v = f2n + (size_t)(y * Wl + x) * p.C;
                float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
                for (int cc = 0; cc < p.C; cc += 32) {
                    const float4 av = __ldg(reinterpret_cast<const float4*>(f1q + cc));
                    const float4 bv = __ldg(reinterpret_cast<const float4*>(v + cc));
                    s4.x = fmaf(av.x, bv.x, s4.x);
                    s4.y = fmaf(av.y, bv.y, s4.y);
                    s4.z = fmaf(av.z, bv.z, s4.z);
                    s4.w = fmaf(av.w, bv.w, s4.w);
                }
                s = (s4.x + s4.y) + (s4.z + s4.w);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (gl == 0 && id < npts) vq[id] = s * p.scale;
        }
        __syncwarp();
        blend(sl, ql, vq, x0, y0, sgx);
        __syncwarp();
    }
    __syncthreads();
    if (p.conv.weight) {
        // ---- fused 1x1 convolution: y[co, q] = act(bias[co] + sum_row W[co, row] * lookup[row, q]) for the tile's 64 queries.
        // The TMA stages are free now: the weight matrix is staged there transposed ([row][co]) so that a thread's 8
        // output channels are two 16-byte loads; a thread owns 8 channels x 4 queries (one row segment of the 8 x 8 tile).
        const int cout = p.conv.cout;
        float* w_s = stage0;
        for (int i = tid; i < cout * rows; i += WT_THREADS) {
            const int co = i / rows, row = i - co * rows;
            w_s[row * cout + co] = __ldg(p.conv.weight + i);
        }
        __syncthreads();
        const int qg = tid & 15, q0 = qg * 4;
        const int qh = ty * 8 + (q0 >> 3), qw = tx * 8 + (q0 & 7);
        for (int co0 = (tid >> 4) * 8; co0 < cout; co0 += (WT_THREADS >> 4) * 8) {
            float acc[8][4];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const float bv = p.conv.bias ? __ldg(p.conv.bias + co0 + a) : 0.f;
#pragma unroll
                for (int b2 = 0; b2 < 4; ++b2) acc[a][b2] = bv;
            }
            for (int row = 0; row < rows; ++row) {
                const float4 w0 = *reinterpret_cast<const float4*>(w_s + row * cout + co0);
                const float4 w1 = *reinterpret_cast<const float4*>(w_s + row * cout + co0 + 4);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                const float* xr = out_tile + (size_t)row * (TQ + 1) + q0;
                const float xv[4] = {xr[0], xr[1], xr[2], xr[3]};
#pragma unroll
                for (int a = 0; a < 8; ++a)
#pragma unroll
                    for (int b2 = 0; b2 < 4; ++b2) acc[a][b2] = fmaf(wv[a], xv[b2], acc[a][b2]);
            }
            if (qh < p.H) {
#pragma unroll
                for (int a = 0; a < 8; ++a) {
                    float* dst = p.conv.out + (((size_t)n * cout + co0 + a) * p.H + qh) * p.W + qw;
#pragma unroll
                    for (int b2 = 0; b2 < 4; ++b2) {
                        const float v = (p.conv.relu && acc[a][b2] < 0.f) ? 0.f : acc[a][b2];  // torch.relu keeps NaN; fmaxf would not
                        if (qw + b2 < p.W) __stcs(dst + b2, v);
                    }
                }
            }
        }
        if (!p.out) return;
    }
    // out[n, row, tile]: 8-float (32-byte sector) runs, 8 of them per row
    float* out_n = p.out + (size_t)n * rows * p.HW;
    for (int row = warp; row < rows; row += WT_WARPS) {
#pragma unroll
        for (int q0 = 0; q0 < TQ; q0 += 32) {
            const int ql = q0 + lane;
            const int qh = ty * 8 + (ql >> 3), qw = tx * 8 + (ql & 7);
            if (qh < p.H && qw < p.W) __stcs(out_n + (size_t)row * p.HW + qh * p.W + qw, out_tile[(size_t)row * (TQ + 1) + ql]);
        }
    }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

// (N, Hm, Wm, C) fp32, box {32, bw, bh, 1}, no swizzle, zero fill outside
int make_map(CUtensorMap* map, const void* base, int N, int Hm, int Wm, int C, int bw, int bh) {
    auto fn = encode_fn();
    if (!fn) return fail(PP_ERR_DEVICE, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wm, (cuuint64_t)Hm, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)Wm * C * 4, (cuuint64_t)Hm * Wm * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)WT_CH, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PP_ERR_LAUNCH, "cuTensorMapEncodeTiled (windowed correlation) failed with CUresult %d", (int)r);
    return PP_OK;
}

template <int R>
int launch(const WTileMaps& maps, const WTileParams& p, cudaStream_t st) {
    const size_t smem = WT<R>::smem_bytes(p.L);
    auto kern = windowed_corr_tiled_kernel<R>;
    PP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PP_CUDA(launch_dependent(kern, dim3(p.N * p.tiles_x * p.tiles_y), dim3(WT_THREADS), smem, st, maps, p));
    PP_LAUNCHED();
    return PP_OK;
}

}  // namespace

// Runs the tiled kernel when it covers the problem (*handled = true); otherwise leaves it to the per-query kernel.
int launch_wcorr_tiled(int radius, const float* f1t, const void* const* f2t_levels, int L, const float* flow, int N, int C,
                       int H, int W, float* out, cudaStream_t st, bool* handled, const WConv* conv) {
    *handled = false;
    if (radius < 1 || radius > 2 || L > WT_MAX_LEVELS || C % WT_CH != 0) return PP_OK;
    const size_t smem = radius == 1 ? WT<1>::smem_bytes(L) : WT<2>::smem_bytes(L);
    if (smem > 220 * 1024) return PP_OK;
    if (conv) {
        // the transposed weight matrix must fit the two TMA stages it reuses; 8 output channels per thread step
        const int D = 2 * radius + 1;
        const size_t stage_words = 2 * (size_t)(WT<2>::KEY_WORDS + WT<2>::F1_WORDS);
        if (conv->cout <= 0 || conv->cout % 8 != 0 || (size_t)conv->cout * L * D * D > stage_words) return PP_OK;
    }
    WTileMaps maps;
    WTileParams p{};
    if (int rc = make_map(&maps.f1, f1t, N, H, W, C, 8, 8)) return rc;
    for (int l = 0; l < L; ++l) {
        if (int rc = make_map(&maps.f2s[l], f2t_levels[l], N, H >> l, W >> l, C, WT<2>::RSS, WT<2>::RSS)) return rc;
        if (int rc = make_map(&maps.f2b[l], f2t_levels[l], N, H >> l, W >> l, C, WT<2>::RSB, WT<2>::RSB)) return rc;
        p.f2t[l] = static_cast<const float*>(f2t_levels[l]);
        p.hl[l] = H >> l;
        p.wl[l] = W >> l;
    }
    p.f1t = f1t;
    p.flow = flow;
    p.out = out;
    p.N = N;
    p.C = C;
    p.H = H;
    p.W = W;
    p.HW = H * W;
    p.L = L;
    p.scale = 1.0f / sqrtf((float)C);
    p.tiles_x = (W + 7) / 8;
    p.tiles_y = (H + 7) / 8;
    if (conv) p.conv = *conv;
    *handled = true;
    return radius == 1 ? launch<1>(maps, p, st) : launch<2>(maps, p, st);
}

}  // namespace pp
