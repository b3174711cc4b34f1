// Fused stage-3 correlation: CorrelationPyramid + CorrLookup without the all-pairs volume.
//
// The reference builds corr[q, key] = <f1[:, q], f2[:, key]> / sqrt(C) for ALL keys (64 MiB per sample at
// 64 x 64, model/stage3/raft_decoder.py:43-47), average-pools it into a pyramid (:49-51) and then reads
// ~(2r+2)^2 entries per query and level (utils/corr_lookup.py:123-130).  Dot product, 2x2 average pooling
// and bilinear sampling are all linear, so
//     lookup[q, l, a, b] = sum_{4 taps} w_tap * < f1[:, q], pool_l(f2)[:, y_tap, x_tap] > / sqrt(C)
// needs only the (D+2)^2 integer neighbours of the window origin in the l-times pooled FEATURE map.
// pp_windowed_correlation_prepare lays features out position-major ((N, H_l*W_l, C), pooled), so one
// neighbour is one contiguous C-vector; windowed_corr_kernel gives each warp a query: 8-lane groups take one
// neighbour each and split its channels (128 contiguous bytes per step, f1[q] staged in shared memory), three
// shuffles finish the dot product, the D*D bilinear samples are blended from the (D+2)^2 grid exactly as
// corr_lookup.cu does (same tap arithmetic), and a block's 32 queries leave through a shared tile so that
// every output channel is stored as one coalesced 128-byte line.
// Bandwidth-bound on L1/L2 (feature maps are a few MB); no GEMM shape to exploit -- 30x fewer FLOPs than
// the all-pairs product and nothing but features and the (B, L*D*D, H, W) result touches HBM.
#include "pp_common.cuh"

#include <cstdlib>
#include <cstring>

namespace pp {

int launch_wcorr_tiled(int radius, const float* f1t, const void* const* f2t_levels, int L, const float* flow, int N, int C,
                       int H, int W, float* out, cudaStream_t st, bool* handled, const WConv* conv = nullptr);  // windowed_corr_tiled.cu

constexpr int WC_MAX_LEVELS = 8;
constexpr int WC_QUERIES = 32;  // queries per block (one output line)
constexpr int WC_WARPS = 8;

struct WCorrParams {
    const float* f1t;                 // (N, H*W, C)
    const float* f2t[WC_MAX_LEVELS];  // level l: (N, hl*wl, C)
    int hl[WC_MAX_LEVELS], wl[WC_MAX_LEVELS];
    const float* flow;                // (N, 2, H, W)
    float* out;                       // (N, L*D*D, H, W)
    int N, C, H, W, HW, L, radius;
    float scale;                      // 1 / sqrt(C)
    int groups_per_n;
};

// same arithmetic as corr_lookup.cu::axis_tap (reference float round trip, clamped for far-away windows)
__device__ __forceinline__ void wc_axis_tap(float p, int size, int& i0, float& w1) {
    float den = (float)(size > 1 ? size - 1 : 1);
    float g = __fsub_rn(__fdiv_rn(__fmul_rn(p, 2.0f), den), 1.0f);
    float i = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
    float f = floorf(i);
    w1 = __fsub_rn(i, f);
    f = fminf(fmaxf(f, -2.0f), (float)size);
    i0 = (f == f) ? (int)f : -2;
    // a non-finite coordinate leaves w1 = NaN: every sample that uses this tap is NaN, as in F.grid_sample
}

// (N, C, H, W) -> (N, (H>>l)*(W>>l), C): 2^l x 2^l average pooling (what l AvgPool2d(2,2) steps do to the
// volume, applied to the features instead) and transposition to position-major.
struct WPrepJobs {
    const float* src[WC_MAX_LEVELS + 1];
    float* dst[WC_MAX_LEVELS + 1];
    int level[WC_MAX_LEVELS + 1];
    int N;
};

// blockIdx.z = job * N + n: job 0 transposes feat1, job j >= 1 pools feat2 to level j-1 and transposes it
__global__ void __launch_bounds__(256) wcorr_prepare_kernel(const WPrepJobs jobs, int C, int H, int W) {
    __shared__ float tile[32][33];
    const int job = blockIdx.z / jobs.N;
    const int n = blockIdx.z - job * jobs.N;
    const float* __restrict__ f = jobs.src[job];
    float* __restrict__ out = jobs.dst[job];
    const int level = jobs.level[job];
    const int hl = H >> level, wl = W >> level, P = hl * wl, s = 1 << level;
    const int pos0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    if (pos0 >= P) return;  // grid.x is sized for level 0
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const float inv = 1.0f / (float)(s * s);
    for (int cy = ty; cy < 32; cy += 8) {
        const int c = c0 + cy, pos = pos0 + tx;
        float v = 0.f;
        if (c < C && pos < P) {
            const int y = pos / wl, x = pos - y * wl;
            const float* src = f + ((size_t)n * C + c) * H * W + (size_t)(y * s) * W + x * s;
            for (int dy = 0; dy < s; ++dy)
                for (int dx = 0; dx < s; ++dx) v += src[dy * W + dx];
            v *= inv;
        }
        tile[cy][tx] = v;
    }
    __syncthreads();
    for (int py = ty; py < 32; py += 8) {
        const int pos = pos0 + py, c = c0 + tx;
        if (pos < P && c < C) out[((size_t)n * P + pos) * C + c] = tile[tx][py];
    }
}

// Faster form of the above for L <= 4, C % 4 == 0, W % 4 == 0: a block takes one 8 x 32 patch of positions and 32
// channels of one sample, reads it ONCE and writes every level: feat1 -> level 0 only, feat2 -> levels 0 .. L-1,
// each level pooled from the one below it like the reference's repeated AvgPool2d(2, 2).
struct WPatchArgs {
    const float* src[2];            // feat1, feat2 (N, C, H, W)
    float* dst[2][4];               // [which][level]: (N, (H>>l)*(W>>l), C)
    int levels[2];                  // 1, L
    int N, C, H, W, patches_x;
};

constexpr int WP_COLS = 32, WP_CH = 32;   // patch: 8 rows x 32 columns of positions, 32 channels
constexpr int WP_PITCH = WP_CH + 4;         // floats per position in shared memory: 16-byte rows, conflict-free both ways
__global__ void __launch_bounds__(256) wcorr_prepare_patch_kernel(const WPatchArgs a) {
    // In: warp = patch row, lane = column; per channel one coalesced 128-byte load per warp; a thread collects four
    // channels of its position and writes them as one 16-byte shared store.  Out: a quarter warp reads the 32
    // channels of a position as 128 contiguous shared bytes and stores them as one full line.  Pooled levels are
    // computed from the staged level below.
    __shared__ __align__(16) float s0[8 * WP_COLS][WP_PITCH];          // level 0: [position][channel]
    __shared__ __align__(16) float s1[4 * (WP_COLS / 2)][WP_PITCH];    // level 1
    __shared__ __align__(16) float s2[2 * (WP_COLS / 4)][WP_PITCH];    // level 2
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    grid_dependency_wait();  // programmatic dependent launch: the features come from the kernel before this one
    const int which = blockIdx.z / a.N, n = blockIdx.z - which * a.N;
    const int py0 = (blockIdx.x / a.patches_x) * 8, px0 = (blockIdx.x % a.patches_x) * WP_COLS;
    const int c0 = blockIdx.y * WP_CH;
    const int C = a.C, H = a.H, W = a.W;
    const int L = a.levels[which];
    const size_t plane = (size_t)H * W;
    {
        const int y = py0 + warp, x = px0 + lane;
        const bool in = y < H && x < W;
        const float* __restrict__ src = a.src[which] + ((size_t)n * C + c0) * plane + (size_t)y * W + x;
        float v[WP_CH];
#pragma unroll
        for (int k = 0; k < WP_CH; ++k) v[k] = (in && c0 + k < C) ? __ldg(src + k * plane) : 0.f;
#pragma unroll
        for (int q = 0; q < WP_CH / 4; ++q)
            *reinterpret_cast<float4*>(&s0[warp * WP_COLS + lane][q * 4]) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    __syncthreads();
    const int q = tid & 7, cc = c0 + q * 4;  // this thread's channel quad on the way out
    for (int pos = tid >> 3; pos < 8 * WP_COLS; pos += 32) {
        const int y = py0 + pos / WP_COLS, x = px0 + pos % WP_COLS;
        if (y < H && x < W && cc < C)
            __stcs(reinterpret_cast<float4*>(a.dst[which][0] + ((size_t)n * plane + (size_t)y * W + x) * C + cc),
                   *reinterpret_cast<const float4*>(&s0[pos][q * 4]));
    }
    if (L <= 1) return;  // block-uniform
    auto pool4 = [](const float4& q0, const float4& q1, const float4& q2, const float4& q3) {
        return make_float4(0.25f * ((q0.x + q1.x) + (q2.x + q3.x)), 0.25f * ((q0.y + q1.y) + (q2.y + q3.y)),
                           0.25f * ((q0.z + q1.z) + (q2.z + q3.z)), 0.25f * ((q0.w + q1.w) + (q2.w + q3.w)));
    };
    {   // level 1: 4 x 16 positions
        const int H1 = H >> 1, W1 = W >> 1;
        for (int pos = tid >> 3; pos < 4 * (WP_COLS / 2); pos += 32) {
            const int Y = pos / (WP_COLS / 2), X = pos % (WP_COLS / 2);
            const int i00 = (2 * Y) * WP_COLS + 2 * X;
            const float4 o = pool4(*reinterpret_cast<const float4*>(&s0[i00][q * 4]), *reinterpret_cast<const float4*>(&s0[i00 + 1][q * 4]),
                                   *reinterpret_cast<const float4*>(&s0[i00 + WP_COLS][q * 4]), *reinterpret_cast<const float4*>(&s0[i00 + WP_COLS + 1][q * 4]));
            *reinterpret_cast<float4*>(&s1[pos][q * 4]) = o;
            const int gy = (py0 >> 1) + Y, gx = (px0 >> 1) + X;
            if (gy < H1 && gx < W1 && cc < C)
                __stcs(reinterpret_cast<float4*>(a.dst[which][1] + ((size_t)n * H1 * W1 + (size_t)gy * W1 + gx) * C + cc), o);
        }
    }
    if (L <= 2) return;
    __syncthreads();
    {   // level 2: 2 x 8 positions
        const int H2 = H >> 2, W2 = W >> 2;
        for (int pos = tid >> 3; pos < 2 * (WP_COLS / 4); pos += 32) {
            const int Y = pos / (WP_COLS / 4), X = pos % (WP_COLS / 4);
            const int i00 = (2 * Y) * (WP_COLS / 2) + 2 * X;
            const float4 o = pool4(*reinterpret_cast<const float4*>(&s1[i00][q * 4]), *reinterpret_cast<const float4*>(&s1[i00 + 1][q * 4]),
                                   *reinterpret_cast<const float4*>(&s1[i00 + WP_COLS / 2][q * 4]), *reinterpret_cast<const float4*>(&s1[i00 + WP_COLS / 2 + 1][q * 4]));
            *reinterpret_cast<float4*>(&s2[pos][q * 4]) = o;
            const int gy = (py0 >> 2) + Y, gx = (px0 >> 2) + X;
            if (gy < H2 && gx < W2 && cc < C)
                __stcs(reinterpret_cast<float4*>(a.dst[which][2] + ((size_t)n * H2 * W2 + (size_t)gy * W2 + gx) * C + cc), o);
        }
    }
    if (L <= 3) return;
    __syncthreads();
    if (tid < 8 * (WP_COLS / 8)) {  // level 3: 1 x 4 positions
        const int H3 = H >> 3, W3 = W >> 3, X = tid >> 3;
        const float4 o = pool4(*reinterpret_cast<const float4*>(&s2[2 * X][q * 4]), *reinterpret_cast<const float4*>(&s2[2 * X + 1][q * 4]),
                               *reinterpret_cast<const float4*>(&s2[WP_COLS / 4 + 2 * X][q * 4]), *reinterpret_cast<const float4*>(&s2[WP_COLS / 4 + 2 * X + 1][q * 4]));
        const int gy = py0 >> 3, gx = (px0 >> 3) + X;
        if (gy < H3 && gx < W3 && cc < C)
            __stcs(reinterpret_cast<float4*>(a.dst[which][3] + ((size_t)n * H3 * W3 + (size_t)gy * W3 + gx) * C + cc), o);
    }
}

// CV > 0: C == 32*CV and the warp keeps f1[:, q] in registers (lane gl of every 8-lane group holds channels
// gl*4 + 32*j); CV == 0: any C % 4 == 0, f1[:, q] staged in shared memory.
template <int R, int CV>
__global__ void __launch_bounds__(WC_WARPS * 32) windowed_corr_kernel(const WCorrParams p) {
    constexpr int D = 2 * R + 1, G = D + 2, DD = D * D;
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane >> 3, gl = lane & 7;
    const int rows = p.L * DD;
    const int f1_words = CV > 0 ? 0 : WC_WARPS * p.C;
    float* out_tile = smem;                                                       // [rows][33]
    float* f1s = smem + (((size_t)rows * 33 + 3) & ~(size_t)3) + (size_t)warp * (CV > 0 ? 0 : p.C);   // [warps][C] (CV == 0)
    float* vals = smem + (((size_t)rows * 33 + 3) & ~(size_t)3) + f1_words + warp * (G * G + 4 * D);
    int* s_xo = reinterpret_cast<int*>(vals + G * G);
    int* s_yo = s_xo + D;
    float* s_xw = reinterpret_cast<float*>(s_yo + D);
    float* s_yw = s_xw + D;
    grid_dependency_wait();  // programmatic dependent launch: the layout pass before this kernel wrote what it reads

    const int n = blockIdx.x / p.groups_per_n;
    const int hw0 = (blockIdx.x - n * p.groups_per_n) * WC_QUERIES;

    for (int qi = 0; qi < WC_QUERIES / WC_WARPS; ++qi) {
        const int ql = qi * WC_WARPS + warp;  // the 8 warps work on 8 neighbouring queries at a time (shared L1 footprint)
        const int hw = hw0 + ql;
        if (hw >= p.HW) continue;  // warp-uniform
        const int qh = hw / p.W, qw = hw - qh * p.W;
        const float cx = __fadd_rn((float)qw, __ldg(p.flow + ((size_t)n * 2 + 0) * p.HW + hw));
        const float cy = __fadd_rn((float)qh, __ldg(p.flow + ((size_t)n * 2 + 1) * p.HW + hw));
        const float* f1q = p.f1t + ((size_t)n * p.HW + hw) * p.C;
        float4 a[CV > 0 ? CV : 1];
        if (CV > 0) {
#pragma unroll
            for (int j = 0; j < CV; ++j) a[j] = __ldg(reinterpret_cast<const float4*>(f1q + gl * 4 + 32 * j));
        } else {
            __syncwarp();
            for (int c = lane * 4; c < p.C; c += 128) *reinterpret_cast<float4*>(f1s + c) = __ldg(reinterpret_cast<const float4*>(f1q + c));
            __syncwarp();
        }

        for (int l = 0; l < p.L; ++l) {
            const int Hl = p.hl[l], Wl = p.wl[l];
            const float inv = 1.0f / (float)(1 << l);
            if (lane < D) {
                int i0;
                float w1;
                wc_axis_tap(__fadd_rn(__fmul_rn(cx, inv), (float)(lane - R)), Wl, i0, w1);
                s_xo[lane] = i0;
                s_xw[lane] = w1;
                wc_axis_tap(__fadd_rn(__fmul_rn(cy, inv), (float)(lane - R)), Hl, i0, w1);
                s_yo[lane] = i0;
                s_yw[lane] = w1;
            }
            __syncwarp();
            const int xmin = s_xo[0], ymin = s_yo[0];
            const int gx = s_xo[D - 1] + 2 - xmin, gy = s_yo[D - 1] + 2 - ymin;  // grid of integer neighbours (<= G each)
            const float rgx = 1.0f / (float)gx;
            const float* f2n = p.f2t[l] + (size_t)n * Hl * Wl * p.C + gl * 4;
            // correlation with every neighbour of the grid: 4 neighbours per trip, 8 lanes each (a lane group reads
            // 128 contiguous bytes of its neighbour's feature vector per step; 3 shuffles finish the dot product)
            const int npts = gx * gy;
#pragma unroll 2
            for (int idx = 0; idx < npts; idx += 4) {
                const int id = idx + grp;
                const int gyi = (int)(((float)id + 0.5f) * rgx);  // id / gx, exact for these small integers
                const int gxi = id - gyi * gx;
                const int x = xmin + gxi, y = ymin + gyi;
                const bool ok = id < npts && (unsigned)x < (unsigned)Wl && (unsigned)y < (unsigned)Hl;
                float acc = 0.f;
                if (CV > 0) {
                    // always load (key 0 when the neighbour is outside the map) and discard: no predicated zero-fill
                    const float4* v = reinterpret_cast<const float4*>(f2n + (size_t)(ok ? y * Wl + x : 0) * p.C);
                    float4 b[CV > 0 ? CV : 1];
#pragma unroll
                    for (int j = 0; j < CV; ++j) b[j] = __ldg(v + 8 * j);
                    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < CV; ++j) {
                        s4.x = fmaf(a[j].x, b[j].x, s4.x);
                        s4.y = fmaf(a[j].y, b[j].y, s4.y);
                        s4.z = fmaf(a[j].z, b[j].z, s4.z);
                        s4.w = fmaf(a[j].w, b[j].w, s4.w);
                    }
                    acc = ok ? (s4.x + s4.y) + (s4.z + s4.w) : 0.f;
                } else if (ok) {
                    const float* v = f2n + (size_t)(y * Wl + x) * p.C;
#pragma unroll 4
                    for (int c = 0; c + gl * 4 < p.C; c += 32) {
                        const float4 av = *reinterpret_cast<const float4*>(f1s + gl * 4 + c);
                        const float4 bv = __ldg(reinterpret_cast<const float4*>(v + c));
                        acc = fmaf(av.x, bv.x, acc);
                        acc = fmaf(av.y, bv.y, acc);
                        acc = fmaf(av.z, bv.z, acc);
                        acc = fmaf(av.w, bv.w, acc);
                    }
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                if (gl == 0 && id < npts) vals[id] = acc * p.scale;
            }
            __syncwarp();
            // D*D bilinear samples from the grid (separable blend, as in corr_lookup.cu)
            for (int k = lane; k < DD; k += 32) {
                const int ai = k / D, bi = k - ai * D;
                const float* t0 = vals + (s_yo[bi] - ymin) * gx + (s_xo[ai] - xmin);
                const float wx1 = s_xw[ai], wx0 = __fsub_rn(1.0f, wx1);
                const float wy1 = s_yw[bi], wy0 = __fsub_rn(1.0f, wy1);
                const float h0 = fmaf(t0[1], wx1, t0[0] * wx0);
                const float h1 = fmaf(t0[gx + 1], wx1, t0[gx] * wx0);
                out_tile[(size_t)(l * DD + k) * 33 + ql] = fmaf(h1, wy1, h0 * wy0);
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // one coalesced line per output channel: out[n, row, hw0 .. hw0+31]
    float* out_n = p.out + (size_t)n * rows * p.HW;
    for (int row = warp; row < rows; row += WC_WARPS) {
        if (hw0 + lane < p.HW) __stcs(out_n + (size_t)row * p.HW + hw0 + lane, out_tile[(size_t)row * 33 + lane]);
    }
}

template <int R, int CV>
static int launch_wcorr_cv(const WCorrParams& p, cudaStream_t st) {
    constexpr int D = 2 * R + 1, G = D + 2;
    const size_t words = (((size_t)p.L * D * D * 33 + 3) & ~(size_t)3) + (CV > 0 ? 0 : (size_t)WC_WARPS * p.C) +
                         (size_t)WC_WARPS * (G * G + 4 * D);
    const size_t smem = words * sizeof(float);
    PP_CHECK_ARG(smem <= 200 * 1024, "pp_windowed_correlation: %zu bytes of shared memory needed (levels x window too large)", smem);
    PP_CUDA(cudaFuncSetAttribute(windowed_corr_kernel<R, CV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PP_CUDA(launch_dependent(windowed_corr_kernel<R, CV>, dim3(p.N * p.groups_per_n), dim3(WC_WARPS * 32), smem, st, p));
    PP_LAUNCHED();
    return PP_OK;
}

template <int R>
static int launch_wcorr(const WCorrParams& p, cudaStream_t st) {
    switch (p.C) {  // the feature widths PicoPose uses get register-resident f1
        case 256: return launch_wcorr_cv<R, 8>(p, st);
        case 128: return launch_wcorr_cv<R, 4>(p, st);
        case 64: return launch_wcorr_cv<R, 2>(p, st);
        case 32: return launch_wcorr_cv<R, 1>(p, st);
        default: return launch_wcorr_cv<R, 0>(p, st);
    }
}

}  // namespace pp

extern "C" int pp_windowed_correlation_prepare(const float* feat, int N, int C, int H, int W, int level, float* out,
                                               void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (N == 0) return PP_OK;
    PP_CHECK_ARG(feat && out, "pp_windowed_correlation_prepare: null pointer");
    PP_CHECK_ARG(N > 0 && N <= 65535 && C > 0 && H > 0 && W > 0 && level >= 0 && (H >> level) > 0 && (W >> level) > 0,
                 "pp_windowed_correlation_prepare: bad shape");
    const int P = (H >> level) * (W >> level);
    dim3 grid((P + 31) / 32, (C + 31) / 32, N);
    WPrepJobs jobs{};
    jobs.src[0] = feat;
    jobs.dst[0] = out;
    jobs.level[0] = level;
    jobs.N = N;
    wcorr_prepare_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(jobs, C, H, W);
    PP_LAUNCHED();
    return PP_OK;
}

extern "C" int pp_windowed_correlation_prepare_all(const float* feat1, const float* feat2, int N, int C, int H, int W, int L,
                                                   float* f1t, void* const* f2t_levels, void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (N == 0) return PP_OK;
    PP_CHECK_ARG(feat1 && feat2 && f1t && f2t_levels, "pp_windowed_correlation_prepare_all: null pointer");
    PP_CHECK_ARG(L >= 1 && L <= WC_MAX_LEVELS && N > 0 && (long long)N * (L + 1) <= 65535 && C > 0 && H > 0 && W > 0 &&
                     (H >> (L - 1)) > 0 && (W >> (L - 1)) > 0,
                 "pp_windowed_correlation_prepare_all: bad shape");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool aligned = ((reinterpret_cast<uintptr_t>(feat1) | reinterpret_cast<uintptr_t>(feat2) | reinterpret_cast<uintptr_t>(f1t)) & 15) == 0;
    if (L <= 4 && C % 4 == 0 && W % 4 == 0 && aligned && 2LL * N <= 65535) {
        WPatchArgs a{};
        a.src[0] = feat1;
        a.src[1] = feat2;
        a.dst[0][0] = f1t;
        a.levels[0] = 1;
        a.levels[1] = L;
        bool ok = true;
        for (int l = 0; l < L; ++l) {
            PP_CHECK_ARG(f2t_levels[l], "pp_windowed_correlation_prepare_all: null level %d", l);
            a.dst[1][l] = static_cast<float*>(f2t_levels[l]);
            ok = ok && (reinterpret_cast<uintptr_t>(f2t_levels[l]) & 15) == 0;
        }
        if (ok) {
            a.N = N;
            a.C = C;
            a.H = H;
            a.W = W;
            a.patches_x = (W + WP_COLS - 1) / WP_COLS;
            dim3 grid(a.patches_x * ((H + 7) / 8), (C + WP_CH - 1) / WP_CH, 2 * N);
            PP_CUDA(launch_dependent(wcorr_prepare_patch_kernel, grid, dim3(256), 0, st, a));
            PP_LAUNCHED();
            return PP_OK;
        }
    }
    WPrepJobs jobs{};
    jobs.N = N;
    jobs.src[0] = feat1;
    jobs.dst[0] = f1t;
    jobs.level[0] = 0;
    for (int l = 0; l < L; ++l) {
        PP_CHECK_ARG(f2t_levels[l], "pp_windowed_correlation_prepare_all: null level %d", l);
        jobs.src[l + 1] = feat2;
        jobs.dst[l + 1] = static_cast<float*>(f2t_levels[l]);
        jobs.level[l + 1] = l;
    }
    dim3 grid((H * W + 31) / 32, (C + 31) / 32, N * (L + 1));
    wcorr_prepare_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(jobs, C, H, W);
    PP_LAUNCHED();
    return PP_OK;
}

extern "C" int pp_windowed_correlation(const float* f1t, const void* const* f2t_levels, int L, const float* flow, int N,
                                       int C, int H, int W, int radius, float* out, void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (N == 0) return PP_OK;
    PP_CHECK_ARG(f1t && f2t_levels && flow && out, "pp_windowed_correlation: null pointer");
    PP_CHECK_ARG(L >= 1 && L <= WC_MAX_LEVELS, "pp_windowed_correlation: 1 <= levels <= %d", WC_MAX_LEVELS);
    PP_CHECK_ARG(radius >= 1 && radius <= 8, "pp_windowed_correlation: 1 <= radius <= 8 (got %d)", radius);
    PP_CHECK_ARG(C > 0 && C % 4 == 0 && C <= 2048, "pp_windowed_correlation: feature dim must be a multiple of 4, <= 2048");
    PP_CHECK_ARG(N > 0 && H > 0 && W > 0 && (H >> (L - 1)) > 0 && (W >> (L - 1)) > 0, "pp_windowed_correlation: bad shape");
    PP_CHECK_ARG((reinterpret_cast<uintptr_t>(f1t) & 15) == 0, "pp_windowed_correlation: features must be 16-byte aligned");
    WCorrParams p{};
    p.f1t = f1t;
    for (int l = 0; l < L; ++l) {
        PP_CHECK_ARG(f2t_levels[l] && (reinterpret_cast<uintptr_t>(f2t_levels[l]) & 15) == 0, "pp_windowed_correlation: bad level %d", l);
        p.f2t[l] = static_cast<const float*>(f2t_levels[l]);
        p.hl[l] = H >> l;
        p.wl[l] = W >> l;
    }
    p.flow = flow;
    p.out = out;
    p.N = N;
    p.C = C;
    p.H = H;
    p.W = W;
    p.HW = H * W;
    p.L = L;
    p.radius = radius;
    p.scale = 1.0f / sqrtf((float)C);
    p.groups_per_n = (p.HW + WC_QUERIES - 1) / WC_QUERIES;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // The tiled kernel (windowed_corr_tiled.cu) when it covers the problem and there are enough 8 x 8 tiles to
    // occupy the GPU; PICOPOSE_WCORR_KERNEL=tiled|direct forces a choice (tests, tools/bench_stage3.py).
    int want_tiled = -1;
    if (const char* e = getenv("PICOPOSE_WCORR_KERNEL")) want_tiled = strcmp(e, "tiled") == 0 ? 1 : strcmp(e, "direct") == 0 ? 0 : -1;
    const long long tiles = (long long)N * ((W + 7) / 8) * ((H + 7) / 8);
    if (want_tiled == 1 || (want_tiled < 0 && tiles >= 48)) {
        bool handled = false;
        if (int rc = launch_wcorr_tiled(radius, f1t, f2t_levels, L, flow, N, C, H, W, out, st, &handled)) return rc;
        if (handled) return PP_OK;
        PP_CHECK_ARG(want_tiled != 1, "pp_windowed_correlation: the tiled kernel does not cover radius %d, C %d, L %d", radius, C, L);
    }
    switch (radius) {
        case 1: return launch_wcorr<1>(p, st);
        case 2: return launch_wcorr<2>(p, st);
        case 3: return launch_wcorr<3>(p, st);
        case 4: return launch_wcorr<4>(p, st);
        case 5: return launch_wcorr<5>(p, st);
        case 6: return launch_wcorr<6>(p, st);
        case 7: return launch_wcorr<7>(p, st);
        default: return launch_wcorr<8>(p, st);
    }
}

// Fused CorrelationPyramid + CorrLookup + first MotionEncoder convolution (model/stage3/flow_decoder.py:59-62 with
// model/stage3/raft_decoder.py:113-116,157): the lookup tile never leaves shared memory.
extern "C" int pp_windowed_correlation_conv1x1(const float* f1t, const void* const* f2t_levels, int L, const float* flow,
                                               int N, int C, int H, int W, int radius, const float* weight, const float* bias,
                                               int cout, int relu, float* out, void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (N == 0) return PP_OK;
    PP_CHECK_ARG(f1t && f2t_levels && flow && weight && out, "pp_windowed_correlation_conv1x1: null pointer");
    PP_CHECK_ARG(L >= 1 && L <= WC_MAX_LEVELS && N > 0 && H > 0 && W > 0 && (H >> (L - 1)) > 0 && (W >> (L - 1)) > 0,
                 "pp_windowed_correlation_conv1x1: bad shape");
    PP_CHECK_ARG((reinterpret_cast<uintptr_t>(f1t) & 15) == 0, "pp_windowed_correlation_conv1x1: features must be 16-byte aligned");
    for (int l = 0; l < L; ++l)
        PP_CHECK_ARG(f2t_levels[l] && (reinterpret_cast<uintptr_t>(f2t_levels[l]) & 15) == 0,
                     "pp_windowed_correlation_conv1x1: bad level %d", l);
    WConv conv;
    conv.weight = weight;
    conv.bias = bias;
    conv.out = out;
    conv.cout = cout;
    conv.relu = relu ? 1 : 0;
    bool handled = false;
    if (int rc = launch_wcorr_tiled(radius, f1t, f2t_levels, L, flow, N, C, H, W, nullptr, static_cast<cudaStream_t>(stream), &handled,
                                    &conv))
        return rc;
    // the fusion lives in the TMA-tiled kernel: r <= 2 (what FlowDecoder uses), C % 32 == 0, L <= 4, cout % 8 == 0 and a
    // weight matrix that fits the kernel's stage buffers; other shapes are the caller's to run unfused
    PP_CHECK_ARG(handled, "pp_windowed_correlation_conv1x1: shape not covered by the fused kernel (radius %d, C %d, L %d, cout %d)",
                 radius, C, L, cout);
    return PP_OK;
}

