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

namespace pp {

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
    if (!(w1 >= 0.0f && w1 <= 1.0f)) w1 = 0.0f;
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

template <int R>
__global__ void __launch_bounds__(WC_WARPS * 32) windowed_corr_kernel(const WCorrParams p) {
    constexpr int D = 2 * R + 1, G = D + 2, DD = D * D;
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows = p.L * DD;
    float* out_tile = smem;                                   // [rows][33]
    float* f1s = smem + (((size_t)rows * 33 + 3) & ~(size_t)3) + (size_t)warp * p.C;   // after the tile: [warps][C]
    float* vals = smem + (((size_t)rows * 33 + 3) & ~(size_t)3) + (size_t)WC_WARPS * p.C + warp * (G * G + 4 * D);
    int* s_xo = reinterpret_cast<int*>(vals + G * G);
    int* s_yo = s_xo + D;
    float* s_xw = reinterpret_cast<float*>(s_yo + D);
    float* s_yw = s_xw + D;

    const int n = blockIdx.x / p.groups_per_n;
    const int hw0 = (blockIdx.x - n * p.groups_per_n) * WC_QUERIES;

    for (int qi = 0; qi < WC_QUERIES / WC_WARPS; ++qi) {
        const int ql = qi * WC_WARPS + warp;  // the 8 warps work on 8 neighbouring queries at a time (shared L1 footprint)
        const int hw = hw0 + ql;
        if (hw >= p.HW) continue;  // warp-uniform
        const int qh = hw / p.W, qw = hw - qh * p.W;
        const float cx = __fadd_rn((float)qw, __ldg(p.flow + ((size_t)n * 2 + 0) * p.HW + hw));
        const float cy = __fadd_rn((float)qh, __ldg(p.flow + ((size_t)n * 2 + 1) * p.HW + hw));
        // stage f1[:, q]: lane keeps channels lane*4 + 128*j in its own 16-byte slots
        const float* f1q = p.f1t + ((size_t)n * p.HW + hw) * p.C;
        __syncwarp();
        for (int c = lane * 4; c < p.C; c += 128) *reinterpret_cast<float4*>(f1s + c) = __ldg(reinterpret_cast<const float4*>(f1q + c));
        __syncwarp();

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
            const float* f2n = p.f2t[l] + (size_t)n * Hl * Wl * p.C;
            // correlation with every neighbour of the grid: 4 neighbours per trip, 8 lanes each (a lane group reads
            // 128 contiguous bytes of its neighbour's feature vector per step; 3 shuffles finish the dot product)
            const int grp = lane >> 3, gl = lane & 7;
            const int npts = gx * gy;
            for (int idx = 0; idx < npts; idx += 4) {
                const int id = idx + grp;
                const int gyi = id / gx, gxi = id - gyi * gx;
                const int x = xmin + gxi, y = ymin + gyi;
                float acc = 0.f;
                if (id < npts && (unsigned)x < (unsigned)Wl && (unsigned)y < (unsigned)Hl) {
                    const float* v = f2n + (size_t)(y * Wl + x) * p.C;
#pragma unroll 4
                    for (int c = gl * 4; c < p.C; c += 32) {
                        const float4 a = *reinterpret_cast<const float4*>(f1s + c);
                        const float4 b = __ldg(reinterpret_cast<const float4*>(v + c));
                        acc = fmaf(a.x, b.x, acc);
                        acc = fmaf(a.y, b.y, acc);
                        acc = fmaf(a.z, b.z, acc);
                        acc = fmaf(a.w, b.w, acc);
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
                const int a = k / D, b = k - a * D;
                const float* t0 = vals + (s_yo[b] - ymin) * gx + (s_xo[a] - xmin);
                const float wx1 = s_xw[a], wx0 = __fsub_rn(1.0f, wx1);
                const float wy1 = s_yw[b], wy0 = __fsub_rn(1.0f, wy1);
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

template <int R>
static int launch_wcorr(const WCorrParams& p, cudaStream_t st) {
    constexpr int D = 2 * R + 1, G = D + 2;
    const size_t words = (((size_t)p.L * D * D * 33 + 3) & ~(size_t)3) + (size_t)WC_WARPS * p.C + (size_t)WC_WARPS * (G * G + 4 * D);
    const size_t smem = words * sizeof(float);
    PP_CHECK_ARG(smem <= 200 * 1024, "pp_windowed_correlation: %zu bytes of shared memory needed (levels x window too large)", smem);
    PP_CUDA(cudaFuncSetAttribute(windowed_corr_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    windowed_corr_kernel<R><<<p.N * p.groups_per_n, WC_WARPS * 32, smem, st>>>(p);
    PP_LAUNCHED();
    return PP_OK;
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
