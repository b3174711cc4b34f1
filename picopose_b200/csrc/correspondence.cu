// Correspondence glue kernels (reference utils/correspondence.py).  Tiny, latency-bound.
#include "pp_common.cuh"

namespace pp {

// compute_init_correspondences, utils/correspondence.py:10-26: patch centres ((i+.5)*patch) through the
// 3x3 affine (utils/torch_utils.py:114-135), / patch, * nearest-resized mask, minus the integer grid.
__global__ void init_corr_kernel(const float* __restrict__ Ms, const float* __restrict__ mask, int B, int Hm,
                                 int Wm, int h, int w, float* __restrict__ flow, float* __restrict__ cert) {
    const int total = B * h * w;
    const int patch = Hm / h;  // utils/correspondence.py:13
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int b = i / (h * w);
        const int rem = i - b * h * w;
        const int y = rem / w, x = rem - y * w;
        const float m = mask[((size_t)b * Hm + nearest_src(y, Hm, h)) * Wm + nearest_src(x, Wm, w)];
        const float cx = (float)(x * patch) + (float)patch * 0.5f;  // utils/torch_utils.py:298-301
        const float cy = (float)(y * patch) + (float)patch * 0.5f;
        const float* M = Ms + (size_t)b * 9;
        const float qx = M[0] * cx + M[1] * cy + M[2];
        const float qy = M[3] * cx + M[4] * cy + M[5];
        const float qz = M[6] * cx + M[7] * cy + M[8];
        const float px = __fdiv_rn(__fdiv_rn(qx, qz), (float)patch);
        const float py = __fdiv_rn(__fdiv_rn(qy, qz), (float)patch);
        flow[((size_t)b * 2 + 0) * h * w + rem] = __fsub_rn(__fmul_rn(px, m), (float)x);
        flow[((size_t)b * 2 + 1) * h * w + rem] = __fsub_rn(__fmul_rn(py, m), (float)y);
        cert[(size_t)b * h * w + rem] = m;
    }
}

// compute_stage3_correspondences, utils/correspondence.py:28-59 (sync-free: no nonzero()).
__global__ void stage3_corr_kernel(const float* __restrict__ flow, const float* __restrict__ cert, int B, int H,
                                   int W, float threshold, long long* __restrict__ tar, long long* __restrict__ src) {
    const int total = B * H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        // thread order follows the OUTPUT layout k = w*H + h so the int64 stores coalesce
        const int b = i / (H * W);
        const int k = i - b * H * W;
        const int x = k / H, y = k - x * H;  // w = x, h = y
        const size_t in = (size_t)y * W + x;
        const float tx = __fadd_rn(flow[((size_t)b * 2 + 0) * H * W + in], (float)x);
        const float ty = __fadd_rn(flow[((size_t)b * 2 + 1) * H * W + in], (float)y);
        const float c = cert[(size_t)b * H * W + in];
        const float sg = 1.0f / (1.0f + expf(-c));
        const bool keep = tx > 0.f && ty > 0.f && tx < (float)(H - 1) && ty < (float)(W - 1) && sg > threshold;
        long long* t = tar + ((size_t)b * H * W + k) * 2;
        long long* s = src + ((size_t)b * H * W + k) * 2;
        // .long() truncates toward zero; kept coordinates are positive and < H
        t[0] = keep ? (long long)tx : -1;
        t[1] = keep ? (long long)ty : -1;
        s[0] = keep ? (long long)x : -1;
        s[1] = keep ? (long long)y : -1;
    }
}

}  // namespace pp

extern "C" int pp_init_correspondences(const float* Ms, const float* tem_mask, int B, int Hm, int Wm, int h, int w,
                                       float* flow, float* certainty, void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (B == 0) return PP_OK;
    PP_CHECK_ARG(Ms && tem_mask && flow && certainty, "pp_init_correspondences: null pointer");
    PP_CHECK_ARG(Hm == Wm, "pp_init_correspondences: mask must be square (reference asserts H == W), got %dx%d", Hm, Wm);
    PP_CHECK_ARG(B >= 0 && h > 0 && w > 0 && Hm >= h, "pp_init_correspondences: bad shape");
    {
        // the reference builds ceil(Hm/patch)^2 patch centres and reshapes them to (w h): sizes must agree
        const int patch = Hm / h;
        const int n = (Hm + patch - 1) / patch;
        PP_CHECK_ARG(n == h && n == w, "pp_init_correspondences: size (%d,%d) does not tile a %d px mask", h, w, Hm);
    }
    if (B == 0) return PP_OK;
    const int total = B * h * w;
    init_corr_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(Ms, tem_mask, B, Hm, Wm, h,
                                                                                       w, flow, certainty);
    PP_LAUNCHED();
    return PP_OK;
}

extern "C" int pp_stage3_correspondences(const float* flow, const float* certainty, int B, int H, int W,
                                         float threshold, int64_t* tar_pts, int64_t* src_pts, void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (B == 0) return PP_OK;
    PP_CHECK_ARG(flow && certainty && tar_pts && src_pts, "pp_stage3_correspondences: null pointer");
    PP_CHECK_ARG(B >= 0 && H > 0 && W > 0, "pp_stage3_correspondences: bad shape");
    if (B == 0) return PP_OK;
    const int total = B * H * W;
    int grid = (total + 255) / 256;
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    stage3_corr_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        flow, certainty, B, H, W, threshold, reinterpret_cast<long long*>(tar_pts),
        reinterpret_cast<long long*>(src_pts));
    PP_LAUNCHED();
    return PP_OK;
}
