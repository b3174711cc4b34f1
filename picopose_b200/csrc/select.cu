// Hypothesis selection: the gathers that hand the top-k template views of stage 1 to stages 2/3.
//
// Replaces Net.select_template_data (reference model/picopose.py:52-70): six torch.gather calls, each on an index
// tensor `pred_id_src[:, k][:, None, ...].repeat(1, 1, ...)` as large as its output, once per hypothesis (:107-110) --
// 6 * k launches plus 6 * k index tensors per detection batch.  Here ONE launch copies, for every (detection,
// hypothesis) and every per-view tensor, the selected view: a pure byte gather, 16-byte vectorised.
#include "pp_common.cuh"

namespace pp {

constexpr int SEL_MAX_TENSORS = 16;

struct SelectParams {
    const char* src[SEL_MAX_TENSORS];   // tensor i: (B, N, view_bytes[i]) contiguous
    char* dst[SEL_MAX_TENSORS];         //           (rows, view_bytes[i])
    long long view_bytes[SEL_MAX_TENSORS];
    int vec[SEL_MAX_TENSORS];           // 16-byte path usable
    int n_tensors, B, N, K, hyp_sel;
    const long long* pred_id;           // (B, K) int64
};

// blockIdx.y = row * n_tensors + tensor; blockIdx.x strides over the view's bytes.
// rows: hyp_sel >= 0 -> row b copies view pred_id[b, hyp_sel];  hyp_sel < 0 -> row k * B + b copies view pred_id[b, k]
__global__ void __launch_bounds__(256) select_views_kernel(const SelectParams p) {
    const int ti = blockIdx.y % p.n_tensors;
    const int row = blockIdx.y / p.n_tensors;
    const int b = p.hyp_sel >= 0 ? row : row % p.B;
    const int k = p.hyp_sel >= 0 ? p.hyp_sel : row / p.B;
    long long v = __ldg(p.pred_id + (size_t)b * p.K + k);
    v = v < 0 ? 0 : (v >= p.N ? p.N - 1 : v);   // indices come from our own top-k; clamped so that a bad one cannot read out of bounds
    const long long nb = p.view_bytes[ti];
    const char* s = p.src[ti] + ((size_t)b * p.N + (size_t)v) * nb;
    char* d = p.dst[ti] + (size_t)row * nb;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p.vec[ti]) {
        const uint4* s4 = reinterpret_cast<const uint4*>(s);
        uint4* d4 = reinterpret_cast<uint4*>(d);
        for (long long i = i0; i < nb / 16; i += stride) d4[i] = __ldg(s4 + i);
    } else {
        for (long long i = i0; i < nb; i += stride) d[i] = s[i];
    }
}

}  // namespace pp

extern "C" int pp_select_templates(const void* const* src_ptrs, void* const* dst_ptrs, const int64_t* view_bytes,
                                   int n_tensors, int B, int N, const int64_t* pred_id, int K, int hyp_sel, void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (B == 0 || n_tensors == 0) return PP_OK;
    PP_CHECK_ARG(src_ptrs && dst_ptrs && view_bytes && pred_id, "pp_select_templates: null pointer");
    PP_CHECK_ARG(n_tensors > 0 && n_tensors <= SEL_MAX_TENSORS, "pp_select_templates: 1..%d tensors per call (got %d)",
                 SEL_MAX_TENSORS, n_tensors);
    PP_CHECK_ARG(B > 0 && N > 0 && K > 0 && hyp_sel < K, "pp_select_templates: bad sizes (B=%d, N=%d, K=%d, hyp=%d)", B, N, K, hyp_sel);
    SelectParams p{};
    long long max_bytes = 0;
    for (int i = 0; i < n_tensors; ++i) {
        PP_CHECK_ARG(src_ptrs[i] && dst_ptrs[i] && view_bytes[i] > 0, "pp_select_templates: bad tensor %d", i);
        p.src[i] = static_cast<const char*>(src_ptrs[i]);
        p.dst[i] = static_cast<char*>(dst_ptrs[i]);
        p.view_bytes[i] = view_bytes[i];
        p.vec[i] = view_bytes[i] % 16 == 0 && (reinterpret_cast<uintptr_t>(src_ptrs[i]) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dst_ptrs[i]) & 15) == 0;
        const long long units = p.vec[i] ? view_bytes[i] / 16 : view_bytes[i];
        max_bytes = units > max_bytes ? units : max_bytes;
    }
    p.n_tensors = n_tensors;
    p.B = B;
    p.N = N;
    p.K = K;
    p.hyp_sel = hyp_sel;
    p.pred_id = reinterpret_cast<const long long*>(pred_id);
    const int rows = hyp_sel >= 0 ? B : B * K;
    PP_CHECK_ARG((long long)rows * n_tensors <= 65535, "pp_select_templates: too many (row, tensor) pairs (%d x %d)", rows, n_tensors);
    long long gx = (max_bytes + 255) / 256;
    gx = gx < 1 ? 1 : (gx > 64 ? 64 : gx);     // up to 64 blocks per (row, tensor): the largest views are ~600 KB
    select_views_kernel<<<dim3((unsigned)gx, (unsigned)(rows * n_tensors)), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    PP_LAUNCHED();
    return PP_OK;
}
