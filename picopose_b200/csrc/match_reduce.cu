// Stage-1 reductions around the tensor-core contraction: score finalisation, top-k, pyramid pooling, and the host side
// of the one-call entry points.  (Reference: utils/matching.py:54-68; the stage-2 volume of :22-25 is written by the
// contraction's own epilogue.)
#include "pp_common.cuh"

#include <cstdlib>

namespace pp {

int run_match_gemm(int epi, const void* q_prep, const void* bank_prep, int64_t n_banks, const int32_t* bank_of_det,
                   int B, int N, int T, int Kp, const float* mrow, const int* tv, const int* rowmap, const float* ra,
                   const float* rb,
                   unsigned long long* rowkey, unsigned long long* colkey, float* emit, float emit_scale, int cluster,
                   cudaStream_t st, int emit_tile_w = 0, const int32_t* det_order = nullptr, const float* cmask = nullptr,
                   int Hm = 0, int Wm = 0, int gh = 0);
int prepare_pair_impl(const float* q_feats, const float* s_feats, int64_t G, int C, int P, int mode, void* prepared, float* rnorm,
                      void* stream);

int prepare_query_impl(const float* tar_feat, const float* tar_mask, int B, int C, int H, int W, int Hm, int Wm, int mode,
                       void* q_prep, float* q_rnorm, void* q_meta, void* clear, size_t clear_bytes, void* stream);

// Stable sort of the detections by bank (B <= 1024, one block): order[i] = detection at position i.  Detections that share
// an object bank become neighbours, and the contraction hands neighbours to one cluster so that they share the bank's
// tiles in L2 instead of each streaming the bank from HBM.
__global__ void __launch_bounds__(1024) det_order_kernel(const int32_t* __restrict__ bank_of_det, int B, int32_t* __restrict__ order) {
    __shared__ int s_bank[1024];
    const int i = threadIdx.x;
    if (i < B) s_bank[i] = bank_of_det[i];
    __syncthreads();
    if (i < B) {
        const int mine = s_bank[i];
        int rank = 0;
        for (int j = 0; j < B; ++j) rank += (s_bank[j] < mine) || (s_bank[j] == mine && j < i);
        order[rank] = i;
    }
}

// One block per (b, n): decode the row / column keys, apply the reference's validity rule
// (utils/matching.py:54-60) and the masked mean (:63-67).
__global__ void __launch_bounds__(256)
finalize_scores_kernel(const unsigned long long* __restrict__ rowkey, const unsigned long long* __restrict__ colkey,
                       const float* __restrict__ mrow, const float* __restrict__ ra, const int* __restrict__ rank,
                       const int* __restrict__ fm, int N, int T, float inv_hh, float* __restrict__ sim_avg,
                       float* __restrict__ score_t2s, int32_t* __restrict__ idx_t2s, int32_t* __restrict__ idx_s2t,
                       uint8_t* __restrict__ mutual, int k, int* __restrict__ done, float* __restrict__ topk_score,
                       long long* __restrict__ topk_idx, const int32_t* __restrict__ bank_of_det, int B, int n_banks,
                       int* __restrict__ fault) {
    const size_t bn = blockIdx.x;
    const int b = (int)(bn / N);
    grid_dependency_wait();  // the keys come from the contraction launched just before (programmatic dependent launch)
    if (bn == 0 && bank_of_det) {
        // range check of the caller's bank indices (the contraction clamps them in its loads): an index outside
        // [0, n_banks) is reported through the fault record (code 6), never trapped
        for (int i = threadIdx.x; i < B; i += blockDim.x) {
            const int v = bank_of_det[i];
            if ((v < 0 || v >= n_banks) && fault) {
                fault[1] = i;
                fault[2] = v;
                fault[3] = n_banks;
                fault[4] = 0;
                __threadfence_system();
                fault[0] = 6;
            }
        }
    }
    // masked query rows never entered the contraction; in the reference they are rows of zeros that take part in
    // every column max (utils/matching.py:48,51): value 0 at the first masked patch wins ties by index
    const int fmb = fm[b];
    const unsigned long long zero_key = fmb >= 0 ? pack_key(0.0f, (uint32_t)fmb) : 0ull;
    float sum = 0.f, cnt = 0.f;
    for (int j = threadIdx.x; j < T; j += blockDim.x) {
        const float m = mrow[(size_t)b * T + j];
        const unsigned long long rk = rowkey[bn * T + j];
        unsigned long long ck = colkey[bn * T + j];
        ck = ck > zero_key ? ck : zero_key;
        // a masked query row is all zeros in the reference: max 0 at index 0
        const bool on = m != 0.f;
        // row key holds max_s of the masked similarity m[t] * sim[t, s] (the query operand carries norm and mask)
        const int rj = rank[(size_t)b * T + j];  // compact row of patch j
        const float sc = (on && rk && rj >= 0) ? key_value(rk) * ra[(size_t)b * T + rj] : 0.f;
        const int it = (on && rk) ? (int)key_index(rk) : 0;
        const int is = ck ? (int)key_index(ck) : 0;
        const float valid = (it != 0 && is != 0) ? m : 0.f;  // tar_mask * (idx_src2tar != 0) * (idx_tar2src != 0)
        sum = fmaf(sc, valid, sum);
        cnt += valid;
        if (score_t2s) score_t2s[bn * T + j] = sc;
        if (idx_t2s) idx_t2s[bn * T + j] = it;
        if (idx_s2t) idx_s2t[bn * T + j] = is;
        if (mutual) {
            // mutual nearest neighbours: query patch j -> template patch `it` and back to j (extra output; the
            // reference only applies its argmax != 0 heuristic)
            unsigned long long back = colkey[bn * T + it];
            back = back > zero_key ? back : zero_key;
            mutual[bn * T + j] = (on && rk && back && (int)key_index(back) == j) ? 1 : 0;
        }
    }
    __shared__ float s_sum[8], s_cnt[8];
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_cnt[threadIdx.x >> 5] = cnt; }
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        float ts = 0.f, tc = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ts += s_sum[w]; tc += s_cnt[w]; }
        sim_avg[bn] = tc > 0.f ? ts * inv_hh : 0.f;  // divisor is H*H, not the valid count (:65-67)
        s_last = 0;
        if (k > 0) {
            // the block that finishes a detection's last view ranks that detection (torch.topk, :68): no extra launch
            __threadfence();
            s_last = atomicAdd(done + b, 1) == N - 1;
        }
    }
    if (k == 0) return;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    extern __shared__ unsigned long long s_keys[];  // N packed keys + 8 partials
    unsigned long long* s_red = s_keys + N;
    const volatile float* sv = sim_avg + (size_t)b * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_keys[i] = pack_key(sv[i] + 0.0f, (uint32_t)i);
    __syncthreads();
    for (int r = 0; r < k; ++r) {
        unsigned long long best = 0ull;
        for (int i = threadIdx.x; i < N; i += blockDim.x) best = s_keys[i] > best ? s_keys[i] : best;
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best = s_red[w] > best ? s_red[w] : best;
            const uint32_t idx = key_index(best);
            topk_score[(size_t)b * k + r] = key_value(best);
            topk_idx[(size_t)b * k + r] = (long long)idx;
            s_keys[idx] = 0ull;
        }
        __syncthreads();
    }
}

// Row-wise top-k by repeated block arg-max; ties resolve to the lowest index (torch leaves them unspecified).
__global__ void __launch_bounds__(256)
topk_kernel(const float* __restrict__ scores, int N, int k, long long idx_offset, float* __restrict__ out_score,
            long long* __restrict__ out_idx, double* __restrict__ out_pair, int kpad) {
    extern __shared__ unsigned long long s_keys[];  // N packed keys + 8 partials
    unsigned long long* s_red = s_keys + N;
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_keys[i] = pack_key(scores[(size_t)b * N + i] + 0.0f, (uint32_t)i);
    __syncthreads();
    for (int r = 0; r < k; ++r) {
        unsigned long long best = 0ull;
        for (int i = threadIdx.x; i < N; i += blockDim.x) best = s_keys[i] > best ? s_keys[i] : best;
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best = s_red[w] > best ? s_red[w] : best;
            const uint32_t idx = key_index(best);
            if (out_score) out_score[(size_t)b * k + r] = key_value(best);
            if (out_idx) out_idx[(size_t)b * k + r] = (long long)idx + idx_offset;
            if (out_pair) {  // (score, index) as two doubles: one tensor to all-gather across ranks
                out_pair[((size_t)b * kpad + r) * 2 + 0] = (double)key_value(best);
                out_pair[((size_t)b * kpad + r) * 2 + 1] = (double)((long long)idx + idx_offset);
            }
            s_keys[idx] = 0ull;  // remove the winner
        }
        __syncthreads();
    }
    if (out_pair && threadIdx.x < kpad - k) {  // shard with fewer than kpad views: pad with (-inf, -1)
        out_pair[((size_t)b * kpad + k + threadIdx.x) * 2 + 0] = -INFINITY;
        out_pair[((size_t)b * kpad + k + threadIdx.x) * 2 + 1] = -1.0;
    }
}

// Merge of per-rank candidate lists (R, B, k_in) of (score, index) pairs into the global top-k of each row:
// one warp per row, ties go to the lowest rank / slot (= lowest global view index for contiguous shards).
__global__ void __launch_bounds__(32)
topk_merge_kernel(const double* __restrict__ pairs, int R, int B, int k_in, int k, float* __restrict__ out_score,
                  long long* __restrict__ out_idx) {
    extern __shared__ unsigned long long s_keys[];
    const int b = blockIdx.x, lane = threadIdx.x, n = R * k_in;
    for (int i = lane; i < n; i += 32) {
        const int r = i / k_in, j = i - r * k_in;
        s_keys[i] = pack_key((float)pairs[(((size_t)r * B + b) * k_in + j) * 2] + 0.0f, (uint32_t)i);
    }
    __syncwarp();
    for (int o = 0; o < k; ++o) {
        unsigned long long best = 0ull;
        for (int i = lane; i < n; i += 32) best = s_keys[i] > best ? s_keys[i] : best;
        for (int sft = 16; sft > 0; sft >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, sft);
            best = other > best ? other : best;
        }
        if (lane == 0) {
            const uint32_t pos = key_index(best);
            const int r = pos / k_in, j = pos - r * k_in;
            out_score[(size_t)b * k + o] = key_value(best);
            out_idx[(size_t)b * k + o] = (long long)pairs[(((size_t)r * B + b) * k_in + j) * 2 + 1];
            s_keys[pos] = 0ull;
        }
        __syncwarp();
    }
}

// AvgPool2d(kernel 2, stride 2) of every (h x w) slice: model/stage3/raft_decoder.py:27,49-51
__global__ void avgpool2_kernel(const float* __restrict__ in, long long slices, int h, int w, float* __restrict__ out) {
    const int ho = h >> 1, wo = w >> 1;
    const long long total = slices * ho * wo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long q = i / (ho * wo);
        const int r = (int)(i - q * ho * wo);
        const int y = r / wo, x = r - y * wo;
        const float* s = in + (q * h + 2 * y) * (long long)w + 2 * x;
        out[i] = (s[0] + s[1] + s[w] + s[w + 1]) * 0.25f;
    }
}

// Offset of key (y, x) inside a slice stored as 4 x 8 tiles (32 floats = one 128-byte line per tile), tiles row-major.
__host__ __device__ __forceinline__ int tiled_offset(int y, int x, int w) {
    return ((y >> 2) * (w >> 3) + (x >> 3)) * 32 + (y & 3) * 8 + (x & 7);
}

// AvgPool2d(2, 2) between two TILED levels: a thread makes one output tile row (8 outputs) from a 2 x 16 patch of the
// level below, i.e. 16-byte loads and two 16-byte stores.
__global__ void avgpool2_tiled_kernel(const float* __restrict__ in, long long slices, int h, int w, float* __restrict__ out) {
    const int ho = h >> 1, wo = w >> 1;
    const int rows_per_slice = ho * (wo >> 3);       // output tile rows of 8 floats
    const long long total = slices * rows_per_slice;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long q = i / rows_per_slice;
        const int r = (int)(i - q * rows_per_slice);
        // enumerate in the OUTPUT's memory order: tile (ty, tx), row-in-tile ry
        const int tpr = wo >> 3;
        const int tile = r >> 2, ry = r & 3;
        const int ty = tile / tpr, tx = tile - ty * tpr;
        const int y = ty * 4 + ry, x = tx * 8;
        const float* s = in + q * (long long)h * w;
        float o[8];
#pragma unroll
        for (int half = 0; half < 2; ++half) {          // input columns 2x .. 2x+15 = two input tiles side by side
            const int xi = 2 * x + 8 * half;
            const float4 a0 = *reinterpret_cast<const float4*>(s + tiled_offset(2 * y, xi, w));
            const float4 a1 = *reinterpret_cast<const float4*>(s + tiled_offset(2 * y, xi, w) + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(s + tiled_offset(2 * y + 1, xi, w));
            const float4 b1 = *reinterpret_cast<const float4*>(s + tiled_offset(2 * y + 1, xi, w) + 4);
            // same association as the row-major kernel: (s[0] + s[1] + s[w] + s[w+1]) * 0.25
            o[4 * half + 0] = (a0.x + a0.y + b0.x + b0.y) * 0.25f;
            o[4 * half + 1] = (a0.z + a0.w + b0.z + b0.w) * 0.25f;
            o[4 * half + 2] = (a1.x + a1.y + b1.x + b1.y) * 0.25f;
            o[4 * half + 3] = (a1.z + a1.w + b1.z + b1.w) * 0.25f;
        }
        float* d = out + q * (long long)ho * wo + tiled_offset(y, x, wo);
        *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
}

// Layout change of a stack of (h x w) slices between the reference's row-major form and the tiled one (a permutation
// inside every slice); a thread moves 8 floats.
__global__ void retile_kernel(const float* __restrict__ in, long long slices, int h, int w, int to_tiled, float* __restrict__ out) {
    const int units = h * (w >> 3);
    const long long total = slices * units;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long q = i / units;
        const int u = (int)(i - q * units);
        // u enumerates the TILED order (tile, row in tile): contiguous on the tiled side
        const int tpr = w >> 3;
        const int tile = u >> 2, ry = u & 3;
        const int ty = tile / tpr, tx = tile - ty * tpr;
        const int y = ty * 4 + ry, x = tx * 8;
        const long long base = q * (long long)h * w;
        const long long lin = base + (long long)y * w + x, til = base + (long long)u * 8;
        const float* s = in + (to_tiled ? lin : til);
        float* d = out + (to_tiled ? til : lin);
        const float4 v0 = *reinterpret_cast<const float4*>(s), v1 = *reinterpret_cast<const float4*>(s + 4);
        *reinterpret_cast<float4*>(d) = v0;
        *reinterpret_cast<float4*>(d + 4) = v1;
    }
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace pp

extern "C" size_t pp_match_scores_workspace(int B, int N, int T) {
    if (B < 0 || N < 0 || T < 0) return 0;
    const size_t keys = (size_t)B * N * T * sizeof(unsigned long long);
    // row keys | column keys | per-detection finished-view counters | bank-sorted detection order
    return 2 * pp::align_up(keys, 256) + 2 * pp::align_up((size_t)B * sizeof(int), 256);
}

namespace pp {
static int match_scores_impl(const void* q_prep, const float* q_rnorm, const void* q_meta, const void* bank_prep,
                             const float* bank_rnorm, int64_t n_banks, const int32_t* bank_of_det, int B, int N, int H,
                             int W, int Kp, float* sim_avg, float* score_t2s, int32_t* idx_t2s, int32_t* idx_s2t,
                             uint8_t* mutual_nn, int k, float* topk_score, int64_t* topk_idx, void* workspace,
                             size_t workspace_bytes, int cluster, void* stream, bool keys_cleared = false) {
    if (int rc = require_sm100()) return rc;
    if (B == 0 || N == 0) return PP_OK;
    PP_CHECK_ARG(q_prep && q_rnorm && q_meta && bank_prep && bank_rnorm && sim_avg, "pp_match_scores: null pointer");
    PP_CHECK_ARG(H == W, "pp_match_scores: the reference asserts a square patch grid (H == W), got %dx%d", H, W);
    PP_CHECK_ARG(B > 0 && N > 0 && H > 0, "pp_match_scores: bad shape");
    PP_CHECK_ARG(bank_of_det != nullptr || n_banks == B, "pp_match_scores: identity bank mapping needs n_banks == B");
    const int T = H * W;
    const size_t need = pp_match_scores_workspace(B, N, T);
    if (!workspace || workspace_bytes < need)
        return fail(PP_ERR_WORKSPACE, "pp_match_scores: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
    PP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "pp_match_scores: workspace must be 256-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const QueryMeta qm = split_query_meta(const_cast<void*>(q_meta), B, T);
    const size_t keys = align_up((size_t)B * N * T * sizeof(unsigned long long), 256);
    unsigned long long* rowkey = reinterpret_cast<unsigned long long*>(workspace);
    unsigned long long* colkey = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + keys);
    int* done = reinterpret_cast<int*>(static_cast<char*>(workspace) + 2 * keys);  // per-detection finished-view counters
    int32_t* order = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) + 2 * keys + align_up((size_t)B * sizeof(int), 256));
    // the key arrays and the per-detection counters start at zero (a caller that has already cleared the scratch,
    // e.g. on a forked stream beside the bank prologue, says so)
    if (!keys_cleared) PP_CUDA(cudaMemsetAsync(rowkey, 0, 2 * keys + align_up((size_t)B * sizeof(int), 256), st));
    // bank-sorted detection order: only for launches long enough to run in group mode (run_match_gemm's rule), where
    // detections of one bank share its tiles; short launches deal single tiles in the caller's order
    const long long tiles_bound = (long long)B * N * ((T + 255) / 256) * ((T + 255) / 256);
    const bool sort_dets = bank_of_det != nullptr && B > 1 && B <= 1024 && tiles_bound >= 64LL * (sm_count() / 2);
    if (sort_dets) {
        det_order_kernel<<<1, 1024, 0, st>>>(bank_of_det, B, order);
        PP_LAUNCHED();
    }
    int* fault = nullptr;
    if (bank_of_det)
        if (int rc = fault_buffer(&fault)) return rc;
    if (int rc = run_match_gemm(0, q_prep, bank_prep, n_banks, bank_of_det, B, N, T, Kp, qm.mrow, qm.tv, qm.rowmap, q_rnorm,
                                bank_rnorm, rowkey, colkey, nullptr, 1.0f, cluster, st, 0, sort_dets ? order : nullptr))
        return rc;
    size_t smem = 0;
    if (k > 0) {
        PP_CHECK_ARG(k <= N && N <= 24000 && topk_score && topk_idx, "top-k: need 0 < k <= N <= 24000 (k=%d, N=%d)", k, N);
        smem = ((size_t)N + 8) * sizeof(unsigned long long);
        if (smem > 48 * 1024)
            PP_CUDA(cudaFuncSetAttribute(finalize_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    PP_CUDA(launch_dependent(finalize_scores_kernel, dim3((unsigned)((size_t)B * N)), dim3(256), smem, st, rowkey, colkey, qm.mrow,
                             q_rnorm, qm.rank, qm.fm, N, T, 1.0f / (float)(H * H), sim_avg, score_t2s, idx_t2s, idx_s2t, mutual_nn, k,
                             done, topk_score, reinterpret_cast<long long*>(topk_idx), bank_of_det, B, (int)n_banks, fault));
    PP_LAUNCHED();
    return PP_OK;
}
}  // namespace pp

extern "C" int pp_match_scores(const void* q_prep, const float* q_rnorm, const void* q_meta, const void* bank_prep,
                               const float* bank_rnorm, int64_t n_banks, const int32_t* bank_of_det, int B, int N, int H,
                               int W, int Kp, float* sim_avg, float* score_t2s, int32_t* idx_t2s, int32_t* idx_s2t,
                               uint8_t* mutual_nn, void* workspace, size_t workspace_bytes, int cluster, void* stream) {
    return pp::match_scores_impl(q_prep, q_rnorm, q_meta, bank_prep, bank_rnorm, n_banks, bank_of_det, B, N, H, W, Kp,
                                 sim_avg, score_t2s, idx_t2s, idx_s2t, mutual_nn, 0, nullptr, nullptr, workspace,
                                 workspace_bytes, cluster, stream);
}

extern "C" int pp_topk(const float* scores, int B, int N, int k, int64_t idx_offset, float* out_score,
                       int64_t* out_idx, void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (B == 0 || k == 0) return PP_OK;
    PP_CHECK_ARG(scores && out_score && out_idx, "pp_topk: null pointer");
    // torch.topk raises when k exceeds the dimension ("selected index k out of range")
    PP_CHECK_ARG(k >= 0 && k <= N, "pp_topk: selected index k out of range (k=%d, N=%d)", k, N);
    PP_CHECK_ARG(N <= 24000, "pp_topk: at most 24000 candidates per row (got %d)", N);
    if (B == 0 || k == 0) return PP_OK;
    const size_t smem = ((size_t)N + 8) * sizeof(unsigned long long);
    if (smem > 48 * 1024)
        PP_CUDA(cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_kernel<<<B, 256, smem, static_cast<cudaStream_t>(stream)>>>(scores, N, k, (long long)idx_offset, out_score,
                                                                    reinterpret_cast<long long*>(out_idx), nullptr, k);
    PP_LAUNCHED();
    return PP_OK;
}

extern "C" int pp_topk_pairs(const float* scores, int B, int N, int k, int64_t idx_offset, double* out_pairs,
                             void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (B == 0 || k == 0) return PP_OK;
    PP_CHECK_ARG(scores && out_pairs, "pp_topk_pairs: null pointer");
    PP_CHECK_ARG(k > 0 && k <= 256 && N >= 0 && N <= 24000, "pp_topk_pairs: bad sizes (k=%d, N=%d)", k, N);
    const int kl = k < N ? k : N;  // a shard may hold fewer than k views: the rest is padding
    const size_t smem = ((size_t)N + 8) * sizeof(unsigned long long);
    if (smem > 48 * 1024)
        PP_CUDA(cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_kernel<<<B, 256, smem, static_cast<cudaStream_t>(stream)>>>(scores, N, kl, (long long)idx_offset, nullptr,
                                                                    nullptr, out_pairs, k);
    PP_LAUNCHED();
    return PP_OK;
}

extern "C" int pp_topk_merge(const double* pairs, int R, int B, int k_in, int k, float* out_score, int64_t* out_idx,
                             void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (B == 0 || k == 0) return PP_OK;
    PP_CHECK_ARG(pairs && out_score && out_idx, "pp_topk_merge: null pointer");
    PP_CHECK_ARG(R > 0 && k_in > 0 && k > 0 && k <= R * k_in && R * k_in <= 4096, "pp_topk_merge: bad sizes");
    topk_merge_kernel<<<B, 32, (size_t)R * k_in * sizeof(unsigned long long), static_cast<cudaStream_t>(stream)>>>(
        pairs, R, B, k_in, k, out_score, reinterpret_cast<long long*>(out_idx));
    PP_LAUNCHED();
    return PP_OK;
}

extern "C" size_t pp_match_similarity_workspace(int B, int T) {
    // the similarity volume is written by the contraction's epilogue; nothing is staged any more (kept for the ABI)
    (void)B;
    (void)T;
    return 0;
}

extern "C" int pp_match_similarity(const void* q_prep, const float* q_rnorm, const void* s_prep, const float* s_rnorm,
                                   const float* src_mask, int B, int H, int W,
                                   int Kp, int Hm, int Wm, float* out, void* workspace, size_t workspace_bytes,
                                   int cluster, void* stream) {
    using namespace pp;
    (void)workspace;
    (void)workspace_bytes;
    if (int rc = require_sm100()) return rc;
    if (B == 0) return PP_OK;
    PP_CHECK_ARG(q_prep && q_rnorm && s_prep && s_rnorm && src_mask && out, "pp_match_similarity: null pointer");
    PP_CHECK_ARG(H == W, "pp_match_similarity: the reference asserts a square patch grid (H == W), got %dx%d", H, W);
    PP_CHECK_ARG(B >= 0 && H > 0 && Hm > 0 && Wm > 0, "pp_match_similarity: bad shape");
    const int T = H * W;
    // one "view" per detection: banks == detections, N = 1; norms, template mask, clamp and the (w h) layout in the epilogue
    return run_match_gemm(3, q_prep, s_prep, B, nullptr, B, 1, T, Kp, nullptr, nullptr, nullptr, q_rnorm, s_rnorm, nullptr, nullptr,
                          out, 1.0f, cluster & ~PP_MATCH_FAST_KEYS, static_cast<cudaStream_t>(stream), 0, nullptr, src_mask, Hm, Wm, H);
}

// matching_features_similarity as the reference calls it (utils/matching.py:6-26, from model/picopose.py:81): fp32 features
// in, TWO launches: one prologue for both operands, one contraction whose epilogue writes the finished volume.
extern "C" size_t pp_match_similarity_dense_workspace(int B, int C, int H, int W, int mode) {
    const int Kp = pp_match_kp(C, mode);
    if (B < 0 || H <= 0 || W <= 0 || Kp <= 0) return 0;
    const size_t T = (size_t)H * W;
    return pp::align_up(2 * (size_t)B * T * Kp * 2, 256) + pp::align_up(2 * (size_t)B * T * 4, 256);
}

extern "C" int pp_match_similarity_dense(const float* src_feat, const float* tar_feat, const float* src_mask, int B, int C,
                                         int H, int W, int Hm, int Wm, int mode, float* out, void* workspace,
                                         size_t workspace_bytes, int cluster, void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (B == 0) return PP_OK;
    PP_CHECK_ARG(src_feat && tar_feat && src_mask && out, "pp_match_similarity_dense: null pointer");
    PP_CHECK_ARG(H == W, "pp_match_similarity_dense: the reference asserts a square patch grid (H == W), got %dx%d", H, W);
    const int Kp = pp_match_kp(C, mode);
    PP_CHECK_ARG(Kp > 0 && Hm > 0 && Wm > 0, "pp_match_similarity_dense: bad feature dim %d / mode %d / mask", C, mode);
    const size_t need = pp_match_similarity_dense_workspace(B, C, H, W, mode);
    if (!workspace || workspace_bytes < need)
        return fail(PP_ERR_WORKSPACE, "pp_match_similarity_dense: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
    PP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "pp_match_similarity_dense: workspace must be 256-byte aligned");
    const int T = H * W;
    char* ws = static_cast<char*>(workspace);
    float* rn = reinterpret_cast<float*>(ws + align_up(2 * (size_t)B * T * Kp * 2, 256));
    // rows [0, B): query (tar) patches, rows [B, 2B): template (src) patches
    if (int rc = prepare_pair_impl(tar_feat, src_feat, B, C, T, mode, ws, rn, stream)) return rc;
    const char* s_prep = ws + (size_t)B * T * Kp * 2;
    return run_match_gemm(3, ws, s_prep, B, nullptr, B, 1, T, Kp, nullptr, nullptr, nullptr, rn, rn + (size_t)B * T, nullptr, nullptr,
                          out, 1.0f, cluster & ~PP_MATCH_FAST_KEYS, static_cast<cudaStream_t>(stream), 0, nullptr, src_mask, Hm, Wm, H);
}

namespace pp {
static int correlation_pyramid_impl(const void* f1_prep, const void* f2_prep, int N, int H, int W, int Kp, float scale,
                                    int num_levels, void* const* level_ptrs, int cluster, void* stream, bool tiled) {
    if (int rc = require_sm100()) return rc;
    if (N == 0) return PP_OK;
    PP_CHECK_ARG(f1_prep && f2_prep && level_ptrs, "pp_correlation_pyramid: null pointer");
    PP_CHECK_ARG(N > 0 && H > 0 && W > 0 && num_levels >= 1 && num_levels <= 8, "pp_correlation_pyramid: bad shape");
    for (int l = 0; l < num_levels; ++l) PP_CHECK_ARG(level_ptrs[l], "pp_correlation_pyramid: null level %d", l);
    PP_CHECK_ARG((H >> (num_levels - 1)) >= 1 && (W >> (num_levels - 1)) >= 1, "pp_correlation_pyramid: too many levels for %dx%d", H, W);
    if (tiled)
        PP_CHECK_ARG((W >> (num_levels - 1)) % 8 == 0 && (H >> (num_levels - 1)) % 4 == 0 && W % (1 << (num_levels - 1)) == 0 &&
                         H % (1 << (num_levels - 1)) == 0,
                     "pp_correlation_pyramid_tiled: every level needs W %% 8 == 0 and H %% 4 == 0 (%dx%d, %d levels)", H, W, num_levels);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int T = H * W;
    // all-pairs products <f1[:, q], f2[:, key]> * scale straight into level 0: (N*H*W, 1, H, W) is row-major [q][key]
    if (int rc = run_match_gemm(1, f1_prep, f2_prep, N, nullptr, N, 1, T, Kp, nullptr, nullptr, nullptr, nullptr, nullptr,
                                nullptr, nullptr, static_cast<float*>(level_ptrs[0]), scale, cluster, st, tiled ? W : 0))
        return rc;
    int h = H, w = W;
    for (int l = 1; l < num_levels; ++l) {
        const long long slices = (long long)N * T;
        const long long total = tiled ? slices * (h >> 1) * (w >> 4) : slices * (h >> 1) * (w >> 1);
        int grid = (int)((total + 255) / 256 < (long long)sm_count() * 32 ? (total + 255) / 256 : (long long)sm_count() * 32);
        if (tiled)
            avgpool2_tiled_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(level_ptrs[l - 1]), slices, h, w,
                                                        static_cast<float*>(level_ptrs[l]));
        else
            avgpool2_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(level_ptrs[l - 1]), slices, h, w,
                                                  static_cast<float*>(level_ptrs[l]));
        PP_LAUNCHED();
        h >>= 1;
        w >>= 1;
    }
    return PP_OK;
}
}  // namespace pp

extern "C" int pp_correlation_pyramid(const void* f1_prep, const void* f2_prep, int N, int H, int W, int Kp, float scale,
                                      int num_levels, void* const* level_ptrs, int cluster, void* stream) {
    return pp::correlation_pyramid_impl(f1_prep, f2_prep, N, H, W, Kp, scale, num_levels, level_ptrs, cluster, stream, false);
}

extern "C" int pp_correlation_pyramid_tiled(const void* f1_prep, const void* f2_prep, int N, int H, int W, int Kp, float scale,
                                            int num_levels, void* const* level_ptrs, int cluster, void* stream) {
    return pp::correlation_pyramid_impl(f1_prep, f2_prep, N, H, W, Kp, scale, num_levels, level_ptrs, cluster, stream, true);
}

extern "C" int pp_volume_retile(const float* in, float* out, int64_t slices, int h, int w, int to_tiled, void* stream) {
    using namespace pp;
    if (int rc = require_sm100()) return rc;
    if (slices == 0) return PP_OK;
    PP_CHECK_ARG(in && out && in != out, "pp_volume_retile: null pointer or in-place call");
    PP_CHECK_ARG(slices > 0 && h > 0 && w > 0 && w % 8 == 0 && h % 4 == 0, "pp_volume_retile: slices of H %% 4 == 0 rows and W %% 8 == 0 columns (got %dx%d)", h, w);
    PP_CHECK_ARG(((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "pp_volume_retile: 16-byte alignment");
    const long long total = (long long)slices * h * (w >> 3);
    int grid = (int)((total + 255) / 256 < (long long)sm_count() * 32 ? (total + 255) / 256 : (long long)sm_count() * 32);
    retile_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, (long long)slices, h, w, to_tiled ? 1 : 0, out);
    PP_LAUNCHED();
    return PP_OK;
}

// ---- one-call stage-1 ranking against prepared banks (keeps the host side to a single library call) ----------
namespace pp {
struct MatchWs {
    size_t q_prep, q_rnorm, q_meta, sim_avg, keys, total;
};
static MatchWs match_templates_layout(int B, int N, int T, int Kp) {
    MatchWs w;
    size_t off = 0;
    w.q_prep = off;  off += align_up((size_t)B * T * Kp * 2, 256);
    w.q_rnorm = off; off += align_up((size_t)B * T * 4, 256);
    w.q_meta = off;  off += align_up(pp_match_query_meta_bytes(B, T), 256);
    w.sim_avg = off; off += align_up((size_t)B * N * 4, 256);
    w.keys = off;    off += pp_match_scores_workspace(B, N, T);
    w.total = off;
    return w;
}
}  // namespace pp

extern "C" size_t pp_match_templates_workspace(int B, int N, int C, int H, int W, int mode) {
    const int Kp = pp_match_kp(C, mode);
    if (B < 0 || N < 0 || H <= 0 || W <= 0 || Kp <= 0) return 0;
    return pp::match_templates_layout(B, N, H * W, Kp).total;
}

extern "C" int pp_match_templates(const float* tar_feat, const float* tar_mask, const void* bank_prep,
                                  const float* bank_rnorm, int64_t n_banks, const int32_t* bank_of_det, int B, int N,
                                  int C, int H, int W, int Hm, int Wm, int mode, int k, float* out_score,
                                  int64_t* out_idx, float* sim_avg_out, void* workspace, size_t workspace_bytes,
                                  int cluster, void* stream) {
    using namespace pp;
    if (B == 0) return PP_OK;
    const int Kp = pp_match_kp(C, mode);
    PP_CHECK_ARG(Kp > 0, "pp_match_templates: bad feature dim %d / mode %d", C, mode);
    PP_CHECK_ARG(k >= 0 && k <= N, "pp_match_templates: selected index k out of range (k=%d, N=%d)", k, N);
    const int T = H * W;
    const MatchWs w = match_templates_layout(B, N, T, Kp);
    if (!workspace || workspace_bytes < w.total)
        return fail(PP_ERR_WORKSPACE, "pp_match_templates: workspace of %zu bytes needed, %zu given", w.total,
                    workspace_bytes);
    PP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "pp_match_templates: workspace must be 256-byte aligned");
    char* ws = static_cast<char*>(workspace);
    float* sim_avg = sim_avg_out ? sim_avg_out : reinterpret_cast<float*>(ws + w.sim_avg);
    // the query prologue also zeroes the contraction's key scratch: one launch instead of a kernel and a memset node
    if (int rc = prepare_query_impl(tar_feat, tar_mask, B, C, H, W, Hm, Wm, mode, ws + w.q_prep,
                                    reinterpret_cast<float*>(ws + w.q_rnorm), ws + w.q_meta, ws + w.keys,
                                    pp_match_scores_workspace(B, N, T), stream))
        return rc;
    const bool rank_it = k > 0 && out_score && out_idx;
    if (mode == PP_MODE_BF16 && !getenv("PICOPOSE_B200_EXACT_KEYS")) cluster |= PP_MATCH_FAST_KEYS;
    return match_scores_impl(ws + w.q_prep, reinterpret_cast<float*>(ws + w.q_rnorm), ws + w.q_meta, bank_prep, bank_rnorm,
                             n_banks, bank_of_det, B, N, H, W, Kp, sim_avg, nullptr, nullptr, nullptr, nullptr,
                             rank_it ? k : 0, out_score, out_idx, ws + w.keys, pp_match_scores_workspace(B, N, T), cluster,
                             stream, /*keys_cleared=*/true);
}

namespace pp {
// one forked stream + fork/join events per device, created on first use
struct ForkJoin {
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static thread_local ForkJoin g_fj[64];  // per host thread: concurrent callers never share the side stream or its events
static int fork_join_for_current_device(ForkJoin** out) {
    int dev = 0;
    PP_CUDA(cudaGetDevice(&dev));
    PP_CHECK_ARG(dev >= 0 && dev < 64, "device index %d out of range", dev);
    ForkJoin& f = g_fj[dev];
    if (!f.side) {
        PP_CUDA(cudaStreamCreateWithFlags(&f.side, cudaStreamNonBlocking));
        PP_CUDA(cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming));
        PP_CUDA(cudaEventCreateWithFlags(&f.join, cudaEventDisableTiming));
    }
    *out = &f;
    return PP_OK;
}
}  // namespace pp

extern "C" int pp_match_templates_dense(const float* src_feats, int64_t G, const float* tar_feat, const float* tar_mask,
                                        const int32_t* bank_of_det, int B, int N, int C, int H, int W, int Hm, int Wm,
                                        int mode, int k, void* bank_prep, float* bank_rnorm, float* out_score,
                                        int64_t* out_idx, float* sim_avg_out, void* workspace, size_t workspace_bytes,
                                        int cluster, void* stream) {
    using namespace pp;
    if (B == 0) return PP_OK;
    const int Kp = pp_match_kp(C, mode);
    PP_CHECK_ARG(Kp > 0, "pp_match_templates_dense: bad feature dim %d / mode %d", C, mode);
    PP_CHECK_ARG(src_feats && bank_prep && bank_rnorm && G > 0, "pp_match_templates_dense: null pointer / no banks");
    PP_CHECK_ARG(bank_of_det || G == B, "pp_match_templates_dense: %lld banks for %d detections need bank_of_det", (long long)G, B);
    PP_CHECK_ARG(k >= 0 && k <= N, "pp_match_templates_dense: selected index k out of range (k=%d, N=%d)", k, N);
    const int T = H * W;
    const MatchWs w = match_templates_layout(B, N, T, Kp);
    if (!workspace || workspace_bytes < w.total)
        return fail(PP_ERR_WORKSPACE, "pp_match_templates_dense: workspace of %zu bytes needed, %zu given", w.total,
                    workspace_bytes);
    PP_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "pp_match_templates_dense: workspace must be 256-byte aligned");
    char* ws = static_cast<char*>(workspace);
    float* sim_avg = sim_avg_out ? sim_avg_out : reinterpret_cast<float*>(ws + w.sim_avg);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ForkJoin* fj = nullptr;
    if (int rc = fork_join_for_current_device(&fj)) return rc;
    // fork: the (small, latency-bound) query prologue runs beside the (large, bandwidth-bound) bank prologue
    PP_CUDA(cudaEventRecord(fj->fork, st));
    PP_CUDA(cudaStreamWaitEvent(fj->side, fj->fork, 0));
    // (it also clears the contraction's key scratch)
    const int rc = prepare_query_impl(tar_feat, tar_mask, B, C, H, W, Hm, Wm, mode, ws + w.q_prep,
                                      reinterpret_cast<float*>(ws + w.q_rnorm), ws + w.q_meta, ws + w.keys,
                                      pp_match_scores_workspace(B, N, T), fj->side);
    const int rc2 = pp_match_prepare(src_feats, G * N, C, T, mode, 0, bank_prep, bank_rnorm, stream);
    // join unconditionally so that the side stream never outlives the call's ordering on `stream`
    const cudaError_t e1 = cudaEventRecord(fj->join, fj->side);
    const cudaError_t e2 = cudaStreamWaitEvent(st, fj->join, 0);
    if (rc) return rc;
    if (rc2) return rc2;
    PP_CUDA(e1);
    PP_CUDA(e2);
    const bool rank_it = k > 0 && out_score && out_idx;
    if (mode == PP_MODE_BF16 && !getenv("PICOPOSE_B200_EXACT_KEYS")) cluster |= PP_MATCH_FAST_KEYS;
    return match_scores_impl(ws + w.q_prep, reinterpret_cast<float*>(ws + w.q_rnorm), ws + w.q_meta, bank_prep, bank_rnorm,
                             G, bank_of_det, B, N, H, W, Kp, sim_avg, nullptr, nullptr, nullptr, nullptr,
                             rank_it ? k : 0, out_score, out_idx, ws + w.keys, pp_match_scores_workspace(B, N, T), cluster,
                             stream, /*keys_cleared=*/true);
}
