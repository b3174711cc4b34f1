// Shared by the stage-3 lookup kernels (corr_lookup.cu, corr_lookup_tma.cu): launch parameters and the tap arithmetic.
#pragma once

#include "pp_common.cuh"

namespace pp {

constexpr int LOOKUP_MAX_LEVELS = 8;
constexpr int LOOKUP_MAX_RADIUS = 16;

struct LookupParams {
    const float* vol[LOOKUP_MAX_LEVELS];
    int vh[LOOKUP_MAX_LEVELS];
    int vw[LOOKUP_MAX_LEVELS];
    int vec_ok[LOOKUP_MAX_LEVELS];  // 16-byte path usable (width % 4 == 0, base aligned)
    int tiled;                      // slices stored as 4 x 8 tiles of 32 floats (one 128-byte line each), tiles row-major
    int L;
    const float* flow;
    float* out;
    int B, H, W, HW;
    int radius;
    int groups_per_b;  // ceil(HW / 32)
    int total_groups;
    int nr_max, nv_max;  // generic kernel: staged rows / 16-byte pieces per row
    int qstride;         // generic kernel: words between two queries' staging areas
};

// pixel coordinate -> (floor index, weight of the upper tap), replicating
//   g = p*2/max(size-1,1) - 1 ; i = ((g+1)/2)*(size-1)
__device__ __forceinline__ void axis_tap(float p, int size, int& i0, float& w1) {
    float den = (float)(size > 1 ? size - 1 : 1);
    float g = __fsub_rn(__fdiv_rn(__fmul_rn(p, 2.0f), den), 1.0f);
    float i = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
    float f = floorf(i);
    w1 = __fsub_rn(i, f);
    // clamp so that far-away (or non-finite) windows stay inside the staged footprint; every tap of
    // a clamped index is out of bounds and contributes zero, exactly as zero padding does.
    f = fminf(fmaxf(f, -2.0f), (float)size);
    i0 = (f == f) ? (int)f : -2;
    // a non-finite coordinate leaves w1 = NaN: every sample that uses this tap is NaN, as in F.grid_sample
}


// Same, plus the floor BEFORE the clamp (limited to +-1e6, NaN -> a sentinel): lets a caller recognise the regular case
// "tap j sits exactly j rows below tap 0" even when some taps were clamped at the border.
__device__ __forceinline__ void axis_tap_raw(float p, int size, int& i0, float& w1, int& raw) {
    float den = (float)(size > 1 ? size - 1 : 1);
    float g = __fsub_rn(__fdiv_rn(__fmul_rn(p, 2.0f), den), 1.0f);
    float i = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
    float f = floorf(i);
    w1 = __fsub_rn(i, f);
    raw = (f == f) ? (int)fminf(fmaxf(f, -1.0e6f), 1.0e6f) : -2000000;
    f = fminf(fmaxf(f, -2.0f), (float)size);
    i0 = (f == f) ? (int)f : -2;
    // a non-finite coordinate leaves w1 = NaN: every sample that uses this tap is NaN, as in F.grid_sample
}

// corr_lookup_tma.cu: tiled volumes fetched with bulk copies (radius 1..8); *handled = false leaves it to the banded kernel
int launch_lookup_tma(const LookupParams& p, cudaStream_t st, bool* handled);

}  // namespace pp
