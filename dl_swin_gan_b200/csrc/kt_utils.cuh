// k-t pre-processing steps either side of the SENSE operator, dl_cs/mri/utils.py:
//   fftmod :7-19, get_mask :69-79, time_average :29-34, sliding_window :37-49
// as used by DataTransform (scripts/reconstruct.py:132-171) and CinePreprocess (dl_cs/data/preprocess.py:140-164).
// All memory-bound one-pass kernels over [outer, T, inner] complex64 data (inner = Y*X contiguous).
#pragma once

#include "cuda_compat.h"

namespace dlsg {

// in place: out[..., y, x] *= -(-1)^(x + y)      (the three sign flips of fftmod combined)
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(256)
fftmod_kernel(float2* __restrict__ data, long long n, int Y, int X) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = (int)(i % X), y = (int)((i / X) % Y);
    if (((x + y) & 1) == 0) {
        float2 v = data[i];
        data[i] = make_float2(-v.x, -v.y);
    }
}

// mask[i] = |data[i]| > eps ? 1 : 0
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(256)
get_mask_kernel(const float2* __restrict__ data, float* __restrict__ mask, long long n, float eps) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 v = __ldg(data + i);
    mask[i] = (sqrtf(v.x * v.x + v.y * v.y) > eps) ? 1.0f : 0.0f;
}

// circular sliding-window time average along T for data [outer, T, inner]:
//   out[o, t, p] = sum_{j<window} d[o, (t - window/2 + j) mod T, p] / (count(|d| > mask_eps) + eps)
// window == T and t_out == 1 gives time_average(keepdim=True).  One thread per (o, p): running sums over the ring.
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(256)
sliding_window_kernel(const float2* __restrict__ data, float2* __restrict__ out, long long outer, int T, long long inner,
                      int window, int t_out, float mask_eps, float eps) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= outer * inner) return;
    const long long o = idx / inner, p = idx - o * inner;
    const float2* d = data + o * T * inner + p;
    float2* q = out + o * (long long)t_out * inner + p;
    const int half = window / 2;
    for (int t = 0; t < t_out; ++t) {
        float sx = 0.f, sy = 0.f, cnt = 0.f;
        for (int j = 0; j < window; ++j) {
            int tt = (t - half + j) % T;
            if (tt < 0) tt += T;
            const float2 v = __ldg(d + (long long)tt * inner);
            sx += v.x; sy += v.y;
            cnt += (sqrtf(v.x * v.x + v.y * v.y) > mask_eps) ? 1.0f : 0.0f;
        }
        const float inv = 1.0f / (cnt + eps);
        q[(long long)t * inner] = make_float2(sx * inv, sy * inv);
    }
}

}  // namespace dlsg
