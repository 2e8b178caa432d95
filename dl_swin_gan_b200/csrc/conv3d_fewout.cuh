// Last layer of the denoisers: nn.Conv3d(Cin -> 2 or 4, kernel 3, 'same') of `final_layer` (dl_cs/models/swin3D.py:395,
// resnet3d.py:270): the output is the re / im channels of the E map sets.  With 2 - 4 output channels there is no GEMM to
// speak of (N = 4): plain FP32 FMAs on the CUDA cores (packed FFMA2, exact fp32 - no operand splitting needed), one pass over
// the input.  Thread layout as the tensor-core kernel's producers: 8 lanes cover the 128 bytes of one voxel's 32-channel chunk
// (coalesced), each thread keeps the partial sums of its channel quad for 8 voxels and all outputs; the 8 lanes of a voxel are
// reduced with shuffles at the end.  The weights sit in shared memory as [tap][ci][co].
#pragma once

#include "cuda_compat.h"
#include "fft_packed.cuh"

#if defined(__CUDACC__)
namespace dlsg {

struct FewOutParams {
    const float* x;        // [NB, T, Y, X, Cin]
    const float* w;        // [27][Cin][CO]
    const float* bias;     // [CO] or NULL
    float* out;            // [NB, T, Y, X, CO]
    int NB, T, Y, X, Cin;
    long long NV;
    int relu_in;
};

constexpr int kFewThreads = 256;

template <int CO>
__global__ void __launch_bounds__(kFewThreads, 2) conv3d_k3_fewout_kernel(FewOutParams p) {
    extern __shared__ __align__(16) unsigned char few_smem[];
    float* sw = reinterpret_cast<float*>(few_smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nw = 27 * p.Cin * CO;
    int* sTapOff = reinterpret_cast<int*>(sw + nw);                          // voxel offset of each tap
    unsigned char* sOk = reinterpret_cast<unsigned char*>(sTapOff + 28);      // [27][kFewThreads]
    for (int i = tid; i < nw / 4; i += kFewThreads) reinterpret_cast<float4*>(sw)[i] = __ldg(reinterpret_cast<const float4*>(p.w) + i);
    if (tid < 27) sTapOff[tid] = ((tid / 9 - 1) * p.Y + ((tid / 3) % 3 - 1)) * p.X + (tid % 3 - 1);
    __syncthreads();
    const int KCH = p.Cin / 32;
    const int k4 = lane & 7, rsub = lane >> 3;
    const long long YX = (long long)p.Y * p.X;
    const long long ntiles = (p.NV + kFewThreads - 1) / kFewThreads;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long v0 = tile * kFewThreads + warp * 32 + rsub;          // this thread's voxels: v0 + 4 i
        int coord[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const long long vi = v0 + 4 * i;
            if (vi < p.NV) {
                const long long r1 = vi % YX;
                coord[i] = (int)((vi / YX) % p.T) << 20 | (int)(r1 / p.X) << 10 | (int)(r1 % p.X);
            } else coord[i] = -1;
        }
        float2 acc[8][CO / 2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < CO / 2; ++c) acc[i][c] = make_float2(0.f, 0.f);
        // in-volume flags of this thread's 8 rows for each tap, kept in shared memory (one byte per tap and thread)
        for (int tap = 0; tap < 27; ++tap) {
            const int dt = tap / 9 - 1, dy = (tap / 3) % 3 - 1, dx = tap % 3 - 1;
            unsigned m = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = coord[i];
                const bool ok = c >= 0 && (unsigned)((c >> 20) + dt) < (unsigned)p.T && (unsigned)(((c >> 10) & 1023) + dy) < (unsigned)p.Y &&
                                (unsigned)((c & 1023) + dx) < (unsigned)p.X;
                m |= (ok ? 1u : 0u) << i;
            }
            sOk[tap * kFewThreads + tid] = (unsigned char)m;
        }
        // channel chunk outermost, taps inside: the three dx taps of a (dt, dy) pair read the same lines one voxel apart, so two of the
        // three come from L1 (with the taps outermost each of the 27 passes over the input came from L2: ~10 TB/s, L2-bound)
        const unsigned rbase = (unsigned)v0 * (unsigned)p.Cin + 4u * k4;
        for (int ch = 0; ch < KCH; ++ch) {
#pragma unroll 1
            for (int tap = 0; tap < 27; ++tap) {
                const unsigned okmask = sOk[tap * kFewThreads + tid];
                const unsigned soff = rbase + (unsigned)(sTapOff[tap] * p.Cin + ch * 32);      // 32-bit on purpose: a row's offset wraps back into range
                const float* wt = sw + ((size_t)tap * p.Cin + ch * 32 + 4 * k4) * CO;
                float4 a[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    a[i] = ((okmask >> i) & 1u) ? __ldg(reinterpret_cast<const float4*>(p.x + (soff + (unsigned)(4 * i) * (unsigned)p.Cin)))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.relu_in) { a[i].x = fmaxf(a[i].x, 0.f); a[i].y = fmaxf(a[i].y, 0.f); a[i].z = fmaxf(a[i].z, 0.f); a[i].w = fmaxf(a[i].w, 0.f); }
                }
                float2 wq[4][CO / 2];                                        // weights of this thread's 4 input channels
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int c = 0; c < CO / 2; ++c) wq[j][c] = *reinterpret_cast<const float2*>(wt + (size_t)j * CO + 2 * c);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int c = 0; c < CO / 2; ++c) {
                        acc[i][c] = p_fma(p_bc(a[i].x), wq[0][c], acc[i][c]);
                        acc[i][c] = p_fma(p_bc(a[i].y), wq[1][c], acc[i][c]);
                        acc[i][c] = p_fma(p_bc(a[i].z), wq[2][c], acc[i][c]);
                        acc[i][c] = p_fma(p_bc(a[i].w), wq[3][c], acc[i][c]);
                    }
            }
        }
        // sum over the 8 channel-quad lanes of each voxel; lane k4 == 0 stores
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < CO / 2; ++c) {
                float2 s = acc[i][c];
#pragma unroll
                for (int m = 1; m < 8; m <<= 1) {
                    s.x += __shfl_xor_sync(0xffffffffu, s.x, m);
                    s.y += __shfl_xor_sync(0xffffffffu, s.y, m);
                }
                acc[i][c] = s;
            }
        if (k4 == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long vi = v0 + 4 * i;
                if (vi < p.NV) {
#pragma unroll
                    for (int c = 0; c < CO / 2; ++c) {
                        float2 r = acc[i][c];
                        if (p.bias) { r.x += __ldg(p.bias + 2 * c); r.y += __ldg(p.bias + 2 * c + 1); }
                        *reinterpret_cast<float2*>(p.out + (size_t)vi * CO + 2 * c) = r;
                    }
                }
            }
        }
    }
}

// Second form of the last layer: z[v, tap*CO + co] = sum_ci w[co, ci, tap] * x[v, ci] is a pointwise 160 -> 27*CO linear (one pass over the
// input on the tensor-core kernel, conv3d_tc.cuh with a single tap), and the convolution is the sum of z over the 27 shifted voxels:
//     out[v, co] = bias[co] + sum_{tap: v + tap inside the volume} z[v + tap, tap*CO + co]
// Every element of z is read once.  Thread = voxel; ZC = channel stride of z (>= 27*CO, padded to the GEMM's tile).
template <int CO>
__global__ void __launch_bounds__(256) conv3d_tapsum_kernel(const float* __restrict__ z, const float* __restrict__ bias, float* __restrict__ out,
                                                           int T, int Y, int X, int ZC, long long NV, int tlo, int tn) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;        // row of the frame window tlo .. tlo + tn - 1 of every sample
    if (w >= NV) return;
    const long long YX = (long long)Y * X;
    const long long v = w + ((w / (YX * tn)) * (T - tn) + tlo) * YX;              // voxel of the volume
    const long long r1 = v % YX;
    const int vx = (int)(r1 % X), vy = (int)(r1 / X), vt = (int)((v / YX) % T);
    float acc[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[c] = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
    for (int tap = 0; tap < 27; ++tap) {
        const int dt = tap / 9 - 1, dy = (tap / 3) % 3 - 1, dx = tap % 3 - 1;
        if ((unsigned)(vt + dt) < (unsigned)T && (unsigned)(vy + dy) < (unsigned)Y && (unsigned)(vx + dx) < (unsigned)X) {
            const float* q = z + (size_t)(v + (long long)dt * YX + (long long)dy * X + dx) * ZC + tap * CO;
            if (CO == 4) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(q));
                acc[0] += a.x; acc[1] += a.y; acc[2 % CO] += a.z; acc[3 % CO] += a.w;
            } else {
                const float2 a = __ldg(reinterpret_cast<const float2*>(q));
                acc[0] += a.x; acc[1] += a.y;
            }
        }
    }
    if (CO == 4) *reinterpret_cast<float4*>(out + (size_t)v * 4) = make_float4(acc[0], acc[1], acc[2 % CO], acc[3 % CO]);
    else *reinterpret_cast<float2*>(out + (size_t)v * 2) = make_float2(acc[0], acc[1]);
}

}  // namespace dlsg
#endif
