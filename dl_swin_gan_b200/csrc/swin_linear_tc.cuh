// Linear layers of the window-attention block (qkv :146, proj :168 of
// dl_cs/models/video_swin_transformer_mri_downsample.py) on the 5th-generation tensor cores:
//     Y[M, N] = X[M, K] * W[N, K]^T + bias[N]        fp16 operands, fp32 accumulation in TMEM.
//
// One persistent CTA per SM (128 threads).  W is converted to fp16 once per CTA and kept in shared memory in
// the canonical no-swizzle K-major UMMA layout ([k/8][n/8][8 rows][8 halves]: core matrices of 8 x 16 bytes,
// SBO = 128 B between row groups, LBO = (rows/8)*128 B between k chunks).  Per 128-row tile: the X rows are
// converted and laid out the same way, one elected thread issues the tcgen05.mma chain (K/16 steps per N half,
// N <= 256 per instruction), tcgen05.commit signals an mbarrier, and the four warps drain their 32 TMEM lanes
// with tcgen05.ld.
//
// Fusions around the GEMM (template modes):
//   A_LN_GATHER  rows are window tokens gathered from the feature volume [B,D,H,W,C] through the cyclic shift,
//                LayerNorm'ed on the fly (norm1 + pad + roll + window_partition, :218-236); padded tokens are zero rows
//   E_QKV16      the epilogue splits the 3C outputs into q (pre-scaled by scale*log2e), k and v, each [B_,heads,N,32] fp16
//                (v carries a column of ones at index head_dim), the layout the attention core stages with 16-byte copies
//   E_SCATTER    the epilogue writes each row to its place in [B,D,H,W,C] (window_reverse + roll back + crop, :238-248)
// CUDA-only (no host emulation: tcgen05 / TMEM have no software model here); validated on the B200 against
// torch in tests/test_swin_gpu.py.
#pragma once

#include "cuda_compat.h"
#include "swin_kernels.cuh"
#include "tc_common.cuh"

#if defined(__CUDACC__)
namespace dlsg {

constexpr int kLinThreads = 256;              // 8 warps: two threads per tile row, two warps per TMEM lane quarter
constexpr int kLinTileM = 128;

enum : int { A_PLAIN = 0, A_LN_GATHER = 1 };
enum : int { E_PLAIN = 0, E_QKV16 = 1, E_SCATTER = 2 };

struct LinearParams {
    const float* x;      // A_PLAIN: [M, K];  A_LN_GATHER: feature volume [B, D, H, W, C = K]
    const float* w;      // [N, K]
    const float* bias;   // [N] or NULL
    float* y;            // E_PLAIN: [M, N];  E_SCATTER: feature volume [B, D, H, W, N]
    const float* resid;  // E_SCATTER: optional [B, D, H, W, N] added to the scattered rows (the block's shortcut, :262)
    int M, N, K;
    int tmem_cols;       // power of two >= N, >= 32
    // A_LN_GATHER / E_SCATTER: rows are window tokens of this geometry (M = B*nW*tokens)
    SwinGeom g;
    const float* gamma; const float* beta; float eps;
    // E_QKV16
    half_t* q16; half_t* k16; half_t* vt16;     // q, k, v: each [B_,heads,Ntok,32]
    int heads, hd, ntok;
    float qscale;        // scale * log2(e), folded into q
};

__host__ __device__ inline size_t linear_a_region_bytes(int K) {      // A tile, reused as 4 x [32][33] fp32 transpose buffers
    const size_t a = (size_t)kLinTileM * K * 2, st = (size_t)(kLinThreads / 32) * 32 * 33 * 4;
    return ((a > st ? a : st) + 127) / 128 * 128;
}
// dynamic smem: W fp16 | A fp16 (reused by the epilogue) | bias fp32 [N] | row offsets int64 [128] | LN partials float2 [256] | ctrl
inline size_t linear_smem_bytes(int N, int K) {
    return (size_t)N * K * 2 + linear_a_region_bytes(K) + (size_t)N * 4 + (size_t)kLinTileM * 8 + (size_t)kLinThreads * 8 + 64;
}

template <int AMODE, int EMODE>
__global__ void __launch_bounds__(kLinThreads, (EMODE == E_QKV16) ? 1 : 2) swin_linear_tc_kernel(LinearParams p) {
    extern __shared__ __align__(128) unsigned char lin_smem[];
    const int N = p.N, K = p.K, KC = K / 8;
    half_t* sW = reinterpret_cast<half_t*>(lin_smem);
    half_t* sA = sW + (size_t)N * K;
    float* sBias = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(sA) + linear_a_region_bytes(K));
    long long* sRowOff = reinterpret_cast<long long*>(sBias + N);           // destination offset of each tile row, -1 = dropped
    float2* sPart = reinterpret_cast<float2*>(sRowOff + kLinTileM);         // per-thread LayerNorm partial sums
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(sPart + kLinThreads);
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned lboW = (unsigned)(N / 8) * 128u, lboA = (unsigned)(kLinTileM / 8) * 128u;

    // ---- one-time setup: TMEM allocation, mbarrier, weights and bias
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int i = tid; i < N * KC; i += kLinThreads) {          // item = (k chunk, row n): consecutive threads -> consecutive n
        const int kc = i / N, n = i - kc * N;
        const float4 a = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)n * K + 8 * kc));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)n * K + 8 * kc + 4));
        uint4 h;
        h.x = pack_h2(a.x, a.y); h.y = pack_h2(a.z, a.w); h.z = pack_h2(b.x, b.y); h.w = pack_h2(b.z, b.w);
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(sW) + (size_t)kc * lboW + (size_t)(n >> 3) * 128 + (n & 7) * 16) = h;
    }
    for (int i = tid; i < N; i += kLinThreads) sBias[i] = p.bias ? __ldg(p.bias + i) : 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem_base = *tmem_slot;

    const int ntok = swin_tokens(p.g), nwin = swin_windows(p.g);   // only meaningful for the gather / scatter modes
    const int ntiles = (p.M + kLinTileM - 1) / kLinTileM;
    unsigned phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- A tile: two threads per row (each one half of K); fp32 -> fp16, canonical layout (rows beyond M are zero)
        const int trow = tid & (kLinTileM - 1), khalf = tid >> 7;
        const int row = tile * kLinTileM + trow;
        const int kc0 = khalf * (KC / 2), kc1 = khalf ? KC : KC / 2;
        const float* xr = nullptr;
        if (AMODE == A_PLAIN) {
            if (row < p.M) xr = p.x + (size_t)row * K;
        }
        if (AMODE == A_LN_GATHER || EMODE == E_SCATTER) {
            long long off = -1;
            if (row < p.M) {
                const int tok = row % ntok;
                const int bw = row / ntok;
                const int win = bw % nwin, b = bw / nwin;
                int d, h, w;
                if (swin_token_coords(p.g, win, tok, d, h, w))
                    off = (((long long)b * p.g.D + d) * p.g.H + h) * p.g.W + w;          // token index in [B, D, H, W]
            }
            if (EMODE == E_SCATTER && khalf == 0) sRowOff[trow] = off;
            if (AMODE == A_LN_GATHER && off >= 0) xr = p.x + (size_t)off * K;
        }
        float mu = 0.f, rstd = 1.f;
        if (AMODE == A_LN_GATHER) {
            // one pass over this thread's half row: shifted sums (shift = first element of the row), combined through smem
            float s1 = 0.f, s2 = 0.f, x0 = 0.f;
            if (xr) {
                x0 = __ldg(xr);
                for (int kc = kc0; kc < kc1; ++kc) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(xr + 8 * kc));
                    const float4 b = __ldg(reinterpret_cast<const float4*>(xr + 8 * kc + 4));
                    const float d0 = a.x - x0, d1 = a.y - x0, d2 = a.z - x0, d3 = a.w - x0;
                    const float d4 = b.x - x0, d5 = b.y - x0, d6 = b.z - x0, d7 = b.w - x0;
                    s1 += ((d0 + d1) + (d2 + d3)) + ((d4 + d5) + (d6 + d7));
                    s2 += ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3)) + ((d4 * d4 + d5 * d5) + (d6 * d6 + d7 * d7));
                }
            }
            sPart[tid] = make_float2(s1, s2);
            __syncthreads();
            const float2 o2 = sPart[tid ^ kLinTileM];
            s1 += o2.x; s2 += o2.y;
            const float m0 = s1 / (float)K;
            mu = x0 + m0;
            rstd = 1.0f / sqrtf(fmaxf(s2 / (float)K - m0 * m0, 0.f) + p.eps);
        }
        for (int kc = kc0; kc < kc1; ++kc) {
            uint4 h = make_uint4(0u, 0u, 0u, 0u);
            if (xr) {
                float4 a = __ldg(reinterpret_cast<const float4*>(xr + 8 * kc));
                float4 b = __ldg(reinterpret_cast<const float4*>(xr + 8 * kc + 4));
                if (AMODE == A_LN_GATHER) {
                    const float4 g0 = p.gamma ? __ldg(reinterpret_cast<const float4*>(p.gamma + 8 * kc)) : make_float4(1.f, 1.f, 1.f, 1.f);
                    const float4 g1 = p.gamma ? __ldg(reinterpret_cast<const float4*>(p.gamma + 8 * kc + 4)) : make_float4(1.f, 1.f, 1.f, 1.f);
                    const float4 b0 = p.beta ? __ldg(reinterpret_cast<const float4*>(p.beta + 8 * kc)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 b1 = p.beta ? __ldg(reinterpret_cast<const float4*>(p.beta + 8 * kc + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    a.x = fmaf((a.x - mu) * rstd, g0.x, b0.x); a.y = fmaf((a.y - mu) * rstd, g0.y, b0.y);
                    a.z = fmaf((a.z - mu) * rstd, g0.z, b0.z); a.w = fmaf((a.w - mu) * rstd, g0.w, b0.w);
                    b.x = fmaf((b.x - mu) * rstd, g1.x, b1.x); b.y = fmaf((b.y - mu) * rstd, g1.y, b1.y);
                    b.z = fmaf((b.z - mu) * rstd, g1.z, b1.z); b.w = fmaf((b.w - mu) * rstd, g1.w, b1.w);
                }
                h.x = pack_h2(a.x, a.y); h.y = pack_h2(a.z, a.w); h.z = pack_h2(b.x, b.y); h.w = pack_h2(b.z, b.w);
            }
            *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(sA) + (size_t)kc * lboA + (size_t)(trow >> 3) * 128 + (trow & 7) * 16) = h;
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // generic-proxy smem writes -> visible to the tensor core
        __syncthreads();
        // ---- MMA chain: D[128, N] = A[128, K] * W[N, K]^T, N split into halves of <= 256 columns
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const int nh = (N > 256) ? N / 2 : N;
            const unsigned idesc = umma_idesc_f16(kLinTileM, nh);
            for (int n0 = 0; n0 < N; n0 += nh) {
                for (int ks = 0; ks < K / 16; ++ks) {
                    const unsigned long long ad = umma_desc(smem_u32(sA) + (unsigned)(2 * ks) * lboA, lboA, 128u);
                    const unsigned long long bd = umma_desc(smem_u32(sW) + (unsigned)(2 * ks) * lboW + (unsigned)(n0 / 8) * 128u, lboW, 128u);
                    umma_f16(tmem_base + (unsigned)n0, ad, bd, idesc, ks > 0 ? 1u : 0u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(mbar)) : "memory");
        }
        // ---- wait for the accumulator (bounded spin: a broken pipeline traps instead of hanging the GPU)
        {
            unsigned done = 0;
            for (long long spin = 0; !done; ++spin) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(done) : "r"(smem_u32(mbar)), "r"(phase) : "memory");
                if (spin > (1ll << 26)) __trap();
            }
            phase ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const int quarter = warp & 3, wsub = warp >> 2;                        // warps q and q+4 share TMEM lanes 32q..32q+31
        const unsigned tlane = tmem_base + ((unsigned)(quarter * 32) << 16);
        if (EMODE == E_QKV16) {
            // thread = token row; per (q|k|v, head) group of hd columns: fp16 rows of 32 for q / k, columns of v^T
            const int C = p.heads * p.hd;
            const int erow = tile * kLinTileM + quarter * 32 + lane;
            const bool live = erow < p.M;                     // tcgen05.ld is warp-collective: only the stores are predicated
            const int tok = live ? erow % p.ntok : 0, bw = live ? erow / p.ntok : 0;
            for (int which = 0; which < 3; ++which) {
                for (int head = wsub; head < p.heads; head += 2) {
                    const int c0 = which * C + head * p.hd;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        if (4 * j4 < p.hd) tmem_ld4(tlane + (unsigned)(c0 + 4 * j4), v + 4 * j4);
                    tmem_ld_wait();
                    const float sc = (which == 0) ? p.qscale : 1.0f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = (j < p.hd) ? (v[j] + sBias[c0 + (j < p.hd ? j : 0)]) * sc : 0.f;
                    if (which == 2 && p.hd < 32) v[p.hd] = 1.0f;          // column of ones: the attention core reads the softmax row sums off the PV MMA
                    const size_t bh = (size_t)bw * p.heads + head;
                    if (live) {
                        half_t* base16 = (which == 0) ? p.q16 : (which == 1 ? p.k16 : p.vt16);
                        uint4* dst = reinterpret_cast<uint4*>(base16 + (bh * p.ntok + tok) * 32);
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4)
                            dst[q4] = make_uint4(pack_h2(v[8 * q4], v[8 * q4 + 1]), pack_h2(v[8 * q4 + 2], v[8 * q4 + 3]),
                                                 pack_h2(v[8 * q4 + 4], v[8 * q4 + 5]), pack_h2(v[8 * q4 + 6], v[8 * q4 + 7]));
                    }
                }
            }
        } else {
            // 32 columns at a time through a shared-memory transpose so that rows are written contiguously
            float* stage = reinterpret_cast<float*>(sA) + warp * (32 * 33);
            const int row0 = tile * kLinTileM + quarter * 32;
            for (int c0 = wsub * 32; c0 < N; c0 += 64) {
                float v[32];
                tmem_ld32(tlane + (unsigned)c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = v[j] + sBias[c0 + j];
                __syncwarp();
                if (EMODE == E_SCATTER && p.resid) {
                    // shortcut rows of this 32 x 32 block: all loads in flight together, then add and store
                    float xr[32];
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const long long off = sRowOff[quarter * 32 + r];
                        xr[r] = (off >= 0) ? __ldg(p.resid + (size_t)off * N + c0 + lane) : 0.f;
                    }
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const long long off = sRowOff[quarter * 32 + r];
                        if (off >= 0) p.y[(size_t)off * N + c0 + lane] = stage[r * 33 + lane] + xr[r];
                    }
                } else
                for (int r = 0; r < 32; ++r) {
                    if (EMODE == E_SCATTER) {
                        const long long off = sRowOff[quarter * 32 + r];
                        if (off >= 0) p.y[(size_t)off * N + c0 + lane] = stage[r * 33 + lane];
                    } else {
                        if (row0 + r < p.M) p.y[(size_t)(row0 + r) * N + c0 + lane] = stage[r * 33 + lane];
                    }
                }
                __syncwarp();
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();                                      // TMEM drained and the A / stage region free for the next tile
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

}  // namespace dlsg
#endif
