// MLP half of the Swin block (forward_part2, dl_cs/models/video_swin_transformer_mri_downsample.py:251-252 with
// Mlp :20-38 and norm2 :208):      out = [x +] fc2( GELU( fc1( LayerNorm(x) ) ) )
// on the 5th-generation tensor cores (fp16 operands, fp32 accumulation in TMEM), as two kernels:
//
//   swin_mlp_fc1_kernel   per 128-token tile: LayerNorm(x) -> fp16 A tile (once); the hidden layer is produced in
//                         chunks of 160 columns: the chunk of W1 (fp16, [H, C] row-major) streams in with cp.async
//                         while the previous chunk's epilogue (bias, exact erf GELU, fp16 store) runs
//   swin_mlp_fc2_kernel   per tile: K = H is walked in chunks of 160: fp16 hidden rows and the W2 chunk ([C, H] row-major)
//                         land with cp.async directly in the canonical UMMA layout, tcgen05.mma accumulates over the
//                         chunks in TMEM; epilogue adds bias (and the residual x) and writes fp32 rows
//
// Both operands use the no-swizzle K-major core-matrix layout of swin_linear_tc.cuh: a 16-byte cp.async is exactly one
// core-matrix row (8 halves of one row), so no conversion pass is needed for fp16 sources.  The weights are converted
// to fp16 once by the host module (they are constants at inference).  C = 160 is the reference's embedding width;
// any C, H that are multiples of 32 with C <= 256 and H % C == 0 are accepted.
// CUDA-only (tcgen05 / TMEM have no host model); validated on the B200 against torch (tests/test_swin_gpu.py).
#pragma once

#include "swin_linear_tc.cuh"

#if defined(__CUDACC__)
namespace dlsg {

struct MlpParams {
    const float* x;        // [M, C] fp32 (tokens, channel fastest)
    const float* gamma; const float* beta; float eps;      // norm2
    const half_t* w1h;     // [H, C] fp16
    const float* b1;       // [H]
    const half_t* w2h;     // [C, H] fp16
    const float* b2;       // [C]
    half_t* hidden;        // [M, H] fp16 workspace
    float* y;              // [M, C]
    int M, C, H;
    int residual;          // 1: y = x + mlp(LN(x))  (the block's `x + forward_part2(x)`, :271);  0: y = mlp(LN(x))
    int tmem_cols;
};

// canonical layout helpers: element (row, k) of a [rows x K] K-major operand, in bytes
__device__ __forceinline__ unsigned canon_off(int row, int k8, unsigned lbo) {
    return (unsigned)k8 * lbo + (unsigned)(row >> 3) * 128u + (unsigned)(row & 7) * 16u;
}

inline size_t mlp_smem_bytes(int C, int H) {
    const size_t a = linear_a_region_bytes(C);                 // A tile (fc1: LN(x); fc2: hidden chunk), reused as transpose stage
    const size_t w = (size_t)C * C * 2;                        // one weight chunk: C rows x C k
    const size_t bias = (size_t)(H > C ? H : C) * 4;
    return a + w + bias + (size_t)kLinThreads * 8 + 64;
}

__device__ __forceinline__ void mbar_wait(unsigned long long* mbar, unsigned& phase) {
    unsigned done = 0;
    for (long long spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smem_u32(mbar)), "r"(phase) : "memory");
        if (spin > (1ll << 26)) __trap();          // a broken pipeline traps instead of hanging the GPU
    }
    phase ^= 1u;
}

// copy a [rows x C] fp16 chunk (row stride ld halves, starting at column k0) into the canonical layout; rows >= nrows are zero
__device__ __forceinline__ void stage_chunk_f16(unsigned char* dst, unsigned lbo, const half_t* src, size_t ld, int k0,
                                                int rows, int nrows, int C, int tid) {
    const int kc = C / 8;
    for (int i = tid; i < rows * kc; i += kLinThreads) {
        const int k8 = i / rows, r = i - k8 * rows;            // consecutive threads -> consecutive rows of one k chunk
        unsigned char* d = dst + canon_off(r, k8, lbo);
        if (r < nrows) cp_async16(d, src + (size_t)r * ld + k0 + 8 * k8);
        else *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);
    }
}

template <int PHASE>    // 1: hidden = GELU(LN(x) W1^T + b1);  2: y = [x +] hidden W2^T + b2
__global__ void __launch_bounds__(kLinThreads, 2) swin_mlp_kernel(MlpParams p) {
    extern __shared__ __align__(128) unsigned char mlp_smem[];
    const int C = p.C, H = p.H, KC = C / 8;
    unsigned char* sA = mlp_smem;
    unsigned char* sW = sA + linear_a_region_bytes(C);
    float* sBias = reinterpret_cast<float*>(sW + (size_t)C * C * 2);
    float2* sPart = reinterpret_cast<float2*>(sBias + (H > C ? H : C));
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(sPart + kLinThreads);
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned lboA = (unsigned)(kLinTileM / 8) * 128u, lboW = (unsigned)(C / 8) * 128u;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(smem_u32(mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    const int nb = (PHASE == 1) ? H : C;
    const float* bsrc = (PHASE == 1) ? p.b1 : p.b2;
    for (int i = tid; i < nb; i += kLinThreads) sBias[i] = bsrc ? __ldg(bsrc + i) : 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem_base = *tmem_slot;
    const unsigned idesc = umma_idesc_f16(kLinTileM, C);
    const int quarter = warp & 3, wsub = warp >> 2;            // warps q and q+4 share TMEM lanes 32q .. 32q+31
    const unsigned tlane = tmem_base + ((unsigned)(quarter * 32) << 16);
    const int nchunks = H / C;

    const int ntiles = (p.M + kLinTileM - 1) / kLinTileM;
    unsigned phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tile * kLinTileM;
        const int nrows = min(kLinTileM, p.M - row0);
        if (PHASE == 1) {
            // ---- A tile = LayerNorm(x rows), two threads per row (each one half of C), one pass with shifted sums
            stage_chunk_f16(sW, lboW, p.w1h, (size_t)C, 0, C, C, C, tid);          // first W1 chunk: rows 0 .. C-1 of [H, C]
            cp_async_commit();
            const int trow = tid & (kLinTileM - 1), khalf = tid >> 7;
            const int kc0 = khalf * (KC / 2), kc1 = khalf ? KC : KC / 2;
            const float* xr = (trow < nrows) ? p.x + (size_t)(row0 + trow) * C : nullptr;
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
            const float m0 = s1 / (float)C;
            const float mu = x0 + m0;
            const float rstd = 1.0f / sqrtf(fmaxf(s2 / (float)C - m0 * m0, 0.f) + p.eps);
            for (int kc = kc0; kc < kc1; ++kc) {
                uint4 h = make_uint4(0u, 0u, 0u, 0u);
                if (xr) {
                    float4 a = __ldg(reinterpret_cast<const float4*>(xr + 8 * kc));
                    float4 b = __ldg(reinterpret_cast<const float4*>(xr + 8 * kc + 4));
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + 8 * kc)), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + 8 * kc + 4));
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + 8 * kc)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta + 8 * kc + 4));
                    a.x = fmaf((a.x - mu) * rstd, g0.x, b0.x); a.y = fmaf((a.y - mu) * rstd, g0.y, b0.y);
                    a.z = fmaf((a.z - mu) * rstd, g0.z, b0.z); a.w = fmaf((a.w - mu) * rstd, g0.w, b0.w);
                    b.x = fmaf((b.x - mu) * rstd, g1.x, b1.x); b.y = fmaf((b.y - mu) * rstd, g1.y, b1.y);
                    b.z = fmaf((b.z - mu) * rstd, g1.z, b1.z); b.w = fmaf((b.w - mu) * rstd, g1.w, b1.w);
                    h.x = pack_h2(a.x, a.y); h.y = pack_h2(a.z, a.w); h.z = pack_h2(b.x, b.y); h.w = pack_h2(b.z, b.w);
                }
                *reinterpret_cast<uint4*>(sA + canon_off(trow, kc, lboA)) = h;
            }
            for (int nc = 0; nc < nchunks; ++nc) {
                cp_async_wait_group0();
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                __syncthreads();                               // A tile (first chunk) and this W1 chunk are in place
                if (tid == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    for (int ks = 0; ks < C / 16; ++ks)
                        umma_f16(tmem_base, umma_desc(smem_u32(sA) + (unsigned)(2 * ks) * lboA, lboA, 128u),
                                 umma_desc(smem_u32(sW) + (unsigned)(2 * ks) * lboW, lboW, 128u), idesc, ks > 0 ? 1u : 0u);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(mbar)) : "memory");
                }
                mbar_wait(mbar, phase);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                if (nc + 1 < nchunks)                          // the MMA has consumed sW: the next chunk lands during the epilogue
                    stage_chunk_f16(sW, lboW, p.w1h + (size_t)(nc + 1) * C * C, (size_t)C, 0, C, C, C, tid);
                cp_async_commit();
                // epilogue: thread = token row; 32 columns at a time: bias, exact GELU (nn.GELU default), fp16
                const int erow = quarter * 32 + lane;
                for (int c0 = wsub * 32; c0 < C; c0 += 64) {
                    float v[32];
                    tmem_ld32(tlane + (unsigned)c0, v);
                    unsigned hw[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const float a = v[j] + sBias[nc * C + c0 + j], b = v[j + 1] + sBias[nc * C + c0 + j + 1];
                        hw[j / 2] = pack_h2(0.5f * a * (1.0f + erff(a * 0.70710678118654752f)),
                                            0.5f * b * (1.0f + erff(b * 0.70710678118654752f)));
                    }
                    if (erow < nrows) {
                        uint4* dst = reinterpret_cast<uint4*>(p.hidden + (size_t)(row0 + erow) * H + nc * C + c0);
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) dst[q4] = make_uint4(hw[4 * q4], hw[4 * q4 + 1], hw[4 * q4 + 2], hw[4 * q4 + 3]);
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
                __syncthreads();                               // TMEM drained before the next chunk's MMA overwrites it
            }
        } else {
            for (int kc = 0; kc < nchunks; ++kc) {
                stage_chunk_f16(sA, lboA, p.hidden + (size_t)row0 * H, (size_t)H, kc * C, kLinTileM, nrows, C, tid);
                stage_chunk_f16(sW, lboW, p.w2h, (size_t)H, kc * C, C, C, C, tid);
                cp_async_commit();
                cp_async_wait_group0();
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                __syncthreads();
                if (tid == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                    for (int ks = 0; ks < C / 16; ++ks)
                        umma_f16(tmem_base, umma_desc(smem_u32(sA) + (unsigned)(2 * ks) * lboA, lboA, 128u),
                                 umma_desc(smem_u32(sW) + (unsigned)(2 * ks) * lboW, lboW, 128u), idesc, (kc > 0 || ks > 0) ? 1u : 0u);
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(mbar)) : "memory");
                }
                mbar_wait(mbar, phase);                        // both operand buffers are free again
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            }
            // epilogue: 32 columns at a time through a shared-memory transpose so that rows are read / written contiguously
            float* stage = reinterpret_cast<float*>(sA) + warp * (32 * 33);
            for (int c0 = wsub * 32; c0 < C; c0 += 64) {
                float v[32];
                tmem_ld32(tlane + (unsigned)c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = v[j] + sBias[c0 + j];
                __syncwarp();
                for (int r = 0; r < 32; ++r) {
                    const int row = row0 + quarter * 32 + r;
                    if (row < p.M) {
                        const size_t g = (size_t)row * C + c0 + lane;
                        p.y[g] = stage[r * 33 + lane] + (p.residual ? __ldg(p.x + g) : 0.f);
                    }
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncthreads();                                   // TMEM drained, stage region free for the next tile's operands
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

}  // namespace dlsg
#endif
