// MLP half of the Swin block (forward_part2, dl_cs/models/video_swin_transformer_mri_downsample.py:251-252 with
// Mlp :20-38 and norm2 :208):      out = [x +] fc2( GELU( fc1( LayerNorm(x) ) ) )
// on the 5th-generation tensor cores (fp16 operands, fp32 accumulation in TMEM), as two kernels:
//
//   swin_ln_gemm_stream_kernel<false> (fc1)   per 128-token tile: LayerNorm(x) -> fp16 A tile (once); the hidden layer is produced in
//                         chunks of 160 columns: the chunk of W1 (fp16, [H, C] row-major) streams in with cp.async
//                         while the previous chunk's epilogue (bias, exact erf GELU, fp16 store) runs
//   swin_mlp_fc2_kernel   per tile: K = H is walked in chunks of 160: fp16 hidden rows and the W2 chunk ([C, H] row-major)
//                         land with cp.async directly in the canonical UMMA layout, tcgen05.mma accumulates over the
//                         chunks in TMEM; epilogue adds bias (and the residual x) and writes fp32 rows
//
// Both operands use the no-swizzle K-major core-matrix layout of swin_linear_tc.cuh ([k/8][row/8][8 rows][8 halves]).  The
// weights are converted to fp16 and put in that order, chunk by chunk, once by the host module (constants at inference), and
// fc1 writes the hidden layer in the same order per (tile, chunk), so every operand is staged by a contiguous block copy.  C = 160 is the reference's embedding width;
// any C, H that are multiples of 32 with C <= 256 and H % C == 0 are accepted.
// CUDA-only (tcgen05 / TMEM have no host model); validated on the B200 against torch (tests/test_swin_gpu.py).
#pragma once

#include "swin_linear_tc.cuh"

#if defined(__CUDACC__)
namespace dlsg {

struct MlpParams {
    const float* x;        // [M, C] fp32 (tokens, channel fastest)
    const float* gamma; const float* beta; float eps;      // norm2
    const half_t* w1h;     // fc1 weight [H, C], fp16, tiled: chunk nc = output rows nc*C..: [H/C][k/8][n/8][n%8][k%8]
    const float* b1;       // [H]
    const half_t* w2h;     // fc2 weight [C, H], fp16, tiled: chunk kc = input columns kc*C..: [H/C][k/8][n/8][n%8][k%8]
    const float* b2;       // [C]
    half_t* hidden;        // fp16 workspace, ceil(M/128) * 128 * H elements: per (tile, chunk) a [128 x C] block in the same order
    float* y;              // [M, C]
    int M, C, H;
    int residual;          // 1: y = x + mlp(LN(x))  (the block's `x + forward_part2(x)`, :271);  0: y = mlp(LN(x))
    int tmem_cols;
    // QKV mode of the streamed LN-GEMM (H = 3C; chunk 0 / 1 / 2 = q / k / v): rows are window tokens of this geometry, gathered from
    // the feature volume x [B,D,H,W,C] through the cyclic shift (padded tokens are zero rows); outputs as in swin_linear_tc.cuh E_QKV16
    SwinGeom g;
    half_t* q16; half_t* k16; half_t* vt16;
    int heads, hd, ntok;
    float qscale;
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

// exact-form GELU 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below the fp16
// rounding of the hidden layer): one MUFU.RCP, one MUFU.EX2 and a degree-5 polynomial instead of erff's ~3x longer sequence
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float e = poly * t * fast_exp2(-1.4426950408889634f * z * z);     // 1 - erf(z)
    const float erf_abs = 1.0f - e;
    return 0.5f * x * (1.0f + copysignf(erf_abs, x));
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

// contiguous block copy global -> shared, 16 bytes per cp.async, consecutive threads on consecutive chunks (coalesced, conflict-free).
// Every operand block below is stored in global memory already in the canonical core-matrix order, so staging is just this.
__device__ __forceinline__ void copy_block_async(unsigned char* dst, const void* src, int nbytes, int tid) {
    const unsigned char* s8 = static_cast<const unsigned char*>(src);
#pragma unroll 4
    for (int i = tid * 16; i < nbytes; i += kLinThreads * 16) cp_async16(dst + i, s8 + i);
}

// Streamed LayerNorm-GEMM:  out chunk nc = epilogue( LN(rows of x) W[nc]^T + bias[nc] ),  nc < H / C.
//   QKV = false (fc1):  rows are the tokens in order; epilogue = exact GELU -> fp16 hidden in fc2's tile order
//   QKV = true        :  rows are window tokens gathered through the shift (norm1 + pad + roll + window_partition, :218-236);
//                        chunks are q, k, v; epilogue writes the attention core's fp16 [B_,heads,N,32] rows (q pre-scaled, ones column in v)
// The LayerNorm'ed fp16 A tile is built once per 128 rows; each weight chunk streams in by cp.async while the previous chunk's
// epilogue runs; two CTAs per SM.
template <bool QKV>
__global__ void __launch_bounds__(kLinThreads, 2) swin_ln_gemm_stream_kernel(MlpParams p) {
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
    for (int i = tid; i < H; i += kLinThreads) sBias[i] = p.b1 ? __ldg(p.b1 + i) : 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem_base = *tmem_slot;
    const unsigned idesc = umma_idesc_f16(kLinTileM, C);
    const int quarter = warp & 3, wsub = warp >> 2;            // warps q and q+4 share TMEM lanes 32q .. 32q+31
    const unsigned tlane = tmem_base + ((unsigned)(quarter * 32) << 16);
    const int nchunks = H / C;
    const int ntok = QKV ? swin_tokens(p.g) : 1, nwin = QKV ? swin_windows(p.g) : 1;

    const int ntiles = (p.M + kLinTileM - 1) / kLinTileM;
    unsigned phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = tile * kLinTileM;
        const int nrows = min(kLinTileM, p.M - row0);
        // ---- A tile = LayerNorm(x rows), two threads per row (each one half of C), one pass with shifted sums
        copy_block_async(sW, p.w1h, C * C * 2, tid);                               // first weight chunk (output rows 0 .. C-1)
        cp_async_commit();
        const int trow = tid & (kLinTileM - 1), khalf = tid >> 7;
        const int kc0 = khalf * (KC / 2), kc1 = khalf ? KC : KC / 2;
        const float* xr = nullptr;
        if (trow < nrows) {
            if (QKV) {
                const int row = row0 + trow, tok = row % ntok, bw = row / ntok;
                const int win = bw % nwin, b = bw / nwin;
                int d, h, w;
                if (swin_token_coords(p.g, win, tok, d, h, w))
                    xr = p.x + ((((size_t)b * p.g.D + d) * p.g.H + h) * p.g.W + w) * C;
            } else {
                xr = p.x + (size_t)(row0 + trow) * C;
            }
        }
        float s1 = 0.f, s2 = 0.f, x0 = 0.f;
        if (xr) {
            x0 = __ldg(xr);
#pragma unroll 5
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
#pragma unroll 5
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
            __syncthreads();                                   // A tile (first chunk) and this weight chunk are in place
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                for (int ks = 0; ks < C / 16; ++ks)
                    umma_f16(tmem_base, umma_desc(smem_u32(sA) + (unsigned)(2 * ks) * lboA, lboA, 128u),
                             umma_desc(smem_u32(sW) + (unsigned)(2 * ks) * lboW, lboW, 128u), idesc, ks > 0 ? 1u : 0u);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(mbar)) : "memory");
            }
            mbar_wait(mbar, phase);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            if (nc + 1 < nchunks)                              // the MMA has consumed sW: the next chunk lands during the epilogue
                copy_block_async(sW, p.w1h + (size_t)(nc + 1) * C * C, C * C * 2, tid);
            cp_async_commit();
            const int erow = quarter * 32 + lane;              // thread = token row of the tile
            if (QKV) {
                // chunk nc is q (0), k (1) or v (2): per head, hd columns -> one fp16 row of 32 (zero padded; v carries a one at index hd)
                const bool live = erow < nrows;                // tcgen05.ld is warp-collective: only the stores are predicated
                const int grow = row0 + erow;
                const int tok = live ? grow % p.ntok : 0, bw = live ? grow / p.ntok : 0;
                half_t* base16 = (nc == 0) ? p.q16 : (nc == 1 ? p.k16 : p.vt16);
                const float sc = (nc == 0) ? p.qscale : 1.0f;
                for (int head = wsub; head < p.heads; head += 2) {
                    const int c0 = head * p.hd;
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4)
                        if (4 * j4 < p.hd) tmem_ld4(tlane + (unsigned)(c0 + 4 * j4), v + 4 * j4);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)               // index hd of v: a one (the attention core reads the softmax row sums off the PV MMA)
                        v[j] = (j < p.hd) ? (v[j] + sBias[nc * C + c0 + (j < p.hd ? j : 0)]) * sc : ((nc == 2 && j == p.hd) ? 1.0f : 0.f);
                    if (live) {
                        uint4* dst = reinterpret_cast<uint4*>(base16 + (((size_t)bw * p.heads + head) * p.ntok + tok) * 32);
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4)
                            dst[q4] = make_uint4(pack_h2(v[8 * q4], v[8 * q4 + 1]), pack_h2(v[8 * q4 + 2], v[8 * q4 + 3]),
                                                 pack_h2(v[8 * q4 + 4], v[8 * q4 + 5]), pack_h2(v[8 * q4 + 6], v[8 * q4 + 7]));
                    }
                }
            } else {
                // 32 columns at a time: bias, exact GELU (nn.GELU default), fp16
                for (int c0 = wsub * 32; c0 < C; c0 += 64) {
                    float v[32];
                    tmem_ld32(tlane + (unsigned)c0, v);
                    unsigned hw[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const float a = v[j] + sBias[nc * C + c0 + j], b = v[j + 1] + sBias[nc * C + c0 + j + 1];
                        hw[j / 2] = pack_h2(gelu_erf(a), gelu_erf(b));
                    }
                    // hidden is stored per (tile, chunk) in the canonical A layout of fc2: for one k8 the 32 rows of a warp are 512 contiguous bytes
                    unsigned char* hb = reinterpret_cast<unsigned char*>(p.hidden + ((size_t)tile * nchunks + nc) * (size_t)(kLinTileM * C));
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4)
                        *reinterpret_cast<uint4*>(hb + canon_off(erow, c0 / 8 + q4, lboA)) =
                            make_uint4(hw[4 * q4], hw[4 * q4 + 1], hw[4 * q4 + 2], hw[4 * q4 + 3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncthreads();                                   // TMEM drained before the next chunk's MMA overwrites it
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

inline size_t mlp_fc2_smem_bytes(int C, int stages) {
    return (size_t)stages * (linear_a_region_bytes(C) + (size_t)C * C * 2) + (size_t)C * 4 + 64;
}

// y = [x +] hidden W2^T + b2.  One CTA per SM; the (tile, K chunk) items run through a two-stage ring: while the
// tensor core works on one stage, the fp16 hidden rows and the W2 chunk of the next item (also the first chunk of the
// next tile, during the epilogue) stream into the other stage with cp.async.
template <int STAGES>     // 2: one CTA per SM with the ring;  1: two CTAs per SM cover each other's staging instead
__global__ void __launch_bounds__(kLinThreads, 3 - STAGES) swin_mlp_fc2_kernel(MlpParams p) {
    extern __shared__ __align__(128) unsigned char mlp_smem[];
    const int C = p.C, H = p.H;
    const size_t a_bytes = linear_a_region_bytes(C), w_bytes = (size_t)C * C * 2;
    auto stA = [&](int st) { return mlp_smem + (size_t)st * (a_bytes + w_bytes); };      // stage st: A region, then the W chunk
    auto stW = [&](int st) { return mlp_smem + (size_t)st * (a_bytes + w_bytes) + a_bytes; };
    float* sBias = reinterpret_cast<float*>(mlp_smem + STAGES * (a_bytes + w_bytes));
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(sBias + C);
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
    for (int i = tid; i < C; i += kLinThreads) sBias[i] = p.b2 ? __ldg(p.b2 + i) : 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem_base = *tmem_slot;
    const unsigned idesc = umma_idesc_f16(kLinTileM, C);
    const int quarter = warp & 3, wsub = warp >> 2;
    const unsigned tlane = tmem_base + ((unsigned)(quarter * 32) << 16);

    const int nk = H / C;
    const int ntiles = (p.M + kLinTileM - 1) / kLinTileM;
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int nitems = my_tiles * nk;
    auto issue = [&](int item) {
        const int tile = blockIdx.x + (item / nk) * gridDim.x, kc = item % nk, s = item % STAGES;
        copy_block_async(stA(s), p.hidden + ((size_t)tile * nk + kc) * (size_t)(kLinTileM * C), kLinTileM * C * 2, tid);
        copy_block_async(stW(s), p.w2h + (size_t)kc * C * C, C * C * 2, tid);
        cp_async_commit();
    };
    if (STAGES == 2 && nitems > 0) issue(0);
    unsigned phase = 0;
    bool pending = false;                                      // an MMA commit that has not been waited for
    for (int i = 0; i < nitems; ++i) {
        if (pending) {                                         // MMA(i-1) done: the other stage is free again
            mbar_wait(mbar, phase);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            pending = false;
        }
        if (STAGES == 1) { issue(i); cp_async_wait_group0(); }
        else if (i + 1 < nitems) { issue(i + 1); cp_async_wait_group1(); }
        else cp_async_wait_group0();
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();                                       // item i's operands are in place for every thread
        const int s = i % STAGES, kc = i % nk;
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            for (int ks = 0; ks < C / 16; ++ks)
                umma_f16(tmem_base, umma_desc(smem_u32(stA(s)) + (unsigned)(2 * ks) * lboA, lboA, 128u),
                         umma_desc(smem_u32(stW(s)) + (unsigned)(2 * ks) * lboW, lboW, 128u), idesc, (kc > 0 || ks > 0) ? 1u : 0u);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(mbar)) : "memory");
        }
        pending = true;
        if (kc == nk - 1) {
            mbar_wait(mbar, phase);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            pending = false;
            // epilogue: 32 columns at a time through a transpose in this stage's (now idle) A region: rows are read / written contiguously
            const int row0 = (blockIdx.x + (i / nk) * gridDim.x) * kLinTileM;
            float* stage = reinterpret_cast<float*>(stA(s)) + warp * (32 * 33);
            for (int c0 = wsub * 32; c0 < C; c0 += 64) {
                // the residual rows of this 32 x 32 block: all 32 loads in flight before the accumulator is touched
                float xr[32];
                const float* xb = p.x + (size_t)(row0 + quarter * 32) * C + c0 + lane;
                const int live = min(32, p.M - (row0 + quarter * 32));        // rows of this block inside the matrix (may be <= 0)
#pragma unroll
                for (int r = 0; r < 32; ++r) xr[r] = (p.residual && r < live) ? __ldg(xb + (size_t)r * C) : 0.f;
                float v[32];
                tmem_ld32(tlane + (unsigned)c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = v[j] + sBias[c0 + j];
                __syncwarp();
                float* yb = p.y + (size_t)(row0 + quarter * 32) * C + c0 + lane;
#pragma unroll
                for (int r = 0; r < 32; ++r)
                    if (r < live) yb[(size_t)r * C] = stage[r * 33 + lane] + xr[r];
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncthreads();                                   // TMEM drained and the scratch free before the ring moves on
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

}  // namespace dlsg
#endif
