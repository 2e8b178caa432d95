// Convolutions of the denoisers (nn.Conv3d with 'same' zero padding: dl_cs/models/swin3D.py:120-134, the 160 -> 160 / 64 -> 64
// layers of ConvBlock :225-273; PatchEmbed3D / PatchUnembed3D of video_swin_transformer_mri_downsample.py:455-520) as implicit
// GEMMs on the 5th-generation tensor cores, channels-last, with the operand split fused into the operand staging:
//
//     out[v, co] = bias[co] + sum_{tap, ci} x[v + tap, ci] * w[co, ci, tap]          v = (n, t, y, x) voxel
//                = sum (x_hi * w_hi + x_lo * w_hi + x_hi * w_lo),   a = a_hi + a_lo + O(2^-18 a),  a_hi = bf16(a),  a_lo = bf16(a - a_hi)
//
// i.e. the "3x" scheme of models.set_conv_precision("tf32x3") on 16-bit pieces: 16 mantissa bits per operand, fp32 accumulation,
// 1e-5 relative on a layer's output - the level the tensor core's own accumulator leaves at K = 27 * 160 - where one TF32 pass
// gives 5e-4.  bf16 pieces instead of TF32 ones because the kernel is bound by shared-memory wavefronts (operand reads of the
// MMAs + the producers' stores share one 128 B/clk pipe; ncu, profiles/r2r_conv_*): 2-byte operands halve both and run at twice
// the MMA rate (first version, TF32 pieces: 4.33 ms for the 160 -> 160 layer of one cine slice; this one: 2.87 ms).
//
// One persistent CTA per SM owns 256 consecutive voxels per tile = two 128-row accumulators [128 x Cout] fp32 in tensor
// memory.  The K loop walks taps x Cin/32 channel chunks through a four-stage shared-memory ring:
//   * 16 producer warps read the 32 input channels of the shifted voxels (8 lanes per 128-byte line, coalesced; zeros outside the
//     volume; loads issued one iteration ahead), split them into the bf16 head and remainder, and store both in the canonical
//     no-swizzle K-major UMMA layout ([k/8][row/8][8 rows][8 bf16], planes padded by 32 bytes against bank conflicts);
//   * the weights arrive already split and already in that layout (prepared once per weight version on the host side)
//     by one bulk copy per stage (cp.async.bulk, completion counted on the stage's mbarrier);
//   * one thread issues 12 tcgen05.mma.kind::f16 (bf16 operands) per stage (2 tiles x 2 k-steps x 3 products) and commits the
//     stage back to the producers.
// After the last stage the producer warps drain the accumulators (tcgen05.ld), add the bias and the optional skip connection and
// store channels-last rows.  The same kernel serves the first layer (Cin = 2 / 4, K gathered over taps), the patch embedding
// (kernel = stride) and the transposed patch unembedding; Cout = 64 has an operand-swapped variant (below).
// CUDA-only: tcgen05 / TMEM have no host emulation (validated on the B200 against fp64 torch, tests/test_conv_gpu.py).
#pragma once

#include "cuda_compat.h"
#include "tc_common.cuh"

#if defined(__CUDACC__)
namespace dlsg {

constexpr int kConvProducers = 512;                // 16 producer warps: two threads per voxel row pair ... four rows per thread and stage
constexpr int kConvThreads = kConvProducers + 32;
constexpr int kConvRows = 4;                       // rows per producer thread and stage
constexpr int kConvStages = 4;
constexpr int kConvKC = 32;                        // input channels per stage
constexpr int kConvTileRows = 256;                 // voxels per CTA tile (two UMMA M = 128 accumulators)
constexpr int kConvLboA = 128 * 16 + 32;           // bytes between the 8-channel planes (8 bf16 = 16 bytes per row) of an A block: +32 so that the 4 planes a half-warp stores hit different banks
constexpr int kConvABytes = (kConvKC / 8) * kConvLboA;   // one [128 x 32] bf16 operand block

enum : int { CONV_GATHER = 0, CONV_UNPATCH = 1 };
struct ConvParams {
    const float* x;        // [NB, T, Y, X, Cin]          (input volume)
    const float* wprep;    // bf16 image [taps][Cin/32][hi|lo][4][Cout/8][8][8] (include/dlsg_b200.h)
    const float* bias;     // [Cout] or NULL
    float* out;            // [NB, To, Yo, Xo, Cout]
    int NB, T, Y, X, Cin, Cout;
    int To, Yo, Xo;        // output volume
    int ksz, stride, pad;  // CONV_GATHER: out[o] = sum over the ksz^3 taps of x[o * stride - pad + tap];  3,1,1 = 'same' 3x3x3;  P,P,0 = patch embedding
                           // CONV_UNPATCH: out[i * ksz + tap] = W_tap x[i]  (transposed convolution with kernel = stride = ksz: patch unembedding)
    int mode;
    int tlo, tn;           // CONV_GATHER: only the output frames tlo .. tlo + tn - 1 of every sample are computed (tn = To: all).  The denoisers
                           // pad the time axis circularly and crop it at the end (swin3D.py:401-417): a layer's outer frames that no later
                           // layer reads are never computed; the frames that are computed are bit-identical to the full convolution
    long long NV;          // GEMM rows: output voxels in the frame window (CONV_GATHER) or input voxels (CONV_UNPATCH)
    int relu_in;           // rectify the input on the fly (ConvBlock's pre-activation)
    const float* resid;    // optional [NB, To, Yo, Xo, Cout]: out = conv + bias + resid_scale * resid (the blocks' skip connections)
    float resid_scale;
    int resid_relu;        // the skip connection is relu(resid): ResBlock's in-place pre-activation rectifies the tensor the skip re-reads (resnet3d.py:240)
};

__host__ __device__ inline size_t conv_stage_bytes(int Cout) { return (size_t)4 * kConvABytes + (size_t)2 * Cout * kConvKC * 2; }
inline size_t conv_smem_bytes(int Cout) { return kConvStages * conv_stage_bytes(Cout) + (size_t)Cout * 4 + 128; }

// kind::f16 with bf16 operands (format 1), fp32 accumulation
__device__ __forceinline__ unsigned umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
// four floats -> bf16 heads (round to nearest) and bf16 remainders, each packed as 4 x 16 bit
__device__ __forceinline__ void bf16_split4(float4 v, uint2& hi, uint2& lo) {
    unsigned h01, h23;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h01) : "f"(v.y), "f"(v.x));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h23) : "f"(v.w), "f"(v.z));
    const float r0 = v.x - __uint_as_float(h01 << 16), r1 = v.y - __uint_as_float(h01 & 0xffff0000u);
    const float r2 = v.z - __uint_as_float(h23 << 16), r3 = v.w - __uint_as_float(h23 & 0xffff0000u);
    unsigned l01, l23;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l01) : "f"(r1), "f"(r0));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l23) : "f"(r3), "f"(r2));
    hi = make_uint2(h01, h23);
    lo = make_uint2(l01, l23);
}
// bounded wait: a broken pipeline traps instead of hanging the GPU
__device__ __forceinline__ void conv_mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    for (unsigned spin = 0; !done; ++spin) {             // try_wait suspends for a while by itself; 2^26 rounds is seconds
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void conv_bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// CS = 0: Cin % 32 == 0, the K loop walks (tap, 32-channel chunk).  CS = 2 / 4: the first layer of a denoiser (Cin = 2 or 4 = re / im of the
// E map sets): K runs over k = tap * Cin + ci, padded with zero weights to a multiple of 32; a 16-byte quad of a stage is one tap (CS = 4)
// or two taps (CS = 2) of one voxel, so a stage gathers 8 / 16 taps and the whole layer is 4 / 2 stages per tile.
template <int CS, bool RELU>
__global__ void __launch_bounds__(kConvThreads, 1) conv3d_k3_tf32x3_kernel(ConvParams p) {
    extern __shared__ __align__(1024) unsigned char conv_smem[];
    const int Cout = p.Cout, KCH = (CS == 0) ? p.Cin / kConvKC : 1;
    const int ksz = p.ksz, ntaps = ksz * ksz * ksz;
    const bool unpatch = (CS == 0) && p.mode == CONV_UNPATCH;
    const int NIT = (CS == 0) ? (unpatch ? KCH : ntaps * KCH) : (27 * CS + kConvKC - 1) / kConvKC;
    const size_t stage_bytes = conv_stage_bytes(Cout);
    const unsigned wblock = (unsigned)Cout * kConvKC * 2;            // one of w_hi / w_lo (bf16) for a stage
    float* sBias = reinterpret_cast<float*>(conv_smem + kConvStages * stage_bytes);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sBias + Cout);   // full[S], empty[S], accfull, accempty
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 2 * kConvStages + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned s0 = smem_u32(conv_smem), bar0 = smem_u32(bars);
    const unsigned bempty = bar0 + 8u * kConvStages, baccfull = bar0 + 16u * kConvStages, baccempty = baccfull + 8u;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < kConvStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bar0 + 8u * s), "r"(kConvProducers + 1) : "memory");        // full
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bempty + 8u * s), "r"(1) : "memory");                      // empty
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(baccfull), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(baccempty), "r"(kConvProducers) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int i = tid; i < Cout; i += kConvThreads) sBias[i] = p.bias ? __ldg(p.bias + i) : 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem_base = *tmem_slot;

    // CONV_UNPATCH: a tile is (256 input voxels, one of the ksz^3 output positions inside their patches); the taps of a row tile are consecutive tiles
    const long long rtiles = (p.NV + kConvTileRows - 1) / kConvTileRows;
    const long long ntiles = unpatch ? rtiles * ntaps : rtiles;
    const long long YX = (long long)p.Y * p.X;                       // input plane
    const long long YXo = (long long)p.Yo * p.Xo;                    // output plane

    if (warp < kConvProducers / 32) {
        // =========================== producers, then epilogue ===========================
        // loads are coalesced: 8 lanes cover the 128 bytes of one voxel's channel chunk, a warp instruction takes 4 voxels;
        // thread (lane) owns channel quad k4 = lane & 7 of rows 4 i + (lane >> 3), i = 0..7, of its warp's 32 rows
        const int ti = warp >> 3;                                // which of the two 128-row blocks this warp fills
        const int k4 = lane & 7, rsub = lane >> 3;
        const int wrow0 = (warp & 7) * 16;                       // first of this warp's 16 rows inside its 128-row block
        unsigned g = 0, tcount = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
            const long long rtile = unpatch ? tile / ntaps : tile;
            const int cls = unpatch ? (int)(tile - rtile * ntaps) : 0;
            const long long v0 = rtile * kConvTileRows + ti * 128 + wrow0 + rsub;     // this thread's GEMM rows: v0 + 4 i
            // per row: packed (t, y, x) + 1 of the first tap's input voxel (-1 = row beyond the problem) and that voxel's index
            int coord[kConvRows], ibase[kConvRows];
#pragma unroll
            for (int i = 0; i < kConvRows; ++i) {
                const long long vi = v0 + 4 * i;
                coord[i] = -1; ibase[i] = 0;
                if (vi < p.NV) {
                    if (unpatch) { coord[i] = 0; ibase[i] = (int)vi; }
                    else {
                        const long long r1 = vi % YXo;
                        const int n = (int)(vi / (YXo * p.tn));
                        const int t0 = ((int)((vi / YXo) % p.tn) + p.tlo) * p.stride - p.pad, y0 = (int)(r1 / p.Xo) * p.stride - p.pad,
                                  x0 = (int)(r1 % p.Xo) * p.stride - p.pad;
                        coord[i] = (t0 + 1) << 20 | (y0 + 1) << 10 | (x0 + 1);
                        ibase[i] = (int)((((long long)n * p.T + t0) * p.Y + y0) * p.X + x0);
                    }
                }
            }
            unsigned okmask = 0;
            int tapoff = 0;
            // shifted voxel of row i for one tap: pointer to its Cin values, or NULL outside the volume (small-Cin modes, 3 x 3 x 3 only)
            auto tap_ptr = [&](int i, int tap) -> const float* {
                const int c = coord[i];
                const int a = tap / 9, b2 = (tap / 3) % 3, c2 = tap % 3;
                const bool ok = tap < 27 && c >= 0 && (unsigned)((c >> 20) + a - 1) < (unsigned)p.T && (unsigned)(((c >> 10) & 1023) + b2 - 1) < (unsigned)p.Y &&
                                (unsigned)((c & 1023) + c2 - 1) < (unsigned)p.X;
                return ok ? p.x + ((long long)ibase[i] + ((long long)a * p.Y + b2) * p.X + c2) * CS : nullptr;
            };
            // the global loads run one iteration ahead of the split / store (their L2 latency is longer than a stage; with 216 KB of
            // shared memory in use there is no L1 to speak of): `fetch(it, dst)` issues iteration it's loads, its tap state is its own
            // element offset of each row's first-tap voxel (this thread's channel quad); all address arithmetic in the loop is 32-bit
            unsigned roff[kConvRows];
#pragma unroll
            for (int i = 0; i < kConvRows; ++i) roff[i] = (unsigned)ibase[i] * (unsigned)((CS == 0) ? p.Cin : CS) + ((CS == 0) ? 4u * k4 : 0u);
            int ftap = 0, fch = 0;                                  // (tap, chunk) of the next fetch
            unsigned foff = 0;                                      // tap offset * Cin + chunk * 32 of the next fetch
            auto tap_state = [&](int tap) {
                okmask = 0;
                if (unpatch) {
#pragma unroll
                    for (int i = 0; i < kConvRows; ++i) okmask |= (coord[i] >= 0 ? 1u : 0u) << i;
                    tapoff = 0;
                } else {
                    const int ta = tap / (ksz * ksz), tb = (tap / ksz) % ksz, tc = tap % ksz;
#pragma unroll
                    for (int i = 0; i < kConvRows; ++i) {
                        const int c = coord[i];
                        const bool ok = c >= 0 && (unsigned)((c >> 20) + ta - 1) < (unsigned)p.T && (unsigned)(((c >> 10) & 1023) + tb - 1) < (unsigned)p.Y &&
                                        (unsigned)((c & 1023) + tc - 1) < (unsigned)p.X;
                        okmask |= (ok ? 1u : 0u) << i;
                    }
                    tapoff = (ta * p.Y + tb) * p.X + tc;
                }
                foff = (unsigned)tapoff * (unsigned)p.Cin;
            };
            if (CS == 0) tap_state(0);
            auto fetch = [&](int it, float4* a) {
                if (CS == 0) {
#pragma unroll
                    for (int i = 0; i < kConvRows; ++i)
                        a[i] = ((okmask >> i) & 1u) ? __ldg(reinterpret_cast<const float4*>(p.x + (roff[i] + foff))) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (++fch == KCH) { fch = 0; if (++ftap < ntaps) tap_state(ftap); } else foff += kConvKC;
                } else if (CS == 4) {
#pragma unroll
                    for (int i = 0; i < kConvRows; ++i) {
                        const float* q = tap_ptr(i, 8 * it + k4);
                        a[i] = q ? __ldg(reinterpret_cast<const float4*>(q)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < kConvRows; ++i) {
                        const float* q0 = tap_ptr(i, 2 * (8 * it + k4));
                        const float* q1 = tap_ptr(i, 2 * (8 * it + k4) + 1);
                        const float2 lo2 = q0 ? __ldg(reinterpret_cast<const float2*>(q0)) : make_float2(0.f, 0.f);
                        const float2 hi2 = q1 ? __ldg(reinterpret_cast<const float2*>(q1)) : make_float2(0.f, 0.f);
                        a[i] = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
                    }
                }
            };
            float4 a0[kConvRows], a1[kConvRows];
            auto stage = [&](int it, const float4* a) {
                    const unsigned s = g % kConvStages, use = g / kConvStages;
                    conv_mbar_wait(bempty + 8u * s, (use & 1u) ^ 1u);                 // stage free (passes at once on first use)
                    unsigned char* st = conv_smem + s * stage_bytes;
                    if (tid == 0) {
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar0 + 8u * s), "r"(2u * wblock) : "memory");
                        conv_bulk_load(s0 + (unsigned)(s * stage_bytes) + 4u * kConvABytes,
                                       reinterpret_cast<const unsigned char*>(p.wprep) + ((size_t)cls * KCH + it) * 2 * (size_t)wblock, 2u * wblock, bar0 + 8u * s);
                    }
                    // this thread's 4 channels are half of a 16-byte core-matrix row: plane k4 / 2, bytes 8 (k4 % 2) ..
                    unsigned char* ahi = st + (size_t)(2 * ti) * kConvABytes + (size_t)(k4 >> 1) * kConvLboA + (k4 & 1) * 8;
                    unsigned char* alo = ahi + kConvABytes;
                    // row = wrow0 + rsub + 4 i with rsub < 4: row i sits (i / 2) * 128 + (i % 2) * 64 bytes after row 0
                    const unsigned off0 = (unsigned)((wrow0 + rsub) >> 3) * 128u + (unsigned)(rsub & 7) * 16u;
#pragma unroll
                    for (int i = 0; i < kConvRows; ++i) {
                        const unsigned off = off0 + (unsigned)(i >> 1) * 128u + (unsigned)(i & 1) * 64u;
                        float4 x4 = a[i];
                        if (RELU) { x4.x = fmaxf(x4.x, 0.f); x4.y = fmaxf(x4.y, 0.f); x4.z = fmaxf(x4.z, 0.f); x4.w = fmaxf(x4.w, 0.f); }
                        uint2 h2, l2;
                        bf16_split4(x4, h2, l2);
                        *reinterpret_cast<uint2*>(ahi + off) = h2;
                        *reinterpret_cast<uint2*>(alo + off) = l2;
                    }
                    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(bar0 + 8u * s) : "memory");
            };
            fetch(0, a0);
            for (int it = 0; it < NIT; it += 2) {                    // two iterations per round: the register sets alternate without copies
                if (it + 1 < NIT) fetch(it + 1, a1);
                stage(it, a0);
                ++g;
                if (it + 1 < NIT) {
                    if (it + 2 < NIT) fetch(it + 2, a0);
                    stage(it + 1, a1);
                    ++g;
                }
            }
            // ---- epilogue: this warp's 32 TMEM lanes of its accumulator
            conv_mbar_wait(baccfull, tcount & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            // epilogue: thread = accumulator row (TMEM lane 32 (warp % 4) + lane) of block eti; warps w and w + 8 share rows and split the columns
            const int eti = (warp >> 2) & 1, ehalf = warp >> 3;
            const long long v = rtile * kConvTileRows + eti * 128 + (warp & 3) * 32 + lane;
            const bool live = v < p.NV;
            const unsigned taddr = tmem_base + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)(eti * 256);
            long long vout = v;
            if (!unpatch && p.tn != p.To) vout = v + ((v / (YXo * p.tn)) * (p.To - p.tn) + p.tlo) * YXo;     // row of the frame window -> voxel of the volume
            if (unpatch && live) {                                                   // input voxel (n, t, y, x) -> output voxel (n, t k + a, y k + b, x k + c)
                const long long r1 = v % YX;
                const long long nt = v / YX;                                          // n * T + t
                const int ta = cls / (ksz * ksz), tb = (cls / ksz) % ksz, tc = cls % ksz;
                vout = (((nt / p.T) * p.To + (nt % p.T) * ksz + ta) * p.Yo + (r1 / p.X) * ksz + tb) * p.Xo + (r1 % p.X) * ksz + tc;
            }
            float* orow = p.out + (size_t)vout * Cout;
            const float* rrow = p.resid ? p.resid + (size_t)vout * Cout : nullptr;
            for (int c0 = 32 * ehalf; c0 < Cout; c0 += 64) {
                float acc[32];
                tmem_ld32(taddr + (unsigned)c0, acc);
                if (live && rrow) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 r4 = __ldg(reinterpret_cast<const float4*>(rrow + c0 + 4 * q));
                        if (p.resid_relu) { r4.x = fmaxf(r4.x, 0.f); r4.y = fmaxf(r4.y, 0.f); r4.z = fmaxf(r4.z, 0.f); r4.w = fmaxf(r4.w, 0.f); }
                        acc[4 * q] = fmaf(p.resid_scale, r4.x, acc[4 * q]); acc[4 * q + 1] = fmaf(p.resid_scale, r4.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(p.resid_scale, r4.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(p.resid_scale, r4.w, acc[4 * q + 3]);
                    }
                }
                if (live) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4*>(orow + c0 + 4 * q) =
                            make_float4(acc[4 * q] + sBias[c0 + 4 * q], acc[4 * q + 1] + sBias[c0 + 4 * q + 1],
                                        acc[4 * q + 2] + sBias[c0 + 4 * q + 2], acc[4 * q + 3] + sBias[c0 + 4 * q + 3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(baccempty) : "memory");
        }
    } else if (lane == 0) {
        // =========================== MMA issuer ===========================
        const unsigned idesc = umma_idesc_bf16(128, Cout);
        const unsigned lboA = (unsigned)kConvLboA, lboW = (unsigned)(Cout / 8) * 128u;
        unsigned g = 0, tcount = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
            conv_mbar_wait(baccempty, (tcount & 1u) ^ 1u);                           // accumulators drained
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            for (int it = 0; it < NIT; ++it, ++g) {
                const unsigned s = g % kConvStages, use = g / kConvStages;
                conv_mbar_wait(bar0 + 8u * s, use & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const unsigned sa = s0 + (unsigned)(s * stage_bytes), sw = sa + 4u * kConvABytes;
#pragma unroll
                for (int ti = 0; ti < 2; ++ti) {
                    const unsigned d = tmem_base + (unsigned)(ti * 256);
#pragma unroll
                    for (int j = 0; j < kConvKC / 16; ++j) {                          // K = 16 per instruction = two 8-channel planes
                        const unsigned long long ahi = umma_desc(sa + (unsigned)(2 * ti) * kConvABytes + (unsigned)(2 * j) * lboA, lboA, 128u);
                        const unsigned long long alo = umma_desc(sa + (unsigned)(2 * ti + 1) * kConvABytes + (unsigned)(2 * j) * lboA, lboA, 128u);
                        const unsigned long long whi = umma_desc(sw + (unsigned)(2 * j) * lboW, lboW, 128u);
                        const unsigned long long wlo = umma_desc(sw + wblock + (unsigned)(2 * j) * lboW, lboW, 128u);
                        umma_f16(d, alo, whi, idesc, (it > 0 || j > 0) ? 1u : 0u);
                        umma_f16(d, ahi, wlo, idesc, 1u);
                        umma_f16(d, ahi, whi, idesc, 1u);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(bempty + 8u * s) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(baccfull) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"(512) : "memory");
    }
}


}  // namespace dlsg
#endif

// =====================================================================================================================
// Cout = 64 (the ResNet denoiser's 64 -> 64 layers, dl_cs/models/resnet3d.py:243-317; configs/config_recon.yaml): with 64 output
// channels the kernel above spends its time re-reading the activation block for every N = 64 instruction (tensor pipe
// 35 % active).  Here the operands swap roles: the MMA's A operand is the weight block [W_hi; W_lo] (2 x 64 = 128 rows) and the
// 256 voxels of the tile are its N dimension,
//     D[128 x 256] += [W_hi; W_lo](128 x 8) * X_hi^T(8 x 256)  and  ... * X_lo^T,
// two N = 256 instructions per k-step (all four products): rows 0-63 collect W_hi X_hi + W_hi X_lo, rows 64-127 W_lo X_hi + W_lo X_lo, and
// out[v, co] = D[co, v] + D[64 + co, v] + bias.  The epilogue reads the accumulator one output channel per thread and 32 voxels
// per load, the lo half goes through shared memory to the warp that holds the hi half, and each store writes one voxel's 32
// channels (128 bytes, coalesced).
#if defined(__CUDACC__)
namespace dlsg {

constexpr int kC64Producers = 512;                      // 16 producer warps (4 per scheduler): the producers are latency-, not throughput-bound
constexpr int kC64Threads = kC64Producers + 32;
constexpr int kC64LboB = 32 * 128 + 32;                 // bytes between the 8-channel planes of the 256-row activation block (+32: bank spread)
constexpr int kC64BBytes = (kConvKC / 8) * kC64LboB;    // one of X_hi / X_lo (bf16) for a stage
constexpr int kC64WBytes = 128 * kConvKC * 2;           // [W_hi; W_lo] (bf16) for a stage
constexpr int kC64Stage = 2 * kC64BBytes + kC64WBytes;
inline size_t conv_c64_smem_bytes() { return (size_t)kConvStages * kC64Stage + 64 * 4 + 128 + 28 * 4 + 27 * kC64Producers; }

template <bool RELU>
__global__ void __launch_bounds__(kC64Threads, 1) conv3d_k3_c64_kernel(ConvParams p) {
    extern __shared__ __align__(1024) unsigned char conv_smem[];
    constexpr int Cout = 64;
    const int KCH = p.Cin / kConvKC, NIT = 27 * KCH;
    float* sBias = reinterpret_cast<float*>(conv_smem + kConvStages * kC64Stage);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sBias + Cout);   // full[S], empty[S], accfull, accempty
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 2 * kConvStages + 2);
    int* sTapOff = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(bars) + 128);      // voxel offset of each of the 27 taps
    unsigned char* sOk = reinterpret_cast<unsigned char*>(sTapOff + 28);                      // [27][producer thread]: in-volume flags of its 4 rows
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned s0 = smem_u32(conv_smem), bar0 = smem_u32(bars);
    const unsigned bempty = bar0 + 8u * kConvStages, baccfull = bar0 + 16u * kConvStages, baccempty = baccfull + 8u;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < kConvStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bar0 + 8u * s), "r"(kC64Producers + 1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(bempty + 8u * s), "r"(1) : "memory");
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(baccfull), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(baccempty), "r"(kC64Producers) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int i = tid; i < Cout; i += kC64Threads) sBias[i] = p.bias ? __ldg(p.bias + i) : 0.f;
    if (tid < 27) sTapOff[tid] = ((tid / 9 - 1) * p.Y + ((tid / 3) % 3 - 1)) * p.X + (tid % 3 - 1);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const unsigned tmem_base = *tmem_slot;
    const long long ntiles = (p.NV + kConvTileRows - 1) / kConvTileRows;
    const long long YX = (long long)p.Y * p.X;

    if (warp < kC64Producers / 32) {
        const int k4 = lane & 7, rsub = lane >> 3;
        const int wrow0 = warp * 16;                            // first of this warp's 16 rows of the 256-row tile
        unsigned g = 0, tcount = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
            const long long v0 = tile * kConvTileRows + wrow0 + rsub;
            const long long win = YX * p.tn;                        // GEMM rows per sample (frame window tlo .. tlo + tn - 1)
            int coord[4];
            long long vact[4];                                      // voxel of the volume behind each row
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long vi = v0 + 4 * i;
                vact[i] = vi + ((vi / win) * (p.T - p.tn) + p.tlo) * YX;
                if (vi < p.NV) {
                    const long long r1 = vi % YX;
                    coord[i] = ((int)((vi / YX) % p.tn) + p.tlo) << 20 | (int)(r1 / p.X) << 10 | (int)(r1 % p.X);
                } else coord[i] = -1;
            }
            // the loads run one iteration ahead of the split / store; 32-bit element offsets, (tap, chunk) advanced incrementally
            unsigned okmask = 0, foff = 0;
            unsigned roff[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) roff[i] = (unsigned)vact[i] * (unsigned)p.Cin + 4u * k4;
            // the in-volume flags of all 27 taps are worked out once per tile (they cost more than the split of a whole stage) and kept in
            // this thread's own shared-memory bytes
            for (int tap = 0; tap < 27; ++tap) {
                const int dt = tap / 9 - 1, dy = (tap / 3) % 3 - 1, dx = tap % 3 - 1;
                unsigned m = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = coord[i];
                    const bool ok = c >= 0 && (unsigned)((c >> 20) + dt) < (unsigned)p.T && (unsigned)(((c >> 10) & 1023) + dy) < (unsigned)p.Y &&
                                    (unsigned)((c & 1023) + dx) < (unsigned)p.X;
                    m |= (ok ? 1u : 0u) << i;
                }
                sOk[tap * kC64Producers + tid] = (unsigned char)m;
            }
            auto tap_state = [&](int tap) {
                okmask = sOk[tap * kC64Producers + tid];
                foff = (unsigned)sTapOff[tap] * (unsigned)p.Cin;
            };
            int ftap = 0, fch = 0;
            auto fetch = [&](float4* dst) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    dst[i] = ((okmask >> i) & 1u) ? __ldg(reinterpret_cast<const float4*>(p.x + (roff[i] + foff))) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (++fch == KCH) { fch = 0; if (++ftap < 27) tap_state(ftap); } else foff += kConvKC;
            };
            const unsigned off0 = (unsigned)((wrow0 + rsub) >> 3) * 128u + (unsigned)(rsub & 7) * 16u;
            auto stage = [&](int it, const float4* a) {
                const unsigned s = g % kConvStages, use = g / kConvStages;
                conv_mbar_wait(bempty + 8u * s, (use & 1u) ^ 1u);
                unsigned char* st = conv_smem + (size_t)s * kC64Stage;
                if (tid == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(bar0 + 8u * s), "r"((unsigned)kC64WBytes) : "memory");
                    conv_bulk_load(s0 + (unsigned)(s * kC64Stage) + 2u * kC64BBytes,
                                   reinterpret_cast<const unsigned char*>(p.wprep) + (size_t)it * kC64WBytes, (unsigned)kC64WBytes, bar0 + 8u * s);
                }
                unsigned char* bhi = st + (size_t)(k4 >> 1) * kC64LboB + (k4 & 1) * 8;
                unsigned char* blo = bhi + kC64BBytes;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const unsigned off = off0 + (unsigned)(i >> 1) * 128u + (unsigned)(i & 1) * 64u;
                    float4 x4 = a[i];
                    if (RELU) { x4.x = fmaxf(x4.x, 0.f); x4.y = fmaxf(x4.y, 0.f); x4.z = fmaxf(x4.z, 0.f); x4.w = fmaxf(x4.w, 0.f); }
                    uint2 h2, l2;
                    bf16_split4(x4, h2, l2);
                    *reinterpret_cast<uint2*>(bhi + off) = h2;
                    *reinterpret_cast<uint2*>(blo + off) = l2;
                }
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(bar0 + 8u * s) : "memory");
            };
            float4 a0[4], a1[4];
            tap_state(0);
            fetch(a0);
            for (int it = 0; it < NIT; it += 2) {
                if (it + 1 < NIT) fetch(a1);
                stage(it, a0);
                ++g;
                if (it + 1 < NIT) {
                    if (it + 2 < NIT) fetch(a0);
                    stage(it + 1, a1);
                    ++g;
                }
            }
            // ---- epilogue.  TMEM lane = (half, co): warp quarter q = warp & 3 holds co = 32 (q & 1) + lane of half q >> 1 (0: W_hi rows, 1: W_lo rows);
            // warps 4 g .. 4 g + 3 take voxel columns 64 g .. 64 g + 63.  The stage buffers are idle now and carry the lo halves.
            conv_mbar_wait(baccfull, tcount & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const int q = warp & 3, half = q >> 1, cbase = 32 * (q & 1), colw = (warp >> 2) * 64;
            const unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16);
            float* xb = reinterpret_cast<float*>(conv_smem) + (size_t)((warp >> 2) * 2 + (q & 1)) * (32 * 33);   // one exchange block per (column half, co half)
            const float bco = sBias[cbase + lane];
            for (int c0 = 0; c0 < 64; c0 += 32) {
                float acc[32];
                tmem_ld32(taddr + (unsigned)(colw + c0), acc);
                if (half == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) xb[j * 33 + lane] = acc[j];
                }
                asm volatile("bar.sync 1, 512;\n" ::: "memory");
                if (half == 0) {
                    const long long vb = tile * kConvTileRows + colw + c0;
                    // rows of the frame window -> voxels of the volume: a block of 32 rows crosses at most one sample boundary
                    const long long win = YX * p.tn, nb0 = vb / win, edge = (nb0 + 1) * win;
                    const long long sh0 = (nb0 * (p.T - p.tn) + p.tlo) * YX, sh1 = sh0 + (long long)(p.T - p.tn) * YX;
                    if (p.resid) {                                            // all 32 skip-connection loads in flight together
                        float rs[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            rs[j] = (vb + j < p.NV) ? __ldg(p.resid + (size_t)(vb + j + (vb + j < edge ? sh0 : sh1)) * Cout + cbase + lane) : 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] = fmaf(p.resid_scale, p.resid_relu ? fmaxf(rs[j], 0.f) : rs[j], acc[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const long long v = vb + j;
                        if (v < p.NV) p.out[(size_t)(v + (v < edge ? sh0 : sh1)) * Cout + cbase + lane] = acc[j] + xb[j * 33 + lane] + bco;
                    }
                }
                asm volatile("bar.sync 1, 512;\n" ::: "memory");
            }
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(baccempty) : "memory");
        }
    } else if (lane == 0) {
        const unsigned idesc = umma_idesc_bf16(128, 256);
        const unsigned lboB = (unsigned)kC64LboB, lboW = 16u * 128u;
        unsigned g = 0, tcount = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
            conv_mbar_wait(baccempty, (tcount & 1u) ^ 1u);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            for (int it = 0; it < NIT; ++it, ++g) {
                const unsigned s = g % kConvStages, use = g / kConvStages;
                conv_mbar_wait(bar0 + 8u * s, use & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const unsigned sb = s0 + (unsigned)(s * kC64Stage), sw = sb + 2u * kC64BBytes;
#pragma unroll
                for (int j = 0; j < kConvKC / 16; ++j) {
                    const unsigned long long wd = umma_desc(sw + (unsigned)(2 * j) * lboW, lboW, 128u);
                    const unsigned long long xhi = umma_desc(sb + (unsigned)(2 * j) * lboB, lboB, 128u);
                    const unsigned long long xlo = umma_desc(sb + (unsigned)kC64BBytes + (unsigned)(2 * j) * lboB, lboB, 128u);
                    umma_f16(tmem_base, wd, xlo, idesc, (it > 0 || j > 0) ? 1u : 0u);
                    umma_f16(tmem_base, wd, xhi, idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(bempty + 8u * s) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(baccfull) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"(256) : "memory");
    }
}

}  // namespace dlsg
#endif
