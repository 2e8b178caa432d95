// Shifted-window attention kernels of the Swin denoiser block (sm_100a).
//
// Replaces, behind the C ABI of include/dlsg_b200.h, the reference's
// dl_cs/models/video_swin_transformer_mri_downsample.py:
//   forward_part1 :215-249   LayerNorm -> zero pad -> torch.roll -> window_partition        => swin_ln_partition_kernel
//   WindowAttention3D.forward :148-167   softmax(q*scale @ k^T + bias (+ mask)) @ v       => swin_attn_core_kernel
//   forward_part1 :240-248   window_reverse -> roll back -> crop                          => swin_window_reverse_kernel
// The two nn.Linear layers (qkv :146, proj :168) stay plain library GEMMs in the host module.
//
// Attention core.  One CTA = (one head, one window); K and V of that (window, head) are staged once in
// shared memory as fp16, each of the 7 warps walks 16-row query tiles and runs a FlashAttention-2
// style online softmax over key blocks of 64 with warp-level tensor-core MMAs (m16n8k16, fp16 in,
// fp32 accumulate).  What the reference materialises in HBM never leaves the chip:
//   * the [B_, heads, N, N] score matrix (193 MB per block at the cine size, 4-5 HBM round trips);
//   * the relative-position bias gather (table[index[N,N]] every forward, :152-155): since
//     index[i,j] = a(i) - a(j) + const with a(t) = d*(2Wh-1)(2Ww-1) + h*(2Ww-1) + w, the head's table
//     column sits in shared memory and one subtraction addresses it;
//   * the {0,-100} shift mask (compute_mask :342-355, 24 MB per shifted block): recomputed from a
//     per-token region id (one byte per token and window), or read densely for arbitrary masks.
// hd = 20 pads to 32 in the QK^T reduction and to 24 in the PV output; the kernel is bound by the
// softmax (exp + bias + mask per score), not by the tensor pipe.
#pragma once

#include "cuda_compat.h"

namespace dlsg {

struct SwinGeom {
    int B, D, H, W, C;          // un-padded feature volume [B, D, H, W, C]
    int wd, wh, ww;             // window actually used (after get_window_size clamping)
    int sd, sh, sw;             // cyclic shift actually used
    int Dp, Hp, Wp;             // padded extents (multiples of the window)
};

DLSG_DEVICE int swin_tokens(const SwinGeom& g) { return g.wd * g.wh * g.ww; }
DLSG_DEVICE int swin_windows(const SwinGeom& g) { return (g.Dp / g.wd) * (g.Hp / g.wh) * (g.Wp / g.ww); }

// source coordinates (in the un-padded, un-shifted volume) of token `tok` of window `win`;
// returns false for tokens that fall into the zero padding
DLSG_DEVICE bool swin_token_coords(const SwinGeom& g, int win, int tok, int& d, int& h, int& w) {
    const int nwh = g.Hp / g.wh, nww = g.Wp / g.ww;
    const int wi_w = win % nww, wi_h = (win / nww) % nwh, wi_d = win / (nww * nwh);
    const int tw = tok % g.ww, th = (tok / g.ww) % g.wh, td = tok / (g.ww * g.wh);
    // position in the shifted (rolled by -shift) padded volume, then undo the roll
    int pd = wi_d * g.wd + td + g.sd; if (pd >= g.Dp) pd -= g.Dp;
    int ph = wi_h * g.wh + th + g.sh; if (ph >= g.Hp) ph -= g.Hp;
    int pw = wi_w * g.ww + tw + g.sw; if (pw >= g.Wp) pw -= g.Wp;
    d = pd; h = ph; w = pw;
    return pd < g.D && ph < g.H && pw < g.W;
}

// ============================================================================
// LayerNorm (eps, biased variance) + zero pad + cyclic shift + window partition.
//   x [B, D, H, W, C] fp32  ->  xw [B*nW, N, C] fp32      one warp per output token
// ============================================================================
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(256)
swin_ln_partition_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                         float eps, float* __restrict__ xw, SwinGeom g) {
    const int lane = threadIdx.x & 31;
    const long long token = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int N = swin_tokens(g), nW = swin_windows(g);
    const long long total = (long long)g.B * nW * N;
    if (token >= total) return;
    const int tok = (int)(token % N);
    const long long bw = token / N;
    const int win = (int)(bw % nW), b = (int)(bw / nW);
    float* dst = xw + token * g.C;
    int d, h, w;
    if (!swin_token_coords(g, win, tok, d, h, w)) {          // padding happens after the norm (:219-225): exact zeros
        for (int c = lane; c < g.C; c += 32) dst[c] = 0.f;
        return;
    }
    const float* src = x + ((((long long)b * g.D + d) * g.H + h) * g.W + w) * g.C;
    float s = 0.f;
    for (int c = lane; c < g.C; c += 32) s += __ldg(src + c);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s / (float)g.C;
    float v = 0.f;
    for (int c = lane; c < g.C; c += 32) { const float t = __ldg(src + c) - mu; v = fmaf(t, t, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = 1.0f / sqrtf(v / (float)g.C + eps);
    for (int c = lane; c < g.C; c += 32)
        dst[c] = (gamma ? __ldg(gamma + c) : 1.f) * ((__ldg(src + c) - mu) * rstd) + (beta ? __ldg(beta + c) : 0.f);
}

// ============================================================================
// window_reverse + roll back + crop:  yw [B*nW, N, C] -> y [B, D, H, W, C]   one warp per output token
// ============================================================================
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(256)
swin_window_reverse_kernel(const float* __restrict__ yw, float* __restrict__ y, SwinGeom g) {
    const int lane = threadIdx.x & 31;
    const long long token = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int N = swin_tokens(g), nW = swin_windows(g);
    const long long total = (long long)g.B * nW * N;
    if (token >= total) return;
    const int tok = (int)(token % N);
    const long long bw = token / N;
    const int win = (int)(bw % nW), b = (int)(bw / nW);
    int d, h, w;
    if (!swin_token_coords(g, win, tok, d, h, w)) return;     // cropped away (:247-248)
    const float* src = yw + token * g.C;
    float* dst = y + ((((long long)b * g.D + d) * g.H + h) * g.W + w) * g.C;
    for (int c = lane; c < g.C; c += 32) dst[c] = __ldg(src + c);
}

// ============================================================================
// Attention core.
// ============================================================================
constexpr int kAttnWarps = 7;                  // 7 x 16 = 112 query rows per CTA: 448 = 4 x 112
constexpr int kAttnRows = 16 * kAttnWarps;
constexpr int kAttnKB = 64;                    // keys per online-softmax block
constexpr int kAttnHdPad = 32;                 // head_dim padded for the QK^T reduction (two k16 steps)
constexpr int kAttnKStride = 40;               // halves per K row in smem (80 B: conflict-free fragment loads)
constexpr float kLog2e = 1.4426950408889634f;

enum : int { SWIN_MASK_NONE = 0, SWIN_MASK_REGION = 1, SWIN_MASK_DENSE = 2 };

struct AttnParams {
    const float* qkv;        // [B_, N, 3, heads, hd] fp32 (output of the qkv Linear, untouched); NULL when the fp16 triple is given
    const half_t* q16;       // [B_, heads, N, 32] fp16, already multiplied by scale*log2(e)   (written by the fused qkv GEMM)
    const half_t* k16;       // [B_, heads, N, 32] fp16
    const half_t* vt16;      // v16 [B_, heads, N, 32] fp16 (same row layout as k16; column head_dim holds 1.0 when head_dim < 32)
    const float* table;      // relative_position_bias_table [(2Wd-1)(2Wh-1)(2Ww-1), heads] fp32
    const unsigned char* region;   // [nW, N] region ids (SWIN_MASK_REGION)
    const float* mask;       // [nW, N, N] additive mask (SWIN_MASK_DENSE)
    float* out;              // [B_, N, heads*hd] fp32
    int B_, N, heads, hd, nW;
    int mwd, mwh, mww;       // the MODULE's window size (bias table geometry, :115-129)
    float scale;
    int npad;                // N rounded up to a multiple of kAttnKB
    int vstride;             // unused (V shares the K row stride)
    int table_len;
};

inline size_t attn_smem_bytes(int npad, int vstride, int table_len, bool fast) {   // host-side sizing of the dynamic shared memory
    (void)vstride;
    return (size_t)npad * kAttnKStride * 2 * 2 +
           (size_t)(fast ? 2 * (table_len + 1) : table_len) * 4 + (size_t)npad * 4;
}

// FAST: (a) the bias table is stored as pairs (tab[i], tab[i-1]) so that the two adjacent key columns a lane owns cost
// one 8-byte load and the bias initialises the MMA accumulator; (b) row sums come out of the PV MMA through a column of
// ones appended to V (head dim hd is free when hd < 32).  Needs an even module window width and hd < 32.
template <int MASK_MODE, bool IN16, bool FAST>
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(32 * kAttnWarps, 2)
swin_attn_core_kernel(AttnParams p) {
    DLSG_DYN_SMEM(unsigned char, smem_raw);
    half_t* sK = reinterpret_cast<half_t*>(smem_raw);                                   // [npad][kAttnKStride]
    half_t* sV = sK + (size_t)p.npad * kAttnKStride;                                    // [npad][kAttnKStride] (row = key, like K)
    float* sTab = reinterpret_cast<float*>(sV + (size_t)p.npad * kAttnKStride);         // [table_len], * log2e
    int* sTok = reinterpret_cast<int*>(sTab + (FAST ? 2 * (p.table_len + 1) : p.table_len));   // [npad]: a(t) | region << 16

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int head = blockIdx.x, bwin = blockIdx.y;
    const int win = bwin % p.nW;
    const int N = p.N, hd = p.hd, C3 = 3 * p.heads * hd;
    const float* base = p.qkv + (size_t)bwin * N * C3 + head * hd;     // + tok*C3 (+ heads*hd for k, 2*heads*hd for v)
    const int koff = p.heads * hd, voff = 2 * p.heads * hd;

    // ---- stage K [tok][hd] and V^T [d][tok] as fp16 (zero padded), the head's bias column, the token table.
    // Lanes run over tokens (conflict-free V^T stores); each item is 4 consecutive head dims.
    const size_t bh = (size_t)bwin * p.heads + head;
    if (IN16) {
        // fp16 q / k / v^T straight from the fused qkv GEMM: plain 16-byte asynchronous copies (N % 8 == 0)
        const half_t* kg = p.k16 + bh * (size_t)N * 32;
        const half_t* vg = p.vt16 + bh * (size_t)N * 32;
        for (int i = tid; i < N * 4; i += blockDim.x) {
            const int tok = i >> 2, ch = i & 3;
            cp_async16(sK + (size_t)tok * kAttnKStride + 8 * ch, kg + (size_t)tok * 32 + 8 * ch);
        }
        for (int i = tid; i < N * 4; i += blockDim.x) {
            const int tok = i >> 2, ch = i & 3;
            cp_async16(sV + (size_t)tok * kAttnKStride + 8 * ch, vg + (size_t)tok * 32 + 8 * ch);
        }
        for (int i = tid; i < (p.npad - N) * kAttnHdPad; i += blockDim.x) {          // padded keys: zero rows
            const int tok = N + i / kAttnHdPad, dd = i % kAttnHdPad;
            sK[(size_t)tok * kAttnKStride + dd] = f2h(0.f);
            sV[(size_t)tok * kAttnKStride + dd] = f2h(0.f);
        }
    } else {
    const bool vec4 = (hd % 4 == 0) && (((size_t)C3 * 4) % 16 == 0) && ((((size_t)head * hd + koff) * 4) % 16 == 0) &&
                      ((((size_t)head * hd + voff) * 4) % 16 == 0) && ((reinterpret_cast<size_t>(p.qkv) % 16) == 0) &&
                      ((((size_t)N * C3) * 4) % 16 == 0);
    for (int i = tid; i < p.npad * (kAttnHdPad / 4); i += blockDim.x) {
        const int ch = i / p.npad, tok = i - ch * p.npad, dd = 4 * ch;
        float kv[4] = {0.f, 0.f, 0.f, 0.f}, vv[4] = {0.f, 0.f, 0.f, 0.f};
        if (tok < N && dd < hd) {
            const float* kp = base + (size_t)tok * C3 + koff + dd;
            const float* vp = base + (size_t)tok * C3 + voff + dd;
            if (vec4) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(kp)), b4 = __ldg(reinterpret_cast<const float4*>(vp));
                kv[0] = a.x; kv[1] = a.y; kv[2] = a.z; kv[3] = a.w;
                vv[0] = b4.x; vv[1] = b4.y; vv[2] = b4.z; vv[3] = b4.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (dd + e < hd) { kv[e] = __ldg(kp + e); vv[e] = __ldg(vp + e); }
            }
        }
        if (FAST && tok < N && dd <= hd && hd < dd + 4) vv[hd - dd] = 1.f;       // the column of ones (row sums through the PV MMA)
        unsigned* kdst = reinterpret_cast<unsigned*>(sK + (size_t)tok * kAttnKStride + dd);
        kdst[0] = pack_h2(kv[0], kv[1]);
        kdst[1] = pack_h2(kv[2], kv[3]);
        unsigned* vdst = reinterpret_cast<unsigned*>(sV + (size_t)tok * kAttnKStride + dd);
        vdst[0] = pack_h2(vv[0], vv[1]);
        vdst[1] = pack_h2(vv[2], vv[3]);
    }
    }
    if (FAST) {                                              // sTab as float2 pairs: [i] = (tab[i], tab[i-1])
        float2* sTabP = reinterpret_cast<float2*>(sTab);
        for (int i = tid; i <= p.table_len; i += blockDim.x) {
            const float hi = (i < p.table_len) ? __ldg(p.table + (size_t)i * p.heads + head) * kLog2e : 0.f;
            const float lo = (i > 0) ? __ldg(p.table + (size_t)(i - 1) * p.heads + head) * kLog2e : 0.f;
            sTabP[i] = make_float2(hi, lo);
        }
    } else {
        for (int i = tid; i < p.table_len; i += blockDim.x) sTab[i] = __ldg(p.table + (size_t)i * p.heads + head) * kLog2e;
    }
    const int c_h = 2 * p.mww - 1, c_d = (2 * p.mwh - 1) * c_h;
    int mixed_local = 0;
    for (int tok = tid; tok < p.npad; tok += blockDim.x) {
        const int tw = tok % p.mww, th = (tok / p.mww) % p.mwh, td = tok / (p.mww * p.mwh);
        const int a = (tok < N) ? td * c_d + th * c_h + tw : 0;   // padded tokens stay inside the table
        int r = 0;
        if (MASK_MODE == SWIN_MASK_REGION && tok < N) r = p.region[(size_t)win * N + tok];
        sTok[tok] = a | (r << 16);
        if (MASK_MODE == SWIN_MASK_REGION && tok < N && r != (int)p.region[(size_t)win * N]) mixed_local = 1;
    }
    if (IN16) cp_async_wait_all();
    // interior windows of a shifted block hold a single region: their mask is all zero and the test is skipped
    const bool masked = (MASK_MODE == SWIN_MASK_REGION) ? (__syncthreads_or(mixed_local) != 0) : (MASK_MODE == SWIN_MASK_DENSE);
    if (MASK_MODE != SWIN_MASK_REGION) __syncthreads();

    // each warp walks the query tiles of this (window, head): rows warp*16 + k*112
    for (int row0 = warp * 16; row0 < N; row0 += kAttnRows) {
    const int r_lo = row0 + g, r_hi = row0 + g + 8;
    // ---- Q fragments (A operand, 16 x 32): q * scale * log2e, fp16
    const float qs = p.scale * kLog2e;
    unsigned qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int half_i = 0; half_i < 4; ++half_i) {
            const int r = (half_i & 1) ? r_hi : r_lo;
            const int c = ks * 16 + 2 * t4 + ((half_i & 2) ? 8 : 0);
            if (IN16) {
                qa[ks][half_i] = (r < N) ? __ldg(reinterpret_cast<const unsigned*>(p.q16 + (bh * (size_t)N + r) * 32 + c)) : 0u;
                continue;
            }
            float v0 = 0.f, v1 = 0.f;
            if (r < N) {
                if (c < hd) v0 = __ldg(base + (size_t)r * C3 + c) * qs;
                if (c + 1 < hd) v1 = __ldg(base + (size_t)r * C3 + c + 1) * qs;
            }
            qa[ks][half_i] = pack_h2(v0, v1);
        }
    }
    const int center = (p.mwd - 1) * c_d + (p.mwh - 1) * c_h + (p.mww - 1);   // index offset of relative (0,0,0)
    const int ti_lo = sTok[min(r_lo, p.npad - 1)], ti_hi = sTok[min(r_hi, p.npad - 1)];
    const int a_lo = (ti_lo & 0xffff) + center, a_hi = (ti_hi & 0xffff) + center;
    const int rg_lo = ti_lo >> 16, rg_hi = ti_hi >> 16;
    const float* mrow_lo = nullptr;
    const float* mrow_hi = nullptr;
    if (MASK_MODE == SWIN_MASK_DENSE) {
        mrow_lo = p.mask + ((size_t)win * N + min(r_lo, N - 1)) * N;
        mrow_hi = p.mask + ((size_t)win * N + min(r_hi, N - 1)) * N;
    }

    const half_t* kfrag = sK + (lane & 7) * kAttnKStride + 8 * (lane >> 3);                        // this lane's ldmatrix row, key 0
    const half_t* vfrag = sV + (((lane >> 3) & 1) * 8 + (lane & 7)) * kAttnKStride + 8 * (lane >> 4);
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
    float o[4][4];                                         // 16 x 32 output accumulator (4 n8 tiles)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    const float kMasked = -100.0f * kLog2e;

    for (int kb = 0; kb < p.npad; kb += kAttnKB) {
        const bool ragged = kb + kAttnKB > N;               // only the last key block can hold padded keys
        // ---- S = bias + Q K^T for 64 keys: 8 n8 tiles, 2 k16 steps (+ mask), running max
        float s[8][4];
        float bm_lo = -INFINITY, bm_hi = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int j0 = kb + nt * 8 + 2 * t4;
            int tj0, tj1;
            if (FAST) {
                const int2 tt = *reinterpret_cast<const int2*>(sTok + j0);       // j0 is even: 8-byte aligned
                tj0 = tt.x; tj1 = tt.y;
                const float2 b_lo = reinterpret_cast<const float2*>(sTab)[a_lo - (tj0 & 0xffff)];
                const float2 b_hi = reinterpret_cast<const float2*>(sTab)[a_hi - (tj0 & 0xffff)];
                s[nt][0] = b_lo.x; s[nt][1] = b_lo.y; s[nt][2] = b_hi.x; s[nt][3] = b_hi.y;
            } else {
                tj0 = sTok[j0]; tj1 = sTok[j0 + 1];
                s[nt][0] = sTab[a_lo - (tj0 & 0xffff)]; s[nt][1] = sTab[a_lo - (tj1 & 0xffff)];
                s[nt][2] = sTab[a_hi - (tj0 & 0xffff)]; s[nt][3] = sTab[a_hi - (tj1 & 0xffff)];
            }
            unsigned kf[4];                                  // (b0, b1) of both k16 steps in one ldmatrix.x4
            ldmatrix_x4(kf, kfrag + (kb + nt * 8) * kAttnKStride);
            mma_m16n8k16_f16(s[nt], qa[0], kf[0], kf[1]);
            mma_m16n8k16_f16(s[nt], qa[1], kf[2], kf[3]);
            if (MASK_MODE == SWIN_MASK_REGION) {
                if (masked) {
                    const int r0 = tj0 >> 16, r1 = tj1 >> 16;
                    if (r0 != rg_lo) s[nt][0] += kMasked;
                    if (r1 != rg_lo) s[nt][1] += kMasked;
                    if (r0 != rg_hi) s[nt][2] += kMasked;
                    if (r1 != rg_hi) s[nt][3] += kMasked;
                }
            } else if (MASK_MODE == SWIN_MASK_DENSE) {
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (j0 + e < N) {
                        s[nt][e] = fmaf(__ldg(mrow_lo + j0 + e), kLog2e, s[nt][e]);
                        s[nt][2 + e] = fmaf(__ldg(mrow_hi + j0 + e), kLog2e, s[nt][2 + e]);
                    }
            }
            // padded keys: with the ones column (FAST) their V row is all zero, so they add nothing to the numerator nor to
            // the row sum whatever their score; only the explicit-sum variant has to push them to -inf (last block only)
            if (!FAST && ragged) {
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (j0 + e >= N) { s[nt][e] = -INFINITY; s[nt][2 + e] = -INFINITY; }
            }
            bm_lo = fmaxf(bm_lo, fmaxf(s[nt][0], s[nt][1]));
            bm_hi = fmaxf(bm_hi, fmaxf(s[nt][2], s[nt][3]));
        }
        bm_lo = fmaxf(bm_lo, __shfl_xor_sync(0xffffffffu, bm_lo, 1));
        bm_lo = fmaxf(bm_lo, __shfl_xor_sync(0xffffffffu, bm_lo, 2));
        bm_hi = fmaxf(bm_hi, __shfl_xor_sync(0xffffffffu, bm_hi, 1));
        bm_hi = fmaxf(bm_hi, __shfl_xor_sync(0xffffffffu, bm_hi, 2));
        const float mn_lo = fmaxf(m_lo, bm_lo), mn_hi = fmaxf(m_hi, bm_hi);   // finite: every block holds >= 1 real key
        const float f_lo = fast_exp2(m_lo - mn_lo), f_hi = fast_exp2(m_hi - mn_hi);
        m_lo = mn_lo; m_hi = mn_hi;
        l_lo *= f_lo; l_hi *= f_hi;
#pragma unroll
        for (int i = 0; i < 4; ++i) { o[i][0] *= f_lo; o[i][1] *= f_lo; o[i][2] *= f_hi; o[i][3] *= f_hi; }
        // ---- P = exp2(S - m), row sums, P -> fp16 A fragments, O += P V
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {                    // 16 keys per step = 2 n8 tiles of S
            unsigned pa[4];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int nt = 2 * kk + h2;
                const float p0 = fast_exp2(s[nt][0] - m_lo), p1 = fast_exp2(s[nt][1] - m_lo);
                const float p2 = fast_exp2(s[nt][2] - m_hi), p3 = fast_exp2(s[nt][3] - m_hi);
                if (!FAST) { l_lo += p0 + p1; l_hi += p2 + p3; }
                pa[2 * h2] = pack_h2(p0, p1);
                pa[2 * h2 + 1] = pack_h2(p2, p3);
            }
            // V fragments: transposing ldmatrix on [key][dim] rows; one x4 = (b0, b1) of two output tiles
            const half_t* vrow = vfrag + (kb + kk * 16) * kAttnKStride;
#pragma unroll
            for (int dp = 0; dp < 2; ++dp) {                // output dims dp*16 .. dp*16+15
                unsigned vf[4];
                ldmatrix_x4_trans(vf, vrow + dp * 16);
                mma_m16n8k16_f16(o[2 * dp], pa, vf[0], vf[1]);
                mma_m16n8k16_f16(o[2 * dp + 1], pa, vf[2], vf[3]);
            }
        }
    }
    // ---- normalise and store  out[bwin, row, head*hd + d]
    if (FAST) {                                              // the row sums sit in output column hd (lanes with 2*t4 == hd % 8 ...)
        float sl = 0.f, sh = 0.f;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                if (dn * 8 + 2 * t4 + e == hd) { sl = o[dn][e]; sh = o[dn][2 + e]; }
        const int src = (lane & ~3) | ((hd & 7) >> 1);
        l_lo = __shfl_sync(0xffffffffu, sl, src);
        l_hi = __shfl_sync(0xffffffffu, sh, src);
    } else {
        l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
        l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    }
    const float inv_lo = 1.0f / l_lo, inv_hi = 1.0f / l_hi;
    float* obase = p.out + (size_t)bwin * N * (p.heads * hd) + head * hd;
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int dcol = dn * 8 + 2 * t4 + e;
            if (dcol < hd) {
                if (r_lo < N) obase[(size_t)r_lo * (p.heads * hd) + dcol] = o[dn][e] * inv_lo;
                if (r_hi < N) obase[(size_t)r_hi * (p.heads * hd) + dcol] = o[dn][2 + e] * inv_hi;
            }
        }
    }
    }   // query tiles
}

}  // namespace dlsg
