// SENSE forward / adjoint / normal-operator kernels (sm_100a).
//
// Replaces, behind the C-ABI of include/dlsg_b200.h, the reference's
//   SenseModel._forward_op   dl_cs/mri/transforms.py:92-98   k = W * F2( sum_e S_ec * x_e )
//   SenseModel._adjoint_op   dl_cs/mri/transforms.py:84-90   x_e = sum_c conj(S_ec) * F2^-1( W * k_c )
//   DC update                dl_cs/models/unrolledswin.py:109 (== unrolled.py:104)
//                            x <- x + s * (A^H A x - A^H y)
// with F2 the un-centred 2-D DFT, norm='ortho' (transforms.py:18,39,41).
//
// Data layout in HBM (all contiguous, X fastest, complex64 = interleaved fp32):
//   x     [B, E, T, Y, X]      image series per ESPIRiT map set
//   maps  [B, E, C, Y, X]      coil sensitivities (the reference's singleton dim 3 dropped)
//   k     [B, C, T, Y, X]      multi-coil k-space
//   W     real fp32, addressed through 5 element strides over (b, c, t, y, x); 0 = broadcast
//   tmp   [Bc, C, Tc, Y, X]    row-transformed intermediate of one chunk (sized to stay in L2)
//
// Decomposition.  Every axis length N = 32*R (R in {1,2,3,4,5,6,8,10}) is split
// as a radix-R DFT and a radix-32 DFT, both done entirely in registers by one
// thread, with one shared-memory exchange between them:
//   "small-first":  n = n1 + 32*n2, k = R*k1 + k2
//       X[R*k1+k2] = sum_n1 W32^(n1 k1) [ WN^(n1 k2) sum_n2 WR^(n2 k2) x[n1+32 n2] ]
//   "32-first" (only inside the fused normal operator, for the way back):
//       n = R*n1 + n2, k = k1 + 32*k2
//       X[k1+32*k2] = sum_n2 WR^(n2 k2) [ WN^(n2 k1) sum_n1 W32^(n1 k1) x[R*n1+n2] ]
// Row pass: one warp owns a batch of NB = 32/R rows so that the radix-32 stage has
// one transform per lane.  Column pass: one CTA owns a 32-column slab of one image.
// Axis lengths that are not 32*R fall back to the direct-DFT kernels at the end of
// this file (still CUDA, O(N^2) per line; correctness path for odd/ragged sizes).
#pragma once

#include "cuda_compat.h"
#include "fft_radix.cuh"
#include "fft_packed.cuh"

// Occupancy knobs (resident CTAs per SM the register allocator must allow); tuned on the B200, see DESIGN.md.
#ifndef DLSG_ROWS_MINB
#define DLSG_ROWS_MINB 3
#endif
#ifndef DLSG_COLS_TARGET_THREADS
#define DLSG_COLS_TARGET_THREADS 448
#endif
// DLSG_PIN_LOADS: keep the compiler from hoisting the next batch's global loads above the radix-32 phase
#ifdef DLSG_PIN_LOADS
#define DLSG_LOAD_FENCE() asm volatile("" ::: "memory")
#else
#define DLSG_LOAD_FENCE() ((void)0)
#endif

namespace dlsg {

// In-register DFTs of the kernels below: the packed-pair butterflies (FFMA2 / FADD2, fft_packed.cuh) unless
// DLSG_SCALAR_DFT selects the scalar ones of fft_radix.cuh (kept for A/B measurements).
#ifdef DLSG_SCALAR_DFT
template <int N, int DIR> using XDft = Dft<N, DIR>;
#else
template <int N, int DIR> using XDft = PDft<N, DIR>;
#endif

struct SenseDims {
    int B, E, C, T, Y, X;   // B = slices in this launch (chunk), the rest full extents
    int t_begin, t_count;   // frames [t_begin, t_begin + t_count) are processed; tmp has t_count frames
};
struct WeightStrides { long long b, c, t, y, x; };

enum : int { MODE_FWD = 0, MODE_ADJ = 1, MODE_NORMAL = 2 };
enum : int { EPI_PLAIN = 0, EPI_DC = 1, EPI_REG = 2 };   // REG: out = acc + step * xin  (A^H A x + lambda x, the MoDL normal equations)

DLSG_DEVICE size_t idx_x(const SenseDims& d, int b, int e, int t, int y) {
    return ((((size_t)b * d.E + e) * d.T + t) * d.Y + y) * (size_t)d.X;
}
DLSG_DEVICE size_t idx_maps(const SenseDims& d, int b, int e, int c, int y) {
    return ((((size_t)b * d.E + e) * d.C + c) * d.Y + y) * (size_t)d.X;
}
DLSG_DEVICE size_t idx_k(const SenseDims& d, int b, int c, int t, int y) {
    return ((((size_t)b * d.C + c) * d.T + t) * d.Y + y) * (size_t)d.X;
}
DLSG_DEVICE size_t idx_tmp(const SenseDims& d, int b, int c, int t, int y) {   // t absolute
    return ((((size_t)b * d.C + c) * d.t_count + (t - d.t_begin)) * d.Y + y) * (size_t)d.X;
}
DLSG_DEVICE float load_w(const float* w, const WeightStrides& ws, int b, int c, int t, int y, int x) {
    return w ? __ldg(w + (b * ws.b + c * ws.c + t * ws.t + y * ws.y + x * ws.x)) : 1.0f;
}

// ============================================================================
// Row pass (axis X), fast path: X = 32*RX.
//
// One CTA (4 warps) owns (slice b, image row y, a chunk of frames).  The coil-sensitivity
// rows maps[b, :, c-group, y, :] are staged once per CTA in shared memory (cp.async) and
// reused for every frame.  Each warp owns a contiguous sub-range of frames and walks the
// flattened (frame, coil) items, coil fastest, in batches of NB = 32/RX rows, so that the
// radix-32 stage has one transform per lane:
//   phase A  lane = x mod 32: radix-RX DFT over x = lane + 32 j, twiddle, to view 1
//   phase B  lane = (row, k2): radix-32 DFT over the 32 lanes' values, to view 2 (natural kx order)
//   phase C  coalesced copy-out (forward) / conj-map multiply-accumulate over coils (adjoint)
// ============================================================================
template <int RX>
struct RowCfg {
    static constexpr int X = 32 * RX;
    static constexpr int NB = 32 / RX;          // rows per warp batch
    static constexpr int S1 = 34;               // view 1: [row*RX + k2][lane], padded for LDS.128
    static constexpr int S2 = 33 * RX;          // view 2: [row][x] natural order, stride = RX mod 16
    static constexpr int BUF = 32 * 34;         // complex elements of smem per warp
    static constexpr int WARPS = 4;
    // forward: exchange buffers + staged maps;  adjoint (staging): + raw-row staging and DC-row buffers per warp
    static size_t smem_bytes(int E, int cgroup, bool staging) {
        return ((size_t)WARPS * (BUF + (staging ? NB * X + 2 * E * X : 0)) + (size_t)E * cgroup * X) * sizeof(float2);
    }
};

struct RowPlan { int tchunk, nchunks, cgroup; };

// lane twiddles exp(DIR*2*pi*i*lane*k2/X), k2 = 0..RX-1
template <int RX, int DIR>
DLSG_DEVICE void lane_twiddles(int lane, float2* tw) {
#pragma unroll
    for (int k2 = 0; k2 < RX; ++k2) {
        int m = (lane * k2) % (32 * RX);            // exact integer reduction, then |angle| <= pi
        if (2 * m > 32 * RX) m -= 32 * RX;
        float s, c;
        sincospif((float)(2 * m) / (float)(32 * RX), &s, &c);
        tw[k2] = make_float2(c, DIR < 0 ? -s : s);
    }
}

// Phase B of the row pass: radix-32 across the batch, view 1 -> view 2.
template <int RX, int DIR>
DLSG_DEVICE void row_radix32(float2* buf, int lane, int nrows) {
    using Cfg = RowCfg<RX>;
    const bool act = lane < nrows * RX;
    float2 u[32];
    if (act) {
        const float4* p = reinterpret_cast<const float4*>(buf + lane * Cfg::S1);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float4 q = p[i];
            u[2 * i] = make_float2(q.x, q.y);
            u[2 * i + 1] = make_float2(q.z, q.w);
        }
    }
    __syncwarp();
    if (act) {
        XDft<32, DIR>::run(u);
        const int row = lane / RX, k2 = lane - row * RX;
        float2* o = buf + row * Cfg::S2 + k2;
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) o[RX * k1] = u[k1];
    }
    __syncwarp();
}

// Stage maps[b, 0..E, cg0..cg0+cgn, y, :] into shared memory as [e][cl][X]; all threads of the CTA call it.
template <int RX>
DLSG_DEVICE void stage_maps(float2* ms, const float2* __restrict__ maps, const SenseDims& d,
                            int b, int y, int cg0, int cgn, int cgroup) {
    constexpr int X = 32 * RX;
    constexpr int CH = X / 2;                    // 16-byte chunks per row
    const int total = d.E * cgn * CH;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int ch = i % CH, rc = i / CH;
        const int cl = rc % cgn, e = rc / cgn;
        cp_async16(ms + ((size_t)(e * cgroup + cl) * X + 2 * ch),
                   maps + idx_maps(d, b, e, cg0 + cl, y) + 2 * ch);
    }
    cp_async_wait_all();
}

// Forward rows:  tmp[b,c,t,y,:] = DFT_X( sum_e maps[b,e,c,y,:] * x[b,e,t,y,:] )   (unnormalised)
// EREG: number of map sets whose x row is cached in registers (1 or 2); 0 = any E, x re-read per row.
template <int RX, int EREG>
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(128, DLSG_ROWS_MINB)
sense_rows_fwd_kernel(const float2* __restrict__ x, const float2* __restrict__ maps,
                      float2* __restrict__ tmp, SenseDims d, RowPlan plan) {
    using Cfg = RowCfg<RX>;
    constexpr int X = Cfg::X, NB = Cfg::NB;
    DLSG_DYN_SMEM(float2, smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* buf = smem + warp * Cfg::BUF;
    float2* ms = smem + Cfg::WARPS * Cfg::BUF;   // [e][cl][X]
    long long unit = blockIdx.x;
    const int chunk = (int)(unit % plan.nchunks); unit /= plan.nchunks;
    const int y = (int)(unit % d.Y), b = (int)(unit / d.Y);
    const int t0 = d.t_begin + chunk * plan.tchunk;
    const int t1 = min(d.t_begin + d.t_count, t0 + plan.tchunk);
    const int per_warp = (t1 - t0 + Cfg::WARPS - 1) / Cfg::WARPS;
    const int tw0 = min(t1, t0 + warp * per_warp), tw1 = min(t1, tw0 + per_warp);

    float2 tw[RX];
    lane_twiddles<RX, -1>(lane, tw);
    constexpr int ER = EREG > 0 ? EREG : 1;
    float2 xr[ER][RX], xn[ER][RX];

    auto load_x = [&](float2 (&dst)[ER][RX], int t) {
#pragma unroll
        for (int e = 0; e < ER; ++e) {
            const float2* xs = x + idx_x(d, b, e, t, y) + lane;
#pragma unroll
            for (int j = 0; j < RX; ++j) dst[e][j] = __ldg(xs + 32 * j);
        }
    };

    for (int cg0 = 0; cg0 < d.C; cg0 += plan.cgroup) {
        const int cgn = min(plan.cgroup, d.C - cg0);
        if (maps) {
            __syncthreads();                     // previous group's rows are no longer read
            stage_maps<RX>(ms, maps, d, b, y, cg0, cgn, plan.cgroup);
            __syncthreads();
        }
        const int nitems = (tw1 - tw0) * cgn;
        int cur_t = -1;
        if (EREG > 0 && nitems > 0) load_x(xn, tw0);
        for (int it0 = 0; it0 < nitems; it0 += NB) {
            const int nrows = min(NB, nitems - it0);
            int tl = it0 / cgn, cl = it0 - tl * cgn;
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                if (r >= nrows) break;
                const int t = tw0 + tl;
                if (EREG > 0 && t != cur_t) {    // new frame: take the prefetched row, prefetch the next one
#pragma unroll
                    for (int e = 0; e < ER; ++e)
#pragma unroll
                        for (int j = 0; j < RX; ++j) xr[e][j] = xn[e][j];
                    cur_t = t;
                    if (t + 1 < tw1) load_x(xn, t + 1);
                }
                float2 v[RX];
#pragma unroll
                for (int j = 0; j < RX; ++j) v[j] = make_float2(0.f, 0.f);
                if (EREG > 0) {
#pragma unroll
                    for (int e = 0; e < ER; ++e) {
                        if (maps) {
                            const float2* m = ms + (size_t)(e * plan.cgroup + cl) * X + lane;
#pragma unroll
                            for (int j = 0; j < RX; ++j) v[j] = cfma(m[32 * j], xr[e][j], v[j]);
                        } else {                 // maps == NULL: unit sensitivities (plain FFT entry point)
#pragma unroll
                            for (int j = 0; j < RX; ++j) v[j] = cadd(v[j], xr[e][j]);
                        }
                    }
                } else {
                    for (int e = 0; e < d.E; ++e) {
                        const float2* xs = x + idx_x(d, b, e, t, y) + lane;
                        const float2* m = ms + (size_t)(e * plan.cgroup + cl) * X + lane;
#pragma unroll
                        for (int j = 0; j < RX; ++j) {
                            const float2 xv = __ldg(xs + 32 * j);
                            v[j] = maps ? cfma(m[32 * j], xv, v[j]) : cadd(v[j], xv);
                        }
                    }
                }
                XDft<RX, -1>::run(v);
                float2* o = buf + (r * RX) * Cfg::S1 + lane;
#pragma unroll
                for (int k2 = 0; k2 < RX; ++k2) o[k2 * Cfg::S1] = (k2 == 0) ? v[0] : cmul(v[k2], tw[k2]);
                if (++cl == cgn) { cl = 0; ++tl; }
            }
            __syncwarp();
            row_radix32<RX, -1>(buf, lane, nrows);
            tl = it0 / cgn; cl = it0 - tl * cgn;
#pragma unroll
            for (int r = 0; r < NB; ++r) {
                if (r >= nrows) break;
                float2* dst = tmp + idx_tmp(d, b, cg0 + cl, tw0 + tl, y) + lane;
                const float2* sv = buf + r * Cfg::S2 + lane;
#pragma unroll
                for (int j = 0; j < RX; ++j) dst[32 * j] = sv[32 * j];
                if (++cl == cgn) { cl = 0; ++tl; }
            }
            __syncwarp();
        }
    }
}

// Adjoint rows:  acc_e[b,t,y,:] = sum_c conj(maps[b,e,c,y,:]) * IDFT_X( tmp[b,c,t,y,:] )
//   EPI_PLAIN: out = acc
//   EPI_DC   : out = xin + step * (acc - aty)          (unrolledswin.py:109)
// EMAX accumulators live in registers; E > EMAX is handled by e-groups (e0 .. e0+EMAX).
// With several coil groups the partial sums go through `out` (each element is owned by one lane).
template <int RX, int EMAX>
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(128, DLSG_ROWS_MINB)
sense_rows_adj_kernel(const float2* __restrict__ tmp, const float2* __restrict__ maps,
                      float2* out, const float2* __restrict__ xin,
                      const float2* __restrict__ aty, float step_host, const float* __restrict__ step_dev,
                      int epi, SenseDims d, RowPlan plan, int e0) {
    using Cfg = RowCfg<RX>;
    constexpr int X = Cfg::X, NB = Cfg::NB;
    DLSG_DYN_SMEM(float2, smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* buf = smem + warp * Cfg::BUF;
    float2* stg = smem + Cfg::WARPS * Cfg::BUF + warp * (NB * X);     // [r][X] raw rows of the next batch
    float2* dcb = smem + Cfg::WARPS * (Cfg::BUF + NB * X) + warp * (2 * EMAX * X);   // [xin|aty][e][X] of the frame being summed
    float2* ms = smem + Cfg::WARPS * (Cfg::BUF + NB * X + 2 * EMAX * X);             // [e][cl][X], e relative to e0
    long long unit = blockIdx.x;
    const int chunk = (int)(unit % plan.nchunks); unit /= plan.nchunks;
    const int y = (int)(unit % d.Y), b = (int)(unit / d.Y);
    const int t0 = d.t_begin + chunk * plan.tchunk;
    const int t1 = min(d.t_begin + d.t_count, t0 + plan.tchunk);
    const int per_warp = (t1 - t0 + Cfg::WARPS - 1) / Cfg::WARPS;
    const int tw0 = min(t1, t0 + warp * per_warp), tw1 = min(t1, tw0 + per_warp);
    const int ne = min(EMAX, d.E - e0);
    const float step = step_dev ? __ldg(step_dev) : step_host;

    float2 tw[RX];
    lane_twiddles<RX, +1>(lane, tw);

    float2 acc[EMAX][RX];
#pragma unroll
    for (int e = 0; e < EMAX; ++e)
#pragma unroll
        for (int j = 0; j < RX; ++j) acc[e][j] = make_float2(0.f, 0.f);

    for (int cg0 = 0; cg0 < d.C; cg0 += plan.cgroup) {
        const int cgn = min(plan.cgroup, d.C - cg0);
        const bool first_group = (cg0 == 0), last_group = (cg0 + cgn >= d.C);
        if (maps) {
            __syncthreads();
            SenseDims de = d;                    // stage only this kernel's map sets: shift e by e0
            constexpr int CH = X / 2;
            const int total = ne * cgn * CH;
            for (int i = threadIdx.x; i < total; i += blockDim.x) {
                const int ch = i % CH, rc = i / CH;
                const int cl = rc % cgn, e = rc / cgn;
                cp_async16(ms + ((size_t)(e * plan.cgroup + cl) * X + 2 * ch),
                           maps + idx_maps(de, b, e0 + e, cg0 + cl, y) + 2 * ch);
            }
            cp_async_wait_all();
            __syncthreads();
        }

        // x and A^H y rows of frame t for the DC epilogue, fetched asynchronously when the frame is started
        auto stage_dc = [&](int t) {
            if (last_group && epi != EPI_PLAIN) {
                __syncwarp();                    // the previous frame's epilogue has read the buffer
                for (int e = 0; e < ne; ++e) {
                    const size_t base = idx_x(d, b, e0 + e, t, y);
#pragma unroll
                    for (int m = 0; m < (X / 2 + 31) / 32; ++m) {
                        const int q = lane + 32 * m;
                        if ((X / 2) % 32 == 0 || q < X / 2) {
                            cp_async16(dcb + e * X + 2 * q, xin + base + 2 * q);
                            if (epi == EPI_DC) cp_async16(dcb + (EMAX + e) * X + 2 * q, aty + base + 2 * q);
                        }
                    }
                }
                cp_async_commit();
            }
        };
        auto flush = [&](int t) {
            if (last_group && epi != EPI_PLAIN) { cp_async_wait_group0(); __syncwarp(); }
#pragma unroll
            for (int e = 0; e < EMAX; ++e) {
                if (e < ne) {
                    const size_t base = idx_x(d, b, e0 + e, t, y) + lane;
#pragma unroll
                    for (int j = 0; j < RX; ++j) {
                        float2 r = acc[e][j];
                        if (!first_group) r = cadd(r, out[base + 32 * j]);
                        if (last_group && epi == EPI_DC) {
                            const float2 xv = dcb[e * X + lane + 32 * j], av = dcb[(EMAX + e) * X + lane + 32 * j];
                            r = make_float2(fmaf(step, r.x - av.x, xv.x), fmaf(step, r.y - av.y, xv.y));
                        } else if (last_group && epi == EPI_REG) {
                            const float2 xv = dcb[e * X + lane + 32 * j];
                            r = make_float2(fmaf(step, xv.x, r.x), fmaf(step, xv.y, r.y));
                        }
                        out[base + 32 * j] = r;
                        acc[e][j] = make_float2(0.f, 0.f);
                    }
                }
            }
        };

        const int nitems = (tw1 - tw0) * cgn;
        int cur_t = tw0;
        // asynchronous staging of one batch of tmp rows (natural order, [r][X]) into this warp's buffer
        auto stage_batch = [&](int it0) {
            const int nrows = min(NB, nitems - it0);
            int tl = it0 / cgn, cl = it0 - tl * cgn;
            for (int r = 0; r < nrows; ++r) {
                const float2* src = tmp + idx_tmp(d, b, cg0 + cl, tw0 + tl, y);
                float2* dstp = stg + r * X;
#pragma unroll
                for (int m = 0; m < (X / 2 + 31) / 32; ++m) {
                    const int q = lane + 32 * m;
                    if ((X / 2) % 32 == 0 || q < X / 2) cp_async16(dstp + 2 * q, src + 2 * q);
                }
                if (++cl == cgn) { cl = 0; ++tl; }
            }
            cp_async_commit();
        };
        if (nitems > 0) { stage_dc(tw0); stage_batch(0); }
        for (int it0 = 0; it0 < nitems; it0 += NB) {
            const int nrows = min(NB, nitems - it0);
            const int tl0 = it0 / cgn, cl0 = it0 - tl0 * cgn;
            cp_async_wait_group0();
            __syncwarp();
#pragma unroll
            for (int r = 0; r < NB; ++r) {       // rows are independent: unrolled so that their chains interleave
                if (r < nrows) {
                    float2 v[RX];
#pragma unroll
                    for (int j = 0; j < RX; ++j) v[j] = stg[r * X + lane + 32 * j];
                    XDft<RX, +1>::run(v);
                    float2* o = buf + (r * RX) * Cfg::S1 + lane;
#pragma unroll
                    for (int k2 = 0; k2 < RX; ++k2) o[k2 * Cfg::S1] = (k2 == 0) ? v[0] : cmul(v[k2], tw[k2]);
                }
            }
            __syncwarp();                        // staging buffer consumed, view 1 complete
            if (it0 + NB < nitems) stage_batch(it0 + NB);   // lands while phases B and C run
            row_radix32<RX, +1>(buf, lane, nrows);
            int tl = tl0, cl = cl0;
            const float2* sv = buf + lane;
            const float2* mrow = ms + cl0 * X + lane;
            const int estride = plan.cgroup * X;
            for (int r = 0; r < nrows; ++r, sv += Cfg::S2) {
                const int t = tw0 + tl;
                if (t != cur_t) { flush(cur_t); cur_t = t; stage_dc(t); }
                float2 v[RX];
#pragma unroll
                for (int j = 0; j < RX; ++j) v[j] = sv[32 * j];
#pragma unroll
                for (int e = 0; e < EMAX; ++e) {
                    if (e < ne) {
                        if (maps) {
                            const float2* m = mrow + e * estride;
#pragma unroll
                            for (int j = 0; j < RX; ++j) acc[e][j] = cfmac(v[j], m[32 * j], acc[e][j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < RX; ++j) acc[e][j] = cadd(acc[e][j], v[j]);
                        }
                    }
                }
                mrow += X;
                if (++cl == cgn) { cl = 0; ++tl; mrow = ms + lane; }
            }
            __syncwarp();
        }
        if (nitems > 0) flush(cur_t);
    }
}

// ============================================================================
// Column pass (axis Y), fast path: Y = 32*RY.  One CTA = 32 columns of one image.
//   MODE_FWD   : k   = W * scale * DFT_Y(tmp)            (weights applied at the store)
//   MODE_ADJ   : tmp = IDFT_Y(W * scale * k)             (weights applied at the load)
//   MODE_NORMAL: tmp = IDFT_Y((W*scale)^2 * DFT_Y(tmp))  in place
// ============================================================================
template <int RY>
struct ColCfg {
    static constexpr int Y = 32 * RY;
    static constexpr int SW = (RY <= 6) ? 32 : 16;      // slab width in columns
    static constexpr int NT = SW * RY;                  // threads per CTA = radix-32 transforms per slab
    static constexpr int ITEMS = (32 + RY - 1) / RY;    // phase-A (l, column) items per thread
    static constexpr int MINB = (DLSG_COLS_TARGET_THREADS / NT) > 0 ? (DLSG_COLS_TARGET_THREADS / NT) : 1;
    static constexpr size_t SMEM = (size_t)(2 * Y * SW + Y) * sizeof(float2);   // two slabs + twiddle table
};

// One CTA owns an SW-column slab of the images (b, c0..c1, t).  The raw slab of the next coil is
// fetched with cp.async while the current one is transformed; phase A works in place on the slab.
template <int RY, int MODE>
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(ColCfg<RY>::NT, ColCfg<RY>::MINB)
sense_cols_kernel(const float2* __restrict__ src, float2* __restrict__ dst,
                  const float* __restrict__ w, WeightStrides ws, float scale,
                  SenseDims d, int cchunk, int ncchunks) {
    using Cfg = ColCfg<RY>;
    constexpr int DIR = (MODE == MODE_ADJ) ? +1 : -1;
    constexpr int ITEMS = Cfg::ITEMS, NT = Cfg::NT, SW = Cfg::SW, Y = Cfg::Y;
    DLSG_DYN_SMEM(float2, smem);
    float2* twy = smem + 2 * Y * SW;             // [l*RY + k2] = exp(-2*pi*i*l*k2/Y)  (forward sign)
    const int tid = threadIdx.x, col = tid % SW, grp = tid / SW;   // grp = k2 of this thread's radix-32

    const int ncolblk = d.X / SW;
    long long u = blockIdx.x;
    const int colblk = (int)(u % ncolblk); u /= ncolblk;
    const int cc = (int)(u % ncchunks); u /= ncchunks;
    const int tl = (int)(u % d.t_count);
    const int b = (int)(u / d.t_count);
    const int t = d.t_begin + tl;
    const int kx0 = colblk * SW, kx = kx0 + col;
    const int c0 = cc * cchunk, c1 = min(d.C, c0 + cchunk);
    const int Xs = d.X;

    auto prefetch = [&](int c, float2* slab) {
        const float2* in = src + ((MODE == MODE_ADJ) ? idx_k(d, b, c, t, 0) : idx_tmp(d, b, c, t, 0)) + kx0;
        for (int q = tid; q < Y * (SW / 2); q += NT) {
            const int yy = q / (SW / 2), ch = q - yy * (SW / 2);
            cp_async16(slab + yy * SW + 2 * ch, in + (size_t)yy * Xs + 2 * ch);
        }
        cp_async_commit();
    };

    if (c0 < c1) prefetch(c0, smem);
    for (int i = tid; i < Y; i += NT) {
        const int l = i / RY, k2 = i - l * RY;
        int m = (l * k2) % Y;
        if (2 * m > Y) m -= Y;
        float sn, cs;
        sincospif((float)(2 * m) / (float)Y, &sn, &cs);
        twy[i] = make_float2(cs, -sn);
    }

    // k-space weights of this thread's elements live in registers: ADJ needs them at its phase-A elements
    // (y = l + 32 j), FWD / NORMAL at its phase-B outputs (ky = RY k1 + grp).  Loaded before the wait on the
    // slab so that the latency overlaps; only once per CTA when the weights are coil-broadcast (ws.c == 0).
    constexpr int NW = (MODE == MODE_ADJ) ? ITEMS * RY : 32;
    float wreg[NW];
    auto load_weights = [&](int c) {
        const float* wp = w + (b * ws.b + c * ws.c + t * ws.t + kx * ws.x);
        if (MODE == MODE_ADJ) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const int l = grp + i * RY;
#pragma unroll
                for (int j = 0; j < RY; ++j) wreg[i * RY + j] = (l < 32) ? scale * __ldg(wp + (l + 32 * j) * ws.y) : 0.f;
            }
        } else {
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) wreg[k1] = scale * __ldg(wp + (RY * k1 + grp) * ws.y);
        }
    };
    if (!w) {
#pragma unroll
        for (int i = 0; i < NW; ++i) wreg[i] = scale;
    } else if (ws.c == 0 && c0 < c1) {
        load_weights(c0);
    }

    for (int c = c0; c < c1; ++c) {
        float2* slab = smem + ((c - c0) & 1) * (Y * SW);          // [k2*32 + l][col] after phase A
        float2* other = smem + (((c - c0) & 1) ^ 1) * (Y * SW);
        float2* outp = dst + ((MODE == MODE_FWD) ? idx_k(d, b, c, t, 0) : idx_tmp(d, b, c, t, 0)) + kx;
        if (w && ws.c != 0) load_weights(c);
        cp_async_wait_group0();
        __syncthreads();                         // this coil's slab has landed; the other slab is no longer read
        if (c + 1 < c1) prefetch(c + 1, other);
        // ---- phase A (in place): radix-RY over y = l + 32*j, twiddle
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int l = grp + i * RY;          // item = tid + i*NT  ->  l = item / SW
            if (l < 32) {
                float2 v[RY];
#pragma unroll
                for (int j = 0; j < RY; ++j) {
                    const int yy = l + 32 * j;
                    float2 a = slab[yy * SW + col];
                    if (MODE == MODE_ADJ) a = cscale(a, wreg[i * RY + j]);
                    v[j] = a;
                }
                XDft<RY, DIR>::run(v);
#pragma unroll
                for (int k2 = 0; k2 < RY; ++k2) {
                    float2 q = v[k2];
                    if (k2 > 0) {
                        const float2 tq = twy[l * RY + k2];
                        q = (DIR < 0) ? cmul(q, tq) : cmulc(q, tq);
                    }
                    slab[(k2 * 32 + l) * SW + col] = q;
                }
            }
        }
        __syncthreads();
        // ---- phase B: radix-32 over l for (k2 = grp, column = col); ky = RY*k1 + k2
        float2 r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = slab[(grp * 32 + i) * SW + col];
        XDft<32, DIR>::run(r);
        if (MODE == MODE_FWD) {
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) {
                outp[(RY * k1 + grp) * Xs] = cscale(r[k1], wreg[k1]);
            }
        } else if (MODE == MODE_ADJ) {
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) outp[(RY * k1 + grp) * Xs] = r[k1];
        } else {
            // weights squared, then the way back with the "32-first" split:
            // thread holds ky = RY*n1 + n2 (n2 = grp): radix-32 over n1 -> k1', twiddle
            // exp(+2*pi*i*n2*k1'/Y), exchange, radix-RY over n2 -> y = k1' + 32*k2'.
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) r[k1] = cscale(r[k1], wreg[k1] * wreg[k1]);
            XDft<32, +1>::run(r);
            __syncthreads();                     // every phase-B read of the slab is done
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) {
                float2 q = r[k1];
                if (grp > 0) q = cmulc(q, twy[k1 * RY + grp]);
                slab[(grp * 32 + k1) * SW + col] = q;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const int k1 = grp + i * RY;
                if (k1 < 32) {
                    float2 g[RY];
#pragma unroll
                    for (int n2 = 0; n2 < RY; ++n2) g[n2] = slab[(n2 * 32 + k1) * SW + col];
                    XDft<RY, +1>::run(g);
#pragma unroll
                    for (int k2 = 0; k2 < RY; ++k2) outp[(k1 + 32 * k2) * Xs] = g[k2];
                }
            }
        }
    }
}

// ============================================================================
// Direct-DFT kernels for axis lengths that are not 32*R (odd, prime, tiny ...).
// O(N^2) per line, twiddle table in shared memory; same tmp layout as above so the
// two axes can mix fast and direct kernels.
// ============================================================================
DLSG_DEVICE void fill_twiddles(float2* tw, int n, int dir, int tid, int nt) {
    for (int i = tid; i < n; i += nt) {
        // angle = dir*2*pi*i/n, reduced to |angle| <= pi so the fp32 argument is accurate
        const int m = (2 * i > n) ? i - n : i;
        float s, c;
        sincospif((float)(2 * m) / (float)n, &s, &c);
        tw[i] = make_float2(c, dir < 0 ? -s : s);
    }
}

// rows, forward.  CTA = (b, t, y); loops over coils.
//   final1d: 0 = unweighted into tmp (a column pass follows); 1 (ndims == 1) = W*scale applied, written
//   as k-space; 2 (ndims == 1, normal operator) = W*scale applied, written in the tmp layout
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(256)
sense_rows_fwd_direct_kernel(const float2* __restrict__ x, const float2* __restrict__ maps,
                             float2* __restrict__ dst, const float* __restrict__ w, WeightStrides ws,
                             float scale, int final1d, SenseDims d) {
    DLSG_DYN_SMEM(float2, smem);
    float2* row = smem;            // [X]
    float2* tw = smem + d.X;       // [X]
    const int tid = threadIdx.x, nt = blockDim.x;
    long long u = blockIdx.x;
    const int y = (int)(u % d.Y); u /= d.Y;
    const int tl = (int)(u % d.t_count);
    const int b = (int)(u / d.t_count);
    const int t = d.t_begin + tl;
    fill_twiddles(tw, d.X, -1, tid, nt);
    for (int c = 0; c < d.C; ++c) {
        __syncthreads();
        for (int i = tid; i < d.X; i += nt) {
            float2 a = make_float2(0.f, 0.f);
            for (int e = 0; e < d.E; ++e) {
                const float2 xv = __ldg(x + idx_x(d, b, e, t, y) + i);
                a = maps ? cfma(__ldg(maps + idx_maps(d, b, e, c, y) + i), xv, a) : cadd(a, xv);
            }
            row[i] = a;
        }
        __syncthreads();
        for (int k = tid; k < d.X; k += nt) {
            float2 a = make_float2(0.f, 0.f);
            int idx = 0;
            for (int n = 0; n < d.X; ++n) {
                a = cfma(row[n], tw[idx], a);
                idx += k; if (idx >= d.X) idx -= d.X;
            }
            if (final1d) a = cscale(a, scale * load_w(w, ws, b, c, t, y, k));
            if (final1d == 1) dst[idx_k(d, b, c, t, y) + k] = a;
            else              dst[idx_tmp(d, b, c, t, y) + k] = a;
        }
    }
}

// rows, adjoint.  CTA = (b, t, y); accumulates over coils in shared memory.
//   final1d: 0 = src is tmp, unweighted; 1 = src is k-space, W*scale at the load; 2 = src is tmp, W*scale at the load.
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(256)
sense_rows_adj_direct_kernel(const float2* __restrict__ src, const float2* __restrict__ maps,
                             float2* __restrict__ out, const float2* __restrict__ xin,
                             const float2* __restrict__ aty, float step_host,
                             const float* __restrict__ step_dev, int epi,
                             const float* __restrict__ w, WeightStrides ws, float scale, int final1d,
                             SenseDims d) {
    const float step = step_dev ? __ldg(step_dev) : step_host;
    DLSG_DYN_SMEM(float2, smem);
    float2* row = smem;                 // [X]
    float2* tw = smem + d.X;            // [X]
    float2* acc = smem + 2 * d.X;       // [E][X]
    const int tid = threadIdx.x, nt = blockDim.x;
    long long u = blockIdx.x;
    const int y = (int)(u % d.Y); u /= d.Y;
    const int tl = (int)(u % d.t_count);
    const int b = (int)(u / d.t_count);
    const int t = d.t_begin + tl;
    fill_twiddles(tw, d.X, +1, tid, nt);
    for (int i = tid; i < d.E * d.X; i += nt) acc[i] = make_float2(0.f, 0.f);
    for (int c = 0; c < d.C; ++c) {
        __syncthreads();
        for (int k = tid; k < d.X; k += nt) {
            float2 a = __ldg(src + (final1d == 1 ? idx_k(d, b, c, t, y) : idx_tmp(d, b, c, t, y)) + k);
            if (final1d) a = cscale(a, scale * load_w(w, ws, b, c, t, y, k));
            row[k] = a;
        }
        __syncthreads();
        for (int n = tid; n < d.X; n += nt) {
            float2 a = make_float2(0.f, 0.f);
            int idx = 0;
            for (int k = 0; k < d.X; ++k) {
                a = cfma(row[k], tw[idx], a);
                idx += n; if (idx >= d.X) idx -= d.X;
            }
            for (int e = 0; e < d.E; ++e)
                acc[e * d.X + n] = maps ? cfmac(a, __ldg(maps + idx_maps(d, b, e, c, y) + n), acc[e * d.X + n])
                                        : cadd(acc[e * d.X + n], a);
        }
    }
    __syncthreads();
    for (int i = tid; i < d.E * d.X; i += nt) {
        const int e = i / d.X, n = i - e * d.X;
        const size_t g = idx_x(d, b, e, t, y) + n;
        float2 r = acc[i];
        if (epi == EPI_DC) {
            const float2 xv = __ldg(xin + g), av = __ldg(aty + g);
            r = make_float2(fmaf(step, r.x - av.x, xv.x), fmaf(step, r.y - av.y, xv.y));
        } else if (epi == EPI_REG) {
            const float2 xv = __ldg(xin + g);
            r = make_float2(fmaf(step, xv.x, r.x), fmaf(step, xv.y, r.y));
        }
        out[g] = r;
    }
}

// columns, direct.  CTA = (b, c, t, block of CW columns).
constexpr int kDirectColW = 8;
template <int MODE>
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(256)
sense_cols_direct_kernel(const float2* __restrict__ src, float2* __restrict__ dst,
                         const float* __restrict__ w, WeightStrides ws, float scale, SenseDims d) {
    constexpr int CW = kDirectColW;
    DLSG_DYN_SMEM(float2, smem);
    float2* col = smem;                     // [Y][CW]
    float2* col2 = smem + (size_t)d.Y * CW; // [Y][CW]  (MODE_NORMAL only)
    float2* tw = smem + (size_t)2 * d.Y * CW;   // [Y] forward sign
    const int tid = threadIdx.x, nt = blockDim.x;
    const int ncolblk = (d.X + CW - 1) / CW;
    long long u = blockIdx.x;
    const int colblk = (int)(u % ncolblk); u /= ncolblk;
    const int tl = (int)(u % d.t_count); u /= d.t_count;
    const int c = (int)(u % d.C);
    const int b = (int)(u / d.C);
    const int t = d.t_begin + tl;
    const int cw = tid % CW, rl = tid / CW, nr = nt / CW;
    const int kx = colblk * CW + cw;
    const bool ok = kx < d.X;
    const size_t img_in = (MODE == MODE_ADJ) ? idx_k(d, b, c, t, 0) : idx_tmp(d, b, c, t, 0);
    const size_t img_out = (MODE == MODE_FWD) ? idx_k(d, b, c, t, 0) : idx_tmp(d, b, c, t, 0);
    fill_twiddles(tw, d.Y, -1, tid, nt);
    for (int yy = rl; yy < d.Y; yy += nr) {
        float2 a = make_float2(0.f, 0.f);
        if (ok) {
            a = __ldg(src + img_in + (size_t)yy * d.X + kx);
            if (MODE == MODE_ADJ) a = cscale(a, scale * load_w(w, ws, b, c, t, yy, kx));
        }
        col[yy * CW + cw] = a;
    }
    __syncthreads();
    for (int k = rl; k < d.Y; k += nr) {
        float2 a = make_float2(0.f, 0.f);
        int idx = 0;
        for (int n = 0; n < d.Y; ++n) {
            const float2 tq = tw[idx];
            a = (MODE == MODE_ADJ) ? cfmac(col[n * CW + cw], tq, a) : cfma(col[n * CW + cw], tq, a);
            idx += k; if (idx >= d.Y) idx -= d.Y;
        }
        if (MODE == MODE_FWD) {
            if (ok) dst[img_out + (size_t)k * d.X + kx] = cscale(a, scale * load_w(w, ws, b, c, t, k, kx));
        } else if (MODE == MODE_ADJ) {
            if (ok) dst[img_out + (size_t)k * d.X + kx] = a;
        } else {
            const float wv = ok ? scale * load_w(w, ws, b, c, t, k, kx) : 0.f;
            col2[k * CW + cw] = cscale(a, wv * wv);
        }
    }
    if (MODE == MODE_NORMAL) {
        __syncthreads();
        for (int n = rl; n < d.Y; n += nr) {
            float2 a = make_float2(0.f, 0.f);
            int idx = 0;
            for (int k = 0; k < d.Y; ++k) {
                a = cfmac(col2[k * CW + cw], tw[idx], a);
                idx += n; if (idx >= d.Y) idx -= d.Y;
            }
            if (ok) dst[img_out + (size_t)n * d.X + kx] = a;
        }
    }
}

}  // namespace dlsg
