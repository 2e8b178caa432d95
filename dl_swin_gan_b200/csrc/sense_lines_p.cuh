// Sampled-line column passes, second generation (sm_100a): packed-pair arithmetic and pruned transforms.
// Same operator as sense_lines.cuh (SenseModel._forward_op / _adjoint_op, dl_cs/mri/transforms.py:84-98; DC update
// dl_cs/models/unrolledswin.py:109) and the same intermediate `tl` [B,C,T,lmax,X], so the row pass on the sampled
// lines (sense_rows_lines_kernel) is shared.
//
// Why a second generation.  The first one was issue- and latency-bound (profiles/r1q_lines_ncu_summary.csv: 12 warps/SM,
// issue-active 56 %, 1594 instructions per thread and coil for ~900 of FP32 work).  Here
//   * every complex add / multiply / fma is one packed instruction on a register pair (FADD2 / FFMA2,
//     fft_packed.cuh): the same FP32 pipe time in half the issue slots;
//   * Y = 16 * R2 is split as an in-register radix-16 DFT (one thread, 16 complex registers instead of 32) and a
//     radix-R2 stage that is only evaluated for the sampled lines ("pruned"):
//       forward  y = R2 l + j :  Z_j[k1] = DFT16_l( img[R2 l + j] ),   tl[i] = sum_j w_i^j Z_j[ky_i mod 16],  w_i = W_Y^ky_i
//       adjoint                  U_j[k1] = sum_{i : ky_i mod 16 = k1} conj(w_i)^j tl[i],
//                                out[R2 l + j] = IDFT16_k1( U_j )[l]
//     so the dense part of a column transform is R2 DFT16s and the rest is ~n * R2 complex fmas per column; the
//     powers of w come from Horner's rule / repeated multiplication on the even and odd terms (one coefficient
//     pair per line, precomputed in the line plan together with the grouping of the lines by ky mod 16);
//   * a CTA is a 16-column slab x R2 threads per column (192 threads for Y = 192), 3 CTAs per SM in the forward kernel;
//   * the coil-map slab [Y rows][16 columns] of each coil arrives by TMA (cp.async.bulk.tensor, one copy per map set
//     issued by one thread, completion on an mbarrier) and every thread reads its own 16 values from it; x stays in
//     registers.  No per-thread global loads in the coil loop, nothing held in registers while a slab is in flight.
// Measurements and the variants that were rejected: DESIGN.md section 3c.
#pragma once

#include "sense_lines.cuh"

namespace dlsg {

template <int R2>
struct PColCfg {
    static constexpr int R1 = 16, Y = 16 * R2, SW = 16, NT = SW * R2;
    static constexpr size_t SLAB = (size_t)16 * NT;                              // complex elements of one [16][NT] block
    static constexpr int NBOX = (Y > 256) ? 2 : 1;                               // TMA boxes per slab (a box has at most 256 rows)
    static size_t tables(int lmax) { return 16 + (size_t)lmax * (2 * sizeof(float4) + 2 * sizeof(int)) + (17 + 16 + 2 + 1) * sizeof(int); }
    static size_t smem_fwd(int E, int lmax, int mb, int xb) { return (size_t)(E * mb + xb) * SLAB * sizeof(float2) + tables(lmax); }
    static size_t smem_adj(int E, int lmax, int xb) {
        return (size_t)(E * 2 + xb) * SLAB * sizeof(float2) + (size_t)2 * lmax * SW * sizeof(float2) + tables(lmax);
    }
};

// block index -> (16-column slab, frame, slice)
struct PColWho { int b, t, bt, kx0; };
DLSG_DEVICE PColWho pcol_who(const SenseDims& d) {
    const int ncolblk = d.X / 16;
    long long u = blockIdx.x;
    const int colblk = (int)(u % ncolblk); u /= ncolblk;
    const int tl = (int)(u % d.t_count);
    PColWho w;
    w.b = (int)(u / d.t_count);
    w.t = d.t_begin + tl;
    w.bt = w.b * d.T + w.t;
    w.kx0 = colblk * 16;
    return w;
}

// (cos, sin) of 2 pi m / Y with the integer reduced to |angle| <= pi first
DLSG_DEVICE float2 unit_root(int m, int Y) {
    m %= Y;
    if (2 * m > Y) m -= Y;
    float sn, cs;
    sincospif((float)(2 * m) / (float)Y, &sn, &cs);
    return make_float2(cs, sn);
}

// ---------------------------------------------------------------------------- per-CTA line tables
// The sampled lines of one (b, t), grouped by class k1 = ky mod 16 (the DFT16 bin they meet):
//   seg[k1] .. seg[k1+1]  sorted positions of class k1 (ascending ky inside a class)
//   order[q]              line index i (row of `tl`) at sorted position q
//   wq[2q], wq[2q+1]      (c, s, -s, c) of w = exp(SIGN * 2 pi i ky / Y) and of w^2
//   clist[0..m)           the non-empty classes, meta[0] = m, meta[1] = bit mask of the non-empty classes
struct LineTables { float4* wq; int* order; int* seg; int* clist; int* meta; int* k1q; };
DLSG_DEVICE LineTables line_tables_at(void* p, int lmax) {
    LineTables t;
    t.wq = reinterpret_cast<float4*>(p);
    t.order = reinterpret_cast<int*>(t.wq + 2 * lmax);
    t.seg = t.order + lmax;
    t.clist = t.seg + 17;
    t.meta = t.clist + 16;
    t.k1q = t.meta + 2;                                      // [lmax] (ky mod 16) * NT of sorted line q
    return t;
}
// copies the tables of (b, t) from the plan (built once per weights tensor by sense_line_plan_kernel) into shared
// memory; SIGN = +1 conjugates the coefficients.  All threads of the CTA call it; the caller's next barrier publishes them.
template <int SIGN>
DLSG_DEVICE void load_line_tables(const LineTables& lt, const LinePlan& lp, int bt, int n, int Y, int tid, int nt) {
    const float4* src = reinterpret_cast<const float4*>(lp.wq + (size_t)bt * Y * 8);
    for (int i = tid; i < 2 * n; i += nt) {
        const float4 w = __ldg(src + i);
        lt.wq[i] = (SIGN < 0) ? w : make_float4(w.x, w.z, w.y, w.w);       // conj: (c, s, -s, c) -> (c, -s, s, c)
    }
    const int* ord = lp.order + (size_t)bt * Y;
    const int* list = lp.list + (size_t)bt * Y;
    for (int i = tid; i < n; i += nt) {
        const int o = __ldg(ord + i);
        lt.order[i] = o;
        lt.k1q[i] = (__ldg(list + o) & 15) * nt;                            // nt = 16 * R2 = row stride of the exchange
    }
    const int* cls = lp.cls + (size_t)bt * kPlanClsInts;
    for (int i = tid; i < 35; i += nt) lt.seg[i] = __ldg(cls + i);          // seg[17], clist[16], meta[2] are contiguous
}

// The coil's map slab [Y rows][16 columns] of every map set arrives by TMA (one bulk tensor copy per map set, issued by
// one thread, completion counted on an mbarrier): no per-thread global loads, no registers held while in flight.
template <int R2, int E>
DLSG_DEVICE void tma_maps(float2* slab, const TileMap* tm, int b, int c, int C, int kx0, mbar_t* bar) {
    constexpr int Y = 16 * R2, NBOX = PColCfg<R2>::NBOX, BR = Y / NBOX;
    mbar_expect_tx(bar, (unsigned)(E * Y * 16 * sizeof(float2)));
#pragma unroll
    for (int e = 0; e < E; ++e)
#pragma unroll
        for (int h = 0; h < NBOX; ++h)
            tile_load_2d(slab + (size_t)e * PColCfg<R2>::SLAB + (size_t)h * BR * 16, tm, 2 * kx0,
                         ((b * E + e) * C + c) * Y + h * BR, bar);
}

// tl[b,c,t,i,kx] = DFT_Y( sum_e maps[b,e,c,:,kx] * x[b,e,t,:,kx] )[ky_i]   for the sampled lines of (b,t).
// Per coil: product of the TMA-staged map slab with the thread's x values, DFT16, the bins of the non-empty classes go
// to shared memory, then one thread per (class, column) reads the R2 bins of its class once and evaluates
// sum_j w^j Z_j for every line of the class (w = W_Y^ky; Horner's rule on the even and the odd terms).
// MB / XB: buffers of the map slab / of the exchange (XB = 2: one CTA barrier per coil, XB = 1: two).
template <int R2, int E, int MB, int XB>
#ifndef DLSG_PCOLS_FWD_MINB
#define DLSG_PCOLS_FWD_MINB(SLABS) ((SLABS) <= 3 ? 3 : 2)      // resident CTAs the register allocation must allow, by shared-memory slabs
#endif
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(16 * R2, DLSG_PCOLS_FWD_MINB(E * MB + XB))
sense_cols_fwd_lines_p_kernel(const float2* __restrict__ x, const DLSG_GRID_CONSTANT TileMap maps_tm,
                              float2* __restrict__ tl_out, LinePlan lp, SenseDims d) {
    using Cfg = PColCfg<R2>;
    constexpr int NT = Cfg::NT, Y = Cfg::Y;
    static_assert(R2 % 2 == 0, "even R2");
    DLSG_DYN_SMEM(float2, smem);
    float2* mslab = smem;                                        // [MB][E][Y][16] map slabs (TMA destinations)
    float2* exch = mslab + (size_t)MB * E * Cfg::SLAB;           // [XB][16][NT] Z_j[k1] at (k1, tid = j*16 + col)
    mbar_t* bars = reinterpret_cast<mbar_t*>(exch + (size_t)XB * Cfg::SLAB);   // [MB]
    const LineTables lt = line_tables_at(bars + 2, lp.lmax);
    const int tid = threadIdx.x, col = tid & 15, j = tid >> 4;
    const PColWho who = pcol_who(d);
    const int b = who.b, t = who.t, kx = who.kx0 + col;
    const int n = __ldg(lp.counts + who.bt);
    if (n == 0) return;                                          // nothing sampled in this frame (uniform over the CTA)

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < MB; ++i) mbar_init(bars + i, 1);
        mbar_fence_init();
    }
    const size_t rstride = (size_t)R2 * d.X;
    float2 xr[E][16];                                            // x[b,e,t, R2 l + j, kx]: the same for every coil
#pragma unroll
    for (int e = 0; e < E; ++e)
#pragma unroll
        for (int l = 0; l < 16; ++l) xr[e][l] = __ldg(x + idx_x(d, b, e, t, j) + kx + l * rstride);
    load_line_tables<-1>(lt, lp, who.bt, n, Y, tid, NT);
    __syncthreads();                                             // tables and mbarrier initialisation visible
    const int m = lt.meta[0], mask = lt.meta[1];
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < MB; ++i)
            if (i < d.C) tma_maps<R2, E>(mslab + (size_t)i * E * Cfg::SLAB, &maps_tm, b, i, d.C, who.kx0, bars + i);
    }

#pragma unroll 1
    for (int c = 0; c < d.C; ++c) {
        const int mb = c % MB;
        const float2* ms = mslab + (size_t)mb * E * Cfg::SLAB + tid;        // slab element (R2 l + j, col) = [l * NT + tid]
        mbar_wait(bars + mb, (unsigned)((c / MB) & 1));
        float2 v[16];
#pragma unroll
        for (int l = 0; l < 16; ++l) {
            v[l] = pc_mul(ms[l * NT], xr[0][l]);
#pragma unroll
            for (int e = 1; e < E; ++e) v[l] = pc_fma(ms[(e * 16 + l) * NT], xr[e][l], v[l]);
        }
        PDft<16, -1>::run(v);
        float2* ex = exch + (size_t)(c % XB) * Cfg::SLAB;
        if (XB == 1) __syncthreads();                            // the previous coil's bins have been consumed
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1)
            if ((mask >> k1) & 1) ex[k1 * NT + tid] = v[k1];
        __syncthreads();                                         // bins visible; everyone has read the map slab
        if (tid == 0 && c + MB < d.C)
            tma_maps<R2, E>(mslab + (size_t)mb * E * Cfg::SLAB, &maps_tm, b, c + MB, d.C, who.kx0, bars + mb);
#ifdef DLSG_PCOLS_GATHER_BY_CLASS
        float2* outp = tl_out + idx_lines(d, lp.lmax, b, c, t, 0) + kx;
        for (int wi = j; wi < m; wi += R2) {                     // this thread's (class, column) items
            const int k1 = lt.clist[wi];
            const float2* zp = ex + k1 * NT + col;
            float2 z[R2];
#pragma unroll
            for (int jj = 0; jj < R2; ++jj) z[jj] = zp[jj * 16];
            const int qe = lt.seg[k1 + 1];
#pragma unroll 1
            for (int q = lt.seg[k1]; q < qe; ++q) {
                const float4 wa = lt.wq[2 * q], wb = lt.wq[2 * q + 1];
                const float2 w1 = make_float2(wa.x, wa.y), w2 = make_float2(wa.z, wa.w);      // w
                const float2 v1 = make_float2(wb.x, wb.y), v2 = make_float2(wb.z, wb.w);      // w^2
                float2 ae = z[R2 - 2], ao = z[R2 - 1];
#pragma unroll
                for (int jj = R2 - 4; jj >= 0; jj -= 2) {
                    ae = pc_fma_w(v1, v2, ae, z[jj]);            // ae <- z[jj] + w^2 ae
                    ao = pc_fma_w(v1, v2, ao, z[jj + 1]);
                }
                outp[(size_t)lt.order[q] * d.X] = pc_fma_w(w1, w2, ao, ae);                   // even + w * odd
            }
        }
#else
        // item = (sorted line q, column): n * 16 items over the NT threads; every thread takes the same number of rounds
        // (the last one partly idle), so the warps reach the next barrier together
        float2* outp = tl_out + idx_lines(d, lp.lmax, b, c, t, 0) + who.kx0;
        for (int item = tid; item < n * 16; item += NT) {
            const int q = item >> 4, cl = item & 15;
            const float2* zp = ex + lt.k1q[q] + cl;
            const float4 wa = lt.wq[2 * q], wb = lt.wq[2 * q + 1];
            const float2 w1 = make_float2(wa.x, wa.y), w2 = make_float2(wa.z, wa.w);          // w
            const float2 v1 = make_float2(wb.x, wb.y), v2 = make_float2(wb.z, wb.w);          // w^2
            float2 ae = zp[(R2 - 2) * 16], ao = zp[(R2 - 1) * 16];
#pragma unroll
            for (int jj = R2 - 4; jj >= 0; jj -= 2) {
                ae = pc_fma_w(v1, v2, ae, zp[jj * 16]);          // ae <- z[jj] + w^2 ae
                ao = pc_fma_w(v1, v2, ao, zp[(jj + 1) * 16]);
            }
            outp[(size_t)lt.order[q] * d.X + cl] = pc_fma_w(w1, w2, ao, ae);                  // even + w * odd
        }
#endif
    }
}

// out_e[b,t,y,kx] = EPI( sum_c conj(maps[b,e,c,y,kx]) * IDFT_Y( scatter of tl[b,c,t,:,kx] to its ky lines )[y] )
// Per coil: one thread per (class, column) forms U_j[k1] = sum_{lines of the class} conj(w)^j tl (powers by repeated
// multiplication with w^2 on the even and odd terms, each line read once), the R2 values go to shared memory; then
// every thread takes its 16 bins (zeros for the empty classes), IDFT16, and accumulates conj(S) * out in registers,
// S from the TMA-staged map slab (two buffers: the slab of coil c+2 is requested once everyone has used coil c's).
template <int R2, int E, int XB>
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(16 * R2, 2)
sense_cols_adj_lines_p_kernel(const float2* __restrict__ tl_in, const DLSG_GRID_CONSTANT TileMap maps_tm, float2* __restrict__ out,
                              const float2* __restrict__ xin, const float2* __restrict__ aty,
                              float step_host, const float* __restrict__ step_dev, int epi, LinePlan lp, SenseDims d) {
    using Cfg = PColCfg<R2>;
    constexpr int NT = Cfg::NT, Y = Cfg::Y, SW = 16, MB = 2;
    DLSG_DYN_SMEM(float2, smem);
    float2* mslab = smem;                                        // [MB][E][Y][16]
    float2* exch = mslab + (size_t)MB * E * Cfg::SLAB;           // [XB][16][NT] U_j[k1] at (k1, j*16 + col)
    float2* tbuf = exch + (size_t)XB * Cfg::SLAB;                // [2][lmax][SW] sampled lines of a coil in sorted order
    mbar_t* bars = reinterpret_cast<mbar_t*>(tbuf + (size_t)2 * lp.lmax * SW);
    const LineTables lt = line_tables_at(bars + 2, lp.lmax);
    const int tid = threadIdx.x, col = tid & 15, j = tid >> 4;
    const PColWho who = pcol_who(d);
    const int b = who.b, t = who.t, kx = who.kx0 + col;
    const int n = __ldg(lp.counts + who.bt);
    const float step = step_dev ? __ldg(step_dev) : step_host;
    const size_t rstride = (size_t)R2 * d.X;

    auto fetch_lines = [&](int c) {                              // sorted position q <- line order[q]
        const float2* src = tl_in + idx_lines(d, lp.lmax, b, c, t, 0) + who.kx0;
        float2* dst = tbuf + (size_t)(c & 1) * lp.lmax * SW;
        for (int u = tid; u < n * 8; u += NT) {
            const int q = u >> 3, ch = u & 7;
            cp_async16(dst + q * SW + 2 * ch, src + (size_t)lt.order[q] * d.X + 2 * ch);
        }
    };

    float2 acc[E][16];
#pragma unroll
    for (int e = 0; e < E; ++e)
#pragma unroll
        for (int l = 0; l < 16; ++l) acc[e][l] = make_float2(0.f, 0.f);

    int m = 0, mask = 0;
    if (n > 0) {
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < MB; ++i) mbar_init(bars + i, 1);
            mbar_fence_init();
        }
        load_line_tables<+1>(lt, lp, who.bt, n, Y, tid, NT);
        __syncthreads();                                         // tables and mbarrier initialisation visible
        m = lt.meta[0]; mask = lt.meta[1];
        if (tid == 0) {
#pragma unroll
            for (int i = 0; i < MB; ++i)
                if (i < d.C) tma_maps<R2, E>(mslab + (size_t)i * E * Cfg::SLAB, &maps_tm, b, i, d.C, who.kx0, bars + i);
        }
        fetch_lines(0);
        cp_async_commit();
        cp_async_wait_group0();
        __syncthreads();                                         // lines of coil 0 visible
        if (d.C > 1) fetch_lines(1);
    }
    cp_async_commit();

#pragma unroll 1
    for (int c = 0; c < d.C && n > 0; ++c) {
        // (1) class phase: lines of coil c (visible since the previous barrier) -> U_j[k1]
        const float2* tb = tbuf + (size_t)(c & 1) * lp.lmax * SW + col;
        float2* ex = exch + (size_t)(c % XB) * Cfg::SLAB;
        if (XB == 1 && c > 0) __syncthreads();                   // everyone has taken the previous coil's bins
        for (int wi = j; wi < m; wi += R2) {
            const int k1 = lt.clist[wi];
            float2 uj[R2];
#pragma unroll
            for (int jj = 0; jj < R2; ++jj) uj[jj] = make_float2(0.f, 0.f);
            const int qe = lt.seg[k1 + 1];
#pragma unroll 1
            for (int q = lt.seg[k1]; q < qe; ++q) {
                const float4 wa = lt.wq[2 * q], wb = lt.wq[2 * q + 1];
                const float2 w1 = make_float2(wa.x, wa.y), w2 = make_float2(wa.z, wa.w);      // w
                const float2 v1 = make_float2(wb.x, wb.y), v2 = make_float2(wb.z, wb.w);      // w^2
                const float2 zero = make_float2(0.f, 0.f);
                float2 pe = tb[q * SW];                                                       // w^0 T
                float2 po = pc_fma_w(w1, w2, pe, zero);                                       // w^1 T
                uj[0] = p_add(uj[0], pe);
                uj[1] = p_add(uj[1], po);
#pragma unroll
                for (int jj = 2; jj < R2; jj += 2) {
                    pe = pc_fma_w(v1, v2, pe, zero);
                    po = pc_fma_w(v1, v2, po, zero);
                    uj[jj] = p_add(uj[jj], pe);
                    uj[jj + 1] = p_add(uj[jj + 1], po);
                }
            }
            float2* up = ex + k1 * NT + col;
#pragma unroll
            for (int jj = 0; jj < R2; ++jj) up[jj * 16] = uj[jj];
        }
        cp_async_wait_group0();                                  // this thread's chunks of the next coil's lines
        __syncthreads();                                         // (2) bins of coil c and lines of coil c+1 visible; tbuf[c & 1] and the slab of coil c-1 are free
        if (c + 2 < d.C) fetch_lines(c + 2);
        cp_async_commit();
        if (tid == 0 && c >= 1 && c - 1 + MB < d.C)
            tma_maps<R2, E>(mslab + (size_t)((c - 1) % MB) * E * Cfg::SLAB, &maps_tm, b, c - 1 + MB, d.C, who.kx0, bars + (c - 1) % MB);
        float2 u[16];
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) u[k1] = ((mask >> k1) & 1) ? ex[k1 * NT + tid] : make_float2(0.f, 0.f);
        PDft<16, +1>::run(u);
        const int mb = c % MB;
        const float2* ms = mslab + (size_t)mb * E * Cfg::SLAB + tid;
        mbar_wait(bars + mb, (unsigned)((c / MB) & 1));
#pragma unroll
        for (int e = 0; e < E; ++e)
#pragma unroll
            for (int l = 0; l < 16; ++l) acc[e][l] = pc_fmac(ms[(e * 16 + l) * NT], u[l], acc[e][l]);
    }
    // ---- epilogue: this thread owns y = R2 l + j of column kx
    // out = a * (acc - av) + bx * xv :  PLAIN (a, bx) = (1, 0);  DC (step, 1) with av = aty;  REG (1, step)
    const float ea = (epi == EPI_DC) ? step : 1.f;
    const float eb = (epi == EPI_DC) ? 1.f : (epi == EPI_REG ? step : 0.f);
#pragma unroll
    for (int e = 0; e < E; ++e) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {                            // two batches of 8 rows: all loads of a batch in flight together
            const size_t g0 = idx_x(d, b, e, t, j) + kx + (size_t)(8 * h) * rstride;
            float2 xv[8], av[8];
#pragma unroll
            for (int l = 0; l < 8; ++l) xv[l] = av[l] = make_float2(0.f, 0.f);
            if (epi != EPI_PLAIN) {
#pragma unroll
                for (int l = 0; l < 8; ++l) xv[l] = __ldg(xin + g0 + l * rstride);
            }
            if (epi == EPI_DC) {
#pragma unroll
                for (int l = 0; l < 8; ++l) av[l] = __ldg(aty + g0 + l * rstride);
            }
#pragma unroll
            for (int l = 0; l < 8; ++l)
                out[g0 + l * rstride] = p_fma(p_bc(ea), p_sub(acc[e][8 * h + l], av[l]), p_mul(p_bc(eb), xv[l]));
        }
    }
}


// ---------------------------------------------------------------------------- row pass of the normal operator, second generation
//   tl[b,c,t,i,:] <- IDFT_X( (W*scale)^2 * DFT_X( tl[b,c,t,i,:] ) )      in place, for the sampled lines of (b, t)
// X = 16 * RX.  A warp takes two lines at a time; both transforms share their middle stage, so a line crosses shared
// memory twice instead of four times:
//   lane (line, j), j < 16 :  a[l] = tl[16 l + j], l < RX          DFT_RX over l, twiddle W_X^(j k1)      -> exchange 1  [k1][j]
//   lane (line, k1), k1 < RX: DFT16 over j  ->  k[k1 + RX k2], k2 < 16;   times (W scale)^2;   IDFT16 over k2 -> n2;
//                             twiddle conj(W_X)^(n2 k1)                                                   -> exchange 2  [n2][k1]
//   lane (line, n2), n2 < 16: IDFT_RX over k1 -> tl[n2 + 16 n1], n1 < RX      (the layout the line was loaded in)
// Global accesses are 128-byte segments (16 lanes x 8 bytes per l), the exchanges are warp-private (only __syncwarp),
// registers stay below 80 so that 6 CTAs of 4 warps share an SM.  Packed butterflies (fft_packed.cuh).
template <int RX>
struct PRowCfg {
    static constexpr int X = 16 * RX, WARPS = 4;
    static constexpr int S1 = 17, S2 = RX + 1;                       // padded strides of the two exchange layouts (conflict-free)
    static constexpr int EX1 = RX * S1, EX2 = 16 * S2;               // complex elements per line
    static constexpr int EXW = 2 * (EX1 + EX2);                      // per warp (two lines)
    static size_t smem() { return ((size_t)WARPS * EXW + X) * sizeof(float2); }   // + twiddle table W_X^m, m = j k1 < X
};

template <int RX>
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(128, 5)
sense_rows_normal_lines_p_kernel(float2* __restrict__ tl, const float* __restrict__ w, WeightStrides ws, float scale,
                                 LinePlan lp, SenseDims d, int nsplit) {
    using Cfg = PRowCfg<RX>;
    constexpr int X = Cfg::X, S1 = Cfg::S1, S2 = Cfg::S2;
    DLSG_DYN_SMEM(float2, smem);
    float2* twt = smem;                                              // [X]: exp(-2 pi i m / X)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* ex1 = smem + X + (size_t)warp * Cfg::EXW;                // [2 lines][RX][S1]  (k1, j)
    float2* ex2 = ex1 + 2 * Cfg::EX1;                                // [2 lines][16][S2]  (n2, k1)
    for (int m = threadIdx.x; m < X; m += blockDim.x) {
        float sn, cs;
        sincospif((float)(2 * (2 * m > X ? m - X : m)) / (float)X, &sn, &cs);
        twt[m] = make_float2(cs, -sn);
    }
    __syncthreads();
    const int split = (int)(blockIdx.x % nsplit), unit = (int)(blockIdx.x / nsplit);
    const int tloc = unit % d.t_count, b = unit / d.t_count, t = d.t_begin + tloc;
    const int bt = b * d.T + t;
    const int n = __ldg(lp.counts + bt);
    const int rows = d.C * n;                                        // (coil, line) rows of this (b, t), line fastest
    const int* list = lp.list + (size_t)bt * d.Y;
    const int half = lane >> 4, j = lane & 15;                       // stages "RX": two lines x 16 lanes
    const float s2 = scale * scale;

    for (int r0 = 2 * (split * Cfg::WARPS + warp); r0 < rows; r0 += 2 * nsplit * Cfg::WARPS) {
        const int nl = min(2, rows - r0);                            // lines of this round (uniform over the warp)
        const int row = r0 + half;
        const bool on = half < nl;
        const int c = on ? row / n : 0, i = on ? row - c * n : 0;
        float2* line = tl + idx_lines(d, lp.lmax, b, c, t, i);
        // ---- load, DFT_RX over l, twiddle, exchange 1
        if (on) {
            float2 a[RX];
#pragma unroll
            for (int l = 0; l < RX; ++l) a[l] = line[16 * l + j];
            PDft<RX, -1>::run(a);
            float2* e1 = ex1 + half * Cfg::EX1 + j;
#pragma unroll
            for (int k1 = 0; k1 < RX; ++k1) e1[k1 * S1] = (k1 == 0) ? a[0] : pc_mul(a[k1], twt[j * k1]);
        }
        __syncwarp();
        // ---- middle, items (line, k1): DFT16 over j, weights, IDFT16 over k2, twiddle, exchange 2
        for (int item = lane; item < nl * RX; item += 32) {
            const int hl = item / RX, k1 = item - hl * RX;
            const int rw = r0 + hl, cc = rw / n, ii = rw - cc * n;
            const float* wr = w + (b * ws.b + cc * ws.c + t * ws.t + __ldg(list + ii) * ws.y + k1 * ws.x);
            float wv[16];
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) wv[k2] = __ldg(wr + (long long)(RX * k2) * ws.x);
            const float2* src = ex1 + hl * Cfg::EX1 + k1 * S1;
            float2 u[16];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) u[jj] = src[jj];
            PDft<16, -1>::run(u);                                    // u[k2] = X[k1 + RX k2]
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) u[k2] = p_mul(u[k2], p_bc(wv[k2] * wv[k2] * s2));
            PDft<16, +1>::run(u);                                    // over k2 -> n2
            float2* dst = ex2 + hl * Cfg::EX2 + k1;
            dst[0] = u[0];
#pragma unroll
            for (int n2 = 1; n2 < 16; ++n2) {
                const float2 tw = twt[n2 * k1];
                dst[n2 * S2] = pc_mul(u[n2], make_float2(tw.x, -tw.y));     // conj(W_X)^(n2 k1)
            }
        }
        __syncwarp();
        // ---- IDFT_RX over k1, store (n2 = j: the layout the line was loaded in)
        if (on) {
            float2 a[RX];
            const float2* e2 = ex2 + half * Cfg::EX2 + j * S2;
#pragma unroll
            for (int k1 = 0; k1 < RX; ++k1) a[k1] = e2[k1];
            PDft<RX, +1>::run(a);
#pragma unroll
            for (int n1 = 0; n1 < RX; ++n1) line[j + 16 * n1] = a[n1];
        }
        __syncwarp();                                                // the exchanges are free for the next round
    }
}

}  // namespace dlsg
