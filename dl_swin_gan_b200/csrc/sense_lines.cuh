// Sampled-line path of the SENSE operator (sm_100a): same operator as sense_kernels.cuh
//   SenseModel._forward_op / _adjoint_op   dl_cs/mri/transforms.py:84-98
//   DC update                              dl_cs/models/unrolledswin.py:109
// for k-space weights W that are zero on whole ky lines, which is what the reference's k-t sampling
// masks are (dl_cs/mri/utils.py:69-79 get_mask of ky-undersampled data; 12x in config_recon.yaml).
//
// The 2-D transform is separable, so the axis order is free.  Here the Y axis goes first:
//   forward :  cols_fwd_lines  tl[b,c,t,i,:] = DFT_Y( sum_e S x )[ky_i, :]      only the sampled lines ky_i are stored
//              rows_lines<FWD> k[b,c,t,ky_i,:] = W * DFT_X( tl[i,:] )            unsampled lines of k are zero-filled
//   adjoint :  rows_lines<ADJ> tl[i,:] = IDFT_X( W * k[ky_i,:] )
//              cols_adj_lines  x = sum_c conj(S) * IDFT_Y( scatter(tl) )  (+ DC / REG epilogue)
//   normal  :  cols_fwd_lines, rows_lines<NORMAL> (DFT_X, W^2, IDFT_X in place), cols_adj_lines
// so the intermediate `tl` [B,C,T,lmax,X] holds n(b,t) <= lmax lines per image instead of Y, the X passes run on
// those lines only, and the fused normal operator moves ~4x less HBM traffic than the dense path at 12x undersampling.
// A line is "sampled" when any weight on it (over coils and kx) is non-zero, so the result equals the dense
// path's for finite data (0 * k contributes exactly 0 there too); the plan is built once per weights tensor.
#pragma once

#include "sense_kernels.cuh"

// Tuning knobs (defaults are what ships; overridden only by the variant builds of tools/build_variant.sh)
#ifndef DLSG_LINES_FWD_ITEM_UNROLL
#define DLSG_LINES_FWD_ITEM_UNROLL 2        // phase-A items unrolled together in cols_fwd (full unrolling overflows the instruction cache there)
#endif
#ifndef DLSG_LINES_ADJ_ITEM_UNROLL
#define DLSG_LINES_ADJ_ITEM_UNROLL 8        // cols_adj: full (ITEMS <= 8) measured 3 % faster than 2
#endif
#ifndef DLSG_LINES_FWD_MINB
#define DLSG_LINES_FWD_MINB(RY) ColCfg<RY>::MINB
#endif
#ifndef DLSG_LINES_ADJ_MINB
#define DLSG_LINES_ADJ_MINB(RY) ColCfg<RY>::MINB
#endif
#ifndef DLSG_LINES_ROWS_MINB
#define DLSG_LINES_ROWS_MINB 3
#endif
#define DLSG_PRAGMA(x) _Pragma(#x)
#define DLSG_UNROLL(n) DLSG_PRAGMA(unroll n)

namespace dlsg {

// Device-resident plan (int32), see dlsg_sense_line_plan_build:
//   counts[b*T + t]          n(b,t)
//   list  [(b*T + t)*Y + i]  ky of the i-th sampled line, ascending
//   pos   [(b*T + t)*Y + ky] i with list[i] == ky, or -1
//   order [(b*T + t)*Y + q]  line index i at sorted position q, lines grouped by class ky mod 16 (ascending ky inside a class)
//   cls   [(b*T + t)*36 + .] seg[17] (first sorted position of each class), clist[16] (the non-empty classes), m, bit mask, pad
//   wq    [((b*T + t)*Y + q)*8 + .] floats (c, s, -s, c) of w = exp(-2 pi i ky / Y) and of w^2, sorted order
// (the last three serve the pruned column kernels of sense_lines_p.cuh)
struct LinePlan { const int* counts; const int* list; const int* pos; int lmax; const int* order; const int* cls; const float* wq; };
constexpr int kPlanClsInts = 36;
// int32 elements in front of wq, rounded up so that wq (read as float4) is 16-byte aligned, and of the whole plan
DLSG_HD_CE size_t line_plan_wq_offset(int B, int T, int Y) { return (((size_t)B * T * (1 + 3 * (size_t)Y + kPlanClsInts)) + 3) / 4 * 4; }
DLSG_HD_CE size_t line_plan_ints(int B, int T, int Y) { return line_plan_wq_offset(B, T, Y) + (size_t)B * T * Y * 8; }

DLSG_DEVICE size_t idx_lines(const SenseDims& d, int lmax, int b, int c, int t, int i) {
    return ((((size_t)b * d.C + c) * d.t_count + (t - d.t_begin)) * lmax + i) * (size_t)d.X;
}

// ---------------------------------------------------------------------------- plan
// CTA = (b, t): flags the ky lines that carry any non-zero (or NaN) weight, then one thread compacts them.
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(256)
sense_line_plan_kernel(const float* __restrict__ w, WeightStrides ws, int C, int T, int Y, int X,
                       int* counts, int* list, int* pos, int* order, int* cls, float* wq, int cap) {
    DLSG_DYN_SMEM(int, flag);                    // [Y]
    const int bt = blockIdx.x, b = bt / T, t = bt - b * T;
    for (int y = threadIdx.x; y < Y; y += blockDim.x) flag[y] = 0;
    __syncthreads();
    const int nc = ws.c ? C : 1, nx = ws.x ? X : 1, ny = ws.y ? Y : 1;   // broadcast axes are visited once
    const float* base = w + (b * ws.b + t * ws.t);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int row = warp; row < nc * ny; row += nwarps) {                  // one (coil, line) row per warp, lanes along kx
        const int c = row / ny, y = row - c * ny;
        const float* wr = base + (c * ws.c + y * ws.y);
        bool nz = false;
        for (int x = lane; x < nx; x += 32) nz |= (__ldg(wr + x * ws.x) != 0.f);
        if (nz) flag[y] = 1;                                              // same value from every writer
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        // cap > 0: the caller sized the intermediate for at most `cap` lines per frame without reading counts[] back; lines beyond
        // the promise are dropped (treated as unsampled) and flagged in cls[.][35], so nothing is ever written out of bounds
        int n = 0, dropped = 0;
        for (int y = 0; y < Y; ++y) {
            bool on = flag[ny == 1 ? 0 : y] != 0;
            if (on && cap > 0 && n >= cap) { on = false; dropped = 1; }
            pos[(size_t)bt * Y + y] = on ? n : -1;
            if (on) list[(size_t)bt * Y + n++] = y;
        }
        for (int i = n; i < Y; ++i) list[(size_t)bt * Y + i] = 0;
        counts[bt] = n;
        // classes ky mod 16: counting sort (stable), the non-empty classes and their bit mask
        int* seg = cls + (size_t)bt * kPlanClsInts;
        int* clist = seg + 17;
        int cnt[16];
        for (int k = 0; k < 16; ++k) cnt[k] = 0;
        const int* lst = list + (size_t)bt * Y;
        for (int i = 0; i < n; ++i) ++cnt[lst[i] & 15];
        int s = 0, m = 0, mask = 0;
        for (int k = 0; k < 16; ++k) {
            seg[k] = s;
            clist[k] = 0;
            if (cnt[k] > 0) { clist[m++] = k; mask |= 1 << k; }
            s += cnt[k];
            cnt[k] = seg[k];                         // running write position of the class
        }
        seg[16] = s;
        clist[16] = m; clist[17] = mask; clist[18] = dropped;
        int* ord = order + (size_t)bt * Y;
        for (int i = 0; i < n; ++i) ord[cnt[lst[i] & 15]++] = i;
        for (int i = n; i < Y; ++i) ord[i] = 0;
        flag[0] = n;                                 // hand n to the other threads
    }
    __syncthreads();
    {
        const int n = flag[0];
        const int* lst = list + (size_t)bt * Y;
        const int* ord = order + (size_t)bt * Y;
        float* wo = wq + (size_t)bt * Y * 8;
        for (int q = threadIdx.x; q < n; q += blockDim.x) {
            const int ky = lst[ord[q]];
            int m1 = ky % Y, m2 = (2 * ky) % Y;
            if (2 * m1 > Y) m1 -= Y;
            if (2 * m2 > Y) m2 -= Y;
            float s1, c1, s2, c2;
            sincospif((float)(2 * m1) / (float)Y, &s1, &c1);
            sincospif((float)(2 * m2) / (float)Y, &s2, &c2);
            float* o = wo + (size_t)q * 8;           // forward sign: w = cos - i sin
            o[0] = c1; o[1] = -s1; o[2] = s1; o[3] = c1;
            o[4] = c2; o[5] = -s2; o[6] = s2; o[7] = c2;
        }
    }
}

// ---------------------------------------------------------------------------- column passes
template <int RY>
struct ColLinesCfg {
    static constexpr int Y = 32 * RY;
    static constexpr int SW = ColCfg<RY>::SW, NT = ColCfg<RY>::NT, ITEMS = ColCfg<RY>::ITEMS, MINB = ColCfg<RY>::MINB;
    static constexpr size_t SLAB = (size_t)Y * SW;                   // complex elements of one slab
    static size_t tables() { return (size_t)Y * sizeof(float2) + (size_t)Y * sizeof(int); }
    static size_t smem_fwd(int e_staged) { return (size_t)(e_staged + 1) * SLAB * sizeof(float2) + tables(); }
    static size_t smem_adj(int E, int lmax) { return ((size_t)(1 + E) * SLAB + (size_t)lmax * SW) * sizeof(float2) + tables(); }
};

// asynchronous copy of the SW-column slab of one image (row stride Xs) into shared memory [Y][SW]
template <int RY>
DLSG_DEVICE void fetch_slab(float2* slab, const float2* img, int Xs, int rows, int tid) {
    constexpr int SW = ColLinesCfg<RY>::SW, NT = ColLinesCfg<RY>::NT;
#pragma unroll 4
    for (int q = tid; q < rows * (SW / 2); q += NT) {
        const int yy = q / (SW / 2), ch = q - yy * (SW / 2);
        cp_async16(slab + yy * SW + 2 * ch, img + (size_t)yy * Xs + 2 * ch);
    }
}

// twiddles of the radix-RY stage and the line table of this (b,t): sp[ky] = pos[ky] * mul, or -1 when ky is not sampled
template <int RY>
DLSG_DEVICE void fill_col_tables(float2* twy, int* sp, const int* pos, int mul, int tid) {
    constexpr int Y = 32 * RY, NT = ColLinesCfg<RY>::NT;
    for (int i = tid; i < Y; i += NT) {
        const int l = i / RY, k2 = i - l * RY;
        int m = (l * k2) % Y;
        if (2 * m > Y) m -= Y;
        float sn, cs;
        sincospif((float)(2 * m) / (float)Y, &sn, &cs);
        twy[i] = make_float2(cs, -sn);           // forward sign; the inverse uses the conjugate
        const int p = __ldg(pos + i);
        sp[i] = (p >= 0) ? p * mul : -1;
    }
}

// tl[b,c,t,i,kx] = DFT_Y( sum_e maps[b,e,c,:,kx] * x[b,e,t,:,kx] )[ky_i]   for the sampled lines of (b,t).
// CTA = (SW-column slab, t, b), all coils.  The coil's map slab arrives by cp.async in the work slab while the
// previous coil's radix-32 runs; XS: the x slabs are staged once per CTA (else read through L1/L2 per coil).
template <int RY, bool XS>
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(ColCfg<RY>::NT, DLSG_LINES_FWD_MINB(RY))
sense_cols_fwd_lines_kernel(const float2* __restrict__ x, const float2* __restrict__ maps,
                            float2* __restrict__ tl_out, LinePlan lp, SenseDims d) {
    using Cfg = ColLinesCfg<RY>;
    constexpr int ITEMS = Cfg::ITEMS, SW = Cfg::SW, Y = Cfg::Y;
    DLSG_DYN_SMEM(float2, smem);
    const int es = XS ? d.E : 0;
    float2* xs = smem;                           // [e][Y][SW]
    float2* work = smem + (size_t)es * Cfg::SLAB;
    float2* twy = work + Cfg::SLAB;
    int* sp = reinterpret_cast<int*>(twy + Y);
    const int tid = threadIdx.x, col = tid % SW, grp = tid / SW;

    const int ncolblk = d.X / SW;
    long long u = blockIdx.x;
    const int colblk = (int)(u % ncolblk); u /= ncolblk;
    const int tl = (int)(u % d.t_count);
    const int b = (int)(u / d.t_count);
    const int t = d.t_begin + tl;
    const int kx0 = colblk * SW, kx = kx0 + col;
    const int bt = b * d.T + t;
    if (__ldg(lp.counts + bt) == 0) return;      // nothing sampled in this frame (uniform over the CTA)

    if (XS)
        for (int e = 0; e < d.E; ++e) fetch_slab<RY>(xs + (size_t)e * Cfg::SLAB, x + idx_x(d, b, e, t, 0) + kx0, d.X, Y, tid);
    fetch_slab<RY>(work, maps + idx_maps(d, b, 0, 0, 0) + kx0, d.X, Y, tid);
    cp_async_commit();
    fill_col_tables<RY>(twy, sp, lp.pos + (size_t)bt * Y, d.X, tid);    // sp = element offset of the line in tl

#pragma unroll 1
    for (int c = 0; c < d.C; ++c) {
        cp_async_wait_group0();
        __syncthreads();                         // slab(s) landed, tables written
        // ---- phase A (in place): coil image column, radix-RY over y = l + 32 j, twiddle
DLSG_UNROLL(DLSG_LINES_FWD_ITEM_UNROLL)
        for (int i = 0; i < ITEMS; ++i) {
            const int l = grp + i * RY;
            if (l < 32) {
                float2 v[RY];
#pragma unroll
                for (int j = 0; j < RY; ++j) {
                    const int yy = l + 32 * j;
                    float2 a = cmul(work[yy * SW + col], XS ? xs[yy * SW + col] : __ldg(x + idx_x(d, b, 0, t, yy) + kx));
#pragma unroll 1
                    for (int e = 1; e < d.E; ++e)    // further map sets: through the cache (rare; keeps the code small)
                        a = cfma(__ldg(maps + idx_maps(d, b, e, c, yy) + kx),
                                 XS ? xs[(size_t)e * Cfg::SLAB + yy * SW + col] : __ldg(x + idx_x(d, b, e, t, yy) + kx), a);
                    v[j] = a;
                }
                XDft<RY, -1>::run(v);
#pragma unroll
                for (int k2 = 0; k2 < RY; ++k2)
                    work[(k2 * 32 + l) * SW + col] = (k2 == 0) ? v[0] : cmul(v[k2], twy[l * RY + k2]);
            }
        }
        __syncthreads();
        // ---- phase B: radix-32 over l for (k2 = grp, column); ky = RY*k1 + k2
        float2 r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = work[(grp * 32 + i) * SW + col];
        __syncthreads();                         // the work slab is free: the next coil's maps land during the radix-32
        if (c + 1 < d.C) fetch_slab<RY>(work, maps + idx_maps(d, b, 0, c + 1, 0) + kx0, d.X, Y, tid);
        cp_async_commit();
        XDft<32, -1>::run(r);
        float2* outp = tl_out + idx_lines(d, lp.lmax, b, c, t, 0) + kx;
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) {
            const int off = sp[RY * k1 + grp];
            if (off >= 0) outp[off] = r[k1];
        }
    }
}

// out_e[b,t,y,kx] = EPI( sum_c conj(maps[b,e,c,y,kx]) * IDFT_Y( scatter of tl[b,c,t,:,kx] to its ky lines )[y] )
// CTA = (SW-column slab, t, b), all coils.  The accumulator of map set 0 stays in registers; the coil's map slab is
// prefetched one coil ahead (cp.async) into its own buffer, the sampled lines of the next coil into a small one.
template <int RY>
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(ColCfg<RY>::NT, DLSG_LINES_ADJ_MINB(RY))
sense_cols_adj_lines_kernel(const float2* __restrict__ tl_in, const float2* __restrict__ maps, float2* __restrict__ out,
                            const float2* __restrict__ xin, const float2* __restrict__ aty,
                            float step_host, const float* __restrict__ step_dev, int epi, LinePlan lp, SenseDims d) {
    using Cfg = ColLinesCfg<RY>;
    constexpr int ITEMS = Cfg::ITEMS, SW = Cfg::SW, Y = Cfg::Y, NT = Cfg::NT;
    DLSG_DYN_SMEM(float2, smem);
    float2* work = smem;                                   // [Y][SW] exchange between the two radix stages
    float2* mbuf = smem + Cfg::SLAB;                       // [Y][SW] maps[b,0,c,:,slab] of the coil being accumulated
    float2* accx = mbuf + Cfg::SLAB;                       // [e-1][Y][SW] accumulators of map sets e >= 1 (thread-private slots)
    float2* cmp = accx + (size_t)(d.E - 1) * Cfg::SLAB;    // [lmax][SW] sampled lines of the coil being transformed
    float2* twy = cmp + (size_t)lp.lmax * SW;
    int* sp = reinterpret_cast<int*>(twy + Y);
    const int tid = threadIdx.x, col = tid % SW, grp = tid / SW;

    const int ncolblk = d.X / SW;
    long long u = blockIdx.x;
    const int colblk = (int)(u % ncolblk); u /= ncolblk;
    const int tl = (int)(u % d.t_count);
    const int b = (int)(u / d.t_count);
    const int t = d.t_begin + tl;
    const int kx0 = colblk * SW, kx = kx0 + col;
    const int bt = b * d.T + t;
    const int n = __ldg(lp.counts + bt);
    const float step = step_dev ? __ldg(step_dev) : step_host;

    auto fetch_cmp = [&](int c) {
        const float2* src = tl_in + idx_lines(d, lp.lmax, b, c, t, 0) + kx0;
        for (int q = tid; q < n * (SW / 2); q += NT) {
            const int i = q / (SW / 2), ch = q - i * (SW / 2);
            cp_async16(cmp + i * SW + 2 * ch, src + (size_t)i * d.X + 2 * ch);
        }
    };
    float2 acc0[32];
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) acc0[k1] = make_float2(0.f, 0.f);

    if (n > 0) {
        fetch_cmp(0);
        fetch_slab<RY>(mbuf, maps + idx_maps(d, b, 0, 0, 0) + kx0, d.X, Y, tid);
    }
    cp_async_commit();
    fill_col_tables<RY>(twy, sp, lp.pos + (size_t)bt * Y, SW, tid);     // sp = element offset of the line in cmp
    cp_async_wait_group0();
    __syncthreads();

#pragma unroll 1
    for (int c = 0; c < d.C && n > 0; ++c) {
        // ---- phase A: gather the sampled lines (zeros elsewhere), radix-RY over ky = l + 32 j, twiddle
DLSG_UNROLL(DLSG_LINES_ADJ_ITEM_UNROLL)
        for (int i = 0; i < ITEMS; ++i) {
            const int l = grp + i * RY;
            if (l < 32) {
                float2 v[RY];
                int any = -1;                        // stays negative when none of the RY lines l + 32 j is sampled
#pragma unroll
                for (int j = 0; j < RY; ++j) {
                    const int off = sp[l + 32 * j];
                    any &= off;
                    v[j] = (off >= 0) ? cmp[off + col] : make_float2(0.f, 0.f);
                }
                if (any < 0) {                       // uniform over the warp (same l for all its columns): the DFT of zeros
#pragma unroll
                    for (int k2 = 0; k2 < RY; ++k2) work[(k2 * 32 + l) * SW + col] = make_float2(0.f, 0.f);
                    continue;
                }
                XDft<RY, +1>::run(v);
#pragma unroll
                for (int k2 = 0; k2 < RY; ++k2)
                    work[(k2 * 32 + l) * SW + col] = (k2 == 0) ? v[0] : cmulc(v[k2], twy[l * RY + k2]);
            }
        }
        __syncthreads();
        if (c + 1 < d.C) fetch_cmp(c + 1);       // the line buffer has been consumed
        cp_async_commit();
        // ---- phase B: radix-32 over l for (k2 = grp, column); y = RY*k1 + k2
        float2 r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = work[(grp * 32 + i) * SW + col];
        XDft<32, +1>::run(r);
        cp_async_wait_group1();                  // this coil's maps (committed before the lines of the next coil)
        __syncthreads();
        {
            const float2* m0 = mbuf + grp * SW + col;    // y = RY*k1 + grp
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) acc0[k1] = cfmac(r[k1], m0[RY * k1 * SW], acc0[k1]);
        }
#pragma unroll 1
        for (int e = 1; e < d.E; ++e) {          // further map sets: maps through the cache (rare; keeps the code small)
            float2* ae = accx + (size_t)(e - 1) * Cfg::SLAB + grp * SW + col;
            const float2* me = maps + idx_maps(d, b, e, c, grp) + kx;
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) {
                const float2 prev = (c == 0) ? make_float2(0.f, 0.f) : ae[RY * k1 * SW];
                ae[RY * k1 * SW] = cfmac(r[k1], __ldg(me + (size_t)RY * k1 * d.X), prev);
            }
        }
        cp_async_wait_group0();
        __syncthreads();                         // map buffer and work slab are free, the next coil's lines have landed
        if (c + 1 < d.C) fetch_slab<RY>(mbuf, maps + idx_maps(d, b, 0, c + 1, 0) + kx0, d.X, Y, tid);
        cp_async_commit();
    }
    // ---- epilogue (each thread owns its outputs: y = RY*k1 + grp, column kx)
    // out = a * (acc - av) + bx * xv :  PLAIN (a, bx) = (1, 0);  DC (step, 1) with av = aty;  REG (1, step)
    const float ea = (epi == EPI_DC) ? step : 1.f;
    const float eb = (epi == EPI_DC) ? 1.f : (epi == EPI_REG ? step : 0.f);
    const size_t rstride = (size_t)RY * d.X;      // consecutive outputs of a thread are RY image rows apart
    // x and A^H y are read in batches of 16 rows with all loads of a batch in flight together (the accumulators
    // hold 64 registers, so the compiler would otherwise serialise one DRAM latency per row)
    auto finish16 = [&](const float2* accv, const float2* xp, const float2* ap, float2* op) {
        float2 xv[16], av[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) xv[i] = av[i] = make_float2(0.f, 0.f);
        if (epi != EPI_PLAIN) {
#pragma unroll
            for (int i = 0; i < 16; ++i) xv[i] = __ldg(xp + i * rstride);
        }
        if (epi == EPI_DC) {
#pragma unroll
            for (int i = 0; i < 16; ++i) av[i] = __ldg(ap + i * rstride);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i)
            op[i * rstride] = make_float2(fmaf(ea, accv[i].x - av[i].x, eb * xv[i].x), fmaf(ea, accv[i].y - av[i].y, eb * xv[i].y));
    };
    {
        const size_t g0 = idx_x(d, b, 0, t, grp) + kx;
        finish16(acc0, xin + g0, aty + g0, out + g0);
        finish16(acc0 + 16, xin + g0 + 16 * rstride, aty + g0 + 16 * rstride, out + g0 + 16 * rstride);
    }
#pragma unroll 1
    for (int e = 1; e < d.E; ++e) {
        const float2* ae = accx + (size_t)(e - 1) * Cfg::SLAB + grp * SW + col;
        const size_t g0 = idx_x(d, b, e, t, grp) + kx;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            float2 a16[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) a16[i] = (n > 0) ? ae[RY * (16 * h + i) * SW] : make_float2(0.f, 0.f);
            finish16(a16, xin + g0 + 16 * h * rstride, aty + g0 + 16 * h * rstride, out + g0 + 16 * h * rstride);
        }
    }
}

// ---------------------------------------------------------------------------- row pass on the sampled lines
//   MODE_FWD   : k[b,c,t,ky_i,:] = W*scale * DFT_X(tl[b,c,t,i,:])
//   MODE_ADJ   : tl[b,c,t,i,:]   = IDFT_X(W*scale * k[b,c,t,ky_i,:])      (src_lines: k already gathered to the tl layout, in place)
//   MODE_NORMAL: tl              = IDFT_X((W*scale)^2 * DFT_X(tl))            in place
// The flattened (coil, line) items of one (b, t) are walked in warp batches of NB = 32/RX rows by the 4 warps of
// `nsplit` CTAs (the split keeps small batches, e.g. the chunks of a pipelined reconstruction, at several waves).
// The raw rows of a warp's next batch arrive by cp.async in its staging buffer while the current batch is transformed;
// the weights of a batch are fetched together (one exposed latency per batch, not per row).
template <int RX>
struct RowLinesCfg {
    static constexpr int PER_WARP = RowCfg<RX>::BUF + RowCfg<RX>::NB * RowCfg<RX>::X;      // exchange buffer + staged rows (complex)
    static constexpr size_t SMEM = (size_t)RowCfg<RX>::WARPS * PER_WARP * sizeof(float2);
};

#ifndef DLSG_LINES_ROWS_NORMAL_MINB
#define DLSG_LINES_ROWS_NORMAL_MINB 2       // the normal-operator variant spills at the 168-register cap of 3 CTAs/SM: 2 CTAs (216 registers) measured 14 % faster
#endif
template <int RX, int MODE>
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(128, (MODE == MODE_NORMAL ? DLSG_LINES_ROWS_NORMAL_MINB : DLSG_LINES_ROWS_MINB))
sense_rows_lines_kernel(const float2* src, float2* dst, const float* __restrict__ w, WeightStrides ws, float scale,
                        LinePlan lp, SenseDims d, int src_lines, int nsplit) {
    using Cfg = RowCfg<RX>;
    constexpr int NB = Cfg::NB, X = Cfg::X;
    constexpr int DIR = (MODE == MODE_ADJ) ? +1 : -1;
    DLSG_DYN_SMEM(float2, smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* buf = smem + warp * RowLinesCfg<RX>::PER_WARP;
    float2* stg = buf + Cfg::BUF;                              // [r][X] raw rows of the batch being started
    const int split = (int)(blockIdx.x % nsplit);             // nsplit CTAs share the items of one (b, t)
    const int unit = (int)(blockIdx.x / nsplit);
    const int tl = unit % d.t_count, b = unit / d.t_count;
    const int t = d.t_begin + tl;
    const int bt = b * d.T + t;
    const int n = __ldg(lp.counts + bt);
    const int items = d.C * n;
    const int* list = lp.list + (size_t)bt * d.Y;
    const bool from_k = (MODE == MODE_ADJ) && !src_lines;

    float2 tw[RX];
    lane_twiddles<RX, DIR>(lane, tw);
    const float* wbt = w + (b * ws.b + t * ws.t + lane * ws.x);      // weights of this lane's first column of (b, t)

    // rows of a batch are consecutive (coil, line) items, line fastest: one division per batch, then an incremental walk
    struct Walk { int c, i; };
    auto walk_begin = [&](int it0) { Walk wk; wk.c = it0 / n; wk.i = it0 - wk.c * n; return wk; };
    auto walk_next = [&](Walk& wk) { if (++wk.i == n) { wk.i = 0; ++wk.c; } };

    auto stage_batch = [&](int it0) {
        const int nrows = min(NB, items - it0);
        Walk wk = walk_begin(it0);
        for (int r = 0; r < nrows; ++r, walk_next(wk)) {
            const float2* row = src + (from_k ? idx_k(d, b, wk.c, t, __ldg(list + wk.i)) : idx_lines(d, lp.lmax, b, wk.c, t, wk.i));
#pragma unroll
            for (int m = 0; m < (X / 2 + 31) / 32; ++m) {
                const int q = lane + 32 * m;
                if ((X / 2) % 32 == 0 || q < X / 2) cp_async16(stg + r * X + 2 * q, row + 2 * q);
            }
        }
        cp_async_commit();
    };
    // scale * W of the lane's RX columns of the rows [r0, r1) of the batch: all loads are issued before the first one is used
    auto load_weights = [&](int it0, int nrows, float (&wv)[NB][RX], int r0, int r1) {
        Walk wk = walk_begin(it0);
#pragma unroll
        for (int r = 0; r < NB; ++r) {
            if (r < nrows) {
                if (r >= r0 && r < r1) {
                    const float* wr = wbt + (wk.c * ws.c + __ldg(list + wk.i) * ws.y);
#pragma unroll
                    for (int j = 0; j < RX; ++j) wv[r][j] = __ldg(wr + 32 * j * ws.x);
                }
                walk_next(wk);
            }
        }
#pragma unroll
        for (int r = 0; r < NB; ++r)
            if (r >= r0 && r < r1) {
#pragma unroll
                for (int j = 0; j < RX; ++j) wv[r][j] *= scale;
            }
    };

    const int stride = nsplit * Cfg::WARPS * NB;
    int it0 = (split * Cfg::WARPS + warp) * NB;
    if (it0 < items) stage_batch(it0);
    for (; it0 < items; it0 += stride) {
        const int nrows = min(NB, items - it0);
        float wv[NB][RX];
        if (MODE == MODE_ADJ) load_weights(it0, nrows, wv, 0, NB);
        cp_async_wait_group0();
        __syncwarp();
#pragma unroll
        for (int r = 0; r < NB; ++r) {
            if (r < nrows) {
                float2 v[RX];
#pragma unroll
                for (int j = 0; j < RX; ++j) {
                    v[j] = stg[r * X + lane + 32 * j];
                    if (MODE == MODE_ADJ) v[j] = cscale(v[j], wv[r][j]);
                }
                XDft<RX, DIR>::run(v);
                float2* o = buf + (r * RX) * Cfg::S1 + lane;
#pragma unroll
                for (int k2 = 0; k2 < RX; ++k2) o[k2 * Cfg::S1] = (k2 == 0) ? v[0] : cmul(v[k2], tw[k2]);
            }
        }
        __syncwarp();                                          // staging buffer consumed, view 1 complete
        if (it0 + stride < items) stage_batch(it0 + stride);   // lands while the rest of this batch runs
        row_radix32<RX, DIR>(buf, lane, nrows);
        if (MODE == MODE_NORMAL) {
            // way back in place: rows in reverse order, because view 1 of row r starts behind the end of view 2 of row r-1
            // (34 RX r >= 33 RX r - RX) and only overlaps rows >= r, which have been consumed
            load_weights(it0, nrows, wv, 0, NB);
#pragma unroll
            for (int rr = 0; rr < NB; ++rr) {
                const int r = NB - 1 - rr;
                if (r < nrows) {
                    float2 v[RX];
#pragma unroll
                    for (int j = 0; j < RX; ++j) v[j] = cscale(buf[r * Cfg::S2 + lane + 32 * j], wv[r][j] * wv[r][j]);
                    __syncwarp();                              // every lane has read row r before its view 1 is written
                    XDft<RX, +1>::run(v);
                    float2* o = buf + (r * RX) * Cfg::S1 + lane;
#pragma unroll
                    for (int k2 = 0; k2 < RX; ++k2) o[k2 * Cfg::S1] = (k2 == 0) ? v[0] : cmulc(v[k2], tw[k2]);
                }
            }
            __syncwarp();
            row_radix32<RX, +1>(buf, lane, nrows);
        }
        if (MODE == MODE_FWD) load_weights(it0, nrows, wv, 0, NB);
        Walk wo = walk_begin(it0);
#pragma unroll
        for (int r = 0; r < NB; ++r) {
            if (r < nrows) {
                const int c = wo.c, i = wo.i;
                walk_next(wo);
                float2* o = dst + ((MODE == MODE_FWD) ? idx_k(d, b, c, t, __ldg(list + i)) : idx_lines(d, lp.lmax, b, c, t, i)) + lane;
                const float2* sv = buf + r * Cfg::S2 + lane;
#pragma unroll
                for (int j = 0; j < RX; ++j) o[32 * j] = (MODE == MODE_FWD) ? cscale(sv[32 * j], wv[r][j]) : sv[32 * j];
            }
        }
        __syncwarp();
    }
}

// kl[b,c,t,i,:] = k[b,c,t,ky_i,:]: copies the sampled lines of a k-space that may live in pinned (mapped) host memory,
// so that only those lines cross PCIe.  Persistent CTAs walk the (c, t, b) images; the copy is PCIe-bound (51 GB/s alone).
// Grid size measured inside the pipelined end-to-end step, where it shares the SMs with the DC kernels of the previous chunk:
// 8 / 16 / 24 / 48 / 96 / 148 / 296 CTAs -> 2560 / 3040 / 3100 / 3500 / 3885 / 3845 / 3640 slices/s.
#ifndef DLSG_GATHER_CTAS
#define DLSG_GATHER_CTAS 96
#endif
constexpr int kGatherCtas = DLSG_GATHER_CTAS;
#ifndef DLSG_GATHER_ILP
#define DLSG_GATHER_ILP 4
#endif
constexpr int kGatherIlp = DLSG_GATHER_ILP;    // 16-byte loads in flight per thread
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(256)
sense_gather_lines_kernel(const float2* __restrict__ k, float2* __restrict__ kl, LinePlan lp, SenseDims d) {
    const long long units = (long long)d.B * d.t_count * d.C;
    for (long long unit = blockIdx.x; unit < units; unit += gridDim.x) {
        long long u = unit;
        const int c = (int)(u % d.C); u /= d.C;
        const int tl = (int)(u % d.t_count);
        const int b = (int)(u / d.t_count);
        const int t = d.t_begin + tl;
        const int bt = b * d.T + t;
        const int n = __ldg(lp.counts + bt);
        const int* list = lp.list + (size_t)bt * d.Y;
        const float4* src = reinterpret_cast<const float4*>(k + idx_k(d, b, c, t, 0));
        float4* dst = reinterpret_cast<float4*>(kl + idx_lines(d, lp.lmax, b, c, t, 0));
        const int per_row = d.X / 2, total = n * per_row;
        for (int q0 = threadIdx.x; q0 < total; q0 += kGatherIlp * blockDim.x) {
            float4 v[kGatherIlp];
#pragma unroll
            for (int m = 0; m < kGatherIlp; ++m) {
                const int q = q0 + m * blockDim.x;
                if (q < total) {
                    const int i = q / per_row, ch = q - i * per_row;
                    v[m] = __ldg(src + (size_t)__ldg(list + i) * per_row + ch);
                }
            }
#pragma unroll
            for (int m = 0; m < kGatherIlp; ++m) {
                const int q = q0 + m * blockDim.x;
                if (q < total) dst[q] = v[m];
            }
        }
    }
}

// Forward operator: the ky lines of k that carry no weight are zero.  CTA = (c, t, b).
DLSG_KERNEL void DLSG_LAUNCH_BOUNDS(256)
sense_zero_unsampled_kernel(float2* __restrict__ k, LinePlan lp, SenseDims d) {
    long long u = blockIdx.x;
    const int c = (int)(u % d.C); u /= d.C;
    const int tl = (int)(u % d.t_count);
    const int b = (int)(u / d.t_count);
    const int t = d.t_begin + tl;
    const int* pos = lp.pos + (size_t)(b * d.T + t) * d.Y;
    float4* img = reinterpret_cast<float4*>(k + idx_k(d, b, c, t, 0));
    const int per_row = d.X / 2;                 // X is a multiple of 32 on this path
    for (int q = threadIdx.x; q < d.Y * per_row; q += blockDim.x)
        if (__ldg(pos + q / per_row) < 0) img[q] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace dlsg
