// C-ABI entry points of the SENSE operator (declared in include/dlsg_b200.h):
// argument checks, chunking of the L2-resident intermediate, kernel selection by
// axis length, launches.  No device allocation and no synchronisation on the normal paths; the only process-wide
// state is a launch counter (statistics for bench.py) and the cached driver entry point of the TMA descriptor encoder.
// One exception, documented in include/dlsg_b200.h: when the caller's workspace is smaller than the problem the dense
// path pipelines chunks on a temporary side stream (created and destroyed inside the call, never during graph capture).
#include "../../include/dlsg_b200.h"
#include "sense_kernels.cuh"
#include "sense_lines.cuh"
#include "sense_lines_p.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace dlsg {
namespace {

constexpr int kMaxDirectX = 4096;
constexpr int kMaxDirectY = 1536;
constexpr long long kTargetRowCtas = 148LL * 8;    // row pass: about two waves of 4-5 CTAs per SM
constexpr long long kTargetColCtas = 148LL * 4;
constexpr long long kTargetLineRowCtas = 148LL * 24;  // sampled-line row pass: ~6 waves of 4 CTAs per SM

inline int radix_of(int n) {   // R if n == 32*R with a compiled radix, else 0
    if (n % 32) return 0;
    const int r = n / 32;
#ifdef DLSG_FAST_BUILD
    return (r == 5 || r == 6) ? r : 0;
#else
    switch (r) { case 1: case 2: case 3: case 4: case 5: case 6: case 8: case 10: return r; default: return 0; }
#endif
}

// One `case R:` per compiled radix (DLSG_FAST_BUILD: cine radices only, for tuning builds; never shipped).
#define DLSG_RADIX_CASE(n, EXPR) case n: { constexpr int R = n; return EXPR; }
#ifdef DLSG_FAST_BUILD
#define DLSG_RADIX_CASES(EXPR) DLSG_RADIX_CASE(5, EXPR) DLSG_RADIX_CASE(6, EXPR)
#else
#define DLSG_RADIX_CASES(EXPR)                                                                         \
    DLSG_RADIX_CASE(1, EXPR) DLSG_RADIX_CASE(2, EXPR) DLSG_RADIX_CASE(3, EXPR) DLSG_RADIX_CASE(4, EXPR) \
    DLSG_RADIX_CASE(5, EXPR) DLSG_RADIX_CASE(6, EXPR) DLSG_RADIX_CASE(8, EXPR) DLSG_RADIX_CASE(10, EXPR)
#endif
#define DLSG_RADIX_SWITCH(r, EXPR) switch (r) { DLSG_RADIX_CASES(EXPR) default: return -1; }

struct Problem {          // what stays fixed over the chunk walk
    WeightStrides ws;
    float2* tmp;
    int ndims;
    float step;
    const float* step_dev;
    int epi;
    stream_t stream;
};

#define DLSG_TRY(expr) do { if ((expr) != 0) return DLSG_ELAUNCH; } while (0)

// ---- row pass: work split ----------------------------------------------------------------
// One CTA (4 warps) per (slice, image row, frame chunk).  The frame chunk is a multiple of 4 so that
// every warp owns whole frames; coil groups are sized so that the staged map rows stay <= 24 KB.
RowPlan make_row_plan(const SenseDims& d, int e_staged) {
    RowPlan p;
    const long long base = (long long)d.B * d.Y;
    long long want = (kTargetRowCtas + base - 1) / base;
    const long long max_chunks = (d.t_count + 3) / 4;
    if (want > max_chunks) want = max_chunks;
    if (want < 1) want = 1;
    int tchunk = (int)((d.t_count + want - 1) / want);
    if (d.t_count >= 4) tchunk = ((tchunk + 3) / 4) * 4;
    p.tchunk = tchunk;
    p.nchunks = (d.t_count + tchunk - 1) / tchunk;
    long long cg = 24576 / ((long long)e_staged * d.X * (long long)sizeof(float2));
    if (cg < 1) cg = 1;
    if (cg > d.C) cg = d.C;
    p.cgroup = (int)cg;
    return p;
}

// ---- row pass, forward --------------------------------------------------------------
template <int RX>
int launch_rows_fwd(const float2* x, const float2* maps, float2* tmp, const SenseDims& d, stream_t s) {
    using Cfg = RowCfg<RX>;
    const RowPlan plan = make_row_plan(d, d.E);
    const unsigned grid = (unsigned)((long long)d.B * d.Y * plan.nchunks);
    const size_t smem = Cfg::smem_bytes(d.E, plan.cgroup, false);
    if (d.E == 1)
        return launch(sense_rows_fwd_kernel<RX, 1>, dim3(grid), dim3(32 * Cfg::WARPS), smem, s, x, maps, tmp, d, plan);
    if (d.E == 2)
        return launch(sense_rows_fwd_kernel<RX, 2>, dim3(grid), dim3(32 * Cfg::WARPS), smem, s, x, maps, tmp, d, plan);
    return launch(sense_rows_fwd_kernel<RX, 0>, dim3(grid), dim3(32 * Cfg::WARPS), smem, s, x, maps, tmp, d, plan);
}

int rows_fwd(const Problem& p, const float2* x, const float2* maps, const float* w, float2* k,
             const SenseDims& d, float scale, int final1d) {
    const int rx = (p.ndims == 2) ? radix_of(d.X) : 0;
    switch (rx) { DLSG_RADIX_CASES((launch_rows_fwd<R>(x, maps, p.tmp, d, p.stream))) default: break; }
    const size_t smem = (size_t)2 * d.X * sizeof(float2);
    const unsigned grid = (unsigned)((long long)d.B * d.t_count * d.Y);
    float2* dst = (final1d == 1) ? k : p.tmp;
    return launch(sense_rows_fwd_direct_kernel, dim3(grid), dim3(256), smem, p.stream,
                  x, maps, dst, w, p.ws, scale, final1d, d);
}

// ---- row pass, adjoint --------------------------------------------------------------
template <int RX, int EMAX>
int launch_rows_adj(const float2* maps, float2* out, const float2* xin, const float2* aty,
                    const Problem& p, const SenseDims& d, int e0) {
    using Cfg = RowCfg<RX>;
    const RowPlan plan = make_row_plan(d, EMAX);
    const unsigned grid = (unsigned)((long long)d.B * d.Y * plan.nchunks);
    const size_t smem = Cfg::smem_bytes(EMAX, plan.cgroup, true);
    return launch(sense_rows_adj_kernel<RX, EMAX>, dim3(grid), dim3(32 * Cfg::WARPS), smem, p.stream,
                  (const float2*)p.tmp, maps, out, xin, aty, p.step, p.step_dev, p.epi, d, plan, e0);
}

template <int RX>
int rows_adj_groups(const float2* maps, float2* out, const float2* xin, const float2* aty,
                    const Problem& p, const SenseDims& d) {
    for (int e0 = 0; e0 < d.E; e0 += 2) {
        if (d.E - e0 >= 2) DLSG_TRY((launch_rows_adj<RX, 2>(maps, out, xin, aty, p, d, e0)));
        else               DLSG_TRY((launch_rows_adj<RX, 1>(maps, out, xin, aty, p, d, e0)));
    }
    return 0;
}

int rows_adj(const Problem& p, const float2* src_k, const float2* maps, const float* w, float2* out,
             const float2* xin, const float2* aty, const SenseDims& d, float scale, int final1d) {
    const int rx = (p.ndims == 2) ? radix_of(d.X) : 0;
    switch (rx) { DLSG_RADIX_CASES((rows_adj_groups<R>(maps, out, xin, aty, p, d))) default: break; }
    const size_t smem = (size_t)(2 + d.E) * d.X * sizeof(float2);
    const unsigned grid = (unsigned)((long long)d.B * d.t_count * d.Y);
    const float2* src = (final1d == 1) ? src_k : (const float2*)p.tmp;
    return launch(sense_rows_adj_direct_kernel, dim3(grid), dim3(256), smem, p.stream,
                  src, maps, out, xin, aty, p.step, p.step_dev, p.epi, w, p.ws, scale, final1d, d);
}

// ---- column pass --------------------------------------------------------------------
template <int RY, int MODE>
int launch_cols(const float2* src, float2* dst, const float* w, const Problem& p, const SenseDims& d, float scale) {
    using Cfg = ColCfg<RY>;
    const long long base = (long long)(d.X / Cfg::SW) * d.t_count * d.B;
    long long want = (kTargetColCtas + base - 1) / base;
    if (want < 1) want = 1;
    if (want > d.C) want = d.C;
    const int cchunk = (int)((d.C + want - 1) / want);
    const int ncchunks = (d.C + cchunk - 1) / cchunk;
    const unsigned grid = (unsigned)(base * ncchunks);
    return launch(sense_cols_kernel<RY, MODE>, dim3(grid), dim3(Cfg::NT), Cfg::SMEM, p.stream,
                  src, dst, w, p.ws, scale, d, cchunk, ncchunks);
}

template <int MODE>
int cols(const Problem& p, const float2* src, float2* dst, const float* w, const SenseDims& d, float scale) {
    // the fast column kernel walks 32-column slabs, so it also needs X % 32 == 0
    const int ry = (d.X % 32 == 0) ? radix_of(d.Y) : 0;
    switch (ry) { DLSG_RADIX_CASES((launch_cols<R, MODE>(src, dst, w, p, d, scale))) default: break; }
    const size_t smem = ((size_t)2 * d.Y * kDirectColW + d.Y) * sizeof(float2);
    const int ncolblk = (d.X + kDirectColW - 1) / kDirectColW;
    const unsigned grid = (unsigned)((long long)d.B * d.C * d.t_count * ncolblk);
    return launch(sense_cols_direct_kernel<MODE>, dim3(grid), dim3(256), smem, p.stream, src, dst, w, p.ws, scale, d);
}

// ---- problem validation and chunk walk ------------------------------------------------
enum Op { OP_FWD, OP_ADJ, OP_NORMAL };

size_t per_bt_bytes(int C, int Y, int X) { return (size_t)C * Y * X * sizeof(float2); }

int check_sizes(int B, int E, int C, int T, int Y, int X, int ndims) {
    if (B <= 0 || E <= 0 || C <= 0 || T <= 0 || Y <= 0 || X <= 0) return DLSG_EINVAL;
    if (ndims != 1 && ndims != 2) return DLSG_EINVAL;
    const bool fast_x = (ndims == 2) && radix_of(X);
    if (!fast_x && X > kMaxDirectX) return DLSG_EUNSUPPORTED;
    if (ndims == 2) {
        const bool fast_y = radix_of(Y) && (X % 32 == 0);
        if (!fast_y && Y > kMaxDirectY) return DLSG_EUNSUPPORTED;
    }
    return DLSG_OK;
}

// in : x (forward / normal) or k (adjoint);  out : k (forward) or x (adjoint / normal)
int run(Op op, const float2* in, const float2* maps, const float* w, const long long* w_strides,
        float2* out, const float2* aty, float step, const float* step_dev, int epi,
        void* workspace, size_t workspace_bytes,
        int B, int E, int C, int T, int Y, int X, int ndims, void* stream) {
    const int st = check_sizes(B, E, C, T, Y, X, ndims);
    if (st != DLSG_OK) return st;
    if (!in || !out) return DLSG_EINVAL;          // maps may be NULL: unit sensitivities
    if (w && !w_strides) return DLSG_EINVAL;
    if (epi == EPI_DC && !aty) return DLSG_EINVAL;

    Problem p{};
    if (w) p.ws = WeightStrides{w_strides[0], w_strides[1], w_strides[2], w_strides[3], w_strides[4]};
    p.tmp = static_cast<float2*>(workspace);
    p.ndims = ndims;
    p.step = step;
    p.step_dev = step_dev;
    p.epi = epi;
    p.stream = static_cast<stream_t>(stream);

    const size_t per_bt = per_bt_bytes(C, Y, X);
    const bool need_tmp = (ndims == 2) || (op == OP_NORMAL);
    if (need_tmp && (!workspace || workspace_bytes < per_bt)) return DLSG_EWORKSPACE;
    const double npts = (ndims == 2) ? (double)Y * (double)X : (double)X;
    const float scale = (float)(1.0 / sqrt(npts));      // norm='ortho' (transforms.py:18)
    const size_t x_slice = (size_t)E * T * Y * X, m_slice = (size_t)E * C * Y * X, k_slice = (size_t)C * T * Y * X;

    // A chunk = (slices b0..b0+bc, frames t0..t0+tc) whose row-transformed intermediate fits one workspace slot.
    // stage 1 fills the slot, stage 2 drains it.
    struct Chunk { long long b0; int bc, t0, tc; };
    auto stage1 = [&](const Chunk& c, const Problem& q) -> int {
        const SenseDims d{c.bc, E, C, T, Y, X, c.t0, c.tc};
        const float2* maps_b = maps ? maps + c.b0 * m_slice : nullptr;
        const float* w_b = w ? w + c.b0 * q.ws.b : nullptr;
        if (op == OP_ADJ) {
            const float2* k_b = in + c.b0 * k_slice;
            if (ndims == 2) return cols<MODE_ADJ>(q, k_b, q.tmp, w_b, d, scale);
            return rows_adj(q, k_b, maps_b, w_b, out + c.b0 * x_slice, nullptr, nullptr, d, scale, 1);
        }
        const float2* x_b = in + c.b0 * x_slice;
        if (op == OP_FWD && ndims == 1) return rows_fwd(q, x_b, maps_b, w_b, out + c.b0 * k_slice, d, scale, 1);
        return rows_fwd(q, x_b, maps_b, w_b, nullptr, d, scale, (ndims == 2) ? 0 : 2);
    };
    auto stage2 = [&](const Chunk& c, const Problem& q) -> int {
        const SenseDims d{c.bc, E, C, T, Y, X, c.t0, c.tc};
        const float2* maps_b = maps ? maps + c.b0 * m_slice : nullptr;
        const float* w_b = w ? w + c.b0 * q.ws.b : nullptr;
        if (op == OP_FWD) return (ndims == 2) ? cols<MODE_FWD>(q, q.tmp, out + c.b0 * k_slice, w_b, d, scale) : 0;
        if (op == OP_ADJ)
            return (ndims == 2) ? rows_adj(q, nullptr, maps_b, w_b, out + c.b0 * x_slice, nullptr, nullptr, d, scale, 0) : 0;
        const float2* x_b = in + c.b0 * x_slice;
        const float2* aty_b = aty ? aty + c.b0 * x_slice : nullptr;
        if (ndims == 2 && cols<MODE_NORMAL>(q, q.tmp, q.tmp, w_b, d, scale) != 0) return -1;
        return rows_adj(q, nullptr, maps_b, w_b, out + c.b0 * x_slice, x_b, aty_b, d, scale, (ndims == 2) ? 0 : 2);
    };
    auto chunk_shape = [&](size_t slot_bytes, long long& bchunk, long long& tcount) {
        bchunk = B; tcount = T;
        if (!need_tmp) return;
        const long long max_bt = (long long)(slot_bytes / per_bt);
        if (max_bt >= T) { bchunk = max_bt / T; if (bchunk > B) bchunk = B; tcount = T; }
        else             { bchunk = 1; tcount = max_bt; }
    };
    auto count_chunks = [&](long long bchunk, long long tcount) {
        return ((B + bchunk - 1) / bchunk) * ((T + tcount - 1) / tcount);
    };

    long long bchunk, tcount;
    chunk_shape(workspace_bytes, bchunk, tcount);
#if defined(__CUDACC__)
    // Pipelined mode: when the workspace cannot hold the whole problem, split it into two slots and let the drain
    // of chunk k (second stream) overlap the fill of chunk k+1, so that a slot sized to a fraction of the 126 MB L2
    // is re-read from L2 instead of HBM and launch gaps / partial waves of one stage are covered by the other.
    // (not while the caller's stream is being captured into a CUDA graph: the side stream would be destroyed inside the
    // capture; the serial chunk walk below is used instead)
    cudaStreamCaptureStatus capture = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(p.stream, &capture) != cudaSuccess) { (void)cudaGetLastError(); capture = cudaStreamCaptureStatusActive; }
    if (capture == cudaStreamCaptureStatusNone && need_tmp && count_chunks(bchunk, tcount) > 1 && workspace_bytes / 2 >= per_bt) {
        long long bch2, tc2;
        chunk_shape(workspace_bytes / 2, bch2, tc2);
        if (count_chunks(bch2, tc2) >= 3) {
            const size_t slot_elems = (size_t)bch2 * C * tc2 * Y * X;
            cudaStream_t sa = p.stream, sb = nullptr;
            cudaEvent_t ev1[2] = {nullptr, nullptr}, ev2[2] = {nullptr, nullptr};
            bool ok = cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking) == cudaSuccess;
            for (int i = 0; i < 2 && ok; ++i)
                ok = cudaEventCreateWithFlags(&ev1[i], cudaEventDisableTiming) == cudaSuccess &&
                     cudaEventCreateWithFlags(&ev2[i], cudaEventDisableTiming) == cudaSuccess;
            int rc = ok ? 0 : -1;
            long long k = 0;
            bool used[2] = {false, false};
            for (long long b0 = 0; b0 < B && rc == 0; b0 += bch2) {
                const int bc = (int)((B - b0 < bch2) ? (B - b0) : bch2);
                for (long long t0 = 0; t0 < T && rc == 0; t0 += tc2, ++k) {
                    const int tc = (int)((T - t0 < tc2) ? (T - t0) : tc2);
                    const int slot = (int)(k & 1);
                    const Chunk c{b0, bc, (int)t0, tc};
                    Problem qa = p, qb = p;
                    qa.tmp = qb.tmp = p.tmp + (size_t)slot * slot_elems;
                    qa.stream = sa; qb.stream = sb;
                    // a failed record / wait would turn into a silent race on the slot: every call is checked
                    if (used[slot] && cudaStreamWaitEvent(sa, ev2[slot], 0) != cudaSuccess) { rc = -1; break; }   // the slot has been drained
                    rc = stage1(c, qa);
                    if (rc == 0 && (cudaEventRecord(ev1[slot], sa) != cudaSuccess || cudaStreamWaitEvent(sb, ev1[slot], 0) != cudaSuccess)) rc = -1;
                    if (rc == 0) rc = stage2(c, qb);
                    if (rc == 0 && cudaEventRecord(ev2[slot], sb) != cudaSuccess) rc = -1;
                    used[slot] = (rc == 0);
                }
            }
            for (int i = 0; i < 2; ++i)
                if (used[i] && cudaStreamWaitEvent(sa, ev2[i], 0) != cudaSuccess) rc = -1;   // join: later work on the caller's stream sees the results
            if (rc != 0) cudaStreamSynchronize(sb);                         // error path: nothing of the side stream may outlive this call
            for (int i = 0; i < 2; ++i) { if (ev1[i]) cudaEventDestroy(ev1[i]); if (ev2[i]) cudaEventDestroy(ev2[i]); }
            if (sb) cudaStreamDestroy(sb);                                    // deferred by the runtime until its work has drained
            return rc == 0 ? DLSG_OK : DLSG_ELAUNCH;
        }
    }
#endif
    for (long long b0 = 0; b0 < B; b0 += bchunk) {
        const int bc = (int)((B - b0 < bchunk) ? (B - b0) : bchunk);
        for (long long t0 = 0; t0 < T; t0 += tcount) {
            const int tc = (int)((T - t0 < tcount) ? (T - t0) : tcount);
            const Chunk c{b0, bc, (int)t0, tc};
            DLSG_TRY(stage1(c, p));
            DLSG_TRY(stage2(c, p));
        }
    }
    return DLSG_OK;
}

// ---- sampled-line path (sense_lines.cuh) ---------------------------------------------------

constexpr size_t kSmemTwoPerSm = 113 * 1024;     // largest dynamic shared memory that still lets two CTAs share an SM
constexpr size_t kSmemMax = 227 * 1024;

template <int RY>
int launch_cols_fwd_lines(const float2* x, const float2* maps, float2* tl, const LinePlan& lp, const SenseDims& d, stream_t s) {
    using Cfg = ColLinesCfg<RY>;
    const unsigned grid = (unsigned)((long long)(d.X / Cfg::SW) * d.t_count * d.B);
#ifndef DLSG_LINES_FWD_NO_XS
    if (Cfg::smem_fwd(d.E) <= kSmemTwoPerSm)
        return launch(sense_cols_fwd_lines_kernel<RY, true>, dim3(grid), dim3(Cfg::NT), Cfg::smem_fwd(d.E), s, x, maps, tl, lp, d);
#endif
    return launch(sense_cols_fwd_lines_kernel<RY, false>, dim3(grid), dim3(Cfg::NT), Cfg::smem_fwd(0), s, x, maps, tl, lp, d);
}
bool pcols_supported(int E, int Y, int X, int lmax);
inline bool pcols_disabled();
int pcols_fwd(const float2* x, const float2* maps, float2* tl, const LinePlan& lp, const SenseDims& d, stream_t s);
int pcols_adj(const float2* tl, const float2* maps, float2* out, const float2* xin, const float2* aty,
              float step, const float* step_dev, int epi, const LinePlan& lp, const SenseDims& d, stream_t s);

int cols_fwd_lines(const float2* x, const float2* maps, float2* tl, const LinePlan& lp, const SenseDims& d, stream_t s) {
    if (!pcols_disabled() && pcols_supported(d.E, d.Y, d.X, lp.lmax)) return pcols_fwd(x, maps, tl, lp, d, s);
    DLSG_RADIX_SWITCH(radix_of(d.Y), (launch_cols_fwd_lines<R>(x, maps, tl, lp, d, s)))
}

template <int RY>
int launch_cols_adj_lines(const float2* tl, const float2* maps, float2* out, const float2* xin, const float2* aty,
                          float step, const float* step_dev, int epi, const LinePlan& lp, const SenseDims& d, stream_t s) {
    using Cfg = ColLinesCfg<RY>;
    const unsigned grid = (unsigned)((long long)(d.X / Cfg::SW) * d.t_count * d.B);
    return launch(sense_cols_adj_lines_kernel<RY>, dim3(grid), dim3(Cfg::NT), Cfg::smem_adj(d.E, lp.lmax), s,
                  tl, maps, out, xin, aty, step, step_dev, epi, lp, d);
}
int cols_adj_lines(const float2* tl, const float2* maps, float2* out, const float2* xin, const float2* aty,
                   float step, const float* step_dev, int epi, const LinePlan& lp, const SenseDims& d, stream_t s) {
    if (!pcols_disabled() && pcols_supported(d.E, d.Y, d.X, lp.lmax))
        return pcols_adj(tl, maps, out, xin, aty, step, step_dev, epi, lp, d, s);
    DLSG_RADIX_SWITCH(radix_of(d.Y), (launch_cols_adj_lines<R>(tl, maps, out, xin, aty, step, step_dev, epi, lp, d, s)))
}
size_t cols_adj_lines_smem(int E, int Y, int lmax) {     // ColLinesCfg<RY>::smem_adj without the template argument
    const size_t sw = (radix_of(Y) <= 6) ? 32 : 16;
    return ((size_t)(1 + E) * Y * sw + (size_t)lmax * sw) * sizeof(float2) + (size_t)Y * (sizeof(float2) + sizeof(int));
}

template <int RX, int MODE>
int launch_rows_lines(const float2* src, float2* dst, const float* w, const WeightStrides& ws, float scale,
                      const LinePlan& lp, const SenseDims& d, stream_t s, int src_lines) {
    using Cfg = RowCfg<RX>;
    const size_t smem = RowLinesCfg<RX>::SMEM;
    const long long units = (long long)d.B * d.t_count;
    const long long batches = ((long long)d.C * lp.lmax + Cfg::NB - 1) / Cfg::NB;          // per (b, t), upper bound
    long long nsplit = (kTargetLineRowCtas + units - 1) / units;
    const long long max_split = (batches + Cfg::WARPS - 1) / Cfg::WARPS;                   // at least one batch per warp
    if (nsplit > max_split) nsplit = max_split;
    if (nsplit < 1) nsplit = 1;
    return launch(sense_rows_lines_kernel<RX, MODE>, dim3((unsigned)(units * nsplit)), dim3(32 * Cfg::WARPS), smem, s,
                  src, dst, w, ws, scale, lp, d, src_lines, (int)nsplit);
}
template <int MODE>
int rows_lines(const float2* src, float2* dst, const float* w, const WeightStrides& ws, float scale,
               const LinePlan& lp, const SenseDims& d, stream_t s, int src_lines = 0) {
    DLSG_RADIX_SWITCH(radix_of(d.X), (launch_rows_lines<R, MODE>(src, dst, w, ws, scale, lp, d, s, src_lines)))
}

// ---- second-generation column kernels (sense_lines_p.cuh): Y = 16 * R2, E <= 2
inline int pradix_of(int y) {      // R2 if y == 16*R2 with a compiled R2, else 0
    if (y % 16) return 0;
    const int r = y / 16;
#ifdef DLSG_FAST_BUILD
    return (r == 10 || r == 12) ? r : 0;
#else
    switch (r) { case 8: case 10: case 12: case 16: case 20: return r; default: return 0; }
#endif
}
#define DLSG_PRADIX_CASE(n, EXPR) case n: { constexpr int R2 = n; return EXPR; }
#ifdef DLSG_FAST_BUILD
#define DLSG_PRADIX_SWITCH(r, EXPR) switch (r) { DLSG_PRADIX_CASE(10, EXPR) DLSG_PRADIX_CASE(12, EXPR) default: return -1; }
#else
#define DLSG_PRADIX_SWITCH(r, EXPR) switch (r) { DLSG_PRADIX_CASE(8, EXPR) DLSG_PRADIX_CASE(10, EXPR) DLSG_PRADIX_CASE(12, EXPR) \
                                                  DLSG_PRADIX_CASE(16, EXPR) DLSG_PRADIX_CASE(20, EXPR) default: return -1; }
#endif

// DLSG_LINES_KERNELS=v1 keeps the first-generation column kernels (A/B measurements, tests of both families)
inline bool pcols_disabled() {
    const char* v = std::getenv("DLSG_LINES_KERNELS");
    return v && std::strcmp(v, "v1") == 0;
}
// buffers of the exchange (xb) in the forward kernel; one map slab (two measured slower: 2 CTAs/SM instead of 3).
// DLSG_PCOLS_XB=2 selects the double-buffered exchange (one CTA barrier per coil; measured equal)
inline void pcols_fwd_cfg(int E, int Y, int lmax, int& mb, int& xb) {
    mb = 1; xb = 1;
    if (const char* v = std::getenv("DLSG_PCOLS_XB")) xb = (v[0] == '2') ? 2 : 1;
    const size_t slab = (size_t)16 * Y * sizeof(float2);
    const size_t tables = 16 + (size_t)lmax * (2 * sizeof(float4) + 2 * sizeof(int)) + 36 * sizeof(int);
    if ((size_t)(E * mb + xb) * slab + tables > kSmemMax) xb = 1;
}
inline int pcols_adj_xb(int E, int Y, int lmax) {
    int xb = 2;
    if (const char* v = std::getenv("DLSG_PCOLS_ADJ_XB")) xb = (v[0] == '2') ? 2 : 1;
    const size_t slab = (size_t)16 * Y * sizeof(float2);
    const size_t rest = 16 + (size_t)2 * lmax * 16 * sizeof(float2) + (size_t)lmax * (2 * sizeof(float4) + 2 * sizeof(int)) + 36 * sizeof(int);
    if ((size_t)(2 * E + xb) * slab + rest > kSmemMax) xb = 1;
    return xb;
}
bool pcols_supported(int E, int Y, int X, int lmax) {
    if (E < 1 || E > 2 || !pradix_of(Y) || X % 16) return false;
    const size_t slab = (size_t)16 * Y * sizeof(float2);
    const size_t tables = 16 + (size_t)lmax * (2 * sizeof(float4) + 2 * sizeof(int)) + 36 * sizeof(int);
    const size_t fwd = (size_t)(E + 1) * slab + tables;
    const size_t adj = (size_t)(2 * E + 1) * slab + (size_t)2 * lmax * 16 * sizeof(float2) + tables;
    return fwd <= kSmemMax && adj <= kSmemMax;
}

// TMA descriptor of the coil maps of one chunk seen as a row-major fp32 matrix [rows = B*E*C*Y][2X]; a tile is
// box_rows rows of 16 complex columns
#if defined(__CUDACC__)
}  // namespace
}  // namespace dlsg
#include <cudaTypedefs.h>
namespace dlsg {
namespace {
bool make_tile_map(TileMap* tm, const float2* base, long long rows, int X, int box_rows) {
    static const PFN_cuTensorMapEncodeTiled_v12000 encode = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }();
    if (!encode) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)2 * X, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)X * sizeof(float2)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estride[2] = {1u, 1u};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float2*>(base), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
#else
bool make_tile_map(TileMap* tm, const float2* base, long long rows, int X, int box_rows) {
    tm->base = reinterpret_cast<const float*>(base); tm->pitch = 2LL * X; tm->rows = rows; tm->box_cols = 32; tm->box_rows = box_rows;
    tm->swizzle128 = 0;
    return true;
}
#endif

template <int R2, int E>
int launch_pcols_fwd_e(const float2* x, const TileMap& tm, float2* tl, const LinePlan& lp, const SenseDims& d, stream_t s) {
    using Cfg = PColCfg<R2>;
    const unsigned grid = (unsigned)((long long)(d.X / 16) * d.t_count * d.B);
    int mb, xb;
    pcols_fwd_cfg(E, d.Y, lp.lmax, mb, xb);
    const size_t smem = Cfg::smem_fwd(E, lp.lmax, mb, xb);
    if (xb == 2) return launch(sense_cols_fwd_lines_p_kernel<R2, E, 1, 2>, dim3(grid), dim3(Cfg::NT), smem, s, x, tm, tl, lp, d);
    return launch(sense_cols_fwd_lines_p_kernel<R2, E, 1, 1>, dim3(grid), dim3(Cfg::NT), smem, s, x, tm, tl, lp, d);
}
template <int R2>
int launch_pcols_fwd(const float2* x, const float2* maps, float2* tl, const LinePlan& lp, const SenseDims& d, stream_t s) {
    TileMap tm;
    if (!make_tile_map(&tm, maps, (long long)d.B * d.E * d.C * d.Y, d.X, d.Y / PColCfg<R2>::NBOX)) return -1;
    if (d.E == 1) return launch_pcols_fwd_e<R2, 1>(x, tm, tl, lp, d, s);
    return launch_pcols_fwd_e<R2, 2>(x, tm, tl, lp, d, s);
}
template <int R2, int E>
int launch_pcols_adj_e(const float2* tl, const TileMap& tm, float2* out, const float2* xin, const float2* aty,
                       float step, const float* step_dev, int epi, const LinePlan& lp, const SenseDims& d, stream_t s) {
    using Cfg = PColCfg<R2>;
    const unsigned grid = (unsigned)((long long)(d.X / 16) * d.t_count * d.B);
    const int xb = pcols_adj_xb(E, d.Y, lp.lmax);
    const size_t smem = Cfg::smem_adj(E, lp.lmax, xb);
    if (xb == 2)
        return launch(sense_cols_adj_lines_p_kernel<R2, E, 2>, dim3(grid), dim3(Cfg::NT), smem, s, tl, tm, out, xin, aty, step, step_dev, epi, lp, d);
    return launch(sense_cols_adj_lines_p_kernel<R2, E, 1>, dim3(grid), dim3(Cfg::NT), smem, s, tl, tm, out, xin, aty, step, step_dev, epi, lp, d);
}
template <int R2>
int launch_pcols_adj(const float2* tl, const float2* maps, float2* out, const float2* xin, const float2* aty,
                     float step, const float* step_dev, int epi, const LinePlan& lp, const SenseDims& d, stream_t s) {
    TileMap tm;
    if (!make_tile_map(&tm, maps, (long long)d.B * d.E * d.C * d.Y, d.X, d.Y / PColCfg<R2>::NBOX)) return -1;
    if (d.E == 1) return launch_pcols_adj_e<R2, 1>(tl, tm, out, xin, aty, step, step_dev, epi, lp, d, s);
    return launch_pcols_adj_e<R2, 2>(tl, tm, out, xin, aty, step, step_dev, epi, lp, d, s);
}

int pcols_fwd(const float2* x, const float2* maps, float2* tl, const LinePlan& lp, const SenseDims& d, stream_t s) {
    DLSG_PRADIX_SWITCH(pradix_of(d.Y), (launch_pcols_fwd<R2>(x, maps, tl, lp, d, s)))
}
int pcols_adj(const float2* tl, const float2* maps, float2* out, const float2* xin, const float2* aty,
              float step, const float* step_dev, int epi, const LinePlan& lp, const SenseDims& d, stream_t s) {
    DLSG_PRADIX_SWITCH(pradix_of(d.Y), (launch_pcols_adj<R2>(tl, maps, out, xin, aty, step, step_dev, epi, lp, d, s)))
}

// row pass of the normal operator: second-generation kernel for X = 16 * RX (sense_lines_p.cuh), else the first generation
template <int RX>
int launch_prows_normal(float2* tl, const float* w, const WeightStrides& ws, float scale, const LinePlan& lp, const SenseDims& d, stream_t s) {
    using Cfg = PRowCfg<RX>;
    const long long units = (long long)d.B * d.t_count;
    const long long rounds = ((long long)d.C * lp.lmax + 1) / 2;                 // warp rounds per (b, t), upper bound
    long long nsplit = (148LL * 5 * 4 + units - 1) / units;                       // about four waves of 5 CTAs per SM
    const long long max_split = (rounds + Cfg::WARPS - 1) / Cfg::WARPS;
    if (nsplit > max_split) nsplit = max_split;
    if (nsplit < 1) nsplit = 1;
    return launch(sense_rows_normal_lines_p_kernel<RX>, dim3((unsigned)(units * nsplit)), dim3(32 * Cfg::WARPS), Cfg::smem(), s,
                  tl, w, ws, scale, lp, d, (int)nsplit);
}
int rows_normal_lines(float2* tl, const float* w, const WeightStrides& ws, float scale, const LinePlan& lp, const SenseDims& d, stream_t s) {
    if (!pcols_disabled()) {
        switch (pradix_of(d.X)) {
            DLSG_PRADIX_CASE(10, (launch_prows_normal<10>(tl, w, ws, scale, lp, d, s)))
            DLSG_PRADIX_CASE(12, (launch_prows_normal<12>(tl, w, ws, scale, lp, d, s)))
#ifndef DLSG_FAST_BUILD
            DLSG_PRADIX_CASE(8, (launch_prows_normal<8>(tl, w, ws, scale, lp, d, s)))
            DLSG_PRADIX_CASE(16, (launch_prows_normal<16>(tl, w, ws, scale, lp, d, s)))
            DLSG_PRADIX_CASE(20, (launch_prows_normal<20>(tl, w, ws, scale, lp, d, s)))
#endif
            default: break;
        }
    }
    return rows_lines<MODE_NORMAL>(tl, tl, w, ws, scale, lp, d, s);
}

bool lines_supported(int E, int Y, int X, int lmax) {
    if (E <= 0 || lmax <= 0 || lmax > Y) return false;
    if (!radix_of(Y) || !radix_of(X)) return false;
    if (!pcols_disabled() && pcols_supported(E, Y, X, lmax)) return true;
    return cols_adj_lines_smem(E, Y, lmax) <= kSmemMax;
}

int run_lines(Op op, const float2* in, const float2* maps, const float* w, const long long* w_strides,
              const int* plan, int lmax, float2* out, const float2* aty, float step, const float* step_dev, int epi,
              void* workspace, size_t workspace_bytes, int B, int E, int C, int T, int Y, int X, void* stream,
              bool gather_k = false) {
    if (B <= 0 || E <= 0 || C <= 0 || T <= 0 || Y <= 0 || X <= 0) return DLSG_EINVAL;
    if (!in || !out || !maps || !w || !w_strides || !plan) return DLSG_EINVAL;
    if (epi == EPI_DC && !aty) return DLSG_EINVAL;
    if (!lines_supported(E, Y, X, lmax)) return DLSG_EUNSUPPORTED;
    const size_t per_slice = (size_t)C * T * lmax * X * sizeof(float2);
    if (!workspace || workspace_bytes < per_slice) return DLSG_EWORKSPACE;
    long long bchunk = (long long)(workspace_bytes / per_slice);
    if (bchunk > B) bchunk = B;
    const WeightStrides ws{w_strides[0], w_strides[1], w_strides[2], w_strides[3], w_strides[4]};
    float2* tl = static_cast<float2*>(workspace);
    stream_t s = static_cast<stream_t>(stream);
    const float scale = (float)(1.0 / sqrt((double)Y * (double)X));
    const size_t x_slice = (size_t)E * T * Y * X, m_slice = (size_t)E * C * Y * X, k_slice = (size_t)C * T * Y * X;
    const int* counts = plan;
    const int* list = plan + (size_t)B * T;
    const int* pos = list + (size_t)B * T * Y;
    const int* order = pos + (size_t)B * T * Y;
    const int* cls = order + (size_t)B * T * Y;
    const float* wq = reinterpret_cast<const float*>(plan + line_plan_wq_offset(B, T, Y));
    for (long long b0 = 0; b0 < B; b0 += bchunk) {
        const int bc = (int)((B - b0 < bchunk) ? (B - b0) : bchunk);
        const SenseDims d{bc, E, C, T, Y, X, 0, T};
        const LinePlan lp{counts + b0 * T, list + b0 * T * Y, pos + b0 * T * Y, lmax,
                          order + b0 * T * Y, cls + b0 * T * kPlanClsInts, wq + b0 * T * Y * 8};
        const float2* maps_b = maps ? maps + b0 * m_slice : nullptr;
        const float* w_b = w + b0 * ws.b;
        if (op == OP_ADJ) {
            if (gather_k) {      // k may be mapped host memory: copy its sampled lines, then transform them in place
                const long long units = (long long)bc * T * C;
                DLSG_TRY(launch(sense_gather_lines_kernel, dim3((unsigned)(units < kGatherCtas ? units : kGatherCtas)), dim3(256), 0, s,
                                in + b0 * k_slice, tl, lp, d));
                DLSG_TRY((rows_lines<MODE_ADJ>(tl, tl, w_b, ws, scale, lp, d, s, 1)));
            } else {
                DLSG_TRY((rows_lines<MODE_ADJ>(in + b0 * k_slice, tl, w_b, ws, scale, lp, d, s)));
            }
            DLSG_TRY(cols_adj_lines(tl, maps_b, out + b0 * x_slice, nullptr, nullptr, 0.f, nullptr, EPI_PLAIN, lp, d, s));
            continue;
        }
        const float2* x_b = in + b0 * x_slice;
        DLSG_TRY(cols_fwd_lines(x_b, maps_b, tl, lp, d, s));
        if (op == OP_FWD) {
            float2* k_b = out + b0 * k_slice;
            DLSG_TRY(launch(sense_zero_unsampled_kernel, dim3((unsigned)((long long)bc * T * C)), dim3(256), 0, s, k_b, lp, d));
            DLSG_TRY((rows_lines<MODE_FWD>(tl, k_b, w_b, ws, scale, lp, d, s)));
        } else {
            DLSG_TRY(rows_normal_lines(tl, w_b, ws, scale, lp, d, s));
            DLSG_TRY(cols_adj_lines(tl, maps_b, out + b0 * x_slice, x_b, aty ? aty + b0 * x_slice : nullptr,
                                    step, step_dev, epi, lp, d, s));
        }
    }
    return DLSG_OK;
}

}  // namespace
}  // namespace dlsg

extern "C" {

int dlsg_version(void) { return 100; }

unsigned long long dlsg_launch_count(void) { return dlsg::launch_counter().load(std::memory_order_relaxed); }

const char* dlsg_status_string(int status) {
    switch (status) {
        case DLSG_OK: return "ok";
        case DLSG_EINVAL: return "invalid argument (null pointer, non-positive extent, or ndims not in {1,2})";
        case DLSG_EWORKSPACE: return "workspace missing or smaller than dlsg_sense_workspace_bytes(..., 0)";
        case DLSG_EUNSUPPORTED: return "axis length not supported (direct-DFT limits: X <= 4096, Y <= 1536 unless 32*R)";
        case DLSG_ELAUNCH: return "CUDA kernel launch failed";
        default: return "unknown status";
    }
}

size_t dlsg_sense_workspace_bytes(int B, int E, int C, int T, int Y, int X, int ndims, size_t budget_bytes) {
    (void)E; (void)ndims;
    if (B <= 0 || C <= 0 || T <= 0 || Y <= 0 || X <= 0) return 0;
    const size_t per_bt = dlsg::per_bt_bytes(C, Y, X);
    if (budget_bytes <= per_bt) return per_bt;
    const size_t per_slice = per_bt * (size_t)T;
    if (budget_bytes >= per_slice) {
        size_t nb = budget_bytes / per_slice;
        if (nb > (size_t)B) nb = (size_t)B;
        return nb * per_slice;
    }
    return (budget_bytes / per_bt) * per_bt;
}

int dlsg_sense_forward_c64(const void* x, const void* maps, const float* weights, const long long* w_strides,
                           void* k_out, void* workspace, size_t workspace_bytes,
                           int B, int E, int C, int T, int Y, int X, int ndims, void* stream) {
    return dlsg::run(dlsg::OP_FWD, (const float2*)x, (const float2*)maps, weights, w_strides, (float2*)k_out,
                     nullptr, 0.f, nullptr, dlsg::EPI_PLAIN, workspace, workspace_bytes, B, E, C, T, Y, X, ndims, stream);
}

int dlsg_sense_adjoint_c64(const void* k, const void* maps, const float* weights, const long long* w_strides,
                           void* x_out, void* workspace, size_t workspace_bytes,
                           int B, int E, int C, int T, int Y, int X, int ndims, void* stream) {
    return dlsg::run(dlsg::OP_ADJ, (const float2*)k, (const float2*)maps, weights, w_strides, (float2*)x_out,
                     nullptr, 0.f, nullptr, dlsg::EPI_PLAIN, workspace, workspace_bytes, B, E, C, T, Y, X, ndims, stream);
}

int dlsg_sense_normal_c64(const void* x, const void* maps, const float* weights, const long long* w_strides,
                          void* x_out, void* workspace, size_t workspace_bytes,
                          int B, int E, int C, int T, int Y, int X, int ndims, void* stream) {
    return dlsg::run(dlsg::OP_NORMAL, (const float2*)x, (const float2*)maps, weights, w_strides, (float2*)x_out,
                     nullptr, 0.f, nullptr, dlsg::EPI_PLAIN, workspace, workspace_bytes, B, E, C, T, Y, X, ndims, stream);
}

int dlsg_sense_normal_reg_c64(const void* x, const void* maps, const float* weights, const long long* w_strides,
                              float lamda, const float* lamda_dev, void* x_out, void* workspace, size_t workspace_bytes,
                              int B, int E, int C, int T, int Y, int X, int ndims, void* stream) {
    return dlsg::run(dlsg::OP_NORMAL, (const float2*)x, (const float2*)maps, weights, w_strides, (float2*)x_out,
                     nullptr, lamda, lamda_dev, dlsg::EPI_REG, workspace, workspace_bytes, B, E, C, T, Y, X, ndims, stream);
}

int dlsg_sense_dc_step_c64(const void* x, const void* aty, const void* maps, const float* weights,
                           const long long* w_strides, float step, const float* step_dev, void* x_out,
                           void* workspace, size_t workspace_bytes,
                           int B, int E, int C, int T, int Y, int X, int ndims, void* stream) {
    return dlsg::run(dlsg::OP_NORMAL, (const float2*)x, (const float2*)maps, weights, w_strides, (float2*)x_out,
                     (const float2*)aty, step, step_dev, dlsg::EPI_DC, workspace, workspace_bytes, B, E, C, T, Y, X, ndims, stream);
}

// ---- sampled-line entry points ---------------------------------------------------------------
size_t dlsg_sense_line_plan_bytes(int B, int T, int Y) {
    if (B <= 0 || T <= 0 || Y <= 0) return 0;
    return sizeof(int) * dlsg::line_plan_ints(B, T, Y);
}

int dlsg_sense_line_plan_build(const float* weights, const long long* w_strides, int B, int C, int T, int Y, int X,
                               void* plan, void* stream) {
    return dlsg_sense_line_plan_build_capped(weights, w_strides, B, C, T, Y, X, 0, plan, stream);
}

int dlsg_sense_line_plan_build_capped(const float* weights, const long long* w_strides, int B, int C, int T, int Y, int X,
                                      int max_lines, void* plan, void* stream) {
    if (B <= 0 || C <= 0 || T <= 0 || Y <= 0 || X <= 0 || max_lines < 0) return DLSG_EINVAL;
    if (!weights || !w_strides || !plan) return DLSG_EINVAL;
    const dlsg::WeightStrides ws{w_strides[0], w_strides[1], w_strides[2], w_strides[3], w_strides[4]};
    int* counts = static_cast<int*>(plan);
    int* list = counts + (size_t)B * T;
    int* pos = list + (size_t)B * T * Y;
    int* order = pos + (size_t)B * T * Y;
    int* cls = order + (size_t)B * T * Y;
    float* wq = reinterpret_cast<float*>(counts + dlsg::line_plan_wq_offset(B, T, Y));
    const int rc = dlsg::launch(dlsg::sense_line_plan_kernel, dim3((unsigned)((long long)B * T)), dim3(256), (size_t)Y * sizeof(int),
                                static_cast<dlsg::stream_t>(stream), weights, ws, C, T, Y, X, counts, list, pos, order, cls, wq, max_lines);
    return rc == 0 ? DLSG_OK : DLSG_ELAUNCH;
}

int dlsg_sense_lines_supported(int E, int Y, int X, int lmax) { return dlsg::lines_supported(E, Y, X, lmax) ? 1 : 0; }

size_t dlsg_sense_lines_workspace_bytes(int B, int C, int T, int X, int lmax, size_t budget_bytes) {
    if (B <= 0 || C <= 0 || T <= 0 || X <= 0 || lmax <= 0) return 0;
    const size_t per_slice = (size_t)C * T * lmax * X * sizeof(float2);
    size_t nb = budget_bytes / per_slice;
    if (nb < 1) nb = 1;
    if (nb > (size_t)B) nb = (size_t)B;
    return nb * per_slice;
}

int dlsg_sense_forward_lines_c64(const void* x, const void* maps, const float* weights, const long long* w_strides,
                                 const void* plan, int lmax, void* k_out, void* workspace, size_t workspace_bytes,
                                 int B, int E, int C, int T, int Y, int X, void* stream) {
    return dlsg::run_lines(dlsg::OP_FWD, (const float2*)x, (const float2*)maps, weights, w_strides, (const int*)plan, lmax,
                           (float2*)k_out, nullptr, 0.f, nullptr, dlsg::EPI_PLAIN, workspace, workspace_bytes, B, E, C, T, Y, X, stream);
}

int dlsg_sense_adjoint_lines_c64(const void* k, const void* maps, const float* weights, const long long* w_strides,
                                 const void* plan, int lmax, void* x_out, void* workspace, size_t workspace_bytes,
                                 int B, int E, int C, int T, int Y, int X, void* stream) {
    return dlsg::run_lines(dlsg::OP_ADJ, (const float2*)k, (const float2*)maps, weights, w_strides, (const int*)plan, lmax,
                           (float2*)x_out, nullptr, 0.f, nullptr, dlsg::EPI_PLAIN, workspace, workspace_bytes, B, E, C, T, Y, X, stream);
}

int dlsg_sense_adjoint_lines_gather_c64(const void* k, const void* maps, const float* weights, const long long* w_strides,
                                        const void* plan, int lmax, void* x_out, void* workspace, size_t workspace_bytes,
                                        int B, int E, int C, int T, int Y, int X, void* stream) {
    return dlsg::run_lines(dlsg::OP_ADJ, (const float2*)k, (const float2*)maps, weights, w_strides, (const int*)plan, lmax,
                           (float2*)x_out, nullptr, 0.f, nullptr, dlsg::EPI_PLAIN, workspace, workspace_bytes, B, E, C, T, Y, X, stream,
                           true);
}

int dlsg_sense_normal_lines_c64(const void* x, const void* aty, const void* maps, const float* weights,
                                const long long* w_strides, const void* plan, int lmax,
                                float step, const float* step_dev, int epilogue, void* x_out,
                                void* workspace, size_t workspace_bytes,
                                int B, int E, int C, int T, int Y, int X, void* stream) {
    if (epilogue < 0 || epilogue > 2) return DLSG_EINVAL;
    return dlsg::run_lines(dlsg::OP_NORMAL, (const float2*)x, (const float2*)maps, weights, w_strides, (const int*)plan, lmax,
                           (float2*)x_out, (const float2*)aty, step, step_dev, epilogue, workspace, workspace_bytes,
                           B, E, C, T, Y, X, stream);
}

}  // extern "C"
