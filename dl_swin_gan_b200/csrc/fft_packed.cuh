// Packed-pair FP32 arithmetic for the SENSE butterflies (sm_100a): a complex number is one 64-bit register
// pair and every add / multiply / fma below is ONE instruction on both halves (PTX add/mul/fma.rn.f32x2 ->
// SASS FADD2 / FMUL2 / FFMA2).  Measured on the B200 (tools/microbench/fp32_issue.cu,
// profiles/r2a_fp32_issue_microbench.log): FFMA2 sustains 1.95 warp-instructions/clk/SM = the same 73 TFLOP/s as
// scalar FFMA at 3.6/clk, i.e. the same FP32 pipe time for HALF the issue slots - and the column kernels were
// issue-bound (profiles/r1q_lines_ncu_summary.csv: issue-active 56 %, 1594 instructions per thread and coil for
// ~900 of FP32 work).  The operand forms ptxas folds for free: pair, pair with swapped halves (.LO_HI), one 32-bit
// register broadcast to both halves (.F32).
//
// The header also compiles as plain host C++ (scalar fallbacks, same rounding: each half is an IEEE fma / add /
// mul), so the host-emulation tests cover the same index math.
//
// Convention as in fft_radix.cuh: DIR = -1 forward (exp(-2 pi i nk/N)), DIR = +1 unnormalised inverse.
// Reference path served: torch.fft.fftn / ifftn, dl_cs/mri/transforms.py:36-44.
#pragma once

#include "fft_radix.cuh"

namespace dlsg {

#if defined(__CUDA_ARCH__)
DLSG_HD float2 p_add(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
DLSG_HD float2 p_sub(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
DLSG_HD float2 p_mul(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mul.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
// a * b + c on both halves
DLSG_HD float2 p_fma(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; "
        "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
#else
DLSG_HD float2 p_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
DLSG_HD float2 p_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
DLSG_HD float2 p_mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
DLSG_HD float2 p_fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
DLSG_HD float2 p_bc(float v) { return make_float2(v, v); }               // broadcast (folds into the .F32 operand form)
DLSG_HD float2 p_sw(float2 a) { return make_float2(a.y, a.x); }          // swapped halves (folds into .LO_HI)
DLSG_HD float2 p_neg(float2 a) { return make_float2(-a.x, -a.y); }

// ---------------------------------------------------------------------------- complex on pairs
// a * b, both run-time: (ax bx - ay by, ax by + ay bx) = ax * b + ay * (-by, bx)
DLSG_HD float2 pc_mul(float2 a, float2 b) {
    const float2 t = p_mul(p_bc(a.y), p_sw(b));                          // (ay by, ay bx)
    return p_fma(p_bc(a.x), b, make_float2(-t.x, t.y));
}
// acc + a * b
DLSG_HD float2 pc_fma(float2 a, float2 b, float2 acc) {
    const float2 t = p_fma(p_bc(a.x), b, acc);
    const float2 nb = make_float2(-b.y, b.x);                            // i * b : one sign flip, the swap is free
    return p_fma(p_bc(a.y), nb, t);
}
// acc + conj(a) * b = acc + ax * b + ay * (by, -bx)
DLSG_HD float2 pc_fmac(float2 a, float2 b, float2 acc) {
    const float2 t = p_fma(p_bc(a.x), b, acc);
    const float2 nb = make_float2(b.y, -b.x);                            // -i * b
    return p_fma(p_bc(a.y), nb, t);
}
// acc + w * b for a coefficient stored as the two pairs  w1 = (c, s), w2 = (-s, c)  (w = c + i s)
DLSG_HD float2 pc_fma_w(float2 w1, float2 w2, float2 b, float2 acc) {
    return p_fma(p_bc(b.y), w2, p_fma(p_bc(b.x), w1, acc));
}

// a + (C + i S) * b with compile-time C, S; the special cases cost one instruction
template <int K, int N, int DIR>
struct PTw {
    static constexpr int k = ((K % N) + N) % N;
    static constexpr float c = float(ct_cos2pi(k, N));
    static constexpr float s = float(double(DIR) * ct_sin2pi(k, N));
    static constexpr bool one = (k == 0), minus_one = (2 * k == N);
    static constexpr bool plus_i = (4 * k == N && DIR > 0) || (4 * k == 3 * N && DIR < 0);      // w = +i
    static constexpr bool minus_i = (4 * k == N && DIR < 0) || (4 * k == 3 * N && DIR > 0);     // w = -i
    static constexpr bool trivial = one || minus_one || plus_i || minus_i;
    // a + w * b
    static DLSG_HD float2 madd(float2 b, float2 a) {
        if constexpr (one) return p_add(a, b);
        else if constexpr (minus_one) return p_sub(a, b);
        else if constexpr (plus_i) return p_fma(p_sw(b), make_float2(-1.f, 1.f), a);            // i b = (-by, bx)
        else if constexpr (minus_i) return p_fma(p_sw(b), make_float2(1.f, -1.f), a);           // -i b = (by, -bx)
        else return p_fma(p_bc(b.y), make_float2(-s, c), p_fma(p_bc(b.x), make_float2(c, s), a));
    }
    // a - w * b
    static DLSG_HD float2 msub(float2 b, float2 a) {
        if constexpr (one) return p_sub(a, b);
        else if constexpr (minus_one) return p_add(a, b);
        else if constexpr (plus_i) return p_fma(p_sw(b), make_float2(1.f, -1.f), a);
        else if constexpr (minus_i) return p_fma(p_sw(b), make_float2(-1.f, 1.f), a);
        else return p_fma(p_bc(b.y), make_float2(s, -c), p_fma(p_bc(b.x), make_float2(-c, -s), a));
    }
    // w * b
    static DLSG_HD float2 mul(float2 b) {
        if constexpr (one) return b;
        else if constexpr (minus_one) return p_neg(b);
        else if constexpr (plus_i) return make_float2(-b.y, b.x);
        else if constexpr (minus_i) return make_float2(b.y, -b.x);
        else return p_fma(p_bc(b.y), make_float2(-s, c), p_mul(p_bc(b.x), make_float2(c, s)));
    }
};

// radix-2 butterfly with the twiddle on the second input: a <- a + w b, b <- a - w b
template <int K, int N, int DIR>
DLSG_HD void p_bfly2(float2& a, float2& b) {
    using W = PTw<K, N, DIR>;
    if constexpr (W::trivial) {
        const float2 lo = W::madd(b, a), hi = W::msub(b, a);
        a = lo; b = hi;
    } else {
        const float2 lo = W::madd(b, a);
        const float2 hi = p_fma(a, p_bc(2.0f), p_neg(lo));               // 2a - lo
        a = lo; b = hi;
    }
}

// ---------------------------------------------------------------------------- PDft<N, DIR>: in-place, natural order
template <int N, int DIR> struct PDft;
template <int DIR> struct PDft<1, DIR> { static DLSG_HD void run(float2*) {} };
template <int DIR> struct PDft<2, DIR> {
    static DLSG_HD void run(float2* v) { const float2 a = v[0], b = v[1]; v[0] = p_add(a, b); v[1] = p_sub(a, b); }
};
template <int DIR> struct PDft<3, DIR> {
    static DLSG_HD void run(float2* v) {
        constexpr float s = float(double(DIR) * 0.86602540378443864676372317075294);
        const float2 t = p_add(v[1], v[2]), d = p_sub(v[1], v[2]);
        const float2 m = p_fma(t, p_bc(-0.5f), v[0]);
        v[0] = p_add(v[0], t);
        v[1] = p_fma(p_sw(d), make_float2(-s, s), m);                    // m + i s d
        v[2] = p_fma(p_sw(d), make_float2(s, -s), m);
    }
};
template <int DIR> struct PDft<4, DIR> {
    static DLSG_HD void run(float2* v) {
        const float2 a = p_add(v[0], v[2]), b = p_sub(v[0], v[2]);
        const float2 c = p_add(v[1], v[3]), d = p_sub(v[1], v[3]);
        constexpr float g = float(DIR);                                  // e = DIR * i * d = g * (-dy, dx)
        v[0] = p_add(a, c); v[2] = p_sub(a, c);
        v[1] = p_fma(p_sw(d), make_float2(-g, g), b);
        v[3] = p_fma(p_sw(d), make_float2(g, -g), b);
    }
};
template <int DIR> struct PDft<5, DIR> {
    static DLSG_HD void run(float2* v) {
        constexpr float c1 = float(ct_cos2pi(1, 5)), c2 = float(ct_cos2pi(2, 5));
        constexpr float s1 = float(double(DIR) * ct_sin2pi(1, 5)), s2 = float(double(DIR) * ct_sin2pi(2, 5));
        const float2 t1 = p_add(v[1], v[4]), t3 = p_sub(v[1], v[4]);
        const float2 t2 = p_add(v[2], v[3]), t4 = p_sub(v[2], v[3]);
        const float2 m1 = p_fma(t1, p_bc(c1), p_fma(t2, p_bc(c2), v[0]));
        const float2 m2 = p_fma(t1, p_bc(c2), p_fma(t2, p_bc(c1), v[0]));
        const float2 n1 = p_fma(t3, p_bc(s1), p_mul(t4, p_bc(s2)));
        const float2 n2 = p_fma(t3, p_bc(s2), p_mul(t4, p_bc(-s1)));
        v[0] = p_add(v[0], p_add(t1, t2));
        v[1] = p_fma(p_sw(n1), make_float2(-1.f, 1.f), m1);              // m1 + i n1
        v[4] = p_fma(p_sw(n1), make_float2(1.f, -1.f), m1);
        v[2] = p_fma(p_sw(n2), make_float2(-1.f, 1.f), m2);
        v[3] = p_fma(p_sw(n2), make_float2(1.f, -1.f), m2);
    }
};
// Composite N = P * M (P = smallest prime factor), decimation in time, as Dft<N, DIR> of fft_radix.cuh
template <int N, int DIR> struct PDft {
    static constexpr int P = smallest_factor(N);
    static constexpr int M = N / P;
    static_assert(P == 2 || P == 3 || P == 5, "PDft: only 2,3,5-smooth sizes are instantiable");
    template <int K> static DLSG_HD void combine(float2 (&sub)[P][M], float2* v) {
        if constexpr (K < M) {
            if constexpr (P == 2) {
                float2 a = sub[0][K], b = sub[1][K];
                p_bfly2<K, N, DIR>(a, b);
                v[K] = a; v[K + M] = b;
            } else {
                float2 g[P];
#pragma unroll
                for (int p = 0; p < P; ++p) g[p] = sub[p][K];
                twiddle_all<K, 1>(g);
                PDft<P, DIR>::run(g);
#pragma unroll
                for (int q = 0; q < P; ++q) v[K + M * q] = g[q];
            }
            combine<K + 1>(sub, v);
        }
    }
    template <int K, int Pi> static DLSG_HD void twiddle_all(float2* g) {
        if constexpr (Pi < P) {
            g[Pi] = PTw<Pi * K, N, DIR>::mul(g[Pi]);
            twiddle_all<K, Pi + 1>(g);
        }
    }
    static DLSG_HD void run(float2* v) {
        float2 sub[P][M];
#pragma unroll
        for (int p = 0; p < P; ++p)
#pragma unroll
            for (int m = 0; m < M; ++m) sub[p][m] = v[P * m + p];
#pragma unroll
        for (int p = 0; p < P; ++p) PDft<M, DIR>::run(sub[p]);
        combine<0>(sub, v);
    }
};

}  // namespace dlsg
