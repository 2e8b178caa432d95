// In-register DFT building blocks for the SENSE operator kernels (sm_100a).
//
// Everything here works on small fixed-size arrays of float2 that live in
// registers: all loops are compile-time unrolled and every twiddle factor is a
// compile-time constant (constexpr trig evaluated in double, rounded once to
// fp32), so the generated SASS is a straight line of FADD/FMUL/FFMA.
//
// The header also compiles as plain host C++ (no CUDA) so that the butterflies
// are unit-tested on the CPU (tests/test_fft_radix_host.py) before any GPU time
// is spent.
//
// Convention: DIR = -1 is the forward transform  X[k] = sum_n x[n] exp(-2*pi*i*n*k/N),
//             DIR = +1 is the unnormalised inverse              exp(+2*pi*i*n*k/N).
// The reference path this serves is dl_cs/mri/transforms.py:36-44
// (torch.fft.fftn / ifftn with norm='ortho'); the 1/sqrt(N) factors are applied
// once by the calling kernels, not here.
#pragma once

#include "cuda_compat.h"

#if defined(__CUDACC__)
#define DLSG_HD __host__ __device__ __forceinline__
#define DLSG_HD_CE __host__ __device__ constexpr
#else
#define DLSG_HD inline
#define DLSG_HD_CE constexpr
#endif

namespace dlsg {

// ----------------------------------------------------------------------------
// constexpr trigonometry: cos/sin of 2*pi*k/n, exact symmetries, double Taylor.
// ----------------------------------------------------------------------------
constexpr double kPi = 3.14159265358979323846264338327950288;

DLSG_HD_CE double ct_sin_small(double x) {   // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i < 12; ++i) {
        term *= -x2 / double((2 * i) * (2 * i + 1));
        sum += term;
    }
    return sum;
}
DLSG_HD_CE double ct_cos_small(double x) {   // |x| <= pi/4
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i < 12; ++i) {
        term *= -x2 / double((2 * i - 1) * (2 * i));
        sum += term;
    }
    return sum;
}
// cos(2*pi*k/n) and sin(2*pi*k/n) for integer k (any sign), n > 0.
// Reduction is done on integers so that multiples of an eighth turn are exact.
DLSG_HD_CE double ct_cos2pi(long long k, long long n) {
    k %= n; if (k < 0) k += n;                // 0 <= k < n
    // octant index o in [0,8): angle = 2*pi*k/n ; 8k/n
    long long k8 = 8 * k;
    long long o = k8 / n;                     // floor(8k/n)
    // r = k/n - o/8 in [0, 1/8)  -> as fraction (8k - o*n) / (8n)
    long long num = k8 - o * n;               // in [0, n)
    double a = 2.0 * kPi * double(num) / double(8 * n);   // in [0, pi/4)
    double b = kPi / 4.0 - a;                               // for odd octants
    switch (o) {
        case 0: return  ct_cos_small(a);
        case 1: return  ct_sin_small(b);
        case 2: return -ct_sin_small(a);
        case 3: return -ct_cos_small(b);
        case 4: return -ct_cos_small(a);
        case 5: return -ct_sin_small(b);
        case 6: return  ct_sin_small(a);
        default: return ct_cos_small(b);
    }
}
DLSG_HD_CE double ct_sin2pi(long long k, long long n) {
    // sin(t) = cos(t - pi/2) = cos(2*pi*(k/n - 1/4)) = cos(2*pi*(4k - n)/(4n))
    return ct_cos2pi(4 * k - n, 4 * n);
}

// ----------------------------------------------------------------------------
// complex helpers
// ----------------------------------------------------------------------------
DLSG_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
DLSG_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
DLSG_HD float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}
// a * conj(b)
DLSG_HD float2 cmulc(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -(a.x * b.y)));
}
// acc + a*b
DLSG_HD float2 cfma(float2 a, float2 b, float2 acc) {
    return make_float2(fmaf(a.x, b.x, fmaf(-a.y, b.y, acc.x)),
                       fmaf(a.x, b.y, fmaf(a.y, b.x, acc.y)));
}
// acc + a*conj(b)
DLSG_HD float2 cfmac(float2 a, float2 b, float2 acc) {
    return make_float2(fmaf(a.x, b.x, fmaf(a.y, b.y, acc.x)),
                       fmaf(a.y, b.x, fmaf(-a.x, b.y, acc.y)));
}
DLSG_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

// Multiply by the compile-time twiddle W = exp(DIR * 2*pi*i * K / N).
template <int K, int N, int DIR>
DLSG_HD float2 mul_tw(float2 a) {
    constexpr int k = ((K % N) + N) % N;
    if constexpr (k == 0) {
        return a;
    } else if constexpr (2 * k == N) {
        return make_float2(-a.x, -a.y);
    } else if constexpr (4 * k == N) {          // W = DIR * i
        if constexpr (DIR < 0) return make_float2(a.y, -a.x);
        else                   return make_float2(-a.y, a.x);
    } else if constexpr (4 * k == 3 * N) {      // W = -DIR * i
        if constexpr (DIR < 0) return make_float2(-a.y, a.x);
        else                   return make_float2(a.y, -a.x);
    } else {
        constexpr float c = float(ct_cos2pi(k, N));
        constexpr float s = float(double(DIR) * ct_sin2pi(k, N));
        return make_float2(fmaf(a.x, c, -(a.y * s)), fmaf(a.x, s, a.y * c));
    }
}

// Radix-2 butterfly with a compile-time twiddle on the second input:
//   lo = a + W*b, hi = a - W*b,  W = exp(DIR*2*pi*i*K/N).
// General case costs 6 FMA (hi = 2a - lo) instead of 4 mul/fma + 4 add.
template <int K, int N, int DIR>
DLSG_HD void bfly2_tw(float2& a, float2& b) {
    constexpr int k = ((K % N) + N) % N;
    if constexpr (k == 0 || 2 * k == N || 4 * k == N || 4 * k == 3 * N) {
        float2 t = mul_tw<K, N, DIR>(b);
        float2 lo = cadd(a, t), hi = csub(a, t);
        a = lo; b = hi;
    } else {
        constexpr float c = float(ct_cos2pi(k, N));
        constexpr float s = float(double(DIR) * ct_sin2pi(k, N));
        float2 lo;
        lo.x = fmaf(b.x, c, fmaf(-b.y, s, a.x));
        lo.y = fmaf(b.x, s, fmaf(b.y, c, a.y));
        float2 hi = make_float2(fmaf(2.0f, a.x, -lo.x), fmaf(2.0f, a.y, -lo.y));
        a = lo; b = hi;
    }
}

// ----------------------------------------------------------------------------
// Dft<N, DIR>::run(v): in-place DFT of v[0..N), natural order in and out.
// ----------------------------------------------------------------------------
template <int N, int DIR> struct Dft;

template <int DIR> struct Dft<1, DIR> {
    static DLSG_HD void run(float2*) {}
};

template <int DIR> struct Dft<2, DIR> {
    static DLSG_HD void run(float2* v) {
        float2 a = v[0], b = v[1];
        v[0] = cadd(a, b); v[1] = csub(a, b);
    }
};

template <int DIR> struct Dft<3, DIR> {
    static DLSG_HD void run(float2* v) {
        constexpr float s = float(double(DIR) * 0.86602540378443864676372317075294);  // DIR*sin(2pi/3)
        float2 t = cadd(v[1], v[2]);
        float2 d = csub(v[1], v[2]);
        float2 m = make_float2(fmaf(-0.5f, t.x, v[0].x), fmaf(-0.5f, t.y, v[0].y));
        v[0] = cadd(v[0], t);
        // W3^1 = -1/2 + i*s (with s carrying DIR): out1 = m + i*s*d, out2 = m - i*s*d
        v[1] = make_float2(fmaf(-s, d.y, m.x), fmaf(s, d.x, m.y));
        v[2] = make_float2(fmaf(s, d.y, m.x), fmaf(-s, d.x, m.y));
    }
};

template <int DIR> struct Dft<4, DIR> {
    static DLSG_HD void run(float2* v) {
        float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
        float2 c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
        float2 e = mul_tw<1, 4, DIR>(d);     // DIR*i * d
        v[0] = cadd(a, c); v[2] = csub(a, c);
        v[1] = cadd(b, e); v[3] = csub(b, e);
    }
};

template <int DIR> struct Dft<5, DIR> {
    static DLSG_HD void run(float2* v) {
        constexpr float c1 = float(ct_cos2pi(1, 5)), c2 = float(ct_cos2pi(2, 5));
        constexpr float s1 = float(double(DIR) * ct_sin2pi(1, 5));
        constexpr float s2 = float(double(DIR) * ct_sin2pi(2, 5));
        float2 t1 = cadd(v[1], v[4]), t3 = csub(v[1], v[4]);
        float2 t2 = cadd(v[2], v[3]), t4 = csub(v[2], v[3]);
        float2 m1 = make_float2(fmaf(c1, t1.x, fmaf(c2, t2.x, v[0].x)), fmaf(c1, t1.y, fmaf(c2, t2.y, v[0].y)));
        float2 m2 = make_float2(fmaf(c2, t1.x, fmaf(c1, t2.x, v[0].x)), fmaf(c2, t1.y, fmaf(c1, t2.y, v[0].y)));
        // n1 = s1*t3 + s2*t4 ; n2 = s2*t3 - s1*t4   (s carry DIR)
        float2 n1 = make_float2(fmaf(s1, t3.x, s2 * t4.x), fmaf(s1, t3.y, s2 * t4.y));
        float2 n2 = make_float2(fmaf(s2, t3.x, -(s1 * t4.x)), fmaf(s2, t3.y, -(s1 * t4.y)));
        v[0] = cadd(v[0], cadd(t1, t2));
        // out1 = m1 + i*n1, out4 = m1 - i*n1, out2 = m2 + i*n2, out3 = m2 - i*n2
        v[1] = make_float2(m1.x - n1.y, m1.y + n1.x);
        v[4] = make_float2(m1.x + n1.y, m1.y - n1.x);
        v[2] = make_float2(m2.x - n2.y, m2.y + n2.x);
        v[3] = make_float2(m2.x + n2.y, m2.y - n2.x);
    }
};

DLSG_HD_CE int smallest_factor(int n) {
    if (n % 2 == 0) return 2;
    if (n % 3 == 0) return 3;
    if (n % 5 == 0) return 5;
    return n;
}

// Composite N = P*M (P = smallest prime factor), decimation in time:
//   sub_p[m] = v[P*m + p];  S_p = Dft<M>(sub_p);
//   out[k + M*q] = sum_p W_P^{p q} * (W_N^{p k} S_p[k])
template <int N, int DIR> struct Dft {
    static constexpr int P = smallest_factor(N);
    static constexpr int M = N / P;
    static_assert(P == 2 || P == 3 || P == 5, "Dft: only 2,3,5-smooth sizes are instantiable");

    template <int K> static DLSG_HD void combine(float2 (&sub)[P][M], float2* v) {
        if constexpr (K < M) {
            if constexpr (P == 2) {
                float2 a = sub[0][K], b = sub[1][K];
                bfly2_tw<K, N, DIR>(a, b);
                v[K] = a; v[K + M] = b;
            } else {
                float2 g[P];
#pragma unroll
                for (int p = 0; p < P; ++p) g[p] = sub[p][K];
                twiddle_all<K, 1>(g);
                Dft<P, DIR>::run(g);
#pragma unroll
                for (int q = 0; q < P; ++q) v[K + M * q] = g[q];
            }
            combine<K + 1>(sub, v);
        }
    }
    template <int K, int Pi> static DLSG_HD void twiddle_all(float2* g) {
        if constexpr (Pi < P) {
            g[Pi] = mul_tw<Pi * K, N, DIR>(g[Pi]);
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
        for (int p = 0; p < P; ++p) Dft<M, DIR>::run(sub[p]);
        combine<0>(sub, v);
    }
};

}  // namespace dlsg
