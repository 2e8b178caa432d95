// Host-side harness: exposes the in-register DFT butterflies of
// dl_swin_gan_b200/csrc/fft_radix.cuh through a C function so that pytest can
// check them against numpy without a GPU.
#include "../../dl_swin_gan_b200/csrc/fft_radix.cuh"

template <int N, int DIR> static void run(float* d) {
    float2 v[N];
    for (int i = 0; i < N; ++i) v[i] = make_float2(d[2 * i], d[2 * i + 1]);
    dlsg::Dft<N, DIR>::run(v);
    for (int i = 0; i < N; ++i) { d[2 * i] = v[i].x; d[2 * i + 1] = v[i].y; }
}
template <int N> static void run_dir(int dir, float* d) { if (dir < 0) run<N, -1>(d); else run<N, 1>(d); }

extern "C" int dft_run(int n, int dir, float* d) {
    switch (n) {
        case 1: run_dir<1>(dir, d); return 0;
        case 2: run_dir<2>(dir, d); return 0;
        case 3: run_dir<3>(dir, d); return 0;
        case 4: run_dir<4>(dir, d); return 0;
        case 5: run_dir<5>(dir, d); return 0;
        case 6: run_dir<6>(dir, d); return 0;
        case 8: run_dir<8>(dir, d); return 0;
        case 10: run_dir<10>(dir, d); return 0;
        case 12: run_dir<12>(dir, d); return 0;
        case 15: run_dir<15>(dir, d); return 0;
        case 16: run_dir<16>(dir, d); return 0;
        case 32: run_dir<32>(dir, d); return 0;
        default: return -1;
    }
}
extern "C" double tw_cos(long long k, long long n) { return dlsg::ct_cos2pi(k, n); }
extern "C" double tw_sin(long long k, long long n) { return dlsg::ct_sin2pi(k, n); }
