// C-ABI entry point of the denoisers' 3 x 3 x 3 convolution (declared in include/dlsg_b200.h).
#include "../../include/dlsg_b200.h"
#include <cstdlib>
#include "conv3d_tc.cuh"
#include "conv3d_fewout.cuh"

extern "C" {

static int conv_kc(int Cin, int Cout) {
    if (Cin <= 0 || Cout <= 0 || (Cin % dlsg::kConvKC != 0 && Cin != 2 && Cin != 4) || Cout % 32 != 0 || Cout > 256) return 0;
    return dlsg::conv_smem_bytes(Cout) <= 227 * 1024 ? dlsg::kConvKC : 0;
}

// 64 -> 64-type layers take the operand-swapped kernel (conv3d_k3_c64_kernel); its weight image stacks the heads and remainders as rows
static bool conv_c64(int Cin, int Cout) {
    const char* e = std::getenv("DLSG_CONV_C64");
    return Cout == 64 && Cin % dlsg::kConvKC == 0 && !(e && e[0] == '0');
}

int dlsg_conv3d_k3_supported(int Cin, int Cout) {
#if defined(__CUDACC__)
    const int kc = conv_kc(Cin, Cout);
    return (kc && conv_c64(Cin, Cout)) ? (kc | 0x100) : kc;
#else
    (void)Cin; (void)Cout;
    return 0;
#endif
}

#if defined(__CUDACC__)
static int conv_launch(dlsg::ConvParams& p, void* stream) {
    using namespace dlsg;
    if (!p.x || !p.wprep || !p.out || p.NB <= 0 || p.T <= 0 || p.Y <= 0 || p.X <= 0 || p.To <= 0 || p.Yo <= 0 || p.Xo <= 0) return DLSG_EINVAL;
    if (!conv_kc(p.Cin, p.Cout) || p.X > 1022 || p.Y > 1022 || p.T > 2046) return DLSG_EUNSUPPORTED;
    // the kernel's address arithmetic is 32-bit in elements
    if ((long long)p.NB * p.T * p.Y * p.X * p.Cin > 0x7fffffffLL || (long long)p.NB * p.To * p.Yo * p.Xo > 0x7fffffffLL) return DLSG_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.wprep) | reinterpret_cast<uintptr_t>(p.out)) & 15) return DLSG_EINVAL;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long ntiles = (p.NV + kConvTileRows - 1) / kConvTileRows;
    if (p.mode == CONV_UNPATCH) ntiles *= (long long)p.ksz * p.ksz * p.ksz;
    const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
    const size_t smem = conv_smem_bytes(p.Cout);
    int rc;
    if (p.relu_in)
        rc = p.Cin == 2 ? launch(conv3d_k3_tf32x3_kernel<2, true>, dim3(grid), dim3(kConvThreads), smem, (stream_t)stream, p)
           : p.Cin == 4 ? launch(conv3d_k3_tf32x3_kernel<4, true>, dim3(grid), dim3(kConvThreads), smem, (stream_t)stream, p)
                        : launch(conv3d_k3_tf32x3_kernel<0, true>, dim3(grid), dim3(kConvThreads), smem, (stream_t)stream, p);
    else
        rc = p.Cin == 2 ? launch(conv3d_k3_tf32x3_kernel<2, false>, dim3(grid), dim3(kConvThreads), smem, (stream_t)stream, p)
           : p.Cin == 4 ? launch(conv3d_k3_tf32x3_kernel<4, false>, dim3(grid), dim3(kConvThreads), smem, (stream_t)stream, p)
                        : launch(conv3d_k3_tf32x3_kernel<0, false>, dim3(grid), dim3(kConvThreads), smem, (stream_t)stream, p);
    return rc == 0 ? DLSG_OK : DLSG_ELAUNCH;
}
#endif

int dlsg_conv3d_k3_tf32x3_f32(const float* x, const void* wprep, const float* bias, const float* resid, float resid_scale, float* out,
                              int NB, int T, int Y, int X, int Cin, int Cout, int relu_in, int t_lo, int t_n, void* stream) {
#if defined(__CUDACC__)
    if (t_n == 0) { t_lo = 0; t_n = T; }
    if (t_lo < 0 || t_n < 1 || t_lo + t_n > T) return DLSG_EINVAL;
    dlsg::ConvParams p{};
    p.x = x; p.wprep = static_cast<const float*>(wprep); p.bias = bias; p.out = out;
    p.NB = NB; p.T = T; p.Y = Y; p.X = X; p.Cin = Cin; p.Cout = Cout;
    p.To = T; p.Yo = Y; p.Xo = X; p.ksz = 3; p.stride = 1; p.pad = 1; p.mode = dlsg::CONV_GATHER;
    p.tlo = t_lo; p.tn = t_n;
    p.NV = (long long)NB * t_n * Y * X;
    p.relu_in = relu_in ? 1 : 0;
    if (resid && (reinterpret_cast<uintptr_t>(resid) & 15)) return DLSG_EINVAL;
    p.resid = resid; p.resid_scale = resid_scale < 0.f ? -resid_scale : resid_scale; p.resid_relu = resid_scale < 0.f ? 1 : 0;
    if (conv_c64(Cin, Cout)) {
        using namespace dlsg;
        if (!p.x || !p.wprep || !p.out || NB <= 0 || T <= 0 || Y <= 0 || X <= 0) return DLSG_EINVAL;
        if (X > 1022 || Y > 1022 || T > 2046 || p.NV > 0x7fffffffLL) return DLSG_EUNSUPPORTED;
        if ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.wprep) | reinterpret_cast<uintptr_t>(p.out)) & 15) return DLSG_EINVAL;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long long ntiles = (p.NV + kConvTileRows - 1) / kConvTileRows;
        const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
        if ((long long)p.NV * Cin > 0x7fffffffLL) return DLSG_EUNSUPPORTED;       // 32-bit element offsets
        const int rc = p.relu_in ? launch(conv3d_k3_c64_kernel<true>, dim3(grid), dim3(kC64Threads), conv_c64_smem_bytes(), (stream_t)stream, p)
                                 : launch(conv3d_k3_c64_kernel<false>, dim3(grid), dim3(kC64Threads), conv_c64_smem_bytes(), (stream_t)stream, p);
        return rc == 0 ? DLSG_OK : DLSG_ELAUNCH;
    }
    return conv_launch(p, stream);
#else
    (void)x; (void)wprep; (void)bias; (void)resid; (void)resid_scale; (void)out; (void)NB; (void)T; (void)Y; (void)X; (void)Cin; (void)Cout; (void)relu_in;
    (void)t_lo; (void)t_n; (void)stream;
    return DLSG_EUNSUPPORTED;      // tcgen05 / TMEM have no host emulation
#endif
}

int dlsg_conv3d_patch_tf32x3_f32(const float* x, const void* wprep, const float* bias, float* out,
                                 int NB, int T, int Y, int X, int Cin, int Cout, int patch, void* stream) {
#if defined(__CUDACC__)
    if (patch < 1 || patch > 8 || Cin % dlsg::kConvKC != 0) return DLSG_EUNSUPPORTED;
    if (T % patch || Y % patch || X % patch) return DLSG_EINVAL;
    dlsg::ConvParams p{};
    p.x = x; p.wprep = static_cast<const float*>(wprep); p.bias = bias; p.out = out;
    p.NB = NB; p.T = T; p.Y = Y; p.X = X; p.Cin = Cin; p.Cout = Cout;
    p.To = T / patch; p.Yo = Y / patch; p.Xo = X / patch; p.ksz = patch; p.stride = patch; p.pad = 0; p.mode = dlsg::CONV_GATHER;
    p.tlo = 0; p.tn = p.To;
    p.NV = (long long)NB * p.To * p.Yo * p.Xo;
    return conv_launch(p, stream);
#else
    (void)x; (void)wprep; (void)bias; (void)out; (void)NB; (void)T; (void)Y; (void)X; (void)Cin; (void)Cout; (void)patch; (void)stream;
    return DLSG_EUNSUPPORTED;
#endif
}

int dlsg_conv3d_unpatch_tf32x3_f32(const float* x, const void* wprep, const float* bias, float* out,
                                   int NB, int T, int Y, int X, int Cin, int Cout, int patch, void* stream) {
#if defined(__CUDACC__)
    if (patch < 1 || patch > 8 || Cin % dlsg::kConvKC != 0) return DLSG_EUNSUPPORTED;
    dlsg::ConvParams p{};
    p.x = x; p.wprep = static_cast<const float*>(wprep); p.bias = bias; p.out = out;
    p.NB = NB; p.T = T; p.Y = Y; p.X = X; p.Cin = Cin; p.Cout = Cout;
    p.To = T * patch; p.Yo = Y * patch; p.Xo = X * patch; p.ksz = patch; p.stride = patch; p.pad = 0; p.mode = dlsg::CONV_UNPATCH;
    p.tlo = 0; p.tn = p.To;
    p.NV = (long long)NB * T * Y * X;
    return conv_launch(p, stream);
#else
    (void)x; (void)wprep; (void)bias; (void)out; (void)NB; (void)T; (void)Y; (void)X; (void)Cin; (void)Cout; (void)patch; (void)stream;
    return DLSG_EUNSUPPORTED;
#endif
}

int dlsg_conv3d_pointwise_tf32x3_f32(const float* x, const void* wprep, const float* bias, float* out,
                                     int NB, int T, int Y, int X, int Cin, int Cout, int relu_in, int t_lo, int t_n, void* stream) {
#if defined(__CUDACC__)
    if (Cin % dlsg::kConvKC != 0) return DLSG_EUNSUPPORTED;
    if (t_n == 0) { t_lo = 0; t_n = T; }
    if (t_lo < 0 || t_n < 1 || t_lo + t_n > T) return DLSG_EINVAL;
    dlsg::ConvParams p{};
    p.x = x; p.wprep = static_cast<const float*>(wprep); p.bias = bias; p.out = out;
    p.NB = NB; p.T = T; p.Y = Y; p.X = X; p.Cin = Cin; p.Cout = Cout;
    p.To = T; p.Yo = Y; p.Xo = X; p.ksz = 1; p.stride = 1; p.pad = 0; p.mode = dlsg::CONV_GATHER;
    p.tlo = t_lo; p.tn = t_n;
    p.NV = (long long)NB * t_n * Y * X;
    p.relu_in = relu_in ? 1 : 0;
    return conv_launch(p, stream);
#else
    (void)x; (void)wprep; (void)bias; (void)out; (void)NB; (void)T; (void)Y; (void)X; (void)Cin; (void)Cout; (void)relu_in; (void)t_lo; (void)t_n; (void)stream;
    return DLSG_EUNSUPPORTED;
#endif
}

int dlsg_conv3d_tapsum_f32(const float* z, const float* bias, float* out, int NB, int T, int Y, int X, int ZC, int Cout,
                           int t_lo, int t_n, void* stream) {
#if defined(__CUDACC__)
    using namespace dlsg;
    if (!z || !out || NB <= 0 || T <= 0 || Y <= 0 || X <= 0) return DLSG_EINVAL;
    if (t_n == 0) { t_lo = 0; t_n = T; }
    if (t_lo < 0 || t_n < 1 || t_lo + t_n > T) return DLSG_EINVAL;
    if ((Cout != 2 && Cout != 4) || ZC < 27 * Cout || ZC % 4 != 0) return DLSG_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(out)) & 15) return DLSG_EINVAL;
    const long long NV = (long long)NB * t_n * Y * X;
    const unsigned grid = (unsigned)((NV + 255) / 256);
    const int rc = Cout == 2 ? launch(conv3d_tapsum_kernel<2>, dim3(grid), dim3(256), 0, (stream_t)stream, z, bias, out, T, Y, X, ZC, NV, t_lo, t_n)
                             : launch(conv3d_tapsum_kernel<4>, dim3(grid), dim3(256), 0, (stream_t)stream, z, bias, out, T, Y, X, ZC, NV, t_lo, t_n);
    return rc == 0 ? DLSG_OK : DLSG_ELAUNCH;
#else
    (void)z; (void)bias; (void)out; (void)NB; (void)T; (void)Y; (void)X; (void)ZC; (void)Cout; (void)t_lo; (void)t_n; (void)stream;
    return DLSG_EUNSUPPORTED;
#endif
}

int dlsg_conv3d_k3_fewout_f32(const float* x, const float* w_tcc, const float* bias, float* out,
                              int NB, int T, int Y, int X, int Cin, int Cout, int relu_in, void* stream) {
#if defined(__CUDACC__)
    using namespace dlsg;
    if (!x || !w_tcc || !out || NB <= 0 || T <= 0 || Y <= 0 || X <= 0) return DLSG_EINVAL;
    if ((Cout != 2 && Cout != 4) || Cin <= 0 || Cin % 32 != 0 || X > 1023 || Y > 1023 || T > 2047) return DLSG_EUNSUPPORTED;
    const size_t smem = (size_t)27 * Cin * Cout * 4 + 28 * 4 + 27 * dlsg::kFewThreads;
    if (smem > 110 * 1024 || (long long)NB * T * Y * X * Cin > 0x7fffffffLL) return DLSG_EUNSUPPORTED;   // two CTAs per SM; 32-bit element offsets
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_tcc) | reinterpret_cast<uintptr_t>(out)) & 15) return DLSG_EINVAL;
    FewOutParams p{};
    p.x = x; p.w = w_tcc; p.bias = bias; p.out = out;
    p.NB = NB; p.T = T; p.Y = Y; p.X = X; p.Cin = Cin;
    p.NV = (long long)NB * T * Y * X;
    p.relu_in = relu_in ? 1 : 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long ntiles = (p.NV + kFewThreads - 1) / kFewThreads;
    const unsigned grid = (unsigned)(ntiles < 2LL * sms ? ntiles : 2LL * sms);
    const int rc = Cout == 2 ? launch(conv3d_k3_fewout_kernel<2>, dim3(grid), dim3(kFewThreads), smem, (stream_t)stream, p)
                             : launch(conv3d_k3_fewout_kernel<4>, dim3(grid), dim3(kFewThreads), smem, (stream_t)stream, p);
    return rc == 0 ? DLSG_OK : DLSG_ELAUNCH;
#else
    (void)x; (void)w_tcc; (void)bias; (void)out; (void)NB; (void)T; (void)Y; (void)X; (void)Cin; (void)Cout; (void)relu_in; (void)stream;
    return DLSG_EUNSUPPORTED;
#endif
}

}  // extern "C"
