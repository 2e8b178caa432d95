// C-ABI entry point of the denoisers' 3 x 3 x 3 convolution (declared in include/dlsg_b200.h).
#include "../../include/dlsg_b200.h"
#include <cstdlib>
#include "conv3d_tc.cuh"

extern "C" {

static int conv_kc(int Cin, int Cout) {
    if (Cin <= 0 || Cout <= 0 || (Cin % dlsg::kConvKC != 0 && Cin != 2 && Cin != 4) || Cout % 32 != 0 || Cout > 256) return 0;
    return dlsg::conv_smem_bytes(Cout) <= 227 * 1024 ? dlsg::kConvKC : 0;
}

int dlsg_conv3d_k3_supported(int Cin, int Cout) {
#if defined(__CUDACC__)
    return conv_kc(Cin, Cout);
#else
    (void)Cin; (void)Cout;
    return 0;
#endif
}

int dlsg_conv3d_k3_tf32x3_f32(const float* x, const float* wprep, const float* bias, float* out,
                              int NB, int T, int Y, int X, int Cin, int Cout, int relu_in, void* stream) {
#if defined(__CUDACC__)
    using namespace dlsg;
    if (!x || !wprep || !out || NB <= 0 || T <= 0 || Y <= 0 || X <= 0) return DLSG_EINVAL;
    if (!dlsg_conv3d_k3_supported(Cin, Cout) || X > 1023 || Y > 1023 || T > 2047) return DLSG_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wprep) | reinterpret_cast<uintptr_t>(out)) & 15) return DLSG_EINVAL;
    ConvParams p{};
    p.x = x; p.wprep = wprep; p.bias = bias; p.out = out;
    p.NB = NB; p.T = T; p.Y = Y; p.X = X; p.Cin = Cin; p.Cout = Cout;
    p.NV = (long long)NB * T * Y * X;
    p.relu_in = relu_in ? 1 : 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long ntiles = (p.NV + kConvTileRows - 1) / kConvTileRows;
    const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
    const int rc = Cin == 2 ? launch(conv3d_k3_tf32x3_kernel<2>, dim3(grid), dim3(kConvThreads), conv_smem_bytes(Cout), (stream_t)stream, p)
                 : Cin == 4 ? launch(conv3d_k3_tf32x3_kernel<4>, dim3(grid), dim3(kConvThreads), conv_smem_bytes(Cout), (stream_t)stream, p)
                            : launch(conv3d_k3_tf32x3_kernel<0>, dim3(grid), dim3(kConvThreads), conv_smem_bytes(Cout), (stream_t)stream, p);
    return rc == 0 ? DLSG_OK : DLSG_ELAUNCH;
#else
    (void)x; (void)wprep; (void)bias; (void)out; (void)NB; (void)T; (void)Y; (void)X; (void)Cin; (void)Cout; (void)relu_in; (void)stream;
    return DLSG_EUNSUPPORTED;      // tcgen05 / TMEM have no host emulation
#endif
}

}  // extern "C"
