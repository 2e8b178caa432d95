// C-ABI entry points of the k-t pre-processing kernels (declared in include/dlsg_b200.h).
#include "../../include/dlsg_b200.h"
#include "kt_utils.cuh"

extern "C" {

int dlsg_fftmod_c64(void* data, long long n_images, int Y, int X, void* stream) {
    if (!data || n_images <= 0 || Y <= 0 || X <= 0) return DLSG_EINVAL;
    const long long n = n_images * Y * X;
    const unsigned grid = (unsigned)((n + 255) / 256);
    return dlsg::launch(dlsg::fftmod_kernel, dim3(grid), dim3(256), 0, (dlsg::stream_t)stream,
                        (float2*)data, n, Y, X) == 0 ? DLSG_OK : DLSG_ELAUNCH;
}

int dlsg_get_mask_c64(const void* data, float* mask, long long n, float eps, void* stream) {
    if (!data || !mask || n <= 0) return DLSG_EINVAL;
    const unsigned grid = (unsigned)((n + 255) / 256);
    return dlsg::launch(dlsg::get_mask_kernel, dim3(grid), dim3(256), 0, (dlsg::stream_t)stream,
                        (const float2*)data, mask, n, eps) == 0 ? DLSG_OK : DLSG_ELAUNCH;
}

int dlsg_sliding_window_c64(const void* data, void* out, long long outer, int T, long long inner, int window,
                            int keep_all_frames, float mask_eps, float eps, void* stream) {
    if (!data || !out || outer <= 0 || T <= 0 || inner <= 0 || window <= 0 || window > T) return DLSG_EINVAL;
    const long long n = outer * inner;
    const unsigned grid = (unsigned)((n + 255) / 256);
    const int t_out = keep_all_frames ? T : 1;
    return dlsg::launch(dlsg::sliding_window_kernel, dim3(grid), dim3(256), 0, (dlsg::stream_t)stream,
                        (const float2*)data, (float2*)out, outer, T, inner, window, t_out, mask_eps, eps) == 0
               ? DLSG_OK : DLSG_ELAUNCH;
}

}  // extern "C"
