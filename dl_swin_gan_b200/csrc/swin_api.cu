// C-ABI entry points of the shifted-window attention path (declared in include/dlsg_b200.h).
#include "../../include/dlsg_b200.h"
#include "swin_kernels.cuh"
#include "swin_linear_tc.cuh"

namespace dlsg {
namespace {

int make_geom(SwinGeom& g, int B, int D, int H, int W, int C, const int* window, const int* shift) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || C <= 0 || !window || !shift) return DLSG_EINVAL;
    for (int i = 0; i < 3; ++i)
        if (window[i] <= 0 || shift[i] < 0 || shift[i] >= window[i]) return DLSG_EINVAL;
    g.B = B; g.D = D; g.H = H; g.W = W; g.C = C;
    g.wd = window[0]; g.wh = window[1]; g.ww = window[2];
    g.sd = shift[0]; g.sh = shift[1]; g.sw = shift[2];
    g.Dp = (D + g.wd - 1) / g.wd * g.wd;
    g.Hp = (H + g.wh - 1) / g.wh * g.wh;
    g.Wp = (W + g.ww - 1) / g.ww * g.ww;
    return DLSG_OK;
}

long long geom_tokens(const SwinGeom& g) {
    return (long long)g.B * (g.Dp / g.wd) * (g.Hp / g.wh) * (g.Wp / g.ww) * g.wd * g.wh * g.ww;
}

}  // namespace
}  // namespace dlsg

extern "C" {

int dlsg_swin_ln_partition_f32(const float* x, const float* gamma, const float* beta, float eps, float* xw,
                               int B, int D, int H, int W, int C, const int* window, const int* shift, void* stream) {
    dlsg::SwinGeom g;
    const int st = dlsg::make_geom(g, B, D, H, W, C, window, shift);
    if (st != DLSG_OK) return st;
    if (!x || !xw) return DLSG_EINVAL;
    const long long tokens = dlsg::geom_tokens(g);
    const unsigned grid = (unsigned)((tokens + 7) / 8);
    return dlsg::launch(dlsg::swin_ln_partition_kernel, dim3(grid), dim3(256), 0, (dlsg::stream_t)stream,
                        x, gamma, beta, eps, xw, g) == 0 ? DLSG_OK : DLSG_ELAUNCH;
}

int dlsg_swin_window_reverse_f32(const float* yw, float* y, int B, int D, int H, int W, int C,
                                 const int* window, const int* shift, void* stream) {
    dlsg::SwinGeom g;
    const int st = dlsg::make_geom(g, B, D, H, W, C, window, shift);
    if (st != DLSG_OK) return st;
    if (!yw || !y) return DLSG_EINVAL;
    const long long tokens = dlsg::geom_tokens(g);
    const unsigned grid = (unsigned)((tokens + 7) / 8);
    return dlsg::launch(dlsg::swin_window_reverse_kernel, dim3(grid), dim3(256), 0, (dlsg::stream_t)stream,
                        yw, y, g) == 0 ? DLSG_OK : DLSG_ELAUNCH;
}

int dlsg_swin_window_attention_f32(const float* qkv, const float* bias_table, float* out,
                                   int B_, int N, int heads, int head_dim, const int* module_window, float scale,
                                   int mask_mode, const void* mask, int nW, void* stream) {
    using namespace dlsg;
    if (!qkv || !bias_table || !out || !module_window) return DLSG_EINVAL;
    if (B_ <= 0 || N <= 0 || heads <= 0 || head_dim <= 0) return DLSG_EINVAL;
    if (mask_mode < SWIN_MASK_NONE || mask_mode > SWIN_MASK_DENSE) return DLSG_EINVAL;
    if (mask_mode != SWIN_MASK_NONE && (!mask || nW <= 0 || B_ % nW != 0)) return DLSG_EINVAL;
    const int mwd = module_window[0], mwh = module_window[1], mww = module_window[2];
    if (mwd <= 0 || mwh <= 0 || mww <= 0 || N > mwd * mwh * mww) return DLSG_EINVAL;
    const long long table_len = (long long)(2 * mwd - 1) * (2 * mwh - 1) * (2 * mww - 1);
    if (head_dim > kAttnHdPad || table_len > 65535 || N > 4096) return DLSG_EUNSUPPORTED;

    AttnParams p{};
    p.qkv = qkv; p.table = bias_table; p.out = out;
    p.region = (mask_mode == SWIN_MASK_REGION) ? static_cast<const unsigned char*>(mask) : nullptr;
    p.mask = (mask_mode == SWIN_MASK_DENSE) ? static_cast<const float*>(mask) : nullptr;
    p.B_ = B_; p.N = N; p.heads = heads; p.hd = head_dim; p.nW = (mask_mode == SWIN_MASK_NONE) ? 1 : nW;
    p.mwd = mwd; p.mwh = mwh; p.mww = mww;
    p.scale = scale;
    p.npad = (N + kAttnKB - 1) / kAttnKB * kAttnKB;
    p.vstride = p.npad + 8;
    p.table_len = (int)table_len;
    const size_t smem = attn_smem_bytes(p.npad, p.vstride, p.table_len);
    if (smem > 227 * 1024) return DLSG_EUNSUPPORTED;
    const dim3 grid(heads, B_), block(32 * kAttnWarps);
    int rc;
    if (mask_mode == SWIN_MASK_NONE)        rc = launch(swin_attn_core_kernel<SWIN_MASK_NONE>, grid, block, smem, (stream_t)stream, p);
    else if (mask_mode == SWIN_MASK_REGION) rc = launch(swin_attn_core_kernel<SWIN_MASK_REGION>, grid, block, smem, (stream_t)stream, p);
    else                                    rc = launch(swin_attn_core_kernel<SWIN_MASK_DENSE>, grid, block, smem, (stream_t)stream, p);
    return rc == 0 ? DLSG_OK : DLSG_ELAUNCH;
}

int dlsg_swin_linear_f32(const float* x, const float* weight, const float* bias, float* y,
                         int M, int N, int K, void* stream) {
#if defined(__CUDACC__)
    using namespace dlsg;
    if (!x || !weight || !y || M <= 0 || N <= 0 || K <= 0) return DLSG_EINVAL;
    if (K % 16 != 0 || N % 32 != 0 || N > 512 || (N > 256 && (N / 2) % 16 != 0)) return DLSG_EUNSUPPORTED;
    if ((reinterpret_cast<size_t>(x) | reinterpret_cast<size_t>(weight)) % 16 != 0) return DLSG_EUNSUPPORTED;
    const size_t smem = linear_smem_bytes(N, K);
    if (smem > 227 * 1024) return DLSG_EUNSUPPORTED;
    LinearParams p{x, weight, bias, y, M, N, K, 32};
    while (p.tmem_cols < N) p.tmem_cols *= 2;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = (M + kLinTileM - 1) / kLinTileM;
    const unsigned grid = (unsigned)(ntiles < sms ? ntiles : sms);
    return launch(swin_linear_tc_kernel, dim3(grid), dim3(kLinThreads), smem, (stream_t)stream, p) == 0 ? DLSG_OK : DLSG_ELAUNCH;
#else
    (void)x; (void)weight; (void)bias; (void)y; (void)M; (void)N; (void)K; (void)stream;
    return DLSG_EUNSUPPORTED;      // tcgen05 / TMEM have no host emulation
#endif
}

}  // extern "C"
