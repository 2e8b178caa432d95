// C-ABI entry points of the shifted-window attention path (declared in include/dlsg_b200.h).
#include "../../include/dlsg_b200.h"
#include <cstdlib>
#include "swin_kernels.cuh"
#include "swin_linear_tc.cuh"
#include "swin_mlp_tc.cuh"

namespace dlsg {

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

}  // namespace dlsg

static int attention_common(const float* qkv, const void* q16, const void* k16, const void* vt16,
                            const float* bias_table, float* out,
                            int B_, int N, int heads, int head_dim, const int* module_window, float scale,
                            int mask_mode, const void* mask, int nW, void* stream) {
    using namespace dlsg;
    const bool in16 = (qkv == nullptr);
    if (in16 ? (!q16 || !k16 || !vt16) : false) return DLSG_EINVAL;
    if (!bias_table || !out || !module_window) return DLSG_EINVAL;
    if (B_ <= 0 || N <= 0 || heads <= 0 || head_dim <= 0) return DLSG_EINVAL;
    if (mask_mode < SWIN_MASK_NONE || mask_mode > SWIN_MASK_DENSE) return DLSG_EINVAL;
    if (mask_mode != SWIN_MASK_NONE && (!mask || nW <= 0 || B_ % nW != 0)) return DLSG_EINVAL;
    const int mwd = module_window[0], mwh = module_window[1], mww = module_window[2];
    if (mwd <= 0 || mwh <= 0 || mww <= 0 || N > mwd * mwh * mww) return DLSG_EINVAL;
    const long long table_len = (long long)(2 * mwd - 1) * (2 * mwh - 1) * (2 * mww - 1);
    if (head_dim > kAttnHdPad || table_len > 65535 || N > 4096) return DLSG_EUNSUPPORTED;
    if (in16 && N % 8 != 0) return DLSG_EUNSUPPORTED;

    AttnParams p{};
    p.qkv = qkv; p.q16 = static_cast<const half_t*>(q16); p.k16 = static_cast<const half_t*>(k16);
    p.vt16 = static_cast<const half_t*>(vt16);
    p.table = bias_table; p.out = out;
    p.region = (mask_mode == SWIN_MASK_REGION) ? static_cast<const unsigned char*>(mask) : nullptr;
    p.mask = (mask_mode == SWIN_MASK_DENSE) ? static_cast<const float*>(mask) : nullptr;
    p.B_ = B_; p.N = N; p.heads = heads; p.hd = head_dim; p.nW = (mask_mode == SWIN_MASK_NONE) ? 1 : nW;
    p.mwd = mwd; p.mwh = mwh; p.mww = mww;
    p.scale = scale;
    p.npad = (N + kAttnKB - 1) / kAttnKB * kAttnKB;
    p.vstride = p.npad + 8;
    p.table_len = (int)table_len;
    const bool fast = (mww % 2 == 0) && (head_dim < kAttnHdPad);
    const size_t smem = attn_smem_bytes(p.npad, p.vstride, p.table_len, fast);
    if (smem > 227 * 1024) return DLSG_EUNSUPPORTED;
    const dim3 grid(heads, B_), block(32 * kAttnWarps);
    stream_t st = (stream_t)stream;
    int rc;
#define DLSG_ATTN_LAUNCH(MM, I16, FS) rc = launch(swin_attn_core_kernel<MM, I16, FS>, grid, block, smem, st, p)
#define DLSG_ATTN_PICK(MM)                                                   \
    do {                                                                     \
        if (in16) { if (fast) DLSG_ATTN_LAUNCH(MM, true, true); else DLSG_ATTN_LAUNCH(MM, true, false); }   \
        else      { if (fast) DLSG_ATTN_LAUNCH(MM, false, true); else DLSG_ATTN_LAUNCH(MM, false, false); } \
    } while (0)
    if (mask_mode == SWIN_MASK_NONE)        DLSG_ATTN_PICK(SWIN_MASK_NONE);
    else if (mask_mode == SWIN_MASK_REGION) DLSG_ATTN_PICK(SWIN_MASK_REGION);
    else                                    DLSG_ATTN_PICK(SWIN_MASK_DENSE);
#undef DLSG_ATTN_PICK
#undef DLSG_ATTN_LAUNCH
    return rc == 0 ? DLSG_OK : DLSG_ELAUNCH;
}

#if defined(__CUDACC__)
static int linear_check(int M, int N, int K, const void* a, const void* w) {
    if (M <= 0 || N <= 0 || K <= 0 || !a || !w) return DLSG_EINVAL;
    if (K % 16 != 0 || N % 32 != 0 || N > 512 || (N > 256 && (N / 2) % 16 != 0)) return DLSG_EUNSUPPORTED;
    if ((reinterpret_cast<size_t>(a) | reinterpret_cast<size_t>(w)) % 16 != 0) return DLSG_EUNSUPPORTED;
    if (dlsg::linear_smem_bytes(N, K) > 227 * 1024) return DLSG_EUNSUPPORTED;
    return DLSG_OK;
}
template <int AMODE, int EMODE>
static int linear_launch(dlsg::LinearParams& p, void* stream) {
    using namespace dlsg;
    p.tmem_cols = 32;
    while (p.tmem_cols < p.N) p.tmem_cols *= 2;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = (p.M + kLinTileM - 1) / kLinTileM;
    // two co-resident CTAs per SM (each with its own <= 256 TMEM columns) when shared memory allows: one CTA's
    // A-load / epilogue then overlaps the other's MMA
    const int per_sm = (linear_smem_bytes(p.N, p.K) + 1024 <= (227 * 1024) / 2 && p.tmem_cols <= 256) ? 2 : 1;
    const int slots = sms * per_sm;
    const unsigned grid = (unsigned)(ntiles < slots ? ntiles : slots);
    return launch(swin_linear_tc_kernel<AMODE, EMODE>, dim3(grid), dim3(kLinThreads), linear_smem_bytes(p.N, p.K),
                  (stream_t)stream, p) == 0 ? DLSG_OK : DLSG_ELAUNCH;
}
#endif

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
    if (!qkv) return DLSG_EINVAL;
    return attention_common(qkv, nullptr, nullptr, nullptr, bias_table, out, B_, N, heads, head_dim, module_window, scale,
                            mask_mode, mask, nW, stream);
}

int dlsg_swin_window_attention_f16(const void* q16, const void* k16, const void* vt16, const float* bias_table, float* out,
                                   int B_, int N, int heads, int head_dim, const int* module_window,
                                   int mask_mode, const void* mask, int nW, void* stream) {
    return attention_common(nullptr, q16, k16, vt16, bias_table, out, B_, N, heads, head_dim, module_window, 1.0f,
                            mask_mode, mask, nW, stream);
}

int dlsg_swin_linear_f32(const float* x, const float* weight, const float* bias, float* y,
                         int M, int N, int K, void* stream) {
#if defined(__CUDACC__)
    const int st = linear_check(M, N, K, x, weight);
    if (st != DLSG_OK) return st;
    if (!y) return DLSG_EINVAL;
    dlsg::LinearParams p{};
    p.x = x; p.w = weight; p.bias = bias; p.y = y; p.M = M; p.N = N; p.K = K;
    return linear_launch<dlsg::A_PLAIN, dlsg::E_PLAIN>(p, stream);
#else
    (void)x; (void)weight; (void)bias; (void)y; (void)M; (void)N; (void)K; (void)stream;
    return DLSG_EUNSUPPORTED;      // tcgen05 / TMEM have no host emulation
#endif
}

int dlsg_swin_qkv_fused_f32(const float* x, const float* gamma, const float* beta, float eps,
                            const float* weight, const float* bias, void* q16, void* k16, void* vt16,
                            int B, int D, int H, int W, int C, const int* window, const int* shift,
                            int heads, float scale, void* stream) {
#if defined(__CUDACC__)
    dlsg::LinearParams p{};
    int st = dlsg::make_geom(p.g, B, D, H, W, C, window, shift);
    if (st != DLSG_OK) return st;
    if (heads <= 0 || C % heads != 0 || !q16 || !k16 || !vt16) return DLSG_EINVAL;
    const int hd = C / heads, ntok = window[0] * window[1] * window[2];
    if (hd > 32 || hd % 4 != 0 || ntok % 8 != 0) return DLSG_EUNSUPPORTED;
    const long long M = dlsg::geom_tokens(p.g);
    if (M > 0x7fffffffLL) return DLSG_EUNSUPPORTED;
    st = linear_check((int)M, 3 * C, C, x, weight);
    if (st != DLSG_OK) return st;
    p.x = x; p.w = weight; p.bias = bias; p.y = nullptr; p.M = (int)M; p.N = 3 * C; p.K = C;
    p.gamma = gamma; p.beta = beta; p.eps = eps;
    p.q16 = static_cast<half_t*>(q16); p.k16 = static_cast<half_t*>(k16); p.vt16 = static_cast<half_t*>(vt16);
    p.heads = heads; p.hd = hd; p.ntok = ntok; p.qscale = scale * dlsg::kLog2e;
    return linear_launch<dlsg::A_LN_GATHER, dlsg::E_QKV16>(p, stream);
#else
    (void)x; (void)gamma; (void)beta; (void)eps; (void)weight; (void)bias; (void)q16; (void)k16; (void)vt16;
    (void)B; (void)D; (void)H; (void)W; (void)C; (void)window; (void)shift; (void)heads; (void)scale; (void)stream;
    return DLSG_EUNSUPPORTED;
#endif
}

int dlsg_swin_qkv_stream_f32(const float* x, const float* gamma, const float* beta, float eps,
                             const void* wqkv_f16, const float* bias, void* q16, void* k16, void* v16,
                             int B, int D, int H, int W, int C, const int* window, const int* shift,
                             int heads, float scale, void* stream) {
#if defined(__CUDACC__)
    dlsg::MlpParams p{};
    int st = dlsg::make_geom(p.g, B, D, H, W, C, window, shift);
    if (st != DLSG_OK) return st;
    if (heads <= 0 || C % heads != 0 || !x || !gamma || !beta || !wqkv_f16 || !q16 || !k16 || !v16) return DLSG_EINVAL;
    const int hd = C / heads, ntok = window[0] * window[1] * window[2];
    if (hd > 32 || hd % 4 != 0 || ntok % 8 != 0 || C % 32 != 0 || C > 256) return DLSG_EUNSUPPORTED;
    const long long M = dlsg::geom_tokens(p.g);
    if (M <= 0) return DLSG_EINVAL;
    if (M > 0x7fffffffLL) return DLSG_EUNSUPPORTED;
    const size_t smem = dlsg::mlp_smem_bytes(C, 3 * C);
    if (smem > 113 * 1024) return DLSG_EUNSUPPORTED;
    p.x = x; p.gamma = gamma; p.beta = beta; p.eps = eps;
    p.w1h = static_cast<const half_t*>(wqkv_f16); p.b1 = bias;
    p.M = (int)M; p.C = C; p.H = 3 * C;
    p.q16 = static_cast<half_t*>(q16); p.k16 = static_cast<half_t*>(k16); p.vt16 = static_cast<half_t*>(v16);
    p.heads = heads; p.hd = hd; p.ntok = ntok; p.qscale = scale * dlsg::kLog2e;
    p.tmem_cols = 32;
    while (p.tmem_cols < C) p.tmem_cols *= 2;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = (p.M + dlsg::kLinTileM - 1) / dlsg::kLinTileM;
    const unsigned grid = (unsigned)(ntiles < 2 * sms ? ntiles : 2 * sms);
    return dlsg::launch(dlsg::swin_ln_gemm_stream_kernel<true>, dim3(grid), dim3(dlsg::kLinThreads), smem, (dlsg::stream_t)stream, p) == 0
               ? DLSG_OK : DLSG_ELAUNCH;
#else
    (void)x; (void)gamma; (void)beta; (void)eps; (void)wqkv_f16; (void)bias; (void)q16; (void)k16; (void)v16;
    (void)B; (void)D; (void)H; (void)W; (void)C; (void)window; (void)shift; (void)heads; (void)scale; (void)stream;
    return DLSG_EUNSUPPORTED;
#endif
}

int dlsg_swin_proj_scatter_residual_f32(const float* attn_out, const float* weight, const float* bias, const float* residual,
                                        float* y, int B, int D, int H, int W, int C, const int* window, const int* shift, void* stream) {
#if defined(__CUDACC__)
    dlsg::LinearParams p{};
    int st = dlsg::make_geom(p.g, B, D, H, W, C, window, shift);
    if (st != DLSG_OK) return st;
    const long long M = dlsg::geom_tokens(p.g);
    if (M > 0x7fffffffLL) return DLSG_EUNSUPPORTED;
    st = linear_check((int)M, C, C, attn_out, weight);
    if (st != DLSG_OK) return st;
    if (!y) return DLSG_EINVAL;
    p.x = attn_out; p.w = weight; p.bias = bias; p.y = y; p.resid = residual; p.M = (int)M; p.N = C; p.K = C;
    return linear_launch<dlsg::A_PLAIN, dlsg::E_SCATTER>(p, stream);
#else
    (void)attn_out; (void)weight; (void)bias; (void)residual; (void)y; (void)B; (void)D; (void)H; (void)W; (void)C;
    (void)window; (void)shift; (void)stream;
    return DLSG_EUNSUPPORTED;
#endif
}

int dlsg_swin_proj_scatter_f32(const float* attn_out, const float* weight, const float* bias, float* y,
                               int B, int D, int H, int W, int C, const int* window, const int* shift, void* stream) {
    return dlsg_swin_proj_scatter_residual_f32(attn_out, weight, bias, nullptr, y, B, D, H, W, C, window, shift, stream);
}

int dlsg_swin_mlp_f32(const float* x, const float* gamma, const float* beta, float eps,
                      const void* w1_f16, const float* b1, const void* w2_f16, const float* b2,
                      void* hidden_f16, float* y, long long M, int C, int H, int residual, void* stream) {
#if defined(__CUDACC__)
    if (M <= 0 || C <= 0 || H <= 0) return DLSG_EINVAL;
    if (!x || !gamma || !beta || !w1_f16 || !w2_f16 || !hidden_f16 || !y) return DLSG_EINVAL;
    if (M > 0x7fffffffLL || C % 32 != 0 || C > 256 || H % C != 0) return DLSG_EUNSUPPORTED;
    const size_t smem = dlsg::mlp_smem_bytes(C, H);
    if (smem > 113 * 1024 || dlsg::mlp_fc2_smem_bytes(C, 2) > 227 * 1024) return DLSG_EUNSUPPORTED;   // fc1: two CTAs per SM; fc2: one
    dlsg::MlpParams p{};
    p.x = x; p.gamma = gamma; p.beta = beta; p.eps = eps;
    p.w1h = static_cast<const half_t*>(w1_f16); p.b1 = b1;
    p.w2h = static_cast<const half_t*>(w2_f16); p.b2 = b2;
    p.hidden = static_cast<half_t*>(hidden_f16); p.y = y;
    p.M = (int)M; p.C = C; p.H = H; p.residual = residual ? 1 : 0;
    p.tmem_cols = 32;
    while (p.tmem_cols < C) p.tmem_cols *= 2;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = (p.M + dlsg::kLinTileM - 1) / dlsg::kLinTileM;
    const unsigned grid = (unsigned)(ntiles < 2 * sms ? ntiles : 2 * sms);
    if (dlsg::launch(dlsg::swin_ln_gemm_stream_kernel<false>, dim3(grid), dim3(dlsg::kLinThreads), smem, (dlsg::stream_t)stream, p) != 0) return DLSG_ELAUNCH;
    // measured (8 cine slices): two CTAs per SM covering each other's staging 0.188 ms, one CTA per SM with the two-stage ring 0.209 ms
    static const int fc2_stages = (getenv("DLSG_MLP_FC2_STAGES") && atoi(getenv("DLSG_MLP_FC2_STAGES")) == 2) ? 2 : 1;   // tuning switch
    if (fc2_stages == 1) {
        if (dlsg::launch(dlsg::swin_mlp_fc2_kernel<1>, dim3(grid), dim3(dlsg::kLinThreads), dlsg::mlp_fc2_smem_bytes(C, 1), (dlsg::stream_t)stream, p) != 0)
            return DLSG_ELAUNCH;
        return DLSG_OK;
    }
    const unsigned grid2 = (unsigned)(ntiles < sms ? ntiles : sms);
    if (dlsg::launch(dlsg::swin_mlp_fc2_kernel<2>, dim3(grid2), dim3(dlsg::kLinThreads), dlsg::mlp_fc2_smem_bytes(C, 2), (dlsg::stream_t)stream, p) != 0)
        return DLSG_ELAUNCH;
    return DLSG_OK;
#else
    (void)x; (void)gamma; (void)beta; (void)eps; (void)w1_f16; (void)b1; (void)w2_f16; (void)b2; (void)hidden_f16; (void)y;
    (void)M; (void)C; (void)H; (void)residual; (void)stream;
    return DLSG_EUNSUPPORTED;      // tcgen05 / TMEM have no host emulation
#endif
}

}  // extern "C"
