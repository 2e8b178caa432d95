// Compatibility layer so that the kernel sources (*.cuh) and the C-ABI dispatch
// (sense_api.cu) compile both with nvcc for sm_100a (the product) and with plain
// g++ (TEST INFRASTRUCTURE ONLY: tests/host builds an emulated copy where every
// CUDA thread of a block is an OS thread and __syncthreads/__syncwarp are real
// barriers; it exists to debug index math without a GPU and is never shipped or
// used as a fallback — the product library is the nvcc build and nothing else).
#pragma once

#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)

#include <cuda_runtime.h>
#define DLSG_KERNEL __global__
#define DLSG_DEVICE __device__ __forceinline__
#define DLSG_DYN_SMEM(type, name)                                       \
    extern __shared__ __align__(1024) unsigned char dlsg_dyn_smem_raw[];  \
    type* name = reinterpret_cast<type*>(dlsg_dyn_smem_raw)
#define DLSG_LAUNCH_BOUNDS(...) __launch_bounds__(__VA_ARGS__)

// 16-byte asynchronous global -> shared copy (LDGSTS), L1 bypassed; both addresses 16-byte aligned
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(sa), "l"(gmem_src) : "memory");
}
// 8-byte variant (one complex64), through L1: used for thread-private staging slots
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" :: "r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_group0() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
// all but the most recently committed group have landed
__device__ __forceinline__ void cp_async_wait_group1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }


// ---- TMA (bulk tensor copy) + mbarrier: a 2-D tile of a row-major fp32 matrix lands in shared memory as
// [box_rows][box_cols] (no swizzle), completion is counted in bytes on a shared-memory mbarrier.
#include <cuda.h>
namespace dlsg {
typedef CUtensorMap TileMap;            // built on the host by make_tile_map (sense_api.cu)
typedef unsigned long long mbar_t;
}
#define DLSG_GRID_CONSTANT __grid_constant__
__device__ __forceinline__ void mbar_init(dlsg::mbar_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
// makes the mbarrier initialisation visible to the async proxy (TMA engine)
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(dlsg::mbar_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
                 :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(dlsg::mbar_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(dlsg::mbar_t* bar, unsigned parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n"
                 :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
// tile whose first element is (row0, col0) of the matrix described by `tm` (col0 in fp32 elements) -> smem_dst (128-byte aligned)
__device__ __forceinline__ void tile_load_2d(void* smem_dst, const dlsg::TileMap* tm, int col0, int row0, dlsg::mbar_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(tm), "r"(col0), "r"(row0),
                    "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

#include <cuda_fp16.h>
typedef __half half_t;
// 2^x on the SFU (MUFU.EX2), flush-to-zero: exp2(-inf) = 0
__device__ __forceinline__ float fast_exp2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ half_t f2h(float v) { return __float2half_rn(v); }
// two floats -> packed fp16 pair (first argument in the low half)
__device__ __forceinline__ unsigned pack_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const unsigned*>(&h);
}
__device__ __forceinline__ unsigned lds_u32(const half_t* p) { return *reinterpret_cast<const unsigned*>(p); }
// four 8x8 b16 matrices from shared memory; lane L supplies the address of row L%8 of matrix L/8 (16-byte aligned).
// r[i] of lane t = (M_i[t/4][2(t%4)], M_i[t/4][2(t%4)+1]);  with .trans = (M_i[2(t%4)][t/4], M_i[2(t%4)+1][t/4])
__device__ __forceinline__ void ldmatrix_x4(unsigned* r, const half_t* row_ptr) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(row_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(unsigned* r, const half_t* row_ptr) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(row_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
// D(16x8, f32) += A(16x16, f16, row) * B(16x8, f16, col); fragment layouts per PTX ISA "mma.m16n8k16"
__device__ __forceinline__ void mma_m16n8k16_f16(float* c, const unsigned* a, unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

#include <atomic>
namespace dlsg {
using stream_t = cudaStream_t;
inline std::atomic<unsigned long long>& launch_counter() {
    static std::atomic<unsigned long long> n{0};   // statistics only (bench.py's gpu_launches); no behaviour depends on it
    return n;
}
template <typename K, typename... A>
inline int launch(K kernel, dim3 grid, dim3 block, size_t smem, stream_t s, A... args) {
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return -1; }
    }
    kernel<<<grid, block, smem, s>>>(args...);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
}  // namespace dlsg

#else  // ---------------------------------------------------------------- host emulation

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <mutex>
#include <condition_variable>

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

namespace dlsg_emu {
class Barrier {
public:
    explicit Barrier(int n) : n_(n), count_(0), gen_(0) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        int g = gen_;
        if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return g != gen_; });
    }
private:
    int n_, count_, gen_;
    std::mutex m_;
    std::condition_variable cv_;
};
struct WarpScratch { unsigned a[32][4]; unsigned b[32][2]; float f[32]; const void* ptr[32]; };
struct BlockCtx { Barrier* block_bar; std::vector<Barrier*> warp_bars; unsigned char* smem; WarpScratch* scratch; };
extern thread_local dim3 t_threadIdx;
extern thread_local dim3 t_blockIdx;
extern thread_local BlockCtx* t_ctx;
extern dim3 g_blockDim, g_gridDim;
}  // namespace dlsg_emu

#define threadIdx (dlsg_emu::t_threadIdx)
#define blockIdx (dlsg_emu::t_blockIdx)
#define blockDim (dlsg_emu::g_blockDim)
#define gridDim (dlsg_emu::g_gridDim)
static inline void __syncthreads() { dlsg_emu::t_ctx->block_bar->wait(); }
static inline void __syncwarp() { dlsg_emu::t_ctx->warp_bars[dlsg_emu::t_threadIdx.x >> 5]->wait(); }
static inline int __syncthreads_or(int pred) {
    static std::atomic<int> acc{0};              // blocks run one at a time in the emulation
    if (pred) acc.store(1);
    __syncthreads();
    const int r = acc.load();
    __syncthreads();
    if (dlsg_emu::t_threadIdx.x == 0) acc.store(0);
    __syncthreads();
    return r;
}
static inline void sincospif(float a, float* s, float* c) {
    const double pi = 3.14159265358979323846;
    *s = (float)std::sin(pi * (double)a); *c = (float)std::cos(pi * (double)a);
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
typedef _Float16 half_t;
static inline float fast_exp2(float x) { return std::exp2(x); }
static inline half_t f2h(float v) { return (half_t)v; }
static inline unsigned pack_h2(float lo, float hi) {
    half_t h[2] = {(half_t)lo, (half_t)hi};
    unsigned u; std::memcpy(&u, h, 4); return u;
}
static inline unsigned lds_u32(const half_t* p) { unsigned u; std::memcpy(&u, p, 4); return u; }
static inline float emu_half_of(unsigned u, int hi) { half_t h[2]; std::memcpy(h, &u, 4); return (float)h[hi]; }
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
    auto* ws = &dlsg_emu::t_ctx->scratch[dlsg_emu::t_threadIdx.x >> 5];
    const int lane = dlsg_emu::t_threadIdx.x & 31;
    ws->f[lane] = v;
    __syncwarp();
    const float r = ws->f[lane ^ lane_mask];
    __syncwarp();
    return r;
}
static inline float __shfl_sync(unsigned, float v, int src_lane) {
    auto* ws = &dlsg_emu::t_ctx->scratch[dlsg_emu::t_threadIdx.x >> 5];
    const int lane = dlsg_emu::t_threadIdx.x & 31;
    ws->f[lane] = v;
    __syncwarp();
    const float r = ws->f[src_lane & 31];
    __syncwarp();
    return r;
}
struct int2 { int x, y; };
// emulation of ldmatrix.m8n8.x4(.trans).b16 with the PTX ISA thread/row mapping
static inline void ldmatrix_x4_impl(unsigned* r, const half_t* row_ptr, bool trans) {
    auto* ws = &dlsg_emu::t_ctx->scratch[dlsg_emu::t_threadIdx.x >> 5];
    const int lane = dlsg_emu::t_threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    ws->ptr[lane] = row_ptr;
    __syncwarp();
    for (int i = 0; i < 4; ++i) {
        half_t e[2];
        for (int h = 0; h < 2; ++h) {
            const int row = trans ? 2 * t4 + h : g, col = trans ? g : 2 * t4 + h;
            e[h] = static_cast<const half_t*>(ws->ptr[8 * i + row])[col];
        }
        std::memcpy(&r[i], e, 4);
    }
    __syncwarp();
}
static inline void ldmatrix_x4(unsigned* r, const half_t* row_ptr) { ldmatrix_x4_impl(r, row_ptr, false); }
static inline void ldmatrix_x4_trans(unsigned* r, const half_t* row_ptr) { ldmatrix_x4_impl(r, row_ptr, true); }
// emulation of mma.sync.m16n8k16 (f16 x f16 -> f32) with the PTX ISA fragment layouts
static inline void mma_m16n8k16_f16(float* c, const unsigned* a, unsigned b0, unsigned b1) {
    auto* ws = &dlsg_emu::t_ctx->scratch[dlsg_emu::t_threadIdx.x >> 5];
    const int lane = dlsg_emu::t_threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    for (int i = 0; i < 4; ++i) ws->a[lane][i] = a[i];
    ws->b[lane][0] = b0; ws->b[lane][1] = b1;
    __syncwarp();
    auto A = [&](int row, int k) { return emu_half_of(ws->a[(row % 8) * 4 + (k % 8) / 2][(row >= 8 ? 1 : 0) + (k >= 8 ? 2 : 0)], k % 2); };
    auto B = [&](int k, int n) { return emu_half_of(ws->b[n * 4 + (k % 8) / 2][k >= 8 ? 1 : 0], k % 2); };
    for (int i = 0; i < 4; ++i) {
        const int row = g + (i >= 2 ? 8 : 0), col = 2 * t4 + (i & 1);
        float acc = c[i];
        for (int k = 0; k < 16; ++k) acc += A(row, k) * B(k, col);
        c[i] = acc;
    }
    __syncwarp();
}




// ---- TMA + mbarrier under the emulation: the copy happens at issue time; the mbarrier word holds
// (completed phases << 32) | (pending arrivals << 20) | pending bytes, so that waiters really wait
namespace dlsg {
struct TileMap { const float* base; long long pitch; long long rows; int box_cols, box_rows, swizzle128; };   // fp32 elements
typedef unsigned long long mbar_t;
}
namespace dlsg_emu {
inline std::mutex& mbar_mutex() { static std::mutex m; return m; }
inline std::vector<std::pair<const void*, unsigned>>& mbar_counts() { static std::vector<std::pair<const void*, unsigned>> v; return v; }
inline unsigned mbar_init_count(const void* bar) {
    for (auto& kv : mbar_counts()) if (kv.first == bar) return kv.second;
    return 1;
}
// caller holds the mutex: completes the phase when no arrival and no byte is pending
inline void mbar_settle(dlsg::mbar_t* bar, unsigned long long v) {
    if (((v >> 20) & 0xfffull) == 0 && (v & 0xfffffull) == 0)
        v = (((v >> 32) + 1) << 32) | ((unsigned long long)mbar_init_count(bar) << 20);
    __atomic_store_n(bar, v, __ATOMIC_SEQ_CST);
}
}  // namespace dlsg_emu
#define DLSG_GRID_CONSTANT
static inline void mbar_init(dlsg::mbar_t* bar, unsigned count) {
    std::lock_guard<std::mutex> g(dlsg_emu::mbar_mutex());
    bool found = false;
    for (auto& kv : dlsg_emu::mbar_counts()) if (kv.first == bar) { kv.second = count; found = true; }
    if (!found) dlsg_emu::mbar_counts().push_back({bar, count});
    __atomic_store_n(bar, (unsigned long long)count << 20, __ATOMIC_SEQ_CST);
}
static inline void mbar_fence_init() {}
static inline void mbar_arrive(dlsg::mbar_t* bar) {
    std::lock_guard<std::mutex> g(dlsg_emu::mbar_mutex());
    dlsg_emu::mbar_settle(bar, __atomic_load_n(bar, __ATOMIC_SEQ_CST) - (1ull << 20));
}
static inline void mbar_expect_tx(dlsg::mbar_t* bar, unsigned bytes) {      // = arrive.expect_tx
    std::lock_guard<std::mutex> g(dlsg_emu::mbar_mutex());
    dlsg_emu::mbar_settle(bar, __atomic_load_n(bar, __ATOMIC_SEQ_CST) - (1ull << 20) + bytes);
}
static inline void mbar_wait(dlsg::mbar_t* bar, unsigned parity) {
    while (((__atomic_load_n(bar, __ATOMIC_SEQ_CST) >> 32) & 1ull) == (unsigned long long)parity) std::this_thread::yield();
}
static inline void tile_load_2d(void* smem_dst, const dlsg::TileMap* tm, int col0, int row0, dlsg::mbar_t* bar) {
    float* dst = static_cast<float*>(smem_dst);
    for (int r = 0; r < tm->box_rows; ++r) {
        const float* src = tm->base + (size_t)(row0 + r) * tm->pitch + col0;
        if (!tm->swizzle128) { std::memcpy(dst + (size_t)r * tm->box_cols, src, sizeof(float) * tm->box_cols); continue; }
        for (int ch = 0; ch < tm->box_cols / 4; ++ch)            // SWIZZLE_128B: 16-byte chunk index XOR (row mod 8), rows of 128 bytes
            std::memcpy(dst + (size_t)r * 32 + 4 * (ch ^ (r & 7)), src + 4 * ch, 16);
    }
    const unsigned long long bytes = (unsigned long long)tm->box_rows * tm->box_cols * sizeof(float);
    std::lock_guard<std::mutex> g(dlsg_emu::mbar_mutex());
    dlsg_emu::mbar_settle(bar, __atomic_load_n(bar, __ATOMIC_SEQ_CST) - bytes);
}
static inline void cp_async16(void* smem_dst, const void* gmem_src) { std::memcpy(smem_dst, gmem_src, 16); }
static inline void cp_async8(void* smem_dst, const void* gmem_src) { std::memcpy(smem_dst, gmem_src, 8); }
static inline void cp_async_wait_all() {}
static inline void cp_async_commit() {}
static inline void cp_async_wait_group0() {}
static inline void cp_async_wait_group1() {}
using std::min;
using std::max;

#define DLSG_KERNEL
#define DLSG_DEVICE inline
#define DLSG_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(dlsg_emu::t_ctx->smem)
#define DLSG_LAUNCH_BOUNDS(...)

#include <atomic>
namespace dlsg {
using stream_t = void*;
inline std::atomic<unsigned long long>& launch_counter() {
    static std::atomic<unsigned long long> n{0};
    return n;
}
template <typename K, typename... A>
inline int launch(K kernel, dim3 grid, dim3 block, size_t smem, stream_t, A... args) {
    using namespace dlsg_emu;
    launch_counter().fetch_add(1, std::memory_order_relaxed);
    if (smem > 227 * 1024 || block.x > 1024 || block.y != 1 || block.z != 1) return -1;
    g_blockDim = block; g_gridDim = grid;
    const size_t smem_alloc = ((smem + 255) / 128) * 128;
    unsigned char* smem_buf = static_cast<unsigned char*>(std::aligned_alloc(128, smem_alloc));
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        Barrier block_bar((int)block.x);
        unsigned nwarps = (block.x + 31) / 32;
        std::vector<Barrier*> wb;
        for (unsigned w = 0; w < nwarps; ++w) {
            unsigned lanes = (w + 1) * 32 <= block.x ? 32 : block.x - w * 32;
            wb.push_back(new Barrier((int)lanes));
        }
        std::vector<WarpScratch> scratch(nwarps);
        BlockCtx ctx{&block_bar, wb, smem_buf, scratch.data()};
        std::memset(smem_buf, 0xCD, smem_alloc);   // poison: catches reads of unwritten smem
        std::vector<std::thread> th;
        for (unsigned t = 0; t < block.x; ++t) {
            th.emplace_back([&, t] {
                t_threadIdx = dim3(t, 0, 0); t_blockIdx = dim3(bx, by, bz); t_ctx = &ctx;
                kernel(args...);
            });
        }
        for (auto& x : th) x.join();
        for (auto* b : wb) delete b;
    }
    std::free(smem_buf);
    return 0;
}
}  // namespace dlsg

#endif
