// Microbenchmark (not product): L2-resident global bandwidth and DSMEM (cluster shared memory) bandwidth on one GPU.
// Decides whether the row<->column transposition of a fused 2-D FFT should go through L2 scratch or through DSMEM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_dsmem_bw l2_dsmem_bw.cu && ./l2_dsmem_bw
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float4 ldcg(const float4* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// every CTA streams over its own private region of `per_cta` float4 (so the footprint is grid * per_cta * 16 B)
__global__ void l2_read(const float4* buf, size_t per_cta, int reps, float* sink) {
    const float4* p = buf + (size_t)blockIdx.x * per_cta;
    float acc = 0.f;
    for (int r = 0; r < reps; ++r)
        for (size_t i = threadIdx.x; i < per_cta; i += blockDim.x) { float4 v = ldcg(p + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) sink[0] = acc;
}
__global__ void l2_write(float4* buf, size_t per_cta, int reps) {
    float4* p = buf + (size_t)blockIdx.x * per_cta;
    for (int r = 0; r < reps; ++r)
        for (size_t i = threadIdx.x; i < per_cta; i += blockDim.x) p[i] = make_float4(r, i, 0.f, 1.f);
}
// transposing pattern: write rows, barrier, read columns (8-byte elements, 32-column slabs), like the FFT passes
__global__ void l2_rw(float4* buf, size_t per_cta, int reps, float* sink) {
    float4* p = buf + (size_t)blockIdx.x * per_cta;
    float acc = 0.f;
    for (int r = 0; r < reps; ++r) {
        for (size_t i = threadIdx.x; i < per_cta; i += blockDim.x) p[i] = make_float4(r, i, 0.f, 1.f);
        __syncthreads();
        for (size_t i = threadIdx.x; i < per_cta; i += blockDim.x) { float4 v = ldcg(p + (i * 17) % per_cta); acc += v.x + v.w; }
        __syncthreads();
    }
    if (acc == 123.456f) sink[0] = acc;
}

__device__ __forceinline__ unsigned map_peer(const void* p, unsigned rank) {
    unsigned a = (unsigned)__cvta_generic_to_shared(p), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
template <int MODE>   // 0 = remote store, 1 = remote load, 2 = local store (control)
__global__ void dsmem_bw(int reps, int n4, float* sink, long long* cycles) {
    extern __shared__ float4 sm[];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned peer = cl.block_rank() ^ 1u;
    const unsigned base = (MODE == 2) ? (unsigned)__cvta_generic_to_shared(sm) : map_peer(sm, peer);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) sm[i] = make_float4(i, 1.f, 2.f, 3.f);
    cl.sync();
    float acc = 0.f;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            const unsigned a = base + 16u * i;
            if (MODE == 1) {
                float4 v;
                asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
                acc += v.x + v.w;
            } else {
                asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(a), "f"((float)r), "f"(1.f), "f"(2.f), "f"(acc) : "memory");
            }
        }
    }
    cl.sync();
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, L2 %d MB\n", prop.name, sms, prop.l2CacheSize >> 20);
    float* sink; CK(cudaMalloc(&sink, 4));
    float4* buf; const size_t cap = (size_t)512 << 20; CK(cudaMalloc(&buf, cap)); CK(cudaMemset(buf, 0, cap));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int kb_list[] = {128, 256, 512};
    for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm)
        for (int kb : kb_list) {
            const int grid = sms * ctas_per_sm, threads = 512, reps = 200;
            const size_t per = (size_t)kb * 1024 / 16;
            const double total = (double)grid * per * 16 * reps;
            for (int mode = 0; mode < 3; ++mode) {
                for (int it = 0; it < 2; ++it) {
                    CK(cudaEventRecord(e0));
                    if (mode == 0) l2_read<<<grid, threads>>>(buf, per, reps, sink);
                    else if (mode == 1) l2_write<<<grid, threads>>>(buf, per, reps);
                    else l2_rw<<<grid, threads>>>(buf, per, reps, sink);
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                }
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                const double bytes = total * (mode == 2 ? 2 : 1);
                printf("L2 %-5s grid %3d x %d thr, %4d KB/CTA (footprint %6.1f MB): %8.1f GB/s\n",
                       mode == 0 ? "read" : mode == 1 ? "write" : "w+r", grid, threads, kb, grid * per * 16 / 1e6, bytes / ms / 1e6);
            }
        }
    // DRAM-sized control (footprint 444 MB)
    {
        const int grid = sms * 2, threads = 512, reps = 4; const size_t per = (size_t)1536 * 1024 / 16;
        CK(cudaEventRecord(e0)); l2_read<<<grid, threads>>>(buf, per, reps, sink); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("DRAM read control (footprint %.0f MB): %.1f GB/s\n", grid * per * 16 / 1e6, (double)grid * per * 16 * reps / ms / 1e6);
    }

    long long* cyc; CK(cudaMalloc(&cyc, sizeof(long long) * 1024));
    long long hc[1024];
    const int n4 = 64 * 1024 / 16, reps = 200;
    auto run = [&](auto kern, const char* name, int threads, int grid) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 64 * 1024;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        float ms = 0;
        for (int it = 0; it < 2; ++it) {
            CK(cudaEventRecord(e0));
            CK(cudaLaunchKernelEx(&cfg, kern, reps, n4, sink, cyc));
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        CK(cudaMemcpy(hc, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
        double mean = 0; for (int i = 0; i < grid; ++i) mean += hc[i]; mean /= grid;
        printf("DSMEM %-12s %4d thr, grid %3d: %6.1f B/clk per SM (mean %0.f clk), aggregate %8.1f GB/s\n", name, threads, grid,
               (double)n4 * 16 * reps / mean, mean, (double)grid * n4 * 16 * reps / ms / 1e6);
    };
    for (int threads : {256, 512, 1024}) {
        run(dsmem_bw<0>, "remote store", threads, sms);
        run(dsmem_bw<1>, "remote load", threads, sms);
        run(dsmem_bw<2>, "local store", threads, sms);
    }
    run(dsmem_bw<0>, "remote store", 512, 2);
    run(dsmem_bw<1>, "remote load", 512, 2);
    return 0;
}
