// Microbenchmark (not product): FP32 issue rates on sm_100a that decide how the SENSE butterflies are written.
//   scalar FFMA with three register operands, FFMA with an immediate, FADD, and the packed FFMA2 / FADD2
//   (PTX fma.rn.f32x2 / add.rn.f32x2) in their pair, broadcast and constant-pair operand forms;
//   also checks the packed instructions' arithmetic against scalar code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_issue fp32_issue.cu && ./fp32_issue
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float2 unpk(u64 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

constexpr int NACC = 8;      // independent accumulators per thread (ILP)
constexpr int INNER = 64;    // unrolled instructions per accumulator per loop trip

// MODE 0: FFMA reg,reg,reg   1: FFMA reg,imm,reg   2: FADD reg,reg   3: FFMA2 pair,pair,pair   4: FFMA2 bcast,pair,pair
// MODE 5: FADD2 pair,pair    6: FFMA2 pair,constpair(regs),pair  7: FMUL reg,imm
template <int MODE>
__global__ void rate(float* out, int trips, float s0, float s1) {
    float a[NACC], b = s0 + threadIdx.x * 1e-9f, c = s1;
    u64 p[NACC], q = pk(s0, s1), r = pk(s1, s0);
#pragma unroll
    for (int i = 0; i < NACC; ++i) { a[i] = 0.001f * (i + threadIdx.x); p[i] = pk(a[i], -a[i]); }
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int j = 0; j < INNER; ++j) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) {
                if (MODE == 0) a[i] = fmaf(a[i], b, c);
                if (MODE == 1) a[i] = fmaf(a[i], 0.999f, c);
                if (MODE == 2) a[i] = a[i] + c;
                if (MODE == 3) p[i] = fma2(p[i], q, r);
                if (MODE == 4) { float2 u = unpk(p[i]); p[i] = fma2(pk(u.x, u.x), q, r); }
                if (MODE == 5) p[i] = add2(p[i], r);
                if (MODE == 6) p[i] = fma2(p[i], pk(0.999f, 0.998f), r);
                if (MODE == 7) a[i] = a[i] * 0.999f;
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { float2 u = unpk(p[i]); s += a[i] + u.x + u.y; }
    if (s == 12345.678f) out[0] = s;
}

// arithmetic check of the packed forms: out[0..] = packed results, ref[0..] = scalar results
__global__ void semantics(float* out, float* ref, float ax, float ay, float bx, float by, float cx, float cy) {
    u64 a = pk(ax, ay), b = pk(bx, by), c = pk(cx, cy);
    float2 r;
    r = unpk(fma2(a, b, c));                       out[0] = r.x; out[1] = r.y;   ref[0] = fmaf(ax, bx, cx); ref[1] = fmaf(ay, by, cy);
    r = unpk(fma2(pk(ax, ax), b, c));              out[2] = r.x; out[3] = r.y;   ref[2] = fmaf(ax, bx, cx); ref[3] = fmaf(ax, by, cy);
    r = unpk(fma2(a, pk(0.5f, 0.25f), c));         out[4] = r.x; out[5] = r.y;   ref[4] = fmaf(ax, 0.5f, cx); ref[5] = fmaf(ay, 0.25f, cy);
    r = unpk(add2(a, b));                          out[6] = r.x; out[7] = r.y;   ref[6] = ax + bx; ref[7] = ay + by;
    r = unpk(fma2(pk(ay, ay), pk(-by, bx), c));    out[8] = r.x; out[9] = r.y;   ref[8] = fmaf(ay, -by, cx); ref[9] = fmaf(ay, bx, cy);
    r = unpk(mul2(a, pk(2.f, -2.f)));              out[10] = r.x; out[11] = r.y; ref[10] = 2.f * ax; ref[11] = -2.f * ay;
}

template <int MODE>
void run(const char* name, int flops_per_instr) {
    float* d; CK(cudaMalloc(&d, 4));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int trips = 200;
    for (int warps : {4, 8, 16, 32}) {        // resident warps per SM (one CTA per SM)
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        rate<MODE><<<pr.multiProcessorCount, warps * 32>>>(d, 2, 1.0f, 0.5f);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        rate<MODE><<<pr.multiProcessorCount, warps * 32>>>(d, trips, 1.0f, 0.5f);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double winstr = (double)trips * INNER * NACC * warps;                  // warp-instructions per SM
        const double clk = pr.clockRate * 1e3;                                       // nominal max SM clock (Hz)
        printf("%-34s warps/SM %2d : %.3f warp-instr/clk/SM (at %.0f MHz nominal), %.1f TFLOP/s\n", name, warps,
               winstr / (ms * 1e-3 * clk), clk / 1e6, winstr * 32 * flops_per_instr * pr.multiProcessorCount / (ms * 1e-3) / 1e12);
    }
    CK(cudaFree(d));
}

int main() {
    float *o, *r; CK(cudaMallocManaged(&o, 64)); CK(cudaMallocManaged(&r, 64));
    semantics<<<1, 1>>>(o, r, 1.25f, -2.5f, 3.0f, 0.75f, 10.f, 20.f);
    CK(cudaDeviceSynchronize());
    int bad = 0;
    for (int i = 0; i < 12; ++i) { if (o[i] != r[i]) ++bad; printf("semantics[%d] packed %g scalar %g %s\n", i, o[i], r[i], o[i] == r[i] ? "ok" : "MISMATCH"); }
    printf("packed f32x2 semantics: %s\n", bad ? "MISMATCH" : "all equal");
    run<0>("FFMA  reg,reg,reg", 2);
    run<1>("FFMA  reg,imm,reg", 2);
    run<2>("FADD  reg,reg", 1);
    run<7>("FMUL  reg,imm", 1);
    run<3>("FFMA2 pair,pair,pair", 4);
    run<4>("FFMA2 bcast,pair,pair", 4);
    run<6>("FFMA2 pair,constpair,pair", 4);
    run<5>("FADD2 pair,pair", 2);
    return 0;
}
