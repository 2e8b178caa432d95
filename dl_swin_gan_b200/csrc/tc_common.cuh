// tcgen05 / tensor-memory helpers shared by the tensor-core kernels (swin_linear_tc.cuh, swin_mlp_tc.cuh, conv3d_tc.cuh):
// shared-memory matrix descriptors, the kind::f16 MMA issue, TMEM loads.  CUDA-only.
#pragma once

#include "cuda_compat.h"

#if defined(__CUDACC__)
namespace dlsg {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle, K-major (cute::UMMA::SmemDescriptor, version 1 = sm_100)
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr, unsigned lbo_bytes, unsigned sbo_bytes) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3fff);
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;                 // descriptor version
    return d;                        // layout_type (bits 61-63) = 0: SWIZZLE_NONE
}

// instruction descriptor for kind::f16: A = B = F16 (0), D = F32 (1), both K-major, dense
__device__ __forceinline__ unsigned umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                         unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_ld32(unsigned taddr, float* v) {
    unsigned* r = reinterpret_cast<unsigned*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_ld4(unsigned taddr, float* v) {
    unsigned* r = reinterpret_cast<unsigned*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

}  // namespace dlsg
#endif
