// TEST INFRASTRUCTURE: builds the product's kernel + dispatch sources with g++ under
// the thread-per-CUDA-thread emulation of csrc/cuda_compat.h, so pytest can run the
// very same index math on the CPU (tests/test_sense_host_emulation.py).  Never shipped.
#include "../../dl_swin_gan_b200/csrc/sense_api.cu"
#include "../../dl_swin_gan_b200/csrc/swin_api.cu"
#include "../../dl_swin_gan_b200/csrc/kt_api.cu"

namespace dlsg_emu {
thread_local dim3 t_threadIdx;
thread_local dim3 t_blockIdx;
thread_local BlockCtx* t_ctx = nullptr;
dim3 g_blockDim, g_gridDim;
}
