// ref_callsites_cuda_shim.cu -- TEST INFRASTRUCTURE ONLY.  Doorway into the reference's CUDA-build call site of the W4A16 GEMV,
// Linear_half_int4 (llm/include/ops/linear.h:186-221, forward: llm/src/ops/cuda/linear.cu:5-40), compiled UNCHANGED with -DQM_CUDA
// against the reference's own kernels/matmul.h and linked with this repo's MatmulOperator definitions (tinychatengine_b200/host/
// matmul_operator.cu -> libtce_b200.so) instead of the reference's kernels/cuda directory.  Together with oracle/ref_modules_shim.cc
// (built into the same library) this is the drop-in proof: the reference's wrappers and modules run on this library without a
// source change.  The constructor loads weight_int4.bin / scaling_factor_int4.bin / zero_point_int4.bin from `weight_path` into
// cudaMallocManaged buffers exactly as the reference does.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstring>
#include <string>

#include "operators.h"
#include "utils.h"

extern "C" __attribute__((visibility("default"))) int ref_linear_half_int4(const char *weight_path, int OC, int IC, int M, const void *x_half_host,
                                                                           void *y_half_host) {
    int *wptr = nullptr;
    allocate_aligned_memory_gpu(wptr, (size_t)OC * (IC / 8) * sizeof(int));  // as Int4llamaAttention's constructor does for its projections
    Matrix3D<int> weight(wptr, 1, OC, IC / 8);
    Linear_half_int4 op(weight, std::string(weight_path));
    float16_t *x = nullptr, *y = nullptr;
    allocate_aligned_memory_gpu(x, (size_t)M * IC * sizeof(float16_t));
    allocate_aligned_memory_gpu(y, (size_t)M * OC * sizeof(float16_t));
    memcpy(x, x_half_host, (size_t)M * IC * sizeof(float16_t));
    Matrix3D<float16_t> X(x, 1, M, IC), Y(y, 1, M, OC);
    for (int m = 0; m < M; m++) {  // the reference's wrapper asserts nothing about M but its kernel grid covers M rows (gemv_cuda.cu:236)
        (void)m;
    }
    op.forward(X, Y);
    const cudaError_t e = cudaDeviceSynchronize();  // the reference synchronises once per forward pass (Int4LlamaForCausalLM::forward)
    memcpy(y_half_host, y, (size_t)M * OC * sizeof(float16_t));
    cudaFree(wptr);
    cudaFree(x);
    cudaFree(y);
    cudaFree(op.scale.m_data);
    cudaFree(op.zero_point.m_data);
    return (int)e;
}
