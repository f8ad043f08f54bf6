// misc_kernels.cu -- the two remaining MatmulOperator methods the reference's kernels/cuda directory provides besides the
// GEMV and the int8 family:
//   * naive_mat_mul_fp16_int4 (kernels/cuda/matmul_int4.cu:8-48): the fp16-accumulate reference in the AWQ-GEMM layout
//     (B int32[IC][OC/8], nibble order 0 2 4 6 1 3 5 7, scales half[IC/G][OC], zero fixed 8).  The reference runs it on
//     the host with half_float::half, i.e. every binary op is evaluated in float and rounded to half; reproduced here
//     op for op (float op + round-to-nearest-even), one thread per output, serial k -> bit-identical results.
//   * mat_mul_accelerator_transposed_fastover_column (kernels/cuda/matmul_ref_fp32.cc): C = A * B^T in fp32, serial k.
#include "common.cuh"
#include "kernels.h"

namespace tce {
namespace {

TCE_DEVINL float rh(float v) { return __half2float(__float2half_rn(v)); }  // round to half, keep as float

__global__ void naive_fp16_int4_kernel(const __half *__restrict__ A, const int32_t *__restrict__ B, const __half *__restrict__ scales,
                                       __half *__restrict__ C, int M, int IC, int OC, int block) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= OC || i >= M) return;
    const int shift_of[8] = {0, 16, 4, 20, 8, 24, 12, 28};
    const int sh = shift_of[j & 7];
    float acc = 0.f;  // always holds a half-representable value
    for (int k = 0; k < IC; k++) {
        const float s = __half2float(scales[(size_t)(k / block) * OC + j]);
        const float in = __half2float(A[(size_t)i * IC + k]);
        const uint32_t word = (uint32_t)B[(size_t)k * (OC / 8) + (j >> 3)];
        const float q = (float)((word >> sh) & 0xF);  // 0..15 exact in half
        const float d = rh(q - 8.0f);
        const float wv = rh(d * s);
        const float prod = rh(in * wv);
        acc = rh(acc + prod);
    }
    C[(size_t)i * OC + j] = __float2half_rn(acc);
}

__global__ void f32_matmul_transposed_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ C, int M, int N, int K) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= N || i >= M) return;
    float acc = 0.f;
    for (int k = 0; k < K; k++) acc = __fadd_rn(acc, __fmul_rn(A[(size_t)i * K + k], B[(size_t)j * K + k]));  // no FMA: as written
    C[(size_t)i * N + j] = acc;
}

}  // namespace

cudaError_t launch_naive_fp16_int4(Ctx *ctx, const __half *A, const int32_t *B, const __half *scales, __half *C, int M, int IC, int OC, int block) {
    dim3 grid((OC + 127) / 128, M);
    naive_fp16_int4_kernel<<<grid, 128, 0, ctx->stream>>>(A, B, scales, C, M, IC, OC, block);
    return cudaGetLastError();
}
cudaError_t launch_f32_matmul_transposed(Ctx *ctx, const float *A, const float *B, float *C, int M, int N, int K) {
    dim3 grid((N + 127) / 128, M);
    f32_matmul_transposed_kernel<<<grid, 128, 0, ctx->stream>>>(A, B, C, M, N, K);
    return cudaGetLastError();
}

}  // namespace tce
