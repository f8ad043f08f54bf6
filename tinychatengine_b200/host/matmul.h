// matmul.h -- drop-in operator surface for the CUDA backend of TinyChatEngine, backed by libtce_b200.so.
//
// Layout-compatible (field for field, QM_CUDA flavour) with the reference's kernels/matmul.h:52-153 so that the
// reference's llm/ call sites (Linear_half_int4::forward llm/src/ops/cuda/linear.cu:5-40, W8A8B8O8Linear::forward
// llm/src/ops/W8A8B8O8Linear.cc:38-78, BMM_S8T_* ...) compile and link against this library unchanged.  Only the
// methods the CUDA backend directory provides are implemented here (matmul_operator.cu); the backend-independent
// ones (naive_mat_mul_int4*, naive_mat_mul_int8, mat_mul_transposed, CHECK_MATRICES*) stay with the reference's own
// kernels/*.cc exactly as in its build (llm/Makefile:24,64-65).
#ifndef TCE_HOST_MATMUL_H
#define TCE_HOST_MATMUL_H
#include <cuda_fp16.h>
#include <stdint.h>

#if defined(__has_include)
#if __has_include("half.hpp")
#include "half.hpp"
typedef half_float::half naive_float16_t;
#define TCE_HAVE_HALF_HPP 1
#endif
#endif
#ifndef TCE_HAVE_HALF_HPP
struct naive_float16_t { uint16_t bits; };  // storage-only stand-in when half.hpp is not on the include path
#endif
typedef half float16_t;  // QM_CUDA (reference matmul.h:14-18)

struct quantization_params {
    float scale;
    bool per_channel = false;
    int32_t zero_point;
    int8_t q_min = -128, q_max = 127;
};

struct matrix {
    int row;
    int column;
    float *data_ptr;
    float16_t *half_data_ptr;
    naive_float16_t *fp16_data_ptr;
    int32_t *int32_data_ptr;
    int8_t *int8_data_ptr;
    uint8_t *uint8_data_ptr;
    uint8_t *int4_data_ptr;
    struct quantization_params qparams;
    int length() { return row * column; }
};

struct optimization_params {
    int blk_size;
    int num_thread = 8;
};

struct matmul_params {
    struct matrix A, B, C, bias;
    struct optimization_params opt_params;
    float alpha, beta;
    float16_t half_alpha;
    float *scales, *offset, *zero_point;   // int4 (CPU formats)
    float16_t *half_scales;                // int4 QM_CUDA: half[OC][zeros_w*8]
    naive_float16_t *fp16_scales;
    int *int32_zero_point;                 // int4 QM_CUDA: uint32[OC][zeros_w]
    int block_size;
    float *A_scales;                       // int8 activations (CPU W4A8)
    int8_t A_zero_point;
};

struct thread_args {
    const struct matrix *A;
    const struct matrix *B;
    const struct matrix *C;
    const struct matmul_params *params;
    int start_i, end_i, blk_size;
};

#ifndef MAX
#define MAX(A, B) ((A) > (B) ? (A) : (B))
#endif
#ifndef MIN
#define MIN(A, B) ((A) < (B) ? (A) : (B))
#endif

namespace matmul {
// Same member list as the reference class (stateless; callers construct a temporary per call).
class MatmulOperator {
   public:
    void mat_mul_transposed(const struct matmul_params *params);
    void mat_mul_accelerator_transposed_fastover_column(const struct matmul_params *params);
    void mat_mul_accelerator_transposed_fastover_column_bias(const struct matmul_params *params);
    void mat_mul_accelerator_untransposed_fastover_column(const struct matmul_params *params);
    void naive_mat_mul_int8(const struct matmul_params *params);
    void mat_mul_accelerator_int8_fast_32unroll_over_column(const struct matmul_params *params);
    void mat_mul_accelerator_int8_fast_2x2_32unroll(const struct matmul_params *params);
    void mat_mul_accelerator_int8_fast_2x2_32unroll_nobias(const struct matmul_params *params);
    void mat_mul_accelerator_int8_fast_2x2_32unroll_nobias_batch(const struct matmul_params *params);
    void mat_mul_accelerator_int8_fast_2x2_32unroll_nobias_ofp32(const struct matmul_params *params);
    void mat_mul_accelerator_int8_fast_2x2_32unroll_nobias_ofp32_batch(const struct matmul_params *params);
    void mat_mul_accelerator_int8_fast_2x2_32unroll_bfp32_ofp32(const struct matmul_params *params);
    void mat_mul_accelerator_int8_fast_2x2_32unroll_bfp32_ofp32_over_column(const struct matmul_params *params);
    void mat_mul_accelerator_int4_fast(const struct matmul_params *params);
    void mat_mul_accelerator_int4_fast_no_offset(const struct matmul_params *params);
    void mat_mul_accelerator_int8_int4_fast_no_offset(struct matmul_params *params);
    void gemv_accelerator_int8_int4_fast_no_offset(struct matmul_params *params);
    void gemm_accelerator_int8_int4_fast_no_offset(struct matmul_params *params);
    void gemm_accelerator_int8_int4_fast_no_offset_v2(struct matmul_params *params);
    void cblas_gemm_accelerator_no_offset(struct matmul_params *params);
    void naive_mat_mul_int4(const struct matmul_params *params);
    void naive_mat_mul_int4_with_offset(const struct matmul_params *params);
    void naive_mat_mul_fp16_int4(const struct matmul_params *params);
    void mat_mul_cuda(const struct matmul_params *params);
    void gemm_forward_cuda(const struct matmul_params *params, int split_k_iters);
    void gemm_forward_cuda_8splits(const struct matmul_params *params, float16_t *split_8_buffer);
    void gemm_forward_cuda_half(const struct matmul_params *params, int split_k_iters);
    void gemm_forward_cuda_half_test(const struct matmul_params *params, int split_k_iters);
    void gemv_forward_cuda(const struct matmul_params *params);

   private:
    float interval_to_us(struct timeval *start, struct timeval *end);
    void CHECK_MATRICES(const struct matrix *A, const struct matrix *B, const struct matrix *C);
    void CHECK_MATRICES_int4weight(const struct matrix *A, const struct matrix *B, const struct matrix *C);
};
}  // namespace matmul

// Library-wide context used by the adapters: device 0, the legacy default stream (the reference launches every
// kernel on stream 0 and relies on in-order semantics, SURVEY.md 8(b) Threading).  tce_host_set_stream() moves it.
struct tce_ctx;
extern "C" tce_ctx *tce_host_ctx(void);
extern "C" void tce_host_set_stream(void *cuda_stream);

#endif
