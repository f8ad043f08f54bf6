// matmul.h -- operator surface of libtce_host.so: the CUDA-backend half of the reference's `matmul::MatmulOperator`, backed by the
// C ABI of libtce_b200.so.
//
// The four parameter structs below are binary-compatible (same fields, same order, same types, QM_CUDA flavour: float16_t = half)
// with the reference's kernels/matmul.h:52-100, so objects filled by the reference's llm/ call sites (Linear_half_int4::forward
// llm/src/ops/cuda/linear.cu:5-40, W8A8B8O8Linear::forward llm/src/ops/W8A8B8O8Linear.cc:38-78, the BMM_S8T_* wrappers, ...) can be
// handed to this library unchanged; tests/test_host_header_abi.py compiles an offsetof/sizeof probe against both headers.
// Only the methods the reference's kernels/cuda/ directory defines are declared (and defined in matmul_operator.cu), plus the
// gemm_forward_cuda* slot it declares without defining.  Backend-independent methods (naive_mat_mul_int4*, naive_mat_mul_int8,
// mat_mul_transposed, CHECK_MATRICES*) keep coming from the reference's own kernels/*.cc and header, see INTEGRATION.md.
#ifndef TCE_HOST_MATMUL_H
#define TCE_HOST_MATMUL_H
#include <cuda_fp16.h>
#include <stdint.h>

// host-side fp16 type of the reference (half_float::half) when its header is reachable, a storage-only stand-in otherwise
#if defined(__has_include)
#if __has_include("half.hpp")
#include "half.hpp"
typedef half_float::half naive_float16_t;
#define TCE_HAVE_HALF_HPP 1
#endif
#endif
#ifndef TCE_HAVE_HALF_HPP
struct naive_float16_t { uint16_t bits; };
#endif
typedef half float16_t;

// ---- parameter blocks (layout contract; do not reorder) ----
struct quantization_params { float scale; bool per_channel = false; int32_t zero_point; int8_t q_min = -128, q_max = 127; };

struct matrix {
    int row, column;                               // meaning is per call site
    float *data_ptr; float16_t *half_data_ptr; naive_float16_t *fp16_data_ptr;   // exactly one typed pointer is set per op
    int32_t *int32_data_ptr; int8_t *int8_data_ptr; uint8_t *uint8_data_ptr, *int4_data_ptr;
    struct quantization_params qparams;
    int length() { return row * column; }
};

struct optimization_params { int blk_size; int num_thread = 8; };

struct matmul_params {
    struct matrix A, B, C, bias;
    struct optimization_params opt_params;
    float alpha, beta; float16_t half_alpha;
    float *scales, *offset, *zero_point;           // W4 CPU formats
    float16_t *half_scales;                        // W4 QM_CUDA: half[OC][zeros_w * 8]
    naive_float16_t *fp16_scales;                  // AWQ-GEMM layout of the fp16 host reference
    int *int32_zero_point;                         // W4 QM_CUDA: uint32[OC][zeros_w]
    int block_size;
    float *A_scales; int8_t A_zero_point;          // W4A8 CPU path
};

namespace matmul {

// X(name): takes `const struct matmul_params *`
#define TCE_MATMUL_OPS(X)                                                                                                            \
    /* W4A16 decode GEMV, kernels/cuda/gemv_cuda.cu:213-260 */ X(gemv_forward_cuda)                                                 \
    /* fp16 host reference in the AWQ-GEMM layout, kernels/cuda/matmul_int4.cu:8-48 */ X(naive_mat_mul_fp16_int4)                    \
    /* fp32 C = A * B^T, kernels/cuda/matmul_ref_fp32.cc:11-34 */ X(mat_mul_accelerator_transposed_fastover_column)                  \
    /* empty in the reference CUDA build, gemv_cuda.cu:262-268 */ X(mat_mul_accelerator_int4_fast) X(mat_mul_accelerator_int4_fast_no_offset) \
    /* W8A8 family, kernels/ref/matmul_ref_int8.cc:11-192 */                                                                       \
    X(mat_mul_accelerator_int8_fast_2x2_32unroll) X(mat_mul_accelerator_int8_fast_32unroll_over_column)                              \
    X(mat_mul_accelerator_int8_fast_2x2_32unroll_nobias) X(mat_mul_accelerator_int8_fast_2x2_32unroll_nobias_batch)                  \
    X(mat_mul_accelerator_int8_fast_2x2_32unroll_nobias_ofp32) X(mat_mul_accelerator_int8_fast_2x2_32unroll_nobias_ofp32_batch)      \
    X(mat_mul_accelerator_int8_fast_2x2_32unroll_bfp32_ofp32) X(mat_mul_accelerator_int8_fast_2x2_32unroll_bfp32_ofp32_over_column)

class MatmulOperator {  // stateless, constructed per call by the reference's wrappers
   public:
#define TCE_DECLARE_OP(name) void name(const struct matmul_params *params);
    TCE_MATMUL_OPS(TCE_DECLARE_OP)
#undef TCE_DECLARE_OP
    // the prefill GEMM slot (kernels/matmul.h:142-145: declared, never defined by the reference): tcgen05 GEMM for M >= 16
    void gemm_forward_cuda(const struct matmul_params *params, int split_k_iters);
    void gemm_forward_cuda_8splits(const struct matmul_params *params, float16_t *split_8_buffer);
    void gemm_forward_cuda_half(const struct matmul_params *params, int split_k_iters);
    void gemm_forward_cuda_half_test(const struct matmul_params *params, int split_k_iters);
};

}  // namespace matmul

// Library-wide context used by the adapters: device 0, the legacy default stream (the reference launches every kernel on stream 0
// and relies on in-order semantics, SURVEY.md 8(b) Threading).  tce_host_set_stream() moves it.
struct tce_ctx;
extern "C" tce_ctx *tce_host_ctx(void);
extern "C" void tce_host_set_stream(void *cuda_stream);

#endif
