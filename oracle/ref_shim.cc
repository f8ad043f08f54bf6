// ref_shim.cc -- extern "C" doorway into the UNMODIFIED reference kernels (compiled in place from
// /root/reference/kernels by oracle/Makefile into oracle/_ref/*.so).  TEST INFRASTRUCTURE ONLY.
//
// Nothing here re-implements arithmetic: each function fills a `matmul_params` exactly the way the
// reference call site does (cited per function) and calls the reference's own MatmulOperator method.
// The oracle restatement (tce_oracle.c) is validated against these, and golden fixtures under
// tests/golden/ are generated from them (tests/golden/make_golden.py).
#include <cstdint>
#include <cstring>

#include "matmul.h"  // -I/root/reference/kernels

using matmul::MatmulOperator;

extern "C" {

// Which reference build is this?  0 = generic (kernels/ref, no QM_*), 1 = AVX (QM_x86)
int ref_build_kind() {
#ifdef QM_x86
    return 1;
#else
    return 0;
#endif
}

// call site: Linear_FP_int4::forward_ref, llm/src/ops/linear.cc:80-117 (B.row=OC, B.column=IC/2)
void ref_naive_mat_mul_int4(const float *A, const uint8_t *B, const float *scales, const float *zero_point,
                            const float *offset, float *C, int M, int IC, int OC, int block_size, int with_offset) {
    struct matmul_params p;
    memset(&p, 0, sizeof(p));
    p.A.row = M;
    p.A.column = IC;
    p.A.data_ptr = const_cast<float *>(A);
    p.B.row = OC;
    p.B.column = IC / 2;
    p.B.int4_data_ptr = const_cast<uint8_t *>(B);
    p.C.row = M;
    p.C.column = OC;
    p.C.data_ptr = C;
    p.scales = const_cast<float *>(scales);
    p.offset = const_cast<float *>(offset);
    p.zero_point = const_cast<float *>(zero_point);
    p.block_size = block_size;
    MatmulOperator op;
    if (with_offset)
        op.naive_mat_mul_int4_with_offset(&p);
    else
        op.naive_mat_mul_int4(&p);
}

// INT8 family.  call sites: W8A8B8O8Linear.cc:38-78, W8A8BFP32OFP32Linear.cc, BMM_S8T_S8N_{F32T,S8T}.cc
// (B.row=K, B.column=N, B stored [N][K]).  variant:
//  0 2x2_32unroll (bias int8, int8 out)      1 32unroll_over_column (same, M==1 flavour)
//  2 nobias                                   3 nobias_batch
//  4 bfp32_ofp32                              5 bfp32_ofp32_over_column
//  6 nobias_ofp32                             7 nobias_ofp32_batch
void ref_int8_matmul(int variant, const int8_t *A, const int8_t *B, const int8_t *bias8, const float *biasf,
                     int8_t *C8, float *Cf, int M, int N, int K, float alpha, float beta, int q_min, int q_max,
                     int num_thread) {
    struct matmul_params p;
    memset(&p, 0, sizeof(p));
    p.A.row = M;
    p.A.column = K;
    p.A.int8_data_ptr = const_cast<int8_t *>(A);
    p.A.qparams.scale = alpha;
    p.A.qparams.zero_point = 0;
    p.B.row = K;
    p.B.column = N;
    p.B.int8_data_ptr = const_cast<int8_t *>(B);
    p.B.qparams.scale = 1.0f;
    p.B.qparams.zero_point = 0;
    p.C.row = M;
    p.C.column = N;
    p.C.int8_data_ptr = C8;
    p.C.data_ptr = Cf;
    p.C.qparams.scale = 1.0f;
    p.C.qparams.zero_point = 0;
    p.C.qparams.q_min = (int8_t)q_min;
    p.C.qparams.q_max = (int8_t)q_max;
    p.bias.row = 1;
    p.bias.column = N;
    p.bias.int8_data_ptr = const_cast<int8_t *>(bias8);
    p.bias.data_ptr = const_cast<float *>(biasf);
    p.alpha = alpha;
    p.beta = beta;
    p.opt_params.blk_size = 256;
    p.opt_params.num_thread = num_thread;
    MatmulOperator op;
    switch (variant) {
        case 0: op.mat_mul_accelerator_int8_fast_2x2_32unroll(&p); break;
        case 1: op.mat_mul_accelerator_int8_fast_32unroll_over_column(&p); break;
        case 2: op.mat_mul_accelerator_int8_fast_2x2_32unroll_nobias(&p); break;
        case 3: op.mat_mul_accelerator_int8_fast_2x2_32unroll_nobias_batch(&p); break;
        case 4: op.mat_mul_accelerator_int8_fast_2x2_32unroll_bfp32_ofp32(&p); break;
        case 5: op.mat_mul_accelerator_int8_fast_2x2_32unroll_bfp32_ofp32_over_column(&p); break;
        case 6: op.mat_mul_accelerator_int8_fast_2x2_32unroll_nobias_ofp32(&p); break;
        case 7: op.mat_mul_accelerator_int8_fast_2x2_32unroll_nobias_ofp32_batch(&p); break;
        default: break;
    }
}

// kernels/matmul_int8.cc:8-30
void ref_naive_mat_mul_int8(const int8_t *A, const int8_t *B, int8_t *C, int M, int N, int K, int A_zp, int C_zp,
                            float A_sc, float B_sc, float C_sc, int q_min, int q_max) {
    struct matmul_params p;
    memset(&p, 0, sizeof(p));
    p.A.row = M;
    p.A.column = K;
    p.A.int8_data_ptr = const_cast<int8_t *>(A);
    p.A.qparams.zero_point = A_zp;
    p.A.qparams.scale = A_sc;
    p.B.row = K;
    p.B.column = N;
    p.B.int8_data_ptr = const_cast<int8_t *>(B);
    p.B.qparams.scale = B_sc;
    p.C.row = M;
    p.C.column = N;
    p.C.int8_data_ptr = C;
    p.C.qparams.zero_point = C_zp;
    p.C.qparams.scale = C_sc;
    p.C.qparams.q_min = (int8_t)q_min;
    p.C.qparams.q_max = (int8_t)q_max;
    MatmulOperator op;
    op.naive_mat_mul_int8(&p);
}

// kernels/matmul_imp.cc:23-35
void ref_mat_mul_transposed(const float *A, const float *B, float *C, int M, int N, int K) {
    struct matmul_params p;
    memset(&p, 0, sizeof(p));
    p.A.row = M;
    p.A.column = K;
    p.A.data_ptr = const_cast<float *>(A);
    p.B.row = N;
    p.B.column = K;
    p.B.data_ptr = const_cast<float *>(B);
    p.C.row = M;
    p.C.column = N;
    p.C.data_ptr = C;
    MatmulOperator op;
    op.mat_mul_transposed(&p);
}

#ifdef REF_HAS_FP16_INT4
// kernels/cuda/matmul_int4.cu:8-48 compiled as host C++; call site Linear_FP16_int4_ref::forward_ref
// (llm/src/ops/cuda/linear.cu:43-76): B int32[IC][OC/8], B.row=IC, B.column=OC/8
void ref_naive_mat_mul_fp16_int4(const uint16_t *A, const int32_t *B, const uint16_t *scales, uint16_t *C, int M,
                                 int IC, int OC, int block_size) {
    struct matmul_params p;
    memset(&p, 0, sizeof(p));
    p.A.row = M;
    p.A.column = IC;
    p.A.fp16_data_ptr = reinterpret_cast<naive_float16_t *>(const_cast<uint16_t *>(A));
    p.B.row = IC;
    p.B.column = OC / 8;
    p.B.int32_data_ptr = const_cast<int32_t *>(B);
    p.C.row = M;
    p.C.column = OC;
    p.C.fp16_data_ptr = reinterpret_cast<naive_float16_t *>(C);
    p.fp16_scales = reinterpret_cast<naive_float16_t *>(const_cast<uint16_t *>(scales));
    p.block_size = block_size;
    MatmulOperator op;
    op.naive_mat_mul_fp16_int4(&p);
}
#endif

#ifdef QM_x86
// The reference's CPU hot path (the timed baseline): Linear_FP_int4::forward, llm/src/ops/linear.cc:171-236
// -> mat_mul_accelerator_int8_int4_fast_no_offset (kernels/avx/matmul_avx_int8_int4.cc:325-357).
// QM_x86 weight format (quantize_row_q4_3), block 32, B.row = IC/2, B.column = OC; x_int8/x_scale are the
// caller-owned scratch the reference keeps in file statics (linear.cc:158-168).  32-byte aligned buffers.
void ref_w4a8_avx(float *A, uint8_t *B, float *scales, float *C, int8_t *x_int8, float *x_scale, int M, int IC,
                  int OC, int num_thread) {
    struct matmul_params p;
    memset(&p, 0, sizeof(p));
    p.A.row = M;
    p.A.column = IC;
    p.A.data_ptr = A;
    p.A.int8_data_ptr = x_int8;
    p.A_scales = x_scale;
    p.B.row = IC / 2;
    p.B.column = OC;
    p.B.int4_data_ptr = B;
    p.C.row = M;
    p.C.column = OC;
    p.C.data_ptr = C;
    p.opt_params.num_thread = num_thread;
    p.scales = scales;
    p.offset = nullptr;
    p.block_size = 32;
    p.bias.data_ptr = nullptr;
    MatmulOperator op;
    op.mat_mul_accelerator_int8_int4_fast_no_offset(&p);
}
#endif

}  // extern "C"
