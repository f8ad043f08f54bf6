// matmul_operator.cu -- MatmulOperator methods of the CUDA backend, forwarding to the C ABI (include/tce_b200.h).
// Field conventions per op are those of the reference call sites (SURVEY.md 8(b) "Field conventions per op").
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime.h>

#include "../../include/tce_b200.h"
// Built against this package's mirror of the operator header by default; the drop-in proof (oracle/Makefile, target callsites) builds
// the same file against the reference's own kernels/matmul.h (-DTCE_REFERENCE_MATMUL_H='"<path>"' -DQM_CUDA) so that the reference's
// unchanged call sites and these definitions share ONE declaration of matmul::MatmulOperator.
#ifdef TCE_REFERENCE_MATMUL_H
#include TCE_REFERENCE_MATMUL_H
struct tce_ctx;
extern "C" tce_ctx *tce_host_ctx(void);
extern "C" void tce_host_set_stream(void *cuda_stream);
#else
#include "matmul.h"
#endif

static tce_ctx *g_ctx = nullptr;

extern "C" tce_ctx *tce_host_ctx(void) {
    if (!g_ctx) {
        if (tce_ctx_create(0, &g_ctx) != TCE_OK) {
            fprintf(stderr, "libtce_b200: %s\n", tce_last_error());
            abort();  // the reference's CHECK_CUDA throws; there is no CPU fallback to continue on
        }
    }
    return g_ctx;
}
extern "C" void tce_host_set_stream(void *s) { tce_ctx_set_stream(tce_host_ctx(), s); }

// ---- operands that live in plain host memory ------------------------------------------------------------------------------------
// The reference's CUDA build keeps the W8A8 operators' buffers in posix_memalign host memory (llm/src/utils.cc:205-220; only the W4
// path uses cudaMallocManaged, cuda/utils.cu:93-96) and its "CUDA" int8 kernels are host loops (kernels/cuda/matmul_ref_int8.cc).  A
// drop-in therefore has to accept host pointers: operands the device cannot reach are staged through grow-only device scratch (one
// slot per operand role) and results are copied back before the call returns, which also preserves the synchronous contract of those
// operators.  Device-reachable pointers (cudaMalloc / cudaMallocManaged / pinned) are passed through untouched.
namespace {
struct Scratch {
    void *dev = nullptr;
    size_t cap = 0;
};
Scratch g_scratch[4];  // 0 A, 1 B, 2 bias, 3 C

bool device_reachable(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged || (at.type == cudaMemoryTypeHost && at.devicePointer != nullptr);
}
void *scratch(int slot, size_t bytes) {
    Scratch &s = g_scratch[slot];
    if (s.cap < bytes) {
        if (s.dev) cudaFree(s.dev);
        if (cudaMalloc(&s.dev, bytes) != cudaSuccess) {
            fprintf(stderr, "libtce_b200: staging allocation of %zu bytes failed\n", bytes);
            abort();
        }
        s.cap = bytes;
    }
    return s.dev;
}
// input operand: device-reachable pointer to the same bytes
const void *in_dev(const void *p, size_t bytes, int slot) {
    if (!p || device_reachable(p)) return p;
    void *d = scratch(slot, bytes);
    if (cudaMemcpy(d, p, bytes, cudaMemcpyHostToDevice) != cudaSuccess) abort();
    return d;
}
struct OutStage {
    void *host = nullptr, *dev = nullptr;
    size_t bytes = 0;
    void *begin(void *p, size_t n, int slot) {
        if (device_reachable(p)) return p;
        host = p;
        bytes = n;
        dev = scratch(slot, n);
        return dev;
    }
    void finish() {  // results in the caller's buffer when the call returns, like the host loops this replaces
        if (host && cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) abort();
    }
};
}  // namespace

static void must(int rc, const char *what) {
    if (rc != TCE_OK) {
        fprintf(stderr, "libtce_b200: %s failed: %s\n", what, tce_last_error());
        exit(1);  // reference behaviour on an unsupported configuration (gemv_cuda.cu:253-257)
    }
}

namespace matmul {

// reference: kernels/cuda/gemv_cuda.cu:213-260.  A.row=M, A.column=IC, B.row=IC/8, B.column=OC, C.row=M, C.column=OC
void MatmulOperator::gemv_forward_cuda(const struct matmul_params *p) {
    const int M = p->A.row, IC = p->A.column, OC = p->C.column;
    assert(p->C.row == M);
    // the reference ignores params->block_size and uses the compile-time QK (=128 under QM_CUDA), gemv_cuda.cu:221
    must(tce_w4a16_gemv(tce_host_ctx(), p->A.half_data_ptr, p->B.int32_data_ptr, p->int32_zero_point, p->half_scales, p->C.half_data_ptr, M, IC, OC, 128),
         "gemv_forward_cuda");
}

// the prefill slot the reference declares but never defines (kernels/matmul.h:142-145); same operand convention
void MatmulOperator::gemm_forward_cuda(const struct matmul_params *p, int /*split_k_iters*/) {
    must(tce_w4a16_gemm(tce_host_ctx(), p->A.half_data_ptr, p->B.int32_data_ptr, p->int32_zero_point, p->half_scales, p->C.half_data_ptr, p->A.row,
                        p->A.column, p->C.column, 128),
         "gemm_forward_cuda");
}
void MatmulOperator::gemm_forward_cuda_8splits(const struct matmul_params *p, float16_t *) { gemm_forward_cuda(p, 8); }
void MatmulOperator::gemm_forward_cuda_half(const struct matmul_params *p, int s) { gemm_forward_cuda(p, s); }
void MatmulOperator::gemm_forward_cuda_half_test(const struct matmul_params *p, int s) { gemm_forward_cuda(p, s); }

// reference: kernels/cuda/matmul_int4.cu:8-48 (host fp16 reference; here on the device, bit-identical).  B.row=IC, B.column=OC/8
void MatmulOperator::naive_mat_mul_fp16_int4(const struct matmul_params *p) {
    must(tce_naive_fp16_int4(tce_host_ctx(), p->A.fp16_data_ptr, p->B.int32_data_ptr, p->fp16_scales, p->C.fp16_data_ptr, p->C.row, p->B.row,
                             p->C.column, p->block_size),
         "naive_mat_mul_fp16_int4");
}

// reference: kernels/cuda/matmul_ref_fp32.cc:11-34.  A.row=M, A.column=K, B stored [N][K] with B.row=K? (Linear_FP: B.row=k, B.column=n)
void MatmulOperator::mat_mul_accelerator_transposed_fastover_column(const struct matmul_params *p) {
    must(tce_f32_matmul_transposed(tce_host_ctx(), p->A.data_ptr, p->B.data_ptr, p->C.data_ptr, p->A.row, p->B.column, p->A.column),
         "mat_mul_accelerator_transposed_fastover_column");
}

// stubs, exactly like the reference CUDA build (gemv_cuda.cu:262-268)
void MatmulOperator::mat_mul_accelerator_int4_fast(const struct matmul_params *) {}
void MatmulOperator::mat_mul_accelerator_int4_fast_no_offset(const struct matmul_params *) {}

// INT8 family (reference kernels/cuda/matmul_ref_int8.cc == kernels/ref/matmul_ref_int8.cc): B.row=K, B.column=N
static void w8(const struct matmul_params *p, int variant, int batch, const void *bias, void *C, const char *who) {
    const int M = p->A.row, K = p->A.column, N = p->B.column;
    assert(p->A.column == p->B.row && p->C.row == M && p->C.column == N);
    const size_t cel = (variant < 2) ? 1 : 4, bel = (variant == 0) ? 1 : 4;
    const void *A = in_dev(p->A.int8_data_ptr, (size_t)M * K, 0);
    const void *B = in_dev(p->B.int8_data_ptr, (size_t)(batch ? M : 1) * N * K, 1);
    const void *bs = bias ? in_dev(bias, (size_t)N * bel, 2) : nullptr;
    OutStage out;
    void *Cd = out.begin(C, (size_t)M * N * cel, 3);
    must(tce_w8a8_matmul(tce_host_ctx(), variant, batch, A, B, bs, Cd, M, N, K, p->alpha, p->beta, p->C.qparams.q_min, p->C.qparams.q_max), who);
    out.finish();
}
void MatmulOperator::mat_mul_accelerator_int8_fast_2x2_32unroll(const struct matmul_params *p) {
    w8(p, 0, 0, p->bias.int8_data_ptr, p->C.int8_data_ptr, "int8_fast_2x2_32unroll");
}
void MatmulOperator::mat_mul_accelerator_int8_fast_32unroll_over_column(const struct matmul_params *p) {
    w8(p, 0, 0, p->bias.int8_data_ptr, p->C.int8_data_ptr, "int8_fast_32unroll_over_column");
}
void MatmulOperator::mat_mul_accelerator_int8_fast_2x2_32unroll_nobias(const struct matmul_params *p) {
    w8(p, 1, 0, nullptr, p->C.int8_data_ptr, "int8_fast_2x2_32unroll_nobias");
}
void MatmulOperator::mat_mul_accelerator_int8_fast_2x2_32unroll_nobias_batch(const struct matmul_params *p) {
    w8(p, 1, 1, nullptr, p->C.int8_data_ptr, "int8_fast_2x2_32unroll_nobias_batch");
}
void MatmulOperator::mat_mul_accelerator_int8_fast_2x2_32unroll_bfp32_ofp32(const struct matmul_params *p) {
    w8(p, 2, 0, p->bias.data_ptr, p->C.data_ptr, "int8_fast_2x2_32unroll_bfp32_ofp32");
}
void MatmulOperator::mat_mul_accelerator_int8_fast_2x2_32unroll_bfp32_ofp32_over_column(const struct matmul_params *p) {
    w8(p, 2, 0, p->bias.data_ptr, p->C.data_ptr, "int8_fast_2x2_32unroll_bfp32_ofp32_over_column");
}
void MatmulOperator::mat_mul_accelerator_int8_fast_2x2_32unroll_nobias_ofp32(const struct matmul_params *p) {
    w8(p, 3, 0, nullptr, p->C.data_ptr, "int8_fast_2x2_32unroll_nobias_ofp32");
}
void MatmulOperator::mat_mul_accelerator_int8_fast_2x2_32unroll_nobias_ofp32_batch(const struct matmul_params *p) {
    w8(p, 3, 1, nullptr, p->C.data_ptr, "int8_fast_2x2_32unroll_nobias_ofp32_batch");
}

}  // namespace matmul
