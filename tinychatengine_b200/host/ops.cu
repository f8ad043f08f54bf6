// ops.cu -- forward() bodies of the L2 op classes: parameter blocks filled as the reference wrappers fill them.
#include <cuda_runtime.h>
#include <errno.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "ops.h"

// read `bytes` of a parameter file into `dst` (host, managed or device memory).  Failure behaviour of the reference's read_to_array
// (llm/src/utils.cc:16-25): print the reason and throw a C string.
static void read_param_file(const std::string &path, void *dst, size_t bytes, bool host_scalar = false) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
        printf("%s: %s\n", strerror(errno), path.c_str());
        throw("Expected error...");
    }
    std::vector<char> host(bytes);
    const size_t got = fread(host.data(), 1, bytes, f);
    fclose(f);
    if (got != bytes) {
        printf("short read (%zu of %zu bytes): %s\n", got, bytes, path.c_str());
        throw("Expected error...");
    }
    if (host_scalar) {  // alpha / beta: members of the op object
        memcpy(dst, host.data(), bytes);
        return;
    }
    if (cudaMemcpy(dst, host.data(), bytes, cudaMemcpyDefault) != cudaSuccess) {
        printf("cudaMemcpy of %zu bytes failed: %s\n", bytes, path.c_str());
        throw("Expected error...");
    }
}

void load_W8A8B8O8Linear_params(W8A8B8O8Linear &op, std::string prefix) {
    read_param_file(prefix + "/weight.bin", op.params.B.int8_data_ptr, (size_t)op.params.B.length());
    read_param_file(prefix + "/bias_int8.bin", op.params.bias.int8_data_ptr, (size_t)op.params.bias.length());
    read_param_file(prefix + "/alpha.bin", &op.alpha, sizeof(float), true);
    read_param_file(prefix + "/beta.bin", &op.beta, sizeof(float), true);
    op.params.alpha = op.alpha;
    op.params.beta = op.beta;
    op.params.A.qparams.scale = op.alpha;
}

void load_W8A8BFP32OFP32Linear_params(W8A8BFP32OFP32Linear &op, std::string prefix) {
    read_param_file(prefix + "/weight.bin", op.params.B.int8_data_ptr, (size_t)op.params.B.length());
    read_param_file(prefix + "/bias.bin", op.params.bias.data_ptr, (size_t)op.params.bias.length() * sizeof(float));
    read_param_file(prefix + "/alpha.bin", &op.alpha, sizeof(float), true);
}

void load_BMM_S8T_S8N_F32T(BMM_S8T_S8N_F32T &op, std::string prefix) { read_param_file(prefix + "/alpha.bin", &op.alpha, sizeof(float), true); }

void load_BMM_S8T_S8N_S8T(BMM_S8T_S8N_S8T &op, std::string prefix) { read_param_file(prefix + "/alpha.bin", &op.alpha, sizeof(float), true); }

void Linear_half_int4::forward(const Matrix3D<float16_t> &x, Matrix3D<float16_t> &output) {
    assert(output.m_dim_x == 1);
    assert(output.m_dim_y == x.m_dim_y);
    struct matmul_params params;
    memset(&params, 0, sizeof(params));
    params.A.row = x.m_dim_y;
    params.A.column = x.m_dim_z;
    params.A.half_data_ptr = x.m_data;
    params.B.row = weight.m_dim_z;     // k / 8
    params.B.column = weight.m_dim_y;  // n
    params.B.int32_data_ptr = weight.m_data;
    params.C.row = output.m_dim_y;
    params.C.column = output.m_dim_z;
    params.C.half_data_ptr = output.m_data;
    params.opt_params.num_thread = 8;
    params.half_scales = scale.m_data;
    params.int32_zero_point = zero_point.m_data;
    params.block_size = QK;
    matmul::MatmulOperator op;
    op.gemv_forward_cuda(&params);
}

W8A8B8O8Linear::W8A8B8O8Linear(W8A8B8O8Linear_params &op, int8_t q_min) {
    memset(&params, 0, sizeof(params));
    alpha = op.alpha;
    beta = op.beta;
    const int k = op.weight.m_dim_z, n = op.weight.m_dim_y;
    params.A.qparams.scale = alpha;
    params.B.qparams.scale = 1.0f;
    params.C.qparams.scale = 1.0f;
    params.B.row = k;
    params.B.column = n;
    params.B.int8_data_ptr = op.weight.m_data;
    params.C.qparams.q_max = 127;
    params.C.qparams.q_min = q_min;
    params.bias.int8_data_ptr = op.bias.m_data;
    params.bias.row = 1;
    params.bias.column = n;
}

void W8A8B8O8Linear::forward(const Matrix3D<int8_t> &x, Matrix3D<int8_t> &output) {
    const int m = x.m_dim_y, k = x.m_dim_z, n = params.B.column;
    assert(output.m_dim_x == x.m_dim_x && output.m_dim_y == x.m_dim_y && output.m_dim_z == n && x.m_dim_z == params.B.row);
    params.A.row = m;
    params.A.column = k;
    params.A.int8_data_ptr = x.m_data;
    params.C.row = m;
    params.C.column = n;
    params.C.int8_data_ptr = output.m_data;
    params.alpha = alpha;
    params.beta = beta;
    matmul::MatmulOperator op;
    for (int bz = 0; bz < x.m_dim_x; bz++) {
        if (m == 1)
            op.mat_mul_accelerator_int8_fast_32unroll_over_column(&params);
        else
            op.mat_mul_accelerator_int8_fast_2x2_32unroll(&params);
        params.A.int8_data_ptr += m * k;
        params.C.int8_data_ptr += m * n;
    }
}

W8A8BFP32OFP32Linear::W8A8BFP32OFP32Linear(W8A8BFP32OFP32Linear_params &op) {
    memset(&params, 0, sizeof(params));
    alpha = op.alpha;
    const int k = op.weight.m_dim_z, n = op.weight.m_dim_y;
    params.B.row = k;
    params.B.column = n;
    params.B.int8_data_ptr = op.weight.m_data;
    params.C.column = n;
    params.bias.data_ptr = op.bias.m_data;
    params.bias.row = 1;
    params.bias.column = op.bias.m_dim_z;
}

void W8A8BFP32OFP32Linear::forward(const Matrix3D<int8_t> &x, Matrix3D<float> &output) {
    const int m = x.m_dim_y, k = x.m_dim_z, n = params.B.column;
    assert(output.m_dim_z == n && x.m_dim_z == params.B.row);
    params.A.row = m;
    params.A.column = k;
    params.A.int8_data_ptr = x.m_data;
    params.C.row = m;
    params.C.column = n;
    params.C.data_ptr = output.m_data;
    params.alpha = alpha;
    matmul::MatmulOperator op;
    for (int bz = 0; bz < x.m_dim_x; bz++) {
        if (m == 1)
            op.mat_mul_accelerator_int8_fast_2x2_32unroll_bfp32_ofp32_over_column(&params);
        else
            op.mat_mul_accelerator_int8_fast_2x2_32unroll_bfp32_ofp32(&params);
        params.A.int8_data_ptr += m * k;
        params.C.data_ptr += m * n;
    }
}

static void bmm_params(struct matmul_params &params, const Matrix3D<int8_t> &x, const Matrix3D<int8_t> &weight, float alpha) {
    memset(&params, 0, sizeof(params));
    const int m = x.m_dim_y, k = x.m_dim_z, n = weight.m_dim_y;
    params.A.row = m;
    params.A.column = k;
    params.A.int8_data_ptr = x.m_data;
    params.B.row = k;
    params.B.column = n;
    params.B.int8_data_ptr = weight.m_data;
    params.C.row = m;
    params.C.column = n;
    params.C.qparams.q_max = 127;
    params.C.qparams.q_min = -128;
    params.alpha = alpha;
}

void BMM_S8T_S8N_F32T::forward(const Matrix3D<int8_t> &x, const Matrix3D<int8_t> &weight, Matrix3D<float> &output) {
    const int m = x.m_dim_y, k = x.m_dim_z, n = weight.m_dim_y;
    assert(output.m_dim_x == x.m_dim_x && output.m_dim_y == m && output.m_dim_z == n && k == weight.m_dim_z);
    struct matmul_params params;
    bmm_params(params, x, weight, alpha);
    params.C.data_ptr = output.m_data;
    matmul::MatmulOperator op;
    if (m == 1 && x.m_dim_x > 1) {
        params.A.row = x.m_dim_x;
        params.C.row = x.m_dim_x;
        op.mat_mul_accelerator_int8_fast_2x2_32unroll_nobias_ofp32_batch(&params);
    } else {
        for (int bz = 0; bz < x.m_dim_x; bz++) {
            op.mat_mul_accelerator_int8_fast_2x2_32unroll_nobias_ofp32(&params);
            params.A.int8_data_ptr += m * k;
            params.B.int8_data_ptr += k * n;
            params.C.data_ptr += m * n;
        }
    }
}

void BMM_S8T_S8N_S8T::forward(const Matrix3D<int8_t> &x, const Matrix3D<int8_t> &weight, Matrix3D<int8_t> &output) {
    const int m = x.m_dim_y, k = x.m_dim_z, n = weight.m_dim_y;
    assert(output.m_dim_x == x.m_dim_x && output.m_dim_y == m && output.m_dim_z == n && k == weight.m_dim_z);
    struct matmul_params params;
    bmm_params(params, x, weight, alpha);
    params.C.int8_data_ptr = output.m_data;
    matmul::MatmulOperator op;
    if (m == 1 && x.m_dim_x > 1) {
        params.A.row = x.m_dim_x;
        params.C.row = x.m_dim_x;
        op.mat_mul_accelerator_int8_fast_2x2_32unroll_nobias_batch(&params);
    } else {
        for (int bz = 0; bz < x.m_dim_x; bz++) {
            op.mat_mul_accelerator_int8_fast_2x2_32unroll_nobias(&params);
            params.A.int8_data_ptr += m * k;
            params.B.int8_data_ptr += k * n;
            params.C.int8_data_ptr += m * n;
        }
    }
}
