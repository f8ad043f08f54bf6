// ops.h -- the reference's L2 op classes on the hot path (boundary B), same names / forward()
// signatures, device buffers owned by the caller as in the reference (llm/include/ops/*.h).  Thin: each forward()
// fills a matmul_params the way the reference wrapper does and calls the MatmulOperator adapter.
#ifndef TCE_HOST_OPS_H
#define TCE_HOST_OPS_H
#include <assert.h>

#include <string>

#include "matmul.h"

#define QK 128  // llm/include/common.h:17-21 under QM_CUDA

template <typename T>
class Matrix3D {  // llm/include/common.h:33-120 (shape + raw pointer; no ownership)
   public:
    Matrix3D(T *data, int dim_x, int dim_y, int dim_z) : m_data(data), m_dim_x(dim_x), m_dim_y(dim_y), m_dim_z(dim_z) {}
    Matrix3D() : m_data(nullptr), m_dim_x(0), m_dim_y(0), m_dim_z(0) {}
    int length() const { return m_dim_x * m_dim_y * m_dim_z; }
    T *m_data;
    int m_dim_x, m_dim_y, m_dim_z;
};

// llm/include/ops/linear.h:186-221, forward: llm/src/ops/cuda/linear.cu:5-40
class Linear_half_int4 {
   public:
    Linear_half_int4(Matrix3D<int> weight_, Matrix3D<float16_t> scale_, Matrix3D<int> zero_point_) : weight(weight_), scale(scale_), zero_point(zero_point_) {}
    Linear_half_int4() {}
    void forward(const Matrix3D<float16_t> &x, Matrix3D<float16_t> &output);
    Matrix3D<int> weight;        // (1, OC, IC/8)
    Matrix3D<float16_t> scale;   // (1, OC, zeros_w*8)
    Matrix3D<int> zero_point;    // (1, OC, zeros_w)
};

struct W8A8B8O8Linear_params { Matrix3D<int8_t> weight; Matrix3D<int8_t> bias; float alpha; float beta; };
class W8A8B8O8Linear {  // llm/src/ops/W8A8B8O8Linear.cc:15-78
   public:
    W8A8B8O8Linear(W8A8B8O8Linear_params &p, int8_t q_min = -128);
    W8A8B8O8Linear() {}
    void forward(const Matrix3D<int8_t> &x, Matrix3D<int8_t> &output);
    struct matmul_params params;
    float alpha, beta;
};
class W8A8B8O8LinearReLU : public W8A8B8O8Linear {  // llm/src/ops/W8A8B8O8LinearReLU.cc (q_min = 0)
   public:
    W8A8B8O8LinearReLU(W8A8B8O8Linear_params &p) : W8A8B8O8Linear(p, 0) {}
    W8A8B8O8LinearReLU() {}
};

struct W8A8BFP32OFP32Linear_params { Matrix3D<int8_t> weight; Matrix3D<float> bias; float alpha; };
class W8A8BFP32OFP32Linear {  // llm/src/ops/W8A8BFP32OFP32Linear.cc:14-81
   public:
    W8A8BFP32OFP32Linear(W8A8BFP32OFP32Linear_params &p);
    W8A8BFP32OFP32Linear() {}
    void forward(const Matrix3D<int8_t> &x, Matrix3D<float> &output);
    struct matmul_params params;
    float alpha;
};

class BMM_S8T_S8N_F32T {  // llm/src/ops/BMM_S8T_S8N_F32T.cc:12-62
   public:
    explicit BMM_S8T_S8N_F32T(float alpha_) : alpha(alpha_) {}
    BMM_S8T_S8N_F32T() : alpha(1.f) {}
    void forward(const Matrix3D<int8_t> &x, const Matrix3D<int8_t> &weight, Matrix3D<float> &output);
    float alpha;
};
class BMM_S8T_S8N_S8T {  // llm/src/ops/BMM_S8T_S8N_S8T.cc:12-64
   public:
    explicit BMM_S8T_S8N_S8T(float alpha_) : alpha(alpha_) {}
    BMM_S8T_S8N_S8T() : alpha(1.f) {}
    void forward(const Matrix3D<int8_t> &x, const Matrix3D<int8_t> &weight, Matrix3D<int8_t> &output);
    float alpha;
};

// Parameter loading with the reference's names, file names and error behaviour (`throw const char *` on I/O failure, llm/src/utils.cc:16-25):
// llm/src/ops/W8A8B8O8Linear.cc:6-13, W8A8BFP32OFP32Linear.cc:6-10, BMM_S8T_S8N_F32T.cc:6-8, BMM_S8T_S8N_S8T.cc:6-8.  The op's buffers may be host,
// managed or device memory (the bytes travel with cudaMemcpyDefault).
void load_W8A8B8O8Linear_params(W8A8B8O8Linear &op, std::string prefix);        // weight.bin, bias_int8.bin, alpha.bin, beta.bin
void load_W8A8BFP32OFP32Linear_params(W8A8BFP32OFP32Linear &op, std::string prefix);  // weight.bin, bias.bin, alpha.bin
void load_BMM_S8T_S8N_F32T(BMM_S8T_S8N_F32T &op, std::string prefix);           // alpha.bin
void load_BMM_S8T_S8N_S8T(BMM_S8T_S8N_S8T &op, std::string prefix);             // alpha.bin

#endif
