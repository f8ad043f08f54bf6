// ref_modules_shim.cc -- extern "C" doorway into the UNMODIFIED reference module Int8OPTAttention (compiled in place from
// /root/reference/llm/src by oracle/Makefile into oracle/_ref/libtce_ref_modules.so).  TEST INFRASTRUCTURE ONLY.
//
// Nothing here re-implements arithmetic: the shim allocates the operator buffers the way the reference's own model code does
// (llm/src/nn_modules/Int8OPTDecoderLayer.cc:130-170), lets the reference constructor load a synthetic parameter tree from
// `param_path` (written by tests/golden/make_golden.py: q_proj|k_proj|v_proj/{weight,bias_int8,alpha,beta}.bin,
// out_proj/{weight,bias,alpha}.bin, qk_bmm|pv_bmm/alpha.bin), and runs Int8OPTAttention::forward for a prefill followed by
// single-token steps fed with the returned past_key_value.  It pins oracle/tce_oracle.c's restatement of the module.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "nn_modules/Int8OPTDecoderLayer.h"  // pulls in Int8OPTAttention.h (the reference headers carry no include guards)
#include "operators.h"
#include "utils.h"

int NUM_THREAD = 4;  // the reference applications define this global (llm/application/chat.cc)

extern "C" {

// hidden: int8 [total_tokens][E] (prefill rows first, then one row per decode step); masks built here as the reference's
// Int8OPTDecoder::prepare_decoder_attention_mask does (0 / lowest float above the diagonal).
// out_fp: float [total_tokens][E] attention outputs in call order; final_k/final_v: int8 [H][total_tokens][hd] after the last call.
int ref_int8_opt_attention(const char *param_path, int E, int H, int max_sqlen, const int8_t *hidden, int prefill, int decode_steps, float *out_fp,
                           int8_t *final_k, int8_t *final_v) {
    struct model_config cfg(1, H, 1, max_sqlen, E, 4 * E, 50272, 1, 0);
    Int8OPTAttention::initialized_memory(cfg);
    const int hd = E / H;
    std::vector<int8_t> wq((size_t)E * E), wk((size_t)E * E), wv((size_t)E * E), wo((size_t)E * E), bq(E), bk(E), bv(E);
    std::vector<float> bo(E);
    struct W8A8B8O8Linear_params pq, pk, pv;
    pq.weight = Matrix3D<int8_t>(wq.data(), 1, E, E);
    pq.bias = Matrix3D<int8_t>(bq.data(), 1, 1, E);
    pk.weight = Matrix3D<int8_t>(wk.data(), 1, E, E);
    pk.bias = Matrix3D<int8_t>(bk.data(), 1, 1, E);
    pv.weight = Matrix3D<int8_t>(wv.data(), 1, E, E);
    pv.bias = Matrix3D<int8_t>(bv.data(), 1, 1, E);
    struct W8A8BFP32OFP32Linear_params po;
    po.weight = Matrix3D<int8_t>(wo.data(), 1, E, E);
    po.bias = Matrix3D<float>(bo.data(), 1, 1, E);
    W8A8B8O8Linear q_proj(pq), k_proj(pk), v_proj(pv);
    W8A8BFP32OFP32Linear out_proj(po);
    BMM_S8T_S8N_F32T qk_bmm;
    BMM_S8T_S8N_S8T pv_bmm;
    Int8OPTAttention attn(std::string(param_path), cfg, qk_bmm, pv_bmm, k_proj, v_proj, q_proj, out_proj);

    Matrix3D<int8_t> past_k, past_v;
    int past = 0, row = 0;
    for (int call = 0; call < 1 + decode_steps; call++) {
        const int sqlen = call == 0 ? prefill : 1, tgz = past + sqlen;
        std::vector<float> mask((size_t)sqlen * tgz, 0.f);
        for (int i = 0; i < sqlen; i++)
            for (int j = past + i + 1; j < tgz; j++) mask[(size_t)i * tgz + j] = -3.402823466e38f;
        Matrix3D<int8_t> hs(const_cast<int8_t *>(hidden) + (size_t)row * E, 1, sqlen, E);
        Matrix3D<float> m(mask.data(), 1, sqlen, tgz);
        struct Int8OPTAttention_output out =
            call == 0 ? attn.forward(Int8OPTAttention_input(hs, m, 0)) : attn.forward(Int8OPTAttention_input(hs, m, past_k, past_v, true, 0));
        memcpy(out_fp + (size_t)row * E, out.attn_output.m_data, (size_t)sqlen * E * sizeof(float));
        past_k = out.past_key_value.first;
        past_v = out.past_key_value.second;
        past = tgz;
        row += sqlen;
    }
    memcpy(final_k, past_k.m_data, (size_t)H * past * hd);
    memcpy(final_v, past_v.m_data, (size_t)H * past * hd);
    return past;
}
}

// Int8OPTDecoderLayer::forward (llm/src/nn_modules/Int8OPTDecoderLayer.cc:24-59), unmodified: LayerNormQ -> Int8OPTAttention -> residual ->
// LayerNormQ -> fc1 (W8A8B8O8LinearReLU) -> fc2 (W8A8BFP32OFP32Linear) -> residual.  Parameter tree under `param_path`: self_attn/...
// (as above), self_attn_layer_norm|final_layer_norm/{weight,bias}.bin, fc1/{weight,bias_int8,alpha,beta}.bin, fc2/{weight,bias,alpha}.bin.
// hidden: fp32 [total_tokens][E]; out_fp: fp32 [total_tokens][E] layer outputs in call order.
extern "C" int ref_int8_opt_decoder_layer(const char *param_path, int E, int H, int F, int max_sqlen, const float *hidden, int prefill, int decode_steps,
                                          float *out_fp, int8_t *final_k, int8_t *final_v) {
    struct model_config cfg(1, H, 1, max_sqlen, E, F, 50272, 1, 0);
    const int hd = E / H;
    std::vector<int8_t> wq((size_t)E * E), wk((size_t)E * E), wv((size_t)E * E), wo((size_t)E * E), bq(E), bk(E), bv(E), w1((size_t)F * E), b1(F),
        w2((size_t)E * F);
    std::vector<float> bo(E), b2(E), ln1w(E), ln1b(E), ln2w(E), ln2b(E);
    struct W8A8B8O8Linear_params pq, pk, pv;
    pq.weight = Matrix3D<int8_t>(wq.data(), 1, E, E);
    pq.bias = Matrix3D<int8_t>(bq.data(), 1, 1, E);
    pk.weight = Matrix3D<int8_t>(wk.data(), 1, E, E);
    pk.bias = Matrix3D<int8_t>(bk.data(), 1, 1, E);
    pv.weight = Matrix3D<int8_t>(wv.data(), 1, E, E);
    pv.bias = Matrix3D<int8_t>(bv.data(), 1, 1, E);
    struct W8A8BFP32OFP32Linear_params po, p2;
    po.weight = Matrix3D<int8_t>(wo.data(), 1, E, E);
    po.bias = Matrix3D<float>(bo.data(), 1, 1, E);
    p2.weight = Matrix3D<int8_t>(w2.data(), 1, E, F);
    p2.bias = Matrix3D<float>(b2.data(), 1, 1, E);
    struct W8A8B8O8LinearReLU_params p1;
    p1.weight = Matrix3D<int8_t>(w1.data(), 1, F, E);
    p1.bias_int8 = Matrix3D<int8_t>(b1.data(), 1, 1, F);
    struct LayerNormQ_params l1, l2;
    l1.weight = Matrix3D<float>(ln1w.data(), 1, 1, E);
    l1.bias = Matrix3D<float>(ln1b.data(), 1, 1, E);
    l2.weight = Matrix3D<float>(ln2w.data(), 1, 1, E);
    l2.bias = Matrix3D<float>(ln2b.data(), 1, 1, E);
    W8A8B8O8Linear q_proj(pq), k_proj(pk), v_proj(pv);
    W8A8BFP32OFP32Linear out_proj(po), fc2(p2);
    W8A8B8O8LinearReLU fc1(p1);
    LayerNormQ self_attn_layer_norm(l1), final_layer_norm(l2);
    BMM_S8T_S8N_F32T qk_bmm;
    BMM_S8T_S8N_S8T pv_bmm;
    Int8OPTDecoderLayer layer(std::string(param_path), cfg, 0, self_attn_layer_norm, final_layer_norm, fc1, fc2, qk_bmm, pv_bmm, k_proj, v_proj, q_proj,
                              out_proj);
    Matrix3D<int8_t> past_k, past_v;
    int past = 0, row = 0;
    for (int call = 0; call < 1 + decode_steps; call++) {
        const int sqlen = call == 0 ? prefill : 1, tgz = past + sqlen;
        std::vector<float> mask((size_t)sqlen * tgz, 0.f);
        for (int i = 0; i < sqlen; i++)
            for (int j = past + i + 1; j < tgz; j++) mask[(size_t)i * tgz + j] = -3.402823466e38f;
        Matrix3D<float> hs(const_cast<float *>(hidden) + (size_t)row * E, 1, sqlen, E);
        Matrix3D<float> m(mask.data(), 1, sqlen, tgz);
        struct Int8OPTDecoderLayer_output out =
            call == 0 ? layer.forward(Int8OPTDecoderLayer_input(hs, m)) : layer.forward(Int8OPTDecoderLayer_input(hs, m, past_k, past_v));
        memcpy(out_fp + (size_t)row * E, out.hidden_states.m_data, (size_t)sqlen * E * sizeof(float));
        past_k = out.past_key_value.first;
        past_v = out.past_key_value.second;
        past = tgz;
        row += sqlen;
    }
    memcpy(final_k, past_k.m_data, (size_t)H * past * hd);
    memcpy(final_v, past_v.m_data, (size_t)H * past * hd);
    return past;
}

// ---- the small fp32 ops either side of the path, as the reference's CPU classes compute them (pins orc_rmsnorm / orc_layernorm_q) ----
extern "C" {

// LlamaRMSNorm::forward (llm/src/ops/LlamaRMSNorm.cc:7-38): x, out fp32 [rows][dim], weight fp32 [dim]
void ref_llama_rmsnorm(const float *x, const float *weight, float *out, int rows, int dim, float eps) {
    LlamaRMSNorm op(Matrix3D<float>(const_cast<float *>(weight), 1, 1, dim));
    Matrix3D<float> X(const_cast<float *>(x), 1, rows, dim), Y(out, 1, rows, dim);
    op.forward(X, Y, eps);
}

// LayerNormQ::forward (llm/src/ops/LayerNormQ.cc:11-51): fp32 in, int8 out = round(layernorm), eps fixed at 1e-5 by the reference
void ref_layernorm_q(const float *x, const float *weight, const float *bias, int8_t *out, int rows, int dim) {
    LayerNormQ_params p;
    p.weight = Matrix3D<float>(const_cast<float *>(weight), 1, 1, dim);
    p.bias = Matrix3D<float>(const_cast<float *>(bias), 1, 1, dim);
    LayerNormQ op(p);
    Matrix3D<float> X(const_cast<float *>(x), 1, rows, dim);
    Matrix3D<int8_t> Y(out, 1, rows, dim);
    op.forward(X, Y);
}
}
