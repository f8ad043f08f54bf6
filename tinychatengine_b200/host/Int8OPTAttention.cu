// Int8OPTAttention.cu -- forward() of the reference module (llm/src/nn_modules/Int8OPTAttention.cc:183-284) on the device:
// three W8A8B8O8Linear projections, the fused int8 attention core (tce_opt_int8_attention: shape, KV concat, QK^T, mask,
// softmax, int8 probabilities, PV, unshape in two kernels) and the W8A8BFP32OFP32Linear output projection.
#include "Int8OPTAttention.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/tce_b200.h"

extern "C" tce_ctx *tce_host_ctx(void);

static int8_t *query_states_unshape_arr, *key_states_unshape_arr, *value_states_unshape_arr, *attn_output_transpose_arr;
static float *attn_output_fp_arr;
static int8_t ***key_states_arr_cache, ***value_states_arr_cache;
static int *cache_num;

template <typename T>
static void device_alloc(T *&p, size_t bytes) {
    if (cudaMalloc((void **)&p, bytes) != cudaSuccess) {
        fprintf(stderr, "Int8OPTAttention: cudaMalloc(%zu) failed\n", bytes);
        exit(1);
    }
}

void Int8OPTAttention::initialized_memory(const struct model_config config) {
    const size_t act = (size_t)config.max_sqlen * config.embed_dim;
    device_alloc(query_states_unshape_arr, act);
    device_alloc(key_states_unshape_arr, act);
    device_alloc(value_states_unshape_arr, act);
    device_alloc(attn_output_transpose_arr, act);
    device_alloc(attn_output_fp_arr, act * sizeof(float));
    cache_num = new int[config.num_layers]();
    key_states_arr_cache = new int8_t **[config.num_layers];
    value_states_arr_cache = new int8_t **[config.num_layers];
    for (int i = 0; i < config.num_layers; i++) {
        key_states_arr_cache[i] = new int8_t *[2];
        value_states_arr_cache[i] = new int8_t *[2];
        for (int j = 0; j < 2; j++) {
            device_alloc(key_states_arr_cache[i][j], act);
            device_alloc(value_states_arr_cache[i][j], act);
        }
    }
}

Int8OPTAttention::Int8OPTAttention(const struct model_config config, BMM_S8T_S8N_F32T &qk_bmm_, BMM_S8T_S8N_S8T &pv_bmm_, W8A8B8O8Linear &k_proj_,
                                   W8A8B8O8Linear &v_proj_, W8A8B8O8Linear &q_proj_, W8A8BFP32OFP32Linear &out_proj_)
    : embed_dim(config.embed_dim), num_heads(config.num_heads), head_dim(config.embed_dim / config.num_heads), q_proj(q_proj_), k_proj(k_proj_),
      v_proj(v_proj_), out_proj(out_proj_), qk_bmm(qk_bmm_), pv_bmm(pv_bmm_) {
    assert(config.embed_dim % config.num_heads == 0);
}

Int8OPTAttention::Int8OPTAttention(std::string param_path, const struct model_config config, BMM_S8T_S8N_F32T &qk_bmm_, BMM_S8T_S8N_S8T &pv_bmm_,
                                   W8A8B8O8Linear &k_proj_, W8A8B8O8Linear &v_proj_, W8A8B8O8Linear &q_proj_, W8A8BFP32OFP32Linear &out_proj_)
    : embed_dim(config.embed_dim), num_heads(config.num_heads), head_dim(config.embed_dim / config.num_heads) {
    assert(config.embed_dim % config.num_heads == 0);
    // load order and directory names of llm/src/nn_modules/Int8OPTAttention.cc:63-68; the caller's operators are filled in place, as there
    load_BMM_S8T_S8N_F32T(qk_bmm_, param_path + "/qk_bmm");
    load_BMM_S8T_S8N_S8T(pv_bmm_, param_path + "/pv_bmm");
    load_W8A8B8O8Linear_params(k_proj_, param_path + "/k_proj");
    load_W8A8B8O8Linear_params(v_proj_, param_path + "/v_proj");
    load_W8A8B8O8Linear_params(q_proj_, param_path + "/q_proj");
    load_W8A8BFP32OFP32Linear_params(out_proj_, param_path + "/out_proj");
    qk_bmm = qk_bmm_;
    pv_bmm = pv_bmm_;
    k_proj = k_proj_;
    v_proj = v_proj_;
    q_proj = q_proj_;
    out_proj = out_proj_;
}

struct Int8OPTAttention_output Int8OPTAttention::forward(const struct Int8OPTAttention_input &input) {
    struct Int8OPTAttention_output output;
    const int sqlen = input.hidden_states.m_dim_y, b = input.hidden_states.m_dim_x;
    assert(b == 1);
    Matrix3D<int8_t> query_states_unshape(query_states_unshape_arr, b, sqlen, embed_dim);
    Matrix3D<int8_t> key_states_unshape(key_states_unshape_arr, b, sqlen, embed_dim);
    Matrix3D<int8_t> value_states_unshape(value_states_unshape_arr, b, sqlen, embed_dim);
    q_proj.forward(input.hidden_states, query_states_unshape);
    k_proj.forward(input.hidden_states, key_states_unshape);
    v_proj.forward(input.hidden_states, value_states_unshape);

    // the reference ping-pongs two cache buffers per layer so that past and final never alias (:191-200)
    const int which = cache_num[input.layer_idx] == 1 ? 1 : 0;
    cache_num[input.layer_idx] = which ? 0 : 1;
    int8_t *ret_value_states = value_states_arr_cache[input.layer_idx][which], *ret_key_states = key_states_arr_cache[input.layer_idx][which];

    int past = 0;
    if (input.has_past_key_value) {
        assert(input.past_key.m_dim_z == head_dim);
        past = input.past_key.m_dim_y;
    }
    const int tgz = sqlen + past;
    Matrix3D<int8_t> attn_output_transpose(attn_output_transpose_arr, 1, sqlen, embed_dim);
    int rc = tce_opt_int8_attention(tce_host_ctx(), query_states_unshape_arr, key_states_unshape_arr, value_states_unshape_arr,
                                    past ? input.past_key.m_data : nullptr, past ? input.past_value.m_data : nullptr, (long long)past * head_dim,
                                    ret_key_states, ret_value_states, (long long)tgz * head_dim, input.attention_mask.m_data, qk_bmm.alpha, pv_bmm.alpha,
                                    sqlen, past, num_heads, head_dim, attn_output_transpose_arr);
    if (rc != TCE_OK) {
        fprintf(stderr, "Int8OPTAttention: %s\n", tce_last_error());
        exit(1);
    }
    Matrix3D<float> attn_output_fp(attn_output_fp_arr, 1, sqlen, embed_dim);
    out_proj.forward(attn_output_transpose, attn_output_fp);

    output.attn_output = attn_output_fp;
    output.past_key_value = {Matrix3D<int8_t>(ret_key_states, num_heads, tgz, head_dim), Matrix3D<int8_t>(ret_value_states, num_heads, tgz, head_dim)};
    return output;
}
