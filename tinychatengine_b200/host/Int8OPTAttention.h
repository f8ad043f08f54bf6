// Int8OPTAttention.h -- mirror of the reference module on the W8A8 hot path (llm/include/nn_modules/Int8OPTAttention.h:1-52,
// llm/src/nn_modules/Int8OPTAttention.cc): same input/output structs, same forward() contract (double-buffered
// [H][tgz][hd] int8 KV returned as past_key_value, fp32 attn_output [1][sqlen][E]).  Buffers live in device memory;
// weight loading from param_path is outside the hot path, so the ops are passed in initialised.
#ifndef TCE_HOST_INT8OPTATTENTION_H
#define TCE_HOST_INT8OPTATTENTION_H
#include <utility>

#include "ops.h"

#ifndef TCE_HAVE_MODEL_CONFIG
struct model_config {  // the fields of llm/include/model.h:5-21 this module reads
    int batch = 1, num_heads = 12, num_kv_heads = 12, num_layers = 12, max_sqlen = 2048, embed_dim = 768, hidden_dim = 3072, vocsize = 50272, padding_idx = 1;
    float rms_norm_eps = 0.f;
};
#endif

struct Int8OPTAttention_output {
    Matrix3D<float> attn_output;
    Matrix3D<int8_t> attn_probs_reshaped;  // never filled by the reference either
    std::pair<Matrix3D<int8_t>, Matrix3D<int8_t>> past_key_value;
};
struct Int8OPTAttention_input {
    Matrix3D<int8_t> hidden_states;
    Matrix3D<float> attention_mask;
    Matrix3D<int8_t> past_key, past_value;
    bool has_past_key_value = false;
    int layer_idx;
    Int8OPTAttention_input(Matrix3D<int8_t> hidden_states_, Matrix3D<float> attention_mask_, int layer_idx_)
        : hidden_states(hidden_states_), attention_mask(attention_mask_), layer_idx(layer_idx_) {}
    Int8OPTAttention_input(Matrix3D<int8_t> hidden_states_, Matrix3D<float> attention_mask_, Matrix3D<int8_t> past_key_, Matrix3D<int8_t> past_value_,
                           bool has_past_key_value_, int layer_idx_)
        : hidden_states(hidden_states_), attention_mask(attention_mask_), past_key(past_key_), past_value(past_value_),
          has_past_key_value(has_past_key_value_), layer_idx(layer_idx_) {}
};

class Int8OPTAttention {
   public:
    Int8OPTAttention(const struct model_config config, BMM_S8T_S8N_F32T &qk_bmm, BMM_S8T_S8N_S8T &pv_bmm, W8A8B8O8Linear &k_proj, W8A8B8O8Linear &v_proj,
                     W8A8B8O8Linear &q_proj, W8A8BFP32OFP32Linear &out_proj);
    Int8OPTAttention() {}
    static void initialized_memory(const struct model_config config);  // llm/src/nn_modules/Int8OPTAttention.cc:27-58
    struct Int8OPTAttention_output forward(const struct Int8OPTAttention_input &input);

   private:
    int embed_dim = 0, num_heads = 0, head_dim = 0;
    BMM_S8T_S8N_F32T qk_bmm;
    BMM_S8T_S8N_S8T pv_bmm;
    W8A8B8O8Linear k_proj, v_proj, q_proj;
    W8A8BFP32OFP32Linear out_proj;
};
#endif
