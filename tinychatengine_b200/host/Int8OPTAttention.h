// Int8OPTAttention.h -- device-side counterpart of the reference module on the W8A8 hot path
// (interface: llm/include/nn_modules/Int8OPTAttention.h:1-52, behaviour: llm/src/nn_modules/Int8OPTAttention.cc:183-284).
//
// Contract kept for the reference's callers (Int8OPTDecoderLayer): the same type names, the same fields in the input / output
// structs, forward() returning fp32 attn_output [1][sqlen][E] and the concatenated int8 K / V [H][tgz][hd] as past_key_value out of
// two alternating per-layer buffers.  Differences: every buffer lives in device memory.  Both constructors exist: the reference's
// (param_path first: loads the six operators from `<param_path>/{qk_bmm,pv_bmm,k_proj,v_proj,q_proj,out_proj}` like
// llm/src/nn_modules/Int8OPTAttention.cc:60-81) and one taking operators that already hold their weights.
#ifndef TCE_HOST_INT8OPTATTENTION_H
#define TCE_HOST_INT8OPTATTENTION_H
#include <utility>

#include "ops.h"

#ifndef TCE_HAVE_MODEL_CONFIG
// the subset of llm/include/model.h:5-21 this module reads (define TCE_HAVE_MODEL_CONFIG when the reference header is included)
struct model_config {
    int batch = 1, num_heads = 12, num_kv_heads = 12, num_layers = 12, max_sqlen = 2048, embed_dim = 768, hidden_dim = 3072, vocsize = 50272, padding_idx = 1;
    float rms_norm_eps = 0.f;
};
#endif

typedef Matrix3D<int8_t> tce_i8_tensor;

struct Int8OPTAttention_output {
    Matrix3D<float> attn_output;                                  // [1][sqlen][embed_dim]
    tce_i8_tensor attn_probs_reshaped;                            // declared by the reference, never filled there either
    std::pair<tce_i8_tensor, tce_i8_tensor> past_key_value;       // K, V: [heads][past + sqlen][head_dim]
};

struct Int8OPTAttention_input {
    tce_i8_tensor hidden_states;         // [1][sqlen][embed_dim]
    Matrix3D<float> attention_mask;      // [1][sqlen][past + sqlen]
    tce_i8_tensor past_key, past_value;  // [heads][past][head_dim], read only when has_past_key_value
    bool has_past_key_value = false;
    int layer_idx;

    // first call of a sequence (no cache yet) / later calls
    Int8OPTAttention_input(tce_i8_tensor hidden, Matrix3D<float> mask, int layer) : hidden_states(hidden), attention_mask(mask), layer_idx(layer) {}
    Int8OPTAttention_input(tce_i8_tensor hidden, Matrix3D<float> mask, tce_i8_tensor k, tce_i8_tensor v, bool has_past, int layer)
        : hidden_states(hidden), attention_mask(mask), past_key(k), past_value(v), has_past_key_value(has_past), layer_idx(layer) {}
};

class Int8OPTAttention {
   public:
    Int8OPTAttention() {}
    // the reference's constructor (llm/include/nn_modules/Int8OPTAttention.h:33-35): fills the operators from the parameter tree, then keeps copies
    Int8OPTAttention(std::string param_path, const struct model_config config, BMM_S8T_S8N_F32T &qk_bmm, BMM_S8T_S8N_S8T &pv_bmm, W8A8B8O8Linear &k_proj,
                     W8A8B8O8Linear &v_proj, W8A8B8O8Linear &q_proj, W8A8BFP32OFP32Linear &out_proj);
    // same operator order, operators already loaded
    Int8OPTAttention(const struct model_config config, BMM_S8T_S8N_F32T &qk_bmm, BMM_S8T_S8N_S8T &pv_bmm, W8A8B8O8Linear &k_proj, W8A8B8O8Linear &v_proj,
                     W8A8B8O8Linear &q_proj, W8A8BFP32OFP32Linear &out_proj);
    // device scratch + the two alternating KV buffers per layer (llm/src/nn_modules/Int8OPTAttention.cc:27-58)
    static void initialized_memory(const struct model_config config);
    struct Int8OPTAttention_output forward(const struct Int8OPTAttention_input &input);

   private:
    int embed_dim = 0, num_heads = 0, head_dim = 0;
    W8A8B8O8Linear q_proj, k_proj, v_proj;
    W8A8BFP32OFP32Linear out_proj;
    BMM_S8T_S8N_F32T qk_bmm;  // only its alpha is used: the product itself runs inside tce_opt_int8_attention
    BMM_S8T_S8N_S8T pv_bmm;
};
#endif
