// ref_llama_model_shim.cc -- extern "C" doorway into the UNMODIFIED reference model Int4LlamaForCausalLM (CPU build: Int4llamaForCausalLM.cc,
// Int4llamaDecoder.cc, Int4llamaDecoderLayer.cc, Int4llamaAttention.cc under llm/src/nn_modules/non_cuda + the ops they are built from,
// compiled in place from /root/reference with the reference's own x86 flags into oracle/_ref/libtce_ref_llama_model.so).
// TEST INFRASTRUCTURE ONLY.  No arithmetic is re-implemented here: the reference constructors load a synthetic parameter tree
// (`<root>/decoder/{embed_tokens,norm}/weight.bin`, `<root>/decoder/layerN/{input_layernorm,post_attention_layernorm}/weight.bin`,
// `<root>/decoder/layerN/self_attn/{q,k,v,o}_proj/…`, `<root>/decoder/layerN/{gate,up,down}_proj/…`, `<root>/lm_head/…`, all linears in the QM_x86
// layout; written by oracle/capi.py::write_llama_model_params) and Int4LlamaForCausalLM::forward runs one prompt pass followed by single-token
// steps fed with the returned past_keys / past_values, exactly like the generate loop does (llm/src/nn_modules/non_cuda/LLaMAGenerate.cc).
// It pins the COMPOSITION of the decode step the GPU tests are checked against (oracle/llama_ref.py::llama_forward): embedding, RMSNorm
// placement, residual order, SiLU(gate)*up, final norm, lm_head.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "nn_modules/Int4llamaForCausalLM.h"
#include "operators.h"
#include "utils.h"

int NUM_THREAD = 2;  // the reference applications define this global (llm/application/chat.cc)

extern "C" {

// tokens: int32 [prefill + decode_steps].  logits: fp32 [prefill + decode_steps][vocab] in call order (the reference's lm_head runs on every position
// of the prompt pass).  max_sqlen must exceed 626 (Int4llamaDecoder's constructor sizes a LLaVA scratch buffer as max_sqlen - 626 rows).
int ref_int4_llama_causal_lm(const char *param_path, int E, int H, int KVH, int L, int F, int vocab, int max_sqlen, float rms_eps, const int *tokens,
                             int prefill, int decode_steps, float *logits) {
    if (max_sqlen <= 626) return -1;
    struct model_config cfg(1, H, KVH, L, max_sqlen, E, F, vocab, 1, rms_eps);
    Int4LlamaForCausalLM model(std::string(param_path), cfg);
    std::vector<Matrix3D<float>> past_k, past_v;
    int row = 0;
    for (int call = 0; call < 1 + decode_steps; call++) {
        const int sqlen = call == 0 ? prefill : 1;
        std::vector<int> ids(tokens + row, tokens + row + sqlen);
        Matrix3D<int> input_ids(ids.data(), 1, 1, sqlen);
        struct Int4LlamaForCausalLM_output o = call == 0 ? model.forward(std::string(param_path), Int4LlamaForCausalLM_input(input_ids))
                                                         : model.forward(std::string(param_path), Int4LlamaForCausalLM_input(input_ids, past_k, past_v));
        memcpy(logits + (size_t)row * vocab, o.logits.m_data, (size_t)sqlen * vocab * sizeof(float));
        past_k = o.past_keys;
        past_v = o.past_values;
        row += sqlen;
    }
    return row;
}
}
