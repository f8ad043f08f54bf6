// Int4llamaForCausalLM.h -- the reference's top-level module of the W4A16 path as a shell over this library's fused decoder
// (interface: llm/include/nn_modules/Int4llamaForCausalLM.h:1-75 under QM_CUDA; behaviour: llm/src/nn_modules/cuda/Int4llamaForCausalLM.cu:7-60
// with Int4llamaDecoder.cu / Int4llamaDecoderLayer.cu / Int4llamaAttention.cu underneath).
//
// Same type names, the same fields in the input / output structs, the same constructor and forward() signatures, so the reference's
// generate loop (LLaMAGenerate.cu:67-110) drives it unchanged:
//   model = Int4LlamaForCausalLM(param_path, config);  out = model.forward(param_path, {input_ids[, past_keys, past_values]});
// Differences, all behind the interface:
//   * the parameter tree is read once, in the constructor, by the C++ loader (tce_llama_load_dir); forward() ignores param_path like the
//     reference's forward does for everything but profiling names;
//   * logits [1][sqlen][vocab] lives in managed memory as in the reference, but only its LAST row is computed -- the only row any caller
//     reads (LLaMAGenerate.cu:160-166); earlier rows are zero;
//   * past_keys / past_values are views [kv_heads][tokens][128] into the library's KV cache (row pitch 128, head pitch max_sqlen * 128); the
//     callers treat them as opaque handles and only read m_dim_y (the number of cached tokens, Int4llamaDecoder.cu:72-75).
#ifndef TCE_HOST_INT4LLAMAFORCAUSALLM_H
#define TCE_HOST_INT4LLAMAFORCAUSALLM_H
#include <string>
#include <vector>

#include "Int8OPTAttention.h"  // model_config (subset of llm/include/model.h)
#include "ops.h"

struct tce_llama;

struct Int4LlamaForCausalLM_output {
    Matrix3D<float> logits;
    std::vector<Matrix3D<float16_t>> past_keys, past_values;
};

struct Int4LlamaForCausalLM_input {
    Matrix3D<int> input_ids;  // host ints, (1, 1, sqlen)
    Matrix3D<float> image_embed;
    Matrix3D<int> second_input_ids;
    bool has_past_keys_values = false;
    bool is_llava = false;
    std::vector<Matrix3D<float16_t>> past_keys, past_values;

    Int4LlamaForCausalLM_input() {}
    Int4LlamaForCausalLM_input(Matrix3D<int> input_ids_) : input_ids(input_ids_) {}
    Int4LlamaForCausalLM_input(Matrix3D<int> input_ids_, std::vector<Matrix3D<float16_t>> past_keys_, std::vector<Matrix3D<float16_t>> past_values_)
        : input_ids(input_ids_), has_past_keys_values(true), past_keys(past_keys_), past_values(past_values_) {}
};

class Int4LlamaForCausalLM {
   public:
    Int4LlamaForCausalLM(std::string param_path, const struct model_config config);
    Int4LlamaForCausalLM() {}
    struct Int4LlamaForCausalLM_output forward(std::string param_path, const struct Int4LlamaForCausalLM_input &input);
    void free_cuda_memory();
    float *logits_output = nullptr;  // managed, [logits_rows][vocsize]

   private:
    struct model_config config_;
    tce_llama *model_ = nullptr;
    int logits_rows_ = 0;
};
#endif
