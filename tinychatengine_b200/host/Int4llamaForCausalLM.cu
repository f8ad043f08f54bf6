// Int4llamaForCausalLM.cu -- see the header.  Reference behaviour: llm/src/nn_modules/cuda/Int4llamaForCausalLM.cu:7-60.
#include "Int4llamaForCausalLM.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/tce_b200.h"

static void die(const char *what) {
    fprintf(stderr, "Int4LlamaForCausalLM: %s: %s\n", what, tce_last_error());
    exit(1);  // the reference asserts / exits on load and launch failures as well (utils.cu CHECK_CUDA)
}

Int4LlamaForCausalLM::Int4LlamaForCausalLM(std::string param_path, const struct model_config config) : config_(config) {
    tce_llama_config c = {};
    c.num_layers = config.num_layers;
    c.num_heads = config.num_heads;
    c.num_kv_heads = config.num_kv_heads;
    c.head_dim = config.embed_dim / config.num_heads;
    c.embed_dim = config.embed_dim;
    c.hidden_dim = config.hidden_dim;
    c.vocab_size = config.vocsize;
    c.max_ctx = config.max_sqlen;
    c.rms_eps = config.rms_norm_eps;
    c.rope_theta = 10000.0f;  // used only when the tree carries no rotary_emb tables
    c.qk_alpha = 0.f;
    c.tp_rank = 0;
    c.tp_size = 1;
    if (tce_llama_load_dir(tce_host_ctx(), param_path.c_str(), &c, &model_) != TCE_OK) die("load");
}

struct Int4LlamaForCausalLM_output Int4LlamaForCausalLM::forward(std::string, const struct Int4LlamaForCausalLM_input &input) {
    const int sqlen = input.input_ids.m_dim_z;
    const int past = input.has_past_keys_values && !input.past_keys.empty() ? input.past_keys[0].m_dim_y : 0;
    if (sqlen > logits_rows_) {
        if (logits_output) cudaFree(logits_output);
        if (cudaMallocManaged((void **)&logits_output, (size_t)sqlen * config_.vocsize * sizeof(float)) != cudaSuccess) die("logits allocation");
        logits_rows_ = sqlen;
    }
    float *last = logits_output + (size_t)(sqlen - 1) * config_.vocsize;
    if (sqlen > 1) cudaMemset(logits_output, 0, (size_t)(sqlen - 1) * config_.vocsize * sizeof(float));
    int rc;
    if (sqlen == 1)
        rc = tce_llama_decode_host(model_, input.input_ids.m_data[0], past, last, nullptr);
    else
        rc = tce_llama_prefill(model_, input.input_ids.m_data, sqlen, past, last, nullptr);
    if (rc != TCE_OK) die("forward");
    struct Int4LlamaForCausalLM_output out;
    out.logits = Matrix3D<float>(logits_output, 1, sqlen, config_.vocsize);
    const int hd = config_.embed_dim / config_.num_heads;
    for (int l = 0; l < config_.num_layers; l++) {
        out.past_keys.push_back(Matrix3D<float16_t>(static_cast<float16_t *>(tce_llama_kv_cache(model_, l, 0)), config_.num_kv_heads, past + sqlen, hd));
        out.past_values.push_back(Matrix3D<float16_t>(static_cast<float16_t *>(tce_llama_kv_cache(model_, l, 1)), config_.num_kv_heads, past + sqlen, hd));
    }
    return out;
}

void Int4LlamaForCausalLM::free_cuda_memory() {
    if (logits_output) cudaFree(logits_output);
    logits_output = nullptr;
    logits_rows_ = 0;
    if (model_) tce_llama_destroy(model_);
    model_ = nullptr;
}
