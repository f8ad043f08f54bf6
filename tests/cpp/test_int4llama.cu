// test_int4llama.cu -- the reference's module-level test for the W4A16 path (tests/cuda/test_Int4llamaForCausalLM.cu in the reference: build the
// model from a parameter tree, run a prompt, then decode token by token through the module API) re-hosted on this library's shell
// (host/Int4llamaForCausalLM.h).  Assets are synthetic: tests/test_gpu_host_cpp.py writes a tiny tree with LlamaModel.save_dir and passes
// its path plus the geometry; the generated ids are printed and compared there with the Python-driven decode of the same weights
// (itself checked against the oracle in tests/test_gpu_llama.py).
//   usage: test_int4llama <dir> <layers> <heads> <kv_heads> <embed> <hidden> <vocab> <max_sqlen> <eps> <n_decode> <prompt ids...>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "Int4llamaForCausalLM.h"

static int argmax(const float *v, int n) {
    int b = 0;
    for (int i = 1; i < n; i++)
        if (v[i] > v[b]) b = i;
    return b;
}

int main(int argc, char **argv) {
    if (argc < 12) {
        fprintf(stderr, "usage: see the header comment\n");
        return 2;
    }
    struct model_config cfg;
    cfg.batch = 1;
    cfg.num_layers = atoi(argv[2]);
    cfg.num_heads = atoi(argv[3]);
    cfg.num_kv_heads = atoi(argv[4]);
    cfg.embed_dim = atoi(argv[5]);
    cfg.hidden_dim = atoi(argv[6]);
    cfg.vocsize = atoi(argv[7]);
    cfg.max_sqlen = atoi(argv[8]);
    cfg.rms_norm_eps = (float)atof(argv[9]);
    const int n_decode = atoi(argv[10]);
    std::vector<int> prompt;
    for (int i = 11; i < argc; i++) prompt.push_back(atoi(argv[i]));
    const std::string path = argv[1];

    Int4LlamaForCausalLM model(path, cfg);
    // prompt pass (sqlen > 1, no past), as LLaMAGenerate.cu:72-90 does on a new prompt
    Matrix3D<int> ids(prompt.data(), 1, 1, (int)prompt.size());
    struct Int4LlamaForCausalLM_input in0(ids);
    struct Int4LlamaForCausalLM_output out = model.forward(path, in0);
    if (out.logits.m_dim_y != (int)prompt.size() || out.logits.m_dim_z != cfg.vocsize || (int)out.past_keys.size() != cfg.num_layers ||
        out.past_keys[0].m_dim_y != (int)prompt.size()) {
        printf("Fail! output shapes\n");
        return 1;
    }
    int tok = argmax(out.logits.m_data + (size_t)(prompt.size() - 1) * cfg.vocsize, cfg.vocsize);
    printf("ids %d", tok);
    for (int i = 1; i < n_decode; i++) {
        int one = tok;
        Matrix3D<int> id1(&one, 1, 1, 1);
        struct Int4LlamaForCausalLM_input in(id1, out.past_keys, out.past_values);
        out = model.forward(path, in);
        if (out.past_keys[0].m_dim_y != (int)prompt.size() + i) {
            printf("\nFail! past length %d at step %d\n", out.past_keys[0].m_dim_y, i);
            return 1;
        }
        tok = argmax(out.logits.m_data, cfg.vocsize);
        printf(" %d", tok);
    }
    printf("\nPassed!\n");
    model.free_cuda_memory();
    return 0;
}
