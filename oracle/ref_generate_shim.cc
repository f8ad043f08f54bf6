// ref_generate_shim.cc -- C entry point over the reference's own sampling functions (llm/src/Generate.cc, compiled from
// /root/reference by oracle/Makefile into oracle/_ref/libtce_ref_generate.so).  TEST INFRASTRUCTURE: pins oracle/sampling.py.
// The chain and its order are LLaMAGenerate.cu:112-166's; the final draw is left out (the caller gets the candidate set and the
// probabilities sample_token would draw from).
#include <vector>

#include "Generate.h"

extern "C" int ref_sample_candidates(const float *logits, int n_vocab, const int *window, int n_window, int top_k, float top_p, float temp,
                                     float repeat_penalty, float frequency_penalty, float presence_penalty, int *out_ids, float *out_probs) {
    std::vector<OPT_token_data> cand;
    cand.reserve(n_vocab);
    for (int i = 0; i < n_vocab; i++) cand.push_back(OPT_token_data{i, logits[i], 0.0f});
    OPT_token_data_array arr = {cand.data(), cand.size(), false};
    sample_repetition_penalty(&arr, window, (size_t)n_window, repeat_penalty);
    sample_frequency_and_presence_penalties(&arr, window, (size_t)n_window, frequency_penalty, presence_penalty);
    if (temp <= 0) {
        out_ids[0] = sample_token_greedy(&arr);
        out_probs[0] = 1.0f;
        return 1;
    }
    const int k = top_k <= 0 ? n_vocab : top_k;
    sample_top_k(&arr, k, 1);
    sample_tail_free(&arr, 1.0f, 1);
    sample_typical(&arr, 1.0f, 1);
    sample_top_p(&arr, top_p, 1);
    sample_temperature(&arr, temp);
    sample_softmax(&arr);
    for (size_t i = 0; i < arr.size; i++) {
        out_ids[i] = arr.data[i].id;
        out_probs[i] = arr.data[i].p;
    }
    return (int)arr.size;
}
