// ref_llama_module_shim.cc -- extern "C" doorway into the UNMODIFIED reference module Int4llamaAttention (CPU build, compiled in
// place from /root/reference/llm/src with the reference's own x86 flags into oracle/_ref/libtce_ref_llama.so).
// TEST INFRASTRUCTURE ONLY.  No arithmetic is re-implemented here: the reference constructor loads a synthetic parameter tree from
// `param_path` (q_proj|k_proj|v_proj|o_proj/{weight_int4,scaling_factor_int4,zero_point_int4}.bin in the QM_x86 layout,
// rotary_emb/{cos,sin}_cached.bin, qk_bmm/alpha.bin, written by oracle/capi.py), and Int4llamaAttention::forward runs a prefill
// followed by single-token steps fed with the returned past_key_value, exactly like Int4llamaDecoderLayer does
// (llm/src/nn_modules/non_cuda/Int4llamaDecoderLayer.cc).  It pins oracle/tce_oracle.c's orc_llama_attention_core.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "nn_modules/Int4llamaAttention.h"
#include "operators.h"
#include "utils.h"

int NUM_THREAD = 2;  // the reference applications define this global (llm/application/chat.cc)

extern "C" {

// hidden: fp32 [T][E] (prefill rows, then one row per decode step).  out: fp32 [T][E] in call order.
// final_k / final_v: fp32 [KVH][T][hd] (post-RoPE keys) after the last call.
// decode_seconds (may be NULL): wall time of the single-token calls alone (the constructor and the prompt pass are outside it) -- bench.py uses it to
// state what the reference's attention module costs per token on the host next to the linears-only CPU baseline.
int ref_int4_llama_attention_timed(const char *param_path, int E, int H, int KVH, int max_sqlen, const float *hidden, int prefill, int decode_steps, float *out,
                                   float *final_k, float *final_v, double *decode_seconds) {
    struct model_config cfg(1, H, KVH, 1, max_sqlen, E, 4 * E, 32000, 1, 1e-5f);
    Int4llamaAttention::initialized_memory(cfg);
    Int4llamaAttention attn(std::string(param_path), cfg, 0);
    const int hd = E / H;
    Matrix3D<float> past_k, past_v;
    int past = 0, row = 0;
    std::chrono::steady_clock::time_point t_decode;
    for (int call = 0; call < 1 + decode_steps; call++) {
        if (call == 1) t_decode = std::chrono::steady_clock::now();
        const int sqlen = call == 0 ? prefill : 1, tgz = past + sqlen;
        // Int4llamaDecoder::prepare_decoder_attention_mask (non_cuda/Int4llamaDecoder.cc): 0 on/below the diagonal, lowest float above
        std::vector<float> mask((size_t)sqlen * tgz, 0.f);
        for (int i = 0; i < sqlen; i++)
            for (int j = past + i + 1; j < tgz; j++) mask[(size_t)i * tgz + j] = std::numeric_limits<float>::lowest();
        std::vector<float> hs(hidden + (size_t)row * E, hidden + (size_t)(row + sqlen) * E);  // the linears quantise their input in scratch, keep ours intact
        Matrix3D<float> x(hs.data(), 1, sqlen, E);
        Matrix3D<float> m(mask.data(), 1, sqlen, tgz);
        struct Int4llamaAttention_output o = call == 0 ? attn.forward(std::string(param_path), Int4llamaAttention_input(x, m, 0))
                                                       : attn.forward(std::string(param_path), Int4llamaAttention_input(x, m, past_k, past_v, true, 0));
        memcpy(out + (size_t)row * E, o.attn_output.m_data, (size_t)sqlen * E * sizeof(float));
        past_k = o.past_key_value.first;
        past_v = o.past_key_value.second;
        past = tgz;
        row += sqlen;
    }
    if (decode_seconds) *decode_seconds = decode_steps > 0 ? std::chrono::duration<double>(std::chrono::steady_clock::now() - t_decode).count() : 0.0;
    memcpy(final_k, past_k.m_data, (size_t)KVH * past * hd * sizeof(float));
    memcpy(final_v, past_v.m_data, (size_t)KVH * past * hd * sizeof(float));
    return past;
}

int ref_int4_llama_attention(const char *param_path, int E, int H, int KVH, int max_sqlen, const float *hidden, int prefill, int decode_steps, float *out,
                             float *final_k, float *final_v) {
    return ref_int4_llama_attention_timed(param_path, E, H, KVH, max_sqlen, hidden, prefill, decode_steps, out, final_k, final_v, nullptr);
}
}
