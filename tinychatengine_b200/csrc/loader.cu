// loader.cu -- reads the reference's INT4 parameter tree from disk and builds the decode model from it.
//
// Tree layout (llm/tools/model_quantizer.py:160-430 writes it, the constructors in llm/src/nn_modules/cuda/Int4llama*.cu:7-90 read it):
//   <dir>/decoder/embed_tokens/weight.bin                      fp32 [vocab][E]       (the reference gathers on the CPU, then float2half)
//   <dir>/decoder/norm/weight.bin                              fp32 [E]
//   <dir>/decoder/layer<i>/{input,post_attention}_layernorm/weight.bin   fp32 [E]
//   <dir>/decoder/layer<i>/self_attn/qkv_proj/                 one INT4 op: q|k|v rows concatenated (llm/tools/llama_qkv_merger.py:13-58)
//        ... or self_attn/{q,k,v}_proj/ when the tree has not been merged
//   <dir>/decoder/layer<i>/self_attn/o_proj/, layer<i>/{gate,up,down}_proj/, <dir>/lm_head/     INT4 ops
//   <dir>/decoder/layer0/self_attn/rotary_emb/{cos,sin}_cached_half.bin   fp16 [max_sqlen][128]   (optional: else computed from rope_theta)
//   <dir>/decoder/layer0/self_attn/qk_bmm/alpha_half.bin                   fp16 [1]                (optional: else cfg.qk_alpha)
// An INT4 op directory in the QM_CUDA flavour (quantize_row_q4_6, llm/tools/quantize_methods.py:370-442; dtypes model_quantizer.py:35-50):
//   weight_int4.bin int32 [OC][IC/8] (8 nibbles per word, element i in bits 4*(i%8)), scaling_factor_int4.bin fp16 [OC][zeros_w*8],
//   zero_point_int4.bin int32 [OC][zeros_w]; zeros_w = calculate_zeros_width(IC, 128).
// The QM_x86 flavour (quantize_row_q4_3, quantize_methods.py:188-243: group 32, byte e of a 64-weight row = w[e] | w[32+e] << 4, fp32
// scales, zero point 8) cannot be carried over losslessly -- the kernels here scale per 128-group like the reference's CUDA build -- so
// tce_w4_import_x86 dequantises it exactly and re-quantises with the QM_CUDA rule, i.e. what re-running the reference's quantizer with
// --method QM_CUDA on the dequantised weights would produce.
#include <sys/stat.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "llama_decoder.h"

namespace tce {

namespace {

int zeros_width_128(int ic) { return (ic / 128 + 7) / 8; }  // calculate_zeros_width(in_features, 128): llm/tools/quantize_methods.py:9-21

bool exists(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

// whole file -> host vector; size must match exactly
bool read_exact(const std::string &path, size_t bytes, std::vector<uint8_t> &buf, std::string *err) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
        *err = "cannot open " + path;
        return false;
    }
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 0 || (size_t)sz != bytes) {
        fclose(f);
        *err = path + ": " + std::to_string(sz) + " bytes on disk, " + std::to_string(bytes) + " expected";
        return false;
    }
    buf.resize(bytes);
    const size_t got = bytes ? fread(buf.data(), 1, bytes, f) : 0;
    fclose(f);
    if (got != bytes) {
        *err = "short read on " + path;
        return false;
    }
    return true;
}

__global__ void f32_to_f16_kernel(const float *in, __half *out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = __float2half(in[i]);
}

struct Loader {
    Ctx *ctx;
    std::vector<void *> owned;
    std::string err;
    cudaError_t cuda = cudaSuccess;

    void *upload(const void *host, size_t bytes) {
        void *d = nullptr;
        cuda = cudaMalloc(&d, bytes ? bytes : 1);
        if (cuda != cudaSuccess) {
            err = "cudaMalloc of " + std::to_string(bytes) + " bytes failed";
            return nullptr;
        }
        owned.push_back(d);
        cuda = cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice);
        if (cuda != cudaSuccess) {
            err = "cudaMemcpy failed";
            return nullptr;
        }
        return d;
    }
    void *file_to_device(const std::string &path, size_t bytes) {
        std::vector<uint8_t> buf;
        if (!read_exact(path, bytes, buf, &err)) return nullptr;
        return upload(buf.data(), bytes);
    }
    // one INT4 op directory (QM_CUDA): rows x ic
    bool w4(const std::string &dir, int oc, int ic, tce_w4_tensor *t) {
        const int zw = zeros_width_128(ic);
        t->w = file_to_device(dir + "/weight_int4.bin", (size_t)oc * (ic / 8) * 4);
        if (!t->w) return false;
        t->scales = file_to_device(dir + "/scaling_factor_int4.bin", (size_t)oc * zw * 8 * 2);
        if (!t->scales) return false;
        t->zeros = file_to_device(dir + "/zero_point_int4.bin", (size_t)oc * zw * 4);
        if (!t->zeros) return false;
        t->oc = oc;
        t->ic = ic;
        return true;
    }
    void release() {
        for (void *p : owned) cudaFree(p);
        owned.clear();
    }
};

tce_w4_tensor rows_of(const tce_w4_tensor &t, int row0, int rows) {  // a row range of a loaded op (q | k | v inside qkv_proj)
    tce_w4_tensor r = t;
    const int zw = zeros_width_128(t.ic);
    r.w = static_cast<const uint8_t *>(t.w) + (size_t)row0 * (t.ic / 8) * 4;
    r.scales = static_cast<const uint8_t *>(t.scales) + (size_t)row0 * zw * 8 * 2;
    r.zeros = static_cast<const uint8_t *>(t.zeros) + (size_t)row0 * zw * 4;
    r.oc = rows;
    return r;
}

}  // namespace

LlamaDecoder *load_llama_dir(Ctx *ctx, int attn_chunk, const char *dir_c, tce_llama_config cfg, std::string *err) {
    const std::string dir(dir_c);
    Loader L{ctx};
    const int E = cfg.embed_dim, H = cfg.num_heads, KVH = cfg.num_kv_heads, hd = cfg.head_dim, F = cfg.hidden_dim, V = cfg.vocab_size, NL = cfg.num_layers;
    auto fail = [&](const std::string &m) -> LlamaDecoder * {
        *err = m;
        L.release();
        return nullptr;
    };
    if (cfg.tp_size > 1) return fail("load_dir: single GPU only (shard the tree first)");
    if (E % 128 || F % 128 || hd != 128) return fail("load_dir: embed/hidden must be multiples of 128 and head_dim 128");
    const std::string dec = dir + "/decoder";
    // embedding: fp32 on disk -> fp16 on the device (the reference converts after the gather, Int4llamaDecoder.cu:62-69)
    __half *embed = nullptr;
    {
        const size_t n = (size_t)V * E;
        float *tmp = static_cast<float *>(L.file_to_device(dec + "/embed_tokens/weight.bin", n * 4));
        if (!tmp) return fail(L.err);
        if (cudaMalloc((void **)&embed, n * 2) != cudaSuccess) return fail("cudaMalloc(embedding) failed");
        L.owned.push_back(embed);
        f32_to_f16_kernel<<<1024, 256, 0, ctx->stream>>>(tmp, embed, n);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail("embedding conversion failed");
        cudaFree(tmp);  // the fp32 copy is not kept
        for (auto &p : L.owned)
            if (p == tmp) p = nullptr;
    }
    std::vector<tce_llama_layer> layers(NL);
    for (int l = 0; l < NL; l++) {
        const std::string lp = dec + "/layer" + std::to_string(l);
        tce_llama_layer &y = layers[l];
        y.input_norm = static_cast<const float *>(L.file_to_device(lp + "/input_layernorm/weight.bin", (size_t)E * 4));
        y.post_norm = y.input_norm ? static_cast<const float *>(L.file_to_device(lp + "/post_attention_layernorm/weight.bin", (size_t)E * 4)) : nullptr;
        if (!y.post_norm) return fail(L.err);
        const std::string sa = lp + "/self_attn";
        if (exists(sa + "/qkv_proj/weight_int4.bin")) {
            tce_w4_tensor qkv;
            if (!L.w4(sa + "/qkv_proj", (H + 2 * KVH) * hd, E, &qkv)) return fail(L.err);
            y.q = rows_of(qkv, 0, H * hd);
            y.k = rows_of(qkv, H * hd, KVH * hd);
            y.v = rows_of(qkv, (H + KVH) * hd, KVH * hd);
        } else {
            if (!L.w4(sa + "/q_proj", H * hd, E, &y.q) || !L.w4(sa + "/k_proj", KVH * hd, E, &y.k) || !L.w4(sa + "/v_proj", KVH * hd, E, &y.v)) return fail(L.err);
        }
        if (!L.w4(sa + "/o_proj", E, H * hd, &y.o) || !L.w4(lp + "/gate_proj", F, E, &y.gate) || !L.w4(lp + "/up_proj", F, E, &y.up) ||
            !L.w4(lp + "/down_proj", E, F, &y.down))
            return fail(L.err);
    }
    tce_llama_weights w{};
    w.embed_f16 = embed;
    w.layers = layers.data();
    w.final_norm = static_cast<const float *>(L.file_to_device(dec + "/norm/weight.bin", (size_t)E * 4));
    if (!w.final_norm) return fail(L.err);
    if (!L.w4(dir + "/lm_head", V, E, &w.lm_head)) return fail(L.err);
    // optional tables of layer 0 (every layer carries the same ones)
    const std::string sa0 = dec + "/layer0/self_attn";
    if (exists(sa0 + "/rotary_emb/cos_cached_half.bin") && exists(sa0 + "/rotary_emb/sin_cached_half.bin")) {
        struct stat st;
        stat((sa0 + "/rotary_emb/cos_cached_half.bin").c_str(), &st);
        const size_t rows = (size_t)st.st_size / ((size_t)hd * 2);
        if (rows < (size_t)cfg.max_ctx) return fail("load_dir: rotary tables cover " + std::to_string(rows) + " positions, max_ctx is " + std::to_string(cfg.max_ctx));
        std::vector<uint8_t> hc, hs;
        if (!read_exact(sa0 + "/rotary_emb/cos_cached_half.bin", rows * hd * 2, hc, &L.err) || !read_exact(sa0 + "/rotary_emb/sin_cached_half.bin", rows * hd * 2, hs, &L.err))
            return fail(L.err);
        std::vector<float> fc((size_t)cfg.max_ctx * hd), fs((size_t)cfg.max_ctx * hd);
        const __half *pc = reinterpret_cast<const __half *>(hc.data()), *ps = reinterpret_cast<const __half *>(hs.data());
        for (size_t i = 0; i < fc.size(); i++) {
            fc[i] = __half2float(pc[i]);
            fs[i] = __half2float(ps[i]);
        }
        w.rope_cos = static_cast<const float *>(L.upload(fc.data(), fc.size() * 4));
        w.rope_sin = w.rope_cos ? static_cast<const float *>(L.upload(fs.data(), fs.size() * 4)) : nullptr;
        if (!w.rope_sin) return fail(L.err);
    }
    if (exists(sa0 + "/qk_bmm/alpha_half.bin")) {
        std::vector<uint8_t> a;
        if (!read_exact(sa0 + "/qk_bmm/alpha_half.bin", 2, a, &L.err)) return fail(L.err);
        cfg.qk_alpha = __half2float(*reinterpret_cast<const __half *>(a.data()));
    }
    LlamaDecoder *d = LlamaDecoder::create(ctx, attn_chunk, cfg, w, err);
    if (!d) {
        L.release();
        return nullptr;
    }
    for (void *p : L.owned)
        if (p) d->adopt(p);
    return d;
}

// QM_x86 op (host arrays) -> QM_CUDA op (host arrays).  Exact dequantisation, then the QM_CUDA rule of quantize_row_q4_6
// (quantize_methods.py:393-442): per 128-group d = (element of largest magnitude) / -8, q = trunc(clip(x / d + 8.5, 0, 15)), zero point 8,
// fp16 scales, padding of the scale / zero rows up to zeros_w * 8 groups (zero nibbles in the padding hold 8 as in the reference's files).
int import_x86(const uint8_t *qs, const float *d32, int oc, int ic, uint32_t *w_out, __half *scales_out, uint32_t *zeros_out) {
    if (ic % 128 || oc < 1) return 1;
    const int zw = zeros_width_128(ic), ng = ic / 128;
    std::vector<float> row(ic);
    for (int r = 0; r < oc; r++) {
        // 64 consecutive weights live in 32 bytes: byte e = w[e] | w[32 + e] << 4; one fp32 scale per 32 weights
        for (int b = 0; b < ic / 64; b++) {
            const uint8_t *p = qs + (size_t)r * (ic / 2) + (size_t)b * 32;
            const float d0 = d32[(size_t)r * (ic / 32) + 2 * b], d1 = d32[(size_t)r * (ic / 32) + 2 * b + 1];
            for (int e = 0; e < 32; e++) {
                row[b * 64 + e] = (float)((int)(p[e] & 0xF) - 8) * d0;
                row[b * 64 + 32 + e] = (float)((int)(p[e] >> 4) - 8) * d1;
            }
        }
        for (int i = 0; i < ic / 8; i++) w_out[(size_t)r * (ic / 8) + i] = 0u;
        for (int i = 0; i < zw * 8; i++) scales_out[(size_t)r * zw * 8 + i] = __float2half(0.f);
        for (int i = 0; i < zw; i++) zeros_out[(size_t)r * zw + i] = 0x88888888u;
        for (int g = 0; g < ng; g++) {
            const float *x = row.data() + g * 128;
            float mv = x[0];
            for (int i = 1; i < 128; i++)
                if (std::fabs(x[i]) > std::fabs(mv)) mv = x[i];  // first element of largest magnitude (np.argmax)
            const float d = mv / -8.0f;
            const float id = d != 0.f ? 1.0f / d : 0.f;
            for (int i = 0; i < 128; i++) {
                float v = x[i] * id + 8.5f;
                v = v < 0.f ? 0.f : (v > 15.f ? 15.f : v);
                const uint32_t q = (uint32_t)(int)v;
                const int k = g * 128 + i;
                w_out[(size_t)r * (ic / 8) + k / 8] |= q << (4 * (k % 8));
            }
            scales_out[(size_t)r * zw * 8 + g] = __float2half(d);
        }
    }
    return 0;
}

}  // namespace tce
