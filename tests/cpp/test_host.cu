// test_host.cu -- reference-style op tests (llm/tests/cuda/test_ops.cu, llm/tests/non_cuda/test_ops.cc) driven through
// the reference's own class / MatmulOperator surface, with synthetic inputs in cudaMallocManaged buffers (what the
// reference allocates) and the CPU oracle as the expected output.  Unlike the reference mains it returns non-zero on
// failure.  TEST CODE: links oracle/libtce_oracle.so as the checker.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "Int8OPTAttention.h"
#include "ops.h"

extern "C" {
int orc_w4a16_gemv(const uint16_t *x, const uint32_t *w, const uint32_t *zeros, const uint16_t *scales, float *y, uint16_t *y_half, int M, int IC,
                   int OC, int group);
int orc_calculate_zeros_width(int in_features, int group_size);
void orc_int8_matmul(const int8_t *A, const int8_t *B, const int8_t *bias, int8_t *C, int M, int N, int K, float alpha, float beta, int q_min, int q_max);
void orc_int8_matmul_nobias(const int8_t *A, const int8_t *B, int8_t *C, int M, int N, int K, float alpha, int q_min, int q_max);
void orc_int8_matmul_nobias_batch(const int8_t *A, const int8_t *B, int8_t *C, int M, int N, int K, float alpha, int q_min, int q_max);
void orc_int8_matmul_bfp32_ofp32(const int8_t *A, const int8_t *B, const float *bias, float *C, int M, int N, int K, float alpha);
void orc_int8_matmul_nobias_ofp32(const int8_t *A, const int8_t *B, float *C, int M, int N, int K, float alpha);
void orc_int8_matmul_nobias_ofp32_batch(const int8_t *A, const int8_t *B, float *C, int M, int N, int K, float alpha);
int orc_opt_int8_attention_core(const int8_t *q8, const int8_t *k8, const int8_t *v8, const int8_t *past_k, const int8_t *past_v, const float *mask,
                                float qk_alpha, float pv_alpha, int sqlen, int past, int H, int hd, int8_t *attn_out, int8_t *final_k, int8_t *final_v);
uint16_t orc_float_to_half(float f);
float orc_half_to_float(uint16_t h);
}

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 11);
}
static float rndf() { return (float)(rnd() & 0xffffff) / (float)0x1000000; }
static float rndn() { return sqrtf(-2.f * logf(rndf() + 1e-7f)) * cosf(6.2831853f * rndf()); }

template <typename T>
static T *managed(size_t n) {
    T *p = nullptr;
    if (cudaMallocManaged(&p, n * sizeof(T)) != cudaSuccess) {
        fprintf(stderr, "cudaMallocManaged failed\n");
        exit(2);
    }
    return p;
}

static int failures = 0;
static void report(const char *name, bool ok) {
    printf("-------- Test of %s: %s --------\n", name, ok ? "Passed!" : "Fail!");
    if (!ok) failures++;
}

static void test_Linear_half_int4(int m, int n, int k) {
    const int zw = orc_calculate_zeros_width(k, QK);
    int *w = managed<int>((size_t)n * k / 8);
    int *z = managed<int>((size_t)n * zw);
    float16_t *s = managed<float16_t>((size_t)n * zw * 8);
    float16_t *x = managed<float16_t>((size_t)m * k);
    float16_t *y = managed<float16_t>((size_t)m * n);
    for (size_t i = 0; i < (size_t)n * k / 8; i++) w[i] = (int)rnd();
    for (size_t i = 0; i < (size_t)n * zw; i++) z[i] = (int)rnd();
    uint16_t *s16 = reinterpret_cast<uint16_t *>(s), *x16 = reinterpret_cast<uint16_t *>(x);
    for (size_t i = 0; i < (size_t)n * zw * 8; i++) s16[i] = orc_float_to_half((0.5f + rndf()) * 0.004f);
    for (size_t i = 0; i < (size_t)m * k; i++) x16[i] = orc_float_to_half(rndn());
    Linear_half_int4 op(Matrix3D<int>(w, 1, n, k / 8), Matrix3D<float16_t>(s, 1, n, zw * 8), Matrix3D<int>(z, 1, n, zw));
    Matrix3D<float16_t> X(x, 1, m, k), Y(y, 1, m, n);
    op.forward(X, Y);
    cudaDeviceSynchronize();
    std::vector<float> ref((size_t)m * n);
    orc_w4a16_gemv(x16, (const uint32_t *)w, (const uint32_t *)z, s16, ref.data(), nullptr, m, k, n, QK);
    double maxref = 0, maxerr = 0;
    for (size_t i = 0; i < ref.size(); i++) {
        const double got = orc_half_to_float(reinterpret_cast<uint16_t *>(y)[i]);
        maxref = fmax(maxref, fabs(ref[i]));
        maxerr = fmax(maxerr, fabs(got - ref[i]));
    }
    char name[128];
    snprintf(name, sizeof(name), "Linear_half_int4 %dx%d->%d (rel err %.2e, bar 1e-2)", m, k, n, maxerr / maxref);
    report(name, maxerr / maxref <= 1e-2);
    cudaFree(w); cudaFree(z); cudaFree(s); cudaFree(x); cudaFree(y);
}

static int8_t *rand_s8(size_t n) {
    int8_t *p = managed<int8_t>(n);
    for (size_t i = 0; i < n; i++) p[i] = (int8_t)((int)(rnd() % 255) - 127);
    return p;
}

static void test_W8A8B8O8Linear(int b, int m, int k, int n, bool relu, float alpha, float beta) {
    int8_t *x = rand_s8((size_t)b * m * k), *w = rand_s8((size_t)n * k), *bias = rand_s8(n), *y = managed<int8_t>((size_t)b * m * n);
    W8A8B8O8Linear_params p = {Matrix3D<int8_t>(w, 1, n, k), Matrix3D<int8_t>(bias, 1, 1, n), alpha, beta};
    Matrix3D<int8_t> X(x, b, m, k), Y(y, b, m, n);
    if (relu) {
        W8A8B8O8LinearReLU op(p);
        op.forward(X, Y);
    } else {
        W8A8B8O8Linear op(p);
        op.forward(X, Y);
    }
    cudaDeviceSynchronize();
    std::vector<int8_t> ref((size_t)b * m * n);
    for (int bz = 0; bz < b; bz++) orc_int8_matmul(x + (size_t)bz * m * k, w, bias, ref.data() + (size_t)bz * m * n, m, n, k, alpha, beta, relu ? 0 : -128, 127);
    bool ok = true;
    for (size_t i = 0; i < ref.size(); i++) ok &= (ref[i] == y[i]);
    report(relu ? "W8A8B8O8LinearReLU (bit exact)" : "W8A8B8O8Linear (bit exact)", ok);
    cudaFree(x); cudaFree(w); cudaFree(bias); cudaFree(y);
}

static void test_W8A8BFP32OFP32Linear(int m, int k, int n, float alpha) {
    int8_t *x = rand_s8((size_t)m * k), *w = rand_s8((size_t)n * k);
    float *bias = managed<float>(n), *y = managed<float>((size_t)m * n);
    for (int i = 0; i < n; i++) bias[i] = rndn();
    W8A8BFP32OFP32Linear_params p = {Matrix3D<int8_t>(w, 1, n, k), Matrix3D<float>(bias, 1, 1, n), alpha};
    W8A8BFP32OFP32Linear op(p);
    Matrix3D<int8_t> X(x, 1, m, k);
    Matrix3D<float> Y(y, 1, m, n);
    op.forward(X, Y);
    cudaDeviceSynchronize();
    std::vector<float> ref((size_t)m * n);
    orc_int8_matmul_bfp32_ofp32(x, w, bias, ref.data(), m, n, k, alpha);
    bool ok = true;
    for (size_t i = 0; i < ref.size(); i++) ok &= (ref[i] == y[i]);
    report("W8A8BFP32OFP32Linear (bit exact fp32)", ok);
    cudaFree(x); cudaFree(w); cudaFree(bias); cudaFree(y);
}

static void test_BMMs(int heads, int m, int t, int d, float alpha) {
    // QK^T: x [heads, m, d], weight [heads, t, d] -> [heads, m, t] fp32 ; PV: p [heads, m, t], V^T [heads, d, t] -> int8
    int8_t *q = rand_s8((size_t)heads * m * d), *kk = rand_s8((size_t)heads * t * d);
    float *s = managed<float>((size_t)heads * m * t);
    BMM_S8T_S8N_F32T qk(alpha);
    Matrix3D<int8_t> Q(q, heads, m, d), K(kk, heads, t, d);
    Matrix3D<float> S(s, heads, m, t);
    qk.forward(Q, K, S);
    int8_t *p = rand_s8((size_t)heads * m * t), *vt = rand_s8((size_t)heads * d * t), *o = managed<int8_t>((size_t)heads * m * d);
    BMM_S8T_S8N_S8T pv(0.0031f);
    Matrix3D<int8_t> P(p, heads, m, t), VT(vt, heads, d, t), O(o, heads, m, d);
    pv.forward(P, VT, O);
    cudaDeviceSynchronize();
    bool ok1 = true, ok2 = true;
    std::vector<float> sref((size_t)m * t);
    std::vector<int8_t> oref((size_t)m * d);
    for (int h = 0; h < heads; h++) {
        orc_int8_matmul_nobias_ofp32(q + (size_t)h * m * d, kk + (size_t)h * t * d, sref.data(), m, t, d, alpha);
        for (size_t i = 0; i < sref.size(); i++) ok1 &= (sref[i] == s[(size_t)h * m * t + i]);
        orc_int8_matmul_nobias(p + (size_t)h * m * t, vt + (size_t)h * d * t, oref.data(), m, d, t, 0.0031f, -128, 127);
        for (size_t i = 0; i < oref.size(); i++) ok2 &= (oref[i] == o[(size_t)h * m * d + i]);
    }
    char name[96];
    snprintf(name, sizeof(name), "BMM_S8T_S8N_F32T heads=%d m=%d (bit exact)", heads, m);
    report(name, ok1);
    snprintf(name, sizeof(name), "BMM_S8T_S8N_S8T heads=%d m=%d (bit exact)", heads, m);
    report(name, ok2);
    cudaFree(q); cudaFree(kk); cudaFree(s); cudaFree(p); cudaFree(vt); cudaFree(o);
}

// Int8OPTAttention::forward, prefill then two decode steps fed with the returned past_key_value (the reference's
// test_Int8OPTAttention / _len512 flow, llm/tests/non_cuda/test_Int8OPTAttention.cc), expected values from the oracle.
static void test_Int8OPTAttention(int E, int H, int prefill) {
    const int hd = E / H;
    struct model_config cfg;
    cfg.num_heads = H; cfg.embed_dim = E; cfg.num_layers = 1; cfg.max_sqlen = 256;
    Int8OPTAttention::initialized_memory(cfg);
    const float a_qkv = 0.0009f, b_qkv = 0.9f, qk_alpha = 0.0007f, pv_alpha = 0.011f, a_out = 0.0008f;
    int8_t *wq = rand_s8((size_t)E * E), *wk = rand_s8((size_t)E * E), *wv = rand_s8((size_t)E * E), *wo = rand_s8((size_t)E * E);
    int8_t *bq = rand_s8(E), *bk = rand_s8(E), *bv = rand_s8(E);
    float *bo = managed<float>(E);
    for (int i = 0; i < E; i++) bo[i] = rndn();
    W8A8B8O8Linear_params pq = {Matrix3D<int8_t>(wq, 1, E, E), Matrix3D<int8_t>(bq, 1, 1, E), a_qkv, b_qkv};
    W8A8B8O8Linear_params pk = {Matrix3D<int8_t>(wk, 1, E, E), Matrix3D<int8_t>(bk, 1, 1, E), a_qkv, b_qkv};
    W8A8B8O8Linear_params pv = {Matrix3D<int8_t>(wv, 1, E, E), Matrix3D<int8_t>(bv, 1, 1, E), a_qkv, b_qkv};
    W8A8BFP32OFP32Linear_params po = {Matrix3D<int8_t>(wo, 1, E, E), Matrix3D<float>(bo, 1, 1, E), a_out};
    W8A8B8O8Linear q_proj(pq), k_proj(pk), v_proj(pv);
    W8A8BFP32OFP32Linear out_proj(po);
    BMM_S8T_S8N_F32T qk_bmm(qk_alpha);
    BMM_S8T_S8N_S8T pv_bmm(pv_alpha);
    Int8OPTAttention attn(cfg, qk_bmm, pv_bmm, k_proj, v_proj, q_proj, out_proj);

    bool ok = true;
    std::vector<int8_t> past_k, past_v;  // oracle-side cache [H][past][hd]
    Matrix3D<int8_t> dev_pk, dev_pv;
    int past = 0;
    for (int step = 0; step < 3; step++) {
        const int sqlen = step == 0 ? prefill : 1, tgz = past + sqlen;
        int8_t *x = rand_s8((size_t)sqlen * E);
        float *mask = managed<float>((size_t)sqlen * tgz);
        for (int i = 0; i < sqlen; i++)
            for (int j = 0; j < tgz; j++) mask[(size_t)i * tgz + j] = j > past + i ? -3.402823466e38f : 0.f;
        Matrix3D<int8_t> X(x, 1, sqlen, E);
        Matrix3D<float> M(mask, 1, sqlen, tgz);
        Int8OPTAttention_output out = step == 0 ? attn.forward(Int8OPTAttention_input(X, M, 0)) : attn.forward(Int8OPTAttention_input(X, M, dev_pk, dev_pv, true, 0));
        cudaDeviceSynchronize();
        // oracle
        std::vector<int8_t> q((size_t)sqlen * E), k((size_t)sqlen * E), v((size_t)sqlen * E), core((size_t)sqlen * E), fk((size_t)H * tgz * hd), fv(fk.size());
        orc_int8_matmul(x, wq, bq, q.data(), sqlen, E, E, a_qkv, b_qkv, -128, 127);
        orc_int8_matmul(x, wk, bk, k.data(), sqlen, E, E, a_qkv, b_qkv, -128, 127);
        orc_int8_matmul(x, wv, bv, v.data(), sqlen, E, E, a_qkv, b_qkv, -128, 127);
        orc_opt_int8_attention_core(q.data(), k.data(), v.data(), past ? past_k.data() : nullptr, past ? past_v.data() : nullptr, mask, qk_alpha, pv_alpha, sqlen,
                                    past, H, hd, core.data(), fk.data(), fv.data());
        std::vector<float> ref((size_t)sqlen * E);
        orc_int8_matmul_bfp32_ofp32(core.data(), wo, bo, ref.data(), sqlen, E, E, a_out);
        std::vector<float> got(ref.size());
        std::vector<int8_t> gk(fk.size()), gv(fv.size());
        cudaMemcpy(got.data(), out.attn_output.m_data, got.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(gk.data(), out.past_key_value.first.m_data, gk.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(gv.data(), out.past_key_value.second.m_data, gv.size(), cudaMemcpyDeviceToHost);
        ok &= out.past_key_value.first.m_dim_y == tgz;
        for (size_t i = 0; i < ref.size(); i++) ok &= (ref[i] == got[i]);
        for (size_t i = 0; i < fk.size(); i++) ok &= (fk[i] == gk[i]) && (fv[i] == gv[i]);
        past_k = fk; past_v = fv; past = tgz;
        dev_pk = out.past_key_value.first; dev_pv = out.past_key_value.second;
        cudaFree(x); cudaFree(mask);
    }
    char name[96];
    snprintf(name, sizeof(name), "Int8OPTAttention E=%d H=%d prefill %d + 2 decode steps (bit exact)", E, H, prefill);
    report(name, ok);
}

// The reference's constructor form: Int8OPTAttention(param_path, config, ops...) fills the six operators from a parameter tree on disk
// (llm/src/nn_modules/Int8OPTAttention.cc:60-81; file names of llm/src/ops/W8A8*.cc / BMM_S8T_*.cc).  A module loaded that way must be
// bit-identical to one built from operators that were handed the same bytes directly.
static void write_file(const std::string &path, const void *data, size_t bytes) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f || fwrite(data, 1, bytes, f) != bytes) {
        fprintf(stderr, "cannot write %s\n", path.c_str());
        exit(2);
    }
    fclose(f);
}
static void test_Int8OPTAttention_param_path(int E, int H, int sqlen) {
    struct model_config cfg;
    cfg.num_heads = H; cfg.embed_dim = E; cfg.num_layers = 1; cfg.max_sqlen = 256;
    Int8OPTAttention::initialized_memory(cfg);
    const float a_qkv = 0.0009f, b_qkv = 0.9f, qk_alpha = 0.0007f, pv_alpha = 0.011f, a_out = 0.0008f;
    char tmpl[] = "/tmp/tce_opt_attn_XXXXXX";
    const char *root_c = mkdtemp(tmpl);
    if (!root_c) { report("Int8OPTAttention(param_path, ...): mkdtemp", false); return; }
    const std::string root(root_c);
    // direct operators (weights in managed memory) and a second, zero-filled set that the constructor must fill from the files
    int8_t *w[4], *w2[4], *b8[3], *b82[3];
    const char *names[4] = {"q_proj", "k_proj", "v_proj", "out_proj"};
    float *bo = managed<float>(E), *bo2 = managed<float>(E);
    for (int i = 0; i < E; i++) { bo[i] = rndn(); bo2[i] = 0.f; }
    for (int i = 0; i < 4; i++) {
        w[i] = rand_s8((size_t)E * E);
        w2[i] = managed<int8_t>((size_t)E * E);
        memset(w2[i], 0, (size_t)E * E);
        if (i < 3) { b8[i] = rand_s8(E); b82[i] = managed<int8_t>(E); memset(b82[i], 0, E); }
        const std::string d = root + "/" + names[i];
        mkdir(d.c_str(), 0700);
        write_file(d + "/weight.bin", w[i], (size_t)E * E);
        if (i < 3) {
            write_file(d + "/bias_int8.bin", b8[i], E);
            write_file(d + "/alpha.bin", &a_qkv, 4);
            write_file(d + "/beta.bin", &b_qkv, 4);
        } else {
            write_file(d + "/bias.bin", bo, (size_t)E * 4);
            write_file(d + "/alpha.bin", &a_out, 4);
        }
    }
    mkdir((root + "/qk_bmm").c_str(), 0700);
    mkdir((root + "/pv_bmm").c_str(), 0700);
    write_file(root + "/qk_bmm/alpha.bin", &qk_alpha, 4);
    write_file(root + "/pv_bmm/alpha.bin", &pv_alpha, 4);
    auto lin = [&](int8_t *wp, int8_t *bp, float a, float b) {
        W8A8B8O8Linear_params p = {Matrix3D<int8_t>(wp, 1, E, E), Matrix3D<int8_t>(bp, 1, 1, E), a, b};
        return W8A8B8O8Linear(p);
    };
    W8A8B8O8Linear q1 = lin(w[0], b8[0], a_qkv, b_qkv), k1 = lin(w[1], b8[1], a_qkv, b_qkv), v1 = lin(w[2], b8[2], a_qkv, b_qkv);
    W8A8B8O8Linear q2 = lin(w2[0], b82[0], 0.f, 0.f), k2 = lin(w2[1], b82[1], 0.f, 0.f), v2 = lin(w2[2], b82[2], 0.f, 0.f);
    W8A8BFP32OFP32Linear_params po1 = {Matrix3D<int8_t>(w[3], 1, E, E), Matrix3D<float>(bo, 1, 1, E), a_out};
    W8A8BFP32OFP32Linear_params po2 = {Matrix3D<int8_t>(w2[3], 1, E, E), Matrix3D<float>(bo2, 1, 1, E), 0.f};
    W8A8BFP32OFP32Linear o1(po1), o2(po2);
    BMM_S8T_S8N_F32T qk1(qk_alpha), qk2;
    BMM_S8T_S8N_S8T pv1(pv_alpha), pv2;
    Int8OPTAttention direct(cfg, qk1, pv1, k1, v1, q1, o1);
    bool ok = true;
    try {
        Int8OPTAttention loaded(root, cfg, qk2, pv2, k2, v2, q2, o2);
        cudaDeviceSynchronize();
        // the caller's operators were filled in place, as in the reference
        ok &= qk2.alpha == qk_alpha && pv2.alpha == pv_alpha && q2.alpha == a_qkv && k2.beta == b_qkv && o2.alpha == a_out;
        for (int i = 0; i < 4; i++) ok &= memcmp(w[i], w2[i], (size_t)E * E) == 0;
        for (int i = 0; i < 3; i++) ok &= memcmp(b8[i], b82[i], E) == 0;
        ok &= memcmp(bo, bo2, (size_t)E * 4) == 0;
        int8_t *x = rand_s8((size_t)sqlen * E);
        float *mask = managed<float>((size_t)sqlen * sqlen);
        for (int i = 0; i < sqlen; i++)
            for (int j = 0; j < sqlen; j++) mask[(size_t)i * sqlen + j] = j > i ? -3.402823466e38f : 0.f;
        Matrix3D<int8_t> X(x, 1, sqlen, E);
        Matrix3D<float> M(mask, 1, sqlen, sqlen);
        std::vector<float> a((size_t)sqlen * E), b(a.size());
        Int8OPTAttention_output o = direct.forward(Int8OPTAttention_input(X, M, 0));
        cudaMemcpy(a.data(), o.attn_output.m_data, a.size() * 4, cudaMemcpyDeviceToHost);
        o = loaded.forward(Int8OPTAttention_input(X, M, 0));
        cudaMemcpy(b.data(), o.attn_output.m_data, b.size() * 4, cudaMemcpyDeviceToHost);
        ok &= memcmp(a.data(), b.data(), a.size() * 4) == 0;
        bool nonzero = false;
        for (float f : a) nonzero |= f != 0.f;
        ok &= nonzero;
        cudaFree(x); cudaFree(mask);
    } catch (const char *msg) {
        fprintf(stderr, "unexpected throw: %s\n", msg);
        ok = false;
    }
    // a missing file throws a C string, like read_to_array (llm/src/utils.cc:16-25)
    bool threw = false;
    remove((root + "/pv_bmm/alpha.bin").c_str());
    try {
        Int8OPTAttention broken(root, cfg, qk2, pv2, k2, v2, q2, o2);
    } catch (const char *) {
        threw = true;
    }
    ok &= threw;
    char name[112];
    snprintf(name, sizeof(name), "Int8OPTAttention(param_path, ...) E=%d H=%d: loaded == direct (bit exact), missing file throws", E, H);
    report(name, ok);
}

int main() {
    // shapes of the reference's op tests (llm/tests/cuda/test_ops.cu:671-724, non_cuda/test_ops.cc:177-478)
    test_Linear_half_int4(1, 11008, 4096);
    test_Linear_half_int4(1, 4096, 11008);
    test_Linear_half_int4(9, 512, 1024);
    test_W8A8B8O8Linear(1, 108, 768, 768, false, 0.00050354f, 0.0213013f);
    test_W8A8B8O8Linear(1, 108, 768, 3072, true, 0.00050354f, 0.0213013f);
    test_W8A8B8O8Linear(2, 1, 2048, 2048, false, 0.00050354f, 0.0213013f);
    test_W8A8BFP32OFP32Linear(512, 768, 768, 0.0012f);
    test_BMMs(12, 64, 64, 64, 0.0021f);
    test_BMMs(12, 1, 300, 64, 0.0021f);
    test_Int8OPTAttention(768, 12, 9);
    test_Int8OPTAttention_param_path(256, 4, 5);
    printf("%d failure(s)\n", failures);
    return failures ? 1 : 0;
}
