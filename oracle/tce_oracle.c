/*
 * tce_oracle.c -- CPU restatement of TinyChatEngine's quantized-linear / KV-attention hot path.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or call it.  The product path (tinychatengine_b200/csrc) never does.
 *
 * Every function restates one reference routine (file:line under /root/reference) in plain C with the
 * same arithmetic order.  Built with -ffp-contract=off -fno-fast-math so float expressions are evaluated
 * exactly as written (no FMA contraction), see oracle/Makefile.
 *
 * Parity pinning: the reference ships no golden vectors for this path (llm/assets is a download), so the
 * oracle is pinned against the reference's own sources compiled in place (oracle/_ref, see
 * oracle/ref_*shim.cc + tests/test_oracle_golden.py) and against fixtures generated from that build
 * (tests/golden/, generator tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * IEEE binary16 helpers (the reference uses half_float::half on the host, llm/half-2.2.0/include/half.hpp,
 * and __half on the device; both are IEEE binary16 with round-to-nearest-even conversions).
 * ---------------------------------------------------------------------------------------------- */
static float h2f(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: normalise */
            int e = -1;
            do {
                e++;
                man <<= 1;
            } while ((man & 0x400u) == 0);
            man &= 0x3ffu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

static uint16_t f2h(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) { /* inf / nan */
        return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? 0x200u : 0));
    }
    if (ax >= 0x477ff000u) { /* rounds to >= 65520 -> inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (ax < 0x33000001u) { /* < 2^-25 (or == 2^-25 tie -> 0) */
        return (uint16_t)sign;
    }
    int32_t e = (int32_t)(ax >> 23) - 127;
    uint32_t m = (ax & 0x7fffffu) | 0x800000u;
    if (e < -14) { /* subnormal half */
        int shift = (-14 - e) + 13;
        uint32_t q = m >> shift;
        uint32_t rem = m & ((1u << shift) - 1);
        uint32_t half = 1u << (shift - 1);
        if (rem > half || (rem == half && (q & 1))) q++;
        return (uint16_t)(sign | q);
    }
    uint32_t q = ((uint32_t)(e + 15) << 10) | ((m >> 13) & 0x3ffu);
    uint32_t rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (q & 1))) q++;
    return (uint16_t)(sign | q);
}

ORC_API float orc_half_to_float(uint16_t h) { return h2f(h); }
ORC_API uint16_t orc_float_to_half(float f) { return f2h(f); }

/* llm/src/nn_modules/cuda/utils.cu:158-178 and llm/tools/quantize_methods.py:6-21 */
ORC_API int orc_calculate_zeros_width(int in_features, int group_size) {
    int mult;
    if (group_size >= 128)
        mult = 1;
    else if (group_size == 64)
        mult = 2;
    else if (group_size == 32)
        mult = 4;
    else
        return -1;
    int base = (in_features / group_size + 7) / 8;
    base = (base + mult - 1) / mult * mult;
    return base;
}

/* ------------------------------------------------------------------------------------------------
 * W4A16, QM_CUDA layout (llm/tools/quantize_methods.py:370-442):
 *   w      uint32[OC][IC/8]        nibble i of word w = weight ic = 8w+i
 *   zeros  uint32[OC][zeros_w]     nibble g = zero point of group g
 *   scales half  [OC][zeros_w*8]
 * Arithmetic of gemv_kernel_g128 (kernels/cuda/gemv_cuda.cu:179-185): deq = s * (q - z) in fp32,
 * psum += deq * x in fp32; k-order is serial here as in naive_mat_mul_int4's generic branch
 * (kernels/matmul_int4.cc:106-127).  Output fp32 (y) and its fp16 rounding (y_half, may be NULL) as
 * stored by gemv_cuda.cu:191-193.
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_w4a16_gemv(const uint16_t *x, const uint32_t *w, const uint32_t *zeros, const uint16_t *scales,
                           float *y, uint16_t *y_half, int M, int IC, int OC, int group) {
    int zeros_w = orc_calculate_zeros_width(IC, group);
    if (zeros_w < 0 || IC % group || IC % 8) return -1;
    int sf_w = zeros_w * 8;
    int wpr = IC / 8;
    for (int m = 0; m < M; m++) {
        for (int oc = 0; oc < OC; oc++) {
            float acc = 0.f;
            for (int ic = 0; ic < IC; ic++) {
                int g = ic / group;
                float s = h2f(scales[(size_t)oc * sf_w + g]);
                float z = (float)((zeros[(size_t)oc * zeros_w + g / 8] >> (4 * (g % 8))) & 0xF);
                float q = (float)((w[(size_t)oc * wpr + ic / 8] >> (4 * (ic % 8))) & 0xF);
                float deq = s * (q - z);
                acc += deq * h2f(x[(size_t)m * IC + ic]);
            }
            y[(size_t)m * OC + oc] = acc;
            if (y_half) y_half[(size_t)m * OC + oc] = f2h(acc);
        }
    }
    return 0;
}

/* kernels/matmul_int4.cc:106-127 (generic branch: sequential nibbles, scalar zero point, any block). */
ORC_API void orc_naive_mat_mul_int4(const float *A, const uint8_t *B, const float *scales, float zero_point,
                                    float *C, int M, int IC, int OC, int block_size) {
    int bcol = IC / 2;
    for (int i = 0; i < M; i++) {
        for (int j = 0; j < OC; j++) {
            float acc = 0;
            for (int k = 0; k < IC; k += block_size) {
                float s = scales[((size_t)j * bcol * 2 + k) / block_size];
                float z = zero_point;
                const float *a = &A[(size_t)i * IC + k];
                const uint8_t *b = &B[(size_t)j * bcol + k / 2];
                for (int qi = 0; qi < block_size / 2; qi++) {
                    uint8_t p = b[qi];
                    float deq_0 = ((float)(p & 0x0F) - z) * s;
                    float deq_1 = ((float)(p >> 4) - z) * s;
                    acc += *a++ * deq_0;
                    acc += *a++ * deq_1;
                }
            }
            C[(size_t)i * OC + j] = acc;
        }
    }
}

/* kernels/matmul_int4.cc:133-166 */
ORC_API void orc_naive_mat_mul_int4_with_offset(const float *A, const uint8_t *B, const float *scales,
                                                const float *offset, float zero_point, float *C, int M, int IC,
                                                int OC, int block_size) {
    int bcol = IC / 2;
    for (int i = 0; i < M; i++) {
        for (int j = 0; j < OC; j++) {
            float acc = 0;
            for (int k = 0; k < IC; k += block_size) {
                float s = scales[((size_t)j * bcol * 2 + k) / block_size];
                float o = offset[((size_t)j * bcol * 2 + k) / block_size];
                float z = zero_point;
                const float *a = &A[(size_t)i * IC + k];
                const uint8_t *b = &B[(size_t)j * bcol + k / 2];
                for (int qi = 0; qi < block_size / 2; qi++) {
                    uint8_t p = b[qi];
                    float deq_0 = ((float)(p & 0x0F) - z) * s + o;
                    float deq_1 = ((float)(p >> 4) - z) * s + o;
                    acc += *a++ * deq_0;
                    acc += *a++ * deq_1;
                }
            }
            C[(size_t)i * OC + j] = acc;
        }
    }
}

/* kernels/ref/matmul_ref_int4.cc:11-38 (legacy format: deq = q*s + offset, block 32). */
ORC_API void orc_int4_fast_ref(const float *A, const uint8_t *B, const float *scale, const float *offset, float *C,
                               int M, int IC, int OC) {
    int brow = IC / 2; /* params->B.row at this call site is IC/2 bytes */
    for (int i = 0; i < M; i++) {
        for (int j = 0; j < OC; j++) {
            float acc = 0;
            for (int k = 0; k < brow; k += 32) {
                /* NOTE: the reference iterates k over B->row (= IC/2) in steps of block_size, reading 16 bytes
                 * (32 weights) per step, so it consumes only the first half of each weight row; that quirk is
                 * restated as-is. */
                float s = scale[(size_t)j * (brow / 16) + k / 32];
                float o = offset[(size_t)j * (brow / 16) + k / 32];
                const uint8_t *wp = &B[(size_t)j * brow + k / 2];
                const float *xp = &A[(size_t)i * IC + k];
                for (int qi = 0; qi < 16; qi++) {
                    uint8_t p = wp[qi];
                    float deq_0 = (float)(p & 0x0F) * s + o;
                    float deq_1 = (float)(p >> 4) * s + o;
                    acc += *xp++ * deq_0;
                    acc += *xp++ * deq_1;
                }
            }
            C[(size_t)i * OC + j] = acc;
        }
    }
}

/* kernels/cuda/matmul_int4.cu:8-48 -- host fp16 reference, AWQ-GEMM layout B int32[IC][OC/8] with nibble
 * order 0 2 4 6 1 3 5 7, scales half[IC/G][OC], zero fixed 8, *fp16 accumulate* (half_float::half ops:
 * every binary op is evaluated in float then rounded to half, round-to-nearest). */
ORC_API void orc_naive_mat_mul_fp16_int4(const uint16_t *A, const int32_t *B, const uint16_t *scales, uint16_t *C,
                                         int M, int IC, int OC, int block_size) {
    static const int shift_of[8] = {0, 16, 4, 20, 8, 24, 12, 28};
    int bcol = OC / 8;
    for (int i = 0; i < M; i++) {
        for (int j = 0; j < OC; j++) {
            uint16_t acc = f2h(0.0f);
            for (int k = 0; k < IC; k++) {
                float s = h2f(scales[(size_t)(k / block_size) * OC + j]);
                float z = 8.0f;
                float in = h2f(A[(size_t)i * IC + k]);
                uint32_t word = (uint32_t)B[(size_t)k * bcol + j / 8];
                float q = h2f(f2h((float)((word >> shift_of[j % 8]) & 0xF))); /* (naive_float16_t)(int) */
                float d = h2f(f2h(q - z));                                   /* half - half */
                float wv = h2f(f2h(d * s));                                  /* half * half */
                float prod = h2f(f2h(in * wv));
                acc = f2h(h2f(acc) + prod);
            }
            C[(size_t)i * OC + j] = acc;
        }
    }
}

/* kernels/matmul_imp.cc:23-35 (mat_mul_transposed) and kernels/ref/matmul_ref_fp32.cc:11-28. */
ORC_API void orc_mat_mul_transposed(const float *A, const float *B, float *C, int M, int N, int K) {
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            float acc = 0;
            for (int k = 0; k < K; k++) acc += A[(size_t)i * K + k] * B[(size_t)j * K + k];
            C[(size_t)i * N + j] = acc;
        }
}

/* ------------------------------------------------------------------------------------------------
 * W8A8 (SmoothQuant) family, kernels/ref/matmul_ref_int8.cc.  B is int8[N][K] (torch [out,in]).
 * std::round = half away from zero = roundf.  Epilogue: (float)acc * alpha + (float)bias * beta, written
 * exactly in that order (matmul_ref_int8.cc:28).
 * ---------------------------------------------------------------------------------------------- */
static inline int32_t dot_s8(const int8_t *a, const int8_t *b, int k) {
    int32_t acc = 0;
    for (int kk = 0; kk < k; kk++) acc += (int32_t)a[kk] * (int32_t)b[kk];
    return acc;
}
static inline int8_t clamp_s8(int32_t v, int8_t q_min, int8_t q_max) {
    if (v < q_min) v = q_min;
    if (v > q_max) v = q_max;
    return (int8_t)v;
}

/* matmul_ref_int8.cc:11-35  int8_ref_matmul (bias int8, int8 out) */
ORC_API void orc_int8_matmul(const int8_t *A, const int8_t *B, const int8_t *bias, int8_t *C, int M, int N, int K,
                             float alpha, float beta, int q_min, int q_max) {
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            int32_t acc = dot_s8(&A[(size_t)i * K], &B[(size_t)j * K], K);
            float v = (float)acc * alpha;
            float bb = (float)bias[j] * beta;
            acc = (int32_t)roundf(v + bb);
            C[(size_t)i * N + j] = clamp_s8(acc, (int8_t)q_min, (int8_t)q_max);
        }
}

/* matmul_ref_int8.cc:37-61  int8_ref_matmul_nobias */
ORC_API void orc_int8_matmul_nobias(const int8_t *A, const int8_t *B, int8_t *C, int M, int N, int K, float alpha,
                                    int q_min, int q_max) {
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            int32_t acc = dot_s8(&A[(size_t)i * K], &B[(size_t)j * K], K);
            acc = (int32_t)roundf((float)acc * alpha);
            C[(size_t)i * N + j] = clamp_s8(acc, (int8_t)q_min, (int8_t)q_max);
        }
}

/* matmul_ref_int8.cc:63-87  int8_ref_matmul_nobias_batch: row i of A uses its own B slab B[i][N][K] */
ORC_API void orc_int8_matmul_nobias_batch(const int8_t *A, const int8_t *B, int8_t *C, int M, int N, int K,
                                          float alpha, int q_min, int q_max) {
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            int32_t acc = dot_s8(&A[(size_t)i * K], &B[(size_t)i * K * N + (size_t)j * K], K);
            acc = (int32_t)roundf((float)acc * alpha);
            C[(size_t)i * N + j] = clamp_s8(acc, (int8_t)q_min, (int8_t)q_max);
        }
}

/* matmul_ref_int8.cc:89-111  int8_ref_matmul_bfp32_ofp32 */
ORC_API void orc_int8_matmul_bfp32_ofp32(const int8_t *A, const int8_t *B, const float *bias, float *C, int M,
                                         int N, int K, float alpha) {
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            int32_t acc = dot_s8(&A[(size_t)i * K], &B[(size_t)j * K], K);
            float v = (float)acc * alpha;
            C[(size_t)i * N + j] = v + bias[j];
        }
}

/* matmul_ref_int8.cc:113-135  int8_ref_matmul_nobias_ofp32 */
ORC_API void orc_int8_matmul_nobias_ofp32(const int8_t *A, const int8_t *B, float *C, int M, int N, int K,
                                          float alpha) {
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            int32_t acc = dot_s8(&A[(size_t)i * K], &B[(size_t)j * K], K);
            C[(size_t)i * N + j] = (float)acc * alpha;
        }
}

/* matmul_ref_int8.cc:137-159  int8_ref_matmul_nobias_ofp32_batch */
ORC_API void orc_int8_matmul_nobias_ofp32_batch(const int8_t *A, const int8_t *B, float *C, int M, int N, int K,
                                                float alpha) {
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            int32_t acc = dot_s8(&A[(size_t)i * K], &B[(size_t)i * K * N + (size_t)j * K], K);
            C[(size_t)i * N + j] = (float)acc * alpha;
        }
}

/* kernels/matmul_int8.cc:8-30  naive_mat_mul_int8 (gemmlowp style; B is [K][N], truncating cast). */
ORC_API void orc_naive_mat_mul_int8(const int8_t *A, const int8_t *B, int8_t *C, int M, int N, int K, int32_t A_zp,
                                    int32_t C_zp, float A_sc, float B_sc, float C_sc, int q_min, int q_max) {
    float effective_scale = A_sc * B_sc / C_sc;
    for (int i = 0; i < M; i++)
        for (int j = 0; j < N; j++) {
            int32_t acc = 0;
            for (int k = 0; k < K; k++) acc += ((int32_t)A[(size_t)i * K + k] - A_zp) * B[(size_t)k * N + j];
            acc = (int32_t)((float)acc * effective_scale);
            acc -= C_zp;
            C[(size_t)i * N + j] = clamp_s8(acc, (int8_t)q_min, (int8_t)q_max);
        }
}

/* ------------------------------------------------------------------------------------------------
 * Small fp32 ops used on either side of the path (oracle for the fused decode block).
 * ---------------------------------------------------------------------------------------------- */

/* llm/src/ops/LlamaRMSNorm.cc:7-36 */
ORC_API void orc_rmsnorm(const float *x, const float *weight, float *out, int rows, int dim, float eps) {
    for (int j = 0; j < rows; j++) {
        float var = 0;
        for (int k = 0; k < dim; k++) var += x[(size_t)j * dim + k] * x[(size_t)j * dim + k];
        var /= (float)dim;
        /* `1.0 / sqrt(var + eps)` with a float argument: C++ picks the float overload of sqrt, the division is in double */
        float variance = (float)(1.0 / (double)sqrtf(var + eps));
        for (int k = 0; k < dim; k++) {
            float value = x[(size_t)j * dim + k];
            out[(size_t)j * dim + k] = (value * variance) * weight[k];
        }
    }
}

/* llm/src/ops/LayerNormQ.cc:12-52 (fp32 -> int8, std::round, wrapping static_cast<int8_t>) */
ORC_API void orc_layernorm_q(const float *x, const float *weight, const float *bias, int8_t *out, int rows,
                             int dim) {
    const float eps = 0.00001f;
    for (int j = 0; j < rows; j++) {
        const float *xr = &x[(size_t)j * dim];
        float mean = 0;
        for (int k = 0; k < dim; k++) mean += xr[k];
        mean /= (float)dim;
        float sq = 0;
        for (int k = 0; k < dim; k++) sq += (xr[k] - mean) * (xr[k] - mean);
        float var = sq / (float)dim;
        float std_dev = sqrtf(var + eps);
        for (int k = 0; k < dim; k++) {
            float fp_out = ((xr[k] - mean) / std_dev * weight[k]) + bias[k];
            out[(size_t)j * dim + k] = (int8_t)(int32_t)roundf(fp_out);
        }
    }
}

/* llm/src/ops/softmax.cc:5-41 (dim 2).  NOTE the reference seeds max with input.m_data[0] (first element of
 * the whole tensor, softmax.cc:13), restated as-is via `seed`; the attention modules call it in place, see the callers. */
static void softmax_row(const float *in, float *out, int n, float seed) {
    float max_value = seed;
    float sum = 0;
    for (int k = 0; k < n; k++)
        if (in[k] > max_value) max_value = in[k];
    for (int k = 0; k < n; k++) sum += expf(in[k] - max_value);
    for (int k = 0; k < n; k++) out[k] = (float)(expf(in[k] - max_value) / (sum + 1e-10));
}

/* llm/src/ops/RotaryPosEmb.cc:7-69: rotate-half with cos/sin tables [max_sqlen][head_dim]. */
static void rope_rows(float *t, int heads, int len, int hd, const float *cosb, const float *sinb, int start) {
    float buf[512];
    int half = hd / 2;
    for (int b = 0; b < heads; b++)
        for (int i = 0; i < len; i++) {
            float *row = &t[((size_t)b * len + i) * hd];
            for (int j = 0; j < half; j++) buf[j] = -1 * row[j + half];
            for (int j = half; j < hd; j++) buf[j] = row[j - half];
            const float *c = &cosb[(size_t)(i + start) * hd], *s = &sinb[(size_t)(i + start) * hd];
            for (int j = 0; j < hd; j++) row[j] = ((row[j] * c[j]) + (buf[j] * s[j]));
        }
}
ORC_API void orc_rope(float *q, float *k, int num_heads, int num_kv_heads, int len, int head_dim, const float *cosb,
                      const float *sinb, int start_idx) {
    rope_rows(q, num_heads, len, head_dim, cosb, sinb, start_idx);
    rope_rows(k, num_kv_heads, len, head_dim, cosb, sinb, start_idx);
}

/* ------------------------------------------------------------------------------------------------
 * fp32 GQA attention core of Int4llamaAttention::forward (llm/src/nn_modules/non_cuda/
 * Int4llamaAttention.cc:288-442) between the q/k/v projections and o_proj:
 *   inputs  q [sqlen][H*hd], k_new/v_new [sqlen][KVH*hd] (projection outputs, "unshape" layout)
 *           past_k/past_v [KVH][past][hd]
 *   shape -> RoPE(q, k_new; start = past) -> concat -> repeat (q head i uses kv head i / (H/KVH), :166-184)
 *   -> S = alpha * Q K^T (BMM_F32T.cc:39-42) -> + mask [sqlen][tgz] (batch_add.cc) -> isinf -> lowest
 *   -> softmax (softmax.cc) -> P V (BMM_F32T.cc:95-108 k-outer order) -> unshape [sqlen][H*hd]
 *   outputs attn_out [sqlen][H*hd], final_k/final_v [KVH][tgz][hd]
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_llama_attention_core(const float *q_in, const float *k_in, const float *v_in, const float *past_k,
                                     const float *past_v, const float *mask, const float *cosb, const float *sinb,
                                     float alpha, int sqlen, int past, int H, int KVH, int hd, float *attn_out,
                                     float *final_k, float *final_v) {
    int tgz = sqlen + past;
    int n_rep = H / KVH;
    float *q = (float *)malloc(sizeof(float) * H * sqlen * hd);
    float *k = (float *)malloc(sizeof(float) * KVH * sqlen * hd);
    float *v = (float *)malloc(sizeof(float) * KVH * sqlen * hd);
    float *S = (float *)malloc(sizeof(float) * (size_t)H * sqlen * tgz);
    float *O = (float *)malloc(sizeof(float) * H * sqlen * hd);
    if (!q || !k || !v || !S || !O) return -1;
    /* shape: [s][h*hd] -> [h][s][hd]  (Int4llamaAttention.cc:128-147) */
    for (int i = 0; i < H; i++)
        for (int j = 0; j < sqlen; j++)
            for (int d = 0; d < hd; d++) q[((size_t)i * sqlen + j) * hd + d] = q_in[(size_t)j * H * hd + i * hd + d];
    for (int i = 0; i < KVH; i++)
        for (int j = 0; j < sqlen; j++)
            for (int d = 0; d < hd; d++) {
                k[((size_t)i * sqlen + j) * hd + d] = k_in[(size_t)j * KVH * hd + i * hd + d];
                v[((size_t)i * sqlen + j) * hd + d] = v_in[(size_t)j * KVH * hd + i * hd + d];
            }
    orc_rope(q, k, H, KVH, sqlen, hd, cosb, sinb, past);
    /* concat with the past (Int4llamaAttention.cc:363-387) */
    for (int i = 0; i < KVH; i++) {
        if (past > 0) {
            memcpy(&final_k[(size_t)i * tgz * hd], &past_k[(size_t)i * past * hd], sizeof(float) * past * hd);
            memcpy(&final_v[(size_t)i * tgz * hd], &past_v[(size_t)i * past * hd], sizeof(float) * past * hd);
        }
        memcpy(&final_k[((size_t)i * tgz + past) * hd], &k[(size_t)i * sqlen * hd], sizeof(float) * sqlen * hd);
        memcpy(&final_v[((size_t)i * tgz + past) * hd], &v[(size_t)i * sqlen * hd], sizeof(float) * sqlen * hd);
    }
    for (int h = 0; h < H; h++) {
        const float *K = &final_k[(size_t)(h / n_rep) * tgz * hd];
        const float *V = &final_v[(size_t)(h / n_rep) * tgz * hd];
        float *Sh = &S[(size_t)h * sqlen * tgz];
        for (int i = 0; i < sqlen; i++)
            for (int j = 0; j < tgz; j++) {
                float acc = 0;
                for (int d = 0; d < hd; d++) acc += q[((size_t)h * sqlen + i) * hd + d] * K[(size_t)j * hd + d];
                Sh[(size_t)i * tgz + j] = acc;
            }
        for (int i = 0; i < sqlen * tgz; i++) Sh[i] *= alpha;
    }
    float lowest = -3.402823466e+38f;
    for (size_t h = 0; h < (size_t)H; h++)
        for (int i = 0; i < sqlen; i++)
            for (int j = 0; j < tgz; j++) {
                float *p = &S[(h * sqlen + i) * tgz + j];
                *p = *p + mask[(size_t)i * tgz + j];
                if (isinf(*p)) *p = lowest;
            }
    /* the module runs softmax IN PLACE (attn_probs aliases attn_weights_arr), so `input.m_data[0]` read at the top of
     * each row (softmax.cc:13) is the raw score only for the first row and p[0][0][0] afterwards */
    for (size_t r = 0; r < (size_t)H * sqlen; r++) softmax_row(&S[r * tgz], &S[r * tgz], tgz, S[0]);
    memset(O, 0, sizeof(float) * H * sqlen * hd);
    for (int h = 0; h < H; h++) {
        const float *V = &final_v[(size_t)(h / n_rep) * tgz * hd];
        for (int i = 0; i < sqlen; i++)
            for (int kk = 0; kk < tgz; kk++) {
                float a = S[((size_t)h * sqlen + i) * tgz + kk];
                for (int d = 0; d < hd; d++) O[((size_t)h * sqlen + i) * hd + d] += a * V[(size_t)kk * hd + d];
            }
    }
    /* unshape (Int4llamaAttention.cc:149-164) */
    for (int i = 0; i < H; i++)
        for (int j = 0; j < sqlen; j++)
            for (int d = 0; d < hd; d++) attn_out[(size_t)j * H * hd + i * hd + d] = O[((size_t)i * sqlen + j) * hd + d];
    free(q);
    free(k);
    free(v);
    free(S);
    free(O);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * int8 attention core of Int8OPTAttention::forward (llm/src/nn_modules/Int8OPTAttention.cc:183-284) between
 * the int8 q/k/v projections and out_proj:
 *   q8,k8,v8 int8 [sqlen][H*hd]; past_k/past_v int8 [H][past][hd]
 *   S = (float)acc * qk_alpha (BMM_S8T_S8N_F32T) + mask -> fp32 softmax -> P8 = (int8)round(p*127) (:262)
 *   -> O8 = clamp(round(acc * pv_alpha)) (BMM_S8T_S8N_S8T) -> unshape int8 [sqlen][H*hd]
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_opt_int8_attention_core(const int8_t *q8, const int8_t *k8, const int8_t *v8, const int8_t *past_k,
                                        const int8_t *past_v, const float *mask, float qk_alpha, float pv_alpha,
                                        int sqlen, int past, int H, int hd, int8_t *attn_out, int8_t *final_k,
                                        int8_t *final_v) {
    int tgz = sqlen + past;
    float *S = (float *)malloc(sizeof(float) * (size_t)H * sqlen * tgz);
    int8_t *P = (int8_t *)malloc((size_t)H * sqlen * tgz);
    if (!S || !P) return -1;
    for (int i = 0; i < H; i++) {
        if (past > 0) {
            memcpy(&final_k[(size_t)i * tgz * hd], &past_k[(size_t)i * past * hd], (size_t)past * hd);
            memcpy(&final_v[(size_t)i * tgz * hd], &past_v[(size_t)i * past * hd], (size_t)past * hd);
        }
        for (int j = 0; j < sqlen; j++)
            for (int d = 0; d < hd; d++) {
                final_k[((size_t)i * tgz + past + j) * hd + d] = k8[(size_t)j * H * hd + i * hd + d];
                final_v[((size_t)i * tgz + past + j) * hd + d] = v8[(size_t)j * H * hd + i * hd + d];
            }
    }
    for (int h = 0; h < H; h++)
        for (int i = 0; i < sqlen; i++)
            for (int j = 0; j < tgz; j++) {
                int32_t acc = 0;
                for (int d = 0; d < hd; d++)
                    acc += (int32_t)q8[(size_t)i * H * hd + h * hd + d] * (int32_t)final_k[((size_t)h * tgz + j) * hd + d];
                float s = (float)acc * qk_alpha;
                S[((size_t)h * sqlen + i) * tgz + j] = s + mask[(size_t)i * tgz + j];
            }
    /* the module runs softmax IN PLACE (attn_probs aliases attn_weights_arr), so `input.m_data[0]` read at the top of
     * each row (softmax.cc:13) is the raw score only for the first row and p[0][0][0] afterwards */
    for (size_t r = 0; r < (size_t)H * sqlen; r++) softmax_row(&S[r * tgz], &S[r * tgz], tgz, S[0]);
    for (size_t i = 0; i < (size_t)H * sqlen * tgz; i++) P[i] = (int8_t)(int32_t)roundf(S[i] * 127);
    for (int h = 0; h < H; h++)
        for (int i = 0; i < sqlen; i++)
            for (int d = 0; d < hd; d++) {
                int32_t acc = 0;
                for (int t = 0; t < tgz; t++)
                    acc += (int32_t)P[((size_t)h * sqlen + i) * tgz + t] * (int32_t)final_v[((size_t)h * tgz + t) * hd + d];
                acc = (int32_t)roundf((float)acc * pv_alpha);
                attn_out[(size_t)i * H * hd + h * hd + d] = clamp_s8(acc, -128, 127);
            }
    free(S);
    free(P);
    return 0;
}
