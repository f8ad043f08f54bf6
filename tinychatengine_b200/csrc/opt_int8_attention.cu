// opt_int8_attention.cu -- the int8 attention core of Int8OPTAttention::forward (llm/src/nn_modules/Int8OPTAttention.cc:183-284)
// between the q/k/v projections and out_proj, as two kernels:
//   1. kv_concat: final_k/final_v[h][0..tgz) = past rows (skipped when the cache is updated in place) ++ this call's rows
//      (the reference's shape() + cat_past_keys_values memcpy, :207-243).
//   2. attn_rows: one CTA per (head, query row): scores (dp4a, exact int32) * qk_alpha + mask -> fp32 softmax -> P8 =
//      round(p*127) -> P8 x V (exact int32) -> clamp(round(acc * pv_alpha)) -> unshape.  Replaces BMM_S8T_S8N_F32T, batch_Add,
//      softmax, the int8 cast loop, transpose_1_2idx, BMM_S8T_S8N_S8T and unshape (:245-275) with no intermediate in HBM.
// Bit-exactness with the CPU reference is the contract: integer parts are exact in any order; the float parts keep the
// reference's evaluation order (serial float sum over the row; p = float(double(e) / (double(sum) + 1e-10))), and exp() is
// evaluated in double and rounded to float, which is what a correctly-rounded expf returns.  The running max of every row
// is seeded with element [0] of the score tensor (softmax.cc:13); the module runs softmax in place, so that element is the
// raw score for row (head 0, query 0) and the PROBABILITY p[0][0][0] for every later row.  Reproduced literally: row (0,0)
// runs first in a one-CTA launch that publishes p[0][0][0], then all other rows run seeded with it.
#include "common.cuh"
#include "kernels_w8a8.h"

namespace tce {
namespace {

constexpr int kThreads = 128;

__global__ void opt_kv_concat_kernel(const int8_t *__restrict__ k8, const int8_t *__restrict__ v8, const int8_t *__restrict__ past_k,
                                     const int8_t *__restrict__ past_v, long long past_hs, int8_t *__restrict__ final_k, int8_t *__restrict__ final_v,
                                     long long final_hs, int sqlen, int past, int H, int hd, int copy_past) {
    const int h = blockIdx.y;
    const int row0 = copy_past ? 0 : past, tgz = past + sqlen;
    const long long n = (long long)(tgz - row0) * hd;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int t = row0 + (int)(e / hd), d = (int)(e % hd);
        int8_t kk, vv;
        if (t < past) {
            kk = past_k[h * past_hs + (long long)t * hd + d];
            vv = past_v[h * past_hs + (long long)t * hd + d];
        } else {
            kk = k8[(size_t)(t - past) * H * hd + h * hd + d];
            vv = v8[(size_t)(t - past) * H * hd + h * hd + d];
        }
        final_k[h * final_hs + (long long)t * hd + d] = kk;
        final_v[h * final_hs + (long long)t * hd + d] = vv;
    }
}

TCE_DEVINL int dot_s8(const int8_t *__restrict__ a, const int8_t *__restrict__ b, int n) {  // n % 4 == 0, both 4-byte aligned
    int acc = 0;
    for (int d = 0; d < n; d += 4) acc = __dp4a(*(const int *)(a + d), *(const int *)(b + d), acc);
    return acc;
}

// dynamic smem: float s[tgz] | int8 p8[tgz rounded to 4] | int8 q[hd] | int acc[hd]
__global__ void __launch_bounds__(kThreads) opt_attn_rows_kernel(const int8_t *__restrict__ q8, const int8_t *__restrict__ final_k,
                                                                const int8_t *__restrict__ final_v, long long hs, const float *__restrict__ mask,
                                                                float qk_alpha, float pv_alpha, int sqlen, int past, int H, int hd,
                                                                int8_t *__restrict__ out, float *seed_ws, int first_row_only) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int h = blockIdx.x, i = blockIdx.y, tgz = past + sqlen, tid = threadIdx.x;
    if (!first_row_only && h == 0 && i == 0) return;  // done by the seed launch
    float *s = (float *)smem_raw;
    int8_t *p8 = (int8_t *)(s + tgz);
    int8_t *q = p8 + ((tgz + 15) & ~15);
    int *acc = (int *)(q + ((hd + 15) & ~15));
    __shared__ float red[kThreads / 32];
    __shared__ float s_sum;

    for (int d = tid; d < hd; d += kThreads) {
        q[d] = q8[(size_t)i * H * hd + h * hd + d];
        acc[d] = 0;
    }
    __syncthreads();
    const int8_t *K = final_k + h * hs, *V = final_v + h * hs;
    const float neg = -3.402823466e38f;
    // softmax.cc:13: row (0,0) is seeded with its own first score (a member of the row: plain max), later rows with p[0][0][0]
    float mx = first_row_only ? -INFINITY : *seed_ws;
    for (int j = tid; j < tgz; j += kThreads) {
        const int a = dot_s8(q, K + (size_t)j * hd, hd);
        const float m = mask ? mask[(size_t)i * tgz + j] : (j > past + i ? neg : 0.f);
        const float v = __fadd_rn(__fmul_rn((float)a, qk_alpha), m);
        s[j] = v;
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < kThreads / 32; w++) mx = fmaxf(mx, red[w]);
    for (int j = tid; j < tgz; j += kThreads) s[j] = (float)exp((double)__fsub_rn(s[j], mx));
    __syncthreads();
    if (tid == 0) {  // the reference's serial float sum (order matters for bit-exactness)
        float sum = 0.f;
        for (int j = 0; j < tgz; j++) sum = __fadd_rn(sum, s[j]);
        s_sum = sum;
    }
    __syncthreads();
    const double denom = (double)s_sum + 1e-10;
    for (int j = tid; j < tgz; j += kThreads) {
        const float p = (float)((double)s[j] / denom);
        p8[j] = (int8_t)(int)roundf(__fmul_rn(p, 127.f));
        if (first_row_only && j == 0) *seed_ws = p;
    }
    __syncthreads();
    // P8 x V: thread (g, c) walks rows t = g, g+G, ... and owns 4 consecutive d's
    const int lanes = hd / 4, G = kThreads / lanes;  // hd % 4 == 0, lanes <= kThreads checked by the launcher
    const int c = tid % lanes, g = tid / lanes;
    if (g < G) {
        int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int t = g; t < tgz; t += G) {
            const int p = p8[t];
            const int v4 = *(const int *)(V + (size_t)t * hd + c * 4);
            a0 += p * (int)(int8_t)(v4 & 0xff);
            a1 += p * (int)(int8_t)((v4 >> 8) & 0xff);
            a2 += p * (int)(int8_t)((v4 >> 16) & 0xff);
            a3 += p * (int)(int8_t)((v4 >> 24) & 0xff);
        }
        atomicAdd(&acc[c * 4 + 0], a0);
        atomicAdd(&acc[c * 4 + 1], a1);
        atomicAdd(&acc[c * 4 + 2], a2);
        atomicAdd(&acc[c * 4 + 3], a3);
    }
    __syncthreads();
    for (int d = tid; d < hd; d += kThreads) {
        int r = (int)roundf(__fmul_rn((float)acc[d], pv_alpha));
        r = max(-128, min(127, r));
        out[(size_t)i * H * hd + h * hd + d] = (int8_t)r;
    }
}

}  // namespace

cudaError_t launch_opt_int8_attention(Ctx *ctx, const OptAttnParams &p) {
    const int tgz = p.past + p.sqlen;
    const bool in_place = (p.past_k == p.final_k && p.past_v == p.final_v && p.past_hs == p.final_hs);
    const int copy_past = (p.past > 0 && !in_place) ? 1 : 0;
    const long long n = (long long)(copy_past ? tgz : p.sqlen) * p.hd;
    const long long nb = (n + 255) / 256;
    dim3 g1((unsigned)(nb < 1024 ? nb : 1024), p.H);
    opt_kv_concat_kernel<<<g1, 256, 0, ctx->stream>>>(p.k8, p.v8, p.past_k, p.past_v, p.past_hs, p.final_k, p.final_v, p.final_hs, p.sqlen, p.past, p.H,
                                                       p.hd, copy_past);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t smem = (size_t)tgz * 4 + ((tgz + 15) & ~15) + ((p.hd + 15) & ~15) + (size_t)p.hd * 4;
    if (smem > 48 * 1024) {
        if (smem > (size_t)ctx->smem_optin) return cudaErrorInvalidValue;
        e = cudaFuncSetAttribute(opt_attn_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    float *seed_ws = ctx->attn_ws;  // one float of the attention workspace; same-stream ordering makes the reuse safe
    opt_attn_rows_kernel<<<dim3(1, 1), kThreads, smem, ctx->stream>>>(p.q8, p.final_k, p.final_v, p.final_hs, p.mask, p.qk_alpha, p.pv_alpha, p.sqlen, p.past,
                                                                      p.H, p.hd, p.out, seed_ws, 1);
    e = cudaGetLastError();
    if (e != cudaSuccess || (p.H == 1 && p.sqlen == 1)) return e;
    dim3 g2(p.H, p.sqlen);
    opt_attn_rows_kernel<<<g2, kThreads, smem, ctx->stream>>>(p.q8, p.final_k, p.final_v, p.final_hs, p.mask, p.qk_alpha, p.pv_alpha, p.sqlen, p.past, p.H,
                                                             p.hd, p.out, seed_ws, 0);
    return cudaGetLastError();
}

}  // namespace tce
