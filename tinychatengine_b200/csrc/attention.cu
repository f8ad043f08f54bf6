// attention.cu -- per-token KV-cache attention for decode (fp16 cache, GQA aware) on sm_100a.
//
// Replaces the attention part of Int4llamaAttention::forward between qkv_proj and o_proj
// (reference llm/src/nn_modules/cuda/Int4llamaAttention.cu:128-217: shape_qkv -> RotaryPosEmb_cuda_forward ->
// 4*heads cudaMemcpyAsync KV concat -> BMM_F16T QK^T -> batch_Add mask -> check_inf -> softmax_cuda ->
// transpose V -> BMM_F16T PV -> unshape) with ONE kernel, with the GQA head mapping of the CPU module
// (llm/src/nn_modules/non_cuda/Int4llamaAttention.cc:166-184: q head i reads kv head i / (H/KVH)).
//
//  * KV cache is [KVH][max_ctx][128] fp16 per layer, appended IN PLACE at `pos` (no ping-pong copy, no V
//    transpose);
//  * grid = (KVH, ceil(max_ctx/chunk)): a CTA owns `chunk` cached positions of one KV head and serves the
//    H/KVH query heads that share it, so each cached byte is read once per token;
//  * its K and V slabs (contiguous chunk*256 B each) are pulled into shared memory by two TMA bulk copies
//    issued up front -- the whole cache of the layer is in flight at once, the kernel is HBM-bound;
//  * fp32 RoPE / scores / softmax statistics / PV accumulation; split results are merged flash-decoding
//    style by the last CTA of each KV head to arrive (fixed order, deterministic).
#include "common.cuh"
#include "kernels.h"
#include "kernels_attn.h"

namespace tce {
namespace {

constexpr int HD = 128;          // head_dim (every Llama config in llm/include/model.h:71-83)
constexpr int kAttnThreads = 256;
constexpr int kMaxRep = 8;

template <int NREP>
__global__ void __launch_bounds__(kAttnThreads, 1) attn_decode_kernel(const __grid_constant__ AttnDecodeArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int chunk = a.chunk;
    __half *sK = reinterpret_cast<__half *>(smem);
    __half *sV = sK + (size_t)chunk * HD;
    float *sQ = reinterpret_cast<float *>(sV + (size_t)chunk * HD);  // [NREP][HD] rotated * alpha
    float *sP = sQ + NREP * HD;                                        // [NREP][chunk] scores -> probabilities
    float *sRed = sP + NREP * chunk;                                   // [16][NREP][HD] PV partials
    float *sStat = sRed + 16 * NREP * HD;                              // [NREP][2 * 8 warps] max/sum scratch
    uint64_t *bar = reinterpret_cast<uint64_t *>(sStat + NREP * 16);
    int *flag = reinterpret_cast<int *>(bar + 2);

    const int kvh = blockIdx.x, split = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    pdl_launch_dependents();
    pdl_wait();  // qkv of this token comes from the previous kernel; the cache was written by earlier steps

    const int pos = *a.pos;       // index of the token being decoded
    const int T = pos + 1;        // visible positions
    const int t0 = split * chunk;
    if (t0 >= T) return;          // empty split
    const int t1 = min(T, t0 + chunk);
    const int nrows = t1 - t0;
    const bool owns_new = (pos >= t0 && pos < t1);
    const int ncached = owns_new ? nrows - 1 : nrows;  // rows that already live in the cache

    __half *Kc = a.k_cache + ((size_t)kvh * a.max_ctx) * HD;
    __half *Vc = a.v_cache + ((size_t)kvh * a.max_ctx) * HD;
    if (tid == 0 && ncached > 0) {
        // default L2 priority: the (evict-first) weight stream then cannot push a short context's cache out of L2
        const uint32_t bytes = (uint32_t)ncached * HD * 2;
        mbar_arrive_expect_tx(&bar[0], bytes);
        bulk_g2s_nohint(sK, Kc + (size_t)t0 * HD, bytes, &bar[0]);
        mbar_arrive_expect_tx(&bar[1], bytes);
        bulk_g2s_nohint(sV, Vc + (size_t)t0 * HD, bytes, &bar[1]);
    }

    // ---- RoPE (llm/src/ops/RotaryPosEmb.cc:7-69 rotate-half) on the NREP query heads and, if this CTA owns
    //      the new position, on the new key; fp32 math, tables [max_ctx][HD] fp32 ----
    const float *cosr = a.cos + (size_t)pos * HD, *sinr = a.sin + (size_t)pos * HD;
    const int H = a.num_heads, KVH = a.num_kv_heads;
    for (int i = tid; i < NREP * HD; i += kAttnThreads) {
        const int r = i / HD, j = i % HD;
        const __half *q = a.qkv + (size_t)(kvh * NREP + r) * HD;
        const float x = __half2float(q[j]);
        const float xr = (j < HD / 2) ? -__half2float(q[j + HD / 2]) : __half2float(q[j - HD / 2]);
        sQ[i] = (x * cosr[j] + xr * sinr[j]) * a.alpha;
    }
    if (owns_new && tid < HD) {
        const int j = tid;
        const __half *k = a.qkv + (size_t)H * HD + (size_t)kvh * HD;
        const __half *v = a.qkv + (size_t)(H + KVH) * HD + (size_t)kvh * HD;
        const float x = __half2float(k[j]);
        const float xr = (j < HD / 2) ? -__half2float(k[j + HD / 2]) : __half2float(k[j - HD / 2]);
        const __half kh = __float2half(x * cosr[j] + xr * sinr[j]);
        sK[(size_t)(nrows - 1) * HD + j] = kh;  // row `pos` of the slab; the bulk copy never touches it
        sV[(size_t)(nrows - 1) * HD + j] = v[j];
        Kc[(size_t)pos * HD + j] = kh;          // in-place append
        Vc[(size_t)pos * HD + j] = v[j];
    }
    __syncthreads();

    // ---- scores: 16 lanes per cached row (8 dims each), 2 rows per warp instruction ----
    const int sub = lane & 15, rsel = lane >> 4;
    float qreg[NREP][8];
#pragma unroll
    for (int r = 0; r < NREP; r++)
#pragma unroll
        for (int d = 0; d < 8; d++) qreg[r][d] = sQ[r * HD + sub * 8 + d];
    if (ncached > 0) mbar_wait(&bar[0], 0);
    for (int row = warp * 2 + rsel; row < nrows + (nrows & 1); row += (kAttnThreads / 32) * 2) {
        float dot[NREP];
#pragma unroll
        for (int r = 0; r < NREP; r++) dot[r] = 0.f;
        if (row < nrows) {
            const uint4 kv = *reinterpret_cast<const uint4 *>(sK + (size_t)row * HD + sub * 8);
            const __half2 *k2 = reinterpret_cast<const __half2 *>(&kv);
#pragma unroll
            for (int d = 0; d < 4; d++) {
                const float2 f = __half22float2(k2[d]);
#pragma unroll
                for (int r = 0; r < NREP; r++) dot[r] += qreg[r][2 * d] * f.x + qreg[r][2 * d + 1] * f.y;
            }
        }
#pragma unroll
        for (int r = 0; r < NREP; r++) {
            float v = dot[r];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if (sub == 0 && row < nrows) sP[r * chunk + row] = v;
        }
    }
    __syncthreads();

    // ---- softmax statistics of this split (fp32): m = max, p = exp(s - m), l = sum p ----
    float m_loc[NREP], l_loc[NREP];
#pragma unroll
    for (int r = 0; r < NREP; r++) {
        float m = -INFINITY;
        for (int i = tid; i < nrows; i += kAttnThreads) m = fmaxf(m, sP[r * chunk + i]);
        m = warp_max(m);
        if (lane == 0) sStat[r * 16 + warp] = m;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NREP; r++) {
        float m = sStat[r * 16];
#pragma unroll
        for (int w = 1; w < kAttnThreads / 32; w++) m = fmaxf(m, sStat[r * 16 + w]);
        m_loc[r] = m;
        float l = 0.f;
        for (int i = tid; i < nrows; i += kAttnThreads) {
            const float p = __expf(sP[r * chunk + i] - m);
            sP[r * chunk + i] = p;
            l += p;
        }
        l = warp_sum(l);
        if (lane == 0) sStat[r * 16 + 8 + warp] = l;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NREP; r++) {
        float l = 0.f;
#pragma unroll
        for (int w = 0; w < kAttnThreads / 32; w++) l += sStat[r * 16 + 8 + w];
        l_loc[r] = l;
    }

    // ---- PV: 16 lanes per row again, each lane accumulates 8 dims for NREP heads over its rows ----
    float acc[NREP][8];
#pragma unroll
    for (int r = 0; r < NREP; r++)
#pragma unroll
        for (int d = 0; d < 8; d++) acc[r][d] = 0.f;
    if (ncached > 0) mbar_wait(&bar[1], 0);
    const int rg = warp * 2 + rsel;  // row group 0..15
    for (int row = rg; row < nrows; row += 16) {
        const uint4 vv = *reinterpret_cast<const uint4 *>(sV + (size_t)row * HD + sub * 8);
        const __half2 *v2 = reinterpret_cast<const __half2 *>(&vv);
        float vf[8];
#pragma unroll
        for (int d = 0; d < 4; d++) {
            const float2 f = __half22float2(v2[d]);
            vf[2 * d] = f.x;
            vf[2 * d + 1] = f.y;
        }
#pragma unroll
        for (int r = 0; r < NREP; r++) {
            const float p = sP[r * chunk + row];
#pragma unroll
            for (int d = 0; d < 8; d++) acc[r][d] += p * vf[d];
        }
    }
#pragma unroll
    for (int r = 0; r < NREP; r++)
#pragma unroll
        for (int d = 0; d < 8; d++) sRed[((size_t)rg * NREP + r) * HD + sub * 8 + d] = acc[r][d];
    __syncthreads();

    // ---- per-split result (unnormalised o, m, l) ----
    const int nsplit_active = (T + chunk - 1) / chunk;
    float *ws = a.ws;  // [H][nsplit_max][HD + 2]
    const int wstride = HD + 2;
    for (int i = tid; i < NREP * HD; i += kAttnThreads) {
        const int r = i / HD, d = i % HD;
        float o = 0.f;
#pragma unroll
        for (int gsel = 0; gsel < 16; gsel++) o += sRed[((size_t)gsel * NREP + r) * HD + d];
        const int head = kvh * NREP + r;
        if (nsplit_active == 1) {
            a.out[(size_t)head * HD + d] = __float2half(o / l_loc[r]);
        } else {
            float *rec = ws + ((size_t)head * a.nsplit_max + split) * wstride;
            rec[d] = o;
            if (d == 0) {
                rec[HD] = m_loc[r];
                rec[HD + 1] = l_loc[r];
            }
        }
    }
    if (nsplit_active == 1) return;

    // ---- merge: the last split of this KV head to arrive combines all of them in split order ----
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&a.counters[kvh], 1u);
        const int last = (prev == (unsigned)(nsplit_active - 1)) ? 1 : 0;
        if (last) a.counters[kvh] = 0;
        *flag = last;
    }
    __syncthreads();
    if (*flag == 0) return;
    __threadfence();
    for (int i = tid; i < NREP * HD; i += kAttnThreads) {
        const int r = i / HD, d = i % HD;
        const int head = kvh * NREP + r;
        const float *base = ws + (size_t)head * a.nsplit_max * wstride;
        float m = -INFINITY;
        for (int s = 0; s < nsplit_active; s++) m = fmaxf(m, ldg_cg_f32(base + (size_t)s * wstride + HD));
        float l = 0.f, o = 0.f;
        for (int s = 0; s < nsplit_active; s++) {
            const float w = __expf(ldg_cg_f32(base + (size_t)s * wstride + HD) - m);
            l += w * ldg_cg_f32(base + (size_t)s * wstride + HD + 1);
            o += w * ldg_cg_f32(base + (size_t)s * wstride + d);
        }
        a.out[(size_t)head * HD + d] = __float2half(o / l);
    }
}

size_t attn_smem_bytes(int nrep, int chunk) {
    size_t b = (size_t)2 * chunk * HD * 2;          // K, V slabs
    b += (size_t)nrep * HD * 4;                     // q
    b += (size_t)nrep * chunk * 4;                  // p
    b += (size_t)16 * nrep * HD * 4;                // PV partials
    b += (size_t)nrep * 16 * 4;                     // stats
    b += 2 * sizeof(uint64_t) + 16;
    return b;
}

template <int NREP>
cudaError_t launch(Ctx *ctx, const AttnDecodeArgs &a, bool pdl) {
    const size_t smem = attn_smem_bytes(NREP, a.chunk);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attn_decode_kernel<NREP>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if ((int)smem > ctx->smem_optin) return cudaErrorInvalidConfiguration;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(a.num_kv_heads, a.nsplit_max);
    cfg.blockDim = dim3(kAttnThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, attn_decode_kernel<NREP>, a);
}

}  // namespace

cudaError_t launch_attn_decode(Ctx *ctx, AttnDecodeArgs a, bool pdl) {
    if (a.head_dim != HD) return cudaErrorNotSupported;
    if (a.num_heads % a.num_kv_heads) return cudaErrorInvalidValue;
    const int nrep = a.num_heads / a.num_kv_heads;
    if (a.chunk <= 0) a.chunk = 128;
    a.nsplit_max = (a.max_ctx + a.chunk - 1) / a.chunk;
    const size_t need = (size_t)a.num_heads * a.nsplit_max * (HD + 2) * sizeof(float);
    if (need > ctx->attn_ws_bytes || a.num_kv_heads > 1024) return cudaErrorInvalidValue;
    a.ws = ctx->attn_ws;
    a.counters = ctx->attn_counters;
    switch (nrep) {
        case 1: return launch<1>(ctx, a, pdl);
        case 2: return launch<2>(ctx, a, pdl);
        case 4: return launch<4>(ctx, a, pdl);
        case 8: return launch<8>(ctx, a, pdl);
        default: return cudaErrorNotSupported;
    }
}

}  // namespace tce
