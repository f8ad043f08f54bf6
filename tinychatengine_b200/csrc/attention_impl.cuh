// attention_impl.cuh -- device-side body of the per-token KV-cache attention (one (kv head, split) work item),
// shared by the stand-alone kernel (attention.cu) and the persistent decode kernel (decode_megakernel.cu).
// See attention.cu for the design notes and reference citations.
#pragma once
#include "common.cuh"
#include "kernels_attn.h"

namespace tce {
namespace attn {

constexpr int HD = 128;  // head_dim (every Llama config in llm/include/model.h:71-83)
constexpr int kAttnThreads = 256;

inline __host__ __device__ size_t smem_bytes(int nrep, int chunk) {
    size_t b = (size_t)2 * chunk * HD * 2;  // K, V slabs
    b += (size_t)nrep * HD * 4;             // q
    b += (size_t)nrep * (chunk > HD + 8 ? chunk : HD + 8) * 4;  // p (cluster mode parks its [nrep][HD + 2] partial here)
    b += (size_t)16 * nrep * HD * 4;        // PV partials
    b += (size_t)nrep * 16 * 4;             // stats
    return (b + 15) & ~(size_t)15;
}

// One (kv head, split) work item executed by kAttnThreads threads that share `smem`.  `sync()` is a barrier over
// exactly those threads (__syncthreads in the stand-alone kernel, a named barrier inside the persistent kernel);
// `bar_parity` tracks the phase of the two TMA mbarriers, which are reused from item to item.
// ---- thread-block-cluster helpers (CL > 0: the splits of one KV head are the CTAs of one cluster) ----
TCE_DEVINL void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
TCE_DEVINL float ld_dsmem_f32(const float *local_ptr, uint32_t cta_rank) {  // the same shared-memory location in CTA `cta_rank` of the cluster
    uint32_t raddr;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_ptr)), "r"(cta_rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(raddr) : "memory");
    return v;
}

// CL == 0: splits are independent CTAs, merged through global partial records by the last CTA of the KV head to arrive.
// CL  > 0: `split` is the CTA's rank in a cluster of CL CTAs that covers the whole context of the KV head (rows are divided evenly
//          at run time); partial results stay in shared memory and are merged over distributed shared memory -- no global round trip,
//          no fence, no atomic, no serial last-CTA tail.
template <int NREP, int CL = 0, typename SyncFn>
TCE_DEVINL void attn_item(const AttnDecodeArgs &a, uint8_t *smem, uint64_t *bar, int *flag, uint32_t &bar_parity, int kvh, int split, int tid, int pos,
                          SyncFn sync) {
    const int chunk = a.chunk;  // shared-memory row capacity (strides); rows actually owned may be fewer in cluster mode
    __half *sK = reinterpret_cast<__half *>(smem);
    __half *sV = sK + (size_t)chunk * HD;
    float *sQ = reinterpret_cast<float *>(sV + (size_t)chunk * HD);  // [NREP][HD] rotated * alpha
    float *sP = sQ + NREP * HD;                                        // [NREP][chunk] scores -> probabilities
    float *sRed = sP + NREP * (chunk > HD + 8 ? chunk : HD + 8);       // [16][NREP][HD] PV partials
    float *sStat = sRed + 16 * NREP * HD;                              // [NREP][2 * 8 warps] max/sum scratch
    const int warp = tid >> 5, lane = tid & 31;
    sync();  // the previous item of this CTA (persistent kernel) may still be reading the shared buffers
    const int T = pos + 1;        // visible positions
    const int rows_per = (CL > 0) ? (T + CL - 1) / CL : chunk;
    const int t0 = min(T, split * rows_per);
    if (CL == 0 && t0 >= T) return;  // empty split
    const int t1 = min(T, t0 + rows_per);
    const int nrows = t1 - t0;       // cluster mode: may be 0 (the CTA then contributes m = -inf, l = 0, o = 0)
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
    sync();

    // ---- scores: 16 lanes per cached row (8 dims each), 2 rows per warp instruction ----
    const int sub = lane & 15, rsel = lane >> 4;
    float qreg[NREP][8];
#pragma unroll
    for (int r = 0; r < NREP; r++)
#pragma unroll
        for (int d = 0; d < 8; d++) qreg[r][d] = sQ[r * HD + sub * 8 + d];
    if (ncached > 0) mbar_wait(&bar[0], bar_parity);
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
    sync();

    // ---- softmax statistics of this split (fp32): m = max, p = exp(s - m), l = sum p ----
    float m_loc[NREP], l_loc[NREP];
#pragma unroll
    for (int r = 0; r < NREP; r++) {
        float m = -INFINITY;
        for (int i = tid; i < nrows; i += kAttnThreads) m = fmaxf(m, sP[r * chunk + i]);
        m = warp_max(m);
        if (lane == 0) sStat[r * 16 + warp] = m;
    }
    sync();
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
    sync();
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
    if (ncached > 0) mbar_wait(&bar[1], bar_parity);
    if (ncached > 0) bar_parity ^= 1;  // both barriers completed one phase
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
    sync();

    if constexpr (CL > 0) {
        // ---- cluster merge over distributed shared memory ----
        // own partial -> sPart[NREP][HD + 2] (the probability buffer is free now): unnormalised o, then m, l
        float *sPart = sP;
        constexpr int PS = HD + 2;
        for (int i = tid; i < NREP * HD; i += kAttnThreads) {
            const int r = i / HD, d = i % HD;
            float o = 0.f;
#pragma unroll
            for (int gsel = 0; gsel < 16; gsel++) o += sRed[((size_t)gsel * NREP + r) * HD + d];
            sPart[r * PS + d] = o;
            if (d == 0) {
                sPart[r * PS + HD] = m_loc[r];
                sPart[r * PS + HD + 1] = l_loc[r];
            }
        }
        cluster_sync_all();  // every CTA's partial is visible cluster-wide
        // gather the CL x NREP (m, l) pairs, derive per-head weights w[r][s] = exp(m_s - M_r) / L_r
        float *sW = sRed;  // [NREP][CL] weights (PV partials are consumed)
        float *sML = sRed + NREP * CL;  // [NREP][CL][2]
        for (int i = tid; i < NREP * CL * 2; i += kAttnThreads) {
            const int r = i / (CL * 2), s = (i / 2) % CL, which = i & 1;
            sML[i] = ld_dsmem_f32(&sPart[r * PS + HD + which], (uint32_t)s);
        }
        sync();
        for (int i = tid; i < NREP * CL; i += kAttnThreads) {
            const int r = i / CL, s = i % CL;
            float M = -INFINITY;
#pragma unroll
            for (int q = 0; q < CL; q++) M = fmaxf(M, sML[(r * CL + q) * 2]);
            float L = 0.f;
#pragma unroll
            for (int q = 0; q < CL; q++) {
                const float mq = sML[(r * CL + q) * 2];
                L += (mq == -INFINITY ? 0.f : __expf(mq - M)) * sML[(r * CL + q) * 2 + 1];
            }
            const float ms = sML[(r * CL + s) * 2];
            sW[i] = (ms == -INFINITY ? 0.f : __expf(ms - M)) / L;
        }
        sync();
        // this CTA finishes outputs [split * NO/CL, (split+1) * NO/CL): LPO lanes per output walk the CL partials
        constexpr int NO = NREP * HD, SL = NO / CL, LPO = kAttnThreads / SL;
        static_assert(NO % CL == 0 && kAttnThreads % SL == 0 && LPO <= 32 && (LPO & (LPO - 1)) == 0, "cluster merge mapping");
        {
            const int oi = split * SL + tid / LPO, j = tid % LPO;
            const int r = oi / HD, d = oi % HD;
            float acc = 0.f;
            for (int s = j; s < CL; s += LPO) acc += sW[r * CL + s] * ld_dsmem_f32(&sPart[r * PS + d], (uint32_t)s);
#pragma unroll
            for (int o = LPO / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (j == 0) a.out[(size_t)(kvh * NREP + r) * HD + d] = __float2half(acc);
        }
        cluster_sync_all();  // nobody leaves while a peer may still read its shared memory
        return;
    }

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
    sync();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&a.counters[kvh], 1u);
        const int last = (prev == (unsigned)(nsplit_active - 1)) ? 1 : 0;
        if (last) a.counters[kvh] = 0;
        *flag = last;
    }
    sync();
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


}  // namespace attn
}  // namespace tce
