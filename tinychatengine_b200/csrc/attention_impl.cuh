// attention_impl.cuh -- device-side body of the per-token KV-cache attention (one (kv head, split) work item),
// shared by the stand-alone kernel (attention.cu) and the persistent decode kernel (decode_persistent.cu).
// See attention.cu for the design notes and reference citations.
#pragma once
#include "common.cuh"
#include "kernels_attn.h"

namespace tce {
namespace attn {

constexpr int HD = 128;  // head_dim (every Llama config in llm/include/model.h:71-83)
constexpr int kAttnThreads = 256;

constexpr int kPitch = HD + 8;  // halves per K/V/Q row in shared memory: 272-byte rows keep ldmatrix and fragment loads conflict free

inline __host__ __device__ int rows16(int chunk) { return (chunk + 15) & ~15; }
inline __host__ __device__ int p_floats(int chunk) { return rows16(chunk) > HD + 8 ? rows16(chunk) : HD + 8; }
inline __host__ __device__ size_t smem_bytes(int nrep, int chunk) {
    size_t b = (size_t)2 * rows16(chunk) * kPitch * 2;  // K, V slabs (padded rows)
    b += (size_t)8 * kPitch * 2;                         // q, fp16, 8 head rows (rows >= nrep are zero)
    b += (size_t)nrep * p_floats(chunk) * 4;             // scores fp32 (cluster mode parks its [nrep][HD + 2] partial here)
    b += (size_t)nrep * rows16(chunk) * 2;               // probabilities fp16
    b += (size_t)nrep * HD * 4;                          // o
    b += (size_t)(2 + 3) * nrep * 16 * 4;                 // stats [2][nrep][16] + cluster-merge scratch [3][nrep][16]
    return (b + 15) & ~(size_t)15;
}

TCE_DEVINL void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
TCE_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
TCE_DEVINL void ldmatrix_x4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, const void *row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(row)));
}
TCE_DEVINL void ldmatrix_x4_t(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, const void *row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(smem_u32(row)));
}

// One (kv head, split) work item executed by kAttnThreads threads that share `smem`.  `sync()` is a barrier over
// exactly those threads (__syncthreads in the stand-alone kernel, a named barrier inside the persistent kernel).
// The H/KVH query heads of the KV head are the (up to 8) columns of one MMA tile: scores and P.V run on mma.sync m16n8k16 with
// fp16 operands (q * alpha, K, P, V) and fp32 accumulation; a scalar version of the two products cost ~6 of the kernel's ~18 us.
// Splits are independent CTAs, merged through global partial records by the last CTA of the KV head to arrive.  (A thread-block-cluster flavour --
// one cluster of 8 / 16 CTAs per KV head, partials merged over distributed shared memory -- was implemented and measured in round 1: 2.2 us per
// layer SLOWER (cluster co-scheduling + two cluster barriers); removed in round 2, where the decode step's attention lives in the persistent kernel.)
template <int NREP, typename SyncFn>
TCE_DEVINL void attn_item(const AttnDecodeArgs &a, uint8_t *smem, uint64_t *bar, int *flag, uint32_t &bar_parity, int kvh, int split, int tid, int pos,
                          SyncFn sync) {
    static_assert(NREP <= 8, "the query heads of one KV head ride in the 8 MMA columns");
    const int chunk = a.chunk;  // shared-memory row capacity; rows actually owned may be fewer in cluster mode
    const int R16 = rows16(chunk);
    __half *sK = reinterpret_cast<__half *>(smem);                       // [R16][kPitch]
    __half *sV = sK + (size_t)R16 * kPitch;                              // [R16][kPitch]
    __half *sQ = sV + (size_t)R16 * kPitch;                              // [8][kPitch] rotated * alpha, rows >= NREP zero
    float *sP = reinterpret_cast<float *>(sQ + 8 * kPitch);              // [NREP][p_floats] scores
    __half *sPh = reinterpret_cast<__half *>(sP + NREP * p_floats(chunk));  // [NREP][R16] probabilities
    float *sO = reinterpret_cast<float *>(sPh + NREP * R16);             // [NREP][HD] unnormalised output of this split
    float *sStat = sO + NREP * HD;                                       // [NREP][16] max | [NREP][16] sum
    float *sRed = sStat + 2 * NREP * 16;                                 // cluster-merge scratch
    (void)bar;
    (void)bar_parity;
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, qd = lane & 3;
    sync();  // the previous item of this CTA (persistent kernel) may still be reading the shared buffers
    const int T = pos + 1;        // visible positions
    const int rows_per = chunk;
    const int t0 = min(T, split * rows_per);
    if (t0 >= T) return;  // empty split
    const int t1 = min(T, t0 + rows_per);
    const int nrows = t1 - t0;       // cluster mode: may be 0 (the CTA then contributes m = -inf, l = 0, o = 0)
    const bool owns_new = (pos >= t0 && pos < t1);
    const int ncached = owns_new ? nrows - 1 : nrows;  // rows that already live in the cache
    const int mtiles = (nrows + 15) >> 4;

    __half *Kc = a.k_cache + ((size_t)kvh * a.max_ctx) * HD;
    __half *Vc = a.v_cache + ((size_t)kvh * a.max_ctx) * HD;
    // ---- cached K/V rows -> padded shared rows (16-byte cp.async, whole slab in flight at once) ----
    for (int e = tid; e < ncached * (HD / 8); e += kAttnThreads) {
        const int r = e >> 4, c8 = e & 15;
        cp_async16(sK + (size_t)r * kPitch + c8 * 8, Kc + (size_t)(t0 + r) * HD + c8 * 8);
        cp_async16(sV + (size_t)r * kPitch + c8 * 8, Vc + (size_t)(t0 + r) * HD + c8 * 8);
    }
    // rows of the last 16-row tile beyond nrows: finite zeros (their probabilities are zero, 0 * garbage must not be NaN)
    for (int e = tid; e < (mtiles * 16 - nrows) * (HD / 8); e += kAttnThreads) {
        const int r = nrows + (e >> 4), c8 = e & 15;
        *reinterpret_cast<uint4 *>(sK + (size_t)r * kPitch + c8 * 8) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(sV + (size_t)r * kPitch + c8 * 8) = make_uint4(0, 0, 0, 0);
    }

    // ---- RoPE (llm/src/ops/RotaryPosEmb.cc:7-69 rotate-half) on the NREP query heads and, if this CTA owns the new position, on
    //      the new key; fp32 math, tables [max_ctx][HD] fp32; q * alpha is rounded to fp16 for the tensor-core product ----
    const float *cosr = a.cos + (size_t)pos * HD, *sinr = a.sin + (size_t)pos * HD;
    const int H = a.num_heads, KVH = a.num_kv_heads;
    for (int i = tid; i < 8 * HD; i += kAttnThreads) {
        const int r = i / HD, j = i % HD;
        float v = 0.f;
        if (r < NREP) {
            const __half *q = a.qkv + (size_t)(kvh * NREP + r) * HD;
            const float x = __half2float(q[j]);
            const float xr = (j < HD / 2) ? -__half2float(q[j + HD / 2]) : __half2float(q[j - HD / 2]);
            v = (x * cosr[j] + xr * sinr[j]) * a.alpha;
        }
        sQ[r * kPitch + j] = __float2half(v);
    }
    if (owns_new && tid < HD) {
        const int j = tid;
        const __half *k = a.qkv + (size_t)H * HD + (size_t)kvh * HD;
        const __half *v = a.qkv + (size_t)(H + KVH) * HD + (size_t)kvh * HD;
        const float x = __half2float(k[j]);
        const float xr = (j < HD / 2) ? -__half2float(k[j + HD / 2]) : __half2float(k[j - HD / 2]);
        const __half kh = __float2half(x * cosr[j] + xr * sinr[j]);
        sK[(size_t)(nrows - 1) * kPitch + j] = kh;  // row `pos` of the slab; the copies above never touch it
        sV[(size_t)(nrows - 1) * kPitch + j] = v[j];
        Kc[(size_t)pos * HD + j] = kh;              // in-place append
        Vc[(size_t)pos * HD + j] = v[j];
    }
    cp_async_wait_all();
    sync();

    // ---- scores on the tensor cores: S^T[key][head] = K[key][:] . q[head][:]  (m16n8k16: 16 keys x 8 head columns x 16 dims) ----
    {
        uint32_t qb[8][2];  // B operand: q[head = g][dims], all 8 k-steps
#pragma unroll
        for (int ks = 0; ks < 8; ks++) {
            qb[ks][0] = *reinterpret_cast<const uint32_t *>(sQ + g * kPitch + ks * 16 + qd * 2);
            qb[ks][1] = *reinterpret_cast<const uint32_t *>(sQ + g * kPitch + ks * 16 + 8 + qd * 2);
        }
        for (int mt = warp; mt < mtiles; mt += kAttnThreads / 32) {
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            const __half *arow = sK + (size_t)(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + (lane >> 4) * 8;
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {
                uint32_t a0, a1, a2, a3;
                ldmatrix_x4(a0, a1, a2, a3, arow + ks * 16);
                mma_m16n8k16(c, a0, a1, a2, a3, qb[ks][0], qb[ks][1]);
            }
            const int key = mt * 16 + g, h0 = qd * 2;  // c0,c1: (key, heads h0, h0+1); c2,c3: (key + 8, ...)
            if (h0 < NREP) {
                if (key < nrows) sP[h0 * p_floats(chunk) + key] = c[0];
                if (key + 8 < nrows) sP[h0 * p_floats(chunk) + key + 8] = c[2];
            }
            if (h0 + 1 < NREP) {
                if (key < nrows) sP[(h0 + 1) * p_floats(chunk) + key] = c[1];
                if (key + 8 < nrows) sP[(h0 + 1) * p_floats(chunk) + key + 8] = c[3];
            }
        }
    }
    sync();

    // ---- softmax statistics of this split (fp32): m = max, p = exp(s - m) (kept as fp16 for the P.V product), l = sum p ----
    float m_loc[NREP], l_loc[NREP];
#pragma unroll
    for (int r = 0; r < NREP; r++) {
        float m = -INFINITY;
        for (int i = tid; i < nrows; i += kAttnThreads) m = fmaxf(m, sP[r * p_floats(chunk) + i]);
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
        for (int i = tid; i < mtiles * 16; i += kAttnThreads) {
            float p = 0.f;
            if (i < nrows) {
                p = __expf(sP[r * p_floats(chunk) + i] - m);
                l += p;
            }
            sPh[r * R16 + i] = __float2half(p);
        }
        l = warp_sum(l);
        if (lane == 0) sStat[NREP * 16 + r * 16 + warp] = l;
    }
    sync();
#pragma unroll
    for (int r = 0; r < NREP; r++) {
        float l = 0.f;
#pragma unroll
        for (int w = 0; w < kAttnThreads / 32; w++) l += sStat[NREP * 16 + r * 16 + w];
        l_loc[r] = l;
    }

    // ---- O[head][dim] = sum_key P[head][key] V[key][dim] on the tensor cores: warp w owns dims 16w..16w+15 over all keys ----
    {
        float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
        const int dbase = warp * 16;
        for (int kt = 0; kt < mtiles; kt++) {
            uint32_t a0 = 0u, a2 = 0u;  // A = P: rows = heads (g < NREP valid, rows 8..15 zero), k = 16 keys
            if (g < NREP) {
                a0 = *reinterpret_cast<const uint32_t *>(sPh + g * R16 + kt * 16 + qd * 2);
                a2 = *reinterpret_cast<const uint32_t *>(sPh + g * R16 + kt * 16 + 8 + qd * 2);
            }
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4_t(b0, b1, b2, b3, sV + (size_t)(kt * 16 + (lane & 15)) * kPitch + dbase + (lane >> 4) * 8);
            mma_m16n8k16(o0, a0, 0u, a2, 0u, b0, b1);
            mma_m16n8k16(o1, a0, 0u, a2, 0u, b2, b3);
        }
        if (g < NREP) {  // c0,c1: (head g, dims 2qd, 2qd+1)
            sO[g * HD + dbase + qd * 2] = o0[0];
            sO[g * HD + dbase + qd * 2 + 1] = o0[1];
            sO[g * HD + dbase + 8 + qd * 2] = o1[0];
            sO[g * HD + dbase + 8 + qd * 2 + 1] = o1[1];
        }
    }
    sync();

    // ---- per-split result (unnormalised o, m, l) ----
    const int nsplit_active = (T + chunk - 1) / chunk;
    float *ws = a.ws;  // [H][nsplit_max][HD + 2]
    const int wstride = HD + 2;
    for (int i = tid; i < NREP * HD; i += kAttnThreads) {
        const int r = i / HD, d = i % HD;
        const float o = sO[i];
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
