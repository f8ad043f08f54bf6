// decode_persistent.cu -- one persistent cooperative kernel per decoded token (Llama, AWQ-INT4, batch 1) on sm_100a.
//
// Call sites restated (reference, CUDA build): Int4LlamaForCausalLM::forward (cuda/Int4llamaForCausalLM.cu:17-50) ->
// Int4llamaDecoder::forward (cuda/Int4llamaDecoder.cu:57-112) -> 32 x Int4llamaDecoderLayer::forward (cuda/Int4llamaDecoderLayer.cu:73-115)
// -> Int4llamaAttention::forward (cuda/Int4llamaAttention.cu:116-229): ~19 kernels + 128 memcpys per layer on stream 0.  Here the whole
// token is ONE kernel of one CTA per SM whose warp roles persist across all phases (5 per layer + lm_head):
//
//   producer (1 warp, 1 elected lane)  walks the phases in order and keeps a ring of `nst` TMA stages full.  A stage is either
//        [16 rows x <=16 groups] of packed int4 weights (ONE 2-D UTMALDG, 16 KiB) + the stage's repacked scales|zeros record (one 640 B
//        UBLKCP), or 64 cached K rows or 64 cached V rows of the attention phase (two 128B-swizzled 2-D boxes, 16 KiB).  It depends
//        on nothing but static data and the token position, so it runs ahead across every phase boundary: while the GPU synchronises or
//        stages activations, up to nst x 16 KiB per SM of the NEXT matrices are already in flight, and HBM never goes idle.
//   consumers (16 warps)  per phase: wait for the grid barrier of the previous phase, quantise the activation vector (fused RMSNorm,
//        four int8 planes per 128-group), run the integer-MMA GEMV over this CTA's stage units (w4a16_gemv_impl.cuh: unit1), or
//        run flash-decoding attention straight out of the ring stages (mma.sync m16n8k16, ldmatrix on the swizzled K/V rows).
//   epilogue (1 warp)  reduces the 16 consumer partials of every tile, applies the fused epilogue (fp16 store, RED.ADD into the fp32
//        residual, SiLU(gate)*up, logits + running arg-max, tensor-parallel scatter to the peers) and signals the grid barrier.
//
// Grid barrier = one monotonic arrival counter per phase (red.release.gpu / ld.acquire.gpu); counters are never reset, the target of
// launch e is (e + 1) * #CTAs.  Data produced by other CTAs inside the kernel is read with ld.global.cg (L2), never through L1.
// All waits are bounded: a protocol bug surfaces as a launch failure within seconds, not as a hung GPU.
#include <stdio.h>

#include "attention_impl.cuh"
#include "persistent.h"
#include "w4a16_gemv_impl.cuh"

namespace tce {
namespace pk {

namespace {

using gemv::Lane1;
using gemv::make_lane1;
using gemv::unit1;

// ------------------------------------------------------------------------------------------------------------ small helpers
TCE_DEVINL unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
TCE_DEVINL unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
TCE_DEVINL void red_release_gpu(unsigned *p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
TCE_DEVINL void red_release_sys(unsigned *p) { asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
TCE_DEVINL uint4 ldcg_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
TCE_DEVINL float4 ldcg_f4(const void *p) {
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
TCE_DEVINL float ldcg_f32(const void *p) {
    float r;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
TCE_DEVINL unsigned short ldcg_u16(const void *p) {
    unsigned short r;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return r;
}
TCE_DEVINL float ldcg_half(const __half *p) { return __half2float(__ushort_as_half(ldcg_u16(p))); }

// wait until *ctr has reached `target` (wrap-safe); one thread
TCE_DEVINL void grid_wait(const unsigned *ctr, unsigned target) {
    if ((int)(ld_acquire_gpu(ctr) - target) >= 0) return;
    const long long t0 = clock64();
    while ((int)(ld_acquire_gpu(ctr) - target) < 0) {
        if (clock64() - t0 > 8000000000LL) __trap();
    }
}
// same for a counter peers arrive on over NVLink; a peer may legitimately lag (separate launch), so the bound is generous
TCE_DEVINL void sys_wait(const unsigned *ctr, unsigned target) {
    if ((int)(ld_acquire_sys(ctr) - target) >= 0) return;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(ctr) - target) < 0) {
        if (clock64() - t0 > 60000000000LL) __trap();
    }
}

TCE_DEVINL void stamp(const Args &a, int cta, int nphase, int p, int k) {  // one thread
    if (a.dbg) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.dbg[((size_t)cta * nphase + p) * 4 + k] = t;
    }
}

TCE_DEVINL unsigned long long argmax_key(float v, int idx) {
    unsigned b = __float_as_uint(v);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // monotone map float -> uint
    return ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);  // ties: lowest index wins (arg_max.cc)
}

struct PSmem {
    uint8_t *ring;      // [nst][kStageBytes], 1024-B aligned
    uint8_t *xs;        // activation planes (4 * IC bytes) | attention scratch
    float *gx;          // [max_ng] group steps
    int *gsum;          // [max_ng][2] group sums
    float *red;         // [kRedBufs][kCW][16] tile partials
    float *rms;         // [kCW] + misc
    uint64_t *full, *empty, *red_full, *red_empty;
    int *aflag;
    int nst;
};

TCE_DEVINL PSmem carve(uint8_t *raw, const Args &a) {
    PSmem s;
    uint8_t *base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    s.nst = a.nst;
    s.ring = base;
    uint8_t *p = base + (size_t)a.nst * kStageBytes;
    s.xs = p;
    p += a.xs_bytes;
    s.gx = reinterpret_cast<float *>(p);
    p += (size_t)a.max_ng * 4;
    s.gsum = reinterpret_cast<int *>(p);
    p += (size_t)a.max_ng * 8;
    s.red = reinterpret_cast<float *>(p);
    p += (size_t)kRedBufs * kCW * 16 * 4;
    s.rms = reinterpret_cast<float *>(p);
    p += 32 * 4;
    s.full = reinterpret_cast<uint64_t *>(p);
    s.empty = s.full + a.nst;
    s.red_full = s.empty + a.nst;
    s.red_empty = s.red_full + kRedBufs;
    s.aflag = reinterpret_cast<int *>(s.red_empty + kRedBufs);
    return s;
}

struct Ring {
    int stage = 0;
    uint32_t phase = 0;
    TCE_DEVINL void advance(int nst) {
        if (++stage == nst) {
            stage = 0;
            phase ^= 1;
        }
    }
};
struct Red {
    int rb = 0;
    uint32_t rphase = 0;
    TCE_DEVINL void advance() {
        if (++rb == kRedBufs) {
            rb = 0;
            rphase ^= 1;
        }
    }
};

// this CTA's stage-unit range [su0, su1) of one GEMV op
TCE_DEVINL void partition(const GemvOp &op, int cta, int ncta, int &su0, int &su1) {
    if (op.aligned) {
        const unsigned T = (unsigned)op.num_tiles;
        su0 = (int)((T * (unsigned)cta) / (unsigned)ncta) * op.S;
        su1 = (int)((T * (unsigned)(cta + 1)) / (unsigned)ncta) * op.S;
    } else {
        const unsigned U = (unsigned)op.SU;
        su0 = (int)(((unsigned long long)U * (unsigned)cta) / (unsigned)ncta);
        su1 = (int)(((unsigned long long)U * (unsigned)(cta + 1)) / (unsigned)ncta);
    }
}

// attention work split: the visible positions [0, T) in chunks of kKvChunk; every KV head gets NS = #CTAs / KVH consecutive CTAs,
// each takes `cps` consecutive chunks
struct AttnSplit {
    int kvh, split, ch0, ch1, nsplit;  // ch0 >= ch1: nothing to do
};
TCE_DEVINL AttnSplit attn_split(int cta, int ncta, int KVH, int pos) {
    AttnSplit s;
    const int T = pos + 1;
    const int nch = (T + kKvChunk - 1) / kKvChunk;
    int NS = ncta / KVH;
    if (NS < 1) NS = 1;
    const int cps = (nch + NS - 1) / NS;
    s.nsplit = (nch + cps - 1) / cps;
    s.kvh = cta / NS;
    s.split = cta - s.kvh * NS;
    if (s.kvh >= KVH || s.split >= s.nsplit) {
        s.ch0 = s.ch1 = 0;
        s.kvh = 0;
    } else {
        s.ch0 = s.split * cps;
        s.ch1 = min(nch, s.ch0 + cps);
    }
    return s;
}

// ------------------------------------------------------------------------------------------------------------ producer
TCE_DEVINL void produce_gemv(const GemvOp &op, const CUtensorMap *m0, const uint8_t *meta, const PSmem &sm, Ring &rs, int cta, int ncta, uint32_t leader,
                             uint64_t policy) {
    int su, su1;
    partition(op, cta, ncta, su, su1);
    int tile = su / op.S;
    int s = su - tile * op.S;
    for (; su < su1; su++) {
        mbar_wait(&sm.empty[rs.stage], rs.phase ^ 1);
        uint64_t *bar = &sm.full[rs.stage];
        uint8_t *dst = sm.ring + (size_t)rs.stage * kStageBytes;
        mbar_arrive_expect_tx_pred(bar, (uint32_t)op.box_bytes + kMetaBytes, leader);
        if (op.pair) {
            tma_load_2d_pred(dst, m0, s * 256, tile * 8, bar, policy, leader);
            tma_load_2d_pred(dst + 8 * op.sg * 64, m0 + 1, s * 256, tile * 8, bar, policy, leader);
        } else {
            int row = tile * 16;
            const CUtensorMap *m = m0;
            if (op.nseg > 1 && row >= op.rows0) {
                row -= op.rows0;
                m = m0 + 1;
                if (op.nseg > 2 && row >= op.rows1) {
                    row -= op.rows1;
                    m = m0 + 2;
                }
            }
            tma_load_2d_pred(dst, m, s * 256, row, bar, policy, leader);
        }
        bulk_g2s_pred(dst + kMetaOff, meta + (size_t)su * kMetaBytes, kMetaBytes, bar, policy, leader);
        __syncwarp();
        rs.advance(sm.nst);
        if (++s == op.S) {
            s = 0;
            tile++;
        }
    }
}

TCE_DEVINL void produce_attn(const Args &a, const LayerDesc &L, const CUtensorMap *kvmap, const PSmem &sm, Ring &rs, int cta, int ncta, int pos,
                             uint32_t leader, uint64_t policy) {
    const AttnSplit sp = attn_split(cta, ncta, a.KVH, pos);
    for (int c = sp.ch0; c < sp.ch1; c++) {
#pragma unroll 1
        for (int kv = 0; kv < 2; kv++) {
            const int row = (kv ? L.v_row0 : L.k_row0) + sp.kvh * a.max_ctx + c * kKvChunk;
            mbar_wait(&sm.empty[rs.stage], rs.phase ^ 1);
            uint64_t *bar = &sm.full[rs.stage];
            uint8_t *dst = sm.ring + (size_t)rs.stage * kStageBytes;
            mbar_arrive_expect_tx_pred(bar, 16384u, leader);
            tma_load_2d_pred(dst, kvmap, 0, row, bar, policy, leader);
            tma_load_2d_pred(dst + 8192, kvmap, 64, row, bar, policy, leader);
            __syncwarp();
            rs.advance(sm.nst);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------ consumers: GEMV
// Quantise the activation vector of one GEMV phase into the plane buffer.  Returns the factor the tile sums must be multiplied by
// (1/rms for the fused RMSNorm: y = inv * W (x . gamma), so the normalisation needs no second pass over x).
TCE_DEVINL float stage_x(const Args &a, const GemvOp &op, int x_mode, const PSmem &sm, const void *xsrc, const float *gamma, const float *tp_in, float *resid_out,
                         int token, int cta, int ctid, int cw, int lane) {
    const int units = op.IC / 8;
    float ss = 0.f;
    if (x_mode == PX_HALF) {
        constexpr int PRE = 4;
        for (int ui0 = 0; ui0 < units; ui0 += PRE * kConsumerThreads) {
            uint4 raw[PRE];
#pragma unroll
            for (int k = 0; k < PRE; k++) {
                const int ui = ui0 + k * kConsumerThreads + ctid;
                raw[k] = make_uint4(0u, 0u, 0u, 0u);
                if (ui < units) raw[k] = ldcg_u4(reinterpret_cast<const __half *>(xsrc) + (size_t)ui * 8);
            }
#pragma unroll
            for (int k = 0; k < PRE; k++) {
                if (ui0 + k * kConsumerThreads >= units) break;  // warp-uniform
                const int ui = ui0 + k * kConsumerThreads + ctid;
                const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw[k]);
                float v[8];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float2 f = __half22float2(h2[i]);
                    v[2 * i] = f.x;
                    v[2 * i + 1] = f.y;
                }
                gemv::emit_unit<1>(sm.xs, op.IC, sm.gx, sm.gsum, ui, ui < units, v, lane);
            }
        }
        named_bar_sync(1, kConsumerThreads);
        return 1.f;
    }
    // fp32 residual stream (or the embedding row of the token) with fused RMSNorm
    for (int ui0 = 0; ui0 < units; ui0 += kConsumerThreads) {  // warp-uniform trip count
        const int ui = ui0 + ctid;
        const bool valid = ui < units;
        float x[8], v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = 0.f;
        float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
        if (valid) {
            g0 = *reinterpret_cast<const float4 *>(gamma + (size_t)ui * 8);
            g1 = *reinterpret_cast<const float4 *>(gamma + (size_t)ui * 8 + 4);
            if (x_mode == PX_EMBED_RMS) {
                const uint4 raw = *reinterpret_cast<const uint4 *>(a.embed + (size_t)token * a.E + (size_t)ui * 8);
                const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float2 f = __half22float2(h2[i]);
                    x[2 * i] = f.x;
                    x[2 * i + 1] = f.y;
                }
            } else {
                const float4 r0 = ldcg_f4(reinterpret_cast<const float *>(xsrc) + (size_t)ui * 8);
                const float4 r1 = ldcg_f4(reinterpret_cast<const float *>(xsrc) + (size_t)ui * 8 + 4);
                x[0] = r0.x; x[1] = r0.y; x[2] = r0.z; x[3] = r0.w;
                x[4] = r1.x; x[5] = r1.y; x[6] = r1.z; x[7] = r1.w;
                if (tp_in) {
                    // tensor-parallel all-reduce, receive side: residual += sum over ranks, in rank order (bit-identical everywhere)
                    for (int pr = 0; pr < a.tp_size; pr++) {
                        const float4 q0 = ldcg_f4(tp_in + (size_t)pr * a.E + (size_t)ui * 8);
                        const float4 q1 = ldcg_f4(tp_in + (size_t)pr * a.E + (size_t)ui * 8 + 4);
                        x[0] += q0.x; x[1] += q0.y; x[2] += q0.z; x[3] += q0.w;
                        x[4] += q1.x; x[5] += q1.y; x[6] += q1.z; x[7] += q1.w;
                    }
                }
            }
            if (resid_out && cta == 0) {  // the embedding row / the reduced residual becomes the residual stream (one writer)
                *reinterpret_cast<float4 *>(resid_out + (size_t)ui * 8) = make_float4(x[0], x[1], x[2], x[3]);
                *reinterpret_cast<float4 *>(resid_out + (size_t)ui * 8 + 4) = make_float4(x[4], x[5], x[6], x[7]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) ss += x[i] * x[i];
        v[0] = x[0] * g0.x; v[1] = x[1] * g0.y; v[2] = x[2] * g0.z; v[3] = x[3] * g0.w;
        v[4] = x[4] * g1.x; v[5] = x[5] * g1.y; v[6] = x[6] * g1.z; v[7] = x[7] * g1.w;
        gemv::emit_unit<1>(sm.xs, op.IC, sm.gx, sm.gsum, ui, valid, v, lane);
    }
    ss = warp_sum(ss);
    if (lane == 0) sm.rms[cw] = ss;
    named_bar_sync(1, kConsumerThreads);
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kCW; w++) tot += sm.rms[w];
    return rsqrtf(tot / (float)op.IC + a.eps);  // LlamaRMSNorm (llm/src/ops/LlamaRMSNorm.cc): x / sqrt(mean(x^2) + eps) * weight
}

TCE_DEVINL void consume_gemv(const GemvOp &op, const PSmem &sm, Ring &rs, Red &cs, float inv, int cta, int ncta, int cw, int lane) {
    const int g = lane >> 2, t = lane & 3;
    int su, su1;
    partition(op, cta, ncta, su, su1);
    const int rp = op.sg * 64;  // dense row pitch of the TMA box
    const uint32_t w_off = (uint32_t)(g * rp + t * 16 + cw * 64);
    const Lane1 L = make_lane1(sm.xs, op.IC, g, t);
    int tile = su / op.S;
    int sb = su - tile * op.S;
    while (su < su1) {
        const int se = min(op.S, sb + (su1 - su));
        float totA = 0.f, totB = 0.f;
        for (int s = sb; s < se; s++) {
            const int n = min(16, op.NG - 16 * s);  // groups this stage carries
            mbar_wait(&sm.full[rs.stage], rs.phase);
            if (cw < n) {
                const uint8_t *base = sm.ring + (size_t)rs.stage * kStageBytes;
                const uint4 wa = *reinterpret_cast<const uint4 *>(base + w_off);
                const uint4 wb = *reinterpret_cast<const uint4 *>(base + w_off + 8 * rp);
                const __half *sc = reinterpret_cast<const __half *>(base + kMetaOff) + cw * 16 + g;
                const uint2 z = *reinterpret_cast<const uint2 *>(base + kMetaOff + 512 + cw * 8);
                const float sAq = __half2float(sc[0]), sBq = __half2float(sc[8]);
                const int zAq = (int)((z.x >> (4 * g)) & 0xFu), zBq = (int)((z.y >> (4 * g)) & 0xFu);
                unit1(L, wa, wb, 16 * s + cw, sAq, sBq, zAq, zBq, sm.gx, sm.gsum, totA, totB);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[rs.stage]);
            rs.advance(sm.nst);
        }
        // ---- hand the (possibly partial) tile sums to the epilogue warp ----
        totA += __shfl_xor_sync(0xffffffffu, totA, 1);  // (hi, mid) share of t = 0 + lo share of t = 1
        totB += __shfl_xor_sync(0xffffffffu, totB, 1);
        mbar_wait(&sm.red_empty[cs.rb], cs.rphase ^ 1);
        float *rbuf = sm.red + ((size_t)cs.rb * kCW + cw) * 16;
        if (t == 0) {
            rbuf[g] = totA * inv;
            rbuf[g + 8] = totB * inv;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.red_full[cs.rb]);
        cs.advance();
        su += se - sb;
        sb = 0;
        tile++;
    }
}

// ------------------------------------------------------------------------------------------------------------ epilogue warp
struct EpiOut {
    void *y;                 // fp16 / fp32 output vector, or the fp32 residual for PE_ADD_F32
    float *logits;
    unsigned long long best;  // PE_LOGITS: running arg-max key of this warp
    int index_base;
};

TCE_DEVINL void epilogue_gemv(const Args &a, const GemvOp &op, const PSmem &sm, Red &es, EpiOut &o, int tp_buf, int cta, int ncta, int lane) {
    int su, su1;
    partition(op, cta, ncta, su, su1);
    int tile = su / op.S;
    int sb = su - tile * op.S;
    while (su < su1) {
        const int se = min(op.S, sb + (su1 - su));
        mbar_wait(&sm.red_full[es.rb], es.rphase);
        const float *rbuf = sm.red + (size_t)es.rb * kCW * 16;
        // lane l < 16 sums consumer warps 0..7 of row l, lane l + 16 warps 8..15
        float v = 0.f;
        {
            const int row = lane & 15, w0 = (lane >> 4) * 8;
#pragma unroll
            for (int w = 0; w < 8; w++) v += rbuf[(w0 + w) * 16 + row];
        }
        v += __shfl_down_sync(0xffffffffu, v, 16);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.red_empty[es.rb]);
        es.advance();
        switch (op.epi) {
            case PE_ADD_F32:
                // residual accumulate: RED.ADD of the tile sum (a split tile's contributors each add their part; fire and forget)
                if (lane < 16) atomicAdd(reinterpret_cast<float *>(o.y) + (size_t)tile * 16 + lane, v);
                break;
            case PE_STORE_HALF:
                if (lane < 16) reinterpret_cast<__half *>(o.y)[(size_t)tile * 16 + lane] = __float2half(v);
                break;
            case PE_SILU_MUL: {
                // rows 0-7 = gate, rows 8-15 = up of the same output channels: y = SiLU(gate) * up
                // (reference SiLuMul_half, llm/src/nn_modules/cuda/Int4llamaDecoderLayer.cu:21-30; fp32 here)
                const float up = __shfl_down_sync(0xffffffffu, v, 8);
                if (lane < 8) reinterpret_cast<__half *>(o.y)[(size_t)tile * 8 + lane] = __float2half(v / (1.f + __expf(-v)) * up);
                break;
            }
            case PE_LOGITS:
                if (lane < 16) {
                    const int idx = tile * 16 + lane;
                    o.logits[idx] = v;
                    const unsigned long long key = argmax_key(v, o.index_base + idx);
                    o.best = key > o.best ? key : o.best;
                }
                break;
            case PE_TP_SCATTER:
                // fused collective: the finished outputs go straight into slot `rank` of every rank's gather buffer over NVLink
                if (lane < 16) {
                    for (int pr = 0; pr < a.tp_size; pr++)
                        a.tp_gather[pr][((size_t)tp_buf * a.tp_size + a.tp_rank) * a.E + (size_t)tile * 16 + lane] = v;
                }
                break;
        }
        su += se - sb;
        sb = 0;
        tile++;
    }
}

// ------------------------------------------------------------------------------------------------------------ consumers: attention
// byte offset of (row r, 16-byte chunk c of the 256-byte row) inside a K or V stage: two [64 rows][128 B] boxes, 128B-swizzled
TCE_DEVINL uint32_t kv_off(int r, int c) { return (uint32_t)((c >> 3) * 8192 + r * 128 + (((c & 7) ^ (r & 7)) << 4)); }

TCE_DEVINL void attention_phase(const Args &a, const LayerDesc &L, const PSmem &sm, Ring &rs, int cta, int ncta, int pos, int ctid, int cw, int lane) {
    const AttnSplit sp = attn_split(cta, ncta, a.KVH, pos);
    if (sp.ch0 >= sp.ch1) return;
    const int nrep = a.nrep;
    const int g = lane >> 2, t = lane & 3;
    // scratch (the activation-plane buffer is idle during this phase)
    __half *sQ = reinterpret_cast<__half *>(sm.xs);                        // [8][136] q * alpha after RoPE, rows >= nrep zero
    float *sO = reinterpret_cast<float *>(sm.xs + 8 * 136 * 2);            // [kCW][nrep][128] per-warp unnormalised outputs
    float *sML = sO + (size_t)kCW * nrep * 128;                            // [kCW][nrep][2] per-warp (max, sum)
    const float *cosr = a.cos + (size_t)pos * 128, *sinr = a.sin + (size_t)pos * 128;
    // ---- RoPE (llm/src/ops/RotaryPosEmb.cc:7-69, rotate-half) on the nrep query heads of this KV head; fp32 math ----
    for (int i = ctid; i < 8 * 128; i += kConsumerThreads) {
        const int r = i >> 7, j = i & 127;
        float v = 0.f;
        if (r < nrep) {
            const __half *q = a.qkv + (size_t)(sp.kvh * nrep + r) * 128;
            const float x = ldcg_half(q + j);
            const float xr = (j < 64) ? -ldcg_half(q + j + 64) : ldcg_half(q + j - 64);
            v = (x * cosr[j] + xr * sinr[j]) * a.alpha;
        }
        sQ[r * 136 + j] = __float2half(v);
    }
    named_bar_sync(1, kConsumerThreads);
    uint32_t qa[8][2];  // A operand: q[head g][dims], all 8 k-steps (rows 8..15 of the MMA tile are zero)
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {
        qa[ks][0] = *reinterpret_cast<const uint32_t *>(sQ + g * 136 + ks * 16 + t * 2);
        qa[ks][1] = *reinterpret_cast<const uint32_t *>(sQ + g * 136 + ks * 16 + 8 + t * 2);
    }
    float m_run = -INFINITY, l_run = 0.f;  // of head row g (replicated over t)
    bool have = false;
    float *myO = sO + (size_t)cw * nrep * 128;
    const int kb = cw & 3;  // 16-key block of the chunk this warp owns
    for (int c = sp.ch0; c < sp.ch1; c++) {
        const bool mine = (((c - sp.ch0) & 3) == (cw >> 2)) && (c * kKvChunk + kb * 16 <= pos);
        const int kbase = c * kKvChunk + kb * 16;          // first key of the block
        const bool has_new = mine && pos >= kbase && pos < kbase + 16;
        // ================= K stage: scores =================
        mbar_wait(&sm.full[rs.stage], rs.phase);
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
        if (mine) {
            uint8_t *kst = sm.ring + (size_t)rs.stage * kStageBytes;
            if (has_new) {
                // the token's own key: RoPE, round to fp16, append to the cache and patch the (stale) row of the stage
                const __half *k = a.qkv + (size_t)a.H * 128 + (size_t)sp.kvh * 128;
                const int r = pos - c * kKvChunk;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int j = lane * 4 + i;
                    const float x = ldcg_half(k + j);
                    const float xr = (j < 64) ? -ldcg_half(k + j + 64) : ldcg_half(k + j - 64);
                    const __half kh = __float2half(x * cosr[j] + xr * sinr[j]);
                    *reinterpret_cast<__half *>(kst + kv_off(r, j >> 3) + (j & 7) * 2) = kh;
                    L.k_cache[((size_t)sp.kvh * a.max_ctx + pos) * 128 + j] = kh;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes into a stage the TMA unit will refill
                __syncwarp();
            }
            const int lr = (lane & 7) + ((lane >> 4) << 3);  // ldmatrix row supplied by this lane (key within the block)
            const int lc = (lane >> 3) & 1;                  // ... and which 8-dim half of the k-step
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {
                uint32_t b0, b1, b2, b3;
                attn::ldmatrix_x4(b0, b1, b2, b3, kst + kv_off(kb * 16 + lr, 2 * ks + lc));
                mma_m16n8k16(s0, qa[ks][0], 0u, qa[ks][1], 0u, b0, b1);  // keys 0..7 of the block
                mma_m16n8k16(s1, qa[ks][0], 0u, qa[ks][1], 0u, b2, b3);  // keys 8..15
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[rs.stage]);
        rs.advance(sm.nst);
        // ================= V stage: softmax + P.V =================
        mbar_wait(&sm.full[rs.stage], rs.phase);
        if (mine) {
            uint8_t *vst = sm.ring + (size_t)rs.stage * kStageBytes;
            if (has_new) {
                const __half *v = a.qkv + (size_t)(a.H + a.KVH) * 128 + (size_t)sp.kvh * 128;
                const int r = pos - c * kKvChunk;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int j = lane * 4 + i;
                    const __half vh = __ushort_as_half(ldcg_u16(v + j));
                    *reinterpret_cast<__half *>(vst + kv_off(r, j >> 3) + (j & 7) * 2) = vh;
                    L.v_cache[((size_t)sp.kvh * a.max_ctx + pos) * 128 + j] = vh;
                }
                // rows of the block beyond the token were never written for this sequence: finite zeros (0 * garbage must not be NaN)
                for (int rr = r + 1; rr < kb * 16 + 16; rr++)
                    *reinterpret_cast<uint2 *>(vst + kv_off(rr, lane >> 1) + (lane & 1) * 8) = make_uint2(0u, 0u);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
            }
            // thread (g, t): head row g, keys kbase + {2t, 2t+1} (s0) and kbase + 8 + {2t, 2t+1} (s1)
            const int k0 = kbase + 2 * t;
            float e0 = (k0 <= pos) ? s0[0] : -INFINITY, e1 = (k0 + 1 <= pos) ? s0[1] : -INFINITY;
            float e2 = (k0 + 8 <= pos) ? s1[0] : -INFINITY, e3 = (k0 + 9 <= pos) ? s1[1] : -INFINITY;
            float mb = fmaxf(fmaxf(e0, e1), fmaxf(e2, e3));
            mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1));
            mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
            const float m_new = fmaxf(m_run, mb);  // finite: key kbase is visible
            e0 = __expf(e0 - m_new);
            e1 = __expf(e1 - m_new);
            e2 = __expf(e2 - m_new);
            e3 = __expf(e3 - m_new);
            float lb = (e0 + e1) + (e2 + e3);
            lb += __shfl_xor_sync(0xffffffffu, lb, 1);
            lb += __shfl_xor_sync(0xffffffffu, lb, 2);
            const float sc_old = have ? __expf(m_run - m_new) : 0.f;
            l_run = l_run * sc_old + lb;
            m_run = m_new;
            const uint32_t pa0 = pack_half2(e0, e1), pa2 = pack_half2(e2, e3);  // A operand: P[head g][keys], rows 8..15 zero
            const int lr = (lane & 7) + (((lane >> 3) & 1) << 3);  // ldmatrix.trans row = key within the block
            const int lc = lane >> 4;                                // ... which of the two 8-dim n-tiles
#pragma unroll
            for (int h = 0; h < 2; h++) {  // dims 64h .. 64h + 63
                float oacc[8][4];
#pragma unroll
                for (int j = 0; j < 8; j++) oacc[j][0] = oacc[j][1] = oacc[j][2] = oacc[j][3] = 0.f;
#pragma unroll
                for (int jp = 0; jp < 4; jp++) {
                    uint32_t b0, b1, b2, b3;
                    attn::ldmatrix_x4_t(b0, b1, b2, b3, vst + kv_off(kb * 16 + lr, 8 * h + 2 * jp + lc));
                    mma_m16n8k16(oacc[2 * jp], pa0, 0u, pa2, 0u, b0, b1);
                    mma_m16n8k16(oacc[2 * jp + 1], pa0, 0u, pa2, 0u, b2, b3);
                }
                if (g < nrep) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        float2 *dst = reinterpret_cast<float2 *>(myO + g * 128 + 64 * h + 8 * j + 2 * t);
                        float2 nv = make_float2(oacc[j][0], oacc[j][1]);
                        if (have) {
                            const float2 old = *dst;
                            nv.x += old.x * sc_old;
                            nv.y += old.y * sc_old;
                        }
                        *dst = nv;
                    }
                }
            }
            have = true;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[rs.stage]);
        rs.advance(sm.nst);
    }
    if (t == 0 && g < nrep) {
        sML[(cw * nrep + g) * 2] = have ? m_run : -INFINITY;
        sML[(cw * nrep + g) * 2 + 1] = have ? l_run : 0.f;
    }
    named_bar_sync(1, kConsumerThreads);
    // ---- merge the 16 warp partials of this CTA: thread i < nrep * 128 owns (head r, dim d) ----
    float o = 0.f, M = -INFINITY, Lsum = 0.f;
    const int r = ctid >> 7, d = ctid & 127;
    const bool owner = ctid < nrep * 128;
    if (owner) {
        for (int w = 0; w < kCW; w++) M = fmaxf(M, sML[(w * nrep + r) * 2]);
        for (int w = 0; w < kCW; w++) {
            const float mw = sML[(w * nrep + r) * 2];
            if (mw != -INFINITY) {
                const float wt = __expf(mw - M);
                Lsum += wt * sML[(w * nrep + r) * 2 + 1];
                o += wt * sO[((size_t)w * nrep + r) * 128 + d];
            }
        }
    }
    const int head = sp.kvh * nrep + r;
    if (sp.nsplit == 1) {
        if (owner) a.attn[(size_t)head * 128 + d] = __float2half(o / Lsum);
        return;
    }
    int NS = ncta / a.KVH;
    if (NS < 1) NS = 1;
    if (owner) {
        float *rec = a.attn_ws + ((size_t)head * NS + sp.split) * 130;
        rec[d] = o;
        if (d == 0) {
            rec[128] = M;
            rec[129] = Lsum;
        }
    }
    // ---- the last split of this KV head to arrive combines all of them in split order ----
    __threadfence();
    named_bar_sync(1, kConsumerThreads);
    if (ctid == 0) {
        const unsigned prev = atomicAdd(&a.attn_cnt[sp.kvh], 1u);
        const int last = (prev == (unsigned)(sp.nsplit - 1)) ? 1 : 0;
        if (last) a.attn_cnt[sp.kvh] = 0;
        *sm.aflag = last;
    }
    named_bar_sync(1, kConsumerThreads);
    if (*sm.aflag == 0) return;
    __threadfence();
    if (owner) {
        const float *base = a.attn_ws + (size_t)head * NS * 130;
        float m = -INFINITY;
        for (int s = 0; s < sp.nsplit; s++) m = fmaxf(m, ldcg_f32(base + (size_t)s * 130 + 128));
        float l = 0.f, acc = 0.f;
        for (int s = 0; s < sp.nsplit; s++) {
            const float w = __expf(ldcg_f32(base + (size_t)s * 130 + 128) - m);
            l += w * ldcg_f32(base + (size_t)s * 130 + 129);
            acc += w * ldcg_f32(base + (size_t)s * 130 + d);
        }
        a.attn[(size_t)head * 128 + d] = __float2half(acc / l);
    }
}

// ------------------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kThreads, 1) decode_persistent_kernel(const __grid_constant__ Args a) {
    extern __shared__ uint8_t smem_raw[];
    const PSmem sm = carve(smem_raw, a);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int Lyr = a.num_layers;
    const int token = a.tokpos[0], pos = a.tokpos[1];
    if (token < 0 || token >= a.embed_rows || pos < 0 || pos >= a.max_ctx) {  // uniform over the grid: nobody starts
        if (cta == 0 && tid == 0) {
            *a.error = 1;
            *a.next_token = -1;
        }
        return;
    }
    if (warp == 0) {
        if (lane < a.nst) {
            mbar_init(&sm.full[lane], 1);
            mbar_init(&sm.empty[lane], kCW);
        } else if (lane >= 16 && lane < 16 + kRedBufs) {
            mbar_init(&sm.red_full[lane - 16], kCW);
            mbar_init(&sm.red_empty[lane - 16], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned epoch = *a.epoch;
    const unsigned target = (epoch + 1u) * (unsigned)ncta;  // every phase counter reaches this when all CTAs have arrived in this launch
    const bool tp = a.tp_size > 1;
    // tensor-parallel arrival counters advance by tp_size * ncta per collective, num_layers collectives per launch on each of the two
    const unsigned tp_per = (unsigned)a.tp_size * (unsigned)ncta;
    const unsigned tp_base = epoch * (unsigned)Lyr * tp_per;

    // phase p = 5 * layer + k, k: 0 RMSNorm + q|k|v, 1 attention, 2 o_proj, 3 RMSNorm + gate|up, 4 down_proj; p = 5 * Lyr: lm_head
    const int nphase = 5 * Lyr + 1;
    if (warp == 0) {
        // ================= producer: every byte this CTA needs from HBM, in consumption order =================
        Ring rs;
        const uint64_t policy = l2_policy_evict_first();
        const uint32_t leader = (lane == 0) ? 1u : 0u;
        const CUtensorMap *kvmap = a.maps + (size_t)Lyr * 7 + 1;
#pragma unroll 1
        for (int p = 0; p < nphase; p++) {
            const int l = p / 5, k = p - 5 * l;
            if (l == Lyr) {
                produce_gemv(a.op[OPI_LMHEAD], a.maps + (size_t)Lyr * 7, a.lm_meta, sm, rs, cta, ncta, leader, policy);
            } else if (k == 1) {
                produce_attn(a, a.layers[l], kvmap, sm, rs, cta, ncta, pos, leader, policy);
            } else {
                const int oi = (k == 0) ? OPI_QKV : (k - 1);       // k = 2,3,4 -> OPI_O, OPI_GATEUP, OPI_DOWN
                const int mi = (k == 0) ? 0 : (k == 2 ? 3 : (k == 3 ? 4 : 6));  // first tensor map of the op within the layer's seven
                produce_gemv(a.op[oi], a.maps + (size_t)l * 7 + mi, a.layers[l].meta[oi], sm, rs, cta, ncta, leader, policy);
            }
        }
        return;
    }
    if (warp == 1) {
        // ================= epilogue warp =================
        Red es;
        EpiOut o;
        o.logits = a.logits;
        o.best = 0ull;
        o.index_base = a.vocab_base;
#pragma unroll 1
        for (int p = 0; p < nphase; p++) {
            const int l = p / 5, k = p - 5 * l;
            if (l < Lyr && k == 1) continue;  // attention: the consumers signal the barrier themselves
            const int oi = (l == Lyr) ? OPI_LMHEAD : ((k == 0) ? OPI_QKV : (k - 1));
            o.y = (oi == OPI_QKV) ? (void *)a.qkv : (oi == OPI_GATEUP ? (void *)a.act : (void *)a.resid);
            epilogue_gemv(a, a.op[oi], sm, es, o, (oi == OPI_DOWN) ? 1 : 0, cta, ncta, lane);
            if (tp && (oi == OPI_O || oi == OPI_DOWN)) {
                // this CTA's peer stores are fenced system-wide, then it checks in with every rank
                __threadfence_system();
                __syncwarp();
                if (lane < a.tp_size) red_release_sys(a.tp_arrive[lane] + (oi == OPI_DOWN ? 1 : 0));
            }
            if (oi == OPI_LMHEAD) {
                unsigned long long key = o.best;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, off);
                    key = other > key ? other : key;
                }
                if (lane == 0 && key) atomicMax(a.argmax_cell, key);
            }
            // everything this warp wrote in phase p is visible device-wide before the arrival
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                red_release_gpu(a.sync + p);
                stamp(a, cta, nphase, p, 3);
            }
        }
        return;
    }

    // ================= consumers =================
    const int ctid = tid - 64;
    const int cw = warp - 2;
    Ring rs;
    Red cs;
    // tensor parallel: residual buffers ping-pong (every rank reduces the same gathered partials into the other buffer)
    float *resid_cur = a.resid, *resid_alt = a.resid + a.E;
#pragma unroll 1
    for (int p = 0; p < nphase; p++) {
        const int l = p / 5, k = p - 5 * l;
        if (p > 0) {  // all CTAs have completed phase p - 1
            if (ctid == 0) grid_wait(a.sync + p - 1, target);
            named_bar_sync(1, kConsumerThreads);
        }
        if (ctid == 0) stamp(a, cta, nphase, p, 0);
        if (l < Lyr && k == 1) {
            // ---- RoPE + KV append + attention ----
            attention_phase(a, a.layers[l], sm, rs, cta, ncta, pos, ctid, cw, lane);
            if (ctid == 0) stamp(a, cta, nphase, p, 2);
            __threadfence();
            named_bar_sync(1, kConsumerThreads);
            if (ctid == 0) {
                red_release_gpu(a.sync + p);
                stamp(a, cta, nphase, p, 3);
            }
            continue;
        }
        const int oi = (l == Lyr) ? OPI_LMHEAD : ((k == 0) ? OPI_QKV : (k - 1));
        int x_mode = a.op[oi].x_mode;
        const void *xsrc = resid_cur;
        const float *gamma = nullptr, *tin = nullptr;
        float *rout = nullptr;
        if (oi == OPI_O) {
            xsrc = a.attn;
        } else if (oi == OPI_DOWN) {
            xsrc = a.act;
        } else {
            gamma = (oi == OPI_LMHEAD) ? a.final_norm : (oi == OPI_QKV ? a.layers[l].input_norm : a.layers[l].post_norm);
            if (p == 0) {
                // the token's embedding row is the residual stream (reference: CPU Embedding, cuda/Int4llamaDecoder.cu:62-69)
                x_mode = PX_EMBED_RMS;
                rout = resid_cur;
            } else if (tp) {
                // tensor-parallel all-reduce, receive side: the collective that feeds this RMSNorm (o_proj of this layer for gate|up,
                // down_proj of the previous layer otherwise) has landed in the local gather buffer once every CTA of every rank arrived
                const int buf = (oi == OPI_GATEUP) ? 0 : 1;
                const unsigned done = (oi == OPI_GATEUP) ? (unsigned)(l + 1) : (unsigned)l;  // collectives completed on that buffer this launch
                if (ctid == 0) sys_wait(a.tp_arrive[a.tp_rank] + buf, tp_base + done * tp_per);
                named_bar_sync(1, kConsumerThreads);
                tin = a.tp_gather[a.tp_rank] + (size_t)buf * a.tp_size * a.E;
                rout = resid_alt;
            }
        }
        const float inv = stage_x(a, a.op[oi], x_mode, sm, xsrc, gamma, tin, rout, token, cta, ctid, cw, lane);
        if (tin) {
            float *tmp = resid_cur;
            resid_cur = resid_alt;
            resid_alt = tmp;
        }
        if (ctid == 0) stamp(a, cta, nphase, p, 1);
        consume_gemv(a.op[oi], sm, rs, cs, inv, cta, ncta, cw, lane);
        if (ctid == 0) stamp(a, cta, nphase, p, 2);
    }
    // ---- greedy token: decoded once every CTA's epilogue has contributed its maximum ----
    if (cta == 0 && ctid == 0) {
        grid_wait(a.sync + 5 * Lyr, target);
        unsigned long long key = *reinterpret_cast<volatile unsigned long long *>(a.argmax_cell);
        if (tp) {
            // vocabulary shards: scatter the local key to every rank, wait for all of them, take the global maximum
            for (int pr = 0; pr < a.tp_size; pr++) *reinterpret_cast<volatile unsigned long long *>(a.tp_keys[pr] + a.tp_rank) = key;
            __threadfence_system();
            for (int pr = 0; pr < a.tp_size; pr++) red_release_sys(a.tp_key_arrive[pr]);
            sys_wait(a.tp_key_arrive[a.tp_rank], (epoch + 1u) * (unsigned)a.tp_size);
            key = 0ull;
            for (int pr = 0; pr < a.tp_size; pr++) {
                const unsigned long long k2 = *reinterpret_cast<volatile unsigned long long *>(a.tp_keys[a.tp_rank] + pr);
                key = k2 > key ? k2 : key;
            }
        }
        *a.next_token = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
        *a.argmax_cell = 0ull;  // re-armed for the next launch (every CTA has arrived: nobody touches it any more)
        *a.epoch = epoch + 1u;
    }
}

// ------------------------------------------------------------------------------------------------------------ repack kernel
// scales half[rows][sf_w] + zeros u32[rows][zeros_w] (QM_CUDA, llm/tools/quantize_methods.py:370-442) -> one 640-byte record per
// (16-row tile, 16-group stage): scales half[16 groups][16 rows], zeros u64[16 groups] (nibble r = zero point of row r).
__global__ void repack_meta_kernel(W4Seg s0, W4Seg s1, W4Seg s2, int nseg, int pair, int NG, int zeros_w, int sf_w, int S, int num_tiles, uint8_t *out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (tile, s, gi)
    if (idx >= num_tiles * S * 16) return;
    const int gi = idx & 15, su = idx >> 4;
    const int tile = su / S, s = su - tile * S;
    const int G = 16 * s + gi;
    __half sc[16];
    unsigned long long z = 0ull;
    for (int r = 0; r < 16; r++) {
        const W4Seg *seg = &s0;
        int row;
        if (pair) {
            seg = (r < 8) ? &s0 : &s1;
            row = tile * 8 + (r & 7);
        } else {
            row = tile * 16 + r;
            if (nseg > 1 && row >= s0.rows) {
                row -= s0.rows;
                seg = &s1;
                if (nseg > 2 && row >= s1.rows) {
                    row -= s1.rows;
                    seg = &s2;
                }
            }
        }
        if (G < NG) {
            sc[r] = seg->scales[(size_t)row * sf_w + G];
            z |= (unsigned long long)((seg->zeros[(size_t)row * zeros_w + (G >> 3)] >> ((G & 7) * 4)) & 0xFu) << (4 * r);
        } else {
            sc[r] = __float2half(0.f);
        }
    }
    uint8_t *rec = out + (size_t)su * kMetaBytes;
    __half *so = reinterpret_cast<__half *>(rec) + gi * 16;
    for (int r = 0; r < 16; r++) so[r] = sc[r];
    reinterpret_cast<unsigned long long *>(rec + 512)[gi] = z;
}

}  // namespace

int attn_scratch_bytes(int nrep) { return 8 * 136 * 2 + kCW * nrep * 128 * 4 + kCW * nrep * 2 * 4; }

static size_t fixed_bytes(int xs_bytes, int max_ng) {
    return (size_t)xs_bytes + (size_t)max_ng * 12 + (size_t)kRedBufs * kCW * 16 * 4 + 32 * 4 + (size_t)(2 * kMaxStages + 2 * kRedBufs) * 8 + 16 + 1024;
}
int pick_stages(int smem_optin, int xs_bytes, int max_ng) {
    const long long avail = (long long)smem_optin - (long long)fixed_bytes(xs_bytes, max_ng);
    long long n = avail / kStageBytes;
    if (n > kMaxStages) n = kMaxStages;
    return n < 2 ? 0 : (int)n;
}
size_t smem_bytes(const Args &a) { return fixed_bytes(a.xs_bytes, a.max_ng) + (size_t)a.nst * kStageBytes; }

cudaError_t repack_meta(Ctx *ctx, const W4Seg *segs, int nseg, int pair, int IC, uint8_t *out, cudaStream_t stream) {
    const int NG = IC / kW4Group, S = (NG + 15) / 16;
    int rows = 0;
    for (int i = 0; i < nseg; i++) rows += segs[i].rows;
    const int num_tiles = rows / 16;
    const int zw = zeros_width(IC, kW4Group);
    const int total = num_tiles * S * 16;
    if (total == 0) return cudaSuccess;
    repack_meta_kernel<<<(total + 127) / 128, 128, 0, stream>>>(segs[0], segs[nseg > 1 ? 1 : 0], segs[nseg > 2 ? 2 : 0], nseg, pair, NG, zw, zw * 8, S, num_tiles, out);
    (void)ctx;
    return cudaGetLastError();
}

cudaError_t encode_kv_tmap(CUtensorMap *out, const void *kv, long long rows) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    if (e != cudaSuccess || !sym) return e != cudaSuccess ? e : cudaErrorNotSupported;
    const cuuint64_t gdim[2] = {128, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {256};
    const cuuint32_t box[2] = {64, (cuuint32_t)kKvChunk};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = reinterpret_cast<EncodeFn>(sym)(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void *>(kv), gdim, gstride, box, estr,
                                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t launch(Ctx *ctx, const Args &a, cudaStream_t stream) {
    const size_t smem = smem_bytes(a);
    if ((int)smem > ctx->smem_optin) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(decode_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctx->num_sms);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs must be co-resident: the kernel synchronises grid-wide
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, decode_persistent_kernel, a);
}

}  // namespace pk
}  // namespace tce
