// decode_megakernel.cu -- one persistent kernel per decoded token (Llama, AWQ-INT4, batch 1) on sm_100a.
//
// The reference runs ~19 kernels + 128 memcpys per layer on stream 0; the graph path of this
// library runs 5 kernels per layer, and measurements (profiles/README.md) show ~3.5 us per kernel during which HBM
// idles (launch gap, cold start, first-byte latency, activation staging).  Here the whole token is ONE cooperative
// kernel of one CTA per SM whose warp roles persist across all phases:
//   * 1 producer warp streams the packed weights of phase after phase through the same 4-stage TMA ring.  It
//     depends on nothing but the (static) weights, so it runs ahead across phase boundaries: while the rest of the
//     GPU synchronises, the first 64 KiB/SM of the next matrix are already landing in shared memory.
//   * 8 consumer warps: per phase wait for the grid barrier (all earlier phases complete), stage + quantise the
//     activations (fused RMSNorm), run the integer-MMA GEMV, or execute attention / embedding / arg-max work items.
//   * 1 epilogue warp: fused epilogues (fp16 store, residual RED.ADD, SiLU*mul, fp32 logits), then signals the grid
//     barrier for the phase.
// Grid barrier = one arrival counter PER PHASE (red.release / ld.acquire at gpu scope); every CTA arrives exactly once
// per phase and phase p starts when counter[p-1] == #CTAs.  (A single running counter would be wrong: the epilogue
// warp of a CTA that owns no tile of a phase arrives for it immediately, possibly phases ahead of everybody else.)  All waits are bounded (trap after ~3 s) so a protocol bug is a launch failure, not a hung GPU.
#include <stdio.h>

#include "attention_impl.cuh"
#include "megakernel.h"
#include "w4a16_gemv_impl.cuh"

namespace tce {

namespace {

using namespace gemv;

constexpr int kCW = 8;
constexpr int kThreads = 32 * (kProducerWarps + 1 + kCW);  // 416
constexpr int kConsumerThreads = kCW * 32;                   // 256 == attn::kAttnThreads
static_assert(kConsumerThreads == attn::kAttnThreads, "attention items run on the consumer warps");

TCE_DEVINL void grid_wait(const unsigned *sync, unsigned target) {  // one thread
    const long long t0 = clock64();
    while (true) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(sync) : "memory");
        if (v >= target) break;
        if (clock64() - t0 > 6000000000LL) __trap();
    }
    // L1 is not coherent: buffers produced by other CTAs earlier in THIS kernel (residual, qkv, attention output...)
    // may still sit in this SM's L1 from a previous phase.  A gpu-scope fence makes ptxas emit CCTL.IVALL, which drops
    // every L1 line of the SM before the consumers (released by the named barrier that follows) read them.
    __threadfence();
}
TCE_DEVINL void grid_arrive(unsigned *sync) {  // one thread, after the role's writes were fenced
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(sync) : "memory");
}

TCE_DEVINL unsigned long long argmax_key(float v, int idx) {
    unsigned b = __float_as_uint(v);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // monotone map float -> uint
    return ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);  // ties: lowest index wins
}

template <int NREP>
TCE_DEVINL void attention_phase(const AttnDecodeArgs &at, uint8_t *asmem, uint64_t *abar, int *aflag, uint32_t &parity, int cta, int ncta, int ctid,
                                int pos) {
    const int T = pos + 1;
    const int nsplit = (T + at.chunk - 1) / at.chunk;
    const int items = at.num_kv_heads * nsplit;
    for (int it = cta; it < items; it += ncta) {
        const int kvh = it % at.num_kv_heads, split = it / at.num_kv_heads;
        attn::attn_item<NREP>(at, asmem, abar, aflag, parity, kvh, split, ctid, pos, [] { named_bar_sync(1, kConsumerThreads); });
    }
}

__global__ void __launch_bounds__(kThreads, 1) decode_megakernel(const __grid_constant__ MegaArgs m) {
    extern __shared__ __align__(128) uint8_t smem[];
    using L = Layout<1, kCW>;
    const Smem sm = carve<1, kCW>(smem, m.max_ic);
    uint8_t *asmem = smem + ((L::bytes(m.max_ic) + 127) & ~(size_t)127);
    uint64_t *abar = reinterpret_cast<uint64_t *>(asmem + attn::smem_bytes(m.attn_nrep, m.attn_chunk));
    int *aflag = reinterpret_cast<int *>(abar + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    if (tid == 0) {
        init_barriers<kCW>(sm);
        mbar_init(&abar[0], 1);
        mbar_init(&abar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    if (warp < kProducerWarps) {
        // ================= producers: weights of every GEMV phase, back to back =================
        RingState rs;
        const uint64_t policy = l2_policy_evict_first();
        for (int p = 0; p < m.nphases; p++) {
            if (m.phases[p].type != PH_GEMV) continue;
            produce(m.phases[p].g, sm, rs, cta, ncta, lane, policy);  // by reference: the TMA unit needs the map's global address
        }
        return;
    }
    if (warp == kProducerWarps) {
        // ================= epilogue warp =================
        RedState es;
        for (int p = 0; p < m.nphases; p++) {
            if (m.phases[p].type != PH_GEMV) continue;
            const KArgs a = m.phases[p].g;
            epilogue<1, kCW>(a, sm, es, cta, ncta, lane);
            __threadfence();
            __syncwarp();
            if (lane == 0) grid_arrive(m.sync + p);
        }
        return;
    }

    // ================= consumers =================
    const int ctid = tid - 32 * (kProducerWarps + 1);
    const int cw = warp - (kProducerWarps + 1);
    RingState rs;
    RedState cs;
    uint32_t aparity = 0;
    for (int p = 0; p < m.nphases; p++) {
        if (p > 0) {
            if (ctid == 0) grid_wait(m.sync + (p - 1), (unsigned)ncta);
            named_bar_sync(1, kConsumerThreads);
        }
        const int type = m.phases[p].type;
        if (type == PH_GEMV) {
            const KArgs a = m.phases[p].g;
            stage_activations<1, kCW>(a, sm, L::x_pitch(a.IC), ctid, cw, lane);
            consume<1, kCW>(a, sm, rs, cs, L::x_pitch(a.IC), cta, ncta, cw, lane);
            continue;  // the epilogue warp signals the barrier for GEMV phases
        }
        if (type == PH_EMBED) {
            // resid = (float) table[token]  (reference: CPU Embedding + float2half, cuda/Int4llamaDecoder.cu:62-69)
            const int tok = m.tokpos[0];
            const __half *row = m.embed + (size_t)tok * m.E;
            for (int i = cta * kConsumerThreads + ctid; i < m.E; i += ncta * kConsumerThreads) m.resid[i] = __half2float(row[i]);
            if (cta == 0 && ctid == 0) *m.argmax_cell = 0ull;
        } else if (type == PH_ATTN) {
            const AttnDecodeArgs at = m.phases[p].at;
            const int pos = m.tokpos[1];
            switch (m.attn_nrep) {
                case 1: attention_phase<1>(at, asmem, abar, aflag, aparity, cta, ncta, ctid, pos); break;
                case 2: attention_phase<2>(at, asmem, abar, aflag, aparity, cta, ncta, ctid, pos); break;
                case 4: attention_phase<4>(at, asmem, abar, aflag, aparity, cta, ncta, ctid, pos); break;
                default: attention_phase<8>(at, asmem, abar, aflag, aparity, cta, ncta, ctid, pos); break;
            }
        } else if (type == PH_ARGMAX) {
            float best = -INFINITY;
            int bi = 0x7fffffff;
            for (int i = cta * kConsumerThreads + ctid; i < m.V; i += ncta * kConsumerThreads) {
                const float v = m.logits[i];
                if (v > best || (v == best && i < bi)) {
                    best = v;
                    bi = i;
                }
            }
            unsigned long long key = (bi == 0x7fffffff) ? 0ull : argmax_key(best, bi);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other > key ? other : key;
            }
            if (lane == 0 && key) atomicMax(m.argmax_cell, key);
        }
        // phases executed by the consumers: everyone's writes fenced, then one arrival per CTA
        __threadfence();
        named_bar_sync(1, kConsumerThreads);
        if (ctid == 0) grid_arrive(m.sync + p);
    }
    // greedy token: decoded once every CTA has contributed its local maximum
    if (cta == 0 && ctid == 0 && m.next_token) {
        grid_wait(m.sync + (m.nphases - 1), (unsigned)ncta);
        const unsigned long long key = *reinterpret_cast<volatile unsigned long long *>(m.argmax_cell);
        *m.next_token = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
    }
}

}  // namespace

size_t megakernel_smem_bytes(int max_ic, int nrep, int chunk) {
    return ((Layout<1, kCW>::bytes(max_ic) + 127) & ~(size_t)127) + attn::smem_bytes(nrep, chunk) + 2 * sizeof(uint64_t) + 16;
}

cudaError_t megakernel_fill_gemv(Ctx *ctx, const W4GemvParams &p, MegaPhase *ph, int ncta) {
    KArgs &a = ph->g;
    ph->type = PH_GEMV;
    for (int i = 0; i < 3; i++) a.seg[i] = p.seg[i < p.nseg ? i : 0];
    a.nseg = p.nseg;
    a.pair_mode = p.pair_mode;
    a.IC = p.IC;
    a.NG = p.IC / kW4Group;
    a.zeros_w = zeros_width(p.IC, kW4Group);
    a.sf_w = a.zeros_w * 8;
    int rows = 0;
    for (int i = 0; i < p.nseg; i++) rows += p.seg[i].rows;
    a.num_tiles = rows / 16;
    a.M = 1;
    a.ldx = p.IC;
    a.x_mode = p.x_mode;
    a.x = p.x;
    a.gamma = p.gamma;
    a.eps = p.eps;
    a.y = p.y;
    a.epi = p.epi;
    a.ldy = p.ldy ? p.ldy : (p.pair_mode ? rows / 2 : rows);
    a.partials = ctx->gemv_partials;
    a.counters = ctx->gemv_counters;
    a.dbg = nullptr;
    a.pdl_early = 0;
    a.nst = kStages;
    a.atomic_add = (p.atomic_residual && p.epi == EPI_ADD_F32 && !p.pair_mode) ? 1 : 0;
    a.aligned = (!a.atomic_add && a.num_tiles >= ncta) ? 1 : 0;
    a.sg = a.NG < kStageGroups ? a.NG : kStageGroups;
    a.full = (a.NG % kStageGroups == 0) ? 1 : 0;
    a.tp_size = 1;
    a.tp_in = nullptr;
    a.tp_flags = nullptr;
    a.tp_step = nullptr;
    a.tp_k = 0;
    a.tp_per_step = 1;
    a.resid_out = nullptr;
    for (int i = 0; i < kMaxTP; i++) a.tp_out[i] = nullptr;
    a.tp_sig_counter = nullptr;
    for (int i = 0; i < kMaxTP; i++) a.tp_sig_flag[i] = nullptr;
    a.tp_sig_k = 0;
    for (int i = 0; i < p.nseg; i++) {
        cudaError_t e = encode_w4_tmap(&a.tmap[i], p.seg[i].w, p.seg[i].rows, p.IC, a.sg, p.pair_mode ? 8 : 16);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_megakernel(Ctx *ctx, const MegaArgs &m, cudaStream_t stream) {
    const size_t smem = megakernel_smem_bytes(m.max_ic, m.attn_nrep, m.attn_chunk);
    if ((int)smem > ctx->smem_optin) return cudaErrorInvalidConfiguration;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(decode_megakernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
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
    return cudaLaunchKernelEx(&cfg, decode_megakernel, m);
}

}  // namespace tce
