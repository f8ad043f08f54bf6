// w4a16_gemv.cu -- W4A16 group-128 GEMV for batch-1..8 decode on sm_100a (HBM-bound).
//
// Replaces MatmulOperator::gemv_forward_cuda + gemv_kernel_g128 (reference kernels/cuda/gemv_cuda.cu:140-260)
// and consumes the reference's QM_CUDA on-disk layout unchanged (llm/tools/quantize_methods.py:370-442):
//     w uint32[OC][IC/8] sequential nibbles, zeros uint32[OC][zeros_w] (nibble g = zero of group g),
//     scales half[OC][zeros_w*8];   y[m][oc] = sum_ic s[oc,g] * (q[oc,ic] - z[oc,g]) * x[m][ic]
//
// Design (see DESIGN.md "W4A16 decode GEMV"):
//  * work unit = (16-row tile, one 128-k group) = 1 KiB of packed weights; the units of a launch are cut
//    into equal contiguous ranges, one per CTA (stream-K), so every CTA streams the same number of bytes no
//    matter the shape; a row tile split between CTAs is finished by whichever CTA arrives last (fixed
//    summation order -> deterministic), no atomics on data, no pre-zeroed outputs.
//  * a producer warp streams [16 rows x <=16 groups] weight slabs with 1-D TMA bulk copies (UBLKCP) into a
//    4-stage shared-memory ring guarded by full/empty mbarriers; weights are tagged L2 evict-first.
//    The ring starts filling BEFORE griddepcontrol.wait, so under programmatic dependent launch the next
//    GEMV's weights are already in flight while the previous kernel drains.
//  * 8 consumer warps: 128-bit conflict-free LDS of packed nibbles (row pitch = 64 mod 128 B), int4 -> fp16
//    by lop3 + magic-number subtract (exact, centred at 8), products on the legacy tensor path
//    (mma.sync m16n8k16, fp32 accumulate; the 8 MMA columns are the <=8 activation rows), per-group epilogue
//    tot += s * (acc - (z-8) * sum_x) in fp32.
//  * optional fused prologue (RMSNorm of an fp32 residual stream) and epilogues (fp16/fp32 store,
//    residual += , SiLU(gate)*up with gate/up rows paired inside one MMA tile).
#include <stdio.h>

#include "common.cuh"
#include "kernels.h"

namespace tce {

namespace {

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = 32 + kConsumerThreads;  // warp 0 = TMA producer
constexpr int kStageGroups = 16;                 // 128-k groups per pipeline stage (per row: 1024 B)
constexpr int kRowPitch = kStageGroups * 64 + 64;  // 1088 B: == 64 (mod 128) -> conflict-free LDS.128
constexpr int kStageBytes = 16 * kRowPitch;        // 17408 B
constexpr int kStages = 4;

struct KArgs {
    W4Seg seg[3];
    int nseg, pair_mode;
    int IC, NG, zeros_w, sf_w;
    int num_tiles;
    int M, ldx, x_mode;
    const void *x;
    const float *gamma;
    float eps;
    void *y;
    int epi, ldy;
    float *partials;
    unsigned *counters;
};

struct RowRef {
    const uint8_t *w;
    const uint32_t *z;
    const __half *s;
};

// source row `l` (0..15) of row tile `rt`
TCE_DEVINL RowRef tile_row(const KArgs &a, int rt, int l) {
    int si = 0, r;
    if (a.pair_mode) {
        si = l >> 3;
        r = rt * 8 + (l & 7);
    } else {
        r = rt * 16 + l;
        if (a.nseg > 1 && r >= a.seg[0].rows) {
            r -= a.seg[0].rows;
            si = 1;
            if (a.nseg > 2 && r >= a.seg[1].rows) {
                r -= a.seg[1].rows;
                si = 2;
            }
        }
    }
    const W4Seg &s = a.seg[si];
    RowRef ref;
    ref.w = reinterpret_cast<const uint8_t *>(s.w) + (size_t)r * (a.IC / 2);
    ref.z = s.zeros + (size_t)r * a.zeros_w;
    ref.s = s.scales + (size_t)r * a.sf_w;
    return ref;
}

template <int NCOLS>
struct Smem {
    static constexpr int kXPad = (NCOLS > 1) ? 64 : 0;  // column pitch = 64 (mod 128) B for the per-column B loads
    static __host__ __device__ int x_pitch(int IC) { return IC * 2 + kXPad; }
    static __host__ __device__ size_t off_xs() { return (size_t)kStages * kStageBytes; }
    static __host__ __device__ size_t off_gx(int IC) { return off_xs() + (size_t)NCOLS * x_pitch(IC); }
    static __host__ __device__ size_t off_red(int IC) { return off_gx(IC) + (size_t)NCOLS * (IC / 128) * sizeof(float); }
    static __host__ __device__ size_t off_rms(int IC) { return off_red(IC) + (size_t)2 * kConsumerWarps * 16 * NCOLS * sizeof(float); }
    static __host__ __device__ size_t off_bar(int IC) {
        return (off_rms(IC) + (size_t)NCOLS * kConsumerWarps * sizeof(float) + 15) & ~(size_t)15;
    }
    static __host__ __device__ size_t bytes(int IC) { return off_bar(IC) + 2 * kStages * sizeof(uint64_t) + 16; }
};

template <int NCOLS>
__global__ void __launch_bounds__(kThreads, 1) w4a16_gemv_kernel(const __grid_constant__ KArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    using SM = Smem<NCOLS>;
    uint8_t *stages = smem;
    uint8_t *xs = smem + SM::off_xs();
    float *gx = reinterpret_cast<float *>(smem + SM::off_gx(a.IC));
    float *red = reinterpret_cast<float *>(smem + SM::off_red(a.IC));
    float *rms = reinterpret_cast<float *>(smem + SM::off_rms(a.IC));
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + SM::off_bar(a.IC));
    uint64_t *empty_bar = full_bar + kStages;
    int *flag = reinterpret_cast<int *>(empty_bar + kStages);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    // let the next kernel in the stream become resident right away: it may only prefetch its (static) weights
    // until its own griddepcontrol.wait releases, which happens when this whole grid has finished.
    pdl_launch_dependents();

    StreamK sk;
    sk.U = (long long)a.num_tiles * a.NG;
    sk.nc = gridDim.x;
    sk.NG = a.NG;
    const long long u0 = sk.start(blockIdx.x), u1 = sk.start(blockIdx.x + 1);

    if (warp == 0) {
        // =========================== TMA producer ===========================
        const uint64_t policy = l2_policy_evict_first();
        int stage = 0;
        uint32_t phase = 0;
        long long u = u0;
        while (u < u1) {
            const int rt = (int)(u / a.NG);
            const int gb = (int)(u % a.NG);
            const int ge = (int)min((long long)a.NG, gb + (u1 - u));
            const uint8_t *src = nullptr;
            if (lane < 16) src = tile_row(a, rt, lane).w;
            for (int g0 = gb; g0 < ge; g0 += kStageGroups) {
                const int n = min(kStageGroups, ge - g0);
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (lane == 0) mbar_arrive_expect_tx(&full_bar[stage], 16u * n * 64u);
                __syncwarp();
                if (lane < 16)
                    bulk_g2s(stages + (size_t)stage * kStageBytes + lane * kRowPitch, src + (size_t)g0 * 64, n * 64, &full_bar[stage],
                             policy);
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            u += ge - gb;
        }
        return;
    }

    // =========================== consumers ===========================
    const int ctid = tid - 32;  // 0..255
    const int cw = warp - 1;    // 0..7
    const int g = lane >> 2, t = lane & 3;

    pdl_wait();  // activations (and any buffer we write) belong to the previous kernel until here

    // ---- activation prologue: x -> fp16, permuted into MMA-B order in smem, plus per-group sums ----
    {
        const int units = a.IC / 8;
#pragma unroll 1
        for (int col = 0; col < NCOLS; col++) {
            uint8_t *xcol = xs + (size_t)col * SM::x_pitch(a.IC);
            float inv = 1.f;
            if (a.x_mode == X_RMSNORM_F32 && col < a.M) {
                const float *xr = reinterpret_cast<const float *>(a.x) + (size_t)col * a.ldx;
                float ss = 0.f;
                for (int ui = ctid; ui < units; ui += kConsumerThreads) {
                    float4 v0 = *reinterpret_cast<const float4 *>(xr + ui * 8);
                    float4 v1 = *reinterpret_cast<const float4 *>(xr + ui * 8 + 4);
                    ss += v0.x * v0.x + v0.y * v0.y + v0.z * v0.z + v0.w * v0.w;
                    ss += v1.x * v1.x + v1.y * v1.y + v1.z * v1.z + v1.w * v1.w;
                }
                ss = warp_sum(ss);
                if (lane == 0) rms[col * kConsumerWarps + cw] = ss;
                named_bar_sync(1, kConsumerThreads);
                float tot = 0.f;
#pragma unroll
                for (int w = 0; w < kConsumerWarps; w++) tot += rms[col * kConsumerWarps + w];
                inv = rsqrtf(tot / (float)a.IC + a.eps);
            }
            // `units` is a multiple of 16, not of 32: the trip count is made warp-uniform so that the half-warp
            // shuffles below always run with all 32 lanes (an idle half-warp contributes zeros to nobody)
            for (int ui0 = 0; ui0 < units; ui0 += kConsumerThreads) {
                const int ui = ui0 + ctid;
                const bool valid = ui < units;
                float v[8];
                if (valid && col < a.M) {
                    if (a.x_mode == X_RMSNORM_F32) {
                        const float *xr = reinterpret_cast<const float *>(a.x) + (size_t)col * a.ldx + ui * 8;
                        float4 v0 = *reinterpret_cast<const float4 *>(xr);
                        float4 v1 = *reinterpret_cast<const float4 *>(xr + 4);
                        float4 g0 = *reinterpret_cast<const float4 *>(a.gamma + ui * 8);
                        float4 g1 = *reinterpret_cast<const float4 *>(a.gamma + ui * 8 + 4);
                        v[0] = (v0.x * inv) * g0.x; v[1] = (v0.y * inv) * g0.y; v[2] = (v0.z * inv) * g0.z; v[3] = (v0.w * inv) * g0.w;
                        v[4] = (v1.x * inv) * g1.x; v[5] = (v1.y * inv) * g1.y; v[6] = (v1.z * inv) * g1.z; v[7] = (v1.w * inv) * g1.w;
                    } else {
                        const __half *xr = reinterpret_cast<const __half *>(a.x) + (size_t)col * a.ldx + ui * 8;
                        uint4 raw = *reinterpret_cast<const uint4 *>(xr);
                        const __half2 *h2 = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            float2 f = __half22float2(h2[i]);
                            v[2 * i] = f.x;
                            v[2 * i + 1] = f.y;
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) v[i] = 0.f;
                }
                // MMA-B order inside a 16-byte unit: (x0,x4)(x1,x5)(x2,x6)(x3,x7); units of a group are stored
                // j-major so that the four t-lanes of one LDS.128 hit consecutive 16-B slots.
                uint4 o;
                o.x = pack_half2(v[0], v[4]);
                o.y = pack_half2(v[1], v[5]);
                o.z = pack_half2(v[2], v[6]);
                o.w = pack_half2(v[3], v[7]);
                const int G = ui >> 4, tj = ui & 15;  // ui = G*16 + 4*t + j
                const int pos = G * 16 + (tj & 3) * 4 + (tj >> 2);
                if (valid) *reinterpret_cast<uint4 *>(xcol + (size_t)pos * 16) = o;
                // group sum of the fp16-rounded values the tensor core will actually see
                const __half2 *oh = reinterpret_cast<const __half2 *>(&o);
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    float2 f = __half22float2(oh[i]);
                    s += f.x + f.y;
                }
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                if (valid && (lane & 15) == 0) gx[col * a.NG + G] = s;
            }
        }
        named_bar_sync(1, kConsumerThreads);
    }

    int stage = 0;
    uint32_t phase = 0;
    int flush_idx = 0;
    long long u = u0;
    while (u < u1) {
        const int rt = (int)(u / a.NG);
        const int gb = (int)(u % a.NG);
        const int ge = (int)min((long long)a.NG, gb + (u1 - u));
        const RowRef rA = tile_row(a, rt, g), rB = tile_row(a, rt, g + 8);
        float tot[(NCOLS == 1) ? 2 : 4];
#pragma unroll
        for (int i = 0; i < ((NCOLS == 1) ? 2 : 4); i++) tot[i] = 0.f;

        for (int g0 = gb; g0 < ge; g0 += kStageGroups) {
            const int n = min(kStageGroups, ge - g0);
            // scales / zeros for this warp's (up to two) groups: issued before the barrier wait
            float sA[2], sB[2], zA[2], zB[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int gi = cw + q * kConsumerWarps;
                if (gi < n) {
                    const int G = g0 + gi;
                    sA[q] = __half2float(__ldg(rA.s + G));
                    sB[q] = __half2float(__ldg(rB.s + G));
                    zA[q] = (float)((__ldg(rA.z + (G >> 3)) >> ((G & 7) * 4)) & 0xF) - 8.f;
                    zB[q] = (float)((__ldg(rB.z + (G >> 3)) >> ((G & 7) * 4)) & 0xF) - 8.f;
                }
            }
            mbar_wait(&full_bar[stage], phase);
            const uint8_t *sbase = stages + (size_t)stage * kStageBytes;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int gi = cw + q * kConsumerWarps;
                if (gi < n) {
                    const int G = g0 + gi;
                    const uint8_t *sp = sbase + gi * 64 + t * 16;
                    const uint4 wa = *reinterpret_cast<const uint4 *>(sp + g * kRowPitch);
                    const uint4 wb = *reinterpret_cast<const uint4 *>(sp + (g + 8) * kRowPitch);
                    const uint32_t wav[4] = {wa.x, wa.y, wa.z, wa.w};
                    const uint32_t wbv[4] = {wb.x, wb.y, wb.z, wb.w};
                    const uint8_t *xp = xs + ((NCOLS == 1) ? 0 : (size_t)g * SM::x_pitch(a.IC)) + ((size_t)G * 16 + t) * 16;
                    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint4 xv = *reinterpret_cast<const uint4 *>(xp + j * 64);
                        const uint32_t a_w = wav[j], b_w = wbv[j];
                        const uint32_t a_w8 = a_w >> 8, b_w8 = b_w >> 8;
                        // nibble -> fp16: low nibbles ride on 1024 (0x6400), high nibbles on 64 (0x5400);
                        // subtracting 1032 / 72 yields q - 8 exactly.
                        const uint32_t p0a = hsub2_u32(lop3_and_or(a_w, 0x000f000fu, 0x64006400u), 0x64086408u);
                        const uint32_t p1a = hsub2_u32(lop3_and_or(a_w, 0x00f000f0u, 0x54005400u), 0x54805480u);
                        const uint32_t p2a = hsub2_u32(lop3_and_or(a_w8, 0x000f000fu, 0x64006400u), 0x64086408u);
                        const uint32_t p3a = hsub2_u32(lop3_and_or(a_w8, 0x00f000f0u, 0x54005400u), 0x54805480u);
                        const uint32_t p0b = hsub2_u32(lop3_and_or(b_w, 0x000f000fu, 0x64006400u), 0x64086408u);
                        const uint32_t p1b = hsub2_u32(lop3_and_or(b_w, 0x00f000f0u, 0x54005400u), 0x54805480u);
                        const uint32_t p2b = hsub2_u32(lop3_and_or(b_w8, 0x000f000fu, 0x64006400u), 0x64086408u);
                        const uint32_t p3b = hsub2_u32(lop3_and_or(b_w8, 0x00f000f0u, 0x54005400u), 0x54805480u);
                        mma_m16n8k16(c, p0a, p0b, p1a, p1b, xv.x, xv.y);
                        mma_m16n8k16(c, p2a, p2b, p3a, p3b, xv.z, xv.w);
                    }
                    if (NCOLS == 1) {
                        const float gxv = gx[G];
                        tot[0] += sA[q] * (c[0] - zA[q] * gxv);
                        tot[1] += sB[q] * (c[2] - zB[q] * gxv);
                    } else {
                        const float gx0 = gx[(2 * t) * a.NG + G], gx1 = gx[(2 * t + 1) * a.NG + G];
                        tot[0] += sA[q] * (c[0] - zA[q] * gx0);
                        tot[1] += sA[q] * (c[1] - zA[q] * gx1);
                        tot[2] += sB[q] * (c[2] - zB[q] * gx0);
                        tot[3] += sB[q] * (c[3] - zB[q] * gx1);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == kStages) {
                stage = 0;
                phase ^= 1;
            }
        }

        // ---------------- flush the tile: cross-warp reduce, then epilogue or stream-K fix-up ----------------
        float *rbuf = red + (size_t)(flush_idx & 1) * kConsumerWarps * 16 * NCOLS + (size_t)cw * 16 * NCOLS;
        flush_idx++;
        if (NCOLS == 1) {
            if (t == 0) {
                rbuf[g] = tot[0];
                rbuf[g + 8] = tot[1];
            }
        } else {
            rbuf[g * NCOLS + 2 * t] = tot[0];
            rbuf[g * NCOLS + 2 * t + 1] = tot[1];
            rbuf[(g + 8) * NCOLS + 2 * t] = tot[2];
            rbuf[(g + 8) * NCOLS + 2 * t + 1] = tot[3];
        }
        named_bar_sync(1, kConsumerThreads);

        constexpr int kVals = 16 * NCOLS;                 // values per tile
        constexpr int kWriterWarps = (kVals + 31) / 32;   // consumer warps 0..kWriterWarps-1 own the flush
        const bool full_tile = (gb == 0 && ge == a.NG);
        if (cw < kWriterWarps) {
            auto writer_sync = [&]() {
                if (kWriterWarps > 1)
                    named_bar_sync(2, kWriterWarps * 32);
                else
                    __syncwarp();
            };
            float *rb0 = red + (size_t)((flush_idx - 1) & 1) * kConsumerWarps * 16 * NCOLS;
            float v = 0.f;
            if (ctid < kVals) {
#pragma unroll
                for (int w = 0; w < kConsumerWarps; w++) v += rb0[w * kVals + ctid];
            }
            bool do_final = full_tile;
            if (!full_tile) {
                // stream-K fix-up: park the partial, the last contributor to arrive sums all of them in CTA order
                const long long tb = (long long)rt * a.NG;
                const int c_first = sk.cta_of(tb);
                const int c_last = sk.cta_of(tb + a.NG - 1);
                const int rec = (u0 >= tb) ? 0 : 1;  // 0: this tile holds my first unit, 1: it is my tail tile
                float *mine = a.partials + ((size_t)blockIdx.x * 2 + rec) * kVals;
                if (ctid < kVals) mine[ctid] = v;
                __threadfence();
                writer_sync();
                if (ctid == 0) {
                    const unsigned prev = atomicAdd(&a.counters[rt], 1u);
                    const int last = (prev == (unsigned)(c_last - c_first)) ? 1 : 0;
                    if (last) a.counters[rt] = 0;  // every contributor has arrived: re-arm for the next launch
                    *flag = last;
                }
                writer_sync();
                do_final = (*reinterpret_cast<volatile int *>(flag) != 0);
                if (do_final) {
                    __threadfence();
                    if (ctid < kVals) {
                        v = 0.f;
                        for (int c = c_first; c <= c_last; c++) {
                            const int r = (sk.start(c) >= tb) ? 0 : 1;
                            v += ldg_cg_f32(a.partials + ((size_t)c * 2 + r) * kVals + ctid);
                        }
                    }
                }
            }
            if (a.pair_mode) {
                // rows 0-7 = gate, rows 8-15 = up of the same output channel: exchange through the (now idle)
                // reduction buffer of this flush, then y = SiLU(gate) * up   (reference: SiLuMul_half,
                // llm/src/nn_modules/cuda/Int4llamaDecoderLayer.cu:21-30, evaluated here in fp32)
                writer_sync();
                if (do_final && ctid < kVals) rb0[ctid] = v;
                writer_sync();
                if (do_final && ctid < 8 * NCOLS) {
                    const int row = ctid / NCOLS, col = ctid % NCOLS;
                    if (col < a.M) {
                        const float gte = rb0[row * NCOLS + col], up = rb0[(row + 8) * NCOLS + col];
                        const float act = gte / (1.f + __expf(-gte));
                        reinterpret_cast<__half *>(a.y)[(size_t)col * a.ldy + (size_t)rt * 8 + row] = __float2half(act * up);
                    }
                }
            } else if (do_final && ctid < kVals) {
                const int row = ctid / NCOLS, col = ctid % NCOLS;
                if (col < a.M) {
                    const size_t o = (size_t)col * a.ldy + (size_t)rt * 16 + row;
                    if (a.epi == EPI_STORE_HALF)
                        reinterpret_cast<__half *>(a.y)[o] = __float2half(v);
                    else if (a.epi == EPI_STORE_F32)
                        reinterpret_cast<float *>(a.y)[o] = v;
                    else
                        reinterpret_cast<float *>(a.y)[o] += v;
                }
            }
        }
        u += ge - gb;
    }
}

// ------------------------------------------------------------------------------------------------------
// Simple cross-check kernel: one warp per output row, 128-bit loads, scalar fp32 math.  Not the product
// path for performance; kept as an independently-written second implementation (TCE_GEMV_IMPL=0).
// ------------------------------------------------------------------------------------------------------
__global__ void w4a16_gemv_simple_kernel(const KArgs a, int total_rows) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= total_rows) return;
    int si = 0, r = row;
    if (a.nseg > 1 && r >= a.seg[0].rows) {
        r -= a.seg[0].rows;
        si = 1;
        if (a.nseg > 2 && r >= a.seg[1].rows) {
            r -= a.seg[1].rows;
            si = 2;
        }
    }
    const W4Seg &s = a.seg[si];
    const uint4 *wrow = reinterpret_cast<const uint4 *>(s.w + (size_t)r * (a.IC / 8));
    const uint32_t *zrow = s.zeros + (size_t)r * a.zeros_w;
    const __half *srow = s.scales + (size_t)r * a.sf_w;
    for (int m = 0; m < a.M; m++) {
        const __half *x = reinterpret_cast<const __half *>(a.x) + (size_t)m * a.ldx;
        float acc = 0.f;
        for (int c = lane; c < a.IC / 32; c += 32) {  // 32 nibbles per uint4
            const uint4 wv = wrow[c];
            const int G = c >> 2;
            const float sc = __half2float(srow[G]);
            const float z = (float)((zrow[G >> 3] >> ((G & 7) * 4)) & 0xF);
            const uint32_t words[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const float q = (float)((words[j] >> (4 * i)) & 0xF);
                    acc += (sc * (q - z)) * __half2float(x[c * 32 + j * 8 + i]);
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            const size_t o = (size_t)m * a.ldy + row;
            if (a.epi == EPI_STORE_HALF)
                reinterpret_cast<__half *>(a.y)[o] = __float2half(acc);
            else if (a.epi == EPI_STORE_F32)
                reinterpret_cast<float *>(a.y)[o] = acc;
            else
                reinterpret_cast<float *>(a.y)[o] += acc;
        }
    }
}

KArgs make_kargs(Ctx *ctx, const W4GemvParams &p, int *total_rows) {
    KArgs a;
    for (int i = 0; i < 3; i++) a.seg[i] = p.seg[i < p.nseg ? i : 0];
    a.nseg = p.nseg;
    a.pair_mode = p.pair_mode;
    a.IC = p.IC;
    a.NG = p.IC / kW4Group;
    a.zeros_w = zeros_width(p.IC, kW4Group);
    a.sf_w = a.zeros_w * 8;
    int rows = 0;
    for (int i = 0; i < p.nseg; i++) rows += p.seg[i].rows;
    *total_rows = rows;
    a.num_tiles = rows / 16;
    a.M = p.M;
    a.ldx = p.ldx ? p.ldx : p.IC;
    a.x_mode = p.x_mode;
    a.x = p.x;
    a.gamma = p.gamma;
    a.eps = p.eps;
    a.y = p.y;
    a.epi = p.epi;
    a.ldy = p.ldy ? p.ldy : (p.pair_mode ? rows / 2 : rows);
    a.partials = ctx->gemv_partials;
    a.counters = ctx->gemv_counters;
    return a;
}

template <int NCOLS>
cudaError_t launch_mma(Ctx *ctx, const KArgs &a, bool pdl) {
    const size_t smem = Smem<NCOLS>::bytes(a.IC);
    if ((int)smem > ctx->smem_optin) return cudaErrorInvalidConfiguration;
    static bool attr_set = false;  // per template instantiation
    static size_t attr_smem = 0;
    if (!attr_set || smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(w4a16_gemv_kernel<NCOLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin);
        if (e != cudaSuccess) return e;
        attr_set = true;
        attr_smem = ctx->smem_optin;
    }
    const long long U = (long long)a.num_tiles * a.NG;
    int nc = ctx->num_sms * ctx->gemv_ctas_per_sm;
    if (nc > ctx->gemv_max_ctas) nc = ctx->gemv_max_ctas;
    if ((long long)nc > U) nc = (int)U;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nc);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, w4a16_gemv_kernel<NCOLS>, a);
}

}  // namespace

cudaError_t launch_w4a16_gemv_simple(Ctx *ctx, const W4GemvParams &p) {
    if (p.pair_mode || p.x_mode != X_HALF) return cudaErrorNotSupported;
    int rows;
    KArgs a = make_kargs(ctx, p, &rows);
    const int warps = 8;
    w4a16_gemv_simple_kernel<<<(rows + warps - 1) / warps, warps * 32, 0, ctx->stream>>>(a, rows);
    return cudaGetLastError();
}

cudaError_t launch_w4a16_gemv(Ctx *ctx, const W4GemvParams &p) {
    if (p.M < 1 || p.M > 8 || p.IC % kW4Group || p.nseg < 1 || p.nseg > 3) return cudaErrorInvalidValue;
    int rows;
    KArgs a = make_kargs(ctx, p, &rows);
    for (int i = 0; i < p.nseg; i++)
        if (p.seg[i].rows % (p.pair_mode ? 8 : 16)) return cudaErrorInvalidValue;
    if (p.pair_mode && (p.nseg != 2 || p.seg[0].rows != p.seg[1].rows)) return cudaErrorInvalidValue;
    if (a.num_tiles > ctx->gemv_max_tiles) return cudaErrorInvalidValue;
    if (p.M == 1) return launch_mma<1>(ctx, a, p.pdl);
    return launch_mma<8>(ctx, a, p.pdl);
}

}  // namespace tce
