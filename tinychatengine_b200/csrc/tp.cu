// tp.cu -- the small kernels of the KERNEL-PER-OP tensor-parallel decode path (the fallback for shapes outside the persistent kernel's
// envelope; Llama-3-8B and the other benchmarked geometries run the persistent kernel, where the all-reduce is peer stores of tagged words
// inside decode_persistent.cu and none of these kernels is launched).  No NCCL on the data path.
//
// Collective = all-reduce of the row-parallel GEMV outputs (o_proj, down_proj), realised as
//   (1) the GEMV epilogue storing its finished fp32 outputs into slot `rank` of EVERY rank's gather buffer
//       (EPI_TP_SCATTER_F32, w4a16_gemv_impl.cuh: peer stores, tile by tile while the GEMV is still running),
//   (2) one release-store of a sequence number into every peer's flag word once the GEMV has finished: by the last CTA of the GEMV itself
//       (fused signal, w4a16_gemv_impl.cuh) -- tp_signal_kernel below is the stand-alone form kept for the arg-max exchange -- and
//   (3) the next GEMV's prologue: acquire-poll the P local flags, then residual += sum over ranks in rank order
//       (X_RMSNORM_F32 + tp_in).
// The greedy token needs one more exchange: every rank scatters the arg-max key of its vocabulary shard.
#include "common.cuh"
#include "kernels_tp.h"

namespace tce {
namespace {

__global__ void tp_signal_kernel(const TpSignalArgs a) {
    if (threadIdx.x == 0) {
        const unsigned v = (unsigned)(*a.step) * (unsigned)a.per_step + (unsigned)a.k + 1u;
        __threadfence_system();
        for (int p = 0; p < a.tp_size; p++) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_flag[p]), "r"(v) : "memory");
    }
}

TCE_DEVINL unsigned long long key_of(float v, int idx) {
    unsigned b = __float_as_uint(v);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);
}

// one block: arg-max of the local logits shard, key (value, GLOBAL index) stored into every rank's key slot
__global__ void __launch_bounds__(1024) tp_argmax_scatter_kernel(const TpArgmaxArgs a) {
    __shared__ unsigned long long sk[32];
    unsigned long long best = 0ull;
    for (int i = threadIdx.x; i < a.n_local; i += blockDim.x) {
        const unsigned long long k = key_of(a.logits[i], a.index_base + i);
        best = k > best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) sk[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        best = threadIdx.x < (blockDim.x >> 5) ? sk[threadIdx.x] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (threadIdx.x == 0)
            for (int p = 0; p < a.tp_size; p++) a.peer_key[p][0] = best;
    }
}

__global__ void tp_argmax_finish_kernel(const TpArgmaxFinishArgs a) {
    if (threadIdx.x == 0) {
        const unsigned expect = (unsigned)(*a.step) * (unsigned)a.per_step + (unsigned)a.k + 1u;
        unsigned long long best = 0ull;
        for (int p = 0; p < a.tp_size; p++) {
            const long long t0 = clock64();
            while (true) {
                unsigned f;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(a.flags + p) : "memory");
                if (f >= expect) break;
                if (clock64() - t0 > 6000000000LL) __trap();
            }
            const unsigned long long k = *reinterpret_cast<const volatile unsigned long long *>(a.keys + p);
            best = k > best ? k : best;
        }
        *a.next_token = (int)(0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFull));
        // last tensor-parallel op of the step: advance the step counter ON THE DEVICE, so that any re-execution (warm-up
        // run, graph replay) uses fresh flag values.  A replayed index would leave every arrival flag already satisfied,
        // ranks could drift apart and overwrite gather buffers a slower peer is still reading.
        *a.step = *a.step + 1;
    }
}

}  // namespace

cudaError_t launch_tp_signal(Ctx *ctx, const TpSignalArgs &a) {
    tp_signal_kernel<<<1, 32, 0, ctx->stream>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_tp_argmax_scatter(Ctx *ctx, const TpArgmaxArgs &a) {
    tp_argmax_scatter_kernel<<<1, 1024, 0, ctx->stream>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_tp_argmax_finish(Ctx *ctx, const TpArgmaxFinishArgs &a) {
    tp_argmax_finish_kernel<<<1, 32, 0, ctx->stream>>>(a);
    return cudaGetLastError();
}

}  // namespace tce
