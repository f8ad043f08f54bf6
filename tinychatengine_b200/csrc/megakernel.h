// megakernel.h -- host/device interface of the persistent decode kernel (decode_megakernel.cu).
#pragma once
#include "kernels.h"
#include "kernels_attn.h"
#include "w4a16_gemv_impl.cuh"

namespace tce {

enum MegaPhaseType : int { PH_EMBED = 0, PH_GEMV = 1, PH_ATTN = 2, PH_ARGMAX = 3 };

struct alignas(64) MegaPhase {
    int type;
    int pad[15];
    gemv::KArgs g;      // PH_GEMV
    AttnDecodeArgs at;  // PH_ATTN
};

struct MegaArgs {
    const MegaPhase *phases;  // device array
    int nphases;
    unsigned *sync;           // grid barrier counters [nphases], zeroed before every launch
    const int *tokpos;        // {token, position}
    const __half *embed;      // [vocab][E]
    float *resid;             // [E]
    int E;
    const float *logits;      // [V]
    int V;
    int *next_token;
    unsigned long long *argmax_cell;
    int max_ic;               // largest IC of any GEMV phase (shared-memory carve)
    int attn_nrep, attn_chunk;
};

size_t megakernel_smem_bytes(int max_ic, int nrep, int chunk);
cudaError_t megakernel_fill_gemv(Ctx *ctx, const W4GemvParams &p, MegaPhase *ph, int ncta);
cudaError_t launch_megakernel(Ctx *ctx, const MegaArgs &m, cudaStream_t stream);

}  // namespace tce
