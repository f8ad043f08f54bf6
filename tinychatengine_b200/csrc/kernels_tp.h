// kernels_tp.h -- launch interface of the tensor-parallel helper kernels (internal).
#pragma once
#include "kernels.h"

namespace tce {

struct TpSignalArgs {
    unsigned *peer_flag[kMaxTP];  // this rank's flag word inside every rank's buffer
    int tp_size;
    const int *step;              // device int: decode step index
    int k, per_step;              // flag value = step * per_step + k + 1
};
struct TpArgmaxArgs {
    const float *logits;          // local vocabulary shard
    int n_local, index_base;      // global index = index_base + i
    unsigned long long *peer_key[kMaxTP];  // this rank's key slot inside every rank's buffer
    int tp_size;
};
struct TpArgmaxFinishArgs {
    const unsigned long long *keys;  // local [tp_size]
    const unsigned *flags;           // local [tp_size]
    int *step;                       // device step counter, incremented here
    int k, per_step, tp_size;
    int *next_token;
};
cudaError_t launch_tp_signal(Ctx *ctx, const TpSignalArgs &a);
cudaError_t launch_tp_argmax_scatter(Ctx *ctx, const TpArgmaxArgs &a);
cudaError_t launch_tp_argmax_finish(Ctx *ctx, const TpArgmaxFinishArgs &a);

}  // namespace tce
