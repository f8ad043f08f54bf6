// persistent.h -- host/device interface of the persistent Llama decode kernel (decode_persistent.cu).
//
// One cooperative kernel per decoded token, one CTA per SM.  Everything a token needs from HBM -- packed weights, their repacked
// scales/zeros, and the K/V cache rows -- flows through ONE ring of TMA stages per CTA that is filled by a producer warp which depends
// on nothing but static data, so it runs ahead across all phase boundaries: while the GPU synchronises (grid barrier between the
// five phases of a layer) or stages activations, the next matrices keep streaming.
#pragma once
#include "kernels.h"

namespace tce {
namespace pk {

constexpr int kCW = 16;                          // consumer warps
constexpr int kThreads = 32 * (2 + kCW);         // producer warp + epilogue warp + consumers = 576
constexpr int kConsumerThreads = 32 * kCW;       // 512
constexpr int kStageBytes = 17408;               // 16 KiB weights / K / V + 640 B repacked scales|zeros, 1 KiB multiple (128B-swizzle atoms)
constexpr int kMetaOff = 16384;                  // scales half[16 groups][16 rows] (512 B) then zeros u64[16 groups] (nibble r = row r)
constexpr int kMetaBytes = 640;
constexpr int kMaxStages = 12;
constexpr int kRedBufs = 3;
constexpr int kKvChunk = 64;                     // cached positions per ring stage (64 x 256 B = 16 KiB of K or of V)

enum XMode : int { PX_HALF = 0, PX_RMS_F32 = 1, PX_EMBED_RMS = 2 };
enum Epi : int { PE_STORE_HALF = 0, PE_ADD_F32 = 1, PE_SILU_MUL = 2, PE_LOGITS = 3, PE_TP_SCATTER = 4 };
enum OpIdx : int { OPI_QKV = 0, OPI_O = 1, OPI_GATEUP = 2, OPI_DOWN = 3, OPI_LMHEAD = 4, OPI_COUNT = 5 };

// shape of one GEMV op; identical for every layer, so it lives in the kernel parameter block
struct GemvOp {
    int IC, NG;          // input channels, 128-groups per row
    int sg;              // groups per stage = min(16, NG): dense stage row pitch = sg * 64 B
    int S;               // stages per 16-row tile = ceil(NG / 16)
    int num_tiles;       // 16-row tiles (pair mode: 8 gate rows + 8 up rows)
    int SU;              // stage units = num_tiles * S
    int nseg, pair;      // row segments (q|k|v = 3); pair = gate/up interleave
    int rows0, rows1;    // rows of segments 0 and 1 (tile -> segment)
    int x_mode, epi;
    int aligned;         // 1: CTA ranges are cut at tile boundaries (one ordered writer per output)
    int box_bytes;       // bytes one stage's weight box(es) deliver = 16 * sg * 64
};

struct LayerDesc {       // per layer, global memory
    const uint8_t *meta[4];      // repacked scales|zeros of qkv, o, gate_up, down: [tile][S][640 B]
    const float *input_norm, *post_norm;
    __half *k_cache, *v_cache;   // [KVH][max_ctx][128] of this layer (append)
    int k_row0, v_row0;          // first row of this layer's K / V slab in the cache tensor map
    int pad[2];
};

struct Args {
    GemvOp op[OPI_COUNT];
    const LayerDesc *layers;
    int num_layers;
    const CUtensorMap *maps;     // [num_layers][7] (q k v o gate up down), then lm_head, then the KV cache map
    const uint8_t *lm_meta;
    const float *final_norm;
    // activations / state
    const __half *embed;         // [rows][E]
    int embed_rows;
    float *resid;                // [E] fp32 residual stream
    __half *qkv, *attn, *act;
    float *logits;
    const int *tokpos;           // {token, position}
    int *next_token;
    unsigned long long *argmax_cell;
    float *attn_ws;              // [H][nsplit_max][130] flash-decode partials
    unsigned *attn_cnt;          // [KVH] arrival counters (self re-arming)
    unsigned *sync;              // [5 * num_layers + 1] monotonic grid-barrier counters
    unsigned *epoch;             // launches completed so far (counter targets = (epoch + 1) * #CTAs)
    int *error;                  // device error word (bad token / position)
    const float *cos, *sin;      // [max_ctx][128]
    float alpha, eps;
    int H, KVH, nrep, max_ctx, E, V;
    int nst;                     // ring depth
    int xs_bytes;                // activation-plane buffer (also the attention scratch)
    int max_ng;
    // tensor parallel (tp_size > 1): see decode_persistent.cu
    int tp_size, tp_rank;
    float *tp_gather[kMaxTP];    // every rank's gather buffer [2][P][E] fp32
    unsigned *tp_arrive[kMaxTP]; // every rank's arrival counters [2]
    unsigned long long *tp_keys[kMaxTP];  // every rank's arg-max key slots [P]
    unsigned *tp_key_arrive[kMaxTP];
    int vocab_base;              // global index of this rank's first vocabulary row
    unsigned long long *dbg;     // optional (TCE_PK_DEBUG=1): globaltimer stamps [cta][phase][4] = barrier passed, staged, consumed, arrived
};

size_t smem_bytes(const Args &a);
// ring depth that fits `smem_optin` next to the fixed buffers (0 = does not fit)
int pick_stages(int smem_optin, int xs_bytes, int max_ng);
int attn_scratch_bytes(int nrep);
cudaError_t launch(Ctx *ctx, const Args &a, cudaStream_t stream);
// one-off repack of a (possibly multi-segment / gate-up paired) matrix's scales and zeros into per-stage records
cudaError_t repack_meta(Ctx *ctx, const W4Seg *segs, int nseg, int pair, int IC, uint8_t *out, cudaStream_t stream);
cudaError_t encode_kv_tmap(CUtensorMap *out, const void *kv, long long rows);

}  // namespace pk
}  // namespace tce
