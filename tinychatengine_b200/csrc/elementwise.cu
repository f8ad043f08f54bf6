// elementwise.cu -- the small ops either side of the hot path: embedding gather, RMSNorm,
// argmax.  RMSNorm / residual add / SiLU*mul of the decode block are fused into the GEMV kernels
// (w4a16_gemv.cu); the standalone versions here serve the op-level API and the tests.
#include "common.cuh"
#include "kernels_attn.h"

namespace tce {
namespace {

// `guard` (optional): {rows of the table, max_ctx} bound the device-resident {token, position}; out-of-range values are clamped before any kernel of the
// step uses them (the attention kernel reads the position from safe[1]) and flagged in safe[2], so a bad id can never index past the table / KV slab
__global__ void embedding_kernel(const __half *__restrict__ table, const int *__restrict__ token, float *__restrict__ resid, int E, int rows, int max_ctx,
                                 int *__restrict__ safe) {
    pdl_launch_dependents();
    pdl_wait();
    int tok = *token;
    if (safe) {
        const int pos = token[1];
        const bool bad = tok < 0 || tok >= rows || pos < 0 || pos >= max_ctx;
        tok = tok < 0 ? 0 : (tok >= rows ? rows - 1 : tok);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            safe[0] = tok;
            safe[1] = pos < 0 ? 0 : (pos >= max_ctx ? max_ctx - 1 : pos);
            safe[2] = bad ? 1 : 0;
        }
    }
    const __half2 *row = reinterpret_cast<const __half2 *>(table + (size_t)tok * E);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < E / 2; i += gridDim.x * blockDim.x) {
        const float2 f = __half22float2(row[i]);
        reinterpret_cast<float2 *>(resid)[i] = f;
    }
}

// two-phase argmax with a last-block finish; ties resolve to the lowest index
struct ArgmaxWs {
    float val[256];
    int idx[256];
    unsigned counter;
};
__device__ ArgmaxWs g_argmax_ws;

__global__ void argmax_kernel(const float *__restrict__ x, int n, int *__restrict__ out) {
    __shared__ float sval[32];
    __shared__ int sidx[32];
    __shared__ int is_last;
    pdl_launch_dependents();
    pdl_wait();
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float v = x[i];
        if (v > best || (v == best && i < bi)) {
            best = v;
            bi = i;
        }
    }
    auto combine = [](float &bv, int &bidx, float ov, int oidx) {
        if (ov > bv || (ov == bv && oidx < bidx)) {
            bv = ov;
            bidx = oidx;
        }
    };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) combine(best, bi, __shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, bi, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sval[warp] = best;
        sidx[warp] = bi;
    }
    __syncthreads();
    if (warp == 0) {
        best = (lane < (blockDim.x >> 5)) ? sval[lane] : -INFINITY;
        bi = (lane < (blockDim.x >> 5)) ? sidx[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) combine(best, bi, __shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, bi, o));
        if (lane == 0) {
            g_argmax_ws.val[blockIdx.x] = best;
            g_argmax_ws.idx[blockIdx.x] = bi;
            __threadfence();
            const unsigned prev = atomicAdd(&g_argmax_ws.counter, 1u);
            is_last = (prev == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (warp == 0) {
        best = -INFINITY;
        bi = 0x7fffffff;
        for (int b = lane; b < (int)gridDim.x; b += 32) {
            const float v = *reinterpret_cast<volatile float *>(&g_argmax_ws.val[b]);
            const int ix = *reinterpret_cast<volatile int *>(&g_argmax_ws.idx[b]);
            combine(best, bi, v, ix);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) combine(best, bi, __shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, bi, o));
        if (lane == 0) {
            *out = bi;
            g_argmax_ws.counter = 0;
        }
    }
}

__global__ void rmsnorm_f16_kernel(const __half *__restrict__ x, const float *__restrict__ gamma, __half *__restrict__ y, int dim, float eps) {
    __shared__ float sred[32];
    const __half *xr = x + (size_t)blockIdx.x * dim;
    __half *yr = y + (size_t)blockIdx.x * dim;
    float ss = 0.f;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) {
        const float v = __half2float(xr[i]);
        ss += v * v;
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += sred[w];
    const float inv = rsqrtf(tot / (float)dim + eps);
    for (int i = threadIdx.x; i < dim; i += blockDim.x) yr[i] = __float2half((__half2float(xr[i]) * inv) * gamma[i]);
}

// LayerNormQ::forward (llm/src/ops/LayerNormQ.cc:12-52): fp32 row -> int8 row, eps 1e-5.  The reference sums the row serially in fp32
// (mean, then squared deviations), and the int8 result depends on those sums to the last bit near rounding ties, so the two reductions
// are done by ONE thread in the reference's order (the loads are independent of the add chain and pipeline); the normalisation of the
// row is spread over the block, with the reference's operation order and no FMA contraction.  One block per row.
__global__ void layernorm_q_kernel(const float *__restrict__ x, const float *__restrict__ weight, const float *__restrict__ bias, int8_t *__restrict__ out,
                                   int dim) {
    extern __shared__ __align__(16) float srow[];  // [dim] + 2
    const float *xr = x + (size_t)blockIdx.x * dim;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) srow[i] = xr[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        // 16-byte loads in front of the dependent add chains: the chain (4 cycles per element) is what remains
        float mean = 0.f;
        int k = 0;
        for (; k + 4 <= dim; k += 4) {
            const float4 v = *reinterpret_cast<const float4 *>(srow + k);
            mean = __fadd_rn(mean, v.x);
            mean = __fadd_rn(mean, v.y);
            mean = __fadd_rn(mean, v.z);
            mean = __fadd_rn(mean, v.w);
        }
        for (; k < dim; k++) mean = __fadd_rn(mean, srow[k]);
        mean = __fdiv_rn(mean, (float)dim);
        float sq = 0.f;
        for (k = 0; k + 4 <= dim; k += 4) {
            const float4 v = *reinterpret_cast<const float4 *>(srow + k);
            const float d0 = __fsub_rn(v.x, mean), d1 = __fsub_rn(v.y, mean), d2 = __fsub_rn(v.z, mean), d3 = __fsub_rn(v.w, mean);
            sq = __fadd_rn(sq, __fmul_rn(d0, d0));
            sq = __fadd_rn(sq, __fmul_rn(d1, d1));
            sq = __fadd_rn(sq, __fmul_rn(d2, d2));
            sq = __fadd_rn(sq, __fmul_rn(d3, d3));
        }
        for (; k < dim; k++) {
            const float d = __fsub_rn(srow[k], mean);
            sq = __fadd_rn(sq, __fmul_rn(d, d));
        }
        const float var = __fdiv_rn(sq, (float)dim);
        srow[dim] = mean;
        srow[dim + 1] = __fsqrt_rn(__fadd_rn(var, 0.00001f));
    }
    __syncthreads();
    const float mean = srow[dim], sd = srow[dim + 1];
    int8_t *orow = out + (size_t)blockIdx.x * dim;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) {
        const float fp = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(srow[i], mean), sd), weight[i]), bias[i]);
        orow[i] = (int8_t)(int)roundf(fp);  // std::round: half away from zero, then the implementation's float -> int8 conversion
    }
}

// fp32 residual add of the OPT decoder layer (Int8OPTDecoderLayer::add, llm/src/nn_modules/Int8OPTDecoderLayer.cc:10-22)
__global__ void add_f32_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = __fadd_rn(a[i], b[i]);
}

// ---- row-wise versions for prompt processing (n tokens at once)
__global__ void embedding_rows_kernel(const __half *__restrict__ table, const int *__restrict__ tokens, float *__restrict__ resid, int E) {
    const __half2 *row = reinterpret_cast<const __half2 *>(table + (size_t)tokens[blockIdx.x] * E);
    float2 *dst = reinterpret_cast<float2 *>(resid + (size_t)blockIdx.x * E);
    for (int i = threadIdx.x; i < E / 2; i += blockDim.x) dst[i] = __half22float2(row[i]);
}

// one 256-thread block per row; the row stays in registers between the two passes (dim <= 256 * 4 * kRmsVec), 16-byte loads, 8-byte stores
constexpr int kRmsVec = 8;  // float4 per thread: rows up to 8192 channels
__global__ void __launch_bounds__(256) rmsnorm_rows_f32_kernel(const float *__restrict__ x, const float *__restrict__ gamma, __half *__restrict__ y, int dim, float eps) {
    __shared__ float sred[8];
    const float4 *xr = reinterpret_cast<const float4 *>(x + (size_t)blockIdx.x * dim);
    const float4 *gr = reinterpret_cast<const float4 *>(gamma);
    uint2 *yr = reinterpret_cast<uint2 *>(y + (size_t)blockIdx.x * dim);
    const int nv = dim >> 2;
    float4 v[kRmsVec];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < kRmsVec; k++) {
        const int i = threadIdx.x + k * 256;
        v[k] = i < nv ? xr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        ss += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) tot += sred[w];
    const float inv = rsqrtf(tot / (float)dim + eps);
#pragma unroll
    for (int k = 0; k < kRmsVec; k++) {
        const int i = threadIdx.x + k * 256;
        if (i < nv) {
            const float4 g = gr[i];
            yr[i] = make_uint2(pack_half2((v[k].x * inv) * g.x, (v[k].y * inv) * g.y), pack_half2((v[k].z * inv) * g.z, (v[k].w * inv) * g.w));
        }
    }
}
// general shapes (dim not a multiple of 4, or longer than the register tile)
__global__ void rmsnorm_rows_f32_generic_kernel(const float *__restrict__ x, const float *__restrict__ gamma, __half *__restrict__ y, int dim, float eps) {
    __shared__ float sred[32];
    const float *xr = x + (size_t)blockIdx.x * dim;
    __half *yr = y + (size_t)blockIdx.x * dim;
    float ss = 0.f;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) ss += xr[i] * xr[i];
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += sred[w];
    const float inv = rsqrtf(tot / (float)dim + eps);
    for (int i = threadIdx.x; i < dim; i += blockDim.x) yr[i] = __float2half((xr[i] * inv) * gamma[i]);
}

// act[r][c] = SiLU(gu[r][c]) * gu[r][F + c]   (SiLuMul_half, cuda/Int4llamaDecoderLayer.cu:12-30; fp32 math); 8 channels per thread
__global__ void silu_mul_rows_kernel(const __half *__restrict__ gu, __half *__restrict__ act, int F, long long total8) {
    const int F8 = F >> 3;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total8; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / F8;
        const int c8 = (int)(e - r * F8);
        const uint4 g4 = *reinterpret_cast<const uint4 *>(gu + r * 2 * F + (size_t)c8 * 8), u4 = *reinterpret_cast<const uint4 *>(gu + r * 2 * F + F + (size_t)c8 * 8);
        const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w}, uw[4] = {u4.x, u4.y, u4.z, u4.w};
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float2 g = __half22float2(*reinterpret_cast<const __half2 *>(&gw[i])), u = __half22float2(*reinterpret_cast<const __half2 *>(&uw[i]));
            o[i] = pack_half2((g.x / (1.f + __expf(-g.x))) * u.x, (g.y / (1.f + __expf(-g.y))) * u.y);
        }
        *reinterpret_cast<uint4 *>(act + e * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}
__global__ void silu_mul_rows_generic_kernel(const __half *__restrict__ gu, __half *__restrict__ act, int F, long long total) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / F;
        const int c = (int)(e % F);
        const float g = __half2float(gu[r * 2 * F + c]), u = __half2float(gu[r * 2 * F + F + c]);
        act[e] = __float2half((g / (1.f + __expf(-g))) * u);
    }
}

cudaError_t launch_cfg(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *attr, dim3 grid, dim3 block, cudaStream_t s, bool pdl) {
    cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaSuccess;
}

}  // namespace

cudaError_t launch_embedding(Ctx *ctx, const __half *table, const int *token, float *resid, int E, bool pdl, int rows, int max_ctx, int *safe) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    launch_cfg(cfg, attr, dim3(4), dim3(256), ctx->stream, pdl);
    return cudaLaunchKernelEx(&cfg, embedding_kernel, table, token, resid, E, rows, max_ctx, safe);
}

cudaError_t launch_argmax(Ctx *ctx, const float *logits, int n, int *out, bool pdl) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    int blocks = (n + 1023) / 1024;
    if (blocks > 128) blocks = 128;
    if (blocks < 1) blocks = 1;
    launch_cfg(cfg, attr, dim3(blocks), dim3(256), ctx->stream, pdl);
    return cudaLaunchKernelEx(&cfg, argmax_kernel, logits, n, out);
}

cudaError_t launch_embedding_rows(Ctx *ctx, const __half *table, const int *tokens, float *resid, int n, int E) {
    embedding_rows_kernel<<<n, 256, 0, ctx->stream>>>(table, tokens, resid, E);
    return cudaGetLastError();
}
cudaError_t launch_rmsnorm_rows_f32(Ctx *ctx, const float *x, const float *gamma, __half *y, int rows, int dim, float eps) {
    if ((dim & 3) == 0 && dim <= 256 * 4 * kRmsVec && !(((uintptr_t)x | (uintptr_t)gamma) & 15) && !((uintptr_t)y & 7))
        rmsnorm_rows_f32_kernel<<<rows, 256, 0, ctx->stream>>>(x, gamma, y, dim, eps);
    else
        rmsnorm_rows_f32_generic_kernel<<<rows, 256, 0, ctx->stream>>>(x, gamma, y, dim, eps);
    return cudaGetLastError();
}
cudaError_t launch_silu_mul_rows(Ctx *ctx, const __half *gu, __half *act, int rows, int F) {
    const long long total = (long long)rows * F;
    if ((F & 7) == 0 && !(((uintptr_t)gu | (uintptr_t)act) & 15)) {
        const long long total8 = total >> 3, nb = (total8 + 255) / 256;
        silu_mul_rows_kernel<<<(unsigned)(nb < 8192 ? nb : 8192), 256, 0, ctx->stream>>>(gu, act, F, total8);
    } else {
        const long long nb = (total + 255) / 256;
        silu_mul_rows_generic_kernel<<<(unsigned)(nb < 4096 ? nb : 4096), 256, 0, ctx->stream>>>(gu, act, F, total);
    }
    return cudaGetLastError();
}

cudaError_t launch_layernorm_q(Ctx *ctx, const float *x, const float *weight, const float *bias, int8_t *out, int rows, int dim) {
    const size_t smem = (size_t)(dim + 2) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(layernorm_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    layernorm_q_kernel<<<rows, 128, smem, ctx->stream>>>(x, weight, bias, out, dim);
    return cudaGetLastError();
}
cudaError_t launch_add_f32(Ctx *ctx, const float *a, const float *b, float *out, long long n) {
    const long long nb = (n + 255) / 256;
    add_f32_kernel<<<(unsigned)(nb < 2368 ? nb : 2368), 256, 0, ctx->stream>>>(a, b, out, n);
    return cudaGetLastError();
}

cudaError_t launch_rmsnorm_f16(Ctx *ctx, const __half *x, const float *gamma, __half *y, int rows, int dim, float eps) {
    rmsnorm_f16_kernel<<<rows, 256, 0, ctx->stream>>>(x, gamma, y, dim, eps);
    return cudaGetLastError();
}

}  // namespace tce
