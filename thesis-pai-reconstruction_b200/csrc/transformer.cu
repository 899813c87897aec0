// Kernels of the Trans U-Net's ViT bottleneck (models/trans_unet.py:120-175): nn.LayerNorm,
// nn.TransformerEncoderLayer(d, 8 heads, d_ff=2048, activation="gelu", post-norm) stacked 12 times.
// The Linear layers run on the tensor-core pointwise GEMM (igemm.cu); here are the row-wise and attention parts.
//
// Tokens are rows of a [m, d] bf16 matrix.  The reference builds the encoder layer WITHOUT batch_first and feeds
// it [n, patches, d], so self-attention runs over the BATCH axis: sequence length S = n (images of the batch on
// this GPU), "batch" B = patches.  Row index = s * B + b.
#include "pai_common.cuh"
#include "pai_kernels.h"

namespace pai {

__device__ __forceinline__ float block_sum(float v, float* red /* [32] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}

// ---- LayerNorm over the last dimension: one CTA per row --------------------------------------
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, int d, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ y,
                     float* __restrict__ mean, float* __restrict__ rstd) {
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const __nv_bfloat16* xr = x + row * d;
    float s = 0.f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) s += __bfloat162float(xr[i]);
    const float mu = block_sum(s, red) / d;
    float q = 0.f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        const float c = __bfloat162float(xr[i]) - mu;
        q = fmaf(c, c, q);
    }
    const float rs = rsqrtf(block_sum(q, red) / d + eps);
    for (int i = threadIdx.x; i < d; i += blockDim.x)
        y[row * d + i] = __float2bfloat16_rn((__bfloat162float(xr[i]) - mu) * rs * gamma[i] + beta[i]);
    if (threadIdx.x == 0) {
        mean[row] = mu;
        rstd[row] = rs;
    }
}
// dx = rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat))
__global__ void __launch_bounds__(256)
layernorm_bwd_dx_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ g, int d,
                        const float* __restrict__ gamma, const float* __restrict__ mean,
                        const float* __restrict__ rstd, __nv_bfloat16* __restrict__ dx) {
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const float mu = mean[row], rs = rstd[row];
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        const float gg = __bfloat162float(g[row * d + i]) * gamma[i];
        const float xh = (__bfloat162float(x[row * d + i]) - mu) * rs;
        a += gg;
        b = fmaf(gg, xh, b);
    }
    const float ma = block_sum(a, red) / d;
    const float mb = block_sum(b, red) / d;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        const float gg = __bfloat162float(g[row * d + i]) * gamma[i];
        const float xh = (__bfloat162float(x[row * d + i]) - mu) * rs;
        dx[row * d + i] = __float2bfloat16_rn(rs * (gg - ma - xh * mb));
    }
}
// dgamma[i] = sum_rows g*xhat, dbeta[i] = sum_rows g: thread <-> column, rows split over blockIdx.y
__global__ void __launch_bounds__(256)
layernorm_bwd_affine_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ g, long long m, int d,
                            const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dgamma,
                            float* __restrict__ dbeta) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= d) return;
    float a = 0.f, b = 0.f;
    for (long long r = blockIdx.y; r < m; r += gridDim.y) {
        const float gg = __bfloat162float(g[r * d + col]);
        a = fmaf(gg, (__bfloat162float(x[r * d + col]) - mean[r]) * rstd[r], a);
        b += gg;
    }
    atomicAdd(dgamma + col, a);
    atomicAdd(dbeta + col, b);
}

// ---- GELU (exact, erf) -------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gelu_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long n, __nv_bfloat16* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = __bfloat162float(x[i]);
        y[i] = __float2bfloat16_rn(0.5f * v * (1.f + erff(v * 0.70710678118654752f)));
    }
}
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ g, long long n,
                __nv_bfloat16* __restrict__ dx) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = __bfloat162float(x[i]);
        const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752f));
        const float pdf = 0.3989422804014327f * __expf(-0.5f * v * v);
        dx[i] = __float2bfloat16_rn(__bfloat162float(g[i]) * (cdf + v * pdf));
    }
}

// ---- multi-head self-attention over the sequence axis S -----------------------------------------
// qkv: [S*B, 3*E] bf16 (row = s*B + b; columns [0,E) = Q, [E,2E) = K, [2E,3E) = V), E = H*hd.
// One CTA per (query i, b*H + h).  probs: [B*H, S, S] fp32 (saved for backward).  S <= 1024.
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, int S, int B, int H, int hd, float scale,
                float* __restrict__ probs, __nv_bfloat16* __restrict__ out /* [S*B, E] */) {
    extern __shared__ float sm[];        // q[hd] | p[S]
    float* qs = sm;
    float* ps = sm + hd;
    __shared__ float red[32];
    const int i = blockIdx.x, bh = blockIdx.y, b = bh / H, h = bh % H;
    const int E = H * hd;
    const long long ld = 3LL * E;
    const __nv_bfloat16* qrow = qkv + ((long long)i * B + b) * ld + h * hd;
    for (int t = threadIdx.x; t < hd; t += blockDim.x) qs[t] = __bfloat162float(qrow[t]);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = warp; j < S; j += nw) {
        const __nv_bfloat16* krow = qkv + ((long long)j * B + b) * ld + E + h * hd;
        float acc = 0.f;
        for (int t = lane; t < hd; t += 32) acc = fmaf(qs[t], __bfloat162float(krow[t]), acc);
        acc = warp_sum(acc);
        if (lane == 0) ps[j] = acc * scale;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < S; j += blockDim.x) mx = fmaxf(mx, ps[j]);
    {   // block max
        float v = mx;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) red[warp] = v;
        __syncthreads();
        v = red[0];
        for (int w = 1; w < nw; ++w) v = fmaxf(v, red[w]);
        mx = v;
    }
    float sum = 0.f;
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
        const float e = __expf(ps[j] - mx);
        ps[j] = e;
        sum += e;
    }
    const float inv = 1.f / block_sum(sum, red);
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
        ps[j] *= inv;
        probs[((long long)bh * S + i) * S + j] = ps[j];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < hd; t += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < S; ++j)
            acc = fmaf(ps[j], __bfloat162float(qkv[((long long)j * B + b) * ld + 2 * E + h * hd + t]), acc);
        out[((long long)i * B + b) * E + h * hd + t] = __float2bfloat16_rn(acc);
    }
}
// backward, query side: dS[i, :] = P[i, :] * (dP[i, :] - sum_j P dP), dP[i, j] = dO_i . V_j;  dQ_i = scale * sum_j dS_ij K_j
__global__ void __launch_bounds__(128)
attn_bwd_q_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout, int S, int B, int H,
                  int hd, float scale, const float* __restrict__ probs, float* __restrict__ ds /* [B*H, S, S] */,
                  __nv_bfloat16* __restrict__ dqkv /* [S*B, 3E] */) {
    extern __shared__ float sm[];        // do[hd] | dsrow[S]
    float* dos = sm;
    float* dsr = sm + hd;
    __shared__ float red[32];
    const int i = blockIdx.x, bh = blockIdx.y, b = bh / H, h = bh % H;
    const int E = H * hd;
    const long long ld = 3LL * E;
    for (int t = threadIdx.x; t < hd; t += blockDim.x)
        dos[t] = __bfloat162float(dout[((long long)i * B + b) * E + h * hd + t]);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const float* prow = probs + ((long long)bh * S + i) * S;
    for (int j = warp; j < S; j += nw) {
        const __nv_bfloat16* vrow = qkv + ((long long)j * B + b) * ld + 2 * E + h * hd;
        float acc = 0.f;
        for (int t = lane; t < hd; t += 32) acc = fmaf(dos[t], __bfloat162float(vrow[t]), acc);
        acc = warp_sum(acc);
        if (lane == 0) dsr[j] = acc;
    }
    __syncthreads();
    float part = 0.f;
    for (int j = threadIdx.x; j < S; j += blockDim.x) part = fmaf(prow[j], dsr[j], part);
    const float delta = block_sum(part, red);
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
        const float v = prow[j] * (dsr[j] - delta);
        dsr[j] = v;
        ds[((long long)bh * S + i) * S + j] = v;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < hd; t += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < S; ++j)
            acc = fmaf(dsr[j], __bfloat162float(qkv[((long long)j * B + b) * ld + E + h * hd + t]), acc);
        dqkv[((long long)i * B + b) * ld + h * hd + t] = __float2bfloat16_rn(acc * scale);
    }
}
// backward, key/value side: dK_j = scale * sum_i dS_ij Q_i;  dV_j = sum_i P_ij dO_i
__global__ void __launch_bounds__(128)
attn_bwd_kv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout, int S, int B, int H,
                   int hd, float scale, const float* __restrict__ probs, const float* __restrict__ ds,
                   __nv_bfloat16* __restrict__ dqkv) {
    extern __shared__ float sm[];        // pcol[S] | dscol[S]
    float* pc = sm;
    float* dc = sm + S;
    const int j = blockIdx.x, bh = blockIdx.y, b = bh / H, h = bh % H;
    const int E = H * hd;
    const long long ld = 3LL * E;
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        pc[i] = probs[((long long)bh * S + i) * S + j];
        dc[i] = ds[((long long)bh * S + i) * S + j];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < hd; t += blockDim.x) {
        float dk = 0.f, dv = 0.f;
        for (int i = 0; i < S; ++i) {
            dk = fmaf(dc[i], __bfloat162float(qkv[((long long)i * B + b) * ld + h * hd + t]), dk);
            dv = fmaf(pc[i], __bfloat162float(dout[((long long)i * B + b) * E + h * hd + t]), dv);
        }
        dqkv[((long long)j * B + b) * ld + E + h * hd + t] = __float2bfloat16_rn(dk * scale);
        dqkv[((long long)j * B + b) * ld + 2 * E + h * hd + t] = __float2bfloat16_rn(dv);
    }
}

}  // namespace pai

using namespace pai;
typedef __nv_bfloat16 bf16;

extern "C" {

int pai_layernorm_fwd(const void* x, long long m, int d, const float* gamma, const float* beta, float eps, void* y,
                      float* mean, float* rstd, void* stream) {
    PAI_REQUIRE(x && gamma && beta && y && mean && rstd && m > 0 && d > 0, "pai_layernorm_fwd: null pointer / empty input");
    layernorm_fwd_kernel<<<(unsigned)m, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, d, gamma, beta, eps, (bf16*)y, mean, rstd);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}
int pai_layernorm_bwd(const void* x, const void* g, long long m, int d, const float* gamma, const float* mean,
                      const float* rstd, void* dx, float* dgamma, float* dbeta, void* stream) {
    PAI_REQUIRE(x && g && gamma && mean && rstd && dx && dgamma && dbeta && m > 0, "pai_layernorm_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    layernorm_bwd_dx_kernel<<<(unsigned)m, 256, 0, st>>>((const bf16*)x, (const bf16*)g, d, gamma, mean, rstd, (bf16*)dx);
    PAI_CUDA_OK(cudaGetLastError());
    PAI_CUDA_OK(cudaMemsetAsync(dgamma, 0, sizeof(float) * d, st));
    PAI_CUDA_OK(cudaMemsetAsync(dbeta, 0, sizeof(float) * d, st));
    int ry = (int)(m < 32 ? m : 32);
    layernorm_bwd_affine_kernel<<<dim3((d + 255) / 256, ry), 256, 0, st>>>((const bf16*)x, (const bf16*)g, m, d, mean, rstd,
                                                                          dgamma, dbeta);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}
int pai_gelu_fwd(const void* x, long long n, void* y, void* stream) {
    PAI_REQUIRE(x && y && n > 0, "pai_gelu_fwd: null pointer");
    long long b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    gelu_fwd_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, n, (bf16*)y);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}
int pai_gelu_bwd(const void* x, const void* g, long long n, void* dx, void* stream) {
    PAI_REQUIRE(x && g && dx && n > 0, "pai_gelu_bwd: null pointer");
    long long b = (n + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    gelu_bwd_kernel<<<(int)b, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)g, n, (bf16*)dx);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}
int pai_attn_fwd(const void* qkv, int s, int b, int heads, int head_dim, float* probs, void* out, void* stream) {
    PAI_REQUIRE(qkv && probs && out && s > 0 && b > 0 && heads > 0 && head_dim > 0, "pai_attn_fwd: null pointer / empty");
    PAI_REQUIRE((head_dim + s) * 4 <= 48 * 1024, "pai_attn_fwd: head_dim + sequence too large for shared memory");
    attn_fwd_kernel<<<dim3(s, b * heads), 128, (head_dim + s) * sizeof(float), (cudaStream_t)stream>>>(
        (const bf16*)qkv, s, b, heads, head_dim, 1.f / sqrtf((float)head_dim), probs, (bf16*)out);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}
int pai_attn_bwd(const void* qkv, const void* dout, int s, int b, int heads, int head_dim, const float* probs,
                 float* ds_work, void* dqkv, void* stream) {
    PAI_REQUIRE(qkv && dout && probs && ds_work && dqkv && s > 0, "pai_attn_bwd: null pointer / empty");
    PAI_REQUIRE((head_dim + s) * 4 <= 48 * 1024 && 2 * s * 4 <= 48 * 1024, "pai_attn_bwd: sizes too large for shared memory");
    cudaStream_t st = (cudaStream_t)stream;
    const float scale = 1.f / sqrtf((float)head_dim);
    attn_bwd_q_kernel<<<dim3(s, b * heads), 128, (head_dim + s) * sizeof(float), st>>>(
        (const bf16*)qkv, (const bf16*)dout, s, b, heads, head_dim, scale, probs, ds_work, (bf16*)dqkv);
    PAI_CUDA_OK(cudaGetLastError());
    attn_bwd_kv_kernel<<<dim3(s, b * heads), 128, 2 * s * sizeof(float), st>>>(
        (const bf16*)qkv, (const bf16*)dout, s, b, heads, head_dim, scale, probs, ds_work, (bf16*)dqkv);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
