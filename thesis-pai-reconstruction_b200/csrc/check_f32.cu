// fp32 check path (north star: "generator outputs within 1e-5 with the fp32 accumulate check path").
//
// Plain CUDA-core kernels on fp32 NCHW tensors in the reference's own layouts -- no bf16 rounding anywhere, one
// thread per output element, fp32 FMA accumulation in a fixed order.  They are the slow, exact twin of the
// tcgen05 path: tests run the same drop-in modules through them (pai_b200.engine.check_path()) to show that the
// index math / layer wiring reproduces the reference (models/pix2pix.py:46-111,198-216, models/wrapper.py:196-238)
// to fp32 rounding, which separates "bf16 operand noise" from "wrong arithmetic" in the 1e-2 bf16 parity bound.
// Forward and backward (data / weight / bias gradients, BatchNorm and activation backward); never used by a training
// or benchmark path.
#include "pai_common.cuh"
#include "pai_kernels.h"

namespace pai {
namespace {

__device__ __forceinline__ float check_act(float v, int act, float slope) {
    switch (act) {
        case PAI_ACT_LEAKY: return v > 0.f ? v : v * slope;
        case PAI_ACT_RELU: return fmaxf(v, 0.f);
        case PAI_ACT_TANH: return tanhf(v);
        default: return v;
    }
}

// y[n,co,oy,ox] = bias[co] + sum_{ci,ky,kx} pre(x[n,ci,oy*s-p+ky,ox*s-p+kx]) * w[co,ci,ky,kx]   (nn.Conv2d)
__global__ void check_conv2d_kernel(const float* __restrict__ x, int n, int cin, int h, int w,
                                    const float* __restrict__ wt, int cout, int k, int stride, int pad,
                                    const float* __restrict__ bias, int pre_act, float slope, int ho, int wo,
                                    float* __restrict__ y) {
    const long long total = (long long)n * cout * ho * wo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % wo);
        const int oy = (int)((i / wo) % ho);
        const int co = (int)((i / ((long long)wo * ho)) % cout);
        const int b = (int)(i / ((long long)wo * ho * cout));
        float acc = bias != nullptr ? bias[co] : 0.f;
        for (int ci = 0; ci < cin; ++ci) {
            const float* xp = x + ((size_t)b * cin + ci) * h * w;
            const float* wp = wt + ((size_t)co * cin + ci) * k * k;
            for (int ky = 0; ky < k; ++ky) {
                const int iy = oy * stride - pad + ky;
                if (iy < 0 || iy >= h) continue;
                for (int kx = 0; kx < k; ++kx) {
                    const int ix = ox * stride - pad + kx;
                    if (ix < 0 || ix >= w) continue;
                    acc = fmaf(check_act(xp[(size_t)iy * w + ix], pre_act, slope), wp[ky * k + kx], acc);
                }
            }
        }
        y[i] = acc;
    }
}

// y[n,co,oy,ox] = bias[co] + sum_{ci,ky,kx : oy = iy*s - p + ky} pre(x[n,ci,iy,ix]) * w[ci,co,ky,kx]  (nn.ConvTranspose2d)
__global__ void check_convT2d_kernel(const float* __restrict__ x, int n, int cin, int h, int w,
                                     const float* __restrict__ wt, int cout, int k, int stride, int pad,
                                     const float* __restrict__ bias, int pre_act, float slope, int ho, int wo,
                                     float* __restrict__ y) {
    const long long total = (long long)n * cout * ho * wo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % wo);
        const int oy = (int)((i / wo) % ho);
        const int co = (int)((i / ((long long)wo * ho)) % cout);
        const int b = (int)(i / ((long long)wo * ho * cout));
        float acc = bias != nullptr ? bias[co] : 0.f;
        for (int ci = 0; ci < cin; ++ci) {
            const float* xp = x + ((size_t)b * cin + ci) * h * w;
            const float* wp = wt + ((size_t)ci * cout + co) * k * k;
            for (int ky = 0; ky < k; ++ky) {
                const int ty = oy + pad - ky;
                if (ty < 0 || ty % stride) continue;
                const int iy = ty / stride;
                if (iy >= h) continue;
                for (int kx = 0; kx < k; ++kx) {
                    const int tx = ox + pad - kx;
                    if (tx < 0 || tx % stride) continue;
                    const int ix = tx / stride;
                    if (ix >= w) continue;
                    acc = fmaf(check_act(xp[(size_t)iy * w + ix], pre_act, slope), wp[ky * k + kx], acc);
                }
            }
        }
        y[i] = acc;
    }
}

// nn.BatchNorm2d on fp32 NCHW, one CTA per channel.  training: biased batch variance for the normalisation, unbiased
// for running_var, running = (1 - momentum) * running + momentum * batch (models/pix2pix.py:70,106 defaults).
__global__ void check_bn_kernel(const float* __restrict__ x, int n, int c, int hw, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* __restrict__ running_mean,
                                float* __restrict__ running_var, int training, float eps, float momentum,
                                float* __restrict__ y) {
    __shared__ double red[256];
    __shared__ float s_mean, s_inv;
    const int ch = blockIdx.x, tid = threadIdx.x;
    const long long cnt = (long long)n * hw;
    if (training) {
        double s = 0.0;
        for (long long i = tid; i < cnt; i += blockDim.x) s += x[((size_t)(i / hw) * c + ch) * hw + (i % hw)];
        red[tid] = s;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (tid < o) red[tid] += red[tid + o];
            __syncthreads();
        }
        const double mean = red[0] / (double)cnt;
        __syncthreads();
        double q = 0.0;
        for (long long i = tid; i < cnt; i += blockDim.x) {
            const double d = (double)x[((size_t)(i / hw) * c + ch) * hw + (i % hw)] - mean;
            q += d * d;
        }
        red[tid] = q;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (tid < o) red[tid] += red[tid + o];
            __syncthreads();
        }
        if (tid == 0) {
            const double var = red[0] / (double)cnt;
            s_mean = (float)mean;
            s_inv = 1.f / sqrtf((float)var + eps);
            if (running_mean != nullptr) {
                const double unbiased = cnt > 1 ? red[0] / (double)(cnt - 1) : var;
                running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * (float)mean;
                running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * (float)unbiased;
            }
        }
    } else if (tid == 0) {
        s_mean = running_mean[ch];
        s_inv = 1.f / sqrtf(running_var[ch] + eps);
    }
    __syncthreads();
    const float mean = s_mean, inv = s_inv, g = gamma[ch], b = beta[ch];
    for (long long i = tid; i < cnt; i += blockDim.x) {
        const size_t o = ((size_t)(i / hw) * c + ch) * hw + (i % hw);
        y[o] = (x[o] - mean) * inv * g + b;
    }
}

__global__ void check_act_kernel(const float* __restrict__ x, long long count, int act, float slope, float* __restrict__ y) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        y[i] = check_act(x[i], act, slope);
}

// dW of nn.Conv2d (transposed = 0, dw [cout,cin,k,k]) or nn.ConvTranspose2d (transposed = 1, dw [cin,cout,k,k]):
// one CTA per (out-channel, in-channel) pair, k*k <= 16 taps per thread, positions strided over the CTA.
//   conv : dW[co,ci,ky,kx] = sum_{b,oy,ox} g[b,co,oy,ox] * pre(x[b,ci,oy*s-p+ky,ox*s-p+kx])
//   convT: dW[ci,co,ky,kx] = sum_{b,iy,ix} pre(x[b,ci,iy,ix]) * g[b,co,iy*s-p+ky,ix*s-p+kx]
__global__ void check_conv_wgrad_kernel(const float* __restrict__ x, int n, int cin, int h, int w,
                                        const float* __restrict__ g, int cout, int ho, int wo, int k, int stride, int pad,
                                        int pre_act, float slope, int transposed, float* __restrict__ dw) {
    __shared__ double red[64];
    const int co = blockIdx.x % cout, ci = blockIdx.x / cout;
    // "small" grid = the tensor indexed directly by the loop, "big" = the one read at s*pos - p + k
    const float* small = transposed ? x : g;
    const float* big = transposed ? g : x;
    const int sc = transposed ? ci : co, sC = transposed ? cin : cout, sh = transposed ? h : ho, sw = transposed ? w : wo;
    const int bc = transposed ? co : ci, bC = transposed ? cout : cin, bh = transposed ? ho : h, bw = transposed ? wo : w;
    float acc[16];
    for (int t = 0; t < 16; ++t) acc[t] = 0.f;
    const long long cnt = (long long)n * sh * sw;
    for (long long i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int px = (int)(i % sw), py = (int)((i / sw) % sh), b = (int)(i / ((long long)sw * sh));
        float sv = small[(((size_t)b * sC + sc) * sh + py) * sw + px];
        if (transposed) sv = check_act(sv, pre_act, slope);
        const float* bp = big + ((size_t)b * bC + bc) * bh * bw;
        for (int ky = 0; ky < k; ++ky) {
            const int yy = py * stride - pad + ky;
            if (yy < 0 || yy >= bh) continue;
            for (int kx = 0; kx < k; ++kx) {
                const int xx = px * stride - pad + kx;
                if (xx < 0 || xx >= bw) continue;
                float bv = bp[(size_t)yy * bw + xx];
                if (!transposed) bv = check_act(bv, pre_act, slope);
                acc[ky * k + kx] = fmaf(sv, bv, acc[ky * k + kx]);
            }
        }
    }
    float* out = dw + (transposed ? ((size_t)ci * cout + co) : ((size_t)co * cin + ci)) * k * k;
    for (int t = 0; t < k * k; ++t) {
        red[threadIdx.x] = (double)acc[t];
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[t] = (float)red[0];
        __syncthreads();
    }
}

// sums[ch] = sum over (n, hw) of g[n, ch, :]   (bias gradients), one CTA per channel
__global__ void check_chansum_kernel(const float* __restrict__ g, int n, int c, int hw, float* __restrict__ sums) {
    __shared__ double red[256];
    const int ch = blockIdx.x, tid = threadIdx.x;
    const long long cnt = (long long)n * hw;
    double s = 0.0;
    for (long long i = tid; i < cnt; i += blockDim.x) s += g[((size_t)(i / hw) * c + ch) * hw + (i % hw)];
    red[tid] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (tid == 0) sums[ch] = (float)red[0];
}

// train-mode nn.BatchNorm2d backward, one CTA per channel:
// dx = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat)),  dgamma = sum g * xhat,  dbeta = sum g
__global__ void check_bn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, int n, int c, int hw,
                                    const float* __restrict__ gamma, float eps, float* __restrict__ dx,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
    __shared__ double red[256];
    __shared__ double s_val[4];
    const int ch = blockIdx.x, tid = threadIdx.x;
    const long long cnt = (long long)n * hw;
    auto at = [&](long long i) { return ((size_t)(i / hw) * c + ch) * hw + (i % hw); };
    auto reduce = [&](double v, int slot) {
        red[tid] = v;
        __syncthreads();
        for (int o = blockDim.x / 2; o > 0; o >>= 1) {
            if (tid < o) red[tid] += red[tid + o];
            __syncthreads();
        }
        if (tid == 0) s_val[slot] = red[0];
        __syncthreads();
    };
    double s = 0.0;
    for (long long i = tid; i < cnt; i += blockDim.x) s += x[at(i)];
    reduce(s, 0);
    const double mean = s_val[0] / (double)cnt;
    double q = 0.0;
    for (long long i = tid; i < cnt; i += blockDim.x) {
        const double d = (double)x[at(i)] - mean;
        q += d * d;
    }
    reduce(q, 1);
    const float inv = 1.f / sqrtf((float)(s_val[1] / (double)cnt) + eps);
    double sg = 0.0, sgx = 0.0;
    for (long long i = tid; i < cnt; i += blockDim.x) {
        const size_t o = at(i);
        const double xh = ((double)x[o] - mean) * (double)inv;
        sg += g[o];
        sgx += (double)g[o] * xh;
    }
    reduce(sg, 2);
    reduce(sgx, 3);
    const double mg = s_val[2] / (double)cnt, mgx = s_val[3] / (double)cnt;
    const float gm = gamma[ch];
    for (long long i = tid; i < cnt; i += blockDim.x) {
        const size_t o = at(i);
        const double xh = ((double)x[o] - mean) * (double)inv;
        dx[o] = (float)((double)gm * (double)inv * ((double)g[o] - mg - xh * mgx));
    }
    if (tid == 0) {
        dgamma[ch] = (float)s_val[3];
        dbeta[ch] = (float)s_val[2];
    }
}

// dx = g * act'(x) with x the value the activation was applied to
__global__ void check_act_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, long long count, int act,
                                     float slope, float* __restrict__ dx) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i];
        float d = 1.f;
        if (act == PAI_ACT_LEAKY) d = v > 0.f ? 1.f : slope;
        else if (act == PAI_ACT_RELU) d = v > 0.f ? 1.f : 0.f;
        else if (act == PAI_ACT_TANH) { const float t = tanhf(v); d = 1.f - t * t; }
        dx[i] = g[i] * d;
    }
}

int grid_for(long long total) {
    long long g = (total + 255) / 256;
    return (int)(g < 148 * 32 ? (g > 0 ? g : 1) : 148 * 32);
}

}  // namespace
}  // namespace pai

using namespace pai;

extern "C" {

int pai_check_conv2d_f32(const float* x, int n, int cin, int h, int w, const float* wt, int cout, int k, int stride,
                         int pad, const float* bias, int pre_act, float slope, int transposed, float* y, void* stream) {
    PAI_REQUIRE(x && wt && y, "pai_check_conv2d_f32: null pointer");
    PAI_REQUIRE(n >= 0 && cin > 0 && cout > 0 && h > 0 && w > 0 && k > 0 && stride > 0 && pad >= 0,
                "pai_check_conv2d_f32: bad shape n=%d cin=%d cout=%d %dx%d k=%d stride=%d pad=%d", n, cin, cout, h, w, k,
                stride, pad);
    const int ho = transposed ? (h - 1) * stride - 2 * pad + k : (h + 2 * pad - k) / stride + 1;
    const int wo = transposed ? (w - 1) * stride - 2 * pad + k : (w + 2 * pad - k) / stride + 1;
    PAI_REQUIRE(ho > 0 && wo > 0, "pai_check_conv2d_f32: empty output (%dx%d)", ho, wo);
    if (n == 0) return 0;
    const long long total = (long long)n * cout * ho * wo;
    cudaStream_t st = (cudaStream_t)stream;
    if (transposed)
        check_convT2d_kernel<<<grid_for(total), 256, 0, st>>>(x, n, cin, h, w, wt, cout, k, stride, pad, bias, pre_act,
                                                             slope, ho, wo, y);
    else
        check_conv2d_kernel<<<grid_for(total), 256, 0, st>>>(x, n, cin, h, w, wt, cout, k, stride, pad, bias, pre_act,
                                                            slope, ho, wo, y);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

int pai_check_batchnorm_f32(const float* x, int n, int c, int hw, const float* gamma, const float* beta,
                            float* running_mean, float* running_var, int training, float eps, float momentum, float* y,
                            void* stream) {
    PAI_REQUIRE(x && gamma && beta && y, "pai_check_batchnorm_f32: null pointer");
    PAI_REQUIRE(training || (running_mean && running_var), "pai_check_batchnorm_f32: eval mode needs running statistics");
    PAI_REQUIRE(n >= 0 && c > 0 && hw > 0, "pai_check_batchnorm_f32: bad shape");
    if (n == 0) return 0;
    check_bn_kernel<<<c, 256, 0, (cudaStream_t)stream>>>(x, n, c, hw, gamma, beta, running_mean, running_var, training,
                                                        eps, momentum, y);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

int pai_check_conv2d_wgrad_f32(const float* x, int n, int cin, int h, int w, const float* g, int cout, int k, int stride,
                               int pad, int pre_act, float slope, int transposed, float* dw, float* dbias, void* stream) {
    PAI_REQUIRE(x && g && dw, "pai_check_conv2d_wgrad_f32: null pointer");
    PAI_REQUIRE(n > 0 && cin > 0 && cout > 0 && k > 0 && k <= 4 && stride > 0 && pad >= 0, "pai_check_conv2d_wgrad_f32: bad shape");
    const int ho = transposed ? (h - 1) * stride - 2 * pad + k : (h + 2 * pad - k) / stride + 1;
    const int wo = transposed ? (w - 1) * stride - 2 * pad + k : (w + 2 * pad - k) / stride + 1;
    cudaStream_t st = (cudaStream_t)stream;
    check_conv_wgrad_kernel<<<cin * cout, 64, 0, st>>>(x, n, cin, h, w, g, cout, ho, wo, k, stride, pad, pre_act, slope,
                                                       transposed, dw);
    PAI_CUDA_OK(cudaGetLastError());
    if (dbias != nullptr) {
        check_chansum_kernel<<<cout, 256, 0, st>>>(g, n, cout, ho * wo, dbias);
        PAI_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int pai_check_batchnorm_bwd_f32(const float* x, const float* g, int n, int c, int hw, const float* gamma, float eps,
                                float* dx, float* dgamma, float* dbeta, void* stream) {
    PAI_REQUIRE(x && g && gamma && dx && dgamma && dbeta && n > 0 && c > 0 && hw > 0, "pai_check_batchnorm_bwd_f32: bad arguments");
    check_bn_bwd_kernel<<<c, 256, 0, (cudaStream_t)stream>>>(x, g, n, c, hw, gamma, eps, dx, dgamma, dbeta);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

int pai_check_act_bwd_f32(const float* x, const float* g, long long count, int act, float slope, float* dx, void* stream) {
    PAI_REQUIRE(x && g && dx && count >= 0, "pai_check_act_bwd_f32: bad arguments");
    if (count == 0) return 0;
    check_act_bwd_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(x, g, count, act, slope, dx);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

int pai_check_act_f32(const float* x, long long count, int act, float slope, float* y, void* stream) {
    PAI_REQUIRE(x && y && count >= 0, "pai_check_act_f32: bad arguments");
    if (count == 0) return 0;
    check_act_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(x, count, act, slope, y);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
