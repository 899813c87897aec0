// fp32 check path (north star: "generator outputs within 1e-5 with the fp32 accumulate check path").
//
// Plain CUDA-core kernels on fp32 NCHW tensors in the reference's own layouts -- no bf16 rounding anywhere, one
// thread per output element, fp32 FMA accumulation in a fixed order.  They are the slow, exact twin of the
// tcgen05 path: tests run the same drop-in modules through them (pai_b200.engine.check_path()) to show that the
// index math / layer wiring reproduces the reference (models/pix2pix.py:46-111,198-216, models/wrapper.py:196-238)
// to fp32 rounding, which separates "bf16 operand noise" from "wrong arithmetic" in the 1e-2 bf16 parity bound.
// Forward only; not used by any training or benchmark path.
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

int pai_check_act_f32(const float* x, long long count, int act, float slope, float* y, void* stream) {
    PAI_REQUIRE(x && y && count >= 0, "pai_check_act_f32: bad arguments");
    if (count == 0) return 0;
    check_act_kernel<<<grid_for(count), 256, 0, (cudaStream_t)stream>>>(x, count, act, slope, y);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
