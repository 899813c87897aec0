// Fused Adam update + bf16 GEMM-operand repack for the 4x4 convolution weights, and a multi-tensor Adam
// for the small parameters (biases, BatchNorm affine).
//
// Reference: torch.optim.Adam(lr=2e-4, betas=(0.5, 0.999), eps=1e-7) built by
// UnetWrapper.configure_optimizers (models/wrapper.py:97-115) and stepped in training_step
// (models/wrapper.py:136,160).  The arithmetic follows torch's single-tensor Adam:
//   m += (g - m) * (1 - b1);  v = v * b2 + (1 - b2) * g * g;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// The implicit-GEMM kernels read bf16 K-major packs of every weight [A, B, 4, 4] (fp32 master, reference
// layout):  P1[a, (ky*4+kx)*B + b]              (Conv2d fprop / ConvTranspose2d dgrad operand)
//           P2[py*2+px, b, (ty*2+tx)*A + a]     (ConvTranspose2d fprop / Conv2d dgrad operand, 4 sub-pixel
//                                                phases, tap k = T[parity][t] of SURVEY.md Appendix B)
// One pass over the weight does the update and rewrites both packs, instead of an optimizer pass plus two
// strided copies.
#include "pai_common.cuh"
#include "pai_kernels.h"

namespace pai {

struct AdamHyper {
    float one_minus_b1, b2, one_minus_b2, step_size, inv_bc2_sqrt, eps;
};

// Bias-correction scalars kept on the device so that a captured CUDA graph of the training step stays valid from one
// replay to the next: ++*step; dyn = {lr / (1 - b1^t), 1 / sqrt(1 - b2^t)} in double like torch's host code.
__global__ void adam_prepare_kernel(int* step, float lr, float beta1, float beta2, float* dyn) {
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int t = *step + 1;
        *step = t;
        dyn[0] = (float)((double)lr / (1.0 - pow((double)beta1, (double)t)));
        dyn[1] = (float)(1.0 / sqrt(1.0 - pow((double)beta2, (double)t)));
    }
}

__device__ __forceinline__ AdamHyper adam_dyn(AdamHyper h, const float* __restrict__ dyn) {
    if (dyn != nullptr) {
        h.step_size = __ldg(dyn);
        h.inv_bc2_sqrt = __ldg(dyn + 1);
    }
    return h;
}

__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, const AdamHyper& h) {
    m = fmaf(g - m, h.one_minus_b1, m);
    v = fmaf(v, h.b2, h.one_minus_b2 * g * g);
    const float denom = fmaf(sqrtf(v), h.inv_bc2_sqrt, h.eps);
    return p - h.step_size * (m / denom);
}

static constexpr int kTA = 32, kTB = 32, kPackThreads = 256;

// one 32 x 32 (a, b) tile of one weight; smem: s1[16][32 a][32 b] and s2[16][32 b][32 a] bf16 (2 x 32 KB)
__device__ __forceinline__ void adam_pack_tile(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                                               float* __restrict__ v, int A, int B, const AdamHyper& hy,
                                               __nv_bfloat16* __restrict__ p1, __nv_bfloat16* __restrict__ p2, int b_pad,
                                               int a0, int b0) {
    extern __shared__ __nv_bfloat16 pk_smem[];
    __nv_bfloat16* s1 = pk_smem;                       // [tap][a][b]
    __nv_bfloat16* s2 = pk_smem + 16 * kTA * kTB;      // [tap][b][a]
    const int tid = threadIdx.x;
    // ---- pass 1: float4 = 4 taps of one (a, b); a row of the tile is 32 b x 16 taps = 128 float4
    // (four iterations = 16 independent 16-byte loads in flight per thread)
#pragma unroll 4
    for (int i = tid; i < kTA * 128; i += kPackThreads) {
        const int al = i >> 7, r = i & 127, bl = r >> 2, t4 = (r & 3) * 4;
        const int a = a0 + al, b = b0 + bl;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a < A && b < B) {
            const size_t off = ((size_t)a * B + b) * 16 + t4;
            x = *reinterpret_cast<const float4*>(w + off);
            if (g != nullptr) {
                const float4 gg = *reinterpret_cast<const float4*>(g + off);
                float4 mm = *reinterpret_cast<const float4*>(m + off);
                float4 vv = *reinterpret_cast<const float4*>(v + off);
                x.x = adam_update(x.x, gg.x, mm.x, vv.x, hy);
                x.y = adam_update(x.y, gg.y, mm.y, vv.y, hy);
                x.z = adam_update(x.z, gg.z, mm.z, vv.z, hy);
                x.w = adam_update(x.w, gg.w, mm.w, vv.w, hy);
                *reinterpret_cast<float4*>(w + off) = x;
                *reinterpret_cast<float4*>(m + off) = mm;
                *reinterpret_cast<float4*>(v + off) = vv;
            }
        }
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat16 hb = __float2bfloat16_rn(xs[j]);
            s1[((t4 + j) * kTA + al) * kTB + bl] = hb;
            s2[((t4 + j) * kTB + bl) * kTA + al] = hb;
        }
    }
    __syncthreads();
    // ---- pass 2: 16-byte chunks (8 bf16).  P1 rows: (tap, a) -> 32 consecutive b
    if (p1 != nullptr) {
        for (int i = tid; i < 16 * kTA * 4; i += kPackThreads) {
            const int c = i & 3, al = (i >> 2) & 31, tap = i >> 7;
            const int a = a0 + al, b = b0 + c * 8;
            if (a < A && b < B) {
                __nv_bfloat16* dst = p1 + (size_t)a * (16 * B) + (size_t)tap * B + b;
                const __nv_bfloat16* src = s1 + (tap * kTA + al) * kTB + c * 8;
                if (b + 8 <= B && (B & 7) == 0) {
                    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
                } else {
                    for (int j = 0; j < 8 && b + j < B; ++j) dst[j] = src[j];
                }
            }
        }
    }
    // P2 rows: (phase, t, b) -> 32 consecutive a;  k = 0,1,2,3 -> (parity, t) = (1,0),(0,0),(1,1),(0,1)
    if (p2 != nullptr) {
        for (int i = tid; i < 16 * kTB * 4; i += kPackThreads) {
            const int c = i & 3, bl = (i >> 2) & 31, tap = i >> 7;
            const int ky = tap >> 2, kx = tap & 3;
            const int py = (ky & 1) ^ 1, ty = ky >> 1, px = (kx & 1) ^ 1, tx = kx >> 1;
            const int a = a0 + c * 8, b = b0 + bl;
            if (a < A && b < B) {
                __nv_bfloat16* dst = p2 + ((size_t)(py * 2 + px) * b_pad + b) * (size_t)(4 * A) + (size_t)(ty * 2 + tx) * A + a;
                const __nv_bfloat16* src = s2 + (tap * kTB + bl) * kTA + c * 8;
                if (a + 8 <= A && (A & 7) == 0) {
                    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
                } else {
                    for (int j = 0; j < 8 && a + j < A; ++j) dst[j] = src[j];
                }
            }
        }
    }
}

// grid (ceil(B/32), ceil(A/32))
__global__ void __launch_bounds__(kPackThreads)
adam_pack_conv4x4_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                         float* __restrict__ v, int A, int B, AdamHyper hy0, const float* __restrict__ dyn,
                         __nv_bfloat16* __restrict__ p1, __nv_bfloat16* __restrict__ p2, int b_pad) {
    adam_pack_tile(w, g, m, v, A, B, adam_dyn(hy0, dyn), p1, p2, b_pad, blockIdx.y * kTA, blockIdx.x * kTB);
}

// All 4x4 convolution weights of an optimizer in ONE launch: block -> (weight, tile) through the prefix of tile counts.
// One launch per weight left the small layers (8 .. 128 tiles) at the ~20 us latency of a single block each, 8 of the 17
// launches of a step.
static constexpr int kMaxPackWeights = 24;
struct AdamPackTable {
    float* w[kMaxPackWeights];
    const float* g[kMaxPackWeights];
    float* m[kMaxPackWeights];
    float* v[kMaxPackWeights];
    __nv_bfloat16* p1[kMaxPackWeights];
    __nv_bfloat16* p2[kMaxPackWeights];
    int A[kMaxPackWeights], B[kMaxPackWeights], b_pad[kMaxPackWeights];
    int first_tile[kMaxPackWeights + 1];
    int count;
};

__global__ void __launch_bounds__(kPackThreads)
adam_pack_multi_kernel(const __grid_constant__ AdamPackTable t, AdamHyper hy0, const float* __restrict__ dyn) {
    pdl_launch_dependents();
    pdl_wait();
    int k = 0;
    while (k + 1 < t.count && (int)blockIdx.x >= t.first_tile[k + 1]) ++k;
    const int tile = blockIdx.x - t.first_tile[k];
    const int tiles_b = (t.B[k] + kTB - 1) / kTB;
    adam_pack_tile(t.w[k], t.g[k], t.m[k], t.v[k], t.A[k], t.B[k], adam_dyn(hy0, dyn), t.p1[k], t.p2[k], t.b_pad[k],
                   (tile / tiles_b) * kTA, (tile % tiles_b) * kTB);
}

// ---- weight gradient: GEMM layout -> parameter layout ----------------------------------------------------------
// The stream-K wgrad kernels accumulate into a zeroed tap-major fp32 buffer dw[16][ab] (ab = A*B weight pairs).  This
// pass writes the gradient in the reference's parameter layout grad[ab][16] (= [A, B, 4, 4]): reads coalesced per tap,
// one 64-byte store per (a, b).  zero_src re-zeroes dw for a persistent accumulation buffer; measured slower overall
// than a fresh fill right before the wgrad kernel, because the fill leaves the lines in L2 for the red.adds.
__global__ void __launch_bounds__(256)
wgrad_finish_kernel(float* __restrict__ dw, long long ab, float* __restrict__ grad, int zero_src) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < ab; i += (long long)gridDim.x * blockDim.x) {
        float v[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            v[t] = dw[t * ab + i];
            if (zero_src) dw[t * ab + i] = 0.f;
        }
        float4* o = reinterpret_cast<float4*>(grad + i * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
}

// ---- multi-tensor Adam for the small parameters --------------------------------------------------
static constexpr int kMaxTensors = 48;
struct AdamTable {
    float* p[kMaxTensors];
    const float* g[kMaxTensors];
    float* m[kMaxTensors];
    float* v[kMaxTensors];
    int n[kMaxTensors];
    int count;
};

__global__ void __launch_bounds__(256)
adam_multi_kernel(const __grid_constant__ AdamTable t, AdamHyper hy0, const float* __restrict__ dyn) {
    pdl_launch_dependents();
    pdl_wait();
    const AdamHyper hy = adam_dyn(hy0, dyn);
    for (int k = blockIdx.y; k < t.count; k += gridDim.y) {
        float* p = t.p[k];
        const float* g = t.g[k];
        float* m = t.m[k];
        float* v = t.v[k];
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < t.n[k]; i += gridDim.x * blockDim.x) {
            float mm = m[i], vv = v[i];
            p[i] = adam_update(p[i], g[i], mm, vv, hy);
            m[i] = mm;
            v[i] = vv;
        }
    }
}

// ---- exponential moving average of the weights (callbacks/ema.py:24-34 -> torch_ema's update rule) ------------------
// shadow <- shadow - (1 - d) * (shadow - param) for `count` tensors in one launch; the table reuses AdamTable
// (p = shadow, g = live parameter).
__global__ void __launch_bounds__(256)
ema_multi_kernel(const __grid_constant__ AdamTable t, float one_minus_decay) {
    for (int k = blockIdx.y; k < t.count; k += gridDim.y) {
        float* sh = t.p[k];
        const float* live = t.g[k];
        const int n = t.n[k];
        const int n4 = ((reinterpret_cast<uintptr_t>(sh) | reinterpret_cast<uintptr_t>(live)) & 15) == 0 ? n >> 2 : 0;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
            float4 s = reinterpret_cast<float4*>(sh)[i];
            const float4 p = __ldg(reinterpret_cast<const float4*>(live) + i);
            s.x -= one_minus_decay * (s.x - p.x), s.y -= one_minus_decay * (s.y - p.y);
            s.z -= one_minus_decay * (s.z - p.z), s.w -= one_minus_decay * (s.w - p.w);
            reinterpret_cast<float4*>(sh)[i] = s;
        }
        for (int i = 4 * n4 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
            sh[i] -= one_minus_decay * (sh[i] - live[i]);
    }
}

}  // namespace pai

using namespace pai;

extern "C" {

int pai_ema_multi(int count, float* const* shadows, const float* const* params, const int* numels, float one_minus_decay,
                  void* stream) {
    PAI_REQUIRE(count >= 0 && (count == 0 || (shadows && params && numels)), "pai_ema_multi: bad arguments");
    for (int base = 0; base < count; base += kMaxTensors) {
        AdamTable t;
        t.count = count - base < kMaxTensors ? count - base : kMaxTensors;
        int big = 0;
        for (int i = 0; i < t.count; ++i) {
            t.p[i] = shadows[base + i], t.g[i] = params[base + i], t.m[i] = nullptr, t.v[i] = nullptr, t.n[i] = numels[base + i];
            if (t.n[i] > big) big = t.n[i];
        }
        int bx = (big / 4 + 255) / 256;
        if (bx > 148 * 4) bx = 148 * 4;
        if (bx < 1) bx = 1;
        ema_multi_kernel<<<dim3(bx, t.count), 256, 0, (cudaStream_t)stream>>>(t, one_minus_decay);
        PAI_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int pai_adam_pack_conv4x4(float* w, const float* grad, float* exp_avg, float* exp_avg_sq, int a, int b, float beta1,
                          float beta2, float step_size, float inv_bias_correction2_sqrt, float eps, void* pack1,
                          void* pack2, int b_pad, const float* dyn, void* stream) {
    PAI_REQUIRE(w != nullptr && a > 0 && b > 0, "pai_adam_pack_conv4x4: null weight / empty shape");
    PAI_REQUIRE(grad == nullptr || (exp_avg != nullptr && exp_avg_sq != nullptr),
                "pai_adam_pack_conv4x4: optimizer state missing");
    PAI_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(exp_avg) & 15) == 0 && (reinterpret_cast<uintptr_t>(exp_avg_sq) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(pack1) & 15) == 0 && (reinterpret_cast<uintptr_t>(pack2) & 15) == 0,
                "pai_adam_pack_conv4x4: pointers must be 16 B aligned");
    PAI_REQUIRE(pack2 == nullptr || b_pad >= b, "pai_adam_pack_conv4x4: b_pad %d < b %d", b_pad, b);
    static DeviceOnce attr_once;
    const int attr_dev = current_device();
    if (attr_dev < 0) return -1;
    const int smem = 2 * 16 * kTA * kTB * (int)sizeof(__nv_bfloat16);
    if (attr_once.need(attr_dev)) {
        PAI_CUDA_OK(cudaFuncSetAttribute(adam_pack_conv4x4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_once.mark(attr_dev);
    }
    AdamHyper hy = {1.f - beta1, beta2, 1.f - beta2, step_size, inv_bias_correction2_sqrt, eps};
    dim3 grid((b + kTB - 1) / kTB, (a + kTA - 1) / kTA);
    adam_pack_conv4x4_kernel<<<grid, kPackThreads, smem, (cudaStream_t)stream>>>(
        w, grad, exp_avg, exp_avg_sq, a, b, hy, dyn, (__nv_bfloat16*)pack1, (__nv_bfloat16*)pack2, b_pad);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

int pai_adam_pack_conv4x4_multi(int count, float* const* ws, const float* const* grads, float* const* exp_avgs,
                                float* const* exp_avg_sqs, const int* as, const int* bs, void* const* pack1s,
                                void* const* pack2s, const int* b_pads, float beta1, float beta2, float step_size,
                                float inv_bias_correction2_sqrt, float eps, const float* dyn, void* stream) {
    PAI_REQUIRE(count >= 0 && (count == 0 || (ws && grads && exp_avgs && exp_avg_sqs && as && bs && pack1s && pack2s && b_pads)),
                "pai_adam_pack_conv4x4_multi: null table");
    static DeviceOnce attr_once;
    const int attr_dev = current_device();
    if (attr_dev < 0) return -1;
    const int smem = 2 * 16 * kTA * kTB * (int)sizeof(__nv_bfloat16);
    if (attr_once.need(attr_dev)) {
        PAI_CUDA_OK(cudaFuncSetAttribute(adam_pack_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_once.mark(attr_dev);
    }
    AdamHyper hy = {1.f - beta1, beta2, 1.f - beta2, step_size, inv_bias_correction2_sqrt, eps};
    for (int base = 0; base < count; base += kMaxPackWeights) {
        AdamPackTable t;
        t.count = count - base < kMaxPackWeights ? count - base : kMaxPackWeights;
        int tiles = 0;
        for (int i = 0; i < t.count; ++i) {
            const int k = base + i;
            PAI_REQUIRE(ws[k] && grads[k] && exp_avgs[k] && exp_avg_sqs[k] && as[k] > 0 && bs[k] > 0,
                        "pai_adam_pack_conv4x4_multi: bad entry %d", k);
            PAI_REQUIRE(((reinterpret_cast<uintptr_t>(ws[k]) | reinterpret_cast<uintptr_t>(grads[k]) |
                          reinterpret_cast<uintptr_t>(exp_avgs[k]) | reinterpret_cast<uintptr_t>(exp_avg_sqs[k]) |
                          reinterpret_cast<uintptr_t>(pack1s[k]) | reinterpret_cast<uintptr_t>(pack2s[k])) & 15) == 0,
                        "pai_adam_pack_conv4x4_multi: entry %d: pointers must be 16 B aligned", k);
            PAI_REQUIRE(pack2s[k] == nullptr || b_pads[k] >= bs[k], "pai_adam_pack_conv4x4_multi: entry %d: b_pad %d < b %d",
                        k, b_pads[k], bs[k]);
            t.w[i] = ws[k], t.g[i] = grads[k], t.m[i] = exp_avgs[k], t.v[i] = exp_avg_sqs[k];
            t.p1[i] = (__nv_bfloat16*)pack1s[k], t.p2[i] = (__nv_bfloat16*)pack2s[k];
            t.A[i] = as[k], t.B[i] = bs[k], t.b_pad[i] = b_pads[k];
            t.first_tile[i] = tiles;
            tiles += ((as[k] + kTA - 1) / kTA) * ((bs[k] + kTB - 1) / kTB);
        }
        t.first_tile[t.count] = tiles;
        PAI_CUDA_OK(launch_pdl(adam_pack_multi_kernel, dim3((unsigned)tiles), dim3(kPackThreads), (size_t)smem, (cudaStream_t)stream, 1, t, hy, dyn));
    }
    return 0;
}

int pai_wgrad_finish(float* dw_tap_major, long long ab, float* grad, int zero_src, void* stream) {
    PAI_REQUIRE(dw_tap_major && grad && ab > 0, "pai_wgrad_finish: null pointer / empty shape");
    PAI_REQUIRE((reinterpret_cast<uintptr_t>(grad) & 15) == 0, "pai_wgrad_finish: grad must be 16 B aligned");
    long long blocks = (ab + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    PAI_CUDA_OK(launch_pdl(wgrad_finish_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, 1, dw_tap_major, ab, grad, zero_src));
    return 0;
}

int pai_adam_prepare(int* step, float lr, float beta1, float beta2, float* dyn, void* stream) {
    PAI_REQUIRE(step != nullptr && dyn != nullptr, "pai_adam_prepare: null pointer");
    PAI_CUDA_OK(launch_pdl(adam_prepare_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, 1, step, lr, beta1, beta2, dyn));
    return 0;
}

int pai_adam_multi(int count, float* const* params, const float* const* grads, float* const* exp_avgs,
                   float* const* exp_avg_sqs, const int* numels, float beta1, float beta2, float step_size,
                   float inv_bias_correction2_sqrt, float eps, const float* dyn, void* stream) {
    PAI_REQUIRE(count >= 0 && (count == 0 || (params && grads && exp_avgs && exp_avg_sqs && numels)),
                "pai_adam_multi: null table");
    AdamHyper hy = {1.f - beta1, beta2, 1.f - beta2, step_size, inv_bias_correction2_sqrt, eps};
    for (int base = 0; base < count; base += kMaxTensors) {
        AdamTable t;
        t.count = count - base < kMaxTensors ? count - base : kMaxTensors;
        int max_n = 1;
        for (int i = 0; i < t.count; ++i) {
            t.p[i] = params[base + i], t.g[i] = grads[base + i], t.m[i] = exp_avgs[base + i], t.v[i] = exp_avg_sqs[base + i];
            t.n[i] = numels[base + i];
            PAI_REQUIRE(t.p[i] && t.g[i] && t.m[i] && t.v[i] && t.n[i] >= 0, "pai_adam_multi: bad entry %d", base + i);
            if (t.n[i] > max_n) max_n = t.n[i];
        }
        int bx = (max_n + 255) / 256;
        if (bx > 148 * 8) bx = 148 * 8;
        PAI_CUDA_OK(launch_pdl(adam_multi_kernel, dim3(bx, t.count), dim3(256), 0, (cudaStream_t)stream, 1, t, hy, dyn));
    }
    return 0;
}

}  // extern "C"
