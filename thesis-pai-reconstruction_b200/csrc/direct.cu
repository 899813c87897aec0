// Direct (CUDA-core) kernels for the four degenerate layers of the hot path, which are HBM-bound and
// have a 1- or 2-wide channel dimension that cannot feed a 64-channel TMA/UMMA K-block:
//   enc0  Conv2d(1, 64, 4, 2, 1)            models/pix2pix.py:141-147
//   D0    Conv2d(2, 64, 4, 2, 1)+LeakyReLU  models/wrapper.py:196-206,229 (input = cat([x, y], 1))
//   dec7  ConvTranspose2d(128, 1, 4, 2, 1)  models/pix2pix.py:186-192   (data / weight gradient)
//   D4    Conv2d(512, 1, 4, 1, 1, no bias)  models/wrapper.py:233       (data / weight gradient)
//
//   pai_smallc_conv_fprop:  out[n,oy,ox,c] = act(bias[c] + sum_{t,j} plane_j[n, s*oy+dy_t, s*ox+dx_t] * W[c][t][j])
//   pai_smallc_conv_wgrad:  dW[c][t][j]   += sum_{n,y,x} A[n,y,x,c] * plane_j[n, s*y+dy_t, s*x+dx_t]
// with (dy_t, dx_t) = (ky-1, kx-1), or (1-ky, 1-kx) when `flip` (gradient of a stride-1 conv).
// Planes are single-channel fp32 images (the network input, target, prediction, or a 1-channel
// gradient); the wide operand is NHWC bf16.
#include "pai_common.cuh"
#include "pai_kernels.h"

namespace pai {

static constexpr int kDcThreads = 256;

struct SmallConvGeom {
    int n, ih, iw;   // plane (fine) grid
    int oh, ow;      // wide-tensor (coarse) grid
    int stride, flip, cin, c;
};

__device__ __forceinline__ void tap_offset(int t, int flip, int* dy, int* dx) {
    const int ky = t >> 2, kx = t & 3;
    *dy = flip ? 1 - ky : ky - 1;
    *dx = flip ? 1 - kx : kx - 1;
}

__global__ void __launch_bounds__(kDcThreads)
smallc_fprop_kernel(const float* __restrict__ p0, const float* __restrict__ p1, SmallConvGeom g,
                    const float* __restrict__ w /* [c][16][cin] */, const float* __restrict__ bias,
                    __nv_bfloat16* __restrict__ o1, int ld1, int act1, __nv_bfloat16* __restrict__ o2, int ld2,
                    int act2, float slope) {
    extern __shared__ float ws[];  // [16][cin][c]
    for (int i = threadIdx.x; i < g.c * 16 * g.cin; i += kDcThreads) {
        const int ch = i / (16 * g.cin), r = i - ch * 16 * g.cin;
        ws[r * g.c + ch] = w[i];
    }
    __syncthreads();
    const int cv = g.c >> 3;
    const long long total = (long long)g.n * g.oh * g.ow * cv;
    for (long long idx = (long long)blockIdx.x * kDcThreads + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * kDcThreads) {
        const int vec = (int)(idx % cv);
        long long pix = idx / cv;
        const int ox = (int)(pix % g.ow);
        const int oy = (int)((pix / g.ow) % g.oh);
        const int n = (int)(pix / ((long long)g.ow * g.oh));
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = bias != nullptr ? bias[vec * 8 + i] : 0.f;
        const size_t plane_off = (size_t)n * g.ih * g.iw;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            int dy, dx;
            tap_offset(t, g.flip, &dy, &dx);
            const int iy = oy * g.stride + dy, ix = ox * g.stride + dx;
            if (iy < 0 || iy >= g.ih || ix < 0 || ix >= g.iw) continue;
            const size_t o = plane_off + (size_t)iy * g.iw + ix;
            const float v0 = __ldg(p0 + o);
            const float* wr = ws + (t * g.cin) * g.c + vec * 8;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(v0, wr[i], acc[i]);
            if (g.cin == 2) {
                const float v1 = __ldg(p1 + o);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(v1, wr[g.c + i], acc[i]);
            }
        }
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float a = acc[2 * i], b = acc[2 * i + 1];
            if (act1 == PAI_ACT_LEAKY) a = a > 0.f ? a : a * slope, b = b > 0.f ? b : b * slope;
            if (act1 == PAI_ACT_RELU) a = fmaxf(a, 0.f), b = fmaxf(b, 0.f);
            h[i] = __floats2bfloat162_rn(a, b);
        }
        *reinterpret_cast<uint4*>(o1 + pix * ld1 + vec * 8) = u;
        if (o2 != nullptr) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float a = acc[2 * i], b = acc[2 * i + 1];
                if (act2 == PAI_ACT_LEAKY) a = a > 0.f ? a : a * slope, b = b > 0.f ? b : b * slope;
                if (act2 == PAI_ACT_RELU) a = fmaxf(a, 0.f), b = fmaxf(b, 0.f);
                h[i] = __floats2bfloat162_rn(a, b);
            }
            *reinterpret_cast<uint4*>(o2 + pix * ld2 + vec * 8) = u;
        }
    }
}

// thread <-> (8 channels, 4 taps); a block covers kDcThreads / (4 * c/8) pixels per iteration
__global__ void __launch_bounds__(kDcThreads)
smallc_wgrad_kernel(const __nv_bfloat16* __restrict__ a, int lda, const float* __restrict__ p0,
                    const float* __restrict__ p1, SmallConvGeom g, float* __restrict__ dw /* [c][16][cin] */) {
    extern __shared__ float red[];  // [kDcThreads][8*4*cin]
    const int cv = g.c >> 3;
    const int tpp = cv * 4;                  // threads per pixel
    const int sub = threadIdx.x % tpp;
    const int vec = sub % cv, tg = sub / cv;  // channel vector, tap group (taps 4*tg .. 4*tg+3)
    const int ppb = kDcThreads / tpp;
    float acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[i][t][0] = acc[i][t][1] = 0.f;
    const long long total = (long long)g.n * g.oh * g.ow;
    for (long long pix = (long long)blockIdx.x * ppb + threadIdx.x / tpp; pix < total;
         pix += (long long)gridDim.x * ppb) {
        const int ox = (int)(pix % g.ow);
        const int oy = (int)((pix / g.ow) % g.oh);
        const int n = (int)(pix / ((long long)g.ow * g.oh));
        const uint4 u = *reinterpret_cast<const uint4*>(a + pix * lda + vec * 8);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
        float av[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h[i]);
            av[2 * i] = f.x, av[2 * i + 1] = f.y;
        }
        const size_t plane_off = (size_t)n * g.ih * g.iw;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            int dy, dx;
            tap_offset(tg * 4 + t, g.flip, &dy, &dx);
            const int iy = oy * g.stride + dy, ix = ox * g.stride + dx;
            if (iy < 0 || iy >= g.ih || ix < 0 || ix >= g.iw) continue;
            const size_t o = plane_off + (size_t)iy * g.iw + ix;
            const float v0 = __ldg(p0 + o);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i][t][0] = fmaf(av[i], v0, acc[i][t][0]);
            if (g.cin == 2) {
                const float v1 = __ldg(p1 + o);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i][t][1] = fmaf(av[i], v1, acc[i][t][1]);
            }
        }
    }
    const int per = 32 * g.cin;  // values per thread
    float* mine = red + (size_t)threadIdx.x * per;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            mine[(i * 4 + t) * g.cin] = acc[i][t][0];
            if (g.cin == 2) mine[(i * 4 + t) * g.cin + 1] = acc[i][t][1];
        }
    __syncthreads();
    // outputs of this block: tpp * per values, each summed over the ppb pixel replicas
    for (int o = threadIdx.x; o < tpp * per; o += kDcThreads) {
        const int s = o / per, e = o - s * per;  // sub-thread, element
        float sum = 0.f;
        for (int r = 0; r < ppb; ++r) sum += red[(size_t)(r * tpp + s) * per + e];
        const int svec = s % cv, stg = s / cv;
        const int i = e / (4 * g.cin), t = (e / g.cin) % 4, j = e % g.cin;
        atomicAdd(dw + ((size_t)(svec * 8 + i) * 16 + stg * 4 + t) * g.cin + j, sum);
    }
}

}  // namespace pai

using namespace pai;

extern "C" {

int pai_smallc_conv_fprop(const float* plane0, const float* plane1, int cin, int n, int ih, int iw, int oh, int ow,
                          int stride, int flip, const float* w, const float* bias, int c, void* out1, int ld1,
                          int act1, void* out2, int ld2, int act2, float slope, void* stream) {
    PAI_REQUIRE(plane0 && w && out1 && (cin == 1 || (cin == 2 && plane1)), "pai_smallc_conv_fprop: bad planes / cin=%d", cin);
    PAI_REQUIRE(c > 0 && c % 8 == 0 && ld1 % 8 == 0 && (out2 == nullptr || ld2 % 8 == 0) && (stride == 1 || stride == 2),
                "pai_smallc_conv_fprop: c=%d must be a multiple of 8, stride 1|2", c);
    const size_t smem = sizeof(float) * 16 * cin * c;
    PAI_REQUIRE(smem <= 96 * 1024, "pai_smallc_conv_fprop: weights do not fit shared memory (c=%d)", c);
    static bool attr = false;
    if (!attr) {
        PAI_CUDA_OK(cudaFuncSetAttribute(smallc_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        attr = true;
    }
    SmallConvGeom g{n, ih, iw, oh, ow, stride, flip, cin, c};
    const long long total = (long long)n * oh * ow * (c / 8);
    long long blocks = (total + kDcThreads - 1) / kDcThreads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    smallc_fprop_kernel<<<(int)blocks, kDcThreads, smem, (cudaStream_t)stream>>>(
        plane0, plane1, g, w, bias, (__nv_bfloat16*)out1, ld1, act1, (__nv_bfloat16*)out2, ld2, act2, slope);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

int pai_smallc_conv_wgrad(const void* a, int lda, int c, const float* plane0, const float* plane1, int cin, int n,
                          int ih, int iw, int oh, int ow, int stride, int flip, float* dw, void* stream) {
    PAI_REQUIRE(a && plane0 && dw && (cin == 1 || (cin == 2 && plane1)), "pai_smallc_conv_wgrad: bad planes / cin=%d", cin);
    const int cv = c / 8;
    PAI_REQUIRE(c > 0 && c % 8 == 0 && lda % 8 == 0 && cv * 4 <= kDcThreads && kDcThreads % (cv * 4) == 0,
                "pai_smallc_conv_wgrad: c=%d must be a multiple of 8 with (c/2) | 256", c);
    const size_t smem = sizeof(float) * kDcThreads * 32 * cin;
    static bool attr = false;
    if (!attr) {
        PAI_CUDA_OK(cudaFuncSetAttribute(smallc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr = true;
    }
    SmallConvGeom g{n, ih, iw, oh, ow, stride, flip, cin, c};
    const int ppb = kDcThreads / (cv * 4);
    long long blocks = ((long long)n * oh * ow + ppb - 1) / ppb;
    if (blocks > 148 * 2) blocks = 148 * 2;
    cudaStream_t st = (cudaStream_t)stream;
    smallc_wgrad_kernel<<<(int)blocks, kDcThreads, smem, st>>>((const __nv_bfloat16*)a, lda, plane0, plane1, g, dw);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
