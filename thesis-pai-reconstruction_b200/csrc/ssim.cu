// SSIM / PSNR / MSE metric + loss kernels (memory-bound path of the north star).
//
// Replaces torchmetrics==0.11.4 `structural_similarity_index_measure`, `peak_signal_noise_ratio`,
// `mean_squared_error` as bound by the reference in models/utils.py:38-47 (data_range=1.0) and used by
// models/wrapper.py:53-63,150-156,166-173 and report.py:78-96,146,188-217 (algorithm: SURVEY.md App. A).
//
// One pass over an image pair yields everything report.py needs: the per-image SSIM (mean of the map
// over rows/cols 5..dim-6), the 16 depth-band SSIMs (rows 16d+5..16d+10), the squared error (PSNR, MSE,
// RMSE) and, optionally, the full reflect-padded SSIM map.
//
// Tile = 22 x 64 map pixels from a 32 x 74 input patch (reflect indexing at the image border):
//   stage A  global -> smem float4 (p, t, p*p + t*t, p*t), coalesced, optional de-normalisation
//   stage B  horizontal 11-tap Gaussian, lane <-> input row, warp <-> run of 8 columns, packed FFMA2
//   stage C  vertical 11-tap Gaussian, lane <-> column, SSIM formula, warp-shuffle reductions
// Only sigma_p^2 + sigma_t^2 enters SSIM, so 4 planes are filtered instead of torchmetrics' 5.
#include <math.h>
#include <stdlib.h>

#include "pai_common.cuh"
#include "pai_kernels.h"

namespace pai {

static constexpr int TH = 22, TW = 64;          // output tile
static constexpr int IH = TH + 10, IW = TW + 10;  // input patch 32 x 74
static constexpr int IN_PITCH = 75;               // float4 units; odd -> conflict-free lane<->row access
static constexpr int H_PITCH = 33;                // float4 units per column of the h-pass result
static constexpr int RG = 6;                      // output rows per thread in stage C (4 groups: 6,6,6,4)
static constexpr int kSsimThreads = 256;
static constexpr size_t kSsimSmem = (size_t)(IH * IN_PITCH + TW * H_PITCH) * sizeof(float4);

__constant__ float c_gauss[11];

struct Gauss {
    float g[11];
};
static Gauss host_gauss() {
    // torchmetrics _gaussian(kernel_size=11, sigma=1.5): exp(-((i-5)/sigma)^2 / 2), normalised
    Gauss k;
    double s = 0, v[11];
    for (int i = 0; i < 11; ++i) {
        double d = (i - 5) / 1.5;
        v[i] = exp(-0.5 * d * d);
        s += v[i];
    }
    for (int i = 0; i < 11; ++i) k.g[i] = (float)(v[i] / s);
    return k;
}
static int upload_gauss() {
    static DeviceOnce done_once;
    const int done_dev = current_device();
    if (done_dev < 0) return -1;
    if (done_once.need(done_dev)) {
        Gauss k = host_gauss();
        PAI_CUDA_OK(cudaMemcpyToSymbol(c_gauss, k.g, sizeof(k.g)));
        done_once.mark(done_dev);
    }
    return 0;
}

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)),
          "l"(reinterpret_cast<unsigned long long&>(c)));
    return reinterpret_cast<float2&>(d);
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

template <typename T>
__device__ __forceinline__ float load_val(const T* p, size_t i);
template <>
__device__ __forceinline__ float load_val<float>(const float* p, size_t i) {
    return __ldg(p + i);
}
template <>
__device__ __forceinline__ float load_val<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) {
    return __bfloat162float(p[i]);
}
__device__ __forceinline__ float denorm(float x) { return fminf(fmaxf(fmaf(x, 0.5f, 0.5f), 0.f), 1.f); }

struct SsimTerms {
    float a1, a2, b1, b2;
};
__device__ __forceinline__ SsimTerms ssim_terms(float4 m) {
    const float c1 = 1e-4f, c2 = 9e-4f;
    const float mpp = m.x * m.x, mtt = m.y * m.y, mpt = m.x * m.y;
    SsimTerms s;
    s.a1 = 2.f * mpt + c1;
    s.a2 = 2.f * (m.w - mpt) + c2;
    s.b1 = mpp + mtt + c1;
    s.b2 = (m.z - mpp - mtt) + c2;
    return s;
}

// ---- stages B and C, shared by the three kernels ------------------------------------------------
// in: [IH][IN_PITCH] float4 planes;  hbuf: [TW][H_PITCH] float4
__device__ __forceinline__ void hpass(const float4* __restrict__ in, float4* __restrict__ hbuf) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4* row = in + lane * IN_PITCH + warp * 8;
    float2 xa[18], xb[18];
#pragma unroll
    for (int j = 0; j < 18; ++j) {
        float4 v = row[j];
        xa[j] = make_float2(v.x, v.y);
        xb[j] = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float2 g = make_float2(c_gauss[k], c_gauss[k]);
            sa = fma2(xa[o + k], g, sa);
            sb = fma2(xb[o + k], g, sb);
        }
        hbuf[(warp * 8 + o) * H_PITCH + lane] = make_float4(sa.x, sa.y, sb.x, sb.y);
    }
}
// thread <-> (column, row group); returns the filtered planes of up to RG rows
__device__ __forceinline__ void vpass(const float4* __restrict__ hbuf, int col, int rg, float4 (&out)[RG]) {
    const float4* c = hbuf + col * H_PITCH + rg * RG;
    float2 xa[RG + 10], xb[RG + 10];
#pragma unroll
    for (int j = 0; j < RG + 10; ++j) {
        float4 v = (rg * RG + j < IH) ? c[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        xa[j] = make_float2(v.x, v.y);
        xb[j] = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int o = 0; o < RG; ++o) {
        float2 sa = make_float2(0.f, 0.f), sb = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float2 g = make_float2(c_gauss[k], c_gauss[k]);
            sa = fma2(xa[o + k], g, sa);
            sb = fma2(xb[o + k], g, sb);
        }
        out[o] = make_float4(sa.x, sa.y, sb.x, sb.y);
    }
}

template <typename T, bool DENORM>
__device__ __forceinline__ void load_pair_patch(const T* __restrict__ pred, const T* __restrict__ target,
                                                size_t img_off, int h, int w, int y0, int x0, float4* in) {
    for (int i = threadIdx.x; i < IH * IW; i += kSsimThreads) {
        const int r = i / IW, c = i - r * IW;
        const int gy = reflect_idx(y0 - 5 + r, h), gx = reflect_idx(x0 - 5 + c, w);
        const size_t o = img_off + (size_t)gy * w + gx;
        float p = load_val<T>(pred, o), t = load_val<T>(target, o);
        if (DENORM) {
            p = denorm(p);
            t = denorm(t);
        }
        in[r * IN_PITCH + c] = make_float4(p, t, fmaf(p, p, t * t), p * t);
    }
}

// =============================================================================================
// forward: ssim_sum[n], band_sum[n][16] (optional), sse[n], full map (optional)
template <typename T, bool DENORM>
__global__ void __launch_bounds__(kSsimThreads)
ssim_fwd_kernel(const T* __restrict__ pred, const T* __restrict__ target, int h, int w, int band_rows,
                float* __restrict__ ssim_sum, float* __restrict__ band_sum, float* __restrict__ sse,
                float* __restrict__ full_map) {
    extern __shared__ float4 smem4[];
    float4* in = smem4;
    float4* hbuf = smem4 + IH * IN_PITCH;
    __shared__ float acc[18];
    const int img = blockIdx.z, y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
    const size_t img_off = (size_t)img * h * w;
    if (threadIdx.x < 18) acc[threadIdx.x] = 0.f;
    load_pair_patch<T, DENORM>(pred, target, img_off, h, w, y0, x0, in);
    __syncthreads();
    hpass(in, hbuf);
    __syncthreads();

    const int col = threadIdx.x & 63, rg = threadIdx.x >> 6;
    float4 m[RG];
    vpass(hbuf, col, rg, m);
    const int gx = x0 + col;
    const bool col_in = gx < w, col_int = gx >= 5 && gx < w - 5;
    float s_sum = 0.f, e_sum = 0.f, b_sum[2] = {0.f, 0.f};
    const int gy_first = y0 + rg * RG;
    const int band0 = band_sum != nullptr ? gy_first / band_rows : 0;
#pragma unroll
    for (int o = 0; o < RG; ++o) {
        const int lr = rg * RG + o, gy = y0 + lr;
        if (lr < TH && gy < h && col_in) {
            const SsimTerms t = ssim_terms(m[o]);
            const float s = __fdividef(t.a1 * t.a2, t.b1 * t.b2);
            if (full_map != nullptr) full_map[img_off + (size_t)gy * w + gx] = s;
            const float4 ctr = in[(lr + 5) * IN_PITCH + col + 5];
            const float d = ctr.x - ctr.y;
            e_sum = fmaf(d, d, e_sum);
            if (col_int && gy >= 5 && gy < h - 5) s_sum += s;
            if (band_sum != nullptr && col_int) {
                const int b = gy / band_rows, br = gy - b * band_rows;
                if (br >= 5 && br < band_rows - 5) b_sum[b - band0] += s;
            }
        }
    }
    s_sum = warp_sum(s_sum);
    e_sum = warp_sum(e_sum);
    if (band_sum != nullptr) {
        b_sum[0] = warp_sum(b_sum[0]);
        b_sum[1] = warp_sum(b_sum[1]);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&acc[0], s_sum);
        atomicAdd(&acc[1], e_sum);
        if (band_sum != nullptr) {
            if (band0 < 16) atomicAdd(&acc[2 + band0], b_sum[0]);
            if (band0 + 1 < 16) atomicAdd(&acc[3 + band0], b_sum[1]);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(ssim_sum + img, acc[0]);
    if (threadIdx.x == 1) atomicAdd(sse + img, acc[1]);
    if (band_sum != nullptr && threadIdx.x >= 2 && threadIdx.x < 18 && acc[threadIdx.x] != 0.f)
        atomicAdd(band_sum + (size_t)img * 16 + (threadIdx.x - 2), acc[threadIdx.x]);
}

// =============================================================================================
// backward, pass 1: coef[n,h,w] = g_ssim[n] * (a, b, c, 0) at the interior window centres, else 0
// (a = dS/dG(p), b = dS/dG(p^2), c = dS/dG(pt); SURVEY.md Appendix A "Gradient").
template <typename T, bool DENORM>
__global__ void __launch_bounds__(kSsimThreads)
ssim_bwd_coef_kernel(const T* __restrict__ pred, const T* __restrict__ target, int h, int w,
                     const float* __restrict__ g_ssim, float4* __restrict__ coef) {
    extern __shared__ float4 smem4[];
    float4* in = smem4;
    float4* hbuf = smem4 + IH * IN_PITCH;
    const int img = blockIdx.z, y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
    const size_t img_off = (size_t)img * h * w;
    load_pair_patch<T, DENORM>(pred, target, img_off, h, w, y0, x0, in);
    __syncthreads();
    hpass(in, hbuf);
    __syncthreads();
    const int col = threadIdx.x & 63, rg = threadIdx.x >> 6;
    float4 m[RG];
    vpass(hbuf, col, rg, m);
    const int gx = x0 + col;
    const float gs = g_ssim[img];
#pragma unroll
    for (int o = 0; o < RG; ++o) {
        const int lr = rg * RG + o, gy = y0 + lr;
        if (lr < TH && gy < h && gx < w) {
            float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gx >= 5 && gx < w - 5 && gy >= 5 && gy < h - 5) {
                const SsimTerms t = ssim_terms(m[o]);
                const float inv = 1.f / (t.b1 * t.b2);
                const float s = t.a1 * t.a2 * inv;
                out.x = gs * (2.f * m[o].y * (t.a2 - t.a1) * inv - 2.f * m[o].x * s * (t.b2 - t.b1) * inv);
                out.y = gs * (-s / t.b2);
                out.z = gs * (2.f * t.a1 * inv);
            }
            coef[img_off + (size_t)gy * w + gx] = out;
        }
    }
}

// backward, pass 2: grad = G*(a) + 2 p G*(b) + t G*(c) + g_sse[n] * 2 (p - t), chained through the
// de-normalisation clamp(0.5 x + 0.5, 0, 1) when DENORM (models/utils.py:11).
template <typename T, bool DENORM>
__global__ void __launch_bounds__(kSsimThreads)
ssim_bwd_apply_kernel(const T* __restrict__ pred, const T* __restrict__ target, int h, int w,
                      const float4* __restrict__ coef, const float* __restrict__ g_sse, T* __restrict__ grad) {
    extern __shared__ float4 smem4[];
    float4* in = smem4;
    float4* hbuf = smem4 + IH * IN_PITCH;
    const int img = blockIdx.z, y0 = blockIdx.y * TH, x0 = blockIdx.x * TW;
    const size_t img_off = (size_t)img * h * w;
    for (int i = threadIdx.x; i < IH * IW; i += kSsimThreads) {
        const int r = i / IW, c = i - r * IW;
        const int gy = y0 - 5 + r, gx = x0 - 5 + c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gy >= 0 && gy < h && gx >= 0 && gx < w) v = __ldg(coef + img_off + (size_t)gy * w + gx);
        in[r * IN_PITCH + c] = v;
    }
    __syncthreads();
    hpass(in, hbuf);
    __syncthreads();
    const int col = threadIdx.x & 63, rg = threadIdx.x >> 6;
    float4 m[RG];
    vpass(hbuf, col, rg, m);
    const int gx = x0 + col;
    const float ge = g_sse != nullptr ? g_sse[img] : 0.f;
#pragma unroll
    for (int o = 0; o < RG; ++o) {
        const int lr = rg * RG + o, gy = y0 + lr;
        if (lr < TH && gy < h && gx < w) {
            const size_t idx = img_off + (size_t)gy * w + gx;
            float xp = load_val<T>(pred, idx), xt = load_val<T>(target, idx);
            float p = xp, t = xt, chain = 1.f;
            if (DENORM) {
                const float u = fmaf(xp, 0.5f, 0.5f);
                chain = (u >= 0.f && u <= 1.f) ? 0.5f : 0.f;
                p = denorm(xp);
                t = denorm(xt);
            }
            const float g = m[o].x + 2.f * p * m[o].y + t * m[o].z + ge * 2.f * (p - t);
            if (sizeof(T) == 4)
                reinterpret_cast<float*>(grad)[idx] = g * chain;
            else
                reinterpret_cast<__nv_bfloat16*>(grad)[idx] = __float2bfloat16_rn(g * chain);
        }
    }
}

// =============================================================================================
// Row-streaming forward for 256-pixel-wide images (the BASELINE.json shape): one CTA walks a chunk of
// output rows of one image pair in batches of 8 rows.
//   * input rows arrive by bulk async copies (cp.async.bulk, 8-row granules, 4-slot ring, mbarrier) --
//     every row is fetched from L2/HBM exactly once per chunk
//   * V phase (thread <-> column): vertical 11-tap filter of (p, t, p^2+t^2, p t) for 8 output rows out
//     of 18 ring rows, packed FFMA2, result to smem (column-major, pitch 9 float4 -> conflict-free)
//   * H phase (thread <-> (row, run of 8 columns)): horizontal 11-tap filter, SSIM formula, sums,
//     optional full map (reflect padding = mirrored smem halo columns / mirrored ring rows)
// FULL = reflect-padded map at all h x 256 positions (report.py:78-84); otherwise only the interior
// windows (rows / columns 5 .. dim-6) are evaluated, which is all the scalar metrics need.
namespace rows {

static constexpr int W = 256, B = 8, NT = 256, RING = 32, VP = 9, VC = W + 10;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return reinterpret_cast<float2&>(d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return reinterpret_cast<float2&>(d);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(reinterpret_cast<unsigned long long&>(a)), "l"(reinterpret_cast<unsigned long long&>(b)));
    return reinterpret_cast<float2&>(d);
}
__device__ __forceinline__ float ring_ld(const float* p) { return *p; }
__device__ __forceinline__ float ring_ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// 11-tap symmetric filter over a register window: out[o] = sum_k g[k] * in[o + k]
template <int NO>
__device__ __forceinline__ void filt11(const float2 (&in)[NO + 10], const float2 (&g)[6], float2 (&out)[NO]) {
#pragma unroll
    for (int o = 0; o < NO; ++o) {
        float2 s = mul2(in[o], g[0]);
#pragma unroll
        for (int k = 1; k < 11; ++k) s = fma2(in[o + k], g[k < 6 ? k : 10 - k], s);
        out[o] = s;
    }
}

template <typename T>
static constexpr size_t smem_bytes() {
    return (size_t)2 * RING * W * sizeof(T) + (size_t)VC * VP * sizeof(float4);
}

template <typename T, bool DENORM, bool FULL>
__global__ void __launch_bounds__(NT, 2)
ssim_fwd_rows_kernel(const T* __restrict__ pred, const T* __restrict__ target, int h, int rows_per_chunk,
                     int band_rows, float* __restrict__ ssim_sum, float* __restrict__ band_sum,
                     float* __restrict__ sse, float* __restrict__ full_map, const Gauss gk) {
    extern __shared__ __align__(128) uint8_t rows_smem[];
    T* ring_p = reinterpret_cast<T*>(rows_smem);
    T* ring_t = ring_p + RING * W;
    float4* vbuf = reinterpret_cast<float4*>(rows_smem + (size_t)2 * RING * W * sizeof(T));
    __shared__ uint64_t full_bar[4];
    __shared__ float acc_s[18];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int img = blockIdx.y;
    const size_t img_off = (size_t)img * h * W;
    const int y_lo = FULL ? 0 : 5, y_hi = FULL ? h : h - 5;
    const int y_begin = y_lo + blockIdx.x * rows_per_chunk;
    const int y_end = min(y_hi, y_begin + rows_per_chunk);
    if (y_begin >= y_end) return;
    const bool first_chunk = blockIdx.x == 0, last_chunk = y_end == y_hi;

    if (tid < 18) acc_s[tid] = 0.f;
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) mbar_init(&full_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    float2 g[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) g[i] = make_float2(gk.g[i], gk.g[i]);

    const int g_first = max(0, y_begin - 5) >> 3, g_last = (h - 1) >> 3;
    int g_issued = g_first - 1, g_waited = g_first - 1;
    float s_acc = 0.f, e_acc = 0.f;
    int band_b0 = band_sum != nullptr ? y_begin / band_rows : 0, band_y0 = band_b0 * band_rows;
    const int hrow = lane & 7, run = (lane >> 3) + 4 * warp;
    const int x = tid;

    for (int y0 = y_begin; y0 < y_end; y0 += B) {
        // ---- ring maintenance: granules up to g_need must have landed, one more is prefetched
        const int g_need = min(h - 1, y0 + 12) >> 3;
        if (tid == 0) {
            const int g_pref = min(g_last, g_need + 1);
            while (g_issued < g_pref) {
                const int gi = ++g_issued;
                const int r0 = gi * 8, nr = min(8, h - r0);
                const uint32_t bytes = (uint32_t)(nr * W * sizeof(T));
                uint64_t* bar = &full_bar[gi & 3];
                mbar_expect_tx(bar, 2 * bytes);
                bulk_g2s(ring_p + (gi & 3) * 8 * W, pred + img_off + (size_t)r0 * W, bytes, bar);
                bulk_g2s(ring_t + (gi & 3) * 8 * W, target + img_off + (size_t)r0 * W, bytes, bar);
            }
        }
        while (g_waited < g_need) {
            const int gi = ++g_waited;
            mbar_wait(&full_bar[gi & 3], ((gi - g_first) >> 2) & 1);
        }

        // ---- V phase
        {
            float2 a[B + 10], q[B + 10];
#pragma unroll
            for (int j = 0; j < B + 10; ++j) {
                int r = y0 - 5 + j;
                if (FULL) r = reflect_idx(r, h);
                const int s = (r & (RING - 1)) * W + x;
                float p = ring_ld(ring_p + s), t = ring_ld(ring_t + s);
                if (DENORM) {
                    p = denorm(p);
                    t = denorm(t);
                }
                a[j] = make_float2(p, t);
                q[j] = make_float2(fmaf(p, p, t * t), p * t);
            }
            // squared error: every image row is owned by exactly one batch of one chunk
            const int own_lo = (first_chunk && y0 == y_begin) ? 0 : y0;
            const int own_hi = (last_chunk && y0 + B >= y_end) ? h : min(y0 + B, y_end);
            if (own_lo == y0 && own_hi == y0 + B) {
#pragma unroll
                for (int o = 0; o < B; ++o) {
                    const float d = a[o + 5].x - a[o + 5].y;
                    e_acc = fmaf(d, d, e_acc);
                }
            } else {
#pragma unroll
                for (int j = 0; j < B + 10; ++j) {
                    const int r = y0 - 5 + j;
                    const float d = a[j].x - a[j].y;
                    if (r >= own_lo && r < own_hi) e_acc = fmaf(d, d, e_acc);
                }
            }
            float2 va[B], vq[B];
            filt11<B>(a, g, va);
            filt11<B>(q, g, vq);
            float4* dst = vbuf + (x + 5) * VP;
#pragma unroll
            for (int o = 0; o < B; ++o) dst[o] = make_float4(va[o].x, va[o].y, vq[o].x, vq[o].y);
            if (x >= 1 && x <= 5) {           // mirrored halo columns (reflect padding of the map's columns)
                float4* d2 = vbuf + (5 - x) * VP;
#pragma unroll
                for (int o = 0; o < B; ++o) d2[o] = make_float4(va[o].x, va[o].y, vq[o].x, vq[o].y);
            }
            if (x >= W - 6 && x <= W - 2) {
                float4* d2 = vbuf + (2 * W + 3 - x) * VP;     // column 255 + j  <-  column 255 - j
#pragma unroll
                for (int o = 0; o < B; ++o) d2[o] = make_float4(va[o].x, va[o].y, vq[o].x, vq[o].y);
            }
        }
        __syncthreads();

        // ---- H phase
        {
            float2 a[B + 10], q[B + 10];
            const float4* src = vbuf + (8 * run) * VP + hrow;
#pragma unroll
            for (int j = 0; j < B + 10; ++j) {
                const float4 v = src[j * VP];
                a[j] = make_float2(v.x, v.y);
                q[j] = make_float2(v.z, v.w);
            }
            float2 ma[B], mq[B];
            filt11<B>(a, g, ma);
            filt11<B>(q, g, mq);
            const int y = y0 + hrow;
            float sv[B];
            float part = 0.f;
            // SSIM formula on pixel pairs (packed f32x2): the same arithmetic as ssim_terms(), half the
            // FMA-pipe instructions
            const float2 c1 = make_float2(1e-4f, 1e-4f), c2 = make_float2(9e-4f, 9e-4f), two = make_float2(2.f, 2.f);
#pragma unroll
            for (int o = 0; o < B; o += 2) {
                const float2 mp = make_float2(ma[o].x, ma[o + 1].x), mt = make_float2(ma[o].y, ma[o + 1].y);
                const float2 eq = make_float2(mq[o].x, mq[o + 1].x), er = make_float2(mq[o].y, mq[o + 1].y);
                const float2 mpt = mul2(mp, mt);
                const float2 ss = fma2(mp, mp, mul2(mt, mt));              // mu_p^2 + mu_t^2
                const float2 a1 = fma2(two, mpt, c1);
                const float2 a2 = fma2(two, sub2(er, mpt), c2);
                const float2 b1 = add2(ss, c1);
                const float2 b2 = add2(sub2(eq, ss), c2);
                const float2 num = mul2(a1, a2), den = mul2(b1, b2);
                sv[o] = __fdividef(num.x, den.x);
                sv[o + 1] = __fdividef(num.y, den.y);
            }
            if (run == 0) {
#pragma unroll
                for (int o = 5; o < B; ++o) part += sv[o];
            } else if (run == 31) {
#pragma unroll
                for (int o = 0; o < 3; ++o) part += sv[o];
            } else {
#pragma unroll
                for (int o = 0; o < B; ++o) part += sv[o];
            }
            if (y < y_end) {
                if (FULL && full_map != nullptr) {
                    float4* o4 = reinterpret_cast<float4*>(full_map + img_off + (size_t)y * W + 8 * run);
                    o4[0] = make_float4(sv[0], sv[1], sv[2], sv[3]);
                    o4[1] = make_float4(sv[4], sv[5], sv[6], sv[7]);
                }
                if (!FULL || (y >= 5 && y < h - 5)) s_acc += part;
            }
            if (band_sum != nullptr) {
                // the 8 rows of a batch touch at most two depth bands (band_rows > 10): two predicated warp
                // reductions, one shared-memory atomic per warp and band
                while (y0 >= band_y0 + band_rows) band_y0 += band_rows, ++band_b0;
                const int b0 = band_b0;
                int br = y - band_y0;
                const int sel = br >= band_rows;
                if (sel) br -= band_rows;
                const bool inband = y < y_end && br >= 5 && br < band_rows - 5;
#pragma unroll
                for (int which = 0; which < 2; ++which) {
                    const bool mine = inband && sel == which;
                    if (__any_sync(0xffffffffu, mine)) {
                        const float v = warp_sum(mine ? part : 0.f);
                        if (lane == 0) atomicAdd(&acc_s[2 + b0 + which], v);
                    }
                }
            }
        }
        __syncthreads();
    }
    s_acc = warp_sum(s_acc);
    e_acc = warp_sum(e_acc);
    if (lane == 0) {
        atomicAdd(&acc_s[0], s_acc);
        atomicAdd(&acc_s[1], e_acc);
    }
    __syncthreads();
    if (tid == 0) atomicAdd(ssim_sum + img, acc_s[0]);
    if (tid == 1) atomicAdd(sse + img, acc_s[1]);
    if (band_sum != nullptr && tid >= 2 && tid < 18 && acc_s[tid] != 0.f)
        atomicAdd(band_sum + (size_t)img * 16 + (tid - 2), acc_s[tid]);
}

template <typename T, bool DENORM, bool FULL>
static int launch(const void* pred, const void* target, int n, int h, int band_rows, float* ssim_sum,
                  float* band_sum, float* sse, float* full_map, cudaStream_t st) {
    static DeviceOnce attr_once;
    const int attr_dev = current_device();
    if (attr_dev < 0) return -1;
    if (attr_once.need(attr_dev)) {
        PAI_CUDA_OK(cudaFuncSetAttribute(ssim_fwd_rows_kernel<T, DENORM, FULL>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<T>()));
        attr_once.mark(attr_dev);
    }
    const int out_rows = FULL ? h : h - 10;
    // enough CTAs for two resident per SM with a few waves; chunks are whole 8-row batches
    int chunks = (4 * 148 + n - 1) / n;
    int rpc = (out_rows + chunks - 1) / chunks;
    rpc = ((rpc + B - 1) / B) * B;
    chunks = (out_rows + rpc - 1) / rpc;
    for (int i0 = 0; i0 < n; i0 += 65535) {
        const int cnt = n - i0 < 65535 ? n - i0 : 65535;
        const size_t off = (size_t)i0 * h * W;
        ssim_fwd_rows_kernel<T, DENORM, FULL><<<dim3(chunks, cnt), NT, smem_bytes<T>(), st>>>(
            (const T*)pred + off, (const T*)target + off, h, rpc, band_rows, ssim_sum + i0,
            band_sum ? band_sum + (size_t)i0 * 16 : nullptr, sse + i0, full_map ? full_map + off : nullptr,
            host_gauss());
        PAI_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

}  // namespace rows


// =============================================================================================
// Streaming forward for large batches of 256-wide images (the report.py sweep, BASELINE.json configs[4]):
// a persistent, warp-specialised CTA per SM walks whole images as ONE continuous stream of rows.
//   * producer warp: cp.async.bulk of 8-row granules of pred / target into a 4-granule ring (mbarrier full / empty)
//   * 8 V warps (thread <-> column): every input row is read from the ring ONCE, its products p^2+t^2, p t are
//     formed once, and it is scattered into the 11 pending vertical-filter accumulators it contributes to
//     (registers, 16-slot circular naming, 16-row unrolled loop) -- 22 FFMA2 per pixel, no re-reads, no window
//     shifting; each completed row goes to a double-buffered smem batch of 8 rows; also the squared error
//   * 8 H warps (thread <-> (row, run of 8 columns)): horizontal 11-tap filter of the batch, SSIM formula on pixel
//     pairs, per-image / per-depth-band sums (warp shuffles + one global atomic per warp and sum)
// V and H warps run concurrently on different batches, so the FMA pipe stays busy across the smem hand-over.
// Accumulation order per output equals the row-batched kernel above (ascending tap index), results are bit-equal.
// Interior windows only (rows / columns 5 .. dim-6): what the scalar metrics need.
namespace stream {

static constexpr int W = 256, GR = 8, NG = 4, NB = 4, VP = 9, VC = W + 10;
static constexpr int NV = 256, NH = 256, NT = 32 + NV + NH;     // warp 0 producer, warps 1-8 V, warps 9-16 H

using rows::bulk_g2s;
using rows::mul2;
using rows::ring_ld;
using rows::sub2;

// Blocking wait without clock reads: try_wait suspends the thread in hardware for up to the hinted time, so a waiting
// warp costs (almost) no issue slots of the SM sub-partition it shares with the computing warps.  A pipeline bug still
// traps instead of hanging the GPU (bounded number of wake-ups).
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (int spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(100000u)
            : "memory");
        if (ok) return;
        if (spins > (1 << 22)) __trap();
    }
}

template <typename T>
static constexpr size_t smem_bytes() {
    return (size_t)2 * NG * GR * W * sizeof(T) + (size_t)NB * VC * VP * sizeof(float4);
}

template <typename T, bool DENORM>
__global__ void __launch_bounds__(NT, 1)
ssim_fwd_stream_kernel(const T* __restrict__ pred, const T* __restrict__ target, int n, int h, int band_rows,
                       float* __restrict__ ssim_sum, float* __restrict__ band_sum, float* __restrict__ sse,
                       const Gauss gk) {
    extern __shared__ __align__(128) uint8_t stream_smem[];
    T* ring_p = reinterpret_cast<T*>(stream_smem);
    T* ring_t = ring_p + NG * GR * W;
    float4* vbuf = reinterpret_cast<float4*>(stream_smem + (size_t)2 * NG * GR * W * sizeof(T));
    __shared__ uint64_t gfull[NG], gempty[NG], vfull[NB], vempty[NB];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_mine = (n - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // images blockIdx.x + i * gridDim.x
    if (n_mine <= 0) return;
    const int gran_per_img = h / GR;
    const int total_gran = n_mine * gran_per_img;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NG; ++i) {
            mbar_init(&gfull[i], 1);
            mbar_init(&gempty[i], NV / 32);
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            mbar_init(&vfull[i], NV / 32);
            mbar_init(&vempty[i], NH / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // halo columns of the batches are never written in interior mode: keep them finite
    for (int i = tid; i < NB * 10 * VP; i += NT) {
        const int b = i / (10 * VP), r = i - b * 10 * VP, c = r / VP, o = r - c * VP;
        vbuf[b * VC * VP + (c < 5 ? c : W + c) * VP + o] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    float2 g[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) g[i] = make_float2(gk.g[i], gk.g[i]);

    if (warp == 0) {
        // ------------------------------------------------------------------ producer warp
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(GR * W * sizeof(T));
            int img_i = 0, gi = 0;
            for (int G = 0; G < total_gran; ++G) {
                const int slot = G & (NG - 1);
                mbar_wait_sleep(&gempty[slot], ((G / NG) & 1) ^ 1);     // fresh barrier: passes for the first NG
                const size_t off = ((size_t)(blockIdx.x + (size_t)img_i * gridDim.x) * h + (size_t)gi * GR) * W;
                mbar_expect_tx(&gfull[slot], 2 * bytes);
                bulk_g2s(ring_p + slot * GR * W, pred + off, bytes, &gfull[slot]);
                bulk_g2s(ring_t + slot * GR * W, target + off, bytes, &gfull[slot]);
                if (++gi == gran_per_img) gi = 0, ++img_i;
            }
        }
    } else if (warp <= NV / 32) {
        // ------------------------------------------------------------------ V warps: thread <-> column
        const int x = tid - 32;
        float2 A[16], Q[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) A[i] = Q[i] = make_float2(0.f, 0.f);
        float e_acc = 0.f;
        int G = 0, vb = 0, vpar = 1, r16 = 0, img_i = 0;      // vpar: parity to wait on vempty[vb] (fresh: passes)
        const int blocks_per_img = h / 16;
        const int iters = n_mine * blocks_per_img;
        float4* vcol = vbuf + (x + 5) * VP;
        // the input row is fetched one row ahead, so its shared-memory latency hides behind the 22 FFMA2 of the row
        // being scattered
        mbar_wait_sleep(&gfull[0], 0);
        float pn = ring_ld(ring_p + x), tn = ring_ld(ring_t + x);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float p = pn, t = tn;
                if (j != 15 || it + 1 < iters) {
                    const bool cross = (j & 7) == 7;               // the next row opens the next granule
                    const int Gn = G + (cross ? 1 : 0);
                    if (cross) mbar_wait_sleep(&gfull[Gn & (NG - 1)], (Gn / NG) & 1);
                    const int s_n = ((Gn & (NG - 1)) * GR + ((j + 1) & 7)) * W + x;
                    pn = ring_ld(ring_p + s_n);
                    tn = ring_ld(ring_t + s_n);
                }
                if (DENORM) {
                    p = denorm(p);
                    t = denorm(t);
                }
                const float2 a = make_float2(p, t);
                const float2 q = make_float2(fmaf(p, p, t * t), p * t);
                const float d = p - t;
                e_acc = fmaf(d, d, e_acc);
                // input row r contributes g[k] to output row r + 5 - k (k = 0 initialises, k = 10 completes)
#pragma unroll
                for (int k = 0; k < 11; ++k) {
                    const int sl = (j + 5 - k) & 15;
                    const float2 gw = g[k < 6 ? k : 10 - k];
                    if (k == 0) {
                        A[sl] = mul2(a, gw);
                        Q[sl] = mul2(q, gw);
                    } else {
                        A[sl] = fma2(a, gw, A[sl]);
                        Q[sl] = fma2(q, gw, Q[sl]);
                    }
                }
                const int sl = (j - 5) & 15, ob = (j + 3) & 7;      // completed output row r - 5, its slot in the batch
                if (it > 0 || j >= 5) {
                    if (ob == 0) mbar_wait_sleep(&vempty[vb], vpar);
                    vcol[ob] = make_float4(A[sl].x, A[sl].y, Q[sl].x, Q[sl].y);
                    if (ob == 7) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&vfull[vb]);
                        vcol += VC * VP;
                        if (++vb == NB) vb = 0, vpar ^= 1, vcol -= NB * VC * VP;
                    }
                }
                if ((j & 7) == 7) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&gempty[G & (NG - 1)]);
                    ++G;
                }
            }
            if (++r16 == blocks_per_img) {          // last row of an image: its squared error is complete
                r16 = 0;
                const float e = warp_sum(e_acc);
                if (lane == 0) atomicAdd(sse + blockIdx.x + (size_t)img_i * gridDim.x, e);
                e_acc = 0.f;
                ++img_i;
            }
        }
        // the stream's last batch holds rows h-8 .. h-6 of the last image only: hand it over as it is
        __syncwarp();
        if (lane == 0) mbar_arrive(&vfull[vb]);
    } else {
        // ------------------------------------------------------------------ H warps: thread <-> (row, 8 columns)
        const int ht = tid - 32 - NV;
        const int hrow = ht & 7, run = ht >> 3;
        const int total_batches = total_gran;       // 8 output rows per batch, h / 8 batches per image
        float s_acc = 0.f, b_acc = 0.f;
        int y0 = 0, br0 = 0, band = 0, img_i = 0, hb = 0, hpar = 0;
        const float4* src = vbuf + (8 * run) * VP + hrow;
        const float2 one_two = make_float2(1.f, 2.f), c1 = make_float2(1e-4f, 1e-4f), c2 = make_float2(9e-4f, 9e-4f);
        for (int Bh = 0; Bh < total_batches; ++Bh) {
            mbar_wait_sleep(&vfull[hb], hpar);
            float2 a[GR + 10], q[GR + 10];
#pragma unroll
            for (int j = 0; j < GR + 10; ++j) {
                const float4 v = src[j * VP];
                a[j] = make_float2(v.x, v.y);
                q[j] = make_float2(v.z, v.w);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&vempty[hb]);     // the batch now lives in registers
            src += VC * VP;
            if (++hb == NB) hb = 0, hpar ^= 1, src -= NB * VC * VP;
            float2 ma[GR], mq[GR];
            rows::filt11<GR>(a, g, ma);
            rows::filt11<GR>(q, g, mq);
            // SSIM on plane pairs: ma = (mu_p, mu_t), mq = (E[p^2 + t^2], E[p t]); u = (mu_p^2 + mu_t^2, mu_p mu_t);
            // (b1, a1) = u * (1, 2) + c1, (b2, a2) = (mq - u) * (1, 2) + c2, (den, num) = (b1 b2, a1 a2) -- the same
            // roundings as ssim_terms(), 8 FMA-pipe instructions per pixel and no register shuffling
            float sv[GR];
#pragma unroll
            for (int o = 0; o < GR; ++o) {
                const float2 sq = mul2(ma[o], make_float2(ma[o].x, ma[o].x));      // (mu_p^2, mu_p mu_t): mu_p broadcast
                const float2 u = make_float2(fmaf(ma[o].y, ma[o].y, sq.x), sq.y);
                const float2 ba1 = fma2(u, one_two, c1);
                const float2 ba2 = fma2(sub2(mq[o], u), one_two, c2);
                const float2 dn = mul2(ba1, ba2);
                float inv;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(dn.x));
                sv[o] = dn.y * inv;
            }
            float part = 0.f;
            if (run == 0) {
#pragma unroll
                for (int o = 5; o < GR; ++o) part += sv[o];
            } else if (run == W / 8 - 1) {
#pragma unroll
                for (int o = 0; o < 3; ++o) part += sv[o];
            } else {
#pragma unroll
                for (int o = 0; o < GR; ++o) part += sv[o];
            }
            const int y = y0 + hrow;
            if (y >= 5 && y < h - 5) s_acc += part;
            if (band_sum != nullptr) {
                const int br = br0 + hrow;
                if (br >= 5 && br < band_rows - 5) b_acc += part;
            }
            y0 += GR;
            br0 += GR;
            if (band_sum != nullptr && br0 == band_rows) {       // band complete (band_rows is a multiple of 8)
                const float v = warp_sum(b_acc);
                if (lane == 0) atomicAdd(band_sum + (blockIdx.x + (size_t)img_i * gridDim.x) * 16 + band, v);
                b_acc = 0.f;
                br0 = 0;
                ++band;
            }
            if (y0 == h) {
                const float v = warp_sum(s_acc);
                if (lane == 0) atomicAdd(ssim_sum + blockIdx.x + (size_t)img_i * gridDim.x, v);
                s_acc = 0.f;
                y0 = 0;
                band = 0;
                br0 = 0;
                ++img_i;
            }
        }
    }
}

template <typename T, bool DENORM>
static int launch(const void* pred, const void* target, int n, int h, int band_rows, float* ssim_sum, float* band_sum,
                  float* sse, cudaStream_t st) {
    static DeviceOnce attr_once;
    const int attr_dev = current_device();
    if (attr_dev < 0) return -1;
    if (attr_once.need(attr_dev)) {
        PAI_CUDA_OK(cudaFuncSetAttribute(ssim_fwd_stream_kernel<T, DENORM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem_bytes<T>()));
        attr_once.mark(attr_dev);
    }
    int dev = 0, sms = 148;
    PAI_CUDA_OK(cudaGetDevice(&dev));
    PAI_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = n < sms ? n : sms;
    ssim_fwd_stream_kernel<T, DENORM><<<grid, NT, smem_bytes<T>(), st>>>((const T*)pred, (const T*)target, n, h, band_rows,
                                                                         ssim_sum, band_sum, sse, host_gauss());
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

// whole images per CTA: worth it once every SM gets several images
static bool eligible(int n, int h, int w, int band_rows, const void* full_map) {
    return w == W && full_map == nullptr && h % 16 == 0 && h >= 32 && (band_rows == 0 || band_rows % GR == 0) &&
           n >= 4 * 148;
}

}  // namespace stream

template <typename T, bool DENORM>
static int fwd_launch(const void* pred, const void* target, int n, int h, int w, int band_rows, float* ssim_sum,
                      float* band_sum, float* sse, float* full_map, cudaStream_t st) {
    static DeviceOnce attr_once;
    const int attr_dev = current_device();
    if (attr_dev < 0) return -1;
    if (attr_once.need(attr_dev)) {
        PAI_CUDA_OK(cudaFuncSetAttribute(ssim_fwd_kernel<T, DENORM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSsimSmem));
        attr_once.mark(attr_dev);
    }
    dim3 grid((w + TW - 1) / TW, (h + TH - 1) / TH, n);
    ssim_fwd_kernel<T, DENORM><<<grid, kSsimThreads, kSsimSmem, st>>>(
        (const T*)pred, (const T*)target, h, w, band_rows, ssim_sum, band_sum, sse, full_map);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

template <typename T, bool DENORM>
static int bwd_launch(const void* pred, const void* target, int n, int h, int w, const float* g_ssim,
                      const float* g_sse, float4* coef, void* grad, cudaStream_t st) {
    static DeviceOnce attr_once;
    const int attr_dev = current_device();
    if (attr_dev < 0) return -1;
    if (attr_once.need(attr_dev)) {
        PAI_CUDA_OK(cudaFuncSetAttribute(ssim_bwd_coef_kernel<T, DENORM>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSsimSmem));
        PAI_CUDA_OK(cudaFuncSetAttribute(ssim_bwd_apply_kernel<T, DENORM>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSsimSmem));
        attr_once.mark(attr_dev);
    }
    dim3 grid((w + TW - 1) / TW, (h + TH - 1) / TH, n);
    ssim_bwd_coef_kernel<T, DENORM><<<grid, kSsimThreads, kSsimSmem, st>>>((const T*)pred, (const T*)target, h, w,
                                                                            g_ssim, coef);
    PAI_CUDA_OK(cudaGetLastError());
    ssim_bwd_apply_kernel<T, DENORM><<<grid, kSsimThreads, kSsimSmem, st>>>((const T*)pred, (const T*)target, h, w,
                                                                             coef, g_sse, (T*)grad);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace pai

using namespace pai;

extern "C" {

int pai_ssim_psnr_fwd(const void* pred, const void* target, int dtype, int n, int h, int w, int denormalize,
                      float* ssim_sum, float* band_sum, float* sse, float* full_map, void* stream) {
    PAI_REQUIRE(pred && target && ssim_sum && sse, "pai_ssim_psnr_fwd: null pointer");
    PAI_REQUIRE(n >= 0 && h > 10 && w > 10, "pai_ssim_psnr_fwd: images must be larger than the 11x11 window (got %dx%d)",
                h, w);
    PAI_REQUIRE(dtype == PAI_DTYPE_F32 || dtype == PAI_DTYPE_BF16, "pai_ssim_psnr_fwd: bad dtype %d", dtype);
    int band_rows = 0;
    if (band_sum != nullptr) {
        PAI_REQUIRE(h % 16 == 0 && h / 16 > 10, "pai_ssim_psnr_fwd: depth bands need h %% 16 == 0 and h/16 > 10 (h=%d)", h);
        band_rows = h / 16;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return 0;
    if (int rc = upload_gauss()) return rc;
    PAI_CUDA_OK(cudaMemsetAsync(ssim_sum, 0, sizeof(float) * n, st));
    PAI_CUDA_OK(cudaMemsetAsync(sse, 0, sizeof(float) * n, st));
    if (band_sum) PAI_CUDA_OK(cudaMemsetAsync(band_sum, 0, sizeof(float) * 16 * n, st));
    const bool aligned16 = (reinterpret_cast<uintptr_t>(pred) & 15) == 0 && (reinterpret_cast<uintptr_t>(target) & 15) == 0;
    // large batches of 256-wide images without the full map (the evaluation sweep): persistent streaming kernel
    if (aligned16 && stream::eligible(n, h, w, band_rows, full_map) && !getenv("PAI_SSIM_NO_STREAM")) {
        if (dtype == PAI_DTYPE_F32)
            return denormalize ? stream::launch<float, true>(pred, target, n, h, band_rows, ssim_sum, band_sum, sse, st)
                               : stream::launch<float, false>(pred, target, n, h, band_rows, ssim_sum, band_sum, sse, st);
        return denormalize
                   ? stream::launch<__nv_bfloat16, true>(pred, target, n, h, band_rows, ssim_sum, band_sum, sse, st)
                   : stream::launch<__nv_bfloat16, false>(pred, target, n, h, band_rows, ssim_sum, band_sum, sse, st);
    }
    // 256-wide images (every shape of the reference's pipeline) take the row-streaming kernel
    if (w == rows::W && (reinterpret_cast<uintptr_t>(pred) & 15) == 0 && (reinterpret_cast<uintptr_t>(target) & 15) == 0 &&
        (full_map == nullptr || (reinterpret_cast<uintptr_t>(full_map) & 15) == 0)) {
#define PAI_ROWS(T, DN, FL) rows::launch<T, DN, FL>(pred, target, n, h, band_rows, ssim_sum, band_sum, sse, full_map, st)
        const bool full = full_map != nullptr;
        if (dtype == PAI_DTYPE_F32) {
            if (denormalize) return full ? PAI_ROWS(float, true, true) : PAI_ROWS(float, true, false);
            return full ? PAI_ROWS(float, false, true) : PAI_ROWS(float, false, false);
        }
        if (denormalize) return full ? PAI_ROWS(__nv_bfloat16, true, true) : PAI_ROWS(__nv_bfloat16, true, false);
        return full ? PAI_ROWS(__nv_bfloat16, false, true) : PAI_ROWS(__nv_bfloat16, false, false);
#undef PAI_ROWS
    }
    if (dtype == PAI_DTYPE_F32)
        return denormalize ? fwd_launch<float, true>(pred, target, n, h, w, band_rows, ssim_sum, band_sum, sse, full_map, st)
                           : fwd_launch<float, false>(pred, target, n, h, w, band_rows, ssim_sum, band_sum, sse, full_map, st);
    return denormalize
               ? fwd_launch<__nv_bfloat16, true>(pred, target, n, h, w, band_rows, ssim_sum, band_sum, sse, full_map, st)
               : fwd_launch<__nv_bfloat16, false>(pred, target, n, h, w, band_rows, ssim_sum, band_sum, sse, full_map, st);
}

int pai_ssim_psnr_bwd(const void* pred, const void* target, int dtype, int n, int h, int w, int denormalize,
                      const float* g_ssim_sum, const float* g_sse, void* workspace, void* grad_pred, void* stream) {
    PAI_REQUIRE(pred && target && g_ssim_sum && workspace && grad_pred, "pai_ssim_psnr_bwd: null pointer");
    PAI_REQUIRE(n >= 0 && h > 10 && w > 10, "pai_ssim_psnr_bwd: images must be larger than the 11x11 window");
    PAI_REQUIRE(dtype == PAI_DTYPE_F32 || dtype == PAI_DTYPE_BF16, "pai_ssim_psnr_bwd: bad dtype %d", dtype);
    PAI_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "pai_ssim_psnr_bwd: workspace must be 16 B aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return 0;
    if (int rc = upload_gauss()) return rc;
    float4* coef = reinterpret_cast<float4*>(workspace);
    if (dtype == PAI_DTYPE_F32)
        return denormalize ? bwd_launch<float, true>(pred, target, n, h, w, g_ssim_sum, g_sse, coef, grad_pred, st)
                           : bwd_launch<float, false>(pred, target, n, h, w, g_ssim_sum, g_sse, coef, grad_pred, st);
    return denormalize ? bwd_launch<__nv_bfloat16, true>(pred, target, n, h, w, g_ssim_sum, g_sse, coef, grad_pred, st)
                       : bwd_launch<__nv_bfloat16, false>(pred, target, n, h, w, g_ssim_sum, g_sse, coef, grad_pred, st);
}

long long pai_ssim_bwd_workspace_bytes(int n, int h, int w) { return (long long)n * h * w * 16; }

}  // extern "C"
