// BatchNorm2d (train / eval), pre-activation and activation-backward kernels on NHWC bf16 tensors.
//
// Reference call sites: nn.BatchNorm2d(eps=1e-5, momentum=0.1) after every Conv / ConvT of the U-Net
// (models/pix2pix.py:70,106), the non-inplace pre-activations nn.LeakyReLU(0.2) / nn.ReLU()
// (models/pix2pix.py:62,98) that consume the same BN output on the encoder and decoder side, the skip
// torch.cat (models/pix2pix.py:212) and the PatchGAN LeakyReLU (models/wrapper.py:196-206).
//
// Layout: x is [m pixels][c channels] bf16 with `ld` elements between pixels (so outputs can land in
// a channel slot of a concat buffer: zero-copy torch.cat).  Every thread owns 8 consecutive channels
// (one 16-byte vector) and walks pixels with a grid stride; all kernels are HBM-bound streams.
#include "pai_common.cuh"
#include "pai_kernels.h"

namespace pai {

static constexpr int kEwThreads = 256;

struct Vec8 {
    float v[8];
};
__device__ __forceinline__ Vec8 load8(const __nv_bfloat16* p) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    Vec8 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        r.v[2 * i] = f.x;
        r.v[2 * i + 1] = f.y;
    }
    return r;
}
__device__ __forceinline__ uint4 load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ Vec8 unpack8(const uint4& u) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    Vec8 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        r.v[2 * i] = f.x;
        r.v[2 * i + 1] = f.y;
    }
    return r;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const Vec8& r) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ Vec8 loadf8(const float* p) {
    Vec8 r;
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    r.v[0] = a.x, r.v[1] = a.y, r.v[2] = a.z, r.v[3] = a.w, r.v[4] = b.x, r.v[5] = b.y, r.v[6] = b.z, r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ float act_fwd(float z, int act, float slope) {
    if (act == PAI_ACT_LEAKY) return z > 0.f ? z : z * slope;
    if (act == PAI_ACT_RELU) return fmaxf(z, 0.f);
    return z;
}
__device__ __forceinline__ float act_grad(float z, int act, float slope) {
    if (act == PAI_ACT_LEAKY) return z > 0.f ? 1.f : slope;
    if (act == PAI_ACT_RELU) return z > 0.f ? 1.f : 0.f;
    return 1.f;
}

// Block-level reduction of per-thread 8-channel partial sums (two quantities) followed by one
// atomicAdd per channel: threads with equal (threadIdx.x % cv) own the same channels.
// Grid of the streaming kernels below: ONE wave -- as many blocks as are resident at once (occupancy x SMs), never
// more than the rows need.  The former fixed 148 x 8 blocks ran as 1.6 - 2.7 waves of 3 - 5 resident blocks: every wave
// pays the cold-load latency and (reduce kernels) the block-reduction + atomics tail again, and the last wave is partly
// empty (bn_stats of 16.8 MB took 20 us).
template <auto Kernel>
static int wave_grid(long long m, int c) {
    static int per_sm[64] = {};
    const int dev = current_device();
    if (dev < 0) return 1;
    int nb = __atomic_load_n(&per_sm[dev & 63], __ATOMIC_RELAXED);
    if (nb == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, Kernel, kEwThreads, 0) != cudaSuccess || nb < 1) nb = 2;
        (void)cudaGetLastError();
        __atomic_store_n(&per_sm[dev & 63], nb, __ATOMIC_RELAXED);
    }
    const long long ppb = kEwThreads / (c >> 3);
    long long blocks = (m + ppb - 1) / ppb;
    const long long cap = (long long)sm_count(dev) * nb;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

__device__ __forceinline__ void block_reduce_2x8(const Vec8& a, const Vec8& b, int cv, int c, float* out) {
    __shared__ float red[kEwThreads * 16];
    float* mine = red + threadIdx.x * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        mine[i] = a.v[i];
        mine[8 + i] = b.v[i];
    }
    __syncthreads();
    // thread t < cv*16 sums quantity (t / (cv*8)), channel (t % (cv*8)) over the kEwThreads/cv replicas
    for (int t = threadIdx.x; t < cv * 16; t += kEwThreads) {
        const int which = t / (cv * 8), ch = t - which * cv * 8;
        const int vec = ch >> 3, lane = ch & 7;
        float s = 0.f;
        for (int r = vec; r < kEwThreads; r += cv) s += red[r * 16 + which * 8 + lane];
        atomicAdd(out + which * c + ch, s);
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kEwThreads)
bn_stats_kernel(const __nv_bfloat16* __restrict__ x, long long m, int c, int ld, float* __restrict__ sums) {
    const int cv = c >> 3;
    const int vec = threadIdx.x % cv;
    const long long ppb = kEwThreads / cv;  // pixels per block-iteration
    Vec8 s, q;
#pragma unroll
    for (int i = 0; i < 8; ++i) s.v[i] = q.v[i] = 0.f;
    // 4 independent 16-byte loads in flight per thread (the single-load loop ran at 1.6 TB/s: latency-bound)
    constexpr int U = 4;
    const long long stride = (long long)gridDim.x * ppb;
    long long pix = (long long)blockIdx.x * ppb + threadIdx.x / cv;
    const __nv_bfloat16* xp = x + vec * 8;
    for (; pix + (U - 1) * stride < m; pix += U * stride) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) raw[u] = load_raw(xp + (pix + u * stride) * ld);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const Vec8 v = unpack8(raw[u]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                s.v[i] += v.v[i];
                q.v[i] = fmaf(v.v[i], v.v[i], q.v[i]);
            }
        }
    }
    for (; pix < m; pix += stride) {
        const Vec8 v = load8(xp + pix * ld);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s.v[i] += v.v[i];
            q.v[i] = fmaf(v.v[i], v.v[i], q.v[i]);
        }
    }
    block_reduce_2x8(s, q, cv, c, sums);
}

// ss layout: [scale | shift | mean | invstd], each c floats
__global__ void bn_finalize_kernel(const float* __restrict__ sums, long long m, int c, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, int training,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ ss, int nparts) {
    pdl_launch_dependents();
    pdl_wait();
    // block = 32 channels x 32 row groups: the partial rows are summed by 32 threads per channel (coalesced 128-byte
    // loads, 5 rows each for the 160 per-CTA rows of the fused statistics), then thread row 0 finishes the channel
    constexpr int RG = 32;
    __shared__ float red_s[RG][32], red_q[RG][32];
    const int chl = threadIdx.x & 31, rg = threadIdx.x >> 5;
    const int ch = blockIdx.x * 32 + chl;
    float mean, var;
    if (training) {
        // nparts > 1: per-CTA partial sums written by the implicit-GEMM epilogue (pai_conv4x4_fprop_bnstats)
        float s = 0.f, q = 0.f;
        if (ch < c)
            for (int r = rg; r < nparts; r += RG) {
                s += sums[(size_t)r * 2 * c + ch];
                q += sums[(size_t)r * 2 * c + c + ch];
            }
        red_s[rg][chl] = s;
        red_q[rg][chl] = q;
        __syncthreads();
        if (rg != 0 || ch >= c) return;
#pragma unroll
        for (int r = 1; r < RG; ++r) {
            s += red_s[r][chl];
            q += red_q[r][chl];
        }
        mean = s / (float)m;
        var = fmaxf(q / (float)m - mean * mean, 0.f);
        if (running_mean != nullptr) {
            const float unbiased = m > 1 ? var * ((float)m / (float)(m - 1)) : var;
            running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * mean;
            running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * unbiased;
        }
    } else {
        if (rg != 0 || ch >= c) return;
        mean = running_mean[ch];
        var = running_var[ch];
    }
    const float invstd = rsqrtf(var + eps);
    const float g = gamma != nullptr ? gamma[ch] : 1.f, b = beta != nullptr ? beta[ch] : 0.f;
    ss[ch] = g * invstd;
    ss[c + ch] = b - mean * g * invstd;
    ss[2 * c + ch] = mean;
    ss[3 * c + ch] = invstd;
}

template <int ACT>
__device__ __forceinline__ float act_fwd_t(float z, float slope) {
    if (ACT == PAI_ACT_LEAKY) return fmaxf(z, z * slope);      // slope in (0, 1)
    if (ACT == PAI_ACT_RELU) return fmaxf(z, 0.f);
    return z;
}

template <int ACT1, int ACT2 /* -1: no second output */>
__global__ void __launch_bounds__(kEwThreads)
bn_apply_act_kernel(const __nv_bfloat16* __restrict__ x, long long m, int c, int ld, const float* __restrict__ ss,
                    __nv_bfloat16* __restrict__ o1, int ld1, __nv_bfloat16* __restrict__ o2, int ld2, float slope) {
    pdl_launch_dependents();
    pdl_wait();
    const int cv = c >> 3;
    const int vec = threadIdx.x % cv;
    const long long ppb = kEwThreads / cv;
    Vec8 sc, sh;
#pragma unroll
    for (int i = 0; i < 8; ++i) sc.v[i] = 1.f, sh.v[i] = 0.f;
    if (ss != nullptr) {
        sc = loadf8(ss + vec * 8);
        sh = loadf8(ss + c + vec * 8);
    }
    for (long long pix = (long long)blockIdx.x * ppb + threadIdx.x / cv; pix < m; pix += (long long)gridDim.x * ppb) {
        const Vec8 v = load8(x + pix * ld + vec * 8);
        Vec8 a, b;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float z = fmaf(v.v[i], sc.v[i], sh.v[i]);
            a.v[i] = act_fwd_t<ACT1>(z, slope);
            if (ACT2 >= 0) b.v[i] = act_fwd_t<(ACT2 >= 0 ? ACT2 : 0)>(z, slope);
        }
        store8(o1 + pix * ld1 + vec * 8, a);
        if (ACT2 >= 0) store8(o2 + pix * ld2 + vec * 8, b);
    }
}

// dz = g1 * act1'(z) + g2 * act2'(z),  z = scale * x + shift,  xhat = (x - mean) * invstd
// MODE 0: sums = (sum dz, sum dz*xhat);  MODE 1: dx from the finished sums;  MODE 2 (no BN): dx = dz and
// sums[0:c] = sum dz in the same pass.  The activations and the presence of g2 / BN are template
// parameters: the kernels are streams whose instruction count per byte decides whether HBM saturates.
template <int ACT>
__device__ __forceinline__ float act_grad_t(float z, float slope) {
    if (ACT == PAI_ACT_LEAKY) return z > 0.f ? 1.f : slope;
    if (ACT == PAI_ACT_RELU) return z > 0.f ? 1.f : 0.f;
    return 1.f;
}

template <int ACT1, int ACT2 /* -1: no g2 */, bool BN, int MODE>
__global__ void __launch_bounds__(kEwThreads)
bn_bwd_kernel(const __nv_bfloat16* __restrict__ x, long long m, int c, int ld, const float* __restrict__ ss,
              const __nv_bfloat16* __restrict__ g1, int ldg1, const __nv_bfloat16* __restrict__ g2, int ldg2,
              float slope, float* __restrict__ sums, const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dx,
              int lddx) {
    pdl_launch_dependents();
    pdl_wait();
    const int cv = c >> 3;
    const int vec = threadIdx.x % cv;
    const long long ppb = kEwThreads / cv;
    // Few per-channel constants stay live in the loop (register pressure decides how many 16-byte loads a thread keeps
    // in flight): z = sc * x + sh only supplies the activation masks; the x-hat terms are folded into
    //   reduce: sum dz * xhat = invstd * (sum dz * x - mean * sum dz)          (finished after the loop)
    //   apply : dx = k0 * dz + kA + kB * x,  k0 = gamma * invstd, kA = -k0 * (mean_dz - mean * invstd * mean_dzxhat),
    //                                         kB = -k0 * invstd * mean_dzxhat
    Vec8 sc, sh;
#pragma unroll
    for (int i = 0; i < 8; ++i) sc.v[i] = 1.f, sh.v[i] = 0.f;
    if (BN) {
        sc = loadf8(ss + vec * 8);
        sh = loadf8(ss + c + vec * 8);
    }
    Vec8 k0, kA, kB;
    Vec8 s0, s1;
#pragma unroll
    for (int i = 0; i < 8; ++i) s0.v[i] = s1.v[i] = 0.f, k0.v[i] = 1.f, kA.v[i] = 0.f, kB.v[i] = 0.f;
    if (MODE == 1 && BN) {
        const Vec8 mu = loadf8(ss + 2 * c + vec * 8), is = loadf8(ss + 3 * c + vec * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float g = gamma != nullptr ? gamma[vec * 8 + i] : 1.f;
            const float mdz = sums[vec * 8 + i] / (float)m, mdzx = sums[c + vec * 8 + i] / (float)m;
            k0.v[i] = g * is.v[i];
            kB.v[i] = -k0.v[i] * is.v[i] * mdzx;
            kA.v[i] = -k0.v[i] * mdz - kB.v[i] * mu.v[i];
        }
    }
    const long long stride = (long long)gridDim.x * ppb;
    // one pixel of this thread's 8 channels
    auto body = [&](const uint4& rv, const uint4& ra, const uint4& rb, long long pix) {
        const Vec8 v = unpack8(rv), a = unpack8(ra);
        Vec8 b;
        if (ACT2 >= 0) b = unpack8(rb);
        Vec8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float z = BN ? fmaf(v.v[i], sc.v[i], sh.v[i]) : v.v[i];
            float dz = a.v[i] * act_grad_t<ACT1>(z, slope);
            if (ACT2 >= 0) dz = fmaf(b.v[i], act_grad_t<(ACT2 >= 0 ? ACT2 : 0)>(z, slope), dz);
            if (BN) {
                if (MODE == 1)
                    o.v[i] = fmaf(k0.v[i], dz, fmaf(kB.v[i], v.v[i], kA.v[i]));
                else
                    s1.v[i] = fmaf(dz, v.v[i], s1.v[i]);
            } else {
                o.v[i] = dz;
            }
            if (MODE != 1) s0.v[i] += dz;
        }
        if (MODE != 0) store8(dx + pix * lddx + vec * 8, o);
    };
    // Software pipeline: two register sets of U pixels (2-3 independent 16-byte loads each); the loads of the NEXT U pixels
    // are issued before the arithmetic of the current ones, so a warp's memory latency overlaps its own ~400 instructions
    // instead of alternating with them (the single-set loop reached 3.0-3.9 TB/s on the read-only reduce modes)
    constexpr int U = (ACT2 >= 0 || MODE == 1) ? 2 : 4;       // (measured per variant: registers decide the occupancy)
    long long pix = (long long)blockIdx.x * ppb + threadIdx.x / cv;
    uint4 rvA[U], raA[U], rbA[U], rvB[U], raB[U], rbB[U];
    auto load_set = [&](uint4 (&rv)[U], uint4 (&ra)[U], uint4 (&rb)[U], long long p0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long pp = p0 + u * stride;
            rv[u] = load_raw(x + pp * ld + vec * 8);
            ra[u] = load_raw(g1 + pp * ldg1 + vec * 8);
            rb[u] = ACT2 >= 0 ? load_raw(g2 + pp * ldg2 + vec * 8) : make_uint4(0, 0, 0, 0);
        }
    };
    auto run_set = [&](const uint4 (&rv)[U], const uint4 (&ra)[U], const uint4 (&rb)[U], long long p0) {
#pragma unroll
        for (int u = 0; u < U; ++u) body(rv[u], ra[u], rb[u], p0 + u * stride);
    };
    const long long step = (long long)U * stride;
    bool have = pix + (U - 1) * stride < m;
    if (have) load_set(rvA, raA, rbA, pix);
    while (have) {
        bool next = pix + step + (U - 1) * stride < m;
        if (next) load_set(rvB, raB, rbB, pix + step);
        run_set(rvA, raA, rbA, pix);
        pix += step;
        if (!next) break;
        next = pix + step + (U - 1) * stride < m;
        if (next) load_set(rvA, raA, rbA, pix + step);
        run_set(rvB, raB, rbB, pix);
        pix += step;
        have = next;
    }
    for (; pix < m; pix += stride) {
        const uint4 v0 = load_raw(x + pix * ld + vec * 8), a0 = load_raw(g1 + pix * ldg1 + vec * 8);
        uint4 b0 = make_uint4(0, 0, 0, 0);
        if (ACT2 >= 0) b0 = load_raw(g2 + pix * ldg2 + vec * 8);
        body(v0, a0, b0, pix);
    }
    if (MODE != 1) {
        if (BN) {       // sum dz * xhat from sum dz * x and sum dz
            const Vec8 mu = loadf8(ss + 2 * c + vec * 8), is = loadf8(ss + 3 * c + vec * 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) s1.v[i] = is.v[i] * (s1.v[i] - mu.v[i] * s0.v[i]);
        }
        block_reduce_2x8(s0, s1, cv, c, sums);
    }
}

// ---------------------------------------------------------------------------------------------
// Small layers (<= 8x8 at batch 64: a few MB, every pass over them is launch latency): BatchNorm is per channel, so ONE
// block that owns 8 channels over ALL pixels needs no grid-wide reduction.  It keeps its [m][8] slice in shared memory
// and does statistics -> finalize (running statistics, scale_shift) -> normalise + activation (+ Dropout2d mask) in
// one launch instead of memset + bn_stats + bn_finalize + bn_apply_act + scale_channels; the backward twin replaces
// scale_channels + memset + bn_bwd_reduce + bn_bwd_apply.  Same arithmetic as those kernels.
static constexpr int kSmallThreads = 1024;      // forward: 4 pixels per thread at 8x8 x batch 64, one batch of loads
static constexpr int kSmallBwdThreads = 512;    // backward keeps ~50 values per thread live

__device__ __forceinline__ void block_sum16(float (&acc)[16], float (*red)[16], float* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        float v = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
        total[threadIdx.x] = v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kSmallThreads)
bn_small_fwd_kernel(const __nv_bfloat16* __restrict__ x, int m, int c, int ld, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                    float* __restrict__ running_var, float* __restrict__ ss, __nv_bfloat16* __restrict__ o1, int ld1,
                    int act1, __nv_bfloat16* __restrict__ o2, int ld2, int act2, float slope,
                    const float* __restrict__ mask, int ppi) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ uint4 small_vals[];
    __shared__ float red[kSmallThreads / 32][16];
    __shared__ float total[16], coef[16];
    const int ch0 = blockIdx.x * 8;
    // the 8 finishing threads fetch their per-channel parameters now: the latency hides under the statistics pass
    float pg = 1.f, pb = 0.f, prm = 0.f, prv = 0.f;
    if (threadIdx.x < 8) {
        const int ch = ch0 + threadIdx.x;
        if (gamma != nullptr) pg = gamma[ch];
        if (beta != nullptr) pb = beta[ch];
        if (running_mean != nullptr) prm = running_mean[ch], prv = running_var[ch];
    }
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll 4
    for (int pix = threadIdx.x; pix < m; pix += kSmallThreads) {
        const uint4 r = load_raw(x + (size_t)pix * ld + ch0);
        small_vals[pix] = r;
        const Vec8 v = unpack8(r);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i] += v.v[i];
            acc[8 + i] = fmaf(v.v[i], v.v[i], acc[8 + i]);
        }
    }
    block_sum16(acc, red, total);
    if (threadIdx.x < 8) {
        const int ch = ch0 + threadIdx.x;
        const float mean = total[threadIdx.x] / (float)m;
        const float var = fmaxf(total[8 + threadIdx.x] / (float)m - mean * mean, 0.f);
        if (running_mean != nullptr) {
            const float unbiased = m > 1 ? var * ((float)m / (float)(m - 1)) : var;
            running_mean[ch] = (1.f - momentum) * prm + momentum * mean;
            running_var[ch] = (1.f - momentum) * prv + momentum * unbiased;
        }
        const float invstd = rsqrtf(var + eps);
        const float g = pg, b = pb;
        ss[ch] = coef[threadIdx.x] = g * invstd;
        ss[c + ch] = coef[8 + threadIdx.x] = b - mean * g * invstd;
        ss[2 * c + ch] = mean;
        ss[3 * c + ch] = invstd;
    }
    __syncthreads();
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sc[i] = coef[i], sh[i] = coef[8 + i];
#pragma unroll 2
    for (int pix = threadIdx.x; pix < m; pix += kSmallThreads) {
        const Vec8 v = unpack8(small_vals[pix]);
        Vec8 mk;
#pragma unroll
        for (int i = 0; i < 8; ++i) mk.v[i] = 1.f;
        if (mask != nullptr) mk = loadf8(mask + (size_t)(pix / ppi) * c + ch0);
        Vec8 a, b;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float z = fmaf(v.v[i], sc[i], sh[i]);
            a.v[i] = act_fwd(z, act1, slope) * mk.v[i];
            b.v[i] = act_fwd(z, act2, slope);
        }
        store8(o1 + (size_t)pix * ld1 + ch0, a);
        if (o2 != nullptr) store8(o2 + (size_t)pix * ld2 + ch0, b);
    }
}

// smem: x slice, g1 slice and (g2 != NULL) g2 slice, [m] uint4 each
__global__ void __launch_bounds__(kSmallBwdThreads)
bn_small_bwd_kernel(const __nv_bfloat16* __restrict__ x, int m, int c, int ld, const float* __restrict__ ss,
                    const __nv_bfloat16* __restrict__ g1, int ldg1, int act1, const __nv_bfloat16* __restrict__ g2,
                    int ldg2, int act2, float slope, const float* __restrict__ mask, int ppi,
                    const float* __restrict__ gamma, float* __restrict__ sums, __nv_bfloat16* __restrict__ dx, int lddx) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ uint4 small_vals[];
    __shared__ float red[kSmallBwdThreads / 32][16];
    __shared__ float total[16], coef[24];
    uint4* xs = small_vals;
    uint4* g1s = small_vals + m;
    uint4* g2s = small_vals + 2 * m;
    const int ch0 = blockIdx.x * 8;
    const Vec8 sc = loadf8(ss + ch0), sh = loadf8(ss + c + ch0);
    // dz of one pixel from the staged operands (the Dropout2d mask scales g1: d(mask * relu(z)))
    auto dz_of = [&](int pix, const Vec8& v) {
        const Vec8 a = unpack8(g1s[pix]);
        Vec8 mk, b, dz;
#pragma unroll
        for (int i = 0; i < 8; ++i) mk.v[i] = 1.f, b.v[i] = 0.f;
        if (mask != nullptr) mk = loadf8(mask + (size_t)(pix / ppi) * c + ch0);
        if (g2 != nullptr) b = unpack8(g2s[pix]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float z = fmaf(v.v[i], sc.v[i], sh.v[i]);
            // g1 * mask is rounded to bf16 like the in-place scale_channels pass this kernel replaces
            const float ga = mask != nullptr ? __bfloat162float(__float2bfloat16_rn(a.v[i] * mk.v[i])) : a.v[i];
            dz.v[i] = ga * act_grad(z, act1, slope);
            if (g2 != nullptr) dz.v[i] = fmaf(b.v[i], act_grad(z, act2, slope), dz.v[i]);
        }
        return dz;
    };
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll 2
    for (int pix = threadIdx.x; pix < m; pix += kSmallBwdThreads) {
        const uint4 rx = load_raw(x + (size_t)pix * ld + ch0);
        const uint4 ra = load_raw(g1 + (size_t)pix * ldg1 + ch0);
        xs[pix] = rx;
        g1s[pix] = ra;
        if (g2 != nullptr) g2s[pix] = load_raw(g2 + (size_t)pix * ldg2 + ch0);
        const Vec8 v = unpack8(rx);
        const Vec8 dz = dz_of(pix, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            acc[i] += dz.v[i];
            acc[8 + i] = fmaf(dz.v[i], v.v[i], acc[8 + i]);
        }
    }
    block_sum16(acc, red, total);
    if (threadIdx.x < 8) {
        const int ch = ch0 + threadIdx.x;
        const float mu = ss[2 * c + ch], is = ss[3 * c + ch];
        const float s0 = total[threadIdx.x], s1 = is * (total[8 + threadIdx.x] - mu * s0);
        sums[ch] = s0;
        sums[c + ch] = s1;
        const float g = gamma != nullptr ? gamma[ch] : 1.f;
        const float k0 = g * is, kB = -k0 * is * (s1 / (float)m), kA = -k0 * (s0 / (float)m) - kB * mu;
        coef[threadIdx.x] = k0, coef[8 + threadIdx.x] = kA, coef[16 + threadIdx.x] = kB;
    }
    __syncthreads();
    float k0[8], kA[8], kB[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) k0[i] = coef[i], kA[i] = coef[8 + i], kB[i] = coef[16 + i];
    for (int pix = threadIdx.x; pix < m; pix += kSmallBwdThreads) {
        const Vec8 v = unpack8(xs[pix]);
        const Vec8 dz = dz_of(pix, v);
        Vec8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = fmaf(k0[i], dz.v[i], fmaf(kB[i], v.v[i], kA[i]));
        store8(dx + (size_t)pix * lddx + ch0, o);
    }
}

template <int MODE>
static int bn_bwd_dispatch(cudaStream_t st, const __nv_bfloat16* x, long long m, int c, int ld, const float* ss,
                           const __nv_bfloat16* g1, int ldg1, int act1, const __nv_bfloat16* g2, int ldg2, int act2,
                           float slope, float* sums, const float* gamma, __nv_bfloat16* dx, int lddx) {
#define PAI_BWD(A1, A2, BNF)                                                                                        \
    do {                                                                                                             \
        const dim3 g_((unsigned)wave_grid<bn_bwd_kernel<A1, A2, BNF, MODE>>(m, c));                                   \
        if (MODE == 1) /* the reduce modes follow a memset node: ordinary launch */                                  \
            PAI_CUDA_OK(launch_pdl(bn_bwd_kernel<A1, A2, BNF, MODE>, g_, dim3(kEwThreads), 0, st, 1, x, m, c, ld, ss, g1, ldg1, g2, \
                                   ldg2, slope, sums, gamma, dx, lddx));                                               \
        else                                                                                                         \
            bn_bwd_kernel<A1, A2, BNF, MODE><<<g_, kEwThreads, 0, st>>>(x, m, c, ld, ss, g1, ldg1, g2, ldg2, slope, sums, gamma, \
                                                                         dx, lddx);                                   \
    } while (0)
#define PAI_BWD_A2(A1, BNF)                                            \
    do {                                                               \
        if (g2 == nullptr) PAI_BWD(A1, -1, BNF);                       \
        else if (act2 == PAI_ACT_RELU) PAI_BWD(A1, PAI_ACT_RELU, BNF); \
        else if (act2 == PAI_ACT_LEAKY) PAI_BWD(A1, PAI_ACT_LEAKY, BNF); \
        else PAI_BWD(A1, PAI_ACT_NONE, BNF);                           \
    } while (0)
#define PAI_BWD_A1(BNF)                                                \
    do {                                                               \
        if (act1 == PAI_ACT_LEAKY) PAI_BWD_A2(PAI_ACT_LEAKY, BNF);     \
        else if (act1 == PAI_ACT_RELU) PAI_BWD_A2(PAI_ACT_RELU, BNF);  \
        else PAI_BWD_A2(PAI_ACT_NONE, BNF);                            \
    } while (0)
    if (ss != nullptr) {
        if (MODE == 2) {
            set_error("bn_bwd: the fused apply+reduce mode is for layers without BatchNorm");
            return -2;
        }
        PAI_BWD_A1(true);
    } else {
        PAI_BWD_A1(false);
    }
#undef PAI_BWD_A1
#undef PAI_BWD_A2
#undef PAI_BWD
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

// per-channel column sum of a bf16 [m, c] matrix (bias gradients)
__global__ void __launch_bounds__(kEwThreads)
colsum_kernel(const __nv_bfloat16* __restrict__ x, long long m, int c, int ld, float* __restrict__ sums) {
    const int cv = c >> 3;
    const int vec = threadIdx.x % cv;
    const long long ppb = kEwThreads / cv;
    Vec8 s, q;
#pragma unroll
    for (int i = 0; i < 8; ++i) s.v[i] = q.v[i] = 0.f;
    constexpr int U = 4;
    const long long stride = (long long)gridDim.x * ppb;
    long long pix = (long long)blockIdx.x * ppb + threadIdx.x / cv;
    const __nv_bfloat16* xp = x + vec * 8;
    for (; pix + (U - 1) * stride < m; pix += U * stride) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) raw[u] = load_raw(xp + (pix + u * stride) * ld);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const Vec8 v = unpack8(raw[u]);
#pragma unroll
            for (int i = 0; i < 8; ++i) s.v[i] += v.v[i];
        }
    }
    for (; pix < m; pix += stride) {
        const Vec8 v = load8(xp + pix * ld);
#pragma unroll
        for (int i = 0; i < 8; ++i) s.v[i] += v.v[i];
    }
    // reuse the 2-quantity reducer; the second quantity is discarded into sums[c..2c)
    block_reduce_2x8(s, q, cv, c, sums);
}

// wide matrices (Linear bias gradients of the ViT, c up to 3 * 4096): thread <-> 8 columns, rows split over grid.y
__global__ void __launch_bounds__(kEwThreads)
colsum_wide_kernel(const __nv_bfloat16* __restrict__ x, long long m, int c, int ld, float* __restrict__ sums) {
    const int vec = blockIdx.x * kEwThreads + threadIdx.x;
    if (vec * 8 >= c) return;
    Vec8 s;
#pragma unroll
    for (int i = 0; i < 8; ++i) s.v[i] = 0.f;
    for (long long r = blockIdx.y; r < m; r += gridDim.y) {
        const Vec8 v = load8(x + r * ld + vec * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) s.v[i] += v.v[i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(sums + vec * 8 + i, s.v[i]);
}

static bool ew_ok(int c, const void* p, int ld) {
    const int cv = c >> 3;
    return c > 0 && c % 8 == 0 && cv <= kEwThreads && kEwThreads % cv == 0 && ld % 8 == 0 &&
           (reinterpret_cast<uintptr_t>(p) & 15) == 0;
}

}  // namespace pai

using namespace pai;
typedef __nv_bfloat16 bf16;

extern "C" {

int pai_bn_stats(const void* x, long long m, int c, int ld, float* sums, void* stream) {
    PAI_REQUIRE(x && sums && m > 0, "pai_bn_stats: null pointer / empty input");
    PAI_REQUIRE(ew_ok(c, x, ld), "pai_bn_stats: c=%d ld=%d must be multiples of 8 with (c/8) | 256, x 16 B aligned", c, ld);
    cudaStream_t st = (cudaStream_t)stream;
    PAI_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * c, st));
    bn_stats_kernel<<<wave_grid<bn_stats_kernel>(m, c), kEwThreads, 0, st>>>((const bf16*)x, m, c, ld, sums);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

int pai_bn_finalize_partials(const float* sums, int nparts, long long m, int c, const float* gamma, const float* beta,
                             float eps, float momentum, int training, float* running_mean, float* running_var,
                             float* scale_shift, void* stream) {
    PAI_REQUIRE(scale_shift && (training ? (sums != nullptr && nparts >= 1) : (running_mean && running_var)),
                "pai_bn_finalize: null pointer");
    PAI_CUDA_OK(launch_pdl(bn_finalize_kernel, dim3((unsigned)((c + 31) / 32)), dim3(1024), 0, (cudaStream_t)stream, 1, sums, m, c, gamma, beta,
                           eps, momentum, training, running_mean, running_var, scale_shift, nparts));
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}
int pai_bn_finalize(const float* sums, long long m, int c, const float* gamma, const float* beta, float eps,
                    float momentum, int training, float* running_mean, float* running_var, float* scale_shift,
                    void* stream) {
    return pai_bn_finalize_partials(sums, 1, m, c, gamma, beta, eps, momentum, training, running_mean, running_var,
                                    scale_shift, stream);
}

int pai_bn_apply_act(const void* x, long long m, int c, int ld, const float* scale_shift, void* out1, int ld1,
                     int act1, void* out2, int ld2, int act2, float slope, void* stream) {
    PAI_REQUIRE(x && out1 && m > 0, "pai_bn_apply_act: null pointer / empty input");
    PAI_REQUIRE(ew_ok(c, x, ld) && ew_ok(c, out1, ld1) && (out2 == nullptr || ew_ok(c, out2, ld2)),
                "pai_bn_apply_act: bad channel count / stride / alignment (c=%d)", c);
    cudaStream_t st = (cudaStream_t)stream;
#define PAI_APPLY(A1, A2)                                                                                              \
    PAI_CUDA_OK(launch_pdl(bn_apply_act_kernel<A1, A2>, dim3((unsigned)wave_grid<bn_apply_act_kernel<A1, A2>>(m, c)),     \
                           dim3(kEwThreads), 0, st, 1, (const bf16*)x, m, c, ld, scale_shift, (bf16*)out1, ld1, (bf16*)out2, \
                           ld2, slope))
#define PAI_APPLY_A2(A1)                                               \
    do {                                                               \
        if (out2 == nullptr) PAI_APPLY(A1, -1);                        \
        else if (act2 == PAI_ACT_RELU) PAI_APPLY(A1, PAI_ACT_RELU);    \
        else if (act2 == PAI_ACT_LEAKY) PAI_APPLY(A1, PAI_ACT_LEAKY);  \
        else PAI_APPLY(A1, PAI_ACT_NONE);                              \
    } while (0)
    PAI_REQUIRE(act1 != PAI_ACT_TANH && act2 != PAI_ACT_TANH, "pai_bn_apply_act: tanh is not a BatchNorm consumer");
    if (act1 == PAI_ACT_LEAKY) PAI_APPLY_A2(PAI_ACT_LEAKY);
    else if (act1 == PAI_ACT_RELU) PAI_APPLY_A2(PAI_ACT_RELU);
    else PAI_APPLY_A2(PAI_ACT_NONE);
#undef PAI_APPLY_A2
#undef PAI_APPLY
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

int pai_bn_bwd_reduce(const void* x, long long m, int c, int ld, const float* scale_shift, const void* g1, int ldg1,
                      int act1, const void* g2, int ldg2, int act2, float slope, float* sums, void* stream) {
    PAI_REQUIRE(x && g1 && sums && m > 0, "pai_bn_bwd_reduce: null pointer / empty input");
    PAI_REQUIRE(ew_ok(c, x, ld) && ew_ok(c, g1, ldg1) && (g2 == nullptr || ew_ok(c, g2, ldg2)),
                "pai_bn_bwd_reduce: bad channel count / stride / alignment (c=%d)", c);
    cudaStream_t st = (cudaStream_t)stream;
    PAI_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * c, st));
    return bn_bwd_dispatch<0>(st, (const bf16*)x, m, c, ld, scale_shift, (const bf16*)g1, ldg1, act1,
                              (const bf16*)g2, ldg2, act2, slope, sums, nullptr, nullptr, 0);
}

int pai_bn_bwd_apply(const void* x, long long m, int c, int ld, const float* scale_shift, const void* g1, int ldg1,
                     int act1, const void* g2, int ldg2, int act2, float slope, const float* sums,
                     const float* gamma, void* dx, int lddx, void* stream) {
    PAI_REQUIRE(x && g1 && dx && m > 0 && (scale_shift == nullptr || sums != nullptr),
                "pai_bn_bwd_apply: null pointer / empty input");
    PAI_REQUIRE(ew_ok(c, x, ld) && ew_ok(c, g1, ldg1) && (g2 == nullptr || ew_ok(c, g2, ldg2)) && ew_ok(c, dx, lddx),
                "pai_bn_bwd_apply: bad channel count / stride / alignment (c=%d)", c);
    return bn_bwd_dispatch<1>((cudaStream_t)stream, (const bf16*)x, m, c, ld, scale_shift, (const bf16*)g1,
                              ldg1, act1, (const bf16*)g2, ldg2, act2, slope, const_cast<float*>(sums), gamma, (bf16*)dx,
                              lddx);
}

int pai_act_bwd(const void* x, long long m, int c, int ld, const void* g1, int ldg1, int act1, const void* g2, int ldg2,
                int act2, float slope, float* sums, void* dx, int lddx, void* stream) {
    PAI_REQUIRE(x && g1 && dx && sums && m > 0, "pai_act_bwd: null pointer / empty input");
    PAI_REQUIRE(ew_ok(c, x, ld) && ew_ok(c, g1, ldg1) && (g2 == nullptr || ew_ok(c, g2, ldg2)) && ew_ok(c, dx, lddx),
                "pai_act_bwd: bad channel count / stride / alignment (c=%d)", c);
    cudaStream_t st = (cudaStream_t)stream;
    PAI_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * c, st));
    return bn_bwd_dispatch<2>(st, (const bf16*)x, m, c, ld, nullptr, (const bf16*)g1, ldg1, act1,
                              (const bf16*)g2, ldg2, act2, slope, sums, nullptr, (bf16*)dx, lddx);
}

static const size_t kSmallSmemMax = 200 * 1024;

int pai_bn_small_ok(long long m, int c, int operands) {
    return m > 0 && c > 0 && c % 8 == 0 && operands >= 1 && operands <= 3 && (size_t)m * 16 * operands <= kSmallSmemMax;
}

int pai_bn_small_fwd(const void* x, long long m, int c, int ld, const float* gamma, const float* beta, float eps,
                     float momentum, float* running_mean, float* running_var, float* scale_shift, void* out1, int ld1,
                     int act1, void* out2, int ld2, int act2, float slope, const float* mask, int pixels_per_image,
                     void* stream) {
    PAI_REQUIRE(x && scale_shift && out1 && m > 0, "pai_bn_small_fwd: null pointer / empty input");
    PAI_REQUIRE(pai_bn_small_ok(m, c, 1), "pai_bn_small_fwd: m=%lld c=%d does not fit one block per 8 channels", m, c);
    PAI_REQUIRE(ew_ok(8, x, ld) && ew_ok(8, out1, ld1) && (out2 == nullptr || ew_ok(8, out2, ld2)),
                "pai_bn_small_fwd: bad stride / alignment (c=%d)", c);
    PAI_REQUIRE(act1 != PAI_ACT_TANH && act2 != PAI_ACT_TANH, "pai_bn_small_fwd: tanh is not a BatchNorm consumer");
    PAI_REQUIRE(mask == nullptr || pixels_per_image > 0, "pai_bn_small_fwd: mask needs pixels_per_image");
    static DeviceOnce once;
    const int dev = current_device();
    if (dev < 0) return -1;
    if (once.need(dev)) {
        PAI_CUDA_OK(cudaFuncSetAttribute(bn_small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSmemMax));
        once.mark(dev);
    }
    PAI_CUDA_OK(launch_pdl(bn_small_fwd_kernel, dim3((unsigned)(c / 8)), dim3(kSmallThreads), (size_t)m * 16, (cudaStream_t)stream, 1,
                           (const bf16*)x, (int)m, c, ld, gamma, beta, eps, momentum, running_mean, running_var, scale_shift,
                           (bf16*)out1, ld1, act1, (bf16*)out2, ld2, act2, slope, mask, pixels_per_image > 0 ? pixels_per_image : 1));
    return 0;
}

int pai_bn_small_bwd(const void* x, long long m, int c, int ld, const float* scale_shift, const void* g1, int ldg1,
                     int act1, const void* g2, int ldg2, int act2, float slope, const float* mask, int pixels_per_image,
                     const float* gamma, float* sums, void* dx, int lddx, void* stream) {
    PAI_REQUIRE(x && scale_shift && g1 && sums && dx && m > 0, "pai_bn_small_bwd: null pointer / empty input");
    const int operands = g2 != nullptr ? 3 : 2;
    PAI_REQUIRE(pai_bn_small_ok(m, c, operands), "pai_bn_small_bwd: m=%lld c=%d does not fit one block per 8 channels", m, c);
    PAI_REQUIRE(ew_ok(8, x, ld) && ew_ok(8, g1, ldg1) && (g2 == nullptr || ew_ok(8, g2, ldg2)) && ew_ok(8, dx, lddx),
                "pai_bn_small_bwd: bad stride / alignment (c=%d)", c);
    PAI_REQUIRE(mask == nullptr || pixels_per_image > 0, "pai_bn_small_bwd: mask needs pixels_per_image");
    static DeviceOnce once;
    const int dev = current_device();
    if (dev < 0) return -1;
    if (once.need(dev)) {
        PAI_CUDA_OK(cudaFuncSetAttribute(bn_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSmemMax));
        once.mark(dev);
    }
    PAI_CUDA_OK(launch_pdl(bn_small_bwd_kernel, dim3((unsigned)(c / 8)), dim3(kSmallBwdThreads), (size_t)m * 16 * operands,
                           (cudaStream_t)stream, 1, (const bf16*)x, (int)m, c, ld, scale_shift, (const bf16*)g1, ldg1, act1,
                           (const bf16*)g2, ldg2, act2, slope, mask, pixels_per_image > 0 ? pixels_per_image : 1, gamma, sums,
                           (bf16*)dx, lddx));
    return 0;
}

int pai_colsum(const void* x, long long m, int c, int ld, float* sums2c, void* stream) {
    PAI_REQUIRE(x && sums2c && m > 0, "pai_colsum: null pointer / empty input");
    cudaStream_t st = (cudaStream_t)stream;
    if (!ew_ok(c, x, ld)) {
        PAI_REQUIRE(c > 0 && c % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
                    "pai_colsum: bad channel count / stride / alignment (c=%d)", c);
        PAI_CUDA_OK(cudaMemsetAsync(sums2c, 0, sizeof(float) * 2 * c, st));
        const int ry = (int)(m < 64 ? m : 64);
        colsum_wide_kernel<<<dim3((c / 8 + kEwThreads - 1) / kEwThreads, ry), kEwThreads, 0, st>>>((const bf16*)x, m, c, ld, sums2c);
        PAI_CUDA_OK(cudaGetLastError());
        return 0;
    }
    PAI_CUDA_OK(cudaMemsetAsync(sums2c, 0, sizeof(float) * 2 * c, st));
    colsum_kernel<<<wave_grid<colsum_kernel>(m, c), kEwThreads, 0, st>>>((const bf16*)x, m, c, ld, sums2c);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// report.py image outputs (report.py:122-141,220-233)
namespace pai {
__device__ __forceinline__ unsigned char to_u8(float x) {
    // torchvision F_t.convert_image_dtype float -> uint8: x * (255 + 1 - 1e-3), truncated; out-of-range values wrap like
    // the host's float -> int32 -> uint8 cast
    return (unsigned char)((int)(x * 255.999f) & 0xff);
}
__global__ void __launch_bounds__(256) to_uint8_kernel(const float* __restrict__ x, long long count, unsigned char* __restrict__ out) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < count; i += (long long)gridDim.x * 256) out[i] = to_u8(x[i]);
}
__global__ void __launch_bounds__(256) afmhot_u8_kernel(const float* __restrict__ img, int n, long long hw, unsigned char* __restrict__ out) {
    const long long total = (long long)n * hw;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long b = i / hw, pix = i - b * hw;
        int idx = (int)(img[i] * 256.f);
        idx = idx < 0 ? 0 : (idx > 255 ? 255 : idx);
        const float v = (float)idx / 255.f;                      // look-up table entry idx of 256
        unsigned char* o = out + b * 3 * hw + pix;
        o[0] = to_u8(fminf(fmaxf(2.f * v, 0.f), 1.f));
        o[hw] = to_u8(fminf(fmaxf(2.f * v - 0.5f, 0.f), 1.f));
        o[2 * hw] = to_u8(fminf(fmaxf(2.f * v - 1.f, 0.f), 1.f));
    }
}
}  // namespace pai

extern "C" {
int pai_to_uint8(const float* x, long long count, unsigned char* out, void* stream) {
    PAI_REQUIRE(x && out && count >= 0, "pai_to_uint8: bad arguments");
    if (count == 0) return 0;
    long long blocks = (count + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pai::to_uint8_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, count, out);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}
int pai_afmhot_u8(const float* img, int n, long long hw, unsigned char* out, void* stream) {
    PAI_REQUIRE(img && out && n >= 0 && hw > 0, "pai_afmhot_u8: bad arguments");
    if (n == 0) return 0;
    long long blocks = ((long long)n * hw + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pai::afmhot_u8_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(img, n, hw, out);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}
}  // extern "C"

// ---------------------------------------------------------------------------------------------
// dataset.py's per-image transform on the device (dataset.py:51-61,126-134): Resize((oh, ow), antialias=True) of a
// uint8 grayscale image -> round back to uint8 levels -> ConvertImageDtype(float32) (/255) -> Normalize(0.5, 0.5).
// The resampling is ATen's _upsample_bilinear2d_aa: triangle filter of half-width max(scale, 1) around the source
// centre (o + 0.5) * scale, taps clipped to the image and renormalised.
namespace pai {
struct AaTaps {
    int lo, n;
    float center, invscale;
};
__device__ __forceinline__ AaTaps aa_taps(int o, int in, int out) {
    const float scale = (float)in / (float)out;
    const float support = scale >= 1.f ? scale : 1.f;
    AaTaps t;
    t.center = scale * ((float)o + 0.5f);
    t.invscale = scale >= 1.f ? 1.f / scale : 1.f;
    t.lo = max((int)(t.center - support + 0.5f), 0);
    t.n = min((int)(t.center + support + 0.5f), in) - t.lo;
    return t;
}
__device__ __forceinline__ float aa_weight(const AaTaps& t, int j) {
    const float x = fabsf(((float)(j + t.lo) - t.center + 0.5f) * t.invscale);
    return x < 1.f ? 1.f - x : 0.f;
}
__global__ void __launch_bounds__(256)
resize_aa_norm_kernel(const unsigned char* __restrict__ img, int n, int ih, int iw, int oh, int ow, int normalize,
                      int round_u8, float* __restrict__ out) {
    const long long total = (long long)n * oh * ow;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int ox = (int)(i % ow), oy = (int)((i / ow) % oh);
        const long long b = i / ((long long)ow * oh);
        const AaTaps ty = aa_taps(oy, ih, oh), tx = aa_taps(ox, iw, ow);
        float wy_sum = 0.f, wx_sum = 0.f;
        for (int j = 0; j < ty.n; ++j) wy_sum += aa_weight(ty, j);
        for (int j = 0; j < tx.n; ++j) wx_sum += aa_weight(tx, j);
        const unsigned char* src = img + b * ih * iw;
        // horizontal pass first, then vertical (the order ATen's separable implementation uses); each intermediate row
        // value is kept in fp32
        float acc = 0.f;
        for (int jy = 0; jy < ty.n; ++jy) {
            const unsigned char* row = src + (long long)(ty.lo + jy) * iw + tx.lo;
            float r = 0.f;
            for (int jx = 0; jx < tx.n; ++jx) r = fmaf(aa_weight(tx, jx) / wx_sum, (float)row[jx], r);
            acc = fmaf(aa_weight(ty, jy) / wy_sum, r, acc);
        }
        if (round_u8) acc = fminf(fmaxf(rintf(acc), 0.f), 255.f);      // torchvision rounds back to uint8 after resizing
        float v = acc / 255.f;                                       // ConvertImageDtype(torch.float32)
        if (normalize) v = (v - 0.5f) / 0.5f;                        // Normalize(0.5, 0.5) for the single channel
        out[i] = v;
    }
}
}  // namespace pai

extern "C" int pai_resize_aa_normalize_u8(const unsigned char* img, int n, int ih, int iw, int oh, int ow, int normalize,
                                          int round_u8, float* out, void* stream) {
    PAI_REQUIRE(img && out && n >= 0 && ih > 0 && iw > 0 && oh > 0 && ow > 0, "pai_resize_aa_normalize_u8: bad arguments");
    if (n == 0) return 0;
    long long blocks = ((long long)n * oh * ow + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pai::resize_aa_norm_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(img, n, ih, iw, oh, ow, normalize, round_u8, out);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}
