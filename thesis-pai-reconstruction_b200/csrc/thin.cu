// The degenerate (1- or 2-channel wide) layers of the hot path as single-pass tcgen05 kernels -- no im2col / col2im
// carrier in HBM.  They are HBM-bound (the wide NHWC tensor is read or written exactly once); the contraction itself is
// a K = 16 / 32 (taps x planes) or N = 16 (taps) tensor-core GEMM whose thin operand is built in shared memory by the
// CTA's own threads in the SWIZZLE_128B layout the UMMA descriptors read.
//
//   enc0  Conv2d(1, 64, 4, 2, 1)            models/pix2pix.py:141-147       thin_conv_fprop   (two fused outputs)
//   D0    Conv2d(2, 64, 4, 2, 1)+LeakyReLU  models/wrapper.py:196-206,229   thin_conv_fprop   (cat([x, y]) never built)
//   dec7  ConvTranspose2d(128, 1, 4, 2, 1)  models/pix2pix.py:186-192       thin_convT_plane  (+ bias + Tanh :195)
//   their gradients: dec7 dgrad == thin_conv_fprop on the gradient plane, D0 dgrad == thin_convT_plane,
//   the three weight gradients == thin_wgrad.
//
// Index math: SURVEY.md Appendix B (stride-2 taps  in = 2*o - 1 + k;  ConvT phases T[0] = {(k=1,d=0),(k=3,d=-1)},
// T[1] = {(k=0,d=+1),(k=2,d=0)}).
#include <stdlib.h>

#include "pai_common.cuh"
#include "pai_epilogue.cuh"
#include "pai_kernels.h"

namespace pai {

#ifdef PAI_PROFILE_ROLES
#define TROLE_T0() const long long _t0 = clock64()
#define TROLE_ADD(var) var += clock64() - _t0
#else
#define TROLE_T0()
#define TROLE_ADD(var)
#endif

// =============================================================================================
// thin_conv_fprop: out[pix, co] = act(bias[co] + sum_{t, j} plane_j[n, 2*oy-1+ky, 2*ox-1+kx] * W[co][t*CIN + j])
//
// Tile = 128 consecutive output pixels, two slots (tile parity).  Warps 0-3 build the [128 x 16*CIN] bf16 im2col rows of
// a tile straight from the fp32 plane(s) into the slot's shared-memory A tile (the loads of tile i+1 are in flight while
// tile i is converted and stored), warp 4 issues the 1-2 MMAs (M = 128, N = cout, K = 16) into the slot's TMEM
// accumulator, epilogue warps 8-11 / 12-15 drain accumulator 0 / 1 through the coalescing transpose tile into one or two
// outputs.  The weight tile [cout x 64] stays resident.
static constexpr int kThinFpropThreads = 16 * 32;   // 4 warps per SM sub-partition: 128 registers per thread

struct ThinFpropParams {
    const float* p0;
    const float* p1;
    int n, ih, iw, oh, ow;
    const __nv_bfloat16* w;      // [cout][64] bf16, column = tap * CIN + j (zero padded)
    int cout;                    // multiple of 64, <= 256
    const float* bias;
    __nv_bfloat16* out1;
    int ld1, act1;
    __nv_bfloat16* out2;
    int ld2, act2;
    float slope;
    long long total_pix;
    int tiles;
    FastDiv fd_ow, fd_oh;
};

struct __align__(8) ThinPipe {
    uint64_t a_full[2], a_empty[2], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

// The warp's 32 rows are 32 consecutive pixels pix0 .. pix0+31 of a tensor with pixel stride ld: the row addresses
// follow from the (warp-uniform) first pixel, no shuffles.
template <int ACT>
__device__ __forceinline__ void thin_store_chunk(const uint32_t (&v)[64], float slope, uint4* tile, int lane,
                                                 __nv_bfloat16* dst, long long pix0, int ld, long long total_pix) {
    bias_act_pack<ACT>(v, nullptr, slope, tile, lane);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = i * 4 + (lane >> 3), ch = lane & 7;
        const uint4 val = tile[row * 8 + (ch ^ (row & 7))];
        const long long pix = pix0 + row;
        if (pix < total_pix) *reinterpret_cast<uint4*>(dst + pix * ld + ch * 8) = val;
    }
    __syncwarp();
}

__device__ __forceinline__ void thin_store_dispatch(const uint32_t (&v)[64], int act, float slope, uint4* tile, int lane,
                                                    __nv_bfloat16* dst, long long pix0, int ld, long long total_pix) {
    if (act == PAI_ACT_LEAKY)
        thin_store_chunk<PAI_ACT_LEAKY>(v, slope, tile, lane, dst, pix0, ld, total_pix);
    else if (act == PAI_ACT_RELU)
        thin_store_chunk<PAI_ACT_RELU>(v, slope, tile, lane, dst, pix0, ld, total_pix);
    else if (act == PAI_ACT_TANH)
        thin_store_chunk<PAI_ACT_TANH>(v, slope, tile, lane, dst, pix0, ld, total_pix);
    else
        thin_store_chunk<PAI_ACT_NONE>(v, slope, tile, lane, dst, pix0, ld, total_pix);
}

//
// STREAM (256-pixel wide images: a tile is one output row): the CTA owns a contiguous range of output rows and the input
// rows arrive through a 16-slot shared-memory ring of two-row granules filled by bulk async copies (one thread, many KB
// in flight), so the producers read shared memory instead of waiting a DRAM round trip per tile and every input row is
// fetched once per CTA instead of twice.  Granule g holds input rows 2g+1 and 2g+2; output row oy needs granules oy-1, oy.
static constexpr int kThinRing = 16;
static constexpr int kThinRowW = 256;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int CIN, bool STREAM>
__global__ void __launch_bounds__(kThinFpropThreads, 1) thin_conv_fprop_kernel(const ThinFpropParams p) {
    pdl_launch_dependents();
    pdl_wait();              // (the set-up below already reads the weight pack written by the optimizer kernels)
    extern __shared__ uint8_t smem_raw[];
    __shared__ ThinPipe ps;
    __shared__ uint64_t ring_full[kThinRing], ring_empty[kThinRing];
    __shared__ uint4 stage_buf[8][32 * 8];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                       // 2 x [128 rows x 128 B]
    uint8_t* sB = smem + 2 * 16384;           // [cout rows x 128 B]
    float* ring = reinterpret_cast<float*>(sB + (size_t)p.cout * 128);   // STREAM: [slot][plane][2 rows][256] fp32
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t acc_cols = tmem_cols_for(p.cout);
    // tiles of this CTA: STREAM -> the output rows [t_begin, t_begin + t_count), else strided over the grid
    const long long t_begin = STREAM ? (long long)p.tiles * blockIdx.x / gridDim.x : (long long)blockIdx.x;
    const long long t_step = STREAM ? 1 : (long long)gridDim.x;
    const int t_count = STREAM ? (int)((long long)p.tiles * (blockIdx.x + 1) / gridDim.x - t_begin)
                               : (int)(((long long)p.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

    // resident weight tile (generic-proxy writes, made visible to the tensor core by the proxy fence below)
    for (int i = threadIdx.x; i < p.cout * 8; i += kThinFpropThreads) {
        const int row = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sB + sw128_off(row, c)) = __ldg(reinterpret_cast<const uint4*>(p.w) + i);
    }
    // The bias rides on the GEMM: K columns 16*CIN, 16*CIN+1 of every A row hold 1.0 and the same columns of the weight
    // tile hold the bias split into two bf16 terms (hi + lo: 16 mantissa bits, error < 1e-5 relative) -- one more MMA
    // k-step per tile instead of 64 loads + adds per row in the epilogue.  The constant A chunks are written once.
    if (p.bias != nullptr) {
        __syncthreads();
        for (int co = threadIdx.x; co < p.cout; co += kThinFpropThreads) {
            const float b = __ldg(p.bias + co);
            const __nv_bfloat16 hi = __float2bfloat16_rn(b);
            const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
            const uint32_t w0 = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
            *reinterpret_cast<uint4*>(sB + sw128_off(co, 2 * CIN)) = make_uint4(w0, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(sB + sw128_off(co, 2 * CIN + 1)) = make_uint4(0u, 0u, 0u, 0u);
        }
        for (int i = threadIdx.x; i < 2 * 128; i += kThinFpropThreads) {
            uint8_t* a_tile = sA + (i >> 7) * 16384;
            const int row = i & 127;
            *reinterpret_cast<uint4*>(a_tile + sw128_off(row, 2 * CIN)) = make_uint4(0x3F803F80u, 0u, 0u, 0u);   // (1.0, 1.0)
            *reinterpret_cast<uint4*>(a_tile + sw128_off(row, 2 * CIN + 1)) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    fence_proxy_async_smem();
    if (warp == 4 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            // one arrival per WARP (lane 0 after __syncwarp): 128 per-thread arrivals on one mbarrier serialise
            mbar_init(&ps.a_full[s], 4);
            mbar_init(&ps.a_empty[s], 1);
            mbar_init(&ps.acc_full[s], 1);
            mbar_init(&ps.acc_empty[s], 4);
        }
        if (STREAM)
            for (int s = 0; s < kThinRing; ++s) {
                mbar_init(&ring_full[s], CIN);           // one arrival (with its byte count) per plane loader
                mbar_init(&ring_empty[s], 4);
            }
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(&ps.tmem_base, 2 * acc_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ps.tmem_base;

    if (warp < 4 && STREAM) {
        // ---------------- producers (streaming): thread r <-> output column r of the CTA's current output row
        const int r = warp * 32 + lane;
        const uint32_t ring_s = smem_u32(ring);
        const uint32_t col_b = (uint32_t)(2 * r) * 4;                 // byte offset of input column 2r inside a row
        const uint32_t lft_b = r > 0 ? 4u : 0u, rgt_b = r < 127 ? 8u : 0u;   // neighbours 2r-1 / 2r+2 (clamped at the image edge)
        int gc = 0;                                   // granule counter of this CTA (same sequence as the loader's)
        int i = 0;
        long long row = t_begin;
        const long long row_end = t_begin + t_count;
        long long w_ring = 0, w_slot = 0;
        (void)w_ring, (void)w_slot;
#ifdef PAI_PROFILE_ROLES
        const long long tstart = clock64();
#endif
        while (row < row_end) {
            const int img = (int)(row / p.oh);
            const int a_lo = (int)(row - (long long)img * p.oh);
            const long long img_end = (long long)(img + 1) * p.oh;
            const int a_hi = (int)((img_end < row_end ? img_end : row_end) - (long long)img * p.oh);
            mbar_wait(&ring_full[gc % kThinRing], (uint32_t)((gc / kThinRing) & 1));          // granule a_lo - 1
            for (int a = a_lo; a < a_hi; ++a, ++i, ++gc) {
                const int g1 = gc + 1;                                                       // granule a
                {
                    TROLE_T0();
                    mbar_wait(&ring_full[g1 % kThinRing], (uint32_t)((g1 / kThinRing) & 1));
                    TROLE_ADD(w_ring);
                }
                uint32_t words[8 * CIN];
                float v[CIN][16];
                // shared-space loads with immediate offsets: one base register per granule, the plane / row-of-granule
                // offsets are compile-time constants (this warp is alone on its scheduler: the loop is a chain of dependent
                // instructions, so every address computation it does not need counts)
                const uint32_t base_lo = ring_s + (uint32_t)((gc % kThinRing) * CIN) * (2 * kThinRowW * 4) + col_b;
                const uint32_t base_hi = ring_s + (uint32_t)((g1 % kThinRing) * CIN) * (2 * kThinRowW * 4) + col_b;
                const bool top_ok = a > 0, bot_ok = 2 * a + 2 < p.ih;       // input rows 2a-1 / 2a+2 inside the image
#pragma unroll
                for (int j = 0; j < CIN; ++j) {
#pragma unroll
                    for (int ky = 0; ky < 4; ++ky) {
                        const uint32_t addr = (ky < 2 ? base_lo : base_hi) + (uint32_t)(j * 2 + (ky & 1)) * (kThinRowW * 4);
                        float m0, m1, lf, rg;
                        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(m0), "=f"(m1) : "r"(addr));
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(lf) : "r"(addr - lft_b));
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(rg) : "r"(addr + rgt_b));
                        const bool rok = ky == 0 ? top_ok : (ky == 3 ? bot_ok : true);
                        v[j][ky * 4 + 0] = (rok && r > 0) ? lf : 0.f;
                        v[j][ky * 4 + 1] = rok ? m0 : 0.f;
                        v[j][ky * 4 + 2] = rok ? m1 : 0.f;
                        v[j][ky * 4 + 3] = (rok && r < 127) ? rg : 0.f;
                    }
                }
                if (CIN == 1) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        __nv_bfloat162 h = __floats2bfloat162_rn(v[0][2 * q], v[0][2 * q + 1]);
                        words[q] = *reinterpret_cast<uint32_t*>(&h);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 16; ++q) {           // word q = (plane 0, plane 1) of tap q
                        __nv_bfloat162 h = __floats2bfloat162_rn(v[0][q], v[CIN - 1][q]);
                        words[q] = *reinterpret_cast<uint32_t*>(&h);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ring_empty[gc % kThinRing]);   // granule a-1 is not needed after this tile
                const int s = i & 1;
                {
                    TROLE_T0();
                    mbar_wait(&ps.a_empty[s], (uint32_t)(((i >> 1) & 1) ^ 1));
                    TROLE_ADD(w_slot);
                }
                uint8_t* a_tile = sA + s * 16384;
#pragma unroll
                for (int c = 0; c < 2 * CIN; ++c)
                    *reinterpret_cast<uint4*>(a_tile + sw128_off(r, c)) =
                        make_uint4(words[4 * c], words[4 * c + 1], words[4 * c + 2], words[4 * c + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ps.a_full[s]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring_empty[gc % kThinRing]);       // last granule of the segment
            ++gc;
            row = (long long)img * p.oh + a_hi;
        }
#ifdef PAI_PROFILE_ROLES
        if (blockIdx.x == 0 && threadIdx.x == 0)
            printf("[thin roles] producer: total %lld cyc, waiting for input rows %lld, for a free A slot %lld, tiles %d\n",
                   clock64() - tstart, w_ring, w_slot, i);
#endif
    } else if ((warp == 5 || (CIN == 2 && warp == 6)) && STREAM) {
        // ---------------- loaders: bulk async copies of the input rows into the granule ring, one thread per plane (an
        // elected thread needs ~250 cycles per cp.async.bulk: with one loader and four 1 KB copies per granule the
        // 2-plane layer was loader-bound).  The two rows of a granule are adjacent in memory: one 2 KB copy when both exist.
        if (elect_one()) {
            const int j = warp - 5;                                  // plane of this loader
            const float* plane = j == 0 ? p.p0 : p.p1;
            int gc = 0;
            long long row = t_begin;
            const long long row_end = t_begin + t_count;
            while (row < row_end) {
                const int img = (int)(row / p.oh);
                const int a_lo = (int)(row - (long long)img * p.oh);
                const long long img_end = (long long)(img + 1) * p.oh;
                const int a_hi = (int)((img_end < row_end ? img_end : row_end) - (long long)img * p.oh);
                const float* src = plane + (size_t)img * p.ih * p.iw;
                for (int g = a_lo - 1; g < a_hi; ++g, ++gc) {
                    const int sl = gc % kThinRing;
                    mbar_wait(&ring_empty[sl], (uint32_t)(((gc / kThinRing) & 1) ^ 1));
                    const int r0 = 2 * g + 1, r1 = 2 * g + 2;
                    const bool ok0 = r0 >= 0, ok1 = r1 < p.ih;
                    float* dst = ring + (size_t)(sl * CIN + j) * 2 * kThinRowW;
                    if (ok0 && ok1) {
                        mbar_expect_tx(&ring_full[sl], 2 * kThinRowW * 4);
                        bulk_g2s(dst, src + (size_t)r0 * p.iw, 2 * kThinRowW * 4, &ring_full[sl]);
                    } else {
                        mbar_expect_tx(&ring_full[sl], kThinRowW * 4);
                        if (ok0) bulk_g2s(dst, src + (size_t)r0 * p.iw, kThinRowW * 4, &ring_full[sl]);
                        if (ok1) bulk_g2s(dst + kThinRowW, src + (size_t)r1 * p.iw, kThinRowW * 4, &ring_full[sl]);
                    }
                }
                row = (long long)img * p.oh + a_hi;
            }
        }
    } else if (warp < 4) {
        // ---------------- producers (gather): one im2col row (= output pixel) per thread, loads straight from global
        const int r = warp * 32 + lane;
        float v[CIN][16];
        auto load_tile = [&](long long tile) {
            const long long pix = tile * 128 + r;
            const bool ok = pix < p.total_pix;
            int t, ox, n, oy;
            p.fd_ow.divmod((int)(ok ? pix : 0), t, ox);
            p.fd_oh.divmod(t, n, oy);
#pragma unroll
            for (int j = 0; j < CIN; ++j) {
                const float* src = (j == 0 ? p.p0 : p.p1) + (size_t)n * p.ih * p.iw;
#pragma unroll
                for (int ky = 0; ky < 4; ++ky) {
                    const int iy = 2 * oy - 1 + ky;
                    const bool rok = ok && iy >= 0 && iy < p.ih;
                    const float* rp = src + (size_t)(rok ? iy : 0) * p.iw;
#pragma unroll
                    for (int kx = 0; kx < 4; ++kx) {
                        const int ix = 2 * ox - 1 + kx;
                        v[j][ky * 4 + kx] = (rok && ix >= 0 && ix < p.iw) ? __ldg(rp + ix) : 0.f;
                    }
                }
            }
        };
        if (t_count > 0) load_tile(t_begin);
        for (int i = 0; i < t_count; ++i) {
            const int s = i & 1;
            uint32_t words[8 * CIN];
            if (CIN == 1) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(v[0][2 * q], v[0][2 * q + 1]);
                    words[q] = *reinterpret_cast<uint32_t*>(&h);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(v[0][q], v[CIN - 1][q]);
                    words[q] = *reinterpret_cast<uint32_t*>(&h);
                }
            }
            if (i + 1 < t_count) load_tile(t_begin + (long long)(i + 1) * t_step);   // in flight during the wait / stores below
            mbar_wait(&ps.a_empty[s], (uint32_t)(((i >> 1) & 1) ^ 1));   // the MMAs that read this slot have completed
            uint8_t* a_tile = sA + s * 16384;
#pragma unroll
            for (int c = 0; c < 2 * CIN; ++c)
                *reinterpret_cast<uint4*>(a_tile + sw128_off(r, c)) =
                    make_uint4(words[4 * c], words[4 * c + 1], words[4 * c + 2], words[4 * c + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ps.a_full[s]);
        }
    } else if (warp == 4) {
        // ---------------- MMA issuer
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(128, p.cout, 0, 0);
            const uint64_t db0 = umma_desc_kmajor_sw128(smem_u32(sB));
            long long w_acc = 0, w_a = 0;
            (void)w_acc, (void)w_a;
#ifdef PAI_PROFILE_ROLES
            const long long tstart = clock64();
#endif
            for (int i = 0; i < t_count; ++i) {
                const int s = i & 1;
                const uint32_t ph = (uint32_t)((i >> 1) & 1);
                {
                    TROLE_T0();
                    mbar_wait(&ps.acc_empty[s], ph ^ 1);
                    TROLE_ADD(w_acc);
                }
                {
                    TROLE_T0();
                    mbar_wait(&ps.a_full[s], ph);
                    TROLE_ADD(w_a);
                }
                tc_fence_after();
                const uint64_t da0 = umma_desc_kmajor_sw128(smem_u32(sA + s * 16384));
                const uint32_t td = tmem_base + s * acc_cols;
                umma_bf16_ss(td, da0, db0, idesc, 0);
                if (CIN == 2) umma_bf16_acc(td, da0 + 2, db0 + 2, idesc);
                if (p.bias != nullptr) umma_bf16_acc(td, da0 + 2 * CIN, db0 + 2 * CIN, idesc);     // the bias k-step
                umma_commit(&ps.a_empty[s]);
                umma_commit(&ps.acc_full[s]);
            }
#ifdef PAI_PROFILE_ROLES
            if (blockIdx.x == 0)
                printf("[thin roles] mma: total %lld cyc, waiting for a drained accumulator %lld, for a built A tile %lld\n",
                       clock64() - tstart, w_acc, w_a);
#endif
        }
    } else if (warp >= 8) {
        // ---------------- epilogue: group g drains accumulator g (tiles of parity g)
        const int q = warp & 3;
        const int g = (warp - 8) >> 2;
        const int r = q * 32 + lane;
        uint4* tile_buf = stage_buf[warp - 8];
        int k = 0;
        long long ep_wait = 0, ep_work = 0;
        (void)ep_wait, (void)ep_work;
        for (int i = g; i < t_count; i += 2, ++k) {
            const long long tile = t_begin + (long long)i * t_step;
            {
                TROLE_T0();
                mbar_wait(&ps.acc_full[g], (uint32_t)(k & 1));
                TROLE_ADD(ep_wait);
            }
#ifdef PAI_PROFILE_ROLES
            const long long tw0 = clock64();
#endif
            tc_fence_after();
            const uint32_t td = tmem_base + g * acc_cols + ((uint32_t)(q * 32) << 16);
            for (int c0 = 0; c0 < p.cout; c0 += 64) {
                uint32_t v[64];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    tmem_ld_16(td + (uint32_t)(c0 + 16 * j), *reinterpret_cast<uint32_t(*)[16]>(&v[16 * j]));
                tmem_ld_wait();
                if (c0 + 64 >= p.cout) {
                    // last chunk in registers: the accumulator goes back to the MMA warp now, with a relaxed arrival (the
                    // default release form would wait for the stores below)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_relaxed(&ps.acc_empty[g]);
                }
                const long long pix0 = tile * 128 + q * 32;          // first pixel of this warp's 32 rows
                thin_store_dispatch(v, p.act1, p.slope, tile_buf, lane, p.out1 + c0, pix0, p.ld1, p.total_pix);
                if (p.out2 != nullptr)
                    thin_store_dispatch(v, p.act2, p.slope, tile_buf, lane, p.out2 + c0, pix0, p.ld2, p.total_pix);
            }
#ifdef PAI_PROFILE_ROLES
            ep_work += clock64() - tw0;
#endif
        }
#ifdef PAI_PROFILE_ROLES
        if (blockIdx.x == 0 && lane == 0 && (warp == 8 || warp == 12))
            printf("[thin roles] epilogue warp %d: waiting for a full accumulator %lld cyc, working %lld cyc over %d tiles\n", warp,
                   ep_wait, ep_work, k);
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * acc_cols);
    }
}

// =============================================================================================
// thin_wgrad: dw[c][t*CIN + j] += sum_pix U[pix, c] * plane_j[n, 2*oy-1+ky, 2*ox-1+kx]
//
// GEMM view: M = channels of the wide tensor U (MN-major straight from NHWC through TMA, like igemm_wgrad), N = 16*CIN
// tap columns built K-major in shared memory by 4 producer warps, K = pixels in blocks of 64 consecutive pixels of one
// output row.  Every CTA owns a contiguous range of pixel blocks and ONE TMEM accumulator; it flushes once at the end
// with vector reductions.
static constexpr int kThinWgradThreads = 256;
static constexpr int kThinWgradStages = 6;

struct ThinWgradParams {
    const float* p0;
    const float* p1;
    int n, ih, iw, oh, ow;
    int c;                       // channels of U (64 or 128)
    long long blocks;            // n * oh * ow / 64
    int blocks_per_row;          // ow / 64
    float* dw;                   // [c][16*CIN] fp32, accumulated
    FastDiv fd_bpr, fd_oh;
};

struct __align__(8) ThinWgradPipe {
    uint64_t full[kThinWgradStages], empty[kThinWgradStages], acc_full;
    uint32_t tmem_base;
};

template <int CIN>
__global__ void __launch_bounds__(kThinWgradThreads, 1)
thin_wgrad_kernel(const __grid_constant__ CUtensorMap tm_u, const ThinWgradParams p) {
    pdl_launch_dependents();
    constexpr int N = 16 * CIN;
    constexpr uint32_t kABytes = 2 * 8192, kBBytes = 4096, kStage = kABytes + kBBytes;
    extern __shared__ uint8_t smem_raw[];
    __shared__ ThinWgradPipe ps;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long b_begin = p.blocks * blockIdx.x / gridDim.x, b_end = p.blocks * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_u);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kThinWgradStages; ++s) {
            mbar_init(&ps.full[s], 5);        // 4 column-producer warps (lane 0 after __syncwarp) + the TMA thread's expect_tx arrival
            mbar_init(&ps.empty[s], 1);
        }
        mbar_init(&ps.acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(&ps.tmem_base, 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ps.tmem_base;
    pdl_wait();

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long b = b_begin; b < b_end; ++b) {
                mbar_wait(&ps.empty[stage], phase ^ 1);
                uint8_t* sa = smem + (size_t)stage * kStage;
                mbar_expect_tx(&ps.full[stage], kABytes);
                // two 64-channel blocks of the same 64 pixels; channels >= c are TMA zero fill (c == 64)
                tma_load_2d(sa, &tm_u, &ps.full[stage], 0, (int)(b * 64));
                tma_load_2d(sa + 8192, &tm_u, &ps.full[stage], 64, (int)(b * 64));
                if (++stage == kThinWgradStages) stage = 0, phase ^= 1;
            }
        }
    } else if (warp == 1) {
        if (elect_one() && b_begin < b_end) {
            const uint32_t idesc = umma_idesc_bf16(128, N, 1, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (long long b = b_begin; b < b_end; ++b) {
                mbar_wait(&ps.full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * kStage);
                const uint64_t da0 = umma_desc_mnmajor_sw128(sa, 8192);
                const uint64_t db0 = umma_desc_kmajor_sw128(sa + kABytes);
                umma_bf16_ss(tmem_base, da0, db0, idesc, b != b_begin);
                umma_bf16_acc(tmem_base, da0 + 128, db0 + 2, idesc);
                umma_bf16_acc(tmem_base, da0 + 256, db0 + 4, idesc);
                umma_bf16_acc(tmem_base, da0 + 384, db0 + 6, idesc);
                umma_commit(&ps.empty[stage]);
                if (++stage == kThinWgradStages) stage = 0, phase ^= 1;
            }
            umma_commit(&ps.acc_full);
        }
    } else if (warp >= 4) {
        // ---------------- tap-column producers: thread <-> (pixel pair q, filter row ky); then the final flush
        const int ky = warp & 3;
        const int qp = lane;                                   // pixels 2*qp, 2*qp + 1 of the block
        constexpr int PF = 4;                                  // loads of PF blocks in flight (plane reads come from DRAM)
        float v[PF][CIN][6];
        auto load_block = [&](long long b, float (&dst)[CIN][6]) {
            int row, bx, n, oy;
            p.fd_bpr.divmod((int)b, row, bx);
            p.fd_oh.divmod(row, n, oy);
            const int iy = 2 * oy - 1 + ky;
            const bool rok = iy >= 0 && iy < p.ih;
            const int ix0 = 2 * (bx * 64 + 2 * qp) - 1;        // leftmost input column of the pair's 6-wide window
#pragma unroll
            for (int j = 0; j < CIN; ++j) {
                const float* rp = (j == 0 ? p.p0 : p.p1) + ((size_t)n * p.ih + (rok ? iy : 0)) * p.iw;
                const float4 mid = rok ? __ldg(reinterpret_cast<const float4*>(rp + ix0 + 1)) : make_float4(0, 0, 0, 0);
                dst[j][0] = (rok && ix0 >= 0) ? __ldg(rp + ix0) : 0.f;
                dst[j][1] = mid.x, dst[j][2] = mid.y, dst[j][3] = mid.z, dst[j][4] = mid.w;
                dst[j][5] = (rok && ix0 + 5 < p.iw) ? __ldg(rp + ix0 + 5) : 0.f;
            }
        };
#pragma unroll
        for (int f = 0; f < PF; ++f)
            if (b_begin + f < b_end) load_block(b_begin + f, v[f]);
        int stage = 0;
        uint32_t phase = 0;
        for (long long b0 = b_begin; b0 < b_end; b0 += PF) {
#pragma unroll
            for (int f = 0; f < PF; ++f) {
                const long long b = b0 + f;
                if (b >= b_end) break;
                uint32_t words[4 * CIN];
#pragma unroll
                for (int kx = 0; kx < 4; ++kx)
#pragma unroll
                    for (int j = 0; j < CIN; ++j) {
                        __nv_bfloat162 h = __floats2bfloat162_rn(v[f][j][kx], v[f][j][kx + 2]);
                        words[kx * CIN + j] = *reinterpret_cast<uint32_t*>(&h);
                    }
                if (b + PF < b_end) load_block(b + PF, v[f]);             // refill this register slot
                mbar_wait(&ps.empty[stage], phase ^ 1);
                uint8_t* sb = smem + (size_t)stage * kStage + kABytes;
#pragma unroll
                for (int kx = 0; kx < 4; ++kx)
#pragma unroll
                    for (int j = 0; j < CIN; ++j) {
                        const int k = (ky * 4 + kx) * CIN + j;     // B row = tap column of dw
                        // row k, pixels (2qp, 2qp+1): 16-byte chunk qp/4 (XOR-swizzled with the row), word qp%4
                        *reinterpret_cast<uint32_t*>(sb + (uint32_t)k * 128u +
                                                     (uint32_t)((((qp >> 2) ^ (k & 7)) << 4) + ((qp & 3) << 2))) = words[kx * CIN + j];
                    }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ps.full[stage]);
                if (++stage == kThinWgradStages) stage = 0, phase ^= 1;
            }
        }
        if (b_begin < b_end) {
            const int q = warp & 3;
            const int ch = q * 32 + lane;
            mbar_wait(&ps.acc_full, 0);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < N; c0 += 16) {
                uint32_t v16[16];
                tmem_ld_16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v16);
                tmem_ld_wait();
                if (ch < p.c) {
                    float* o = p.dw + (size_t)ch * N + c0;
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        red_add_v4(o + j, __uint_as_float(v16[j]), __uint_as_float(v16[j + 1]), __uint_as_float(v16[j + 2]),
                                   __uint_as_float(v16[j + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 32);
    }
}

// =============================================================================================
// thin_convT_plane: out[n, 2a+py, 2b+px] = act(bias + sum_c sum_{(ky,dy) in T[py]} sum_{(kx,dx) in T[px]}
//                                               x[n, a+dy, b+dx, c] * W[c][ky*4+kx])          (one output channel)
//
// Row streaming: the CTA owns a contiguous range of input rows (of w == 128 pixels).  For every row one GEMM
// (M = 128 pixels, N = 16 taps, K = C) yields the 16 per-tap partial products P[a][b][t]; they go through a 3-slot
// shared-memory ring, and as soon as P[a] exists the output rows 2a-1 and 2a are complete (they need P[a-1], P[a]):
// every input element is read once, nothing but the fp32 output plane is written.  One halo row is recomputed at each
// end of a CTA's range.
static constexpr int kThinPlaneThreads = 256;
static constexpr int kThinPlaneStages = 8;
static constexpr int kPlaneW = 128;

struct ThinPlaneParams {
    int n, h, c;                 // input [n, h, 128, c]
    int kc;                      // c / 64
    const float* bias;           // 1 element or null
    int act;
    float* out;                  // [n, 2h, 256]
};

struct __align__(8) ThinPlanePipe {
    uint64_t full[kThinPlaneStages], empty[kThinPlaneStages], acc_full[2], acc_empty[2], w_full;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThinPlaneThreads, 1)
thin_convT_plane_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                        const ThinPlaneParams p) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    __shared__ ThinPlanePipe ps;
    __shared__ float ring[3][16][kPlaneW];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;                                   // kc x [16 rows x 128 B]
    uint8_t* sA = smem + 4 * 2048;                        // stages x [128 rows x 128 B]   (kc <= 4)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long rows = (long long)p.n * p.h;
    const long long r_begin = rows * blockIdx.x / gridDim.x, r_end = rows * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kThinPlaneStages; ++s) {
            mbar_init(&ps.full[s], 1);
            mbar_init(&ps.empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&ps.acc_full[a], 1);
            mbar_init(&ps.acc_empty[a], 4);
        }
        mbar_init(&ps.w_full, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(&ps.tmem_base, 32);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ps.tmem_base;
    pdl_wait();

    // Every role walks the same list of P rows: for each image segment [a_lo, a_hi) of the CTA's range, the rows
    // a_lo-1 .. a_hi clipped to the image.
    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(&ps.w_full, (uint32_t)p.kc * 2048u);
            for (int kc = 0; kc < p.kc; ++kc) tma_load_2d(sW + kc * 2048, &tm_w, &ps.w_full, kc * 64, 0);
            int stage = 0;
            uint32_t phase = 0;
            long long r = r_begin;
            while (r < r_end) {
                const int img = (int)(r / p.h);
                const int a_lo = (int)(r - (long long)img * p.h);
                const long long img_end = (long long)(img + 1) * p.h;
                const int a_hi = (int)((img_end < r_end ? img_end : r_end) - (long long)img * p.h);
                const int a0 = a_lo > 0 ? a_lo - 1 : 0, a1 = a_hi < p.h ? a_hi : p.h - 1;
                for (int a = a0; a <= a1; ++a)
                    for (int kc = 0; kc < p.kc; ++kc) {
                        mbar_wait(&ps.empty[stage], phase ^ 1);
                        mbar_expect_tx(&ps.full[stage], 16384);
                        tma_load_2d(sA + (size_t)stage * 16384, &tm_x, &ps.full[stage], kc * 64,
                                    (int)(((long long)img * p.h + a) * kPlaneW));
                        if (++stage == kThinPlaneStages) stage = 0, phase ^= 1;
                    }
                r = (long long)img * p.h + a_hi;
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(128, 16, 0, 0);
            mbar_wait(&ps.w_full, 0);
            int stage = 0;
            uint32_t phase = 0;
            int cnt = 0;
            long long r = r_begin;
            while (r < r_end) {
                const int img = (int)(r / p.h);
                const int a_lo = (int)(r - (long long)img * p.h);
                const long long img_end = (long long)(img + 1) * p.h;
                const int a_hi = (int)((img_end < r_end ? img_end : r_end) - (long long)img * p.h);
                const int a0 = a_lo > 0 ? a_lo - 1 : 0, a1 = a_hi < p.h ? a_hi : p.h - 1;
                for (int a = a0; a <= a1; ++a, ++cnt) {
                    const int acc = cnt & 1;
                    mbar_wait(&ps.acc_empty[acc], (uint32_t)(((cnt >> 1) & 1) ^ 1));
                    tc_fence_after();
                    const uint32_t td = tmem_base + acc * 16;
                    for (int kc = 0; kc < p.kc; ++kc) {
                        mbar_wait(&ps.full[stage], phase);
                        tc_fence_after();
                        const uint64_t da0 = umma_desc_kmajor_sw128(smem_u32(sA + (size_t)stage * 16384));
                        const uint64_t db0 = umma_desc_kmajor_sw128(smem_u32(sW + kc * 2048));
                        umma_bf16_ss(td, da0, db0, idesc, kc != 0);
                        umma_bf16_acc(td, da0 + 2, db0 + 2, idesc);
                        umma_bf16_acc(td, da0 + 4, db0 + 4, idesc);
                        umma_bf16_acc(td, da0 + 6, db0 + 6, idesc);
                        umma_commit(&ps.empty[stage]);
                        if (++stage == kThinPlaneStages) stage = 0, phase ^= 1;
                    }
                    umma_commit(&ps.acc_full[acc]);
                }
                r = (long long)img * p.h + a_hi;
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int b = q * 32 + lane;                        // input column == TMEM lane
        const float bias = p.bias != nullptr ? __ldg(p.bias) : 0.f;
        int cnt = 0;
        long long r = r_begin;
        while (r < r_end) {
            const int img = (int)(r / p.h);
            const int a_lo = (int)(r - (long long)img * p.h);
            const long long img_end = (long long)(img + 1) * p.h;
            const int a_hi = (int)((img_end < r_end ? img_end : r_end) - (long long)img * p.h);
            float* out_img = p.out + (size_t)img * (2 * p.h) * (2 * kPlaneW);
            asm volatile("bar.sync 1, 128;" ::: "memory");     // ring slots of the previous segment are no longer read
            for (int a = a_lo - 1; a <= a_hi; ++a) {
                const bool cur_ok = a >= 0 && a < p.h;
                const bool prev_ok = a - 1 >= 0 && a - 1 >= a_lo - 1;      // P[a-1] was produced in this segment
                const int sc = (a + 3) % 3, sp = (a + 2) % 3;
                if (cur_ok) {
                    const int acc = cnt & 1;
                    mbar_wait(&ps.acc_full[acc], (uint32_t)((cnt >> 1) & 1));
                    tc_fence_after();
                    uint32_t v[16];
                    tmem_ld_16(tmem_base + acc * 16 + ((uint32_t)(q * 32) << 16), v);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_relaxed(&ps.acc_empty[acc]);
                    ++cnt;
#pragma unroll
                    for (int t = 0; t < 16; ++t) ring[sc][t][b] = __uint_as_float(v[t]);
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                // P[a][.][t] with ky in {0,1} and P[a-1][.][t] with ky in {2,3} complete output rows 2a-1 and 2a
                const bool lft = b > 0, rgt = b + 1 < kPlaneW;
                if (a >= a_lo && a < a_hi) {                  // row 2a (py = 0): ky = 1 from P[a], ky = 3 from P[a-1]
                    float e = bias, o = bias;                  // x = 2b (kx 1 @ b, kx 3 @ b-1), x = 2b+1 (kx 0 @ b+1, kx 2 @ b)
                    if (cur_ok) {
                        e += ring[sc][5][b] + (lft ? ring[sc][7][b - 1] : 0.f);
                        o += (rgt ? ring[sc][4][b + 1] : 0.f) + ring[sc][6][b];
                    }
                    if (prev_ok) {
                        e += ring[sp][13][b] + (lft ? ring[sp][15][b - 1] : 0.f);
                        o += (rgt ? ring[sp][12][b + 1] : 0.f) + ring[sp][14][b];
                    }
                    if (p.act == PAI_ACT_TANH) e = tanhf(e), o = tanhf(o);
                    *reinterpret_cast<float2*>(out_img + (size_t)(2 * a) * (2 * kPlaneW) + 2 * b) = make_float2(e, o);
                }
                if (a - 1 >= a_lo && a - 1 < a_hi) {          // row 2a-1 (py = 1 of a-1): ky = 0 from P[a], ky = 2 from P[a-1]
                    float e = bias, o = bias;
                    if (cur_ok) {
                        e += ring[sc][1][b] + (lft ? ring[sc][3][b - 1] : 0.f);
                        o += (rgt ? ring[sc][0][b + 1] : 0.f) + ring[sc][2][b];
                    }
                    if (prev_ok) {
                        e += ring[sp][9][b] + (lft ? ring[sp][11][b - 1] : 0.f);
                        o += (rgt ? ring[sp][8][b + 1] : 0.f) + ring[sp][10][b];
                    }
                    if (p.act == PAI_ACT_TANH) e = tanhf(e), o = tanhf(o);
                    *reinterpret_cast<float2*>(out_img + (size_t)(2 * a - 1) * (2 * kPlaneW) + 2 * b) = make_float2(e, o);
                }
            }
            r = (long long)img * p.h + a_hi;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 32);
    }
}

// =============================================================================================
// gather of a stride-1 4x4 convolution with ONE output channel from per-tap partial products (PatchGAN head,
// models/wrapper.py:233):  out[n, oy, ox] = sum_{ky,kx} P[n, oy-1+ky, ox-1+kx, ky*4+kx],  oy < h-1, ox < w-1
__global__ void __launch_bounds__(256)
col2im4x4s1_kernel(const float* __restrict__ P, int ldp, int n, int h, int w, float* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int oh = h - 1, ow = w - 1;
    const int total = n * oh * ow;
    for (int idx = blockIdx.x * 256 + threadIdx.x; idx < total; idx += gridDim.x * 256) {
        const int ox = idx % ow;
        const int r = idx / ow;
        const int oy = r % oh, img = r / oh;
        const float* base = P + (size_t)img * h * w * ldp;
        float s = 0.f;
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
            const int iy = oy - 1 + ky;
            if (iy < 0 || iy >= h) continue;
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) {
                const int ix = ox - 1 + kx;
                if (ix >= 0 && ix < w) s += __ldg(base + ((size_t)iy * w + ix) * ldp + ky * 4 + kx);
            }
        }
        out[idx] = s;
    }
}

static bool aligned16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace pai

using namespace pai;

extern "C" {

int pai_thin_conv4x4s2_fprop(const float* plane0, const float* plane1, int cin, int n, int ih, int iw,
                             const void* w_packed, int cout, const float* bias, void* out1, int ld1, int act1, void* out2,
                             int ld2, int act2, float slope, void* stream_) {
    PAI_REQUIRE(plane0 && w_packed && out1 && (cin == 1 || (cin == 2 && plane1)), "pai_thin_conv4x4s2_fprop: bad planes / cin=%d", cin);
    PAI_REQUIRE(ih % 2 == 0 && iw % 2 == 0 && n > 0, "pai_thin_conv4x4s2_fprop: even image sizes (got %dx%d)", ih, iw);
    PAI_REQUIRE(cout >= 64 && cout <= 256 && cout % 64 == 0, "pai_thin_conv4x4s2_fprop: cout (%d) must be 64..256, %% 64", cout);
    PAI_REQUIRE(ld1 % 8 == 0 && aligned16p(out1) && aligned16p(w_packed) && (out2 == nullptr || (ld2 % 8 == 0 && aligned16p(out2))) &&
                    (bias == nullptr || aligned16p(bias)),
                "pai_thin_conv4x4s2_fprop: outputs / weights / bias must be 16 B aligned with ld %% 8 == 0");
    ThinFpropParams p;
    p.p0 = plane0, p.p1 = plane1, p.n = n, p.ih = ih, p.iw = iw, p.oh = ih / 2, p.ow = iw / 2;
    p.w = (const __nv_bfloat16*)w_packed, p.cout = cout, p.bias = bias;
    p.out1 = (__nv_bfloat16*)out1, p.ld1 = ld1, p.act1 = act1, p.out2 = (__nv_bfloat16*)out2, p.ld2 = ld2, p.act2 = act2;
    p.slope = slope;
    p.total_pix = (long long)n * p.oh * p.ow;
    PAI_REQUIRE(p.total_pix < (1LL << 31) - 128, "pai_thin_conv4x4s2_fprop: tensor too large");
    p.tiles = (int)((p.total_pix + 127) / 128);
    p.fd_ow = make_fastdiv(p.ow), p.fd_oh = make_fastdiv(p.oh);
    const int dev = current_device(), sms = persistent_ctas(dev);
    if (sms < 0) return -1;
    // 256-pixel wide images take the row-streaming form (bulk-copied input rows in a shared-memory ring)
    const bool stream = iw == kThinRowW && (reinterpret_cast<uintptr_t>(plane0) & 15) == 0 &&
                        (plane1 == nullptr || (reinterpret_cast<uintptr_t>(plane1) & 15) == 0) && getenv("PAI_THIN_GATHER") == nullptr;
    const size_t smem = 2 * 16384 + (size_t)cout * 128 + (stream ? (size_t)kThinRing * cin * 2 * kThinRowW * 4 : 0) + 1024;
    static DeviceOnce once;
    if (once.need(dev)) {
        const int cap = 2 * 16384 + 256 * 128 + kThinRing * 2 * 2 * kThinRowW * 4 + 1024;
        PAI_CUDA_OK(cudaFuncSetAttribute(thin_conv_fprop_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        PAI_CUDA_OK(cudaFuncSetAttribute(thin_conv_fprop_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        PAI_CUDA_OK(cudaFuncSetAttribute(thin_conv_fprop_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        PAI_CUDA_OK(cudaFuncSetAttribute(thin_conv_fprop_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        once.mark(dev);
    }
    const int grid = p.tiles < sms ? p.tiles : sms;
    cudaStream_t st = (cudaStream_t)stream_;
    if (cin == 1 && stream)
        PAI_CUDA_OK(launch_pdl(thin_conv_fprop_kernel<1, true>, dim3((unsigned)grid), dim3(kThinFpropThreads), smem, st, 1, p));
    else if (cin == 1)
        PAI_CUDA_OK(launch_pdl(thin_conv_fprop_kernel<1, false>, dim3((unsigned)grid), dim3(kThinFpropThreads), smem, st, 1, p));
    else if (stream)
        PAI_CUDA_OK(launch_pdl(thin_conv_fprop_kernel<2, true>, dim3((unsigned)grid), dim3(kThinFpropThreads), smem, st, 1, p));
    else
        PAI_CUDA_OK(launch_pdl(thin_conv_fprop_kernel<2, false>, dim3((unsigned)grid), dim3(kThinFpropThreads), smem, st, 1, p));
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

int pai_thin_conv4x4s2_wgrad(const void* u, int u_ld, int c, const float* plane0, const float* plane1, int cin, int n,
                             int ih, int iw, float* dw, void* stream) {
    PAI_REQUIRE(u && plane0 && dw && (cin == 1 || (cin == 2 && plane1)), "pai_thin_conv4x4s2_wgrad: bad planes / cin=%d", cin);
    PAI_REQUIRE((c == 64 || c == 128) && u_ld % 8 == 0 && u_ld >= c && aligned16p(u),
                "pai_thin_conv4x4s2_wgrad: c (%d) must be 64 or 128, u 16 B aligned", c);
    PAI_REQUIRE(ih % 2 == 0 && iw % 128 == 0 && (reinterpret_cast<uintptr_t>(plane0) & 15) == 0 &&
                    (plane1 == nullptr || (reinterpret_cast<uintptr_t>(plane1) & 15) == 0),
                "pai_thin_conv4x4s2_wgrad: image width (%d) must be a multiple of 128, planes 16 B aligned", iw);
    ThinWgradParams p;
    p.p0 = plane0, p.p1 = plane1, p.n = n, p.ih = ih, p.iw = iw, p.oh = ih / 2, p.ow = iw / 2, p.c = c, p.dw = dw;
    const long long pixels = (long long)n * p.oh * p.ow;
    PAI_REQUIRE(pixels < (1LL << 31), "pai_thin_conv4x4s2_wgrad: tensor too large");
    p.blocks = pixels / 64;
    p.blocks_per_row = p.ow / 64;
    p.fd_bpr = make_fastdiv(p.blocks_per_row), p.fd_oh = make_fastdiv(p.oh);
    CUtensorMap tm_u;
    uint64_t dims[2] = {(uint64_t)c, (uint64_t)pixels};
    uint64_t str[1] = {(uint64_t)u_ld * 2};
    uint32_t box[2] = {64, 64};
    int rc = encode_tmap_bf16(&tm_u, u, 2, dims, str, box);
    if (rc) return rc;
    const int dev = current_device(), sms = persistent_ctas(dev);
    if (sms < 0) return -1;
    const size_t smem = (size_t)kThinWgradStages * (2 * 8192 + 4096) + 1024;
    static DeviceOnce once;
    if (once.need(dev)) {
        PAI_CUDA_OK(cudaFuncSetAttribute(thin_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PAI_CUDA_OK(cudaFuncSetAttribute(thin_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        once.mark(dev);
    }
    long long grid = p.blocks / 4 > 0 ? p.blocks / 4 : 1;
    if (grid > sms) grid = sms;
    if (cin == 1)
        PAI_CUDA_OK(launch_pdl(thin_wgrad_kernel<1>, dim3((unsigned)grid), dim3(kThinWgradThreads), smem, (cudaStream_t)stream, 1, tm_u, p));
    else
        PAI_CUDA_OK(launch_pdl(thin_wgrad_kernel<2>, dim3((unsigned)grid), dim3(kThinWgradThreads), smem, (cudaStream_t)stream, 1, tm_u, p));
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

int pai_thin_convT4x4s2_plane(const void* x, int n, int h, int w, int c, int x_ld, const void* w_taps, const float* bias,
                              int act, float* out, void* stream) {
    PAI_REQUIRE(x && w_taps && out, "pai_thin_convT4x4s2_plane: null pointer");
    PAI_REQUIRE(w == kPlaneW, "pai_thin_convT4x4s2_plane: rows must be %d pixels wide (got %d)", kPlaneW, w);
    PAI_REQUIRE(c % 64 == 0 && c >= 64 && c <= 256 && x_ld % 8 == 0 && x_ld >= c && aligned16p(x) && aligned16p(w_taps),
                "pai_thin_convT4x4s2_plane: c (%d) must be 64..256, %% 64; operands 16 B aligned", c);
    PAI_REQUIRE(act == PAI_ACT_NONE || act == PAI_ACT_TANH, "pai_thin_convT4x4s2_plane: act must be none or tanh");
    const long long pixels = (long long)n * h * w;
    PAI_REQUIRE(pixels < (1LL << 31), "pai_thin_convT4x4s2_plane: tensor too large");
    CUtensorMap tm_x, tm_w;
    uint64_t dims[2] = {(uint64_t)c, (uint64_t)pixels};
    uint64_t str[1] = {(uint64_t)x_ld * 2};
    uint32_t box[2] = {64, 128};
    int rc = encode_tmap_bf16(&tm_x, x, 2, dims, str, box);
    if (rc) return rc;
    uint64_t wd[2] = {(uint64_t)c, 16};
    uint64_t ws[1] = {(uint64_t)c * 2};
    uint32_t wb[2] = {64, 16};
    rc = encode_tmap_bf16(&tm_w, w_taps, 2, wd, ws, wb);
    if (rc) return rc;
    ThinPlaneParams p;
    p.n = n, p.h = h, p.c = c, p.kc = c / 64, p.bias = bias, p.act = act, p.out = out;
    const int dev = current_device(), sms = persistent_ctas(dev);
    if (sms < 0) return -1;
    const size_t smem = 4 * 2048 + (size_t)kThinPlaneStages * 16384 + 1024;
    static DeviceOnce once;
    if (once.need(dev)) {
        PAI_CUDA_OK(cudaFuncSetAttribute(thin_convT_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        once.mark(dev);
    }
    const long long rows = (long long)n * h;
    long long grid = rows / 4 > 0 ? rows / 4 : 1;          // >= 4 rows per CTA keeps the halo recompute <= 50 %
    if (grid > sms) grid = sms;
    PAI_CUDA_OK(launch_pdl(thin_convT_plane_kernel, dim3((unsigned)grid), dim3(kThinPlaneThreads), smem, (cudaStream_t)stream, 1, tm_x, tm_w, p));
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

int pai_col2im4x4s1(const float* p, int ldp, int n, int h, int w, float* out, void* stream) {
    PAI_REQUIRE(p && out && ldp >= 16 && h >= 2 && w >= 2, "pai_col2im4x4s1: null pointer / ldp < 16 / tiny grid");
    const long long total = (long long)n * (h - 1) * (w - 1);
    PAI_REQUIRE(total < (1LL << 31), "pai_col2im4x4s1: tensor too large");
    const int sms = sm_count(current_device());
    if (sms < 0) return -1;
    long long blocks = (total + 255) / 256;
    if (blocks > sms * 8) blocks = sms * 8;
    PAI_CUDA_OK(launch_pdl(col2im4x4s1_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, 1, p, ldp, n, h, w, out));
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
