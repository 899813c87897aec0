// tcgen05 / TMEM implicit-GEMM convolution kernels for sm_100a.
//
// One kernel family covers every dense contraction of the Pix2Pix U-Net + PatchGAN hot path
// (reference call sites: models/pix2pix.py:63-69 Conv2d(4,2,1), :99-105 ConvTranspose2d(4,2,1),
// models/wrapper.py:229-233 discriminator convs):
//
//   igemm_fprop_kernel : out[pix, co] = act( bias[co] + sum_{tap, ci} A[pix + tap, ci] * Wp[co, tap, ci] )
//       A is an NHWC bf16 tensor seen through a 5-D TMA map (C', W', P, H', N).  A stride-2 4x4
//       convolution uses the "parity split" view C' = 2C, W' = W/2, P = 2, H' = H/2, so every tap is
//       a dense, unit-stride box and zero padding is the TMA out-of-bounds fill.  A transposed
//       convolution runs as 4 sub-pixel phases (blockIdx.z), each a 2x2 unit-stride conv.  The same
//       kernel therefore also is Conv dgrad (== ConvT fprop) and ConvT dgrad (== Conv fprop).
//   igemm_wgrad_kernel : dW[tap][cu][cs] += sum_pix U[pix, cu] * S[pix + tap, cs]
//       both operands MN-major straight out of the NHWC activations / gradients.
//
// Tile: M = 128 pixels (TMEM lanes), N = n_tile channels (TMEM columns), K-block = 64 channels
// (one SWIZZLE_128B row).  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator,
// 4..7 = epilogue (TMEM -> registers -> global).
#include <stdio.h>

#include "pai_common.cuh"
#include "pai_kernels.h"

namespace pai {

static constexpr int kThreads = 256;
static constexpr int kMaxStages = 8;

struct __align__(8) PipeSmem {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t acc_full[2];
    uint64_t acc_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t tmem_cols_for(int n) {
    uint32_t c = 32;
    while ((int)c < n) c <<= 1;
    return c;
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == PAI_ACT_LEAKY) return v > 0.f ? v : v * slope;
    if (act == PAI_ACT_RELU) return fmaxf(v, 0.f);
    if (act == PAI_ACT_TANH) return tanhf(v);
    return v;
}

// 16 consecutive channels of one pixel -> bf16, vectorised when aligned and fully inside the tensor
__device__ __forceinline__ void store_bf16_16(__nv_bfloat16* o, const float (&f)[16], int act, float slope, bool vec,
                                              int valid) {
    if (vec) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(apply_act(f[2 * j], act, slope), apply_act(f[2 * j + 1], act, slope));
            w[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        reinterpret_cast<uint4*>(o)[0] = make_uint4(w[0], w[1], w[2], w[3]);
        reinterpret_cast<uint4*>(o)[1] = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (j < valid) o[j] = __float2bfloat16_rn(apply_act(f[j], act, slope));
    }
}

// =============================================================================================
// Persistent: one CTA per SM walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...  The fp32
// accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
// Tile order: output-channel tile fastest, so CTAs running at the same time share the A tile in L2.
__global__ void __launch_bounds__(kThreads, 1)
igemm_fprop_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                   const IgemmFpropParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ PipeSmem ps;
    __shared__ uint4 stage_buf[4][32 * 8];   // per epilogue warp: 32 rows x 128 B, XOR-swizzled 16-byte chunks
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = p.n_tile;
    const uint32_t a_bytes = 128 * 128;
    const uint32_t b_bytes = (uint32_t)n_tile * 128;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const int stages = p.stages;
    const int num_kb = p.ntaps * p.kc_per_tap;
    const uint32_t acc_cols = tmem_cols_for(n_tile);
    const int total_tiles = p.n_tiles * p.m_tiles * p.phases;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&ps.full[s], 1);
            mbar_init(&ps.empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&ps.acc_full[a], 1);
            mbar_init(&ps.acc_empty[a], 128);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(&ps.tmem_base, 2 * acc_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ps.tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int nt = t % p.n_tiles;
                int r = t / p.n_tiles;
                int mt = r % p.m_tiles;
                const int phase_idx = r / p.m_tiles;
                const int tw = mt % p.tiles_w;
                mt /= p.tiles_w;
                const int th = mt % p.tiles_h;
                const int tn = mt / p.tiles_h;
                const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
                const int col0 = nt * n_tile;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int tap = kb / p.kc_per_tap;
                    const int kc = kb - tap * p.kc_per_tap;
                    const int ti = phase_idx * p.ntaps + tap;
                    mbar_wait(&ps.empty[stage], phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    uint8_t* sb = sa + a_bytes;
                    mbar_expect_tx(&ps.full[stage], stage_bytes);
                    tma_load_5d(sa, &tm_a, &ps.full[stage], p.tap_c[ti] + kc * 64, w0 + p.tap_w[ti], p.tap_p[ti],
                                h0 + p.tap_h[ti], n0);
                    tma_load_2d(sb, &tm_b, &ps.full[stage], kb * 64, phase_idx * p.b_rows_per_phase + col0);
                    if (++stage == stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, n_tile, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const int a = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&ps.acc_empty[a], acc_phase ^ 1);   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + a * acc_cols;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&ps.full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint32_t sb = sa + a_bytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_bf16_ss(tmem_d, umma_desc_kmajor_sw128(sa + k * 32), umma_desc_kmajor_sw128(sb + k * 32),
                                     idesc, (kb | k) != 0);
                    }
                    umma_commit(&ps.empty[stage]);
                    if (++stage == stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&ps.acc_full[a]);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        const int r = q * 32 + lane;
        const int wi = r % p.bw;
        const int hi = (r / p.bw) % p.bh;
        const int ni = r / (p.bw * p.bh);
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int nt = t % p.n_tiles;
            int rr = t / p.n_tiles;
            int mt = rr % p.m_tiles;
            const int phase_idx = rr / p.m_tiles;
            const int tw = mt % p.tiles_w;
            mt /= p.tiles_w;
            const int th = mt % p.tiles_h;
            const int tn = mt / p.tiles_h;
            const int col0 = nt * n_tile;
            const int gw = tw * p.bw + wi, gh = th * p.bh + hi, gn = tn * p.bn + ni;
            const bool row_ok = (gw < p.gw) && (gh < p.gh) && (gn < p.gn);
            const long long off = p.out_phase_off[phase_idx] + (long long)gn * p.out_sn + (long long)gh * p.out_sh +
                                  (long long)gw * p.out_sw + col0;
            const long long off2 = (long long)gn * p.out2_sn + (long long)gh * p.out2_sh + (long long)gw * p.out2_sw + col0;
            const int a = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            mbar_wait(&ps.acc_full[a], acc_phase);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + a * acc_cols + ((uint32_t)(q * 32) << 16);
            const bool vec_ok = ((p.cout & 7) == 0) && ((off & 7) == 0) && ((off2 & 7) == 0);
            const bool all_vec = __all_sync(0xffffffffu, vec_ok || !row_ok);
            for (int c = 0; c < n_tile; c += 64) {
                const int cw = min(64, n_tile - c);
                // ---- coalesced path: 64 channels of 32 rows are transposed through a swizzled smem tile so
                // that 8 lanes write one full 128-byte row segment (4 rows per store instruction)
                if (!p.out_f32 && cw == 64 && all_vec && col0 + c + 64 <= p.cout) {
                    float f[64];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t v[16];
                        tmem_ld_16(tmem_d + (uint32_t)(c + 16 * j), v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float x = __uint_as_float(v[i]);
                            if (p.bias != nullptr) x += __ldg(p.bias + col0 + c + 16 * j + i);
                            f[16 * j + i] = x;
                        }
                    }
                    uint4* tile = stage_buf[q];
#pragma unroll 1
                    for (int which = 0; which < 2; ++which) {
                        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(which == 0 ? p.out : p.out2);
                        if (dst == nullptr) break;
                        const int act = which == 0 ? p.act : p.act2;
                        const long long my_off = which == 0 ? off : off2;
#pragma unroll
                        for (int ch = 0; ch < 8; ++ch) {
                            uint32_t w[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                __nv_bfloat162 h = __floats2bfloat162_rn(apply_act(f[8 * ch + 2 * k], act, p.slope),
                                                                         apply_act(f[8 * ch + 2 * k + 1], act, p.slope));
                                w[k] = *reinterpret_cast<uint32_t*>(&h);
                            }
                            tile[lane * 8 + (ch ^ (lane & 7))] = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int row = i * 4 + (lane >> 3), ch = lane & 7;
                            const uint4 val = tile[row * 8 + (ch ^ (row & 7))];
                            const long long roff = __shfl_sync(0xffffffffu, my_off, row);
                            const int rok = __shfl_sync(0xffffffffu, (int)row_ok, row);
                            if (rok) *reinterpret_cast<uint4*>(dst + roff + c + ch * 8) = val;
                        }
                        __syncwarp();
                    }
                    continue;
                }
                // ---- generic path (fp32 output, narrow or ragged tiles)
                for (int cc = c; cc < c + cw; cc += 16) {
                    uint32_t v[16];
                    __syncwarp();
                    tmem_ld_16(tmem_d + (uint32_t)cc, v);
                    tmem_ld_wait();
                    if (!row_ok || col0 + cc >= p.cout) continue;
                    float f[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float x = __uint_as_float(v[j]);
                        if (p.bias != nullptr && col0 + cc + j < p.cout) x += __ldg(p.bias + col0 + cc + j);
                        f[j] = x;
                    }
                    if (p.out_f32) {
                        float* o = reinterpret_cast<float*>(p.out) + off + cc;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (col0 + cc + j < p.cout) o[j] = apply_act(f[j], p.act, p.slope);
                    } else {
                        store_bf16_16(reinterpret_cast<__nv_bfloat16*>(p.out) + off + cc, f, p.act, p.slope,
                                      vec_ok && col0 + cc + 16 <= p.cout, p.cout - (col0 + cc));
                    }
                    if (p.out2 != nullptr)
                        store_bf16_16(reinterpret_cast<__nv_bfloat16*>(p.out2) + off2 + cc, f, p.act2, p.slope,
                                      vec_ok && col0 + cc + 16 <= p.cout, p.cout - (col0 + cc));
                }
            }
            __syncwarp();
            tc_fence_before();
            mbar_arrive(&ps.acc_empty[a]);      // 128 arrivals release the accumulator to the MMA warp
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * acc_cols);
    }
}

// =============================================================================================
// wgrad: D[cu (128 rows), cs (n_tile cols)] = sum over pixels of U[pix, cu] * S[pix + tap, cs]
// K-block = 64 pixels.  grid = (cu_blocks * cs_blocks, ntaps, splitk)
__global__ void __launch_bounds__(kThreads, 1)
igemm_wgrad_kernel(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_s,
                   const IgemmWgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ PipeSmem ps;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = p.n_tile;
    const uint32_t blk_bytes = 64 * 128;  // one [64 pixels x 64 channels] box
    const uint32_t a_bytes = 2 * blk_bytes;
    const uint32_t b_bytes = (uint32_t)(n_tile / 64) * blk_bytes;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    const int stages = p.stages;

    const int cs_blocks = p.cs / n_tile;
    const int cu0 = (blockIdx.x / cs_blocks) * 128;   // rows >= cu are TMA zero fill and never stored
    const int cs0 = (blockIdx.x % cs_blocks) * n_tile;
    const int tap = blockIdx.y;
    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
    const int per = (total_tiles + gridDim.z - 1) / gridDim.z;
    const int t_begin = blockIdx.z * per;
    const int t_end = min(total_tiles, t_begin + per);
    const int num_kb = t_end - t_begin;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_u);
        tma_prefetch_desc(&tm_s);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&ps.full[s], 1);
            mbar_init(&ps.empty[s], 1);
        }
        mbar_init(&ps.acc_full[0], 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(&ps.tmem_base, tmem_cols_for(n_tile));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ps.tmem_base;

    if (num_kb > 0) {
        if (warp == 0) {
            if (lane == 0) {
                int stage = 0;
                uint32_t phase = 0;
                for (int t = t_begin; t < t_end; ++t) {
                    int mt = t;
                    const int tw = mt % p.tiles_w;
                    mt /= p.tiles_w;
                    const int th = mt % p.tiles_h;
                    const int tn = mt / p.tiles_h;
                    const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
                    mbar_wait(&ps.empty[stage], phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    uint8_t* sb = sa + a_bytes;
                    mbar_expect_tx(&ps.full[stage], stage_bytes);
                    tma_load_5d(sa, &tm_u, &ps.full[stage], cu0, w0, 0, h0, n0);
                    tma_load_5d(sa + blk_bytes, &tm_u, &ps.full[stage], cu0 + 64, w0, 0, h0, n0);
                    for (int nb = 0; nb < n_tile / 64; ++nb)
                        tma_load_5d(sb + nb * blk_bytes, &tm_s, &ps.full[stage], p.tap_c[tap] + cs0 + nb * 64,
                                    w0 + p.tap_w[tap], p.tap_p[tap], h0 + p.tap_h[tap], n0);
                    if (++stage == stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t idesc = umma_idesc_bf16(128, n_tile, 1, 1);
                int stage = 0;
                uint32_t phase = 0;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&ps.full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint32_t sb = sa + a_bytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // 16 pixels (= 16 rows of 128 B) per MMA
                        umma_bf16_ss(tmem_base, umma_desc_mnmajor_sw128(sa + k * 2048, blk_bytes),
                                     umma_desc_mnmajor_sw128(sb + k * 2048, blk_bytes), idesc, (kb | k) != 0);
                    }
                    umma_commit(&ps.empty[stage]);
                    if (++stage == stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&ps.acc_full[0]);
            }
        } else if (warp >= 4) {
            const int q = warp & 3;
            const int r = q * 32 + lane;
            float* o = p.out + ((size_t)tap * p.cu + (cu0 + r)) * (size_t)p.cs + cs0;
            const bool row_ok = (cu0 + r) < p.cu;
            mbar_wait(&ps.acc_full[0], 0);
            tc_fence_after();
            for (int c = 0; c < n_tile; c += 16) {
                uint32_t v[16];
                __syncwarp();
                tmem_ld_16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
                tmem_ld_wait();
                if (!row_ok) continue;
#pragma unroll
                for (int j = 0; j < 16; ++j) atomicAdd(o + c + j, __uint_as_float(v[j]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols_for(n_tile));
    }
}

// =============================================================================================
// host launchers
static int pick_stages(size_t stage_bytes, size_t budget) {
    int s = (int)(budget / stage_bytes);
    if (s > kMaxStages) s = kMaxStages;
    if (s < 2) s = 2;
    return s;
}

int launch_igemm_fprop(const CUtensorMap& tm_a, const CUtensorMap& tm_b, IgemmFpropParams p, int m_tiles,
                       int n_tiles, int phases, cudaStream_t stream) {
    const size_t stage_bytes = 128 * 128 + (size_t)p.n_tile * 128;
    p.stages = pick_stages(stage_bytes, 192 * 1024);
    p.m_tiles = m_tiles, p.n_tiles = n_tiles, p.phases = phases;
    const size_t smem = stage_bytes * p.stages + 1024;
    static bool attr_done = false;
    static int num_sms = 148;
    if (!attr_done) {
        PAI_CUDA_OK(cudaFuncSetAttribute(igemm_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        int dev = 0;
        PAI_CUDA_OK(cudaGetDevice(&dev));
        PAI_CUDA_OK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        attr_done = true;
    }
    const long long total = (long long)m_tiles * n_tiles * phases;
    const int grid = (int)(total < num_sms ? total : num_sms);
    igemm_fprop_kernel<<<grid, kThreads, smem, stream>>>(tm_a, tm_b, p);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_igemm_wgrad(const CUtensorMap& tm_u, const CUtensorMap& tm_s, IgemmWgradParams p, int ntaps, int splitk,
                       cudaStream_t stream) {
    const size_t stage_bytes = 2 * 8192 + (size_t)(p.n_tile / 64) * 8192;
    p.stages = pick_stages(stage_bytes, 196 * 1024);
    const size_t smem = stage_bytes * p.stages + 1024;
    static bool attr_done = false;
    if (!attr_done) {
        PAI_CUDA_OK(cudaFuncSetAttribute(igemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    dim3 grid(((p.cu + 127) / 128) * (p.cs / p.n_tile), ntaps, splitk);
    igemm_wgrad_kernel<<<grid, kThreads, smem, stream>>>(tm_u, tm_s, p);
    PAI_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace pai
