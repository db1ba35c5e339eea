// Warp-specialised fused attention for SAM's ViT on tcgen05 (round-2 kernel; supersedes attn_tc_kernel for every shape it
// accepts; the two round-1 pipelining drafts attn_tc_v2 / v3 were validated this round, gained 1 % each and were removed).  Operands as in attn_tc.cu: Q' = [q*scale | rel-pos dot products | 0], K' = [k | one-hots | 0]
// (the decomposed relative-position bias rides inside the QK^T contraction), V^T; softmax(Q'K'^T) V per (batch*window*head).
//
// Why the round-1 kernels sat at 6 % of the tensor peak (ncu + CUPTI, profiles/r02_attention.md): ONE softmax warpgroup, one
// warp per SM sub-partition, walking a ~2000-instruction dependent chain per 128x208 tile -- IPC 0.2 -- with the tensor pipe
// idle behind it; pipelining loads and MMAs around that chain (v2, v3) bought 1 %.  This kernel attacks the chain itself:
//
//   * TWO softmax warpgroups (8 warps, 2 per sub-partition), each owning one 128-query tile of the same (window, head) --
//     or of the same global-attention query pair -- so K'/V^T are loaded once for both and the tensor pipe works for one
//     warpgroup while the other is in its softmax;
//   * P never touches shared memory: the softmax warps write fp16 P back into TENSOR MEMORY over the S columns they just read
//     (tcgen05.st) and the P.V MMA takes its A operand from TMEM (tcgen05.mma ... [a_tmem]); O_j lands in the dead upper half of
//     the S region.  That frees 64-128 KB of shared memory and the st.shared / fence.proxy.async traffic per tile;
//   * the softmax inner loops carry no masking on full tiles, read S 64 columns per tcgen05.wait::ld, use one FFMA + one
//     MUFU.EX2 per element, and whole warps whose query rows lie beyond Lq (68 of 128 rows in the second tile of a 14x14
//     window) skip the arithmetic;
//   * persistent: one CTA per SM walks its (bh, query-pair) items; Q' is reloaded as soon as the item's last QK^T has
//     completed, K'_{t+1} as soon as both QK^T_t have, V^T_{t+1} as soon as both P.V_t have.
//
// Roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer (one thread), warps 2-5 softmax WG0, warps 6-9 softmax WG1.
// MMA issue order is staggered  QK0 QK1 | PV0 QK0' PV1 QK1' | ...  so that WG0's next S is being computed while WG1 is still
// in its softmax.  TMEM (512 columns): region g = [256 g, 256 g + 256): S_g [0, NT) fp32, P_g [0, NT/2) fp16 pairs written in
// place, O_g [round32(NT/2), + HD).
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"
#include "kernels.cuh"
#include "tc_api.cuh"

namespace sampt {
using namespace tc;

struct AttnWsParams {
  int Lq, Lk;
  int NT;            // keys per tile (multiple of 16, <= 256)
  int DKB;           // DK / 64
  int HD;            // head dim (multiple of 16, <= 96)
  int nheads;
  int n_pairs;       // query-tile pairs per batch-head = ceil(Lq / 256)
  int n_items;       // n_pairs * BH
  int ntiles;        // key tiles per item = ceil(Lk / NT)
  __half* out;       // [BH/nheads * Lq, ld_out]
  int ld_out, split_off;
  int out_f8;        // with split_off > 0: fp8 correction operands of the proj GEMM (tc_api.cuh) instead of the fp16 remainder
};

constexpr int WS_THREADS = 320;
constexpr float WS_LOG2E = 1.4426950408889634f;

// 64 columns of S for this thread's row: two x32 loads in flight, one wait
__device__ __forceinline__ void ld64(uint32_t taddr, uint32_t (&a)[32], uint32_t (&b)[32]) {
  tmem_ld32(taddr, a);
  tmem_ld32(taddr + 32u, b);
  tmem_ld_wait();
}

template <int HDT>   // compile-time head-dim bound (64 | 80 | 96): sizes the running-output registers
__global__ void __launch_bounds__(WS_THREADS, 1)
attn_ws_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, AttnWsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NTB = (p.NT + 63) / 64;
  const int q_blk_bytes = 128 * 128;          // one 64-column block of a 128-row Q' tile
  const int k_blk_bytes = p.NT * 128;
  const int v_blk_bytes = p.HD * 128;
  uint8_t* sQ = smem;                                   // [2 tiles][DKB] blocks
  uint8_t* sK = sQ + 2 * p.DKB * q_blk_bytes;           // [DKB] blocks
  uint8_t* sV = sK + p.DKB * k_blk_bytes;               // [NTB] blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NTB * v_blk_bytes);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;
  uint64_t* k_empty = bars + 3;
  uint64_t* v_full = bars + 4;
  uint64_t* v_empty = bars + 5;
  uint64_t* s_full = bars + 6;    // [2]
  uint64_t* p_full = bars + 8;    // [2]
  uint64_t* o_full = bars + 10;   // [2]
  uint64_t* s_free = bars + 12;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_local = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // items of this CTA
  const int T = n_local * p.ntiles;                                                           // flat (item, key tile) count

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1); mbar_init(q_empty, 1);
      mbar_init(k_full, 1); mbar_init(k_empty, 1);
      mbar_init(v_full, 1); mbar_init(v_empty, 1);
      for (int g = 0; g < 2; ++g) {
        mbar_init(s_full + g, 1);
        mbar_init(p_full + g, 128);
        mbar_init(o_full + g, 1);
        mbar_init(s_free + g, 128);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int t = 0; t < T; ++t) {
        const int il = t / p.ntiles, j = t % p.ntiles;
        const int item = (int)blockIdx.x + il * (int)gridDim.x;
        const int bh = item / p.n_pairs, pair = item % p.n_pairs;
        if (j == 0) {
          if (il >= 1) mbar_wait(q_empty, (il - 1) & 1);          // every QK^T of the previous item has completed
          mbar_expect_tx(q_full, 2 * p.DKB * q_blk_bytes);
          for (int g = 0; g < 2; ++g)
            for (int kb = 0; kb < p.DKB; ++kb)
              tma_load_3d(sQ + (g * p.DKB + kb) * q_blk_bytes, &tmQ, q_full, kb * 64, pair * 256 + g * 128, bh);
        }
        if (t >= 1) mbar_wait(k_empty, (t - 1) & 1);              // both QK^T of tile t-1 have completed
        mbar_expect_tx(k_full, p.DKB * k_blk_bytes);
        for (int kb = 0; kb < p.DKB; ++kb) tma_load_3d(sK + kb * k_blk_bytes, &tmK, k_full, kb * 64, j * p.NT, bh);
        if (t >= 1) mbar_wait(v_empty, (t - 1) & 1);              // both P.V of tile t-1 have completed
        mbar_expect_tx(v_full, NTB * v_blk_bytes);
        for (int nb = 0; nb < NTB; ++nb) tma_load_3d(sV + nb * v_blk_bytes, &tmV, v_full, j * p.NT + nb * 64, 0, bh);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0 && T > 0) {
      const uint32_t idesc_qk = make_idesc_f16(128, p.NT, 0);
      const uint32_t idesc_pv = make_idesc_f16(128, p.HD, 0);
      const int nk16 = p.NT / 16;
      auto issue_qk = [&](int t, int g) {
        const int il = t / p.ntiles, j = t % p.ntiles;
        if (g == 0) {
          if (j == 0) mbar_wait(q_full, il & 1);
          mbar_wait(k_full, t & 1);
        }
        if (t >= 1) mbar_wait(s_free + g, (t - 1) & 1);           // WG g has read O_{t-1} out of its TMEM region
        tc_fence_after();
        const uint32_t tS = tmem_base + (uint32_t)(g * 256);
        for (int kb = 0; kb < p.DKB; ++kb) {
          const uint64_t ad = make_smem_desc_sw128(smem_u32(sQ + (g * p.DKB + kb) * q_blk_bytes));
          const uint64_t bd = make_smem_desc_sw128(smem_u32(sK + kb * k_blk_bytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tS, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_qk, (kb | k) != 0);
        }
        umma_commit(s_full + g);
        if (g == 1) {
          umma_commit(k_empty);
          if (j == p.ntiles - 1) umma_commit(q_empty);
        }
      };
      auto issue_pv = [&](int t, int g) {
        mbar_wait(p_full + g, t & 1);
        tc_fence_after();
        const uint32_t tP = tmem_base + (uint32_t)(g * 256);
        const uint32_t tO = tP + (uint32_t)(((p.NT / 2 + 31) / 32) * 32);
        for (int kk = 0; kk < nk16; ++kk) {
          const uint64_t bd = make_smem_desc_sw128(smem_u32(sV + (kk >> 2) * v_blk_bytes)) + (uint64_t)(2 * (kk & 3));
          umma_f16_ts(tO, tP + (uint32_t)(kk * 8), bd, idesc_pv, kk != 0);
        }
        umma_commit(o_full + g);
      };
      // Issue order  QK0 QK1 | PV0 QK0' PV1 QK1' | ...  (blocking waits).  An event-driven variant that polled the four candidate
      // operations with try_wait was measured SLOWER (windowed 0.169 vs 0.150 ms, global 2.20 vs 1.84 ms per launch): a poll sweep
      // costs several try_wait latencies, a blocking wait wakes at once.  The softmax chain (IPC 0.4 per sub-partition at 2 warps),
      // not the issue order, is what bounds the kernel (profiles/r02_attention.md).
      issue_qk(0, 0);
      issue_qk(0, 1);
      for (int t = 0; t < T; ++t) {
        mbar_wait(v_full, t & 1);
        issue_pv(t, 0);
        if (t + 1 < T) issue_qk(t + 1, 0);
        issue_pv(t, 1);
        umma_commit(v_empty);
        if (t + 1 < T) issue_qk(t + 1, 1);
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------------ softmax warpgroups
    const int g = (warp - 2) >> 2;                // warpgroup = query tile of the pair
    const int q4 = warp & 3;                      // TMEM lane quarter this warp may access
    const int r = q4 * 32 + lane;                 // row inside the 128-query tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const uint32_t tS = tmem_base + (uint32_t)(g * 256) + lane_addr;
    const uint32_t tO = tS + (uint32_t)(((p.NT / 2 + 31) / 32) * 32);
    float m_run = -INFINITY, l_run = 0.f;
    float o[HDT];
    for (int t = 0; t < T; ++t) {
      const int il = t / p.ntiles, j = t % p.ntiles;
      const int item = (int)blockIdx.x + il * (int)gridDim.x;
      const int bh = item / p.n_pairs, pair = item % p.n_pairs;
      const int qbase = pair * 256 + g * 128;
      const bool warp_live = qbase + q4 * 32 < p.Lq;            // warp-uniform: any valid query row in this warp?
      const int valid = min(p.NT, p.Lk - j * p.NT);             // valid keys in this tile
      if (j == 0) {
        m_run = -INFINITY; l_run = 0.f;
#pragma unroll
        for (int i = 0; i < HDT; ++i) o[i] = 0.f;
      }
      mbar_wait(s_full + g, t & 1);
      tc_fence_after();
      float alpha = 0.f, lsum = 0.f;
      if (warp_live) {
        // ---- pass 1: row maximum over the valid keys
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
        int c0 = 0;
        for (; c0 + 64 <= p.NT; c0 += 64) {
          if (c0 >= valid) break;
          uint32_t a[32], b[32];
          ld64(tS + (uint32_t)c0, a, b);
          if (c0 + 64 <= valid) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              mx0 = fmaxf(mx0, __uint_as_float(a[i])); mx1 = fmaxf(mx1, __uint_as_float(a[i + 1]));
              mx2 = fmaxf(mx2, __uint_as_float(b[i])); mx3 = fmaxf(mx3, __uint_as_float(b[i + 1]));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (c0 + i < valid) mx0 = fmaxf(mx0, __uint_as_float(a[i]));
              if (c0 + 32 + i < valid) mx1 = fmaxf(mx1, __uint_as_float(b[i]));
            }
          }
        }
        for (; c0 < p.NT && c0 < valid; c0 += 16) {               // tail of NT % 64 columns, 16 at a time
          uint32_t v[16];
          tmem_ld16(tS + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (c0 + i < valid) mx0 = fmaxf(mx0, __uint_as_float(v[i]));
        }
        const float m_new = fmaxf(m_run, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)));
        alpha = ex2_approx((m_run - m_new) * WS_LOG2E);            // 0 on the first tile (m_run = -inf)
        m_run = m_new;
        const float mb = -m_new * WS_LOG2E;
        // ---- pass 2: P = exp(S - m) as packed fp16, written back to TMEM over the consumed S columns; fp32 row sum
        float s0 = 0.f, s1 = 0.f;
        c0 = 0;
        for (; c0 + 64 <= p.NT; c0 += 64) {
          uint32_t a[32], b[32];
          ld64(tS + (uint32_t)c0, a, b);
          uint32_t pa[16], pb[16];
          if (c0 + 64 <= valid) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float x0 = ex2_approx(fmaf(__uint_as_float(a[i]), WS_LOG2E, mb));
              const float x1 = ex2_approx(fmaf(__uint_as_float(a[i + 1]), WS_LOG2E, mb));
              const float y0 = ex2_approx(fmaf(__uint_as_float(b[i]), WS_LOG2E, mb));
              const float y1 = ex2_approx(fmaf(__uint_as_float(b[i + 1]), WS_LOG2E, mb));
              s0 += x0 + x1; s1 += y0 + y1;
              __half2 hx = __floats2half2_rn(x0, x1), hy = __floats2half2_rn(y0, y1);
              pa[i >> 1] = *reinterpret_cast<uint32_t*>(&hx);
              pb[i >> 1] = *reinterpret_cast<uint32_t*>(&hy);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float x0 = (c0 + i < valid) ? ex2_approx(fmaf(__uint_as_float(a[i]), WS_LOG2E, mb)) : 0.f;
              const float x1 = (c0 + i + 1 < valid) ? ex2_approx(fmaf(__uint_as_float(a[i + 1]), WS_LOG2E, mb)) : 0.f;
              const float y0 = (c0 + 32 + i < valid) ? ex2_approx(fmaf(__uint_as_float(b[i]), WS_LOG2E, mb)) : 0.f;
              const float y1 = (c0 + 33 + i < valid) ? ex2_approx(fmaf(__uint_as_float(b[i + 1]), WS_LOG2E, mb)) : 0.f;
              s0 += x0 + x1; s1 += y0 + y1;
              __half2 hx = __floats2half2_rn(x0, x1), hy = __floats2half2_rn(y0, y1);
              pa[i >> 1] = *reinterpret_cast<uint32_t*>(&hx);
              pb[i >> 1] = *reinterpret_cast<uint32_t*>(&hy);
            }
          }
          tmem_st16(tS + (uint32_t)(c0 >> 1), pa);
          tmem_st16(tS + (uint32_t)((c0 >> 1) + 16), pb);
        }
        for (; c0 < p.NT; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(tS + (uint32_t)c0, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float x0 = (c0 + i < valid) ? ex2_approx(fmaf(__uint_as_float(v[i]), WS_LOG2E, mb)) : 0.f;
            const float x1 = (c0 + i + 1 < valid) ? ex2_approx(fmaf(__uint_as_float(v[i + 1]), WS_LOG2E, mb)) : 0.f;
            s0 += x0 + x1;
            __half2 hx = __floats2half2_rn(x0, x1);
            pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hx);
            pk[8 + (i >> 1)] = 0u;
          }
          // 16 S columns -> 8 packed P columns: write them with the low half of an x16 store (the upper 8 columns it touches are
          // S columns [c0/2 + 8, c0/2 + 16) < c0, already consumed)
          tmem_st16(tS + (uint32_t)(c0 >> 1), pk);
        }
        lsum = s0 + s1;
        tmem_st_wait();
      }
      l_run = l_run * alpha + lsum;
      tc_fence_before();
      mbar_arrive(p_full + g);
      // ---- O_t = P_t . V_t  ->  running output in registers
      mbar_wait(o_full + g, t & 1);
      tc_fence_after();
      if (warp_live) {
#pragma unroll
        for (int d0 = 0; d0 < HDT; d0 += 16) {
          if (d0 < p.HD) {   // warp-uniform
            uint32_t v[16];
            tmem_ld16(tO + (uint32_t)d0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[d0 + i] = fmaf(o[d0 + i], alpha, __uint_as_float(v[i]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(s_free + g);
      if (j == p.ntiles - 1) {
        const int qrow = qbase + r;
        if (qrow < p.Lq) {
          const float inv = 1.0f / l_run;
          const size_t orow = (size_t)(bh / p.nheads) * p.Lq + qrow;
          __half* op = p.out + orow * p.ld_out + (size_t)(bh % p.nheads) * p.HD;
#pragma unroll
          for (int d0 = 0; d0 < HDT; d0 += 8) {
            if (d0 < p.HD) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int i = 0; i < 8; i += 2) {
                const float a = o[d0 + i] * inv, b = o[d0 + i + 1] * inv;
                __half2 h = __floats2half2_rn(a, b);
                const float2 hf = __half22float2(h);
                __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
                hi[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
                lo[i >> 1] = *reinterpret_cast<uint32_t*>(&l);
              }
              *reinterpret_cast<uint4*>(op + d0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              if (p.split_off > 0 && p.out_f8) {
                // e4m3(remainder * 2^12) | e4m3(value * 2^-3): split_off BYTES each, right behind the split_off hi halves of the row
                uint8_t* ob = reinterpret_cast<uint8_t*>(p.out + orow * p.ld_out + p.split_off) + (size_t)(bh % p.nheads) * p.HD + d0;
                uint32_t l8[2], h8[2];
#pragma unroll
                for (int i = 0; i < 8; i += 4) {
                  const float a0 = o[d0 + i] * inv, a1 = o[d0 + i + 1] * inv, a2 = o[d0 + i + 2] * inv, a3 = o[d0 + i + 3] * inv;
                  const float2 f0 = __half22float2(*reinterpret_cast<__half2*>(&hi[i >> 1]));
                  const float2 f1 = __half22float2(*reinterpret_cast<__half2*>(&hi[(i >> 1) + 1]));
                  l8[i >> 2] = cvt_e4m3x4((a0 - f0.x) * F8_LO_SCALE, (a1 - f0.y) * F8_LO_SCALE, (a2 - f1.x) * F8_LO_SCALE, (a3 - f1.y) * F8_LO_SCALE);
                  h8[i >> 2] = cvt_e4m3x4(a0 * F8_HI_SCALE, a1 * F8_HI_SCALE, a2 * F8_HI_SCALE, a3 * F8_HI_SCALE);
                }
                *reinterpret_cast<uint2*>(ob) = make_uint2(l8[0], l8[1]);
                *reinterpret_cast<uint2*>(ob + p.split_off) = make_uint2(h8[0], h8[1]);
              } else if (p.split_off > 0) {
                *reinterpret_cast<uint4*>(op + p.split_off + d0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static size_t attn_ws_smem(int NT, int DK, int HD) {
  const int DKB = DK / 64, NTB = (NT + 63) / 64;
  return (size_t)2 * DKB * 128 * 128 + (size_t)DKB * NT * 128 + (size_t)NTB * HD * 128 + 1024 + 256;
}

bool attn_ws_applicable(int Lk, int DK, int HD, int NT) {
  static const int enabled = [] { const char* e = std::getenv("SAMPT_ATTN_WS"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();
  if (!enabled) return false;
  if (DK % 64 != 0 || DK > 256 || HD % 16 != 0 || HD > 96 || NT % 16 != 0 || NT > 256 || NT < 16) return false;
  if (((NT / 2 + 31) / 32) * 32 + HD > 256) return false;    // O_g must fit behind P_g inside the 256-column region
  return attn_ws_smem(NT, DK, HD) <= 227 * 1024;
}

int attn_ws(Ctx* c, cudaStream_t st, const __half* Qx, const __half* Kx, const __half* Vt, int BH, int Lq, int Lk, int Lkp, int DK,
            int HD, int NT, int nheads, __half* out, int ld_out, int split_off, int out_f8) {
  SAMPT_CHECK(Lkp % 8 == 0 && Lkp >= Lk, "attn_ws: Lkp=%d must be a multiple of 8 and >= Lk", Lkp);
  CUtensorMap tmQ, tmK, tmV;
  SAMPT_TRY(make_tmap_3d_f16(&tmQ, Qx, DK, Lq, BH, (uint64_t)DK * 2, (uint64_t)Lq * DK * 2, 64, 128, 1));
  SAMPT_TRY(make_tmap_3d_f16(&tmK, Kx, DK, Lk, BH, (uint64_t)DK * 2, (uint64_t)Lk * DK * 2, 64, NT, 1));
  SAMPT_TRY(make_tmap_3d_f16(&tmV, Vt, Lkp, HD, BH, (uint64_t)Lkp * 2, (uint64_t)HD * Lkp * 2, 64, HD, 1));
  AttnWsParams p;
  p.Lq = Lq; p.Lk = Lk; p.NT = NT; p.DKB = DK / 64; p.HD = HD; p.nheads = nheads;
  p.n_pairs = (Lq + 255) / 256;
  p.n_items = p.n_pairs * BH;
  p.ntiles = (Lk + NT - 1) / NT;
  p.out = out; p.ld_out = ld_out; p.split_off = split_off; p.out_f8 = out_f8;
  const size_t smem = attn_ws_smem(NT, DK, HD);
  const int grid = std::min(p.n_items, c->num_sms);
  if (HD <= 64) {
    SAMPT_TRY(ensure_func_smem(c, "attn_ws_kernel<64>", attn_ws_kernel<64>, 227 * 1024));
    attn_ws_kernel<64><<<grid, WS_THREADS, smem, st>>>(tmQ, tmK, tmV, p);
  } else if (HD <= 80) {
    SAMPT_TRY(ensure_func_smem(c, "attn_ws_kernel<80>", attn_ws_kernel<80>, 227 * 1024));
    attn_ws_kernel<80><<<grid, WS_THREADS, smem, st>>>(tmQ, tmK, tmV, p);
  } else {
    SAMPT_TRY(ensure_func_smem(c, "attn_ws_kernel<96>", attn_ws_kernel<96>, 227 * 1024));
    attn_ws_kernel<96><<<grid, WS_THREADS, smem, st>>>(tmQ, tmK, tmV, p);
  }
  c->launches++;
  SAMPT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sampt
