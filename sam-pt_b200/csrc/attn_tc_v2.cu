// ON by default (SAMPT_ATTN_V2=0 selects attn_tc_kernel); validated on hardware in round 2 (gpurun_out/exp_attention_v2_*.log):
// software-pipelined variant of attn_tc_kernel for attention with several key tiles (the ViT's global blocks: Lk = 4096,
// DK = 256 because the one-hot rel-pos extension adds 2 x 64 columns).
//
// attn_tc_kernel runs  load K/V -> S = Q'K'^T -> softmax -> P.V -> O  strictly in sequence per 256-key tile.  Here the key tile
// is 128 wide and the hand-overs are arranged so that the softmax warpgroup (the critical path) does not wait for loads or
// for the QK^T MMA in steady state:
//
//   TMA warp   : K' tile j+1 is loaded as soon as the QK^T MMA of tile j has completed (K' is dead from then on; one buffer,
//                released early); V^T tiles use a 2-stage ring released by the completion of P.V
//   MMA thread : S_{j+1} is issued BEFORE P.V_j (S has two TMEM buffers), so it executes while the softmax of tile j runs
//   softmax WG : tile j: row max;  DEFERRED accumulation O = O*alpha_{j-1} + (P.V)_{j-1} (finished in the meantime);
//                exp -> P_j (single shared buffer: free once (P.V)_{j-1} has been observed complete)
//
// TMEM columns: S0 [0,128) S1 [128,256) O [256,256+HD).  Shared memory (ViT-H global: DK=256, HD=80): Q' 64 KB + K' 64 KB +
// V^T 2x20 KB + P 32 KB = 200 KB.  Numerics are identical to attn_tc_kernel except for the tile size (the running max is
// updated every 128 instead of 256 keys).
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"
#include "kernels.cuh"
#include "tc_api.cuh"

namespace sampt {
using namespace tc;

struct AttnV2Params {
  int Lq, Lk;
  int DKB;           // DK / 64
  int HD;
  int nheads;
  __half* out;
  int ld_out;
  int split_off;
};

constexpr int V2_NT = 128;       // keys per tile
constexpr int V2_NTB = 2;        // 64-key blocks per tile
constexpr int V2_THREADS = 192;

__global__ void __launch_bounds__(V2_THREADS, 1)
attn_tc_v2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, AttnV2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int q_blk_bytes = 128 * 128;            // 128 rows x 64 halves
  const int k_blk_bytes = V2_NT * 128;          // 128 keys x 64 halves
  const int v_blk_bytes = p.HD * 128;           // HD rows x 64 keys
  const int k_stage = p.DKB * k_blk_bytes, v_stage = V2_NTB * v_blk_bytes, p_buf = V2_NTB * q_blk_bytes;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + p.DKB * q_blk_bytes;       // one buffer, released right after the QK^T MMA
  uint8_t* sV = sK + k_stage;                   // [2 stages]
  uint8_t* sP = sV + 2 * v_stage;               // one buffer
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + p_buf);
  uint64_t* barQ = bars + 0;
  uint64_t* barK_full = bars + 1;
  uint64_t* barK_empty = bars + 2;
  uint64_t* barV_full = bars + 3;     // [2]
  uint64_t* barV_empty = bars + 5;    // [2]
  uint64_t* barS_full = bars + 7;     // [2]
  uint64_t* barS_empty = bars + 9;    // [2]
  uint64_t* barP_full = bars + 11;
  uint64_t* barO_full = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, bh = blockIdx.y;
  const int ntiles = (p.Lk + V2_NT - 1) / V2_NT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(barQ, 1);
      mbar_init(barK_full, 1);
      mbar_init(barK_empty, 1);
      mbar_init(barP_full, 128);
      mbar_init(barO_full, 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(barV_full + b, 1);
        mbar_init(barV_empty + b, 1);
        mbar_init(barS_full + b, 1);
        mbar_init(barS_empty + b, 128);
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
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(barQ, p.DKB * q_blk_bytes);
      for (int kb = 0; kb < p.DKB; ++kb) tma_load_3d(sQ + kb * q_blk_bytes, &tmQ, barQ, kb * 64, qt * 128, bh);
      for (int j = 0; j < ntiles; ++j) {
        const int b = j & 1, u = j >> 1;
        if (j >= 1) mbar_wait(barK_empty, (j - 1) & 1);        // QK^T of tile j-1 has completed
        mbar_expect_tx(barK_full, k_stage);
        for (int kb = 0; kb < p.DKB; ++kb) tma_load_3d(sK + kb * k_blk_bytes, &tmK, barK_full, kb * 64, j * V2_NT, bh);
        if (u >= 1) mbar_wait(barV_empty + b, (u - 1) & 1);   // P.V of tile j-2 has completed
        mbar_expect_tx(barV_full + b, v_stage);
        for (int nb = 0; nb < V2_NTB; ++nb)
          tma_load_3d(sV + b * v_stage + nb * v_blk_bytes, &tmV, barV_full + b, j * V2_NT + nb * 64, 0, bh);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_f16(128, V2_NT, 0);
      const uint32_t idesc2 = make_idesc_f16(128, p.HD, 0);
      auto issue_S = [&](int j) {
        const int b = j & 1, u = j >> 1;
        mbar_wait(barK_full, j & 1);
        if (u >= 1) mbar_wait(barS_empty + b, (u - 1) & 1);   // the softmax warps have read S of tile j-2
        tc_fence_after();
        const uint32_t tS = tmem_base + (uint32_t)(b * 128);
        for (int kb = 0; kb < p.DKB; ++kb) {
          const uint64_t ad = make_smem_desc_sw128(smem_u32(sQ + kb * q_blk_bytes));
          const uint64_t bd = make_smem_desc_sw128(smem_u32(sK + kb * k_blk_bytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tS, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc1, (kb | k) != 0);
        }
        umma_commit(barS_full + b);
        umma_commit(barK_empty);                               // K' buffer free for tile j+1
      };
      mbar_wait(barQ, 0);
      issue_S(0);
      const uint32_t tO = tmem_base + 256u;
      for (int j = 0; j < ntiles; ++j) {
        const int b = j & 1, u = j >> 1;
        if (j + 1 < ntiles) issue_S(j + 1);                    // runs on the tensor pipe while softmax(j) is in progress
        // P_j is complete; every softmax thread has accumulated O of tile j-1 before writing P_j, so O may be overwritten
        mbar_wait(barP_full, j & 1);
        mbar_wait(barV_full + b, u & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < V2_NT / 16; ++kk) {
          const int nb = kk >> 2, k = kk & 3;
          const uint64_t ad = make_smem_desc_sw128(smem_u32(sP + nb * q_blk_bytes)) + (uint64_t)(2 * k);
          const uint64_t bd = make_smem_desc_sw128(smem_u32(sV + b * v_stage + nb * v_blk_bytes)) + (uint64_t)(2 * k);
          umma_f16(tO, ad, bd, idesc2, kk != 0);
        }
        umma_commit(barO_full);
        umma_commit(barV_empty + b);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax + output (thread = query row)
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const int qrow = qt * 128 + r;
    const bool row_ok = qrow < p.Lq;
    float m_run = -INFINITY, l_run = 0.f;
    float alpha_prev = 0.f;  // alpha of the tile whose accumulation is still pending
    float o[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) o[i] = 0.f;
    const float LOG2E = 1.4426950408889634f;

    auto accumulate = [&](int t, float alpha) {  // O = O*alpha_t + (P.V)_t
      mbar_wait(barO_full, t & 1);
      tc_fence_after();
      const uint32_t tO = tmem_base + 256u;
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 16) {
        if (c0 < p.HD) {  // warp-uniform
          uint32_t v[16];
          __syncwarp();
          tmem_ld16(tO + lane_addr + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[c0 + i] = o[c0 + i] * alpha + __uint_as_float(v[i]);
        }
      }
    };

    for (int j = 0; j < ntiles; ++j) {
      const int b = j & 1, u = j >> 1;
      const int valid = min(V2_NT, p.Lk - j * V2_NT);
      const uint32_t tS = tmem_base + (uint32_t)(b * 128);
      mbar_wait(barS_full + b, u & 1);
      tc_fence_after();
      float mx = -INFINITY;
#pragma unroll
      for (int c0 = 0; c0 < V2_NT; c0 += 16) {
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(tS + lane_addr + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f((m_run - m_new) * LOG2E);
      // deferred accumulation of tile j-1: its P.V has had the whole max pass to finish.  Observing barO_full here also
      // proves that the tensor core is done reading the (single) P buffer and the O columns, so both may be reused below.
      if (j >= 1) accumulate(j - 1, alpha_prev);
      alpha_prev = alpha;
      float lsum = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < V2_NT; c0 += 16) {
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(tS + lane_addr + (uint32_t)c0, v);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float p0 = (c0 + i < valid) ? exp2f((__uint_as_float(v[i]) - m_new) * LOG2E) : 0.f;
          float p1 = (c0 + i + 1 < valid) ? exp2f((__uint_as_float(v[i + 1]) - m_new) * LOG2E) : 0.f;
          __half2 h = __floats2half2_rn(p0, p1);
          float2 hf = __half22float2(h);
          lsum += hf.x + hf.y;
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
        const int nb = c0 >> 6;
        const int ch = (c0 & 63) >> 3;
        uint8_t* rowp = sP + nb * q_blk_bytes + r * 128;
        *reinterpret_cast<uint4*>(rowp + (((ch + 0) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      l_run = l_run * alpha + lsum;
      m_run = m_new;
      tc_fence_before();
      mbar_arrive(barS_empty + b);   // S buffer b may be overwritten by tile j+2
      fence_proxy_async();           // generic-proxy writes of P -> visible to the tensor core
      mbar_arrive(barP_full);
    }
    accumulate(ntiles - 1, alpha_prev);

    const float inv = 1.0f / l_run;
    if (row_ok) {
      const size_t orow = (size_t)(bh / p.nheads) * p.Lq + qrow;
      __half* op = p.out + orow * p.ld_out + (size_t)(bh % p.nheads) * p.HD;
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 8) {
        if (c0 < p.HD) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            float a = o[c0 + i] * inv, bb = o[c0 + i + 1] * inv;
            __half2 h = __floats2half2_rn(a, bb);
            float2 hf = __half22float2(h);
            __half2 l = __floats2half2_rn(a - hf.x, bb - hf.y);
            hi[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
            lo[i >> 1] = *reinterpret_cast<uint32_t*>(&l);
          }
          *reinterpret_cast<uint4*>(op + c0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          if (p.split_off > 0) *reinterpret_cast<uint4*>(op + p.split_off + c0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
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

// true when the experimental kernel is enabled and applicable (several key tiles, fits shared memory)
bool attn_tc_v2_applicable(int Lk, int DK, int HD) {
  static const int enabled = [] { const char* e = std::getenv("SAMPT_ATTN_V2"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();   // validated on hardware in round 2: on unless =0
  if (!enabled || Lk <= 2 * V2_NT) return false;
  const int DKB = DK / 64;
  const size_t smem = (size_t)DKB * 128 * 128 + (size_t)DKB * V2_NT * 128 + 2 * (size_t)V2_NTB * HD * 128 + (size_t)V2_NTB * 128 * 128 + 1024 + 256;
  return smem <= 227 * 1024;
}

int attn_tc_v2(Ctx* c, cudaStream_t st, const __half* Qx, const __half* Kx, const __half* Vt, int BH, int Lq, int Lk, int Lkp, int DK,
               int HD, int nheads, __half* out, int ld_out, int split_off) {
  SAMPT_CHECK(DK % 64 == 0 && DK <= 256 && HD % 16 == 0 && HD <= 128 && Lkp % 8 == 0 && Lkp >= Lk, "attn_tc_v2: unsupported shape");
  CUtensorMap tmQ, tmK, tmV;
  SAMPT_TRY(make_tmap_3d_f16(&tmQ, Qx, DK, Lq, BH, (uint64_t)DK * 2, (uint64_t)Lq * DK * 2, 64, 128, 1));
  SAMPT_TRY(make_tmap_3d_f16(&tmK, Kx, DK, Lk, BH, (uint64_t)DK * 2, (uint64_t)Lk * DK * 2, 64, V2_NT, 1));
  SAMPT_TRY(make_tmap_3d_f16(&tmV, Vt, Lkp, HD, BH, (uint64_t)Lkp * 2, (uint64_t)HD * Lkp * 2, 64, HD, 1));
  AttnV2Params p;
  p.Lq = Lq; p.Lk = Lk; p.DKB = DK / 64; p.HD = HD; p.nheads = nheads; p.out = out; p.ld_out = ld_out; p.split_off = split_off;
  const size_t smem = (size_t)p.DKB * 128 * 128 + (size_t)p.DKB * V2_NT * 128 + 2 * (size_t)V2_NTB * HD * 128 +
                      (size_t)V2_NTB * 128 * 128 + 1024 + 256;
  SAMPT_CHECK(smem <= 227 * 1024, "attn_tc_v2: needs %zu B of shared memory", smem);
  SAMPT_TRY(ensure_func_smem(c, "attn_tc_v2_kernel", attn_tc_v2_kernel, 227 * 1024));
  dim3 grid((Lq + 127) / 128, BH);
  attn_tc_v2_kernel<<<grid, V2_THREADS, smem, st>>>(tmQ, tmK, tmV, p);
  c->launches++;
  SAMPT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sampt
