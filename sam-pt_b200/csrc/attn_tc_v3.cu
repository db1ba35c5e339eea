// ON by default (SAMPT_ATTN_V3=0 selects attn_tc_kernel); validated on hardware in round 2 (gpurun_out/exp_attention_v3.log):
// persistent, software-pipelined variant of attn_tc_kernel for SINGLE-tile attention (the ViT's 14x14 windowed blocks:
// Lq = Lk = 196, one 208-key tile).
//
// attn_tc_kernel launches one CTA per (window*head, 128-query tile) and runs  load -> QK^T -> softmax -> P.V -> store  in
// sequence, one CTA per SM (190 KB shared memory): ~54 waves of ~8 us.  Here one CTA per SM loops over its work items and the
// hand-overs between the three roles are arranged so that the softmax warpgroup (the critical path) never waits for a load
// or for the QK^T MMA in steady state:
//
//   TMA warp   : Q'/K' of item i+1 are loaded as soon as the QK^T MMA of item i has completed (the Q'K' buffer is dead from
//                then on); V^T of item i+1 as soon as P.V of item i has completed
//   MMA thread : S_{i+1} is issued BEFORE P.V_i (two S buffers in TMEM), i.e. it executes during the softmax of item i;
//                O_i = P_i.V_i is written over the columns of S_i (dead once P_i is published)
//   softmax WG : item i: row max, exp -> P (single buffer: P.V_{i-1} was observed complete by the epilogue of item i-1),
//                wait O_i, normalise, store, release the S/O columns
//
// Items are ordered (window*head major, query tile minor) so that the second query tile of a window re-reads K'/V^T from L2.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"
#include "kernels.cuh"
#include "tc_api.cuh"

namespace sampt {
using namespace tc;

struct AttnV3Params {
  int Lq, Lk;
  int NT;            // keys per tile (multiple of 16, <= 256), Lk <= NT
  int DKB, HD, nheads;
  int n_qt, n_items; // query tiles per batch-head, total items = n_qt * BH
  __half* out;
  int ld_out, split_off;
};

constexpr int V3_THREADS = 192;

__global__ void __launch_bounds__(V3_THREADS, 1)
attn_tc_v3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, AttnV3Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int NTB = (p.NT + 63) / 64;
  const int q_blk_bytes = 128 * 128;
  const int k_blk_bytes = p.NT * 128;
  const int v_blk_bytes = p.HD * 128;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + p.DKB * q_blk_bytes;
  uint8_t* sV = sK + p.DKB * k_blk_bytes;
  uint8_t* sP = sV + NTB * v_blk_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + NTB * q_blk_bytes);
  uint64_t* barQK_full = bars + 0;
  uint64_t* barQK_empty = bars + 1;
  uint64_t* barV_full = bars + 2;
  uint64_t* barV_empty = bars + 3;
  uint64_t* barS_full = bars + 4;     // [2]
  uint64_t* barS_empty = bars + 6;    // [2]
  uint64_t* barP_full = bars + 8;
  uint64_t* barO_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // items of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...  (local counter i = 0, 1, 2, ...)
  const int n_local = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(barQK_full, 1);
      mbar_init(barQK_empty, 1);
      mbar_init(barV_full, 1);
      mbar_init(barV_empty, 1);
      mbar_init(barP_full, 128);
      mbar_init(barO_full, 1);
      for (int b = 0; b < 2; ++b) {
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
      for (int i = 0; i < n_local; ++i) {
        const int item = (int)blockIdx.x + i * (int)gridDim.x;
        const int bh = item / p.n_qt, qt = item % p.n_qt;
        if (i >= 1) mbar_wait(barQK_empty, (i - 1) & 1);       // QK^T of item i-1 has completed
        mbar_expect_tx(barQK_full, p.DKB * (q_blk_bytes + k_blk_bytes));
        for (int kb = 0; kb < p.DKB; ++kb) tma_load_3d(sQ + kb * q_blk_bytes, &tmQ, barQK_full, kb * 64, qt * 128, bh);
        for (int kb = 0; kb < p.DKB; ++kb) tma_load_3d(sK + kb * k_blk_bytes, &tmK, barQK_full, kb * 64, 0, bh);
        if (i >= 1) mbar_wait(barV_empty, (i - 1) & 1);        // P.V of item i-1 has completed
        mbar_expect_tx(barV_full, NTB * v_blk_bytes);
        for (int nb = 0; nb < NTB; ++nb) tma_load_3d(sV + nb * v_blk_bytes, &tmV, barV_full, nb * 64, 0, bh);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0 && n_local > 0) {
      const uint32_t idesc1 = make_idesc_f16(128, p.NT, 0);
      const uint32_t idesc2 = make_idesc_f16(128, p.HD, 0);
      auto issue_S = [&](int i) {
        const int b = i & 1, u = i >> 1;
        mbar_wait(barQK_full, i & 1);
        if (u >= 1) mbar_wait(barS_empty + b, (u - 1) & 1);   // item i-2 has been stored: its S/O columns are free
        tc_fence_after();
        const uint32_t tS = tmem_base + (uint32_t)(b * 256);
        for (int kb = 0; kb < p.DKB; ++kb) {
          const uint64_t ad = make_smem_desc_sw128(smem_u32(sQ + kb * q_blk_bytes));
          const uint64_t bd = make_smem_desc_sw128(smem_u32(sK + kb * k_blk_bytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tS, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc1, (kb | k) != 0);
        }
        umma_commit(barS_full + b);
        umma_commit(barQK_empty);
      };
      issue_S(0);
      for (int i = 0; i < n_local; ++i) {
        const int b = i & 1;
        if (i + 1 < n_local) issue_S(i + 1);
        mbar_wait(barP_full, i & 1);     // P_i published: every softmax thread has finished reading S_i
        mbar_wait(barV_full, i & 1);
        tc_fence_after();
        const uint32_t tO = tmem_base + (uint32_t)(b * 256);   // O_i overwrites the first HD columns of S_i
        const int nk16 = p.NT / 16;
        for (int kk = 0; kk < nk16; ++kk) {
          const int nb = kk >> 2, k = kk & 3;
          const uint64_t ad = make_smem_desc_sw128(smem_u32(sP + nb * q_blk_bytes)) + (uint64_t)(2 * k);
          const uint64_t bd = make_smem_desc_sw128(smem_u32(sV + nb * v_blk_bytes)) + (uint64_t)(2 * k);
          umma_f16(tO, ad, bd, idesc2, kk != 0);
        }
        umma_commit(barO_full);
        umma_commit(barV_empty);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax + output (thread = query row)
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    const int valid = min(p.NT, p.Lk);
    for (int i = 0; i < n_local; ++i) {
      const int item = (int)blockIdx.x + i * (int)gridDim.x;
      const int bh = item / p.n_qt, qt = item % p.n_qt;
      const int b = i & 1, u = i >> 1;
      const uint32_t tS = tmem_base + (uint32_t)(b * 256);
      const int qrow = qt * 128 + r;
      mbar_wait(barS_full + b, u & 1);
      tc_fence_after();
      float mx = -INFINITY;
      for (int c0 = 0; c0 < p.NT; c0 += 16) {
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(tS + lane_addr + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < valid) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      float lsum = 0.f;
      // the single P buffer is free: this thread observed barO_full of item i-1 (P.V_{i-1} complete) in its epilogue
      for (int c0 = 0; c0 < p.NT; c0 += 16) {
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(tS + lane_addr + (uint32_t)c0, v);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float p0 = (c0 + j < valid) ? exp2f((__uint_as_float(v[j]) - mx) * LOG2E) : 0.f;
          float p1 = (c0 + j + 1 < valid) ? exp2f((__uint_as_float(v[j + 1]) - mx) * LOG2E) : 0.f;
          __half2 h = __floats2half2_rn(p0, p1);
          float2 hf = __half22float2(h);
          lsum += hf.x + hf.y;
          pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
        const int nb = c0 >> 6;
        const int ch = (c0 & 63) >> 3;
        uint8_t* rowp = sP + nb * q_blk_bytes + r * 128;
        *reinterpret_cast<uint4*>(rowp + (((ch + 0) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(barP_full);
      // ---- epilogue of item i: O_i (over the columns of S_i) -> normalise -> global
      mbar_wait(barO_full, i & 1);
      tc_fence_after();
      const float inv = 1.0f / lsum;
      const bool row_ok = qrow < p.Lq;
      const size_t orow = (size_t)(bh / p.nheads) * p.Lq + qrow;
      __half* op = p.out + orow * p.ld_out + (size_t)(bh % p.nheads) * p.HD;
#pragma unroll 1
      for (int c0 = 0; c0 < p.HD; c0 += 16) {
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(tS + lane_addr + (uint32_t)c0, v);
        tmem_ld_wait();
        if (row_ok) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float a = __uint_as_float(v[j]) * inv, bb = __uint_as_float(v[j + 1]) * inv;
            __half2 h = __floats2half2_rn(a, bb);
            float2 hf = __half22float2(h);
            __half2 l = __floats2half2_rn(a - hf.x, bb - hf.y);
            hi[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
            lo[j >> 1] = *reinterpret_cast<uint32_t*>(&l);
          }
          *reinterpret_cast<uint4*>(op + c0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(op + c0 + 8) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          if (p.split_off > 0) {
            *reinterpret_cast<uint4*>(op + p.split_off + c0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4*>(op + p.split_off + c0 + 8) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(barS_empty + b);   // the S/O columns of buffer b may be overwritten by item i+2
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

bool attn_tc_v3_applicable(int Lk, int NT) {
  static const int enabled = [] { const char* e = std::getenv("SAMPT_ATTN_V3"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();   // validated on hardware in round 2: on unless =0
  return enabled && Lk <= NT;
}

int attn_tc_v3(Ctx* c, cudaStream_t st, const __half* Qx, const __half* Kx, const __half* Vt, int BH, int Lq, int Lk, int Lkp, int DK,
               int HD, int NT, int nheads, __half* out, int ld_out, int split_off) {
  SAMPT_CHECK(DK % 64 == 0 && DK <= 256 && HD % 16 == 0 && HD <= 128 && NT % 16 == 0 && NT <= 256 && Lk <= NT && Lkp % 8 == 0 && Lkp >= Lk,
              "attn_tc_v3: unsupported shape");
  CUtensorMap tmQ, tmK, tmV;
  SAMPT_TRY(make_tmap_3d_f16(&tmQ, Qx, DK, Lq, BH, (uint64_t)DK * 2, (uint64_t)Lq * DK * 2, 64, 128, 1));
  SAMPT_TRY(make_tmap_3d_f16(&tmK, Kx, DK, Lk, BH, (uint64_t)DK * 2, (uint64_t)Lk * DK * 2, 64, NT, 1));
  SAMPT_TRY(make_tmap_3d_f16(&tmV, Vt, Lkp, HD, BH, (uint64_t)Lkp * 2, (uint64_t)HD * Lkp * 2, 64, HD, 1));
  AttnV3Params p;
  p.Lq = Lq; p.Lk = Lk; p.NT = NT; p.DKB = DK / 64; p.HD = HD; p.nheads = nheads;
  p.n_qt = (Lq + 127) / 128; p.n_items = p.n_qt * BH;
  p.out = out; p.ld_out = ld_out; p.split_off = split_off;
  const int NTB = (NT + 63) / 64;
  const size_t smem = (size_t)p.DKB * 128 * 128 + (size_t)p.DKB * NT * 128 + (size_t)NTB * HD * 128 + (size_t)NTB * 128 * 128 + 1024 + 256;
  SAMPT_CHECK(smem <= 227 * 1024, "attn_tc_v3: needs %zu B of shared memory", smem);
  SAMPT_TRY(ensure_func_smem(c, "attn_tc_v3_kernel", attn_tc_v3_kernel, 227 * 1024));
  const int grid = std::min(p.n_items, c->num_sms);
  attn_tc_v3_kernel<<<grid, V3_THREADS, smem, st>>>(tmQ, tmK, tmV, p);
  c->launches++;
  SAMPT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sampt
