// Fused attention for SAM's ViT (windowed 14x14 and global 64x64) on tcgen05 tensor cores.
//
// Decomposed relative position bias (upstream segment_anything image_encoder.add_decomposed_rel_pos)
// is folded INTO the QK^T contraction by extending the head dimension:
//     Q' = [ q*scale | rel_h(q, 0..S-1) | rel_w(q, 0..S-1) | 0 ]      rel_h(q,j) = q . Rh[qy - j + S-1]
//     K' = [ k       | onehot(ky)       | onehot(kx)       | 0 ]
//     Q'.K'^T = scale*q.k + rel_h(q, ky) + rel_w(q, kx)
// so the kernel is plain flash attention with K-dim DK = pad64(hd + 2S) for QK^T and hd for PV.  Q'/K'/V^T are
// produced by vit_attn_prep (vit_kernels.cu).  Zero-padded window tokens are ordinary keys (no masking), exactly as in
// the reference; only the tile padding (keys >= Lk) is masked.
//
// One CTA = one (batch*window*head, 128-query tile).  warp 0: TMA producer; warp 1: single-thread tcgen05.mma issuer;
// warps 2-5: softmax (one thread per query row; S read from TMEM twice: max pass, exp pass), P written to shared memory
// in the 128B-swizzled K-major layout the second MMA consumes, running output kept in registers (O = O*alpha + P.V).
#include "common.cuh"
#include "tc_common.cuh"
#include "kernels.cuh"
#include "tc_api.cuh"
#include "../../include/sampt_b200.h"

namespace sampt {
using namespace tc;

struct AttnParams {
  int Lq, Lk;        // valid queries / keys per batch-head
  int NT;            // keys per tile (multiple of 16, <= 256)
  int DKB;           // DK / 64
  int HD;            // head dim (multiple of 16, <= 128)
  int nheads;
  __half* out;       // [BH/nheads * Lq, ld_out]: row = (bh / nheads) * Lq + q, col = (bh % nheads) * HD + d
  int ld_out;
  int split_off;     // >0: also write the fp16 residual lo at this column offset
};

constexpr int A_THREADS = 192;

__global__ void __launch_bounds__(A_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, AttnParams p) {
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
  uint64_t* barQ = bars + 0;
  uint64_t* barKV_full = bars + 1;
  uint64_t* barKV_empty = bars + 2;
  uint64_t* barS_full = bars + 3;
  uint64_t* barP_ready = bars + 4;
  uint64_t* barO_full = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, bh = blockIdx.y;
  const int ntiles = (p.Lk + p.NT - 1) / p.NT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(barQ, 1);
      mbar_init(barKV_full, 1);
      mbar_init(barKV_empty, 1);
      mbar_init(barS_full, 1);
      mbar_init(barP_ready, 128);
      mbar_init(barO_full, 1);
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
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(barQ, p.DKB * q_blk_bytes);
      for (int kb = 0; kb < p.DKB; ++kb) tma_load_3d(sQ + kb * q_blk_bytes, &tmQ, barQ, kb * 64, qt * 128, bh);
      for (int j = 0; j < ntiles; ++j) {
        mbar_wait(barKV_empty, (j & 1) ^ 1);
        mbar_expect_tx(barKV_full, p.DKB * k_blk_bytes + NTB * v_blk_bytes);
        for (int kb = 0; kb < p.DKB; ++kb) tma_load_3d(sK + kb * k_blk_bytes, &tmK, barKV_full, kb * 64, j * p.NT, bh);
        for (int nb = 0; nb < NTB; ++nb) tma_load_3d(sV + nb * v_blk_bytes, &tmV, barKV_full, j * p.NT + nb * 64, 0, bh);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_f16(128, p.NT, 0);
      const uint32_t idesc2 = make_idesc_f16(128, p.HD, 0);
      mbar_wait(barQ, 0);
      for (int j = 0; j < ntiles; ++j) {
        mbar_wait(barKV_full, j & 1);
        tc_fence_after();
        // S = Q' K'^T
        for (int kb = 0; kb < p.DKB; ++kb) {
          const uint64_t ad = make_smem_desc_sw128(smem_u32(sQ + kb * q_blk_bytes));
          const uint64_t bd = make_smem_desc_sw128(smem_u32(sK + kb * k_blk_bytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_S, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc1, (kb | k) != 0);
        }
        umma_commit(barS_full);
        // PV = P V
        mbar_wait(barP_ready, j & 1);
        tc_fence_after();
        const int nk16 = p.NT / 16;
        for (int kk = 0; kk < nk16; ++kk) {
          const int nb = kk >> 2, k = kk & 3;
          const uint64_t ad = make_smem_desc_sw128(smem_u32(sP + nb * q_blk_bytes)) + (uint64_t)(2 * k);
          const uint64_t bd = make_smem_desc_sw128(smem_u32(sV + nb * v_blk_bytes)) + (uint64_t)(2 * k);
          umma_f16(tmem_O, ad, bd, idesc2, kk != 0);
        }
        umma_commit(barO_full);
        umma_commit(barKV_empty);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax + output (thread = query row)
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;           // row inside the 128-query tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const int qrow = qt * 128 + r;
    const bool row_ok = qrow < p.Lq;
    float m_run = -INFINITY, l_run = 0.f;
    float o[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) o[i] = 0.f;
    const float LOG2E = 1.4426950408889634f;
    for (int j = 0; j < ntiles; ++j) {
      const int valid = min(p.NT, p.Lk - j * p.NT);
      mbar_wait(barS_full, j & 1);
      tc_fence_after();
      // pass 1: row max
      float mx = -INFINITY;
      for (int c0 = 0; c0 < p.NT; c0 += 16) {
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(tmem_S + lane_addr + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (c0 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f((m_run - m_new) * LOG2E);  // exp2f(-inf) = 0 on the first tile
      // pass 2: P = exp(S - m_new) -> smem (fp16, swizzled K-major), row sum
      float lsum = 0.f;
      for (int c0 = 0; c0 < p.NT; c0 += 16) {
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(tmem_S + lane_addr + (uint32_t)c0, v);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float p0 = (c0 + i < valid) ? exp2f((__uint_as_float(v[i]) - m_new) * LOG2E) : 0.f;
          float p1 = (c0 + i + 1 < valid) ? exp2f((__uint_as_float(v[i + 1]) - m_new) * LOG2E) : 0.f;
          __half2 h = __floats2half2_rn(p0, p1);
          // accumulate the ROUNDED probabilities so that numerator (fp16 P in the MMA) and denominator agree
          float2 hf = __half22float2(h);
          lsum += hf.x + hf.y;
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
        const int nb = c0 >> 6;
        const int ch = (c0 & 63) >> 3;  // 16-byte chunk index inside the 128 B row (two chunks per 16 columns)
        uint8_t* rowp = sP + nb * q_blk_bytes + r * 128;
        *reinterpret_cast<uint4*>(rowp + (((ch + 0) ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      l_run = l_run * alpha + lsum;
      m_run = m_new;
      tc_fence_before();
      fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      mbar_arrive(barP_ready);
      // accumulate O = O*alpha + PV
      mbar_wait(barO_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 16) {
        if (c0 < p.HD) {  // warp-uniform
          uint32_t v[16];
          __syncwarp();
          tmem_ld16(tmem_O + lane_addr + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[c0 + i] = o[c0 + i] * alpha + __uint_as_float(v[i]);
        }
      }
    }
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
            float a = o[c0 + i] * inv, b = o[c0 + i + 1] * inv;
            __half2 h = __floats2half2_rn(a, b);
            float2 hf = __half22float2(h);
            __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
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

int attn_tc(Ctx* c, cudaStream_t st, const __half* Qx, const __half* Kx, const __half* Vt, int BH, int Lq, int Lk, int Lkp,
            int DK, int HD, int NT, int nheads, __half* out, int ld_out, int split_off, int out_f8) {
  if (attn_ws_applicable(Lk, DK, HD, NT))
    return attn_ws(c, st, Qx, Kx, Vt, BH, Lq, Lk, Lkp, DK, HD, NT, nheads, out, ld_out, split_off, out_f8);
  SAMPT_CHECK(!out_f8, "attn_tc: the fp8 output layout is written by attn_ws_kernel only");
  SAMPT_CHECK(DK % 64 == 0 && DK <= 256, "attn_tc: DK=%d must be a multiple of 64 and <= 256", DK);
  SAMPT_CHECK(HD % 16 == 0 && HD <= 128, "attn_tc: HD=%d must be a multiple of 16 and <= 128", HD);
  SAMPT_CHECK(NT % 16 == 0 && NT <= 256, "attn_tc: NT=%d must be a multiple of 16 and <= 256", NT);
  SAMPT_CHECK(Lkp % 8 == 0 && Lkp >= Lk, "attn_tc: Lkp=%d must be a multiple of 8 and >= Lk", Lkp);
  CUtensorMap tmQ, tmK, tmV;
  SAMPT_TRY(make_tmap_3d_f16(&tmQ, Qx, DK, Lq, BH, (uint64_t)DK * 2, (uint64_t)Lq * DK * 2, 64, 128, 1));
  SAMPT_TRY(make_tmap_3d_f16(&tmK, Kx, DK, Lk, BH, (uint64_t)DK * 2, (uint64_t)Lk * DK * 2, 64, NT, 1));
  SAMPT_TRY(make_tmap_3d_f16(&tmV, Vt, Lkp, HD, BH, (uint64_t)Lkp * 2, (uint64_t)HD * Lkp * 2, 64, HD, 1));
  AttnParams p;
  p.Lq = Lq; p.Lk = Lk; p.NT = NT; p.DKB = DK / 64; p.HD = HD; p.nheads = nheads;
  p.out = out; p.ld_out = ld_out; p.split_off = split_off;
  const int NTB = (NT + 63) / 64;
  size_t smem = (size_t)p.DKB * 128 * 128 + (size_t)p.DKB * NT * 128 + (size_t)NTB * HD * 128 + (size_t)NTB * 128 * 128 + 1024 + 256;
  SAMPT_CHECK(smem <= 227 * 1024, "attn_tc: tile configuration needs %zu B of shared memory (> 227 KB)", smem);
  SAMPT_TRY(ensure_func_smem(c, "attn_tc_kernel", attn_tc_kernel, 227 * 1024));
  dim3 grid((Lq + 127) / 128, BH);
  attn_tc_kernel<<<grid, A_THREADS, smem, st>>>(tmQ, tmK, tmV, p);
  c->launches++;
  SAMPT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sampt

using namespace sampt;

// Unit-test entry: softmax(Qx Kx^T) V with pre-extended operands (see header comment).
extern "C" int sampt_attention_f16(sampt_ctx* ctx, const void* Qx, const void* Kx, const void* Vt, int BH, int Lq, int Lk, int Lkp,
                                   int DK, int HD, int NT, int nheads, void* out, int ld_out, int split_off, void* stream) {
  return attn_tc(reinterpret_cast<Ctx*>(ctx), reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const __half*>(Qx),
                 reinterpret_cast<const __half*>(Kx), reinterpret_cast<const __half*>(Vt), BH, Lq, Lk, Lkp, DK, HD, NT, nheads,
                 reinterpret_cast<__half*>(out), ld_out, split_off);
}
