// Attention operand preparation for the ViT, round-2 kernel (supersedes attn_prep_kernel of vit_kernels.cu).
//
// Same outputs (see attn_tc.cu / attn_ws.cu):  Q' = [q*scale | rel_h(q, 0..S-1) | rel_w(q, 0..S-1) | 0]   [BH, L, DK]
//                                               K' = [k       | onehot(ky)       | onehot(kx)       | 0]   [BH, L, DK]
//                                               V^T                                                        [BH, HD, Lkp]
// with rel_h(q, j) = q . Rh[qy - j + S-1], rel_w(q, j) = q . Rw[qx - j + S-1] (upstream add_decomposed_rel_pos, unscaled q).
//
// Round 1 computed the 2S dot products per query on the CUDA cores (440 k FMA per 14x14 window-head, fed from shared memory at
// ~1 load per 5 FMA): 0.66 ms per 10-frame launch for 0.9 GB of traffic, 1.3 TB/s.  Here they are ONE small tensor-core product
// per CTA:  T = q [TC x HD] . Rcat^T [HD x 2(2S-1)],  Rcat = [Rh ; Rw], as legacy mma.sync m16n8k16 tiles (the operands are tiny
// and live in shared memory; tcgen05 + TMEM would be set-up cost only), with Rcat carried as fp16 hi + lo so that the table is
// exact to ~2^-22 -- the fp32 table of the reference -- and q as the fp16 it already is.  The shifted pick
// Q'ext[t][j] = T[t][qy - j + S-1] is then a shared-memory gather while the rows are assembled.  Everything else is coalesced
// 16-byte traffic, so the kernel is bound by its algorithmic bytes (read qkv once, write Q', K', V^T once).
#include <cstdlib>
#include <mma.h>

#include "common.cuh"
#include "kernels.cuh"

namespace sampt {

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// grid (chunks of TC tokens, heads, windows*frames); 256 threads.  TC % 16 == 0.
template <int HD>
__global__ void __launch_bounds__(256)
attn_prep2_kernel(const __half* __restrict__ qkv, int ldq, const float* __restrict__ relh, const float* __restrict__ relw,
                  __half* __restrict__ Qx, __half* __restrict__ Kx, __half* __restrict__ Vt, int S, int L, int Lkp, int DK, int D,
                  int nheads, float scale, int TC, int NRP /* padded table rows: multiple of 8 >= 2(2S-1) */) {
  constexpr int QP = HD + 8;            // row pitch in halves: 16 B aligned, conflict-free for ldmatrix
  constexpr int KS = HD / 16;           // k-steps of the MMA
  extern __shared__ __align__(16) unsigned char smraw[];
  __half* sq = reinterpret_cast<__half*>(smraw);                 // [TC][QP]   q (unscaled)
  __half* sv = sq + (size_t)TC * QP;                             // [TC][QP]   v, later re-used for T
  const int TP = NRP + 8;                                        // T row pitch (halves)
  const size_t svt = (size_t)TC * (size_t)max(QP, TP);          // the v tile and (later) T share this region
  __half* sRh = sv + svt;                                        // [NRP][QP]  Rcat hi
  __half* sRl = sRh + (size_t)NRP * QP;                          // [NRP][QP]  Rcat lo
  __half* sT = sv;                                               // [TC][TP]   T = q . Rcat^T as fp16 (aliases sv once V^T is out)
  const int chunk = blockIdx.x, h = blockIdx.y, wb = blockIdx.z;
  const int t0 = chunk * TC;
  const int nt = min(TC, L - t0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t bh = (size_t)wb * nheads + h;
  constexpr int SEG = HD / 8;           // 16-byte segments per head row
  const int NR = 2 * (2 * S - 1);

  // ---- phase 1: coalesced 16 B loads of q / k / v head rows.  k goes straight back out as the dot part of K'.
  for (int i = tid; i < TC * SEG; i += 256) {
    const int t = i / SEG, sgm = i % SEG;
    uint4 qv = make_uint4(0u, 0u, 0u, 0u), vv = make_uint4(0u, 0u, 0u, 0u);
    if (t < nt) {
      const __half* rowp = qkv + (size_t)((size_t)wb * L + t0 + t) * ldq + h * HD + sgm * 8;
      qv = *reinterpret_cast<const uint4*>(rowp);
      const uint4 kv = *reinterpret_cast<const uint4*>(rowp + D);
      vv = *reinterpret_cast<const uint4*>(rowp + 2 * D);
      *reinterpret_cast<uint4*>(Kx + (bh * L + t0 + t) * DK + sgm * 8) = kv;
    }
    *reinterpret_cast<uint4*>(sq + (size_t)t * QP + sgm * 8) = qv;
    *reinterpret_cast<uint4*>(sv + (size_t)t * QP + sgm * 8) = vv;
  }
  // Rcat = [Rh ; Rw] as fp16 hi / lo (rows >= NR zero)
  for (int i = tid; i < NRP * HD; i += 256) {
    const int r = i / HD, d = i % HD;
    float v = 0.f;
    if (r < 2 * S - 1) v = relh[(size_t)r * HD + d];
    else if (r < NR) v = relw[(size_t)(r - (2 * S - 1)) * HD + d];
    const __half hi = __float2half_rn(v);
    sRh[(size_t)r * QP + d] = hi;
    sRl[(size_t)r * QP + d] = __float2half_rn(v - __half2float(hi));
  }
  // K' extension: one-hots of (ky, kx) + zero padding, 16 B at a time
  const int EXT = DK - HD;
  for (int i = tid; i < TC * (EXT / 8); i += 256) {
    const int t = i / (EXT / 8), e0 = (i % (EXT / 8)) * 8;
    if (t < nt) {
      const int tt = t0 + t, ty = tt / S, tx = tt % S;
      __half hv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = e0 + j;
        hv[j] = __float2half_rn((e == ty || e == S + tx) ? 1.f : 0.f);
      }
      *reinterpret_cast<uint4*>(Kx + (bh * L + tt) * DK + HD + e0) = *reinterpret_cast<uint4*>(hv);
    }
  }
  __syncthreads();
  // ---- phase 2: V^T -- thread = (d, group of 8 consecutive tokens) -> one 16 B store
  {
    const int ngrp = (TC + 7) / 8;
    for (int i = tid; i < HD * ngrp; i += 256) {
      const int g = i / HD, d = i % HD;  // consecutive threads -> consecutive d: conflict-free shared-memory reads
      __half hv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int t = g * 8 + j;
        hv[j] = (t < nt) ? sv[(size_t)t * QP + d] : __float2half_rn(0.f);
      }
      if (t0 + g * 8 < Lkp) *reinterpret_cast<uint4*>(Vt + (bh * HD + d) * Lkp + t0 + g * 8) = *reinterpret_cast<uint4*>(hv);
    }
    if (t0 + TC >= L) {   // tile padding of V^T (keys in [L, Lkp)) beyond the last written group: zeros, written by the last chunk
      const int first = ((L - t0 + 7) / 8) * 8 + t0;
      for (int i = tid; i < HD * max(0, Lkp - first); i += 256) {
        const int d = i / (Lkp - first), t = first + i % (Lkp - first);
        Vt[(bh * HD + d) * Lkp + t] = __float2half_rn(0.f);
      }
    }
  }
  __syncthreads();   // sv is dead from here: T goes over it
  // ---- phase 3: T[TC x NRP] = q . Rcat^T on the tensor cores (mma.sync m16n8k16, fp32 accumulate, Rcat = hi + lo)
  {
    const int mt = TC / 16, ntile = NRP / 8;
    for (int m = warp; m < mt; m += 8) {
      uint32_t af[KS][4];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
        ldmatrix_x4(af[ks], sq + (size_t)(m * 16 + (lane & 15)) * QP + ks * 16 + (lane >> 4) * 8);
      for (int n = 0; n < ntile; ++n) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t bh_[2], bl_[2];
          const size_t off = (size_t)(n * 8 + (lane & 7)) * QP + ks * 16 + ((lane >> 3) & 1) * 8;
          ldmatrix_x2(bh_, sRh + off);
          ldmatrix_x2(bl_, sRl + off);
          mma_16816(acc, af[ks], bh_);
          mma_16816(acc, af[ks], bl_);
        }
        const int r0 = m * 16 + (lane >> 2), c0 = n * 8 + 2 * (lane & 3);
        *reinterpret_cast<__half2*>(sT + (size_t)r0 * TP + c0) = __floats2half2_rn(acc[0], acc[1]);
        *reinterpret_cast<__half2*>(sT + (size_t)(r0 + 8) * TP + c0) = __floats2half2_rn(acc[2], acc[3]);
      }
    }
  }
  __syncthreads();
  // ---- phase 4: Q' rows, 16 B per thread-step: [q*scale (HD) | T[t][ty - j + S-1] (S) | T[t][(2S-1) + tx - j + S-1] (S) | 0]
  {
    const int cpr = DK / 8;  // 16-byte chunks per row
    for (int i = tid; i < TC * cpr; i += 256) {
      const int t = i / cpr, c8 = (i % cpr) * 8;
      if (t >= nt) continue;
      const int tt = t0 + t, ty = tt / S, tx = tt % S;
      __half hv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = c8 + j;
        __half v;
        if (col < HD) v = __float2half_rn(__half2float(sq[(size_t)t * QP + col]) * scale);
        else if (col < HD + S) v = sT[(size_t)t * TP + (ty - (col - HD) + S - 1)];
        else if (col < HD + 2 * S) v = sT[(size_t)t * TP + (2 * S - 1) + (tx - (col - HD - S) + S - 1)];
        else v = __float2half_rn(0.f);
        hv[j] = v;
      }
      *reinterpret_cast<uint4*>(Qx + (bh * L + tt) * DK + c8) = *reinterpret_cast<uint4*>(hv);
    }
  }
}

static bool attn_prep2_enabled() {
  static const int on = [] { const char* e = std::getenv("SAMPT_ATTN_PREP2"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();
  return on != 0;
}

// returns 1 when the shape is not covered (caller falls back to attn_prep_kernel), 0 on success, < 0 on error
int attn_prep2(Ctx* c, cudaStream_t st, const __half* qkv, int ldq, const float* relh, const float* relw, __half* Qx, __half* Kx,
               __half* Vt, int nwb, int nheads, int S, int Lkp, int DK, int D, int HD, float scale) {
  if (!attn_prep2_enabled()) return 1;
  if (!(HD == 80 || HD == 64) || DK % 8 != 0 || DK < HD + 2 * S || (DK - HD) % 8 != 0 || Lkp % 8 != 0) return 1;
  const int L = S * S;
  const int TC = (L <= 256) ? ((L + 15) / 16) * 16 : 128;        // a whole 14x14 window (208 rows) or 128 tokens of the global grid
  const int NRP = ((2 * (2 * S - 1) + 7) / 8) * 8;
  const int QP = HD + 8, TP = NRP + 8;
  // shared memory: sq + max(sv, sT) + Rcat hi + lo
  const size_t sv_or_t = std::max((size_t)TC * QP, (size_t)TC * TP);
  const size_t smem = ((size_t)TC * QP + sv_or_t + 2 * (size_t)NRP * QP) * sizeof(__half);
  if (smem > 200 * 1024) return 1;
  dim3 grid(cdiv(L, TC), nheads, nwb);
  if (HD == 80) {
    SAMPT_TRY(ensure_func_smem(c, "attn_prep2_kernel<80>", attn_prep2_kernel<80>, 200 * 1024));
    attn_prep2_kernel<80><<<grid, 256, smem, st>>>(qkv, ldq, relh, relw, Qx, Kx, Vt, S, L, Lkp, DK, D, nheads, scale, TC, NRP);
  } else {
    SAMPT_TRY(ensure_func_smem(c, "attn_prep2_kernel<64>", attn_prep2_kernel<64>, 200 * 1024));
    attn_prep2_kernel<64><<<grid, 256, smem, st>>>(qkv, ldq, relh, relw, Qx, Kx, Vt, S, L, Lkp, DK, D, nheads, scale, TC, NRP);
  }
  c->launches++;
  SAMPT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sampt
