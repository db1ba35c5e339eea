// ON by default (SAMPT_GEMM_2CTA=0 selects gemm_tc_kernel); validated on hardware in round 2 (gpurun_out/exp_gemm_cta_pair.log):
// CTA-pair variant of gemm_tc_kernel:  C[M,N] = epilogue( A[M,K] . B[N,K]^T ) with tcgen05.mma.cta_group::2.
//
// Why: the 1-CTA kernel (M128 x N256 x K16 per instruction) reads 12 KB of shared memory per 128 tensor-pipe cycles while TMA
// writes the same 12 KB — the shared-memory port is the limiter (ncu: tensor pipe 81 % active, l1tex 72 %).  A CTA pair on one
// TPC computes a 256 x 256 tile; each CTA stages its own 128 rows of A and only HALF of the B tile (128 of the 256 columns),
// the tensor cores of both SMs read the two B halves through the pair link.  Per CTA and k-block: 16 KB + 16 KB instead of
// 16 KB + 32 KB, which also makes room for 6 pipeline stages instead of 4.
//
// Protocol (CUTLASS sm100 2-SM convention; PTX strings as in cute/arch/copy_sm100_tma.hpp, mma_sm100_umma.hpp, cutlass/arch/barrier.h):
//   * cluster (2,1,1); rank 0 = leader.  TMEM is allocated with cta_group::2 by warp 1 of both CTAs.
//   * full[s]   lives in the LEADER: the leader's producer arms it with the bytes of BOTH CTAs; both producers issue
//               cp.async.bulk.tensor...cta_group::2 with the barrier address' peer bit cleared (-> leader's barrier).
//   * empty[s]  one per CTA, released by tcgen05.commit.cta_group::2...multicast::cluster (mask 0b11) from the leader's MMA thread.
//   * tfull[a]  one per CTA (multicast commit): accumulator `a` is complete; each CTA's epilogue drains its own 128 TMEM lanes.
//   * tempty[a] lives in the leader, 8 arrivals (4 epilogue warps x 2 CTAs; the peer arrives remotely through shared::cluster).
// The epilogue is the one of gemm_tc_kernel.  Requires N % 256 == 0 (the ViT's linear layers: 3840, 1280, 5120, 256).
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"
#include "kernels.cuh"
#include "tc_api.cuh"

namespace sampt {

using namespace tc;

constexpr int P_BM = 128;          // rows per CTA (256 per pair)
constexpr int P_BN = 256;          // tile columns (128 staged per CTA)
constexpr int P_BK = 64;
constexpr int P_A_BYTES = P_BM * P_BK * 2;          // 16 KB
constexpr int P_B_BYTES = (P_BN / 2) * P_BK * 2;    // 16 KB (this CTA's half of the B tile)
constexpr int P_STAGE_BYTES = P_A_BYTES + P_B_BYTES;
// pipeline depth: 6 stages (192 KB; default) or 7 (224 KB, SAMPT_GEMM_STAGES=7).  A stage is refilled every STAGES x 512 tensor
// cycles, which has to cover commit -> producer wake-up -> TMA issue -> L2 round trip -> full barrier -> MMA issue (~1.4 us).
// Measured on one box (gpurun_out/c10_bench*.log): 7 stages 933 / 931 ms per C2 step, 6 stages 924 ms -- the seventh stage buys
// nothing and takes the shared memory that lets small decode kernels share the SM
constexpr int p_smem_bytes(int stages) { return stages * P_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/; }
constexpr int P_THREADS = 192;
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;     // clears the CTA-pair peer bit of a shared::cluster address (-> even CTA)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load whose completion bytes are credited to the LEADER CTA's mbarrier at the same shared-memory offset
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// same tile shape with e4m3 operands: K = 32 per instruction (the same 32 bytes per row), twice the fp16 rate; the instruction
// descriptor is bit-identical (format code 0 = F16 for kind::f16, E4M3 for kind::f8f6f4)
__device__ __forceinline__ void umma_f8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// arrive on the LEADER's barrier (from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}

template <int P_STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K, GemmSeg seg,
                GemmEpi ep) {
  if (ep.skip != nullptr && *ep.skip != 0) return;   // uniform over the grid (both CTAs of every pair): nothing allocated yet
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + P_STAGES * P_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + P_STAGES;
  uint64_t* tfull_bar = empty_bar + P_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int m_tiles = (M + 2 * P_BM - 1) / (2 * P_BM), n_tiles = N / P_BN;
  const int num_tiles = m_tiles * n_tiles;
  // k-blocks of 128 bytes per operand row: 64 fp16 or 128 e4m3 elements
  int seg_kb[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) seg_kb[i] = i < seg.nseg ? (seg.f8[i] ? K / (2 * P_BK) : K / P_BK) : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < P_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit / pair TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs: own A rows, own half of B)
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int sg = 0; sg < seg.nseg; ++sg) {
          for (int kb = 0; kb < seg_kb[sg]; ++kb, ++it) {
            const int s = it % P_STAGES;
            const uint32_t ph = (it / P_STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            const int kk = kb * P_BK;   // in fp16 units of the tensor map (an e4m3 block is the same 128 bytes)
            uint8_t* sa = smem + s * P_STAGE_BYTES;
            uint8_t* sb = sa + P_A_BYTES;
            if (leader) mbar_expect_tx(&full_bar[s], 2 * P_STAGE_BYTES);   // bytes of both CTAs land on the leader's barrier
            tma_load_2d_pair(sa, &tmA, &full_bar[s], seg.a_off[sg] + kk, m_blk * 2 * P_BM + (int)rank * P_BM);
            tma_load_2d_pair(sb, &tmB, &full_bar[s], seg.b_off[sg] + kk, n_blk * P_BN + (int)rank * (P_BN / 2));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only)
    if (lane == 0 && leader) {
      const uint32_t idesc = make_idesc_f16(2 * P_BM, P_BN, ep.is_bf16);
      uint32_t it = 0, tl = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++tl) {
        const int acc = tl & 1;
        const uint32_t aph = (tl >> 1) & 1;
        mbar_wait(&tempty_bar[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * P_BN);
        uint32_t first = 0;   // 0 until the first MMA of the tile has been issued (it overwrites the accumulator)
        for (int sg = 0; sg < seg.nseg; ++sg) {
          const bool f8 = seg.f8[sg] != 0;
          for (int kb = 0; kb < seg_kb[sg]; ++kb, ++it) {
            const int s = it % P_STAGES;
            const uint32_t ph = (it / P_STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * P_STAGE_BYTES);
            const uint64_t adesc = make_smem_desc_sw128(sa);
            const uint64_t bdesc = make_smem_desc_sw128(sa + P_A_BYTES);
            if (f8) {
#pragma unroll
              for (int k = 0; k < P_BK / 16; ++k) {   // 4 x K=32 bytes
                umma_f8_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, first);
                first = 1;
              }
            } else {
#pragma unroll
              for (int k = 0; k < P_BK / 16; ++k) {   // 4 x K=16 halves
                umma_f16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, first);
                first = 1;
              }
            }
            umma_commit_pair(&empty_bar[s]);
          }
        }
        umma_commit_pair(&tfull_bar[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
    const int q = warp & 3;
    uint32_t tl = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++tl) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int acc = tl & 1;
      const uint32_t aph = (tl >> 1) & 1;
      mbar_wait(&tfull_bar[acc], aph);
      tc_fence_after();
      const int m = m_blk * 2 * P_BM + (int)rank * P_BM + q * 32 + lane;
      const bool row_ok = m < M;
      long long drow = m;
      if (ep.rowmap && row_ok) drow = ep.rowmap[m];
      const bool store_ok = row_ok && drow >= 0;
      const float acc_scale = ep.acc_scale ? __ldg(ep.acc_scale) : 1.0f;
      // (loading the next 32-column chunk from TMEM while this one is processed was tried: no change of the isolated GEMM time,
      //  the epilogue is hidden behind the MMAs of the other accumulator)
#pragma unroll 1
      for (int ch = 0; ch < P_BN / 32; ++ch) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * P_BN + ch * 32), r);
        tmem_ld_wait();
        const int n0 = n_blk * P_BN + ch * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * acc_scale;
        if (ep.bias) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __ldg(ep.bias + n0 + j);
        }
        if (ep.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        } else if (ep.act == 3) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_tanh(v[j]);
        }
        if (!store_ok) {
          // nothing to write for this row (tail of M, or a padding row dropped by rowmap)
        } else if (ep.out32) {
          float* o = ep.out32 + (size_t)drow * ep.ldc + n0;
          if (ep.resid) {
            const long long rrow = ep.resid_mod > 0 ? (drow % ep.resid_mod) : drow;
            const float* rs = ep.resid + (size_t)rrow * ep.ldc + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 t = *reinterpret_cast<const float4*>(rs + j);
              v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          __half* o = ep.out16 + (size_t)drow * ep.ldc + n0;
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (ep.is_bf16) {
              __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              hi[j] = *reinterpret_cast<uint32_t*>(&h);
              lo[j] = 0;
            } else {
              __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
              hi[j] = *reinterpret_cast<uint32_t*>(&h);
              float2 hf = __half22float2(h);
              __half2 l = __floats2half2_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
              lo[j] = *reinterpret_cast<uint32_t*>(&l);
            }
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<uint4*>(o + 2 * j) = make_uint4(hi[j], hi[j + 1], hi[j + 2], hi[j + 3]);
          if (ep.split_off > 0 && ep.out_f8) {
            // fp8 correction operands of the next GEMM (tc_api.cuh): remainder * 2^12 and value * 2^-3 as e4m3 bytes
            uint8_t* ob = reinterpret_cast<uint8_t*>(ep.out16 + (size_t)drow * ep.ldc + ep.split_off) + n0;
            uint32_t l8[8], h8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const __half2 ha = *reinterpret_cast<__half2*>(&hi[2 * j]), hb = *reinterpret_cast<__half2*>(&hi[2 * j + 1]);
              const float2 fa = __half22float2(ha), fb = __half22float2(hb);
              l8[j] = cvt_e4m3x4((v[4 * j] - fa.x) * F8_LO_SCALE, (v[4 * j + 1] - fa.y) * F8_LO_SCALE,
                                 (v[4 * j + 2] - fb.x) * F8_LO_SCALE, (v[4 * j + 3] - fb.y) * F8_LO_SCALE);
              h8[j] = cvt_e4m3x4(v[4 * j] * F8_HI_SCALE, v[4 * j + 1] * F8_HI_SCALE, v[4 * j + 2] * F8_HI_SCALE, v[4 * j + 3] * F8_HI_SCALE);
            }
            *reinterpret_cast<uint4*>(ob) = make_uint4(l8[0], l8[1], l8[2], l8[3]);
            *reinterpret_cast<uint4*>(ob + 16) = make_uint4(l8[4], l8[5], l8[6], l8[7]);
            *reinterpret_cast<uint4*>(ob + ep.split_off) = make_uint4(h8[0], h8[1], h8[2], h8[3]);
            *reinterpret_cast<uint4*>(ob + ep.split_off + 16) = make_uint4(h8[4], h8[5], h8[6], h8[7]);
          } else if (ep.split_off > 0) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<uint4*>(o + ep.split_off + 2 * j) = make_uint4(lo[j], lo[j + 1], lo[j + 2], lo[j + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be draining TMEM / receiving multicast arrivals
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

bool gemm_tc2_applicable(int M, int N, int K, const GemmEpi& ep) {
  static const int enabled = [] { const char* e = std::getenv("SAMPT_GEMM_2CTA"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();   // validated on hardware in round 2: on unless =0
  (void)ep;
  return enabled && N % P_BN == 0 && K % P_BK == 0 && M >= 2 * P_BM;
}
// the fp8-corrected segments additionally need whole 128-element k-blocks
bool gemm_f8c_applicable(int M, int N, int K) {
  GemmEpi ep{};
  return gemm_tc2_applicable(M, N, K, ep) && K % (2 * P_BK) == 0;
}

int gemm_tc2(Ctx* c, cudaStream_t st, const void* A, int lda, const void* B, int ldb, int M, int N, int K, const GemmSeg& seg,
             const GemmEpi& ep) {
  static const int stages = [] { const char* e = std::getenv("SAMPT_GEMM_STAGES"); return (e != nullptr && e[0] == '7') ? 7 : 6; }();
  CUtensorMap tmA, tmB;
  SAMPT_TRY(make_tmap_2d_f16(&tmA, A, (uint64_t)lda, (uint64_t)M, (uint64_t)lda * 2, P_BK, P_BM));
  SAMPT_TRY(make_tmap_2d_f16(&tmB, B, (uint64_t)ldb, (uint64_t)N, (uint64_t)ldb * 2, P_BK, P_BN / 2));
  const int m_tiles = (M + 2 * P_BM - 1) / (2 * P_BM), n_tiles = N / P_BN;
  const int pairs = std::min(m_tiles * n_tiles, c->num_sms / 2);
  if (stages == 7) {   // cluster dims are static (2,1,1)
    SAMPT_TRY(ensure_func_smem(c, "gemm_tc2_kernel<7>", gemm_tc2_kernel<7>, p_smem_bytes(7)));
    gemm_tc2_kernel<7><<<2 * pairs, P_THREADS, p_smem_bytes(7), st>>>(tmA, tmB, M, N, K, seg, ep);
  } else {
    SAMPT_TRY(ensure_func_smem(c, "gemm_tc2_kernel<6>", gemm_tc2_kernel<6>, p_smem_bytes(6)));
    gemm_tc2_kernel<6><<<2 * pairs, P_THREADS, p_smem_bytes(6), st>>>(tmA, tmB, M, N, K, seg, ep);
  }
  c->launches++;
  SAMPT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sampt
