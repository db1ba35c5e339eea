// Tensor-core GEMM for the SAM ViT encoder:  C[M,N] = epilogue( A[M,K] . B[N,K]^T ),  fp16/bf16 operands, fp32 accumulate.
//
// sm_100a design (one CTA per SM, persistent over 128x256 output tiles):
//   warp 0      TMA producer   : cp.async.bulk.tensor 2-D boxes (64 halves x rows, 128B swizzle) into a 4-stage smem ring
//   warp 1      MMA issuer     : one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=256, K=16), fp32
//                                accumulators in TMEM (2 x 256 columns, double buffered so the epilogue of tile i overlaps
//                                the main loop of tile i+1); tcgen05.commit releases smem stages / publishes accumulators
//   warps 2..5  epilogue       : tcgen05.ld (32 lanes x 32 columns) -> bias / GELU / residual / fp16|fp32|split store
//
// "Split" precision (accuracy dial): an operand x is carried as fp16 hi + fp16 lo
// (lo = fp16(x - hi)); the K loop then runs over up to three segments  A_hi.B_hi + A_lo.B_hi + A_hi.B_lo  accumulating
// into the same TMEM tile.  Segments are described by column offsets into the A / B matrices, so the kernel is the same.
#include "common.cuh"
#include "tc_common.cuh"
#include "kernels.cuh"
#include "tc_api.cuh"
#include "../../include/sampt_b200.h"

namespace sampt {

using namespace tc;

constexpr int G_BM = 128, G_BN = 256, G_BK = 64, G_STAGES = 4;
constexpr int G_A_BYTES = G_BM * G_BK * 2;  // 16 KB
constexpr int G_B_BYTES = G_BN * G_BK * 2;  // 32 KB
constexpr int G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr int G_SMEM_BYTES = G_STAGES * G_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int G_THREADS = 192;

__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
               GemmSeg seg, GemmEpi ep) {
  if (ep.skip != nullptr && *ep.skip != 0) return;   // uniform over the grid: nothing has been allocated yet
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G_STAGES * G_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + G_STAGES;
  uint64_t* tfull_bar = empty_bar + G_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + G_BM - 1) / G_BM, n_tiles = (N + G_BN - 1) / G_BN;
  const int num_tiles = m_tiles * n_tiles;
  const int kb_per_seg = K / G_BK;
  const int num_kb = kb_per_seg * seg.nseg;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < G_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
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
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % G_STAGES;
          const uint32_t ph = (it / G_STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          const int sg = kb / kb_per_seg, kk = (kb % kb_per_seg) * G_BK;
          uint8_t* sa = smem + s * G_STAGE_BYTES;
          uint8_t* sb = sa + G_A_BYTES;
          mbar_expect_tx(&full_bar[s], G_STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[s], seg.a_off[sg] + kk, m_blk * G_BM);
          tma_load_2d(sb, &tmB, &full_bar[s], seg.b_off[sg] + kk, n_blk * G_BN);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
        // narrow problems (N < 256, e.g. 64-channel convolutions) issue a narrower UMMA instead of multiplying zero padding
        const int n_rem = N - (tile % n_tiles) * G_BN;
        const int mma_n = n_rem >= G_BN ? G_BN : ((n_rem + 15) / 16) * 16;
        const uint32_t idesc = make_idesc_f16(G_BM, mma_n, ep.is_bf16);
        const int acc = tl & 1;
        const uint32_t aph = (tl >> 1) & 1;
        mbar_wait(&tempty_bar[acc], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * G_BN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % G_STAGES;
          const uint32_t ph = (it / G_STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * G_STAGE_BYTES);
          const uint64_t adesc = make_smem_desc_sw128(sa);
          const uint64_t bdesc = make_smem_desc_sw128(sa + G_A_BYTES);
#pragma unroll
          for (int k = 0; k < G_BK / 16; ++k) {
            // advance 16 halves = 32 B inside the 128 B swizzle atom: +2 in the (addr >> 4) field
            umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
    const int q = warp & 3;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tl) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int acc = tl & 1;
      const uint32_t aph = (tl >> 1) & 1;
      mbar_wait(&tfull_bar[acc], aph);
      tc_fence_after();
      const int m = m_blk * G_BM + q * 32 + lane;
      const bool row_ok = m < M;
      long long drow = m;
      if (ep.rowmap && row_ok) drow = ep.rowmap[m];
      const bool store_ok = row_ok && drow >= 0;
#pragma unroll 1
      for (int ch = 0; ch < G_BN / 32; ++ch) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * G_BN + ch * 32), r);
        tmem_ld_wait();
        const int n0 = n_blk * G_BN + ch * 32;
        if (n0 >= N) continue;  // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
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
          if (ep.split_off > 0) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<uint4*>(o + ep.split_off + 2 * j) = make_uint4(lo[j], lo[j + 1], lo[j + 2], lo[j + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int gemm_tc(Ctx* c, cudaStream_t st, const void* A, int lda, const void* B, int ldb, int M, int N, int K, const GemmSeg& seg,
            const GemmEpi& ep) {
  SAMPT_CHECK(K % G_BK == 0 && K > 0, "gemm_tc: K (%d) must be a positive multiple of %d", K, G_BK);
  SAMPT_CHECK(N % 32 == 0, "gemm_tc: N (%d) must be a multiple of 32", N);
  SAMPT_CHECK(seg.nseg >= 1 && seg.nseg <= 3, "gemm_tc: nseg out of range");
  SAMPT_CHECK((ep.out16 != nullptr) != (ep.out32 != nullptr), "gemm_tc: exactly one of out16/out32 must be set");
  SAMPT_CHECK(ep.ldc % 8 == 0, "gemm_tc: ldc must be a multiple of 8");
  if (gemm_tc2_applicable(M, N, K, ep)) {
    SAMPT_CHECK(!(seg.f8[0] | seg.f8[1] | seg.f8[2]) || K % 128 == 0, "gemm_tc: e4m3 segments need K %% 128 == 0 (K = %d)", K);
    return gemm_tc2(c, st, A, lda, B, ldb, M, N, K, seg, ep);
  }
  SAMPT_CHECK(!(seg.f8[0] | seg.f8[1] | seg.f8[2]) && !ep.out_f8 && ep.acc_scale == nullptr,
              "gemm_tc: the fp8-corrected GEMM runs on the CTA-pair kernel only (M >= 256, N %% 256 == 0, K %% 128 == 0)");
  SAMPT_TRY(ensure_func_smem(c, "gemm_tc_kernel", gemm_tc_kernel, G_SMEM_BYTES));
  CUtensorMap tmA, tmB;
  // the A/B matrices may carry several K segments side by side (hi | lo): inner extent = lda / ldb
  SAMPT_TRY(make_tmap_2d_f16(&tmA, A, (uint64_t)lda, (uint64_t)M, (uint64_t)lda * 2, G_BK, G_BM));
  SAMPT_TRY(make_tmap_2d_f16(&tmB, B, (uint64_t)ldb, (uint64_t)N, (uint64_t)ldb * 2, G_BK, G_BN));
  const int m_tiles = (M + G_BM - 1) / G_BM, n_tiles = (N + G_BN - 1) / G_BN;
  int grid = std::min(m_tiles * n_tiles, c->num_sms);
  gemm_tc_kernel<<<grid, G_THREADS, G_SMEM_BYTES, st>>>(tmA, tmB, M, N, K, seg, ep);
  c->launches++;
  SAMPT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sampt

using namespace sampt;

// Unit-test / building-block entry: C = act(A B^T + bias), fp16 operands (bf16 if is_bf16), fp32 accumulation.
// precision: 1 = single pass; 2 = B (weights) carried as hi|lo (B is [N, 2K], lo at column K), A plain; 3 = both carried as hi|lo.
extern "C" int sampt_gemm_f16(sampt_ctx* ctx, const void* A, int lda, const void* B, int ldb, int M, int N, int K, int precision,
                              int is_bf16, const float* bias, int act, void* out16, float* out32, const float* resid, int ldc,
                              int split_off, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  GemmSeg seg{};
  seg.nseg = precision;
  SAMPT_CHECK(precision >= 1 && precision <= 3, "precision must be 1, 2 or 3");
  // precision 1: A.B ; 2: A.(B_hi + B_lo)  (weights carried as hi|lo) ; 3: A_hi.B_hi + A_lo.B_hi + A_hi.B_lo
  if (precision == 2) {
    seg.a_off[0] = 0; seg.b_off[0] = 0;
    seg.a_off[1] = 0; seg.b_off[1] = K;
  } else {
    seg.a_off[0] = 0; seg.b_off[0] = 0;
    seg.a_off[1] = K; seg.b_off[1] = 0;
    seg.a_off[2] = 0; seg.b_off[2] = K;
  }
  GemmEpi ep{};
  ep.out16 = reinterpret_cast<__half*>(out16);
  ep.out32 = out32;
  ep.resid = resid;
  ep.bias = bias;
  ep.rowmap = nullptr;
  ep.ldc = ldc;
  ep.act = act;
  ep.split_off = split_off;
  ep.is_bf16 = is_bf16;
  return gemm_tc(c, reinterpret_cast<cudaStream_t>(stream), A, lda, B, ldb, M, N, K, seg, ep);
}

// fp8-corrected split GEMM (tc_api.cuh, "precision 6"): C = act((A_hi.B_hi + A_lo8.B_hi8 + A_hi8.B_lo8) * acc_scale + bias).
// A [M, 2K] and B [N, 2K] fp16 units in the layouts of tc_api.cuh (sampt_split_f8c builds A; the host builds B once per weight).
extern "C" int sampt_gemm_f8c(sampt_ctx* ctx, const void* A, const void* B, int M, int N, int K, const float* acc_scale_dev,
                              const float* bias, int act, void* out16, float* out32, const float* resid, int ldc, int split_off,
                              int out_f8, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  SAMPT_CHECK(gemm_f8c_applicable(M, N, K), "sampt_gemm_f8c: needs M >= 256, N %% 256 == 0, K %% 128 == 0 (got %d x %d x %d)", M, N, K);
  GemmEpi ep{};
  ep.out16 = reinterpret_cast<__half*>(out16);
  ep.out32 = out32;
  ep.resid = resid;
  ep.bias = bias;
  ep.ldc = ldc;
  ep.act = act;
  ep.split_off = split_off;
  ep.out_f8 = out_f8;
  ep.acc_scale = acc_scale_dev;
  return gemm_tc(c, reinterpret_cast<cudaStream_t>(stream), A, 2 * K, B, 2 * K, M, N, K, make_seg_f8(K), ep);
}

// x [M, K] fp32 -> the A operand of sampt_gemm_f8c: [fp16(x) | e4m3((x - fp16(x)) * 2^12) | e4m3(x * 2^-3)], 2K fp16 units per row
extern "C" int sampt_split_f8c(sampt_ctx* ctx, const float* x, int M, int K, void* out, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  return ln_rows(c, reinterpret_cast<cudaStream_t>(stream), x, K, nullptr, nullptr, nullptr, 0.f, reinterpret_cast<__half*>(out), 2 * K, K,
                 M, K, 0, 1);
}
