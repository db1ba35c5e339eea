// FP32 CUDA-core GEMM  Y[M,N] = act(X[M,K] . W[N,K]^T + bias) (+ residual)   -- the strict-fp32 path.
//
// Used where the reference's numerics must be kept at float32 (PIPS MLP-Mixer: the 1e-3 px tolerance with a
// 6-iteration feedback loop, SURVEY §7 "parity under chaos"; SAM prompt/mask decoder: 12 mask->box->mask
// refinement iterations).  Tensor-core GEMMs (tcgen05) live in gemm_tc.cu and serve the ViT encoder.
#include "common.cuh"
#include "kernels.cuh"

namespace sampt {

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_nt_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
                const float* __restrict__ bias, const float* residual, int ldr, float* Y, int ldy,
                int M, int N, int K, int act, const int* skip) {
  if (skip != nullptr && *skip != 0) return;
  constexpr int NT = (BM / TM) * (BN / TN);
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);

  constexpr int A_F4 = BM * BK / 4, B_F4 = BN * BK / 4;
  constexpr int A_PER = (A_F4 + NT - 1) / NT, B_PER = (B_F4 + NT - 1) / NT;
  float4 ra[A_PER], rb[B_PER];

  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int idx = tid + i * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < A_F4) {
        int r = idx / (BK / 4), c = (idx % (BK / 4)) * 4;
        int gm = m0 + r, gk = k0 + c;
        if (gm < M && gk < K) v = *reinterpret_cast<const float4*>(X + (size_t)gm * ldx + gk);
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      int idx = tid + i * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < B_F4) {
        int r = idx / (BK / 4), c = (idx % (BK / 4)) * 4;
        int gn = n0 + r, gk = k0 + c;
        if (gn < N && gk < K) v = *reinterpret_cast<const float4*>(W + (size_t)gn * ldw + gk);
      }
      rb[i] = v;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int idx = tid + i * NT;
      if (idx < A_F4) {
        int r = idx / (BK / 4), c = (idx % (BK / 4)) * 4;
        As[buf][c + 0][r] = ra[i].x; As[buf][c + 1][r] = ra[i].y; As[buf][c + 2][r] = ra[i].z; As[buf][c + 3][r] = ra[i].w;
      }
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      int idx = tid + i * NT;
      if (idx < B_F4) {
        int r = idx / (BK / 4), c = (idx % (BK / 4)) * 4;
        Bs[buf][c + 0][r] = rb[i].x; Bs[buf][c + 1][r] = rb[i].y; Bs[buf][c + 2][r] = rb[i].z; Bs[buf][c + 3][r] = rb[i].w;
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN + j]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int gn = n0 + tx * TN + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      if (act == 1) v = gelu_erf(v);
      else if (act == 2) v = fmaxf(v, 0.f);
      if (residual) v += residual[(size_t)gm * ldr + gn];
      Y[(size_t)gm * ldy + gn] = v;
    }
  }
}

// Small-M variant (M <= MAXM rows, e.g. the prompt tokens of the mask decoder or the N*S rows of the PIPS mixer):
// weight-bandwidth bound, so the grid is spread over the N (output column) axis and every weight row is read exactly once,
// coalesced.  One warp = one output column x all M rows; lanes split K (float4 per lane per step); X is staged through
// shared memory in K chunks of 128; a warp-shuffle reduction per row finishes the dot products.
template <int MAXM>
__global__ void __launch_bounds__(256)
sgemm_smallm_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw, const float* __restrict__ bias,
                    const float* residual, int ldr, float* Y, int ldy, int M, int N, int K, int act, const int* skip) {
  if (skip != nullptr && *skip != 0) return;
  constexpr int KC = 128;
  __shared__ __align__(16) float xs[MAXM][KC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  float acc[MAXM];
#pragma unroll
  for (int m = 0; m < MAXM; ++m) acc[m] = 0.f;
  for (int k0 = 0; k0 < K; k0 += KC) {
    __syncthreads();
    for (int i = threadIdx.x; i < MAXM * (KC / 4); i += 256) {
      int m = i / (KC / 4), c = (i % (KC / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M && k0 + c < K) v = *reinterpret_cast<const float4*>(X + (size_t)m * ldx + k0 + c);
      *reinterpret_cast<float4*>(&xs[m][c]) = v;
    }
    __syncthreads();
    if (n < N) {
      const int kk = k0 + lane * 4;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kk < K) w = *reinterpret_cast<const float4*>(W + (size_t)n * ldw + kk);
#pragma unroll
      for (int m = 0; m < MAXM; ++m) {
        float4 x = *reinterpret_cast<const float4*>(&xs[m][lane * 4]);
        acc[m] = fmaf(w.x, x.x, acc[m]);
        acc[m] = fmaf(w.y, x.y, acc[m]);
        acc[m] = fmaf(w.z, x.z, acc[m]);
        acc[m] = fmaf(w.w, x.w, acc[m]);
      }
    }
  }
  if (n >= N) return;
#pragma unroll
  for (int m = 0; m < MAXM; ++m) {
    float v = warp_sum(acc[m]);
    if (lane == (m & 31) && m < M) {
      if (bias) v += bias[n];
      if (act == 1) v = gelu_erf(v);
      else if (act == 2) v = fmaxf(v, 0.f);
      if (residual) v += residual[(size_t)m * ldr + n];
      Y[(size_t)m * ldy + n] = v;
    }
  }
}

// M <= 64 rows (PIPS mixer: N*S = 64 rows for 8 points): register-tiled so that shared-memory traffic does not bound it.
// CTA = 8 warps = 4 row groups (16 rows) x 2 column groups (8 columns) -> 64 x 16 outputs; lanes split K (float4 per
// lane per step, X staged in smem in chunks of 128, W read coalesced straight from global/L1); per k-step a lane does
// 16 LDS.128 + 8 LDG.128 for 512 FMAs; a shuffle tree finishes the 128 dot products of the warp.
__global__ void __launch_bounds__(256)
sgemm_m64_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw, const float* __restrict__ bias,
                 const float* residual, int ldr, float* Y, int ldy, int M, int N, int K, int act, const int* skip) {
  if (skip != nullptr && *skip != 0) return;
  constexpr int KC = 128;
  __shared__ __align__(16) float xs[64][KC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mg = warp & 3, cg = warp >> 2;
  const int nbase = blockIdx.x * 16 + cg * 8;
  float acc[16][8];
#pragma unroll
  for (int i = 0; i < 16; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += KC) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * (KC / 4); i += 256) {
      int m = i / (KC / 4), c = (i % (KC / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M && k0 + c < K) v = *reinterpret_cast<const float4*>(X + (size_t)m * ldx + k0 + c);
      *reinterpret_cast<float4*>(&xs[m][c]) = v;
    }
    __syncthreads();
    const int kk = k0 + lane * 4;
    float4 w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      w[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (nbase + j < N && kk < K) w[j] = __ldg(reinterpret_cast<const float4*>(W + (size_t)(nbase + j) * ldw + kk));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 x = *reinterpret_cast<const float4*>(&xs[mg * 16 + i][lane * 4]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[i][j] = fmaf(x.x, w[j].x, acc[i][j]);
        acc[i][j] = fmaf(x.y, w[j].y, acc[i][j]);
        acc[i][j] = fmaf(x.z, w[j].z, acc[i][j]);
        acc[i][j] = fmaf(x.w, w[j].w, acc[i][j]);
      }
    }
  }
  // reduce over lanes: after the tree, lane l holds output (i, j) with i*8 + j == l (mod 32) for i*8+j in [32r, 32r+32)
#pragma unroll
  for (int i = 0; i < 16; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = warp_sum(acc[i][j]);
      const int m = mg * 16 + i, n = nbase + j;
      if (lane == ((i * 8 + j) & 31) && m < M && n < N) {
        if (bias) v += bias[n];
        if (act == 1) v = gelu_erf(v);
        else if (act == 2) v = fmaxf(v, 0.f);
        if (residual) v += residual[(size_t)m * ldr + n];
        Y[(size_t)m * ldy + n] = v;
      }
    }
  }
}

int sgemm_nt(Ctx* c, cudaStream_t st, const float* X, int ldx, const float* W, int ldw, const float* bias,
             const float* residual, int ldr, float* Y, int ldy, int M, int N, int K, int act) {
  return sgemm_nt_skip(c, st, X, ldx, W, ldw, bias, residual, ldr, Y, ldy, M, N, K, act, nullptr);
}

int sgemm_nt_skip(Ctx* c, cudaStream_t st, const float* X, int ldx, const float* W, int ldw, const float* bias,
                  const float* residual, int ldr, float* Y, int ldy, int M, int N, int K, int act, const int* skip) {
  SAMPT_CHECK((K % 4) == 0 && (ldx % 4) == 0 && (ldw % 4) == 0, "sgemm_nt: K/ldx/ldw must be multiples of 4 (K=%d ldx=%d ldw=%d)", K, ldx, ldw);
  if (M <= 0 || N <= 0) return 0;
  if (M <= 16) {
    sgemm_smallm_kernel<16><<<cdiv(N, 8), 256, 0, st>>>(X, ldx, W, ldw, bias, residual, ldr, Y, ldy, M, N, K, act, skip);
  } else if (M <= 32) {
    sgemm_smallm_kernel<32><<<cdiv(N, 8), 256, 0, st>>>(X, ldx, W, ldw, bias, residual, ldr, Y, ldy, M, N, K, act, skip);
  } else if (M <= 64) {
    sgemm_m64_kernel<<<cdiv(N, 16), 256, 0, st>>>(X, ldx, W, ldw, bias, residual, ldr, Y, ldy, M, N, K, act, skip);
  } else {
    // tile choice: the largest tile that still yields ~a wave of CTAs on 148 SMs
    const long long t128 = (long long)cdiv(M, 128) * cdiv(N, 64), t64 = (long long)cdiv(M, 64) * cdiv(N, 64);
    if (t128 >= c->num_sms) {
      dim3 grid(cdiv(N, 64), cdiv(M, 128));
      sgemm_nt_kernel<128, 64, 16, 8, 4><<<grid, 256, 0, st>>>(X, ldx, W, ldw, bias, residual, ldr, Y, ldy, M, N, K, act, skip);
    } else if (t64 >= c->num_sms / 2) {
      dim3 grid(cdiv(N, 64), cdiv(M, 64));
      sgemm_nt_kernel<64, 64, 16, 4, 4><<<grid, 256, 0, st>>>(X, ldx, W, ldw, bias, residual, ldr, Y, ldy, M, N, K, act, skip);
    } else {
      dim3 grid(cdiv(N, 32), cdiv(M, 32));
      sgemm_nt_kernel<32, 32, 16, 4, 4><<<grid, 64, 0, st>>>(X, ldx, W, ldw, bias, residual, ldr, Y, ldy, M, N, K, act, skip);
    }
  }
  c->launches++;
  SAMPT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sampt
