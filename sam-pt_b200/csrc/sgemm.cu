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

int sgemm_nt(Ctx* c, cudaStream_t st, const float* X, int ldx, const float* W, int ldw, const float* bias,
             const float* residual, int ldr, float* Y, int ldy, int M, int N, int K, int act) {
  return sgemm_nt_skip(c, st, X, ldx, W, ldw, bias, residual, ldr, Y, ldy, M, N, K, act, nullptr);
}

int sgemm_nt_skip(Ctx* c, cudaStream_t st, const float* X, int ldx, const float* W, int ldw, const float* bias,
                  const float* residual, int ldr, float* Y, int ldy, int M, int N, int K, int act, const int* skip) {
  SAMPT_CHECK((K % 4) == 0 && (ldx % 4) == 0 && (ldw % 4) == 0, "sgemm_nt: K/ldx/ldw must be multiples of 4 (K=%d ldx=%d ldw=%d)", K, ldx, ldw);
  if (M <= 0 || N <= 0) return 0;
  // tile choice: big tiles when there is enough work to fill 148 SMs, else smaller tiles for more CTAs
  long long tiles_big = (long long)cdiv(M, 128) * cdiv(N, 64);
  if (tiles_big >= 2 * c->num_sms) {
    dim3 grid(cdiv(N, 64), cdiv(M, 128));
    sgemm_nt_kernel<128, 64, 16, 8, 4><<<grid, 256, 0, st>>>(X, ldx, W, ldw, bias, residual, ldr, Y, ldy, M, N, K, act, skip);
  } else if ((long long)cdiv(M, 64) * cdiv(N, 64) >= c->num_sms) {
    dim3 grid(cdiv(N, 64), cdiv(M, 64));
    sgemm_nt_kernel<64, 64, 16, 4, 4><<<grid, 256, 0, st>>>(X, ldx, W, ldw, bias, residual, ldr, Y, ldy, M, N, K, act, skip);
  } else {
    dim3 grid(cdiv(N, 16), cdiv(M, 32));
    sgemm_nt_kernel<32, 16, 16, 4, 4><<<grid, 32, 0, st>>>(X, ldx, W, ldw, bias, residual, ldr, Y, ldy, M, N, K, act, skip);
  }
  c->launches++;
  SAMPT_LAUNCH_CHECK();
  return 0;
}

}  // namespace sampt
