// Patch-similarity filtering of tracked points: reference sam_pt/modeling/sam_pt.py:597-682.
//   rgbs_lab = skimage.color.rgb2lab(rgbs[:, [2, 1, 0]])            (:645  -- note the channel swap: B,G,R is fed as R,G,B)
//   patches  = grid_sample(rgbs_lab, (xy + offsets + 0.5) / (w, h) * 2 - 1, bilinear, align_corners=False)   (:604-622)
//   sim      = exp(-||patch_traj - patch_query||_2 / (2 * patch_size^2))                                      (:628-638)
//   vis[(vis == 1) & ~(sim > thr)] = PATCH_NON_SIMILAR ; everything after (before) the first non-similar frame in the forward
//   (backward) direction = REJECTED_AFTER_PATCH_WAS_NON_SIMILAR                                               (:655-682)
// The reference converts the WHOLE clip to Lab on the host (float64, 3 x 8 B per pixel); here only the <= 4 * ps^2 pixels a
// patch touches are converted, on the fly, in fp64 (pow / cbrt) and rounded to fp32 exactly where the reference rounds
// (`torch.as_tensor(rgbs_lab, dtype=float32)`), then sampled in fp32 like grid_sample.  One thread per (frame, point).
#include "common.cuh"
#include "../../include/sampt_b200.h"

namespace sampt {

// skimage.color.rgb2lab (illuminant D65, observer 2) of one uint8 pixel given as (r, g, b) = the values skimage sees
__device__ __forceinline__ void rgb2lab_f64(uint8_t r8, uint8_t g8, uint8_t b8, float* lab) {
  double c[3] = {r8 / 255.0, g8 / 255.0, b8 / 255.0};
#pragma unroll
  for (int i = 0; i < 3; ++i) c[i] = c[i] > 0.04045 ? pow((c[i] + 0.055) / 1.055, 2.4) : c[i] / 12.92;
  double x = 0.412453 * c[0] + 0.357580 * c[1] + 0.180423 * c[2];
  double y = 0.212671 * c[0] + 0.715160 * c[1] + 0.072169 * c[2];
  double z = 0.019334 * c[0] + 0.119193 * c[1] + 0.950227 * c[2];
  x /= 0.95047; z /= 1.08883;
  double f[3] = {x, y, z};
#pragma unroll
  for (int i = 0; i < 3; ++i) f[i] = f[i] > 0.008856 ? cbrt(f[i]) : 7.787 * f[i] + 16.0 / 116.0;
  lab[0] = (float)(116.0 * f[1] - 16.0);
  lab[1] = (float)(500.0 * (f[0] - f[1]));
  lab[2] = (float)(200.0 * (f[1] - f[2]));
}

// bilinear sample (align_corners=False with the +0.5 shift of the reference == pixel coordinates, zero padding) of the Lab image
__device__ __forceinline__ void sample_lab(const uint8_t* frame /*[3,H,W] planar RGB*/, int H, int W, float x, float y, float* out) {
  const float x0f = floorf(x), y0f = floorf(y);
  const int x0 = (int)x0f, y0 = (int)y0f;
  const float wx1 = x - x0f, wy1 = y - y0f, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
  out[0] = out[1] = out[2] = 0.f;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const int xx = x0 + dx, yy = y0 + dy;
      if (xx < 0 || yy < 0 || xx >= W || yy >= H) continue;
      const size_t o = (size_t)yy * W + xx, P = (size_t)H * W;
      float lab[3];
      // the reference feeds channels [2, 1, 0] to rgb2lab: skimage's "r" is our B plane
      rgb2lab_f64(frame[2 * P + o], frame[P + o], frame[o], lab);
      const float w = (dy ? wy1 : wy0) * (dx ? wx1 : wx0);
      out[0] += lab[0] * w; out[1] += lab[1] * w; out[2] += lab[2] * w;
    }
}

__global__ void __launch_bounds__(128)
patch_similarity_kernel(const uint8_t* __restrict__ frames, int T, int H, int W, const float* __restrict__ query /*[N,3] t,x,y*/,
                        const float* __restrict__ traj /*[T,N,2]*/, int N, int ps, float* __restrict__ sim /*[T,N]*/) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T * N) return;
  const int t = idx / N, n = idx % N;
  const int tq = (int)query[3 * n];
  const float qx = query[3 * n + 1], qy = query[3 * n + 2];
  const float px = traj[((size_t)t * N + n) * 2], py = traj[((size_t)t * N + n) * 2 + 1];
  const uint8_t* fq = frames + (size_t)tq * 3 * H * W;
  const uint8_t* ft = frames + (size_t)t * 3 * H * W;
  float ss = 0.f;
  const int h = ps / 2;
  for (int a = -h; a <= h; ++a)        // meshgrid(ij): first component added to x, second to y (sam_pt.py:607-610)
    for (int b = -h; b <= h; ++b) {
      float l1[3], l2[3];
      sample_lab(fq, H, W, qx + (float)a, qy + (float)b, l1);
      sample_lab(ft, H, W, px + (float)a, py + (float)b, l2);
#pragma unroll
      for (int c = 0; c < 3; ++c) { const float d = l2[c] - l1[c]; ss += d * d; }
    }
  sim[idx] = expf(-sqrtf(ss) / (2.f * (float)(ps * ps)));
}

// vis codes: 1 visible, -3 PATCH_NON_SIMILAR, -4 REJECTED_AFTER_PATCH_WAS_NON_SIMILAR (sam_pt/utils/util.py:267-282)
__global__ void patch_reject_kernel(float* __restrict__ vis /*[T,N]*/, const float* __restrict__ sim, const float* __restrict__ query,
                                    int T, int N, float thr) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  for (int t = 0; t < T; ++t) {
    float& v = vis[(size_t)t * N + n];
    if (v == 1.f && !(sim[(size_t)t * N + n] > thr)) v = -3.f;
  }
  const int tq = (int)query[3 * n];
  for (int t = tq + 1; t < T; ++t)
    if (vis[(size_t)t * N + n] == -3.f) {
      for (int u = t + 1; u < T; ++u) vis[(size_t)u * N + n] = -4.f;
      break;
    }
  for (int t = tq - 1; t >= 0; --t)
    if (vis[(size_t)t * N + n] == -3.f) {
      for (int u = 0; u < t; ++u) vis[(size_t)u * N + n] = -4.f;
      break;
    }
}

}  // namespace sampt

using namespace sampt;

// frames [T,3,H,W] u8; query [N,3] = (t,x,y); traj [T,N,2]; vis [T,N] float codes (in/out); sim [T,N] scratch/out
extern "C" int sampt_patch_filter(sampt_ctx* ctx, const uint8_t* frames, int T, int H, int W, const float* query, const float* traj,
                                  int N, int patch_size, float threshold, float* vis, float* sim, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SAMPT_CHECK(patch_size >= 1 && patch_size <= 15, "sampt_patch_filter: patch_size %d outside [1, 15]", patch_size);
  patch_similarity_kernel<<<cdiv((long long)T * N, 128), 128, 0, st>>>(frames, T, H, W, query, traj, N, patch_size, sim);
  c->launches++;
  patch_reject_kernel<<<cdiv(N, 128), 128, 0, st>>>(vis, sim, query, T, N, threshold);
  c->launches++;
  SAMPT_LAUNCH_CHECK();
  return 0;
}
