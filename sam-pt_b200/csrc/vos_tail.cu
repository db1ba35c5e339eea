// VOS harness tail fused into ONE kernel: reference sam_pt/vos_eval/eval.py:304-345
//   logits = stack([zeros] + per-object logits, dim=1)              background channel of zero logits        (:304)
//   logits[:gt_ti, i+1] = -1e8                                      nothing before an object's query frame   (:321-322)
//   logits[gt_ti, i+1]  = where(nearest(gt_mask) , 1e8, -1e8)       ground truth overwrites the query frame  (:324-326)
//   probs = softmax(logits, dim=1)                                                                           (:327)
//   per frame: bilinear(probs -> original shape, align_corners=False) if need_resize ; flip ; argmax -> uint8 (:343-355)
// The reference materialises (T, 1+M, H, W) fp32 logits AND probabilities (2 x 82 MB x (1+M) at 480x854) and walks them frame by
// frame on the host; here every output pixel reads its <= 4 source pixels of the M object logit maps once and writes one byte.
// HBM-bound: algorithmic bytes = M*T*H*W*4 read + T*Ho*Wo written.
#include "common.cuh"
#include "../../include/sampt_b200.h"

namespace sampt {

struct VosTailArgs {
  const float* logits;        // [M, T, H, W]
  const float* gt;            // [M, Hg, Wg] {0,1}
  const int* gt_ti;           // [M]
  int M, T, H, W, Hg, Wg;
  int Ho, Wo, need_resize, flip;
  uint8_t* out;               // [T, Ho, Wo]
};

// value of channel c (0 = background) at source pixel (y, x) of frame t, after the -1e8 / ground-truth overwrites
__device__ __forceinline__ float vos_logit(const VosTailArgs& a, int c, int t, int y, int x) {
  if (c == 0) return 0.f;
  const int i = c - 1;
  const int ti = a.gt_ti[i];
  if (t < ti) return -1e8f;
  if (t == ti) {
    // F.interpolate(gt[None, None], target_hw, mode="nearest"): src = min(floor(dst * (in / out)), in - 1), float32 scale
    const float sy = (float)a.Hg / (float)a.H, sx = (float)a.Wg / (float)a.W;
    const int gy = min((int)floorf((float)y * sy), a.Hg - 1), gx = min((int)floorf((float)x * sx), a.Wg - 1);
    return a.gt[((size_t)i * a.Hg + gy) * a.Wg + gx] != 0.f ? 1e8f : -1e8f;
  }
  return a.logits[(((size_t)i * a.T + t) * a.H + y) * a.W + x];
}

__global__ void __launch_bounds__(256)
vos_tail_kernel(VosTailArgs a) {
  const int t = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.Ho * a.Wo) return;
  const int oy = p / a.Wo, ox_out = p % a.Wo;
  const int ox = a.flip ? (a.Wo - 1 - ox_out) : ox_out;   // torch.flip(prob, dims=[-1]) after the resize
  const int C = a.M + 1;
  int ys[2], xs[2];
  float wy[2], wx[2];
  int ny = 1, nx = 1;
  if (a.need_resize) {
    // F.interpolate(bilinear, align_corners=False): src = max((dst + 0.5) * scale - 0.5, 0), scale = in / out (float32)
    const float sh = (float)a.H / (float)a.Ho, sw = (float)a.W / (float)a.Wo;
    float fy = fmaxf(((float)oy + 0.5f) * sh - 0.5f, 0.f), fx = fmaxf(((float)ox + 0.5f) * sw - 0.5f, 0.f);
    ys[0] = (int)fy; xs[0] = (int)fx;
    ys[1] = ys[0] + (ys[0] < a.H - 1 ? 1 : 0); xs[1] = xs[0] + (xs[0] < a.W - 1 ? 1 : 0);
    wy[1] = fy - (float)ys[0]; wy[0] = 1.f - wy[1];
    wx[1] = fx - (float)xs[0]; wx[0] = 1.f - wx[1];
    ny = nx = 2;
  } else {
    ys[0] = oy; xs[0] = ox; wy[0] = wx[0] = 1.f;
  }
  // pass 1: softmax statistics of each source pixel (max, sum of exp)
  float mx[2][2], den[2][2];
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      float m = 0.f;  // channel 0
      for (int c = 1; c < C; ++c) m = fmaxf(m, vos_logit(a, c, t, ys[j], xs[i]));
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += expf(vos_logit(a, c, t, ys[j], xs[i]) - m);
      mx[j][i] = m; den[j][i] = s;
    }
  // pass 2: blended probability per channel, first-maximum argmax (torch.argmax)
  float best = -1.f;
  int arg = 0;
  for (int c = 0; c < C; ++c) {
    float pr;
    if (a.need_resize) {
      float q[2][2];
      for (int j = 0; j < 2; ++j)
        for (int i = 0; i < 2; ++i) q[j][i] = expf(vos_logit(a, c, t, ys[j], xs[i]) - mx[j][i]) / den[j][i];
      pr = wy[0] * (wx[0] * q[0][0] + wx[1] * q[0][1]) + wy[1] * (wx[0] * q[1][0] + wx[1] * q[1][1]);
    } else {
      pr = expf(vos_logit(a, c, t, ys[0], xs[0]) - mx[0][0]) / den[0][0];
    }
    if (pr > best) { best = pr; arg = c; }
  }
  a.out[((size_t)t * a.Ho + oy) * a.Wo + ox_out] = (uint8_t)arg;
}

}  // namespace sampt

using namespace sampt;

extern "C" int sampt_vos_index_masks(sampt_ctx* ctx, const float* logits, int M, int T, int H, int W, const float* gt_masks, int Hg,
                                     int Wg, const int* gt_ti, int Ho, int Wo, int need_resize, int flip, uint8_t* out, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  SAMPT_CHECK(M >= 1 && M <= 254, "sampt_vos_index_masks: %d objects (uint8 index masks hold at most 254 + background)", M);
  SAMPT_CHECK(need_resize || (Ho == H && Wo == W), "sampt_vos_index_masks: output %dx%d != %dx%d without need_resize", Ho, Wo, H, W);
  VosTailArgs a{logits, gt_masks, gt_ti, M, T, H, W, Hg, Wg, Ho, Wo, need_resize, flip, out};
  vos_tail_kernel<<<dim3(cdiv((long long)Ho * Wo, 256), T), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  c->launches++;
  SAMPT_LAUNCH_CHECK();
  return 0;
}
