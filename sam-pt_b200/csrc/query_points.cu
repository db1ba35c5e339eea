// Query-point extraction from masks: k-medoids on <= 2048 mask pixels, on the GPU.
// Reference: sam_pt/utils/query_points.py:64-104 -> sklearn_extra.cluster.KMedoids(n_clusters=N).fit(pixels).cluster_centers_
// (scikit-learn-extra, un-vendored; metric "euclidean", method "alternate", init "heuristic", max_iter 300).
//
// The package works on a float32 distance matrix with numpy reductions; medoid selection is an argmin over float32 sums, so
// near-ties are decided by the SUMMATION ORDER.  To pick the same medoids the kernels below reproduce numpy's float32
// `add.reduce` order exactly (pairwise summation: 8 interleaved accumulators on blocks <= 128, recursive halving above,
// numpy/core/src/umath/loops_utils.h.src) and numpy's first-occurrence argmin.  The one order that numpy leaves unspecified
// -- `np.argpartition` of the row sums for the "heuristic" initialisation -- stays on the host: phase 1 returns the n row
// sums, the Python side calls numpy itself on them and hands the k initial medoids to phase 2.
//
//   phase 1  kmed_dist_kernel     D[i][j] = sqrtf(max(float(|x_i|^2 + |x_j|^2 - 2 x_i.x_j  in fp64), 0)), D[i][i] = 0
//            kmed_rowsum_kernel   rowsum[i] = pairwise_f32(D[i][:])            (D is symmetric: column reads are coalesced)
//   phase 2  kmed_iterate_kernel  ONE CTA, loops to convergence on the device (no host round trips):
//              labels = argmin_k D[medoid_k][i]; member lists (ascending index, one warp per cluster, ballot compaction);
//              cost[i] = pairwise_f32(D[members(label_i)][i]); per cluster: first-min argmin, strict `<` against the
//              current medoid's cost; stop when no medoid moved.
#include "common.cuh"
#include "../../include/sampt_b200.h"

namespace sampt {

constexpr int KMED_MAX_N = 2048;
constexpr int KMED_MAX_K = 256;

__global__ void __launch_bounds__(256)
kmed_dist_kernel(const float* __restrict__ pts, int n, float* __restrict__ D) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= n) return;
  const double yi = pts[2 * i], xi = pts[2 * i + 1], yj = pts[2 * j], xj = pts[2 * j + 1];
  const double d2 = (yi * yi + xi * xi) + (yj * yj + xj * xj) - 2.0 * (yi * yj + xi * xj);
  float f = (float)d2;
  f = fmaxf(f, 0.f);
  if (i == j) f = 0.f;
  D[(size_t)i * n + j] = sqrtf(f);
}

// numpy's float32 pairwise sum over v(lo .. lo+m-1); `at(j)` returns element j.  The recursion depth is a template parameter
// (m <= 128 * 2^DEPTH): compile-time bounded, so the compiler sizes the stack itself (true device recursion overflowed the
// default 1 KB stack for n = 1800).
template <typename F>
__device__ __forceinline__ float np_pairwise_leaf(F at, int lo, int m) {
  if (m < 8) {
    float r = 0.f;
    for (int i = 0; i < m; ++i) r = __fadd_rn(r, at(lo + i));
    return r;
  }
  float r[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) r[q] = at(lo + q);
  int i = 8;
  for (; i < m - (m % 8); i += 8) {
#pragma unroll
    for (int q = 0; q < 8; ++q) r[q] = __fadd_rn(r[q], at(lo + i + q));
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < m; ++i) res = __fadd_rn(res, at(lo + i));
  return res;
}
template <int DEPTH, typename F>
__device__ float np_pairwise_sum_d(F at, int lo, int m) {
  if constexpr (DEPTH == 0) {
    return np_pairwise_leaf(at, lo, m);
  } else {
    if (m <= 128) return np_pairwise_leaf(at, lo, m);
    int m2 = m / 2;
    m2 -= m2 % 8;
    const float a = np_pairwise_sum_d<DEPTH - 1>(at, lo, m2);
    const float b = np_pairwise_sum_d<DEPTH - 1>(at, lo + m2, m - m2);
    return __fadd_rn(a, b);
  }
}
template <typename F>
__device__ __forceinline__ float np_pairwise_sum(F at, int lo, int m) { return np_pairwise_sum_d<5>(at, lo, m); }   // m <= 4096

__global__ void __launch_bounds__(128)
kmed_rowsum_kernel(const float* __restrict__ D, int n, float* __restrict__ rowsum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  auto at = [&](int j) { return D[(size_t)j * n + i]; };  // == D[i][j] (symmetric), coalesced across the warp
  rowsum[i] = np_pairwise_sum(at, 0, n);
}

// scratch (global, ints): labels[n] | members[n] | offsets[k+1] ; costs[n] floats
__global__ void __launch_bounds__(1024, 1)
kmed_iterate_kernel(const float* __restrict__ D, int n, int k, int max_iter, int* __restrict__ medoids, int* __restrict__ labels,
                    int* __restrict__ members, float* __restrict__ costs, int* __restrict__ n_iter_out) {
  __shared__ int s_med[KMED_MAX_K];
  __shared__ int s_cnt[KMED_MAX_K];
  __shared__ int s_off[KMED_MAX_K + 1];
  __shared__ int s_changed;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int c = tid; c < k; c += blockDim.x) s_med[c] = medoids[c];
  __syncthreads();
  int it = 0;
  for (; it < max_iter; ++it) {
    // labels = np.argmin(D[medoid_idxs, :], axis=0): first minimum over clusters in order
    for (int i = tid; i < n; i += blockDim.x) {
      float best = D[(size_t)s_med[0] * n + i];
      int bc = 0;
      for (int c = 1; c < k; ++c) {
        const float d = D[(size_t)s_med[c] * n + i];
        if (d < best) { best = d; bc = c; }
      }
      labels[i] = bc;
    }
    if (tid == 0) s_changed = 0;
    __syncthreads();
    // member counts, then offsets, then ascending member lists (np.where(labels == c)[0])
    for (int c = warp; c < k; c += nwarps) {
      int cnt = 0;
      for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const unsigned m = __ballot_sync(0xffffffffu, i < n && labels[i] == c);
        cnt += __popc(m);
      }
      if (lane == 0) s_cnt[c] = cnt;
    }
    __syncthreads();
    if (tid == 0) {
      int o = 0;
      for (int c = 0; c < k; ++c) { s_off[c] = o; o += s_cnt[c]; }
      s_off[k] = o;
    }
    __syncthreads();
    for (int c = warp; c < k; c += nwarps) {
      int pos = s_off[c];
      for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const bool in = i < n && labels[i] == c;
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (in) members[pos + __popc(m & ((1u << lane) - 1))] = i;
        pos += __popc(m);
      }
    }
    __syncthreads();
    // costs[i] = np.sum(D[members][:, members], axis=1)[position of i]
    for (int i = tid; i < n; i += blockDim.x) {
      const int c = labels[i];
      const int* mem = members + s_off[c];
      auto at = [&](int j) { return D[(size_t)mem[j] * n + i]; };
      costs[i] = np_pairwise_sum(at, 0, s_cnt[c]);
    }
    __syncthreads();
    // per cluster: first-min argmin over members; move if strictly cheaper than the current medoid
    for (int c = warp; c < k; c += nwarps) {
      const int cnt = s_cnt[c];
      if (cnt == 0) continue;                       // empty cluster: medoid unchanged (the package only warns)
      const int* mem = members + s_off[c];
      float best = INFINITY;
      int bpos = 0x7fffffff;
      for (int p = lane; p < cnt; p += 32) {
        const float v = costs[mem[p]];
        if (v < best) { best = v; bpos = p; }        // lanes see ascending positions: first occurrence per lane
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
        if (ob < best || (ob == best && op < bpos)) { best = ob; bpos = op; }
      }
      if (lane == 0) {
        const int cur = s_med[c];
        // np.argmax(members == medoid): position of the current medoid in its cluster, 0 if it is not a member
        float cur_cost = costs[labels[cur] == c ? cur : mem[0]];
        if (best < cur_cost) { s_med[c] = mem[bpos]; s_changed = 1; }
      }
    }
    __syncthreads();
    const int changed = s_changed;
    __syncthreads();
    if (!changed) break;
  }
  for (int c = tid; c < k; c += blockDim.x) medoids[c] = s_med[c];
  if (tid == 0) *n_iter_out = (it < max_iter) ? it + 1 : max_iter;
}

}  // namespace sampt

using namespace sampt;

// phase 1: pts [n,2] float32 (y,x) -> D [n,n] float32, rowsum [n] float32
extern "C" int sampt_kmedoids_distances(sampt_ctx* ctx, const float* pts, int n, float* D, float* rowsum, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SAMPT_CHECK(n >= 1 && n <= KMED_MAX_N, "sampt_kmedoids_distances: n = %d outside [1, %d]", n, KMED_MAX_N);
  kmed_dist_kernel<<<dim3(cdiv(n, 256), n), 256, 0, st>>>(pts, n, D);
  c->launches++;
  kmed_rowsum_kernel<<<cdiv(n, 128), 128, 0, st>>>(D, n, rowsum);
  c->launches++;
  SAMPT_LAUNCH_CHECK();
  return 0;
}

// phase 2: medoids [k] int32 (in: initial medoids, out: converged), scratch_i [2n] int32, scratch_f [n] float32, n_iter [1] int32
extern "C" int sampt_kmedoids_iterate(sampt_ctx* ctx, const float* D, int n, int k, int max_iter, int* medoids, int* scratch_i,
                                      float* scratch_f, int* n_iter, void* stream) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SAMPT_CHECK(n >= 1 && n <= KMED_MAX_N, "sampt_kmedoids_iterate: n = %d outside [1, %d]", n, KMED_MAX_N);
  SAMPT_CHECK(k >= 1 && k <= KMED_MAX_K && k <= n, "sampt_kmedoids_iterate: k = %d outside [1, min(%d, n)]", k, KMED_MAX_K);
  kmed_iterate_kernel<<<1, 1024, 0, st>>>(D, n, k, max_iter, medoids, scratch_i, scratch_i + n, scratch_f, n_iter);
  c->launches++;
  SAMPT_LAUNCH_CHECK();
  return 0;
}
