// Micro-benchmark behind profiles/r02_roofline_summary.md §1: how many SM cycles does one tcgen05.mma take when the same thread
// issues a long run of them, (a) all accumulating into ONE TMEM tile (what a GEMM K loop does), (b) alternating between TWO TMEM
// tiles, so that consecutive instructions are independent?  Shapes M128 x N{256,128} x K16 (kind::f16) and K32 (kind::f8f6f4,
// e4m3), operands resident in shared memory (128B-swizzled K-major tiles; contents irrelevant), one CTA per SM on every SM so that the
// chip runs at its loaded clock.  No TMA, no epilogue: only the tensor pipe and its operand fetch.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sam-pt_b200/csrc -o /tmp/mma_probe tools/mma_cadence_probe.cu && /tmp/mma_probe
#include <algorithm>
#include <cstdio>
#include <vector>

#include "tc_common.cuh"

using namespace sampt::tc;

__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mode 0: every instruction accumulates into tile 0;  1: instruction i goes to tile (i & 1);  2: k-block (4 instructions) b goes to tile (b & 1)
// distinct_smem: 0 = every k-block reads the same shared-memory stage, 1 = walks NSTAGE different stages (as a pipelined GEMM does)
__global__ void __launch_bounds__(128, 1)
probe_kernel(int n_kblocks, int mode, int N, int f8, int distinct_smem, long long* cycles_out) {
  constexpr int NSTAGE = 4, A_BYTES = 128 * 128, B_BYTES = 256 * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * (A_BYTES + B_BYTES));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NSTAGE * (A_BYTES + B_BYTES) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (warp == 0) {
    if (lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_f16(128, N, 0);
    // warm-up run (first-touch effects), then the timed run
    for (int rep = 0; rep < 2; ++rep) {
      const long long t0 = clock64();
      uint32_t n_issued = 0;
      for (int kb = 0; kb < n_kblocks; ++kb) {
        const int st = distinct_smem ? (kb % NSTAGE) : 0;
        const uint32_t sa = smem_u32(smem + st * (A_BYTES + B_BYTES));
        const uint64_t adesc = make_smem_desc_sw128(sa), bdesc = make_smem_desc_sw128(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k, ++n_issued) {
          const int tile = mode == 0 ? 0 : (mode == 1 ? (int)(n_issued & 1) : (kb & 1));
          const uint32_t d = tmem_base + (uint32_t)(tile * 256);
          if (f8) umma_f8(d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
          else umma_f16(d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
        }
      }
      umma_commit(bar);
      mbar_wait(bar, (uint32_t)rep & 1);
      tc_fence_after();
      const long long t1 = clock64();
      if (rep == 1) cycles_out[blockIdx.x] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int main() {
  int dev = 0, sms = 0;
  cudaSetDevice(dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = 4 * (128 * 128 + 256 * 128) + 1024 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long* d_cyc;
  cudaMalloc(&d_cyc, sizeof(long long) * sms);
  std::vector<long long> h(sms);
  const int n_kblocks = 2048;   // 8192 instructions per CTA and run
  printf("SMs %d, %d k-blocks x 4 instructions per CTA; cycles per instruction (median over CTAs; min .. max)\n", sms, n_kblocks);
  printf("%-10s %-5s %-28s %-14s %s\n", "kind", "N", "accumulators", "operand stages", "cycles / instruction");
  for (int f8 = 0; f8 < 2; ++f8)
    for (int N : {256, 128})
      for (int distinct = 0; distinct < 2; ++distinct)
        for (int mode = 0; mode < 3; ++mode) {
          probe_kernel<<<sms, 128, smem>>>(n_kblocks, mode, N, f8, distinct, d_cyc);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h.data(), d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
          std::sort(h.begin(), h.end());
          const double per = 1.0 / (4.0 * n_kblocks);
          printf("%-10s %-5d %-28s %-14s %.1f  (%.1f .. %.1f)\n", f8 ? "e4m3 K32" : "f16 K16", N,
                 mode == 0 ? "one (dependent)" : (mode == 1 ? "two, alternating per instr" : "two, alternating per k-block"),
                 distinct ? "4 rotating" : "1 fixed", h[sms / 2] * per, h[0] * per, h[sms - 1] * per);
        }
  cudaFree(d_cyc);
  return 0;
}
