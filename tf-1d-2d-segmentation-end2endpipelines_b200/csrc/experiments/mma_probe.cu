// Probe: sustained tcgen05.mma throughput (bf16, cta_group::1) per instruction shape with operands resident in shared
// memory (no loads at all), one CTA per SM.  Answers: how much of the tensor peak do M=128 x N in {32,64,128,256}
// (and M=64) reach, K-major vs MN-major operands?  The wgrad / small-Cout conv tiling is chosen from these numbers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../ptx.cuh"

using namespace b2;

template <int M, int N, int AMN, int BMN>
__global__ void __launch_bounds__(128) mma_rate_kernel(int iters, long long* clk_out, int b_shift_rows, int b_sbo) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(M, N, AMN, BMN);
    // A tile: 128 rows x 64 k (16 KB) at 0 ; B tile: up to 256 rows x 64 k (32 KB) at 32 KB.  4 k-steps per "stage".
    const uint64_t adesc0 = AMN ? make_smem_desc(smem_u32(smem), 64 * 128, 1024) : make_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t bdesc0 = BMN ? make_smem_desc(smem_u32(smem) + 32768 + b_shift_rows * 128, 64 * 128, b_sbo) : make_smem_desc(smem_u32(smem) + 32768 + b_shift_rows * 128, 16, b_sbo);
    const uint32_t astep = AMN ? 128 : 2, bstep = BMN ? 128 : 2;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem + ((it & 1) ? (N <= 256 ? 256 : 0) : 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(d, adesc0 + k * astep, bdesc0 + k * bstep, idesc, 1u);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    clk_out[blockIdx.x] = t1 - t0;
  }
  __syncwarp();
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int M, int N, int AMN, int BMN>
static void run(const char* name, int n_sms, int b_shift_rows = 0, int b_sbo = 1024) {
  const int iters = 4096;
  long long* d_clk; cudaMalloc(&d_clk, n_sms * sizeof(long long));
  const int smem = 98 * 1024 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel<M, N, AMN, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  mma_rate_kernel<M, N, AMN, BMN><<<n_sms, 128, smem>>>(64, d_clk, b_shift_rows, b_sbo);   // warm-up
  cudaEventRecord(e0);
  mma_rate_kernel<M, N, AMN, BMN><<<n_sms, 128, smem>>>(iters, d_clk, b_shift_rows, b_sbo);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  if (err != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(err)); exit(1); }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(n_sms); cudaMemcpy(h.data(), d_clk, n_sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (auto v : h) avg += (double)v; avg /= n_sms;
  const double n_mma = 4.0 * iters;
  const double flops = 2.0 * M * N * 16 * n_mma * n_sms;
  printf("%-28s M=%3d N=%3d a_mn=%d b_mn=%d : %7.1f clk/MMA  %7.0f MAC/clk/SM  %8.1f TFLOP/s (all SMs, event time %.3f ms)\n", name, M, N, AMN, BMN,
         avg / n_mma, (double)M * N * 16 / (avg / n_mma), flops / (ms * 1e-3) / 1e12, ms);
  cudaFree(d_clk);
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int n = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, n);
  run<128, 256, 0, 0>("K-major/K-major", n);
  run<128, 128, 0, 0>("K-major/K-major", n);
  run<128, 64, 0, 0>("K-major/K-major", n);
  run<128, 32, 0, 0>("K-major/K-major", n);
  run<128, 16, 0, 0>("K-major/K-major", n);
  run<64, 256, 0, 0>("K-major/K-major", n);
  run<64, 128, 0, 0>("K-major/K-major", n);
  run<64, 64, 0, 0>("K-major/K-major", n);
  run<128, 256, 1, 1>("MN-major/MN-major (wgrad)", n);
  run<128, 128, 1, 1>("MN-major/MN-major (wgrad)", n);
  run<128, 64, 1, 1>("MN-major/MN-major (wgrad)", n);
  run<128, 32, 1, 1>("MN-major/MN-major (wgrad)", n);
  run<128, 256, 0, 1>("K-major/MN-major (dgrad)", n);
  run<128, 128, 0, 1>("K-major/MN-major (dgrad)", n);
  run<128, 64, 0, 1>("K-major/MN-major (dgrad)", n);
  // halo-style B operands: descriptor starts on an arbitrary 128-byte row, 8-row groups a window row apart
  run<128, 256, 1, 1>("MN/MN B shifted 3 rows, SBO 9 rows", n, 3, 9 * 128);
  run<128, 256, 1, 1>("MN/MN B shifted 1 row, SBO 10 rows", n, 1, 10 * 128);
  run<128, 128, 1, 1>("MN/MN B shifted 1 row, SBO 10 rows", n, 1, 10 * 128);
  run<128, 256, 0, 0>("K/K B shifted 3 rows, SBO 10 rows", n, 3, 10 * 128);
  run<128, 256, 1, 1>("MN/MN B shifted 4 rows, SBO 8 rows", n, 4, 8 * 128);
  return 0;
}
