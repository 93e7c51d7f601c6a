// Probe: how does tcgen05.mma (SWIZZLE_128B, K-major) treat an A descriptor whose start address is a 128-byte row
// that is NOT 1024-byte aligned (a shifted window into a larger TMA-written tile)?
// Loads a [256 rows][64 bf16] tile by TMA, then for each (shift, base_offset variant, sbo) computes
// D[128 x 64] = A_window * I (B = identity 64x64) and reports which variant reproduces rows [shift .. shift+128).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o swz_probe swz_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../ptx.cuh"

using namespace b2;

struct Params {
  CUtensorMap amap;   // 2D: (64 ch, 512 rows), box (64, 256)
  CUtensorMap bmap;   // 2D: (64, 64) identity
  int shift_rows;     // window start row
  int base_off_mode;  // 0: base_offset = 0 ; 1: base_offset = (start >> 7) & 7
  int sbo_bytes;      // stride between 8-row groups
  int group_pitch_rows;  // for the reference: rows between consecutive 8-row groups
  float* out;         // [128][64]
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                 // 512 rows * 128 B = 64 KB
  uint8_t* sb = smem + 512 * 128;     // 64 rows * 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 64 * 128);
  uint64_t* dbar = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(dbar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(dbar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 512 * 128 + 64 * 128);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(sa)), "l"(reinterpret_cast<uint64_t>(&p.amap)), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(sa + 256 * 128)), "l"(reinterpret_cast<uint64_t>(&p.amap)), "r"(smem_u32(bar)), "r"(0), "r"(256) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(sb)), "l"(reinterpret_cast<uint64_t>(&p.bmap)), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    for (int k = 0; k < 4; ++k) {
      const uint32_t a_addr = smem_u32(sa) + p.shift_rows * 128 + k * 32;
      uint64_t adesc = make_smem_desc(a_addr, 16, p.sbo_bytes);
      if (p.base_off_mode == 1) adesc |= (uint64_t)((a_addr >> 7) & 7) << 49;
      const uint64_t bdesc = make_smem_desc(smem_u32(sb) + k * 32, 16, 1024);
      umma_bf16(tmem, adesc, bdesc, idesc, k != 0);
    }
    umma_commit(dbar);
  }
  __syncwarp();
  mbar_wait(dbar, 0);
  tc_fence_after();
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const int ROWS = 512;
  std::vector<__nv_bfloat16> ha(ROWS * 64), hb(64 * 64);
  for (int r = 0; r < ROWS; ++r) for (int c = 0; c < 64; ++c) ha[r * 64 + c] = __float2bfloat16((float)(r * 64 + c) / 16.0f - 1000.f);  // exact in bf16? use small ints instead
  for (int r = 0; r < ROWS; ++r) for (int c = 0; c < 64; ++c) ha[r * 64 + c] = __float2bfloat16((float)((r * 7 + c * 3) % 251));
  for (int r = 0; r < 64; ++r) for (int c = 0; c < 64; ++c) hb[r * 64 + c] = __float2bfloat16(r == c ? 1.f : 0.f);
  __nv_bfloat16 *da, *db; float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  Params p;
  {
    cuuint64_t dims[2] = {64, (cuuint64_t)ROWS}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, 256}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&p.amap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode a failed %d\n", r); return 1; }
    cuuint64_t dimb[2] = {64, 64}; cuuint32_t boxb[2] = {64, 64};
    r = enc(&p.bmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, dimb, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode b failed %d\n", r); return 1; }
  }
  p.out = dout;
  const int smem_bytes = 1024 + 512 * 128 + 64 * 128 + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  std::vector<float> ho(128 * 64);
  // group pitch (rows between consecutive 8-row groups): 8 = dense, 10 / 16 / 18 = halo-style pitches
  const int pitches[] = {8, 10, 16, 18};
  for (int pi = 0; pi < 4; ++pi) {
    for (int shift = 0; shift <= 19; ++shift) {
      if (shift > 11 && shift != 17 && shift != 19) continue;
      for (int mode = 0; mode < 2; ++mode) {
        p.shift_rows = shift; p.base_off_mode = mode; p.group_pitch_rows = pitches[pi]; p.sbo_bytes = pitches[pi] * 128;
        cudaMemset(dout, 0, 128 * 64 * 4);
        probe_kernel<<<1, 128, smem_bytes>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("pitch %d shift %d mode %d: CUDA error %s\n", pitches[pi], shift, mode, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m) {
          const int src_row = shift + (m / 8) * pitches[pi] + (m % 8);
          for (int c = 0; c < 64; ++c) {
            const float want = src_row < ROWS ? __bfloat162float(ha[src_row * 64 + c]) : 0.f;
            if (ho[m * 64 + c] != want) ++bad;
          }
        }
        printf("pitch_rows %2d shift %2d base_offset_mode %d : %s (%d mismatches)\n", pitches[pi], shift, mode, bad == 0 ? "OK" : "WRONG", bad);
      }
    }
  }
  return 0;
}
