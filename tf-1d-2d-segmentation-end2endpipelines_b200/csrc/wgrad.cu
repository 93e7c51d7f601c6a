// Weight-gradient (TF Conv2DBackpropFilter) as a tcgen05 GEMM with both operands MN-major:
//   dW[co][tap][ci] = sum_pixels dY[pixel][co] * X[pixel + tap][ci]
// GEMM-M = co (128 per CTA), GEMM-N = ci (BLOCK_N), GEMM-K = pixels, streamed 64 at a time by 4-D TMA boxes
// (zero OOB fill = SAME padding).  One (m-tile, n-tile, tap, k-split) work item per CTA; split-K partial
// sums are combined with fp32 reductions into the pre-zeroed gradient buffer.
#include "common.h"
#include "ptx.cuh"

namespace b2 {

constexpr int kWgPix = 64;                  // pixels (GEMM-K) per stage
constexpr int kWgABytes = 2 * kWgPix * 128; // two 64-channel boxes of dY
constexpr int kWgThreads = 128;

struct alignas(64) WgradKParams {
  CUtensorMap ymap[B2SEG_MAX_SRC];
  CUtensorMap xmap[B2SEG_MAX_SRC];
  int taps[B2SEG_MAX_TAPS][6];  // pair, dyh, dyw, dh, dw, widx
  int n_taps;
  int bw, bh, bn, tiles_w, tiles_h, tiles_n, k_chunks;
  int ksplit, chunks_per_split;
  int m_tiles, n_tiles;
  float* dw;
  int w_cout, w_taps, w_cin;
  int atomic;
};

template <int BLOCK_N>
struct WgCfg {
  static constexpr int kBBytes = (BLOCK_N / 64) * kWgPix * 128;
  static constexpr int kStageBytes = kWgABytes + kBBytes;
  static constexpr int kStages = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 3 : 4);
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 256;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kWgThreads) wgrad_kernel(const __grid_constant__ WgradKParams p) {
  using Cfg = WgCfg<BLOCK_N>;
  pdl_launch_dependents();   // the next kernel of the stream may become resident as SMs drain
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler

  // work item
  int wi = blockIdx.x;
  const int m_tile = wi % p.m_tiles; wi /= p.m_tiles;
  const int n_tile = wi % p.n_tiles; wi /= p.n_tiles;
  const int tap_i = wi % p.n_taps; wi /= p.n_taps;
  const int split = wi;
  const int c_begin = split * p.chunks_per_split;
  const int c_end = min(c_begin + p.chunks_per_split, p.k_chunks);
  const int n_iter = c_end - c_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, BLOCK_N); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                // barriers / TMEM are set up; from here on global memory of earlier kernels is read
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const int pair = p.taps[tap_i][0], dyh = p.taps[tap_i][1], dyw = p.taps[tap_i][2];
  const int dh = p.taps[tap_i][3], dw = p.taps[tap_i][4], widx = p.taps[tap_i][5];

  if (warp == 0) {
    if (lane == 0 && n_iter > 0) {
      tma_prefetch_desc(&p.ymap[pair]);
      tma_prefetch_desc(&p.xmap[pair]);
      uint32_t stage = 0, phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        const int w0 = (c % p.tiles_w) * p.bw;
        const int h0 = ((c / p.tiles_w) % p.tiles_h) * p.bh;
        const int n0 = (c / (p.tiles_w * p.tiles_h)) * p.bn;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + kWgABytes;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
#pragma unroll
        for (int q = 0; q < 2; ++q)
          tma_load_4d(&p.ymap[pair], &full_bar[stage], sa + q * (kWgPix * 128), m_tile * 128 + q * 64, w0 + dyw, h0 + dyh, n0);
#pragma unroll
        for (int q = 0; q < BLOCK_N / 64; ++q)
          tma_load_4d(&p.xmap[pair], &full_bar[stage], sb + q * (kWgPix * 128), n_tile * BLOCK_N + q * 64, w0 + dw, h0 + dh, n0);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // whole warp, one elected lane issues (uniform-register descriptors, see conv_halo.cu)
    if (n_iter > 0) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 1, 1);
      const uint64_t desc0 = make_smem_desc(0, kWgPix * 128, 1024);  // both operands MN-major: LBO = one 64-channel box
      const uint32_t smem16 = smem_u32(smem) >> 4;
      uint32_t stage = 0, phase = 0;
      for (int it = 0; it < n_iter; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a16 = smem16 + stage * (Cfg::kStageBytes >> 4);
        const uint64_t ad = desc0 + a16, bd = desc0 + (a16 + (kWgABytes >> 4));
#pragma unroll
        for (int k = 0; k < kWgPix / 16; ++k) umma_bf16_elect(tmem_base, ad + k * 128, bd + k * 128, idesc, (it | k) != 0 ? 1u : 0u);
        umma_commit_elect(&empty_bar[stage]);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit_elect(tfull_bar);
    }
    __syncwarp();
  }

  if (n_iter > 0) {
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const int co = m_tile * 128 + warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_base + lane_base + c * 32, v);
      tmem_ld_wait();
      const int ci0 = n_tile * BLOCK_N + c * 32;
      if (co < p.w_cout && ci0 < p.w_cin) {
        float* dst = p.dw + ((size_t)co * p.w_taps + widx) * p.w_cin + ci0;
        if (p.atomic) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ci0 + j < p.w_cin) atomicAdd(dst + j, __uint_as_float(v[j]));
        } else if (ci0 + 32 <= p.w_cin && (p.w_cin & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            reinterpret_cast<float4*>(dst)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                            __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ci0 + j < p.w_cin) dst[j] = __uint_as_float(v[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, BLOCK_N); }
}

struct WgradLaunch : PreparedOp {
  WgradKParams kp;
  int block_n, grid;
  int launch(cudaStream_t s) override;
};

template <int BLOCK_N>
static int launch_wgrad_t(const WgradKParams& kp, int grid, cudaStream_t s) {
  using Cfg = WgCfg<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    B2_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  B2_CUDA_OK(launch_k(wgrad_kernel<BLOCK_N>, dim3(grid), dim3(kWgThreads), Cfg::kSmemBytes, s, kp));
  return 0;
}

int WgradLaunch::launch(cudaStream_t s) {
  if (block_n == 64) return launch_wgrad_t<64>(kp, grid, s);
  if (block_n == 128) return launch_wgrad_t<128>(kp, grid, s);
  return launch_wgrad_t<256>(kp, grid, s);
}

PreparedOp* prepare_wgrad(const b2seg_wgrad_desc* d) {
  if (d->n_pair < 1 || d->n_pair > B2SEG_MAX_SRC || d->n_taps < 1 || d->n_taps > B2SEG_MAX_TAPS) {
    set_error("wgrad: bad pair/tap counts");
    return nullptr;
  }
  if (d->w_cin % 8 != 0 || d->w_cout % 8 != 0) {
    set_error("wgrad: channel extents must be multiples of 8");
    return nullptr;
  }
  {
    bool hard_error = false;
    PreparedOp* fused = prepare_wgrad_halo(d, &hard_error);  // fused-tap halo kernel (wgrad_halo.cu) when eligible
    if (fused || hard_error) return fused;
  }
  WgradLaunch* L = new WgradLaunch();
  WgradKParams& kp = L->kp;
  memset(&kp, 0, sizeof(kp));
  pick_box(d->gN, d->gH, d->gW, kWgPix, &kp.bw, &kp.bh, &kp.bn);
  kp.tiles_w = (d->gW + kp.bw - 1) / kp.bw;
  kp.tiles_h = (d->gH + kp.bh - 1) / kp.bh;
  kp.tiles_n = (d->gN + kp.bn - 1) / kp.bn;
  kp.k_chunks = kp.tiles_w * kp.tiles_h * kp.tiles_n;
  L->block_n = d->w_cin <= 64 ? 64 : (d->w_cin <= 128 ? 128 : 256);
  kp.m_tiles = (d->w_cout + 127) / 128;
  kp.n_tiles = (d->w_cin + L->block_n - 1) / L->block_n;
  kp.n_taps = d->n_taps;
  const int base_items = kp.m_tiles * kp.n_tiles * kp.n_taps;
  int ksplit = d->ksplit;
  if (ksplit <= 0) {
    // aim for ~3 waves of CTAs, but keep at least 8 K-chunks (512 pixels) per CTA
    const int target = 3 * num_sms();
    ksplit = (target + base_items - 1) / base_items;
    const int max_split = kp.k_chunks / 8 > 0 ? kp.k_chunks / 8 : 1;
    if (ksplit > max_split) ksplit = max_split;
    if (ksplit < 1) ksplit = 1;
  }
  kp.chunks_per_split = (kp.k_chunks + ksplit - 1) / ksplit;
  ksplit = (kp.k_chunks + kp.chunks_per_split - 1) / kp.chunks_per_split;
  kp.ksplit = ksplit;
  kp.atomic = (ksplit > 1 || d->accumulate) ? 1 : 0;
  for (int i = 0; i < d->n_pair; ++i) {
    if (encode_act_map(&kp.ymap[i], d->dy[i], 64, kp.bw, kp.bh, kp.bn) != 0 ||
        encode_act_map(&kp.xmap[i], d->x[i], 64, kp.bw, kp.bh, kp.bn) != 0) {
      delete L;
      return nullptr;
    }
  }
  for (int i = d->n_pair; i < B2SEG_MAX_SRC; ++i) { kp.ymap[i] = kp.ymap[0]; kp.xmap[i] = kp.xmap[0]; }
  for (int t = 0; t < d->n_taps; ++t) {
    const b2seg_wgrad_tap& tp = d->taps[t];
    if (tp.pair < 0 || tp.pair >= d->n_pair || tp.widx < 0 || tp.widx >= d->w_taps) {
      set_error("wgrad: tap %d out of range", t);
      delete L;
      return nullptr;
    }
    kp.taps[t][0] = tp.pair; kp.taps[t][1] = tp.dyh; kp.taps[t][2] = tp.dyw;
    kp.taps[t][3] = tp.dh; kp.taps[t][4] = tp.dw; kp.taps[t][5] = tp.widx;
  }
  kp.dw = reinterpret_cast<float*>(d->dw);
  kp.w_cout = d->w_cout; kp.w_taps = d->w_taps; kp.w_cin = d->w_cin;
  L->grid = base_items * ksplit;
  return L;
}

}  // namespace b2
