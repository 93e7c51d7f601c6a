// Implicit-GEMM convolution for sm_100a, halo path: for dense k x k windows on large feature maps the A operand of all
// taps comes from ONE halo tile per 64-channel block — (bh+kh-1) x (bw+kw-1) pixels loaded by a single 4-D TMA box —
// and every tap is just a different start row of the shared-memory matrix descriptor.  This works because the
// tcgen05 128B swizzle is a function of the absolute shared-memory address (measured: csrc/experiments/swz_probe.cu),
// so a descriptor may start on any 128-byte row and use any 8-row-group pitch (SBO = halo row pitch).
// L2 -> smem traffic of the activations drops by ~k*k*128/((bh+kh-1)(bw+kw-1)) (6.4x for 3x3) versus one box per tap.
// Optionally the whole weight slice stays resident in shared memory (small layers), so the steady state streams
// activations only.  Tile = 16 rows x 8 columns of one image (2D) or 128 consecutive samples (1D).
//
// Same descriptor, scheduler and epilogue as conv_gemm.cu (include/b2seg.h: b2seg_conv).
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "conv_common.cuh"

namespace b2 {

constexpr int kHaloMaxAStages = 6;      // A (halo tile) pipeline depth is chosen per layer: (stages-1) x MMA time per stage must cover the TMA latency
constexpr int kHaloABytes = 24576;          // >= (bh+kh-1)*(bw+kw-1)*128
constexpr int kHaloMaxBStages = 16;
constexpr int kHaloMaxWins = 16;
constexpr int kHaloSmemBudget = 232448;     // 227 KiB

__device__ unsigned long long g_halo_trace[48 * 8];

struct HaloWin { int map, oh0, ow0, tap_begin, tap_end, pad0, pad1, pad2; };

struct alignas(64) HaloKParams {
  CUtensorMap amap[B2SEG_MAX_SRC];
  CUtensorMap bmap;
  HaloWin wins[kHaloMaxWins];
  int4 taps[B2SEG_MAX_TAPS];   // x = row of the tap inside the window, y = column, w = widx
  int wins_per_group, kc_blocks;
  int ww, a_bytes, sbo;        // halo width (pixels), bytes per A stage actually loaded, 8-row-group pitch in bytes
  int a_stages;                // halo-tile pipeline stages (2..kHaloMaxAStages)
  int b_stages;                // non-resident: number of B pipeline stages
  int b_region_bytes;
  int b_resident_tiles;        // resident: number of B tiles (taps * channel blocks)
  int k16_last;                // K = 16 steps that hold real channels in the last 64-channel block (1..4): the first layer
                               // (3 -> 8 input channels) issues one MMA per tap instead of four on TMA zero fill
  ConvEpiParams e;
};

template <int BLOCK_N, bool B_MN, bool B_RES>
__global__ void __launch_bounds__(kConvThreads, 1) conv_halo_kernel(const __grid_constant__ HaloKParams p) {
  constexpr int kBStageBytes = BLOCK_N * kBlockK * 2;
  constexpr int kTmemCols = 2 * BLOCK_N;
  pdl_launch_dependents();   // the next kernel of the stream may become resident as SMs drain
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  const int a_stages = p.a_stages;
  uint8_t* sB = smem + a_stages * kHaloABytes;
  uint8_t* staging = sB + p.b_region_bytes;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(staging + kStgBytes);
  uint64_t* emptyA = fullA + kHaloMaxAStages;
  uint64_t* fullB = emptyA + kHaloMaxAStages;
  uint64_t* emptyB = fullB + kHaloMaxBStages;
  uint64_t* tfull_bar = emptyB + kHaloMaxBStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* bres_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);
  uint32_t* s_tapoff = tmem_slot + 4;           // per tap: start-row offset inside the halo tile, in 16-byte units
  float* colpart = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(s_tapoff + B2SEG_MAX_TAPS) + 15) & ~uintptr_t(15));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler (uniform datapath)
  const int lane = threadIdx.x & 31;
  const ConvEpiParams& e = p.e;

  if (threadIdx.x == 0) {
    for (int i = 0; i < B2SEG_MAX_SRC; ++i) tma_prefetch_desc(&p.amap[i]);
    tma_prefetch_desc(&p.bmap);
    for (int s = 0; s < a_stages; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
    for (int s = 0; s < kHaloMaxBStages; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], kEpiWarps); }
    mbar_init(bres_bar, 1);
    for (int t = 0; t < B2SEG_MAX_TAPS; ++t) s_tapoff[t] = (uint32_t)(p.taps[t].x * p.ww + p.taps[t].y) * 8u;
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                // barriers / TMEM are set up; from here on global memory of earlier kernels is read
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      if (B_RES) {
        // the whole weight slice of this layer, once, in consumption order (window, channel block, tap)
        mbar_arrive_expect_tx(bres_bar, p.b_resident_tiles * kBStageBytes);
        int idx = 0;
        const int my_group = e.cta_groups > 1 ? (int)blockIdx.x % e.cta_groups : 0;
        for (int wi = 0; wi < p.wins_per_group; ++wi) {
          const HaloWin win = p.wins[my_group * p.wins_per_group + wi];
          for (int cb = 0; cb < p.kc_blocks; ++cb)
            for (int t = win.tap_begin; t < win.tap_end; ++t, ++idx) {
              uint8_t* sb = sB + idx * kBStageBytes;
              if (!B_MN) {
                tma_load_3d(&p.bmap, bres_bar, sb, cb * kBlockK, p.taps[t].w, 0);
              } else {
#pragma unroll
                for (int q = 0; q < BLOCK_N / 64; ++q) tma_load_3d(&p.bmap, bres_bar, sb + q * 8192, q * 64, p.taps[t].w, cb * kBlockK);
              }
            }
        }
      }
      uint32_t sa = 0, pa = 0, sb_i = 0, pb = 0;
      const TileRange tr = tile_range(e);
      for (int tile = tr.first; tile < tr.end; tile += tr.step) {
        const TileCoord tc = tile_coord(e, tile);
        const int n_tile = tc.n_tile, g = tc.g, w0 = tc.w0, h0 = tc.h0, n0 = tc.n0;
        {
          // pull the halo tiles this CTA needs two tiles from now into L2 (each activation byte is read from HBM once,
          // so without this every A stage pays full HBM latency with only two stages in flight)
          const int ptile = tile + 2 * tr.step;
          if (ptile < tr.end) {
            // (magic-number divisions: this single thread's per-tile latency is the floor of the whole kernel on the thin
            // high-resolution layers — with plain / and % it was ~3300 clocks per tile whatever K and the epilogue cost)
            const TileCoord pc = tile_coord(e, ptile);
            const int pg = pc.g, pw0 = pc.w0, ph0 = pc.h0, pn0 = pc.n0;
            if (e.n_tiles == 1 || pc.n_tile == 0)
            for (int wi = 0; wi < p.wins_per_group; ++wi) {
              const HaloWin win = p.wins[pg * p.wins_per_group + wi];
              for (int cb = 0; cb < p.kc_blocks; ++cb) tma_prefetch_4d(&p.amap[win.map], cb * kBlockK, pw0 + win.ow0, ph0 + win.oh0, pn0);
            }
          }
        }
        const int p_it = (tile - tr.first) / tr.step;
        for (int wi = 0; wi < p.wins_per_group; ++wi) {
          const HaloWin win = p.wins[g * p.wins_per_group + wi];
          for (int cb = 0; cb < p.kc_blocks; ++cb) {
            mbar_wait(&emptyA[sa], pa ^ 1);
            if (e.trace && blockIdx.x == 0 && p_it < 48 && wi == 0 && cb == 0) e.trace[p_it * 8 + 3] = (unsigned long long)clock64();
            mbar_arrive_expect_tx(&fullA[sa], p.a_bytes);
            tma_load_4d(&p.amap[win.map], &fullA[sa], sA + sa * kHaloABytes, cb * kBlockK, w0 + win.ow0, h0 + win.oh0, n0);
            if (++sa == (uint32_t)a_stages) { sa = 0; pa ^= 1; }
            if (!B_RES) {
              for (int t = win.tap_begin; t < win.tap_end; ++t) {
                mbar_wait(&emptyB[sb_i], pb ^ 1);
                uint8_t* sb = sB + sb_i * kBStageBytes;
                mbar_arrive_expect_tx(&fullB[sb_i], kBStageBytes);
                if (!B_MN) {
                  tma_load_3d(&p.bmap, &fullB[sb_i], sb, cb * kBlockK, p.taps[t].w, n_tile * BLOCK_N);
                } else {
#pragma unroll
                  for (int q = 0; q < BLOCK_N / 64; ++q)
                    tma_load_3d(&p.bmap, &fullB[sb_i], sb + q * 8192, n_tile * BLOCK_N + q * 64, p.taps[t].w, cb * kBlockK);
                }
                if (++sb_i == (uint32_t)p.b_stages) { sb_i = 0; pb ^= 1; }
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs this loop with warp-uniform control flow and one elected lane issues each tcgen05 operation,
    // so descriptors live in uniform registers (under `if (lane == 0)` every MMA paid ELECT + 8 R2UR, ~100 clocks:
    // issue-bound for N <= 128, profiles/r1_prof_conv2d_12_issue.txt).
    {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 0, B_MN ? 1 : 0);
      // descriptors are built once; per MMA only the 14-bit start-address field (16-byte units) changes
      const uint64_t adesc0 = make_smem_desc(0, 16, p.sbo);
      const uint64_t bdesc0 = B_MN ? make_smem_desc(0, 8192, 1024) : make_smem_desc(0, 16, 1024);
      const uint32_t sA16 = smem_u32(sA) >> 4, sB16 = smem_u32(sB) >> 4;
      const uint32_t ad_lo0 = (uint32_t)adesc0, ad_hi = (uint32_t)(adesc0 >> 32), bd_lo0 = (uint32_t)bdesc0, bd_hi = (uint32_t)(bdesc0 >> 32);
      const int ww_ = p.ww;
      const int wins_per_group = p.wins_per_group, kc_blocks = p.kc_blocks, b_stages = p.b_stages, k16_last = p.k16_last;
      uint32_t sa = 0, pa = 0, sb_i = 0, pb = 0, acc = 0, acc_phase = 0;
      if (B_RES) {
        mbar_wait(bres_bar, 0);
        tc_fence_after();
      }
      const TileRange tr = tile_range(e);
      for (int tile = tr.first; tile < tr.end; tile += tr.step) {
        const int g = (int)fast_div(fast_div((uint32_t)tile, e.fd_n_tiles), e.fd_m_tiles);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const int m_it = (tile - tr.first) / tr.step;
        const bool mtrace = e.trace != nullptr && blockIdx.x == 0 && lane == 0 && m_it < 48;
        if (mtrace) e.trace[m_it * 8 + 0] = (unsigned long long)clock64();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        uint32_t accumulate = 0;
        uint32_t res16 = sB16;
        for (int wi = 0; wi < wins_per_group; ++wi) {
          const int tap_begin = p.wins[g * wins_per_group + wi].tap_begin, tap_end = p.wins[g * wins_per_group + wi].tap_end;
          for (int cb = 0; cb < kc_blocks; ++cb) {
            mbar_wait(&fullA[sa], pa);
            tc_fence_after();
            if (mtrace && wi == 0 && cb == 0) e.trace[m_it * 8 + 1] = (unsigned long long)clock64();
            const uint32_t ad_stage = ad_lo0 + (sA16 + sa * (kHaloABytes >> 4));
            // the tap's start row inside the halo tile is fetched one tap AHEAD (an indexed constant load feeding a chain
            // of dependent uniform-datapath ops was ~265 clocks per tap: more than the 4 MMAs of an N <= 128 tap take)
            uint32_t off_next = (uint32_t)(p.taps[tap_begin].x * ww_ + p.taps[tap_begin].y) * 8u;
            for (int t = tap_begin; t < tap_end; ++t) {
              const uint32_t off = off_next;
              if (t + 1 < tap_end) off_next = (uint32_t)(p.taps[t + 1].x * ww_ + p.taps[t + 1].y) * 8u;
              uint32_t bd;
              if (B_RES) {
                bd = bd_lo0 + res16;
                res16 += kBStageBytes >> 4;
              } else {
                mbar_wait(&fullB[sb_i], pb);
                tc_fence_after();
                bd = bd_lo0 + (sB16 + sb_i * (kBStageBytes >> 4));
              }
              const uint32_t ad = ad_stage + off;
              if (cb + 1 < kc_blocks || k16_last == kBlockK / 16) {
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                  umma_bf16_elect_lh(d_tmem, ad + k * 2, ad_hi, bd + k * (B_MN ? 128 : 2), bd_hi, idesc, accumulate);
                  accumulate = 1;
                }
              } else {
                for (int k = 0; k < k16_last; ++k) {
                  umma_bf16_elect_lh(d_tmem, ad + k * 2, ad_hi, bd + k * (B_MN ? 128 : 2), bd_hi, idesc, accumulate);
                  accumulate = 1;
                }
              }
              if (!B_RES) {
                umma_commit_elect(&emptyB[sb_i]);
                if (++sb_i == (uint32_t)b_stages) { sb_i = 0; pb ^= 1; }
              }
            }
            umma_commit_elect(&emptyA[sa]);
            if (++sa == (uint32_t)a_stages) { sa = 0; pa ^= 1; }
          }
        }
        umma_commit_elect(&tfull_bar[acc]);
        if (mtrace) e.trace[m_it * 8 + 2] = (unsigned long long)clock64();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    conv_epilogue<BLOCK_N>(e, staging, colpart, tfull_bar, tempty_bar, tmem_base);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
struct WinBuild {
  int group, src, dh0, dh1, dw0, dw1;
  std::vector<int> taps;
};

static int build_windows(const b2seg_conv_desc* d, std::vector<WinBuild>* wins, int* KH, int* KW) {
  wins->clear();
  for (int g = 0; g < d->n_groups; ++g)
    for (int t = 0; t < d->taps_per_group; ++t) {
      const int ti = g * d->taps_per_group + t;
      const b2seg_tap& tp = d->taps[ti];
      WinBuild* w = nullptr;
      for (auto& c : *wins)
        if (c.group == g && c.src == tp.src) w = &c;
      if (!w) {
        wins->push_back(WinBuild{g, tp.src, tp.dh, tp.dh, tp.dw, tp.dw, {}});
        w = &wins->back();
      }
      w->dh0 = std::min(w->dh0, tp.dh); w->dh1 = std::max(w->dh1, tp.dh);
      w->dw0 = std::min(w->dw0, tp.dw); w->dw1 = std::max(w->dw1, tp.dw);
      w->taps.push_back(ti);
    }
  *KH = 0; *KW = 0;
  for (auto& c : *wins) {
    *KH = std::max(*KH, c.dh1 - c.dh0 + 1);
    *KW = std::max(*KW, c.dw1 - c.dw0 + 1);
  }
  return 0;
}

int halo_geometry(const b2seg_conv_desc* d, int* bw, int* bh, int* bn) {
  static const bool disabled = getenv("B2SEG_NO_HALO") != nullptr;
  if (disabled) return 1;
  std::vector<WinBuild> wins;
  int KH, KW;
  build_windows(d, &wins, &KH, &KW);
  if ((int)wins.size() > kHaloMaxWins || (int)wins.size() % d->n_groups != 0) return 1;
  if (KH * KW < 2) return 1;                       // single-tap windows: nothing to reuse
  for (int g = 0; g < d->n_groups; ++g) {
    int cnt = 0;
    for (auto& c : wins) cnt += (c.group == g);
    if (cnt != (int)wins.size() / d->n_groups) return 1;
  }
  const b2seg_view& o = d->out[0];
  if (o.H == 1) {
    if (KH != 1 || o.W % 128 != 0) return 1;
    *bw = 128; *bh = 1; *bn = 1;
  } else {
    if (o.W % 8 != 0 || o.H % 16 != 0) return 1;
    *bw = 8; *bh = 16; *bn = 1;
  }
  if ((*bh + KH - 1) * (*bw + KW - 1) * 128 > kHaloABytes) return 1;
  if (*bw + KW - 1 > 256 || *bh + KH - 1 > 256) return 1;
  for (int i = 0; i < d->n_src; ++i)
    if (d->src[i].N != o.N) return 1;
  return 0;
}

struct HaloLaunch : PreparedOp {
  HaloKParams kp;
  int block_n, grid, smem_bytes;
  bool b_mn, b_res;
  int launch(cudaStream_t s) override;
};

template <int BLOCK_N, bool B_MN, bool B_RES>
static int launch_halo_t(const HaloKParams& kp, int grid, int smem_bytes, cudaStream_t s) {
  static int attr_smem = 0;
  if (smem_bytes > attr_smem) {
    B2_CUDA_OK(cudaFuncSetAttribute(conv_halo_kernel<BLOCK_N, B_MN, B_RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmemBudget));
    attr_smem = kHaloSmemBudget;
  }
  B2_CUDA_OK(launch_k(conv_halo_kernel<BLOCK_N, B_MN, B_RES>, dim3(grid), dim3(kConvThreads), smem_bytes, s, kp));
  return 0;
}

template <int BLOCK_N>
static int launch_halo_n(const HaloLaunch& L, cudaStream_t s) {
  if (L.b_mn) return L.b_res ? launch_halo_t<BLOCK_N, true, true>(L.kp, L.grid, L.smem_bytes, s)
                             : launch_halo_t<BLOCK_N, true, false>(L.kp, L.grid, L.smem_bytes, s);
  return L.b_res ? launch_halo_t<BLOCK_N, false, true>(L.kp, L.grid, L.smem_bytes, s)
                 : launch_halo_t<BLOCK_N, false, false>(L.kp, L.grid, L.smem_bytes, s);
}

int HaloLaunch::launch(cudaStream_t s) {
  if (block_n == 64) return launch_halo_n<64>(*this, s);
  if (block_n == 128) return launch_halo_n<128>(*this, s);
  return launch_halo_n<256>(*this, s);
}

PreparedOp* prepare_conv_halo(const b2seg_conv_desc* d) {
  int bw, bh, bn;
  if (halo_geometry(d, &bw, &bh, &bn) != 0) { set_error("conv_halo: descriptor not eligible"); return nullptr; }
  std::vector<WinBuild> wins;
  int KH, KW;
  build_windows(d, &wins, &KH, &KW);
  HaloLaunch* L = new HaloLaunch();
  HaloKParams& kp = L->kp;
  memset(&kp, 0, sizeof(kp));
  L->b_mn = d->b_mn_major != 0;
  L->block_n = select_block_n(d);
  if (fill_epi_params(d, L->block_n, bw, bh, bn, &kp.e) != 0) { delete L; return nullptr; }
  const int hh = bh + KH - 1, ww = bw + KW - 1;
  kp.ww = ww;
  kp.a_bytes = hh * ww * 128;
  kp.sbo = (bh == 1) ? 1024 : ww * 128;
  const int k_ch = L->b_mn ? d->w_cout : d->w_cin;
  kp.kc_blocks = (k_ch + kBlockK - 1) / kBlockK;
  kp.k16_last = (k_ch - (kp.kc_blocks - 1) * kBlockK + 15) / 16;
  kp.wins_per_group = (int)wins.size() / d->n_groups;
  // windows are created group by group, so wins[] is already ordered by group
  int tcount = 0;
  for (size_t wi = 0; wi < wins.size(); ++wi) {
    const WinBuild& c = wins[wi];
    HaloWin& hw = kp.wins[wi];
    hw.map = c.src; hw.oh0 = c.dh0; hw.ow0 = c.dw0;
    hw.tap_begin = tcount;
    for (int ti : c.taps) {
      const b2seg_tap& tp = d->taps[ti];
      kp.taps[tcount++] = make_int4(tp.dh - c.dh0, tp.dw - c.dw0, 0, tp.widx);
    }
    hw.tap_end = tcount;
  }
  for (int i = 0; i < d->n_src; ++i)
    if (encode_act_map(&kp.amap[i], d->src[i], kBlockK, ww, hh, 1) != 0) { delete L; return nullptr; }
  for (int i = d->n_src; i < B2SEG_MAX_SRC; ++i) kp.amap[i] = kp.amap[0];
  if (!L->b_mn) {
    if (encode_weight_map(&kp.bmap, d->weights, d->w_cout, d->w_taps, d->w_cin, kBlockK, L->block_n) != 0) { delete L; return nullptr; }
  } else {
    if (encode_weight_map(&kp.bmap, d->weights, d->w_cout, d->w_taps, d->w_cin, 64, kBlockK) != 0) { delete L; return nullptr; }
  }
  const int b_stage = L->block_n * kBlockK * 2;
  const int fixed = 1024 + kStgBytes + (2 * kHaloMaxAStages + 2 * kHaloMaxBStages + 5) * 8 + 16 + B2SEG_MAX_TAPS * 4 + kColPartBytes + 64;
  const int budget = kHaloSmemBudget - fixed;   // A stages + B region
  // A pipeline depth.  Measured (profiles/r1_convbench_astages.txt): deeper halo pipelines do not help any cfg2 layer —
  // narrow layers are bound by re-streaming the weights through L2 -> shared memory (64 B/clk/SM needed, ~42 available),
  // not by halo latency — and giving up resident weights for more stages costs 1.7x on the Cout = 64 layers.
  static const int env_a = getenv("B2SEG_HALO_ASTAGES") ? atoi(getenv("B2SEG_HALO_ASTAGES")) : 0;
  int want_a = 2;
  if (env_a >= 2 && env_a <= kHaloMaxAStages) want_a = env_a;
  // Several groups (transposed conv: one per output parity): give every CTA ONE group, so that group's weight slice can stay
  // resident, and let the G CTAs of a quad walk the same m tiles in step — the input tile they share is then read from HBM
  // once instead of once per group (ncu on conv2d_transpose_4 before: 535 MB read for a 134 MB input, L2 hit rate 48 %).
  static const bool no_ctag = getenv("B2SEG_NO_CTA_GROUPS") != nullptr;
  const int sms = num_sms();
  const bool cta_groups = !no_ctag && d->n_groups > 1 && sms % d->n_groups == 0 && kp.e.n_tiles == 1 &&
                          kp.e.m_tiles >= 2 * (sms / d->n_groups);
  const int res_tiles = d->taps_per_group * kp.kc_blocks;   // per group
  static const bool no_res = getenv("B2SEG_NO_BRES") != nullptr;
  const bool res_ok = !no_res && (d->n_groups == 1 || cta_groups) && kp.e.n_tiles == 1 && kp.e.total_tiles >= 2 * sms;
  // resident weights leave room for how many A stages?
  const int a_with_res = res_ok ? std::min(want_a, (budget - res_tiles * b_stage) / kHaloABytes) : 0;
  L->b_res = res_ok && a_with_res >= std::min(want_a, 3);
  if (L->b_res) {
    kp.a_stages = a_with_res;
    kp.b_resident_tiles = res_tiles;
    kp.b_region_bytes = res_tiles * b_stage;
    kp.b_stages = 1;
  } else {
    int st = L->block_n == 256 ? 4 : (L->block_n == 128 ? 6 : 8);
    const int min_st = L->block_n == 256 ? 3 : 4;
    kp.a_stages = want_a;
    while (st > min_st && kp.a_stages * kHaloABytes + st * b_stage > budget) --st;
    while (kp.a_stages > 2 && kp.a_stages * kHaloABytes + st * b_stage > budget) --kp.a_stages;
    while (st > 1 && kp.a_stages * kHaloABytes + st * b_stage > budget) --st;
    kp.b_stages = st;
    kp.b_region_bytes = st * b_stage;
  }
  L->smem_bytes = fixed + kp.a_stages * kHaloABytes + kp.b_region_bytes;
  L->grid = kp.e.total_tiles < sms ? kp.e.total_tiles : sms;
  kp.e.cta_groups = (L->b_res && cta_groups) ? d->n_groups : 0;   // only together with resident weights
  static const bool trace_on = getenv("B2SEG_TRACE") != nullptr;
  if (trace_on) {
    void* sym = nullptr;
    if (cudaGetSymbolAddress(&sym, g_halo_trace) == cudaSuccess) kp.e.trace = reinterpret_cast<unsigned long long*>(sym);
  }
  return L;
}

int read_halo_trace(unsigned long long* out, int n) {
  if (n > 48 * 8) n = 48 * 8;
  return cudaMemcpyFromSymbol(out, g_halo_trace, (size_t)n * 8) == cudaSuccess ? n : -1;
}

}  // namespace b2
