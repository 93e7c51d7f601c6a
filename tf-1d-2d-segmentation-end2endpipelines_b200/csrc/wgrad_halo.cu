// Weight gradient with fused taps: the halo-tile, persistent form of wgrad.cu.
//
//   dW[co][tap][ci] = sum_pixels dY[pixel][co] * X[pixel + tap][ci]
//
// wgrad.cu gives every tap its own CTA, so both operand tiles are fetched once per tap (768 B of L2 -> shared-memory
// traffic per pixel per 128x256 accumulator).  Here one CTA owns a GROUP of taps that read the same dY tile: per pixel
// tile it loads dY once and ONE halo window of X that covers all the group's shifts, then issues one MMA chain per tap
// into separate TMEM accumulators.  A tap's B operand is just a descriptor whose start address is the window row of its
// shift (tcgen05's 128-byte swizzle is a function of the absolute shared-memory address, experiments/swz_probe.cu;
// shifted / pitched descriptors run at full MMA rate, experiments/mma_probe.cu), with the stride between 8-pixel groups
// = one window row.  For Cin <= 64 one MMA covers several taps: consecutive 64-column chunks of the MN-major B operand
// are taken LBO = one pixel (128 B) apart, i.e. the same window shifted by one tap.
//
// GEMM-M = co (128 per CTA), GEMM-N = ci chunks x taps (<= 512 TMEM columns), GEMM-K = pixels (tiles of 64 or 128).
// Scheduling is persistent: the host cuts the (item = m-tile x ci-tile x tap-group, pixel tile) space into one list of
// segments per CTA with equal MMA time (no wave quantisation); a segment that does not cover its item's whole pixel
// range is combined with fp32 red.global into the pre-zeroed gradient.
//
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (TMEM -> global), so the producer keeps
// filling the pipeline for the next segment while the accumulators of the previous one drain.
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace b2 {

constexpr int kWhMaxGroups = 16;
constexpr int kWhMaxSlots = 8;
constexpr int kWhMaxXMaps = 8;
constexpr int kWhMaxStages = 6;
constexpr int kWhThreads = 192;
constexpr int kWhTailBytes = 1024;                                // barriers + the MMA warp's slot table
constexpr int kWhSmemBudget = 227 * 1024 - 1024 - kWhTailBytes;   // stages only (1 KB alignment slack)

struct WhSlot {
  uint32_t blo;    // low descriptor word without the stage address: LBO field | window row of the first tap (16-B units)
  uint32_t bhi;    // high descriptor word (SBO, version, swizzle)
  uint32_t idesc;
  uint32_t col;    // first TMEM column
};

struct WhGroup {
  // producer
  int pair, dyh, dyw;   // dY tile: source pair and coordinate offset
  int hmin, wmin;       // X window origin relative to the pixel tile
  int xmap;             // tensor map index
  int xchunk_bytes;     // 1024-aligned size of one 64-channel window in shared memory
  int xbox_bytes;       // bytes one TMA box delivers
  // MMA issuer (copied to shared memory per segment)
  int n_slots, pad0, pad1, pad2;
  uint32_t krow8[8];    // window row of pixel 16k of the tile, in 16-B units (tap shift excluded)
  WhSlot slot[kWhMaxSlots];
  // epilogue
  int cols, pad3, pad4, pad5;
  int chunk_widx[8];    // per 64 TMEM columns: weight tap index
  int chunk_ci[8];      //                      64-channel chunk inside the ci block
};
constexpr int kWhMmaPartOffset = 8 * 4;                                  // byte offset of n_slots
constexpr int kWhMmaPartBytes = 4 * 4 + 8 * 4 + kWhMaxSlots * 16;        // n_slots .. slot[]

struct WhSeg { int m_tile, n_tile, group, c_begin, c_end, atomic; };

struct alignas(64) WhParams {
  CUtensorMap ymap[B2SEG_MAX_SRC];
  CUtensorMap xmap[kWhMaxXMaps];
  WhGroup groups[kWhMaxGroups];   // in the parameter (constant) bank: the MMA warp reads it through the uniform datapath
  const WhSeg* segs;
  const int* cta_seg;      // [n_ctas + 1] first segment of every CTA
  int bw, bh, bn, tiles_w, tiles_h;
  int a_bytes;             // dY tile: 2 x tile_px x 128
  int stage_bytes, n_stages, xchunks, block_ci;
  float* dw;
  int w_cout, w_taps, w_cin;
};

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint64_t mk64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
__host__ __device__ constexpr uint32_t smem_desc_hi(uint32_t sbo_bytes) {   // matches make_smem_desc (ptx.cuh)
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
}

template <int KSTEPS>
__global__ void __launch_bounds__(kWhThreads) wgrad_halo_kernel(const __grid_constant__ WhParams p) {
  pdl_launch_dependents();   // the next kernel of the stream may become resident as SMs drain
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tail = smem + p.n_stages * p.stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kWhMaxStages;
  uint64_t* tfull_bar = empty_bar + kWhMaxStages;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  const int seg_begin = p.cta_seg[blockIdx.x], seg_end = p.cta_seg[blockIdx.x + 1];

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.n_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 4);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                // barriers / TMEM are set up; from here on global memory of earlier kernels is read
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // uniform register for the MMA operands

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      const int half_a = p.a_bytes >> 1;
      for (int si = seg_begin; si < seg_end; ++si) {
        const WhSeg sg = p.segs[si];
        const WhGroup* G = &p.groups[sg.group];
        const CUtensorMap* ym = &p.ymap[G->pair];
        const CUtensorMap* xm = &p.xmap[G->xmap];
        const int dyh = G->dyh, dyw = G->dyw, hmin = G->hmin, wmin = G->wmin, xchunk = G->xchunk_bytes;
        const uint32_t tx = p.a_bytes + p.xchunks * G->xbox_bytes;
        const int co0 = sg.m_tile * 128, ci0 = sg.n_tile * p.block_ci;
        for (int c = sg.c_begin; c < sg.c_end; ++c) {
          const int w0 = (c % p.tiles_w) * p.bw;
          const int h0 = ((c / p.tiles_w) % p.tiles_h) * p.bh;
          const int n0 = (c / (p.tiles_w * p.tiles_h)) * p.bn;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * p.stage_bytes;
          uint8_t* sb = sa + p.a_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          tma_load_4d(ym, &full_bar[stage], sa, co0, w0 + dyw, h0 + dyh, n0);
          tma_load_4d(ym, &full_bar[stage], sa + half_a, co0 + 64, w0 + dyw, h0 + dyh, n0);
          for (int q = 0; q < p.xchunks; ++q)
            tma_load_4d(xm, &full_bar[stage], sb + q * xchunk, ci0 + q * 64, w0 + wmin, h0 + hmin, n0);
          if (++stage == (uint32_t)p.n_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    // dY tile: dense MN-major, 8-pixel groups 1024 B apart, the second 64-channel chunk half a tile further
    const uint32_t a_lo0 = (uint32_t)((p.a_bytes >> 1) >> 4) << 16;
    constexpr uint32_t a_hi = smem_desc_hi(1024);
    const uint32_t smem16 = smem_u32(smem) >> 4;
    const uint32_t stage16 = p.stage_bytes >> 4, ab16 = p.a_bytes >> 4;
    uint32_t stage = 0, phase = 0;
    for (int si = seg_begin; si < seg_end; ++si) {
      const int group = __shfl_sync(0xffffffffu, p.segs[si].group, 0);
      const int n_iter = __shfl_sync(0xffffffffu, p.segs[si].c_end - p.segs[si].c_begin, 0);
      const WhGroup& G = p.groups[group];
      if (si > seg_begin) { mbar_wait(tempty_bar, (uint32_t)(si - seg_begin - 1) & 1); tc_fence_after(); }
      const int n_slots = G.n_slots;
      for (int it = 0; it < n_iter; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t a16 = smem16 + stage * stage16;
        const uint32_t b16 = a16 + ab16;
        const uint32_t a_lo = a_lo0 + a16;
        for (int s = 0; s < n_slots; ++s) {
          const uint32_t blo = G.slot[s].blo + b16;
          const uint32_t bhi = G.slot[s].bhi, idesc = G.slot[s].idesc;
          const uint32_t d_tmem = tmem_base + G.slot[s].col;
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k)
            umma_bf16_elect(d_tmem, mk64(a_lo + k * 128, a_hi), mk64(blo + G.krow8[k], bhi), idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit_elect(&empty_bar[stage]);
        if (++stage == (uint32_t)p.n_stages) { stage = 0; phase ^= 1; }
      }
      umma_commit_elect(tfull_bar);
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> dW
    const int q = warp & 3;   // TMEM lane quarter this warp may read
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    for (int si = seg_begin; si < seg_end; ++si) {
      const WhSeg sg = p.segs[si];
      const WhGroup* G = &p.groups[sg.group];
      const int n32 = G->cols >> 5;
      const int co = sg.m_tile * 128 + q * 32 + lane;
      mbar_wait(tfull_bar, (uint32_t)(si - seg_begin) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < n32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_base + c * 32, v);
        tmem_ld_wait();
        const int ci0 = sg.n_tile * p.block_ci + G->chunk_ci[c >> 1] * 64 + (c & 1) * 32;
        if (co < p.w_cout && ci0 < p.w_cin) {
          float* dst = p.dw + ((size_t)co * p.w_taps + G->chunk_widx[c >> 1]) * p.w_cin + ci0;
          if (ci0 + 32 <= p.w_cin) {
            if (sg.atomic) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                red_add_v4(dst + 4 * j, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                           __uint_as_float(v[4 * j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                reinterpret_cast<float4*>(dst)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (ci0 + j < p.w_cin) {
                if (sg.atomic) atomicAdd(dst + j, __uint_as_float(v[j]));
                else dst[j] = __uint_as_float(v[j]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

struct WgradHaloLaunch : PreparedOp {
  WhParams kp;
  void* d_tables = nullptr;
  int grid = 0, smem_bytes = 0, ksteps = 8;
  ~WgradHaloLaunch() override { if (d_tables) cudaFree(d_tables); }
  int launch(cudaStream_t s) override {
    static int attr_bytes[2] = {0, 0};
    const int which = ksteps == 8 ? 1 : 0;
    if (smem_bytes > attr_bytes[which]) {
      if (which) B2_CUDA_OK(cudaFuncSetAttribute(wgrad_halo_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      else B2_CUDA_OK(cudaFuncSetAttribute(wgrad_halo_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      attr_bytes[which] = smem_bytes;
    }
    if (which) B2_CUDA_OK(launch_k(wgrad_halo_kernel<8>, dim3(grid), dim3(kWhThreads), smem_bytes, s, kp));
    else B2_CUDA_OK(launch_k(wgrad_halo_kernel<4>, dim3(grid), dim3(kWhThreads), smem_bytes, s, kp));
    return 0;
  }
};

static int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

namespace {
struct Tap { int pair, dyh, dyw, dh, dw, widx; };
struct Slot { std::vector<Tap> taps; };   // >1 tap only when one MMA spans several taps (Cin <= 64)
struct GroupPlan { std::vector<Slot> slots; };
struct Item { int m, n, g; double cost; };  // cost = MMA clocks per pixel tile
}  // namespace

static double mma_clocks(int n) { return n <= 64 ? 54.5 : n * 0.5; }   // experiments/mma_probe.cu, M = 128

// Returns nullptr WITHOUT setting an error when the descriptor is simply not eligible (caller falls back to wgrad.cu).
PreparedOp* prepare_wgrad_halo(const b2seg_wgrad_desc* d, bool* hard_error) {
  *hard_error = false;
  static const bool disabled = getenv("B2SEG_NO_WGRAD_HALO") != nullptr;
  if (disabled) return nullptr;
  static const int env_block_ci = getenv("B2SEG_WG_BLOCK_CI") ? atoi(getenv("B2SEG_WG_BLOCK_CI")) : 0;
  static const int env_tile_px = getenv("B2SEG_WG_TILE_PX") ? atoi(getenv("B2SEG_WG_TILE_PX")) : 0;
  static const int env_ctas = getenv("B2SEG_WG_CTAS") ? atoi(getenv("B2SEG_WG_CTAS")) : 0;
  if (d->n_taps < 2) return nullptr;  // 1x1 convolutions: nothing to fuse

  // ci block: 128 channels x 3-4 fused taps measured best for 3x3 / 1xk kernels (1250-1350 TFLOP/s on the cfg2 layers
  // vs 890-950 for 256 channels x tap pairs, profiles/r1_wgrad_sweep.txt); the 2x2-window parity groups of a
  // transposed convolution (4 taps per dY tile) prefer 256 channels x tap pairs
  int taps_per_y = 0;
  for (int a = 0; a < d->n_taps; ++a) {
    int same = 0;
    for (int b = 0; b < d->n_taps; ++b)
      same += d->taps[a].pair == d->taps[b].pair && d->taps[a].dyh == d->taps[b].dyh && d->taps[a].dyw == d->taps[b].dyw;
    taps_per_y = std::max(taps_per_y, same);
  }
  int xchunks = std::min(taps_per_y <= 4 ? 4 : 2, (d->w_cin + 63) / 64);
  if (env_block_ci == 64 || env_block_ci == 128 || env_block_ci == 192 || env_block_ci == 256) xchunks = std::min(xchunks, env_block_ci / 64);
  const int block_ci = xchunks * 64;

  // ---- slots: MMA chains.  Taps sorted so that taps sharing a dY tile are adjacent, row-major in (dh, dw).
  std::vector<Tap> taps;
  for (int t = 0; t < d->n_taps; ++t) {
    const b2seg_wgrad_tap& tp = d->taps[t];
    if (tp.pair < 0 || tp.pair >= d->n_pair || tp.widx < 0 || tp.widx >= d->w_taps) return nullptr;  // wgrad.cu reports it
    taps.push_back({tp.pair, tp.dyh, tp.dyw, tp.dh, tp.dw, tp.widx});
  }
  std::sort(taps.begin(), taps.end(), [](const Tap& a, const Tap& b) {
    if (a.pair != b.pair) return a.pair < b.pair;
    if (a.dyh != b.dyh) return a.dyh < b.dyh;
    if (a.dyw != b.dyw) return a.dyw < b.dyw;
    if (a.dh != b.dh) return a.dh < b.dh;
    return a.dw < b.dw;
  });
  auto same_y = [](const Tap& a, const Tap& b) { return a.pair == b.pair && a.dyh == b.dyh && a.dyw == b.dyw; };
  std::vector<std::vector<Slot>> ygroups;  // slots per dY tile
  for (size_t i = 0; i < taps.size(); ++i) {
    if (i == 0 || !same_y(taps[i - 1], taps[i])) ygroups.emplace_back();
    std::vector<Slot>& sl = ygroups.back();
    bool extend = false;
    if (xchunks == 1 && !sl.empty()) {
      const Tap& last = sl.back().taps.back();
      extend = last.dh == taps[i].dh && last.dw + 1 == taps[i].dw && sl.back().taps.size() < 4;
    }
    if (extend) sl.back().taps.push_back(taps[i]);
    else { sl.emplace_back(); sl.back().taps.push_back(taps[i]); }
  }

  // ---- CTA groups: split each dY tile's slots evenly into groups of <= 512 TMEM columns
  std::vector<GroupPlan> plans;
  for (auto& sl : ygroups) {
    size_t i = 0;
    while (i < sl.size()) {
      int cols_left = 0;
      for (size_t j = i; j < sl.size(); ++j) cols_left += xchunks == 1 ? 64 * (int)sl[j].taps.size() : block_ci;
      const int groups_left = (cols_left + 511) / 512;
      const int target = (cols_left + groups_left - 1) / groups_left;
      GroupPlan g;
      int cols = 0;
      while (i < sl.size() && (int)g.slots.size() < kWhMaxSlots) {
        const int n = xchunks == 1 ? 64 * (int)sl[i].taps.size() : block_ci;
        if (cols + n > 512 || (cols >= target && cols > 0)) break;
        g.slots.push_back(sl[i]);
        cols += n;
        ++i;
      }
      plans.push_back(g);
    }
  }
  if ((int)plans.size() > kWhMaxGroups) return nullptr;

  // ---- pixel tile and pipeline depth
  int tile_px = (env_tile_px == 64 || env_tile_px == 128) ? env_tile_px : 128;
  WhParams kp;
  std::vector<WhGroup> groups;
  std::vector<double> group_cost;
  struct XKey { int pair, ww, wh; };
  std::vector<XKey> xkeys;
  int ksteps = 8;
  for (int attempt = 0; attempt < 2; ++attempt) {
    memset(&kp, 0, sizeof(kp));
    groups.clear();
    group_cost.clear();
    xkeys.clear();
    if (d->gH > 1) {
      kp.bw = 8;
      kp.bh = std::min(tile_px / 8, pow2ceil(d->gH));
    } else {
      kp.bw = std::min(tile_px, pow2ceil(std::max(d->gW, 8)));
      kp.bh = 1;
    }
    kp.bn = tile_px / (kp.bw * kp.bh);
    if (kp.bn > 256) return nullptr;
    ksteps = tile_px / 16;
    kp.a_bytes = 2 * tile_px * 128;
    int max_stage = 0;
    bool ok = true;
    for (const GroupPlan& gp : plans) {
      WhGroup g;
      memset(&g, 0, sizeof(g));
      const Tap& t0 = gp.slots[0].taps[0];
      g.pair = t0.pair; g.dyh = t0.dyh; g.dyw = t0.dyw;
      int hmin = 1 << 30, hmax = -(1 << 30), wmin = 1 << 30, wmax = -(1 << 30);
      for (const Slot& s : gp.slots)
        for (const Tap& t : s.taps) {
          hmin = std::min(hmin, t.dh); hmax = std::max(hmax, t.dh);
          wmin = std::min(wmin, t.dw); wmax = std::max(wmax, t.dw);
        }
      g.hmin = hmin; g.wmin = wmin;
      const int ww = kp.bw + (wmax - wmin), wh = kp.bh + (hmax - hmin);
      if (ww > 256 || wh > 256) { ok = false; break; }
      g.xbox_bytes = ww * wh * kp.bn * 128;
      g.xchunk_bytes = (g.xbox_bytes + 1023) / 1024 * 1024;
      auto row = [&](int pix) {  // window row of tile pixel `pix` (w fastest, then h, then n)
        const int w = pix % kp.bw, h = (pix / kp.bw) % kp.bh, n = pix / (kp.bw * kp.bh);
        return (n * wh + h) * ww + w;
      };
      const int sbo_rows = row(8) - row(0);
      for (int k = 0; k < ksteps; ++k) {
        g.krow8[k] = (uint32_t)row(16 * k) * 8u;
        if (row(16 * k + 8) - row(16 * k) != sbo_rows || row(16 * k + 7) - row(16 * k) != 7) ok = false;
      }
      const int sbo_bytes = sbo_rows * 128;
      if (!ok || sbo_bytes <= 0 || (sbo_bytes >> 4) > 0x3FFF) { ok = false; break; }
      int col = 0, chunk = 0;
      double cost = 0;
      for (const Slot& s : gp.slots) {
        WhSlot& ws = g.slot[g.n_slots++];
        const Tap& f = s.taps[0];
        const int tapoff_rows = (f.dh - hmin) * ww + (f.dw - wmin);
        int n, lbo16;
        if (xchunks == 1) {
          n = 64 * (int)s.taps.size();
          lbo16 = 128 >> 4;  // next 64-column chunk = same window one pixel further (the next tap)
          for (const Tap& t : s.taps) { g.chunk_widx[chunk] = t.widx; g.chunk_ci[chunk] = 0; ++chunk; }
        } else {
          n = block_ci;
          lbo16 = g.xchunk_bytes >> 4;
          for (int q = 0; q < xchunks; ++q) { g.chunk_widx[chunk] = f.widx; g.chunk_ci[chunk] = q; ++chunk; }
        }
        if (lbo16 > 0x3FFF) ok = false;
        ws.blo = ((uint32_t)lbo16 << 16) + (uint32_t)tapoff_rows * 8u;
        ws.bhi = smem_desc_hi((uint32_t)sbo_bytes);
        ws.idesc = make_idesc_bf16(128, n, 1, 1);
        ws.col = (uint32_t)col;
        col += n;
        cost += ksteps * mma_clocks(n);
      }
      g.cols = col;
      if (!ok || col > 512) { ok = false; break; }
      int xi = -1;
      for (size_t i = 0; i < xkeys.size(); ++i)
        if (xkeys[i].pair == g.pair && xkeys[i].ww == ww && xkeys[i].wh == wh) xi = (int)i;
      if (xi < 0) { xkeys.push_back({g.pair, ww, wh}); xi = (int)xkeys.size() - 1; }
      g.xmap = xi;
      max_stage = std::max(max_stage, kp.a_bytes + xchunks * g.xchunk_bytes);
      groups.push_back(g);
      group_cost.push_back(cost);
    }
    if (!ok || (int)xkeys.size() > kWhMaxXMaps) return nullptr;
    kp.stage_bytes = max_stage;
    kp.n_stages = std::min(kWhMaxStages, kWhSmemBudget / max_stage);
    if (kp.n_stages >= 2 || tile_px == 64 || env_tile_px) break;
    tile_px = 64;
  }
  if (kp.n_stages < 2) return nullptr;

  kp.tiles_w = (d->gW + kp.bw - 1) / kp.bw;
  kp.tiles_h = (d->gH + kp.bh - 1) / kp.bh;
  const int tiles_n = (d->gN + kp.bn - 1) / kp.bn;
  const int k_chunks = kp.tiles_w * kp.tiles_h * tiles_n;
  kp.xchunks = xchunks;
  kp.block_ci = block_ci;
  const int m_tiles = (d->w_cout + 127) / 128;
  const int n_tiles = (d->w_cin + block_ci - 1) / block_ci;
  kp.dw = reinterpret_cast<float*>(d->dw);
  kp.w_cout = d->w_cout; kp.w_taps = d->w_taps; kp.w_cin = d->w_cin;

  // ---- schedule: one segment list per CTA, equal MMA time
  std::vector<Item> items;
  for (int g = 0; g < (int)groups.size(); ++g)
    for (int n = 0; n < n_tiles; ++n)
      for (int m = 0; m < m_tiles; ++m) items.push_back({m, n, g, group_cost[g]});
  const int P = env_ctas > 0 ? std::min(env_ctas, num_sms()) : num_sms();
  std::vector<std::vector<WhSeg>> per_cta;
  double total = 0, max_cost = 0;
  for (const Item& it : items) { total += it.cost * k_chunks; max_cost = std::max(max_cost, it.cost); }
  auto push_seg = [&](std::vector<WhSeg>& v, const Item& it, int c0, int c1) {
    if (c1 <= c0) return;
    const int atomic = (c0 != 0 || c1 != k_chunks || d->accumulate) ? 1 : 0;
    v.push_back({it.m, it.n, it.g, c0, c1, atomic});
  };
  bool proportional = false;
  std::vector<int> share(items.size(), 0);
  if (d->ksplit <= 0 && (int)items.size() <= P) {
    // every item gets CTAs in proportion to its cost; items of equal cost get the same pixel ranges, so the CTAs of
    // different tap groups walk the same dY / X tiles at the same time (they meet in L2)
    double worst = 0;
    for (size_t i = 0; i < items.size(); ++i) {
      int s = (int)(P * items[i].cost * k_chunks / total);
      s = std::max(1, std::min(s, std::max(1, k_chunks / 2)));
      share[i] = s;
      worst = std::max(worst, items[i].cost * ((k_chunks + s - 1) / s));
    }
    proportional = (total / P) / worst >= 0.85;
  }
  if (d->ksplit > 0) {
    // explicit split-K (tests): ksplit CTAs per item
    const int ks = std::min(d->ksplit, k_chunks);
    for (const Item& it : items)
      for (int s = 0; s < ks; ++s) {
        per_cta.emplace_back();
        push_seg(per_cta.back(), it, (int)((int64_t)k_chunks * s / ks), (int)((int64_t)k_chunks * (s + 1) / ks));
      }
  } else if (proportional) {
    for (size_t i = 0; i < items.size(); ++i)
      for (int s = 0; s < share[i]; ++s) {
        per_cta.emplace_back();
        push_seg(per_cta.back(), items[i], (int)((int64_t)k_chunks * s / share[i]), (int)((int64_t)k_chunks * (s + 1) / share[i]));
      }
  } else {
    // stream-K: cut the linearised (item, pixel tile) cost axis into P equal spans
    per_cta.resize(P);
    size_t item = 0;
    int chunk = 0;
    double done = 0;  // cost before (item, chunk)
    for (int j = 0; j < P && item < items.size(); ++j) {
      const double until = total * (j + 1) / P;
      while (item < items.size()) {
        const Item& it = items[item];
        const double item_end = done + it.cost * (k_chunks - chunk);
        if (item_end <= until + 1e-6 || j == P - 1) {
          push_seg(per_cta[j], it, chunk, k_chunks);
          done = item_end;
          ++item;
          chunk = 0;
          continue;
        }
        int take = (int)((until - done) / it.cost);
        // do not leave slivers: a segment costs a full accumulator drain
        if (take < 2) take = 0;
        if (k_chunks - (chunk + take) < 2) take = k_chunks - chunk;
        push_seg(per_cta[j], it, chunk, chunk + take);
        done += it.cost * take;
        chunk += take;
        if (chunk == k_chunks) { ++item; chunk = 0; }
        break;
      }
    }
  }
  std::vector<WhSeg> segs;
  std::vector<int> cta_seg;
  for (auto& v : per_cta) {
    if (v.empty()) continue;
    cta_seg.push_back((int)segs.size());
    segs.insert(segs.end(), v.begin(), v.end());
  }
  cta_seg.push_back((int)segs.size());
  const int n_ctas = (int)cta_seg.size() - 1;
  if (n_ctas < 1) return nullptr;

  WgradHaloLaunch* L = new WgradHaloLaunch();
  *hard_error = true;  // from here on a failure is a real error (set_error has been called)
  for (int i = 0; i < d->n_pair; ++i)
    if (encode_act_map(&kp.ymap[i], d->dy[i], 64, kp.bw, kp.bh, kp.bn) != 0) { delete L; return nullptr; }
  for (int i = d->n_pair; i < B2SEG_MAX_SRC; ++i) kp.ymap[i] = kp.ymap[0];
  for (size_t i = 0; i < xkeys.size(); ++i)
    if (encode_act_map(&kp.xmap[i], d->x[xkeys[i].pair], 64, xkeys[i].ww, xkeys[i].wh, kp.bn) != 0) { delete L; return nullptr; }
  for (size_t i = xkeys.size(); i < (size_t)kWhMaxXMaps; ++i) kp.xmap[i] = kp.xmap[0];
  for (size_t i = 0; i < groups.size(); ++i) kp.groups[i] = groups[i];
  const size_t gbytes = 0;
  const size_t sbytes = (segs.size() * sizeof(WhSeg) + 255) / 256 * 256;
  const size_t cbytes = cta_seg.size() * sizeof(int);
  std::vector<uint8_t> host(gbytes + sbytes + cbytes, 0);
  memcpy(host.data() + gbytes, segs.data(), segs.size() * sizeof(WhSeg));
  memcpy(host.data() + gbytes + sbytes, cta_seg.data(), cbytes);
  if (cudaMalloc(&L->d_tables, host.size()) != cudaSuccess ||
      cudaMemcpy(L->d_tables, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("wgrad_halo: table upload failed");
    delete L;
    return nullptr;
  }
  uint8_t* base = reinterpret_cast<uint8_t*>(L->d_tables);
  kp.segs = reinterpret_cast<const WhSeg*>(base + gbytes);
  kp.cta_seg = reinterpret_cast<const int*>(base + gbytes + sbytes);
  L->kp = kp;
  L->grid = n_ctas;
  L->ksteps = ksteps;
  L->smem_bytes = 1024 + kp.n_stages * kp.stage_bytes + kWhTailBytes;
  *hard_error = false;
  return L;
}

}  // namespace b2
