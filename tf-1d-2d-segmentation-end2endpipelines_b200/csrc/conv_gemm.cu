// Implicit-GEMM convolution for sm_100a: TMA (4-D tiled, zero OOB fill = SAME padding) -> 128B-swizzled smem
// -> tcgen05.mma (M=128, N=BLOCK_N, K=16, bf16 -> fp32 in TMEM) -> epilogue (bias, activation, BN statistics,
// strided / concat-slot store).  Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer,
// warps 2..5 = epilogue; two TMEM accumulators so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Replaces tf.keras.layers.Conv2D/Conv1D (reference TensorFlow/2DCNN/models/unet_variants.py:9,
// TensorFlow/1DCNN/Models/unet_variants.py:55), Conv2DTranspose/Conv1DTranspose (:19 / :104) and their
// input-gradient kernels (TF Conv2DBackpropInput) — see include/b2seg.h.
#include "common.h"
#include "ptx.cuh"

namespace b2 {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kAStageBytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kStgPitch = 144;                        // staging row pitch (64 bf16 + 16 B pad: conflict-free 16 B stores)
constexpr int kStgBytes = kBlockM * kStgPitch;
constexpr int kConvThreads = 192;

struct alignas(64) ConvKParams {
  CUtensorMap amap[B2SEG_MAX_SRC];
  CUtensorMap bmap;
  int4 taps[B2SEG_MAX_TAPS];  // x = src, y = dh, z = dw, w = widx
  int n_groups, taps_per_group, kc_blocks;
  int gN, gH, gW;
  int bw, bh, bn;
  int tiles_w, tiles_h, tiles_n, m_tiles, n_tiles, total_tiles;
  int n_extent;
  unsigned long long out_ptr[B2SEG_MAX_GROUPS];
  long long out_sn, out_sh, out_sw;
  const float* bias;
  int act;
  float* stats;
  unsigned long long mul_ptr;
  long long mul_sn, mul_sh, mul_sw;
  int mul_mode, mul_c;
  int lbw, lbwh;       // log2(bw), log2(bw*bh): tile rows -> pixel coordinates by shifts
  int stats_per_cta;   // 1: one statistics row per CTA (n_tiles == 1), else one per (group, m_tile)
};

template <int BLOCK_N>
struct ConvCfg {
  static constexpr int kBStageBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kAStageBytes + kBStageBytes;
  static constexpr int kStages = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kColPartBytes = 2 * 64 * 2 * 4;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kStgBytes + kBarBytes + kColPartBytes;
};

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case B2SEG_ACT_RELU: return fmaxf(x, 0.f);
    case B2SEG_ACT_LEAKY: return x > 0.f ? x : 0.3f * x;
    case B2SEG_ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
    default: return x;
  }
}

template <int BLOCK_N, bool B_MN>
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvKParams p) {
  using Cfg = ConvCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kStgBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* colpart = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + Cfg::kBarBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < B2SEG_MAX_SRC; ++i) tma_prefetch_desc(&p.amap[i]);
    tma_prefetch_desc(&p.bmap);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nkb = p.taps_per_group * p.kc_blocks;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int rest = tile / p.n_tiles;
        const int m_tile = rest % p.m_tiles;
        const int g = rest / p.m_tiles;
        const int w0 = (m_tile % p.tiles_w) * p.bw;
        const int h0 = ((m_tile / p.tiles_w) % p.tiles_h) * p.bh;
        const int n0 = (m_tile / (p.tiles_w * p.tiles_h)) * p.bn;
        for (int t = 0; t < p.taps_per_group; ++t) {
          const int4 tap = p.taps[g * p.taps_per_group + t];
          for (int cb = 0; cb < p.kc_blocks; ++cb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + kAStageBytes;
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_4d(&p.amap[tap.x], &full_bar[stage], sa, cb * kBlockK, w0 + tap.z, h0 + tap.y, n0);
            if (!B_MN) {
              tma_load_3d(&p.bmap, &full_bar[stage], sb, cb * kBlockK, tap.w, n_tile * BLOCK_N);
            } else {
#pragma unroll
              for (int q = 0; q < BLOCK_N / 64; ++q)
                tma_load_3d(&p.bmap, &full_bar[stage], sb + q * 8192, n_tile * BLOCK_N + q * 64, tap.w, cb * kBlockK);
            }
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 0, B_MN ? 1 : 0);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + kAStageBytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t adesc = make_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc(b_addr + k * 2048, 8192, 1024)
                                        : make_smem_desc(b_addr + k * 32, 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (128 threads)
    const int et = threadIdx.x - 64;           // 0..127
    const int row = (warp & 3) * 32 + lane;    // TMEM lane == tile row owned by this thread
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int bwm = p.bw - 1, bhm = p.bh - 1;
    constexpr int kChunks = BLOCK_N / 64;
    float cta_s[kChunks], cta_q[kChunks];      // per-CTA BN statistics (threads et < 64)
#pragma unroll
    for (int c = 0; c < kChunks; ++c) { cta_s[c] = 0.f; cta_q[c] = 0.f; }
    const int vq = et & 7, r0 = et >> 3;       // this thread's 16-byte column slot and first row of the store pass
    uint32_t acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      const int rest = tile / p.n_tiles;
      const int m_tile = rest % p.m_tiles;
      const int g = rest / p.m_tiles;
      const int w0 = (m_tile % p.tiles_w) * p.bw;
      const int h0 = ((m_tile / p.tiles_w) % p.tiles_h) * p.bh;
      const int n0 = (m_tile / (p.tiles_w * p.tiles_h)) * p.bn;
      const bool my_valid = (n0 + (row >> p.lbwh)) < p.gN && (h0 + ((row >> p.lbw) & bhm)) < p.gH && (w0 + (row & bwm)) < p.gW;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int col0 = n_tile * BLOCK_N + c * 64;
        const int cc_st = col0 + vq * 8;
        // dgrad fusion: fetch the forward activations whose sign masks this chunk early, so the loads overlap the TMEM read
        uint4 yv[8];
        if (p.mul_mode != 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = r0 + 16 * i;
            const int pn = n0 + (r >> p.lbwh), ph = h0 + ((r >> p.lbw) & bhm), pw = w0 + (r & bwm);
            yv[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);  // bf16 1.0 -> derivative 1
            if (pn < p.gN && ph < p.gH && pw < p.gW && cc_st < p.mul_c)
              yv[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.mul_ptr) + pn * p.mul_sn + ph * p.mul_sh + pw * p.mul_sw + cc_st));
          }
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          tmem_ld32(tmem_base + lane_base + acc * BLOCK_N + c * 64 + half * 32, v);
          tmem_ld_wait();
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int cc = col0 + half * 32 + 2 * j;
            float x0 = __uint_as_float(v[2 * j]);
            float x1 = __uint_as_float(v[2 * j + 1]);
            if (p.bias != nullptr) {
              if (cc < p.n_extent) x0 += __ldg(p.bias + cc);
              if (cc + 1 < p.n_extent) x1 += __ldg(p.bias + cc + 1);
            }
            x0 = apply_act(x0, p.act);
            x1 = apply_act(x1, p.act);
            if (!my_valid) { x0 = 0.f; x1 = 0.f; }
            packed[j] = pack_bf16x2(x0, x1);
          }
          uint4* dst = reinterpret_cast<uint4*>(staging + row * kStgPitch + half * 64);
#pragma unroll
          for (int q = 0; q < 4; ++q) dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        }
        if (c == kChunks - 1) {
          // accumulator fully read: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        named_bar_sync(1, 128);
        // ---- coalesced store of the 128 x 64 chunk (+ optional derivative-mask multiply)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 16 * i;
          const int pn = n0 + (r >> p.lbwh), ph = h0 + ((r >> p.lbw) & bhm), pw = w0 + (r & bwm);
          if (pn < p.gN && ph < p.gH && pw < p.gW && cc_st < p.n_extent) {
            uint4 val = *reinterpret_cast<const uint4*>(staging + r * kStgPitch + vq * 16);
            if (p.mul_mode != 0) {
              const __nv_bfloat16* ye = reinterpret_cast<const __nv_bfloat16*>(&yv[i]);
              __nv_bfloat16* ve = reinterpret_cast<__nv_bfloat16*>(&val);
              const float neg = p.mul_mode == B2SEG_ACT_LEAKY ? 0.3f : 0.f;
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (!(__bfloat162float(ye[e]) > 0.f)) ve[e] = __float2bfloat16(__bfloat162float(ve[e]) * neg);
            }
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out_ptr[g]) + pn * p.out_sn + ph * p.out_sh + pw * p.out_sw + cc_st;
            *reinterpret_cast<uint4*>(op) = val;
          }
        }
        // ---- BatchNorm statistics of the stored values: column sum / sum of squares
        if (p.stats != nullptr) {
          const int col = et & 63, hf = et >> 6;
          float s = 0.f, ss = 0.f;
          const uint8_t* sp = staging + (hf * 64) * kStgPitch + col * 2;
#pragma unroll 16
          for (int r = 0; r < 64; ++r) {
            const float x = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(sp + r * kStgPitch));
            s += x;
            ss += x * x;
          }
          colpart[(hf * 64 + col) * 2 + 0] = s;
          colpart[(hf * 64 + col) * 2 + 1] = ss;
          named_bar_sync(2, 128);
          if (et < 64) {
            const float s2 = colpart[et * 2] + colpart[(64 + et) * 2];
            const float ss2 = colpart[et * 2 + 1] + colpart[(64 + et) * 2 + 1];
            if (p.stats_per_cta) {
              cta_s[c] += s2;
              cta_q[c] += ss2;
            } else if (col0 + et < p.n_extent) {
              float* st = p.stats + (size_t)(g * p.m_tiles + m_tile) * 2 * p.n_extent;
              st[col0 + et] = s2;
              st[p.n_extent + col0 + et] = ss2;
            }
          }
        }
        named_bar_sync(1, 128);  // staging is reused by the next chunk
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (p.stats != nullptr && p.stats_per_cta && et < 64) {
      float* st = p.stats + (size_t)blockIdx.x * 2 * p.n_extent;
#pragma unroll
      for (int c = 0; c < kChunks; ++c)
        if (c * 64 + et < p.n_extent) {
          st[c * 64 + et] = cta_s[c];
          st[p.n_extent + c * 64 + et] = cta_q[c];
        }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
void pick_box(int N, int H, int W, int pixels, int* bw, int* bh, int* bn) {
  auto pow2_cover = [](int x, int cap) {  // smallest power of two >= x, capped
    int p = 1;
    while (p < x && p < cap) p <<= 1;
    return p;
  };
  auto best = [&](int extent, int cap) {
    // largest power of two <= cap dividing extent if it is at least 8 (or the whole extent), else a cover
    int d = 1;
    while (d * 2 <= cap && extent % (d * 2) == 0) d <<= 1;
    if (d >= 8 || d == extent) return d;
    return pow2_cover(extent, cap);
  };
  int w = best(W, pixels);
  int h = best(H, pixels / w);
  int n = pixels / (w * h);
  *bw = w; *bh = h; *bn = n;
}

struct ConvLaunch : PreparedOp {
  ConvKParams kp;
  int block_n;
  bool b_mn;
  int grid;
  int launch(cudaStream_t s) override;
};

template <int BLOCK_N, bool B_MN>
static int launch_conv_t(const ConvKParams& kp, int grid, cudaStream_t s) {
  using Cfg = ConvCfg<BLOCK_N>;
  static bool attr_set = false;
  if (!attr_set) {
    B2_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  conv_gemm_kernel<BLOCK_N, B_MN><<<grid, kConvThreads, Cfg::kSmemBytes, s>>>(kp);
  B2_CUDA_OK(cudaGetLastError());
  return 0;
}

int ConvLaunch::launch(cudaStream_t s) {
  if (!b_mn) {
    if (block_n == 64) return launch_conv_t<64, false>(kp, grid, s);
    if (block_n == 128) return launch_conv_t<128, false>(kp, grid, s);
    return launch_conv_t<256, false>(kp, grid, s);
  }
  if (block_n == 64) return launch_conv_t<64, true>(kp, grid, s);
  if (block_n == 128) return launch_conv_t<128, true>(kp, grid, s);
  return launch_conv_t<256, true>(kp, grid, s);
}

static int conv_geometry(const b2seg_conv_desc* d, int* bw, int* bh, int* bn, int* tw, int* th, int* tn) {
  const b2seg_view& o = d->out[0];
  pick_box(o.N, o.H, o.W, kBlockM, bw, bh, bn);
  *tw = (o.W + *bw - 1) / *bw;
  *th = (o.H + *bh - 1) / *bh;
  *tn = (o.N + *bn - 1) / *bn;
  return (*tw) * (*th) * (*tn);
}

int conv_num_mtiles(const b2seg_conv_desc* d) {
  int bw, bh, bn, tw, th, tn;
  return conv_geometry(d, &bw, &bh, &bn, &tw, &th, &tn);
}

PreparedOp* prepare_conv(const b2seg_conv_desc* d) {
  if (d->n_src < 1 || d->n_src > B2SEG_MAX_SRC || d->n_groups < 1 || d->n_groups > B2SEG_MAX_GROUPS ||
      d->taps_per_group < 1 || d->n_groups * d->taps_per_group > B2SEG_MAX_TAPS) {
    set_error("conv: bad src/group/tap counts");
    return nullptr;
  }
  const b2seg_view& o = d->out[0];
  if (o.C % 8 != 0 || d->w_cin % 8 != 0) {
    set_error("conv: channel extents must be multiples of 8 (out.C=%d w_cin=%d)", o.C, d->w_cin);
    return nullptr;
  }
  ConvLaunch* L = new ConvLaunch();
  ConvKParams& kp = L->kp;
  memset(&kp, 0, sizeof(kp));
  L->b_mn = d->b_mn_major != 0;
  const int n_extent = o.C;
  int bn_sel = d->block_n;
  if (bn_sel == 0) bn_sel = n_extent <= 64 ? 64 : (n_extent <= 128 ? 128 : 256);
  if (bn_sel != 64 && bn_sel != 128 && bn_sel != 256) {
    set_error("conv: block_n must be 64/128/256");
    delete L;
    return nullptr;
  }
  L->block_n = bn_sel;
  int tw, th, tn;
  kp.m_tiles = conv_geometry(d, &kp.bw, &kp.bh, &kp.bn, &tw, &th, &tn);
  kp.tiles_w = tw; kp.tiles_h = th; kp.tiles_n = tn;
  kp.gN = o.N; kp.gH = o.H; kp.gW = o.W;
  kp.n_extent = n_extent;
  kp.n_tiles = (n_extent + bn_sel - 1) / bn_sel;
  kp.n_groups = d->n_groups;
  kp.taps_per_group = d->taps_per_group;
  kp.total_tiles = kp.n_groups * kp.m_tiles * kp.n_tiles;
  const int k_ch = L->b_mn ? d->w_cout : d->w_cin;  // GEMM-K channels per tap = channels of the A sources
  kp.kc_blocks = (k_ch + kBlockK - 1) / kBlockK;
  for (int i = 0; i < d->n_src; ++i) {
    if (encode_act_map(&kp.amap[i], d->src[i], kBlockK, kp.bw, kp.bh, kp.bn) != 0) { delete L; return nullptr; }
  }
  for (int i = d->n_src; i < B2SEG_MAX_SRC; ++i) kp.amap[i] = kp.amap[0];
  if (!L->b_mn) {
    if (encode_weight_map(&kp.bmap, d->weights, d->w_cout, d->w_taps, d->w_cin, kBlockK, bn_sel) != 0) { delete L; return nullptr; }
  } else {
    if (encode_weight_map(&kp.bmap, d->weights, d->w_cout, d->w_taps, d->w_cin, 64, kBlockK) != 0) { delete L; return nullptr; }
  }
  for (int t = 0; t < d->n_groups * d->taps_per_group; ++t) {
    const b2seg_tap& tp = d->taps[t];
    if (tp.src < 0 || tp.src >= d->n_src || tp.widx < 0 || tp.widx >= d->w_taps) {
      set_error("conv: tap %d out of range", t);
      delete L;
      return nullptr;
    }
    kp.taps[t] = make_int4(tp.src, tp.dh, tp.dw, tp.widx);
  }
  for (int g = 0; g < d->n_groups; ++g) {
    const b2seg_view& og = d->out[g];
    if (og.N != o.N || og.H != o.H || og.W != o.W || og.C != o.C || og.sn != o.sn || og.sh != o.sh || og.sw != o.sw) {
      set_error("conv: group outputs must share geometry");
      delete L;
      return nullptr;
    }
    kp.out_ptr[g] = og.ptr;
  }
  kp.out_sn = o.sn; kp.out_sh = o.sh; kp.out_sw = o.sw;
  kp.bias = reinterpret_cast<const float*>(d->bias);
  kp.act = d->act;
  kp.stats = reinterpret_cast<float*>(d->stats);
  kp.mul_mode = d->mul_mode;
  if (d->mul_mode != 0) {
    kp.mul_ptr = d->mul_view.ptr;
    kp.mul_sn = d->mul_view.sn; kp.mul_sh = d->mul_view.sh; kp.mul_sw = d->mul_view.sw;
    kp.mul_c = d->mul_view.C;
  }
  const int sms = num_sms();
  L->grid = kp.total_tiles < sms ? kp.total_tiles : sms;
  kp.lbw = 0; while ((1 << kp.lbw) < kp.bw) ++kp.lbw;
  kp.lbwh = 0; while ((1 << kp.lbwh) < kp.bw * kp.bh) ++kp.lbwh;
  kp.stats_per_cta = (kp.n_tiles == 1) ? 1 : 0;
  return L;
}

// rows of the statistics buffer the kernel writes: one per CTA when a CTA always covers the same columns
int conv_num_stat_rows(const b2seg_conv_desc* d) {
  int bw, bh, bn, tw, th, tn;
  const int m_tiles = conv_geometry(d, &bw, &bh, &bn, &tw, &th, &tn);
  const int n_extent = d->out[0].C;
  int bn_sel = d->block_n;
  if (bn_sel == 0) bn_sel = n_extent <= 64 ? 64 : (n_extent <= 128 ? 128 : 256);
  const int n_tiles = (n_extent + bn_sel - 1) / bn_sel;
  const int total = d->n_groups * m_tiles * n_tiles;
  const int sms = num_sms();
  return n_tiles == 1 ? (total < sms ? total : sms) : d->n_groups * m_tiles;
}

}  // namespace b2
