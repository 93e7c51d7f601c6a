// Implicit-GEMM convolution for sm_100a, generic path: one 4-D TMA box per (tap, channel block) (zero OOB fill = SAME
// padding) -> 128B-swizzled smem -> tcgen05.mma (M=128, N=BLOCK_N, K=16, bf16 -> fp32 in TMEM) -> shared epilogue.
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue; two TMEM
// accumulators so the epilogue of tile i overlaps the main loop of tile i+1.
// Dense k x k windows on large feature maps go through conv_halo.cu instead (one halo tile reused by every tap).
//
// Replaces tf.keras.layers.Conv2D/Conv1D (reference TensorFlow/2DCNN/models/unet_variants.py:9,
// TensorFlow/1DCNN/Models/unet_variants.py:55), Conv2DTranspose/Conv1DTranspose (:19 / :104) and their
// input-gradient kernels (TF Conv2DBackpropInput) — see include/b2seg.h.
#include "conv_common.cuh"

namespace b2 {

constexpr int kAStageBytes = kBlockM * kBlockK * 2;  // 16 KiB

struct alignas(64) ConvKParams {
  CUtensorMap amap[B2SEG_MAX_SRC];
  CUtensorMap bmap;
  int4 taps[B2SEG_MAX_TAPS];  // x = src, y = dh, z = dw, w = widx
  int taps_per_group, kc_blocks;
  ConvEpiParams e;
};

template <int BLOCK_N>
struct ConvCfg {
  static constexpr int kBStageBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kAStageBytes + kBStageBytes;
  static constexpr int kStages = BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kStgBytes + kBarBytes + kColPartBytes;
};

template <int BLOCK_N, bool B_MN>
__global__ void __launch_bounds__(kConvThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvKParams p) {
  using Cfg = ConvCfg<BLOCK_N>;
  pdl_launch_dependents();   // the next kernel of the stream may become resident as SMs drain
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + kStgBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* colpart = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + Cfg::kBarBytes);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler (uniform datapath)
  const int lane = threadIdx.x & 31;
  const ConvEpiParams& e = p.e;

  if (threadIdx.x == 0) {
    for (int i = 0; i < B2SEG_MAX_SRC; ++i) tma_prefetch_desc(&p.amap[i]);
    tma_prefetch_desc(&p.bmap);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], kEpiWarps);
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
  pdl_wait();                // barriers / TMEM are set up; from here on global memory of earlier kernels is read
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  const int nkb = p.taps_per_group * p.kc_blocks;
  const TileRange tr = tile_range(e);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = tr.first; tile < tr.end; tile += tr.step) {
        const TileCoord tc = tile_coord(e, tile);
        const int n_tile = tc.n_tile, g = tc.g, w0 = tc.w0, h0 = tc.h0, n0 = tc.n0;
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
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues:
    // warp-uniform control flow keeps the descriptors in uniform registers, see conv_halo.cu)
    {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 0, B_MN ? 1 : 0);
      // descriptors are built once; per MMA only the 14-bit start-address field (16-byte units) changes
      const uint64_t adesc0 = make_smem_desc(0, 16, 1024);
      const uint64_t bdesc0 = B_MN ? make_smem_desc(0, 8192, 1024) : make_smem_desc(0, 16, 1024);
      const uint32_t smem16 = smem_u32(smem) >> 4;
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = tr.first; tile < tr.end; tile += tr.step) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a16 = smem16 + stage * (Cfg::kStageBytes >> 4);
          const uint64_t ad = adesc0 + a16;
          const uint64_t bd = bdesc0 + (a16 + (kAStageBytes >> 4));
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16_elect(d_tmem, ad + k * 2, bd + k * (B_MN ? 128 : 2), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_elect(&empty_bar[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_elect(&tfull_bar[acc]);
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

int select_block_n(const b2seg_conv_desc* d) {
  const int n_extent = d->out[0].C;
  int bn_sel = d->block_n;
  if (bn_sel == 0) bn_sel = n_extent <= 64 ? 64 : (n_extent <= 128 ? 128 : 256);
  return bn_sel;
}

// Fills the tile-scheduler / epilogue parameters shared by both kernels.  Geometry (bw, bh, bn) is chosen by the caller.
int fill_epi_params(const b2seg_conv_desc* d, int block_n, int bw, int bh, int bn, ConvEpiParams* e) {
  const b2seg_view& o = d->out[0];
  memset(e, 0, sizeof(*e));
  e->bw = bw; e->bh = bh; e->bn = bn;
  e->tiles_w = (o.W + bw - 1) / bw;
  e->tiles_h = (o.H + bh - 1) / bh;
  e->tiles_n = (o.N + bn - 1) / bn;
  e->m_tiles = e->tiles_w * e->tiles_h * e->tiles_n;
  e->gN = o.N; e->gH = o.H; e->gW = o.W;
  e->n_extent = o.C;
  e->n_tiles = (o.C + block_n - 1) / block_n;
  e->n_groups = d->n_groups;
  e->total_tiles = e->n_groups * e->m_tiles * e->n_tiles;
  e->fd_n_tiles = make_fastdiv((uint32_t)e->n_tiles);
  e->fd_m_tiles = make_fastdiv((uint32_t)e->m_tiles);
  e->fd_tiles_w = make_fastdiv((uint32_t)e->tiles_w);
  e->fd_tiles_h = make_fastdiv((uint32_t)e->tiles_h);
  for (int g = 0; g < d->n_groups; ++g) {
    const b2seg_view& og = d->out[g];
    if (og.N != o.N || og.H != o.H || og.W != o.W || og.C != o.C || og.sn != o.sn || og.sh != o.sh || og.sw != o.sw)
      return fail(B2SEG_ERR_ARG, "conv: group outputs must share geometry");
    e->out_ptr[g] = og.ptr;
  }
  e->out_sn = o.sn; e->out_sh = o.sh; e->out_sw = o.sw;
  e->bias = reinterpret_cast<const float*>(d->bias);
  e->act = d->act;
  e->stats = reinterpret_cast<float*>(d->stats);
  e->stats_atomic = d->stats_atomic;
  e->mul_mode = d->mul_mode;
  if (d->mul_mode != 0) {
    e->mul_ptr = d->mul_view.ptr;
    e->mul_sn = d->mul_view.sn; e->mul_sh = d->mul_view.sh; e->mul_sw = d->mul_view.sw;
    e->mul_c = d->mul_view.C;
  }
  {
    auto span = [&](long long sn, long long sh, long long sw) { return (bn - 1) * llabs(sn) + (bh - 1) * llabs(sh) + (bw - 1) * llabs(sw); };
    if (span(o.sn, o.sh, o.sw) >= (1ll << 31) || (d->mul_mode != 0 && span(d->mul_view.sn, d->mul_view.sh, d->mul_view.sw) >= (1ll << 31)))
      return fail(B2SEG_ERR_ARG, "conv: a tile spans more than 2^31 elements of the output view");
  }
  e->lbw = 0; while ((1 << e->lbw) < bw) ++e->lbw;
  e->lbwh = 0; while ((1 << e->lbwh) < bw * bh) ++e->lbwh;
  e->stats_per_cta = (e->n_tiles == 1) ? 1 : 0;
  return 0;
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
  B2_CUDA_OK(launch_k(conv_gemm_kernel<BLOCK_N, B_MN>, dim3(grid), dim3(kConvThreads), Cfg::kSmemBytes, s, kp));
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

int conv_num_mtiles(const b2seg_conv_desc* d) {
  int bw, bh, bn;
  const b2seg_view& o = d->out[0];
  if (halo_geometry(d, &bw, &bh, &bn) != 0) pick_box(o.N, o.H, o.W, kBlockM, &bw, &bh, &bn);
  return ((o.W + bw - 1) / bw) * ((o.H + bh - 1) / bh) * ((o.N + bn - 1) / bn);
}

// rows of the statistics buffer the kernel writes: one per CTA when a CTA always covers the same columns
int conv_num_stat_rows(const b2seg_conv_desc* d) {
  if (d->stats_atomic) return 1;
  const int m_tiles = conv_num_mtiles(d);
  const int bn_sel = select_block_n(d);
  const int n_tiles = (d->out[0].C + bn_sel - 1) / bn_sel;
  const int total = d->n_groups * m_tiles * n_tiles;
  const int sms = num_sms();
  return n_tiles == 1 ? (total < sms ? total : sms) : d->n_groups * m_tiles;
}

static bool check_conv_desc(const b2seg_conv_desc* d) {
  if (d->n_src < 1 || d->n_src > B2SEG_MAX_SRC || d->n_groups < 1 || d->n_groups > B2SEG_MAX_GROUPS ||
      d->taps_per_group < 1 || d->n_groups * d->taps_per_group > B2SEG_MAX_TAPS) {
    set_error("conv: bad src/group/tap counts");
    return false;
  }
  if (d->out[0].C % 8 != 0 || d->w_cin % 8 != 0) {
    set_error("conv: channel extents must be multiples of 8 (out.C=%d w_cin=%d)", d->out[0].C, d->w_cin);
    return false;
  }
  const int bn_sel = select_block_n(d);
  if (bn_sel != 64 && bn_sel != 128 && bn_sel != 256) {
    set_error("conv: block_n must be 64/128/256");
    return false;
  }
  for (int t = 0; t < d->n_groups * d->taps_per_group; ++t) {
    const b2seg_tap& tp = d->taps[t];
    if (tp.src < 0 || tp.src >= d->n_src || tp.widx < 0 || tp.widx >= d->w_taps) {
      set_error("conv: tap %d out of range", t);
      return false;
    }
  }
  return true;
}

PreparedOp* prepare_conv(const b2seg_conv_desc* d) {
  if (!check_conv_desc(d)) return nullptr;
  int hb[3];
  if (halo_geometry(d, &hb[0], &hb[1], &hb[2]) == 0) return prepare_conv_halo(d);  // dense windows on large maps
  const b2seg_view& o = d->out[0];
  ConvLaunch* L = new ConvLaunch();
  ConvKParams& kp = L->kp;
  memset(&kp, 0, sizeof(kp));
  L->b_mn = d->b_mn_major != 0;
  L->block_n = select_block_n(d);
  int bw, bh, bn;
  pick_box(o.N, o.H, o.W, kBlockM, &bw, &bh, &bn);
  if (fill_epi_params(d, L->block_n, bw, bh, bn, &kp.e) != 0) { delete L; return nullptr; }
  kp.taps_per_group = d->taps_per_group;
  const int k_ch = L->b_mn ? d->w_cout : d->w_cin;  // GEMM-K channels per tap = channels of the A sources
  kp.kc_blocks = (k_ch + kBlockK - 1) / kBlockK;
  for (int i = 0; i < d->n_src; ++i) {
    if (encode_act_map(&kp.amap[i], d->src[i], kBlockK, bw, bh, bn) != 0) { delete L; return nullptr; }
  }
  for (int i = d->n_src; i < B2SEG_MAX_SRC; ++i) kp.amap[i] = kp.amap[0];
  if (!L->b_mn) {
    if (encode_weight_map(&kp.bmap, d->weights, d->w_cout, d->w_taps, d->w_cin, kBlockK, L->block_n) != 0) { delete L; return nullptr; }
  } else {
    if (encode_weight_map(&kp.bmap, d->weights, d->w_cout, d->w_taps, d->w_cin, 64, kBlockK) != 0) { delete L; return nullptr; }
  }
  for (int t = 0; t < d->n_groups * d->taps_per_group; ++t) {
    const b2seg_tap& tp = d->taps[t];
    kp.taps[t] = make_int4(tp.src, tp.dh, tp.dw, tp.widx);
  }
  const int sms = num_sms();
  L->grid = kp.e.total_tiles < sms ? kp.e.total_tiles : sms;
  return L;
}

}  // namespace b2
