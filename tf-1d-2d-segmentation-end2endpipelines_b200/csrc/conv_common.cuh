// Shared pieces of the implicit-GEMM convolution kernels (conv_gemm.cu: one TMA box per tap; conv_halo.cu: one halo
// tile per channel block reused by every tap): tile geometry, epilogue parameters and the epilogue itself
// (TMEM -> registers -> bias/activation -> bf16 staging in smem -> coalesced, strided store + BN statistics).
//
// The epilogue is written to be cheap in *issued instructions* (round-1 ncu: the first version spent ~3000 SASS
// instructions per warp per 128x64 tile and capped the high-resolution layers at 13 % tensor-pipe activity):
// bias comes from shared memory by 128-bit broadcast loads, bounds checks are hoisted to per-tile / per-chunk
// uniform flags, output addresses are tile-invariant offsets plus one per-tile base, and the BatchNorm statistics
// are accumulated from the registers of the store pass and reduced with two shuffle steps.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace b2 {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kStgPitch = 144;                        // staging row pitch (64 bf16 + 16 B pad: conflict-free 16 B stores)
constexpr int kStgBytes = kBlockM * kStgPitch;
#ifndef B2_EPI_WARPS
#define B2_EPI_WARPS 16
#endif
// Epilogue warps: kEpiWarps / 4 per TMEM lane quadrant, each converting kEpiCols of the 64 columns of a chunk.  ncu (source
// view, profiles/r1_epilogue_stalls.txt): with 8 warps each scheduler holds 2 epilogue warps that issue one instruction
// every ~5 clocks (stalls: fixed-latency dependencies 30 %, shared-memory scoreboard 18 %, barrier 9 %) -- latency-bound, so
// the high-resolution layers took ~3300 clocks per 128x64 chunk whatever the MMA time.  16 warps double the warps per scheduler.
constexpr int kEpiWarps = B2_EPI_WARPS;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kEpiCols = 64 / (kEpiWarps / 4);        // columns of a chunk converted by one warp: 32 (8 warps) or 16 (16 warps)
constexpr int kEpiRows = kBlockM * 8 / kEpiThreads;   // rows per thread in the store pass: 4 or 2
constexpr int kEpiRowStep = kEpiThreads / 8;          // 32 or 64
constexpr int kConvThreads = 64 + kEpiThreads;        // warp 0 = TMA producer, warp 1 = MMA issuer, warps 2.. = epilogue
constexpr int kTileInfoBytes = 48;                    // per-tile scalars computed once by one epilogue thread (two slots)
constexpr int kColPartBytes = kEpiWarps * 64 * 2 * 4 + 256 * 4 + 2 * kTileInfoBytes + 32;  // [warps][64 cols][sum, sumsq] + bias tile [256] + tile info

// everything the tile scheduler and the epilogue need (embedded as `e` in each kernel's parameter struct)
struct ConvEpiParams {
  int n_groups;
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
  int stats_atomic;    // 1: every CTA adds its sums into ONE caller-zeroed row (red.add) instead of writing its own
  unsigned long long* trace;   // debug (B2SEG_TRACE=1): per-tile clock64() stamps of CTA 0, [tile][8]
  int cta_groups;      // G > 1: CTA b only walks the tiles of group b % G (its weights stay resident in shared memory); gridDim.x % G == 0
  FastDiv fd_n_tiles, fd_m_tiles, fd_tiles_w, fd_tiles_h;   // tile index -> coordinates without integer division
};

// the tile indices a CTA walks: all of them round-robin, or (cta_groups) those of its own group only, with the CTAs
// b, b+1, .. b+G-1 of a quad visiting the same m tiles in step so that the shared input tile is fetched from HBM once
struct TileRange { int first, end, step; };
__device__ __forceinline__ TileRange tile_range(const ConvEpiParams& p) {
  TileRange r;
  if (p.cta_groups > 1) {
    const int G = p.cta_groups, per_group = p.m_tiles * p.n_tiles, g = (int)blockIdx.x % G;
    r.first = g * per_group + (int)blockIdx.x / G;
    r.end = (g + 1) * per_group;
    r.step = (int)gridDim.x / G;
  } else {
    r.first = (int)blockIdx.x; r.end = p.total_tiles; r.step = (int)gridDim.x;
  }
  return r;
}

// tile -> (group, n_tile, m_tile, w0, h0, n0); the epilogue used to spend ~20 % of its instructions on these divisions
struct TileCoord { int g, n_tile, m_tile, w0, h0, n0; };
__device__ __forceinline__ TileCoord tile_coord(const ConvEpiParams& p, int tile) {
  TileCoord c;
  const uint32_t rest = fast_div((uint32_t)tile, p.fd_n_tiles);
  c.n_tile = tile - (int)rest * p.n_tiles;
  c.g = (int)fast_div(rest, p.fd_m_tiles);
  c.m_tile = (int)rest - c.g * p.m_tiles;
  const uint32_t q = fast_div((uint32_t)c.m_tile, p.fd_tiles_w);
  c.w0 = (c.m_tile - (int)q * p.tiles_w) * p.bw;
  const uint32_t q2 = fast_div(q, p.fd_tiles_h);
  c.h0 = ((int)q - (int)q2 * p.tiles_h) * p.bh;
  c.n0 = (int)q2 * p.bn;
  return c;
}

template <int ACT>
__device__ __forceinline__ float act_t(float x) {
  if (ACT == B2SEG_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == B2SEG_ACT_LEAKY) return fmaxf(x, 0.3f * x);
  if (ACT == B2SEG_ACT_SIGMOID) return 1.f / (1.f + __expf(-x));
  return x;
}

template <int ACT, int NC>
__device__ __forceinline__ void bias_act_pack(const uint32_t (&v)[NC], uint32_t sb_addr, uint32_t (&packed)[NC / 2]) {
#pragma unroll
  for (int j = 0; j < NC / 4; ++j) {
    const uint4 braw = lds128(sb_addr + 16 * j);  // same address in every lane: shared-memory broadcast
    const float4 b = make_float4(__uint_as_float(braw.x), __uint_as_float(braw.y), __uint_as_float(braw.z), __uint_as_float(braw.w));
    const float x0 = act_t<ACT>(__uint_as_float(v[4 * j + 0]) + b.x);
    const float x1 = act_t<ACT>(__uint_as_float(v[4 * j + 1]) + b.y);
    const float x2 = act_t<ACT>(__uint_as_float(v[4 * j + 2]) + b.z);
    const float x3 = act_t<ACT>(__uint_as_float(v[4 * j + 3]) + b.w);
    packed[2 * j] = pack_bf16x2(x0, x1);
    packed[2 * j + 1] = pack_bf16x2(x2, x3);
  }
}
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld32(taddr, v); }
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld16(taddr, v); }

// Runs on warps 2.. (threads 64..): warp w reads TMEM lanes 32*(w%4).. and the kEpiCols-column slice (w-2)/4 of each chunk.  tfull/tempty: the two-deep TMEM accumulator hand-shake with the MMA warp.
// scratch: kColPartBytes of shared memory.
template <int BLOCK_N>
__device__ __forceinline__ void conv_epilogue(const ConvEpiParams& p, uint8_t* staging, float* scratch, uint64_t* tfull_bar,
                                              uint64_t* tempty_bar, uint32_t tmem_base) {
  constexpr int kChunks = BLOCK_N / 64;
  float* colpart = scratch;                    // [kEpiWarps][64][2]
  float* sbias = scratch + kEpiWarps * 64 * 2; // [BLOCK_N]
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int et = threadIdx.x - 64;           // 0..kEpiThreads-1
  const int ew = et >> 5;                    // epilogue warp
  const int half = ew >> 2;                  // which kEpiCols columns of each 64-column chunk this warp converts
  const int row = (warp & 3) * 32 + lane;    // TMEM lane == tile row owned by this thread
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  // kernel-invariant scalars (keep them in registers instead of re-reading the constant bank)
  const int bw = p.bw, bh = p.bh, bn = p.bn, lbw = p.lbw, lbwh = p.lbwh;
  const int gN = p.gN, gH = p.gH, gW = p.gW, n_extent = p.n_extent, act = p.act;
  const int n_tiles = p.n_tiles, m_tiles = p.m_tiles, tiles_w = p.tiles_w, tiles_h = p.tiles_h;
  const bool has_stats = p.stats != nullptr, per_cta = p.stats_per_cta != 0;
  const int mul_mode = p.mul_mode, mul_c = p.mul_c;
  const long long osn = p.out_sn, osh = p.out_sh, osw = p.out_sw;
  const long long msn = p.mul_sn, msh = p.mul_sh, msw = p.mul_sw;
  const float* bias = p.bias;
  const uint32_t stg_addr = smem_u32(staging);
  const int vq = et & 7, r0 = et >> 3;       // this thread's 16-byte column slot and first row (of kEpiRows, stride kEpiRowStep) of the store pass
  // tile-invariant decomposition of this thread's rows
  const int mdn = row >> lbwh, mdh = (row >> lbw) & (bh - 1), mdw = row & (bw - 1);
  int rel[kEpiRows], mrel[kEpiRows];         // element offsets inside a tile: < 2^31 (checked by fill_epi_params)
#pragma unroll
  for (int i = 0; i < kEpiRows; ++i) {
    const int r = r0 + kEpiRowStep * i;
    const int dn = r >> lbwh, dh = (r >> lbw) & (bh - 1), dw = r & (bw - 1);
    rel[i] = (int)(dn * osn + dh * osh + dw * osw);
    mrel[i] = (int)(dn * msn + dh * msh + dw * msw);
  }
  // dgrad + derivative mask + statistics = "column sums of the masked channels" (bias gradient of the transposed conv whose
  // LeakyReLU' was fused).  When they fit one 64-column chunk and the CTA keeps its columns (n_tiles == 1) each thread just
  // keeps 8 running sums in registers for the whole kernel and the CTA reduces them once at the end: the per-chunk
  // shuffle + barrier reduction of the BatchNorm path cost +0.12 ms on the epilogue-bound 256x256 layer.
  const bool colsum_only = has_stats && mul_mode != 0 && n_tiles == 1 && mul_c <= 64 && per_cta;
  // Same idea for the BatchNorm statistics of a one-chunk tile (BLOCK_N = 64, the epilogue-bound high-resolution layers):
  // sum and sum of squares stay in 16 registers across all tiles of the CTA.
  const bool reg_stats = kChunks == 1 && has_stats && per_cta && !colsum_only;
  const bool persist = colsum_only || reg_stats;
  float s[8], q2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; q2[e] = 0.f; }
  float cta_s[kChunks], cta_q[kChunks];      // per-CTA BN statistics (threads et < 64)
#pragma unroll
  for (int c = 0; c < kChunks; ++c) { cta_s[c] = 0.f; cta_q[c] = 0.f; }
  if (n_tiles == 1) {                         // one column range for the whole kernel: stage the bias once
    for (int i = et; i < BLOCK_N; i += kEpiThreads) sbias[i] = (bias != nullptr && i < n_extent) ? __ldg(bias + i) : 0.f;
  }
  // Per-tile scalars (tile coordinates, 64-bit output / mask offsets, "tile lies fully inside the image").  ncu's source view
  // showed every epilogue warp spending ~100 of its ~350 instructions per tile on recomputing them (profiles/
  // r1_epilogue_stalls.txt), which is what bounds the thin high-resolution layers.  One thread now computes them for the
  // NEXT tile while the others wait for the accumulator, and publishes them in shared memory (two slots; the named
  // barriers of the chunk loop order the write of slot i+1 against its readers).
  const uint32_t tinfo = smem_u32(scratch + kEpiWarps * 64 * 2 + 256 + 4) & ~15u;
  auto publish_tile = [&](int tile, int slot) {
    const TileCoord tc = tile_coord(p, tile);
    const bool full = (tc.n0 + bn <= gN) && (tc.h0 + bh <= gH) && (tc.w0 + bw <= gW);
    const unsigned long long ob = p.out_ptr[tc.g] + 2ull * (unsigned long long)(tc.n0 * osn + tc.h0 * osh + tc.w0 * osw);
    const unsigned long long mo = (unsigned long long)(tc.n0 * msn + tc.h0 * msh + tc.w0 * msw);
    const uint32_t a = tinfo + slot * kTileInfoBytes;
    sts128(a, make_uint4((uint32_t)ob, (uint32_t)(ob >> 32), (uint32_t)mo, (uint32_t)(mo >> 32)));
    sts128(a + 16, make_uint4((uint32_t)tc.n_tile, (uint32_t)(tc.g * m_tiles + tc.m_tile), full ? 1u : 0u, 0u));
    sts128(a + 32, make_uint4((uint32_t)tc.n0, (uint32_t)tc.h0, (uint32_t)tc.w0, 0u));
  };
  uint32_t acc = 0, acc_phase = 0;
  const TileRange tr = tile_range(p);
  const bool publisher = et == kEpiThreads - 1;
  if (publisher && tr.first < tr.end) publish_tile(tr.first, 0);
  named_bar_sync(1, kEpiThreads);             // bias tile and the first tile's scalars are visible
  uint32_t slot = 0;
  int trace_it = 0;
  for (int tile = tr.first; tile < tr.end; tile += tr.step, slot ^= 1) {
    if (publisher && tile + tr.step < tr.end) publish_tile(tile + tr.step, slot ^ 1);
    const uint4 ti0 = lds128(tinfo + slot * kTileInfoBytes), ti1 = lds128(tinfo + slot * kTileInfoBytes + 16);
    __nv_bfloat16* const out_base = reinterpret_cast<__nv_bfloat16*>(((unsigned long long)ti0.y << 32) | ti0.x);
    const long long mtile_off = (long long)(((unsigned long long)ti0.w << 32) | ti0.z);
    const int n_tile = (int)ti1.x, stat_row = (int)ti1.y;
    const bool tile_full = ti1.z != 0;
    int n0 = 0, h0 = 0, w0 = 0;
    if (!tile_full) {
      const uint4 ti2 = lds128(tinfo + slot * kTileInfoBytes + 32);
      n0 = (int)ti2.x; h0 = (int)ti2.y; w0 = (int)ti2.z;
    }
    const bool my_valid = tile_full || ((n0 + mdn) < gN && (h0 + mdh) < gH && (w0 + mdw) < gW);
    if (n_tiles > 1) {
      named_bar_sync(1, kEpiThreads);         // previous tile's readers of sbias are done
      for (int i = et; i < BLOCK_N; i += kEpiThreads) {
        const int cc = n_tile * BLOCK_N + i;
        sbias[i] = (bias != nullptr && cc < n_extent) ? __ldg(bias + cc) : 0.f;
      }
      named_bar_sync(1, kEpiThreads);
    }

    // dgrad fusion: the forward activations whose sign masks a chunk are fetched from global memory.  Chunk 0's loads are
    // issued BEFORE waiting for the accumulator (their HBM latency hides behind the MMA main loop), later chunks' at the
    // top of the chunk (overlapping the TMEM read).
    uint4 yv[kEpiRows];
    auto load_mask = [&](int col) {
      const __nv_bfloat16* mb = reinterpret_cast<const __nv_bfloat16*>(p.mul_ptr) + mtile_off + col;
#pragma unroll
      for (int i = 0; i < kEpiRows; ++i) {
        const int r = r0 + kEpiRowStep * i;
        const bool ok = tile_full || ((n0 + (r >> lbwh)) < gN && (h0 + ((r >> lbw) & (bh - 1))) < gH && (w0 + (r & (bw - 1))) < gW);
        yv[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);  // bf16 1.0 -> derivative 1
        if (ok) yv[i] = __ldg(reinterpret_cast<const uint4*>(mb + mrel[i]));
      }
    };
    if (mul_mode != 0 && n_tile * BLOCK_N + vq * 8 < mul_c) load_mask(n_tile * BLOCK_N + vq * 8);

    mbar_wait(&tfull_bar[acc], acc_phase);
    tc_fence_after();
    const bool tracing = p.trace != nullptr && blockIdx.x == 0 && et == 0 && trace_it < 48;
    if (tracing) p.trace[trace_it * 8 + 4] = (unsigned long long)clock64();
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int col0 = n_tile * BLOCK_N + c * 64;
      const int cc_st = col0 + vq * 8;
      const bool col_ok = cc_st < n_extent;
      const bool do_mul = mul_mode != 0 && cc_st < mul_c;
      if (c > 0 && do_mul) load_mask(cc_st);
      {
        uint32_t v[kEpiCols];
        tmem_ld_cols(tmem_base + lane_base + acc * BLOCK_N + c * 64 + half * kEpiCols, v);
        tmem_ld_wait();
        if (tracing && c == 0) p.trace[trace_it * 8 + 5] = (unsigned long long)clock64();
        uint32_t packed[kEpiCols / 2];
        const uint32_t sb = smem_u32(sbias + c * 64 + half * kEpiCols);
        if (act == B2SEG_ACT_NONE) bias_act_pack<B2SEG_ACT_NONE, kEpiCols>(v, sb, packed);
        else if (act == B2SEG_ACT_RELU) bias_act_pack<B2SEG_ACT_RELU, kEpiCols>(v, sb, packed);
        else if (act == B2SEG_ACT_LEAKY) bias_act_pack<B2SEG_ACT_LEAKY, kEpiCols>(v, sb, packed);
        else bias_act_pack<B2SEG_ACT_SIGMOID, kEpiCols>(v, sb, packed);
        if (!my_valid) {
#pragma unroll
          for (int j = 0; j < kEpiCols / 2; ++j) packed[j] = 0u;
        }
        const uint32_t dst = stg_addr + row * kStgPitch + half * (kEpiCols * 2);
#pragma unroll
        for (int q = 0; q < kEpiCols / 8; ++q) sts128(dst + 16 * q, make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]));
      }
      if (c == kChunks - 1) {
        // accumulator fully read: hand the TMEM buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
      named_bar_sync(1, kEpiThreads);
      // ---- coalesced store of the 128 x 64 chunk (+ derivative mask) and column statistics from the same registers
      if (!persist) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { s[e] = 0.f; q2[e] = 0.f; }
      }
#pragma unroll
      for (int i = 0; i < kEpiRows; ++i) {
        const int r = r0 + kEpiRowStep * i;
        uint4 val = lds128(stg_addr + r * kStgPitch + vq * 16);
        const bool ok = tile_full || ((n0 + (r >> lbwh)) < gN && (h0 + ((r >> lbw) & (bh - 1))) < gH && (w0 + (r & (bw - 1))) < gW);
        if (ok && col_ok) {
          if (do_mul) {
            const uint32_t y4[4] = {yv[i].x, yv[i].y, yv[i].z, yv[i].w};
            uint32_t o4[4] = {val.x, val.y, val.z, val.w};
            const float neg = mul_mode == B2SEG_ACT_LEAKY ? 0.3f : 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // bf16 sign tests on the raw bits: positive and non-zero <=> (bits & 0x7fff) != 0 and sign bit clear
              const uint32_t ylo = y4[e] & 0xffffu, yhi = y4[e] >> 16;
              float lo = __uint_as_float(o4[e] << 16), hi = __uint_as_float(o4[e] & 0xffff0000u);
              if (!((ylo & 0x8000u) == 0 && (ylo & 0x7fffu) != 0)) lo *= neg;
              if (!((yhi & 0x8000u) == 0 && (yhi & 0x7fffu) != 0)) hi *= neg;
              o4[e] = pack_bf16x2(lo, hi);
            }
            val = make_uint4(o4[0], o4[1], o4[2], o4[3]);
          }
          *reinterpret_cast<uint4*>(out_base + rel[i] + cc_st) = val;
          if (colsum_only) {
            if (c == 0) {
              const uint32_t w4[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                s[2 * e] += __uint_as_float(w4[e] << 16);
                s[2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
              }
            }
          } else if (has_stats) {   // statistics of exactly what was stored (rows outside the image hold zeros and add nothing)
            const uint32_t w4[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float lo = __uint_as_float(w4[e] << 16), hi = __uint_as_float(w4[e] & 0xffff0000u);
              s[2 * e] += lo; q2[2 * e] = fmaf(lo, lo, q2[2 * e]);
              s[2 * e + 1] += hi; q2[2 * e + 1] = fmaf(hi, hi, q2[2 * e + 1]);
            }
          }
        }
      }
      if (has_stats && !persist) {
        // lanes l, l^8, l^16, l^24 own the same 8 columns
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          s[e] += __shfl_xor_sync(0xffffffffu, s[e], 8);
          q2[e] += __shfl_xor_sync(0xffffffffu, q2[e], 8);
          s[e] += __shfl_xor_sync(0xffffffffu, s[e], 16);
          q2[e] += __shfl_xor_sync(0xffffffffu, q2[e], 16);
        }
        if (lane < 8) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            colpart[(ew * 64 + lane * 8 + e) * 2 + 0] = s[e];
            colpart[(ew * 64 + lane * 8 + e) * 2 + 1] = q2[e];
          }
        }
        named_bar_sync(2, kEpiThreads);
        if (et < 64) {
          float s2 = 0.f, ss2 = 0.f;
#pragma unroll
          for (int w = 0; w < kEpiWarps; ++w) { s2 += colpart[(w * 64 + et) * 2]; ss2 += colpart[(w * 64 + et) * 2 + 1]; }
          if (per_cta) {
            cta_s[c] += s2;
            cta_q[c] += ss2;
          } else if (col0 + et < n_extent) {
            if (p.stats_atomic) {
              atomicAdd(p.stats + col0 + et, s2);
              atomicAdd(p.stats + n_extent + col0 + et, ss2);
            } else {
              float* st = p.stats + (size_t)stat_row * 2 * n_extent;
              st[col0 + et] = s2;
              st[n_extent + col0 + et] = ss2;
            }
          }
        }
      }
      named_bar_sync(1, kEpiThreads);  // staging / colpart are reused by the next chunk
    }
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1;
    if (tracing) p.trace[trace_it * 8 + 6] = (unsigned long long)clock64();
    ++trace_it;
  }
  if (persist) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s[e] += __shfl_xor_sync(0xffffffffu, s[e], 8);
      q2[e] += __shfl_xor_sync(0xffffffffu, q2[e], 8);
      s[e] += __shfl_xor_sync(0xffffffffu, s[e], 16);
      q2[e] += __shfl_xor_sync(0xffffffffu, q2[e], 16);
    }
    if (lane < 8) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        colpart[(ew * 64 + lane * 8 + e) * 2 + 0] = s[e];
        colpart[(ew * 64 + lane * 8 + e) * 2 + 1] = q2[e];
      }
    }
    named_bar_sync(2, kEpiThreads);
    if (et < 64) {
      float s2 = 0.f, ss2 = 0.f;
#pragma unroll
      for (int w = 0; w < kEpiWarps; ++w) { s2 += colpart[(w * 64 + et) * 2]; ss2 += colpart[(w * 64 + et) * 2 + 1]; }
      cta_s[0] = s2;    // colsum_only: sums of squares are 0 and unused (b2seg_rowsum reads the column sums)
      cta_q[0] = ss2;
    }
  }
  if (has_stats && per_cta && et < 64) {
    float* st = p.stats + (p.stats_atomic ? (size_t)0 : (size_t)blockIdx.x * 2 * n_extent);
#pragma unroll
    for (int c = 0; c < kChunks; ++c)
      if (c * 64 + et < n_extent) {
        if (p.stats_atomic) {
          atomicAdd(st + c * 64 + et, cta_s[c]);
          atomicAdd(st + n_extent + c * 64 + et, cta_q[c]);
        } else {
          st[c * 64 + et] = cta_s[c];
          st[n_extent + c * 64 + et] = cta_q[c];
        }
      }
  }
}

// host-side helpers shared by conv_gemm.cu / conv_halo.cu
int select_block_n(const b2seg_conv_desc* d);
int fill_epi_params(const b2seg_conv_desc* d, int block_n, int bw, int bh, int bn, ConvEpiParams* e);
int halo_geometry(const b2seg_conv_desc* d, int* bw, int* bh, int* bn);  // 0 when the halo kernel applies
PreparedOp* prepare_conv_halo(const b2seg_conv_desc* d);

}  // namespace b2
