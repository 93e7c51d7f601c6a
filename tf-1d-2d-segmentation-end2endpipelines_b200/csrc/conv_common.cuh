// Shared pieces of the implicit-GEMM convolution kernels (conv_gemm.cu: one TMA box per tap; conv_halo.cu: one halo
// tile per channel block reused by every tap): tile geometry, epilogue parameters and the epilogue itself
// (TMEM -> registers -> bias/activation -> bf16 staging in smem -> coalesced, strided store + BN statistics).
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace b2 {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kStgPitch = 144;                        // staging row pitch (64 bf16 + 16 B pad: conflict-free 16 B stores)
constexpr int kStgBytes = kBlockM * kStgPitch;
constexpr int kConvThreads = 192;
constexpr int kColPartBytes = 2 * 64 * 2 * 4;

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
};

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case B2SEG_ACT_RELU: return fmaxf(x, 0.f);
    case B2SEG_ACT_LEAKY: return x > 0.f ? x : 0.3f * x;
    case B2SEG_ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
    default: return x;
  }
}

// Runs on warps 2..5 (threads 64..191).  tfull/tempty: the two-deep TMEM accumulator hand-shake with the MMA warp.
template <int BLOCK_N>
__device__ __forceinline__ void conv_epilogue(const ConvEpiParams& p, uint8_t* staging, float* colpart, uint64_t* tfull_bar,
                                              uint64_t* tempty_bar, uint32_t tmem_base) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

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

// host-side helpers shared by conv_gemm.cu / conv_halo.cu
int select_block_n(const b2seg_conv_desc* d);
int fill_epi_params(const b2seg_conv_desc* d, int block_n, int bw, int bh, int bn, ConvEpiParams* e);
int halo_geometry(const b2seg_conv_desc* d, int* bw, int* bh, int* bn);  // 0 when the halo kernel applies
PreparedOp* prepare_conv_halo(const b2seg_conv_desc* d);

}  // namespace b2
