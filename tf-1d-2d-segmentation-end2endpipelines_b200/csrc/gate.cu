// Fused additive attention gate for sm_100a (include/b2seg.h: b2seg_gate_fwd / b2seg_gate_bwd) — the streaming half of
// 2DCNN/models/unet_variants.py:67-82 Attention_Block.  The two 1x1 projections run on the tcgen05 convolution kernels and add
// their BatchNorm column sums into [2][C] accumulators (stats_atomic); everything between them and the concat slot is here:
//
//   forward   gate_mid_fwd :  scale/shift of both BatchNorms from the sums, c = ReLU(BN(za) + BN(zb)), z = c . w3 + b3 per pixel,
//                             sum z / sum z^2 (the third BatchNorm's statistics)                     reads za, zb once; writes N*h*w floats
//             gate_out_fwd :  m = sigmoid(BN(z)); r = bilinear_x2(m) + LeakyReLU(ConvT4x4s2(m));  out = skip * r   reads skip, writes out
//   backward  gate_out_bwd :  dskip = dout * r,  dr = sum_c dout * skip                              reads dout, skip; writes dskip
//             gate_low_bwd :  dm = (bilinear^T + ConvT^T LeakyReLU')(dr), g3 = dm m (1 - m), its two BatchNorm sums,
//                             the 16 + 1 transposed-conv gradients                                    low-resolution maps only
//             gate_mid_bwd0:  dz3 = BN3'(g3); g = dz3 w3 ReLU'(c): per-channel sums of g, g zhat_a, g zhat_b, dz3 c  reads za, zb
//             gate_mid_bwd1:  dza, dzb (BatchNorm backward of both branches)                          reads za, zb; writes dza, dzb
//
// None of BN(za), BN(zb), their sum, the one-channel maps at either resolution or the resampler is ever written as a tensor: the
// unfused lowering moved ~14 tensors per gate through HBM in 14 forward and ~25 backward launches.  Grid-wide BatchNorm
// reductions are the kernel boundaries (three in each direction); no further fusion is possible in training mode.
#include <math.h>

#include "stream_common.cuh"

namespace b2 {

struct GateK {
  DView za, zb, skip, out, dout, dskip, dza, dzb, da_low;
  const float *sums_a, *sums_b, *gamma_a, *beta_a, *gamma_b, *beta_b;
  float *mm_a, *mv_a, *mm_b, *mv_b, *vec_a, *vec_b;
  const float *w3, *b3;
  float* z;
  float* m;               // sigmoid(BN3(z)), written by the forward's output kernel (one low-resolution pixel per even/even thread)
  float* sums3;
  const float *gamma3, *beta3;
  float *mm3, *mv3;
  const float *wt, *bt;
  int wt_stride;
  float inv_count, count, eps, momentum;
  int training, bessel;
  int C, h, w, npix;      // low-resolution grid, npix = N*h*w
  float *dr, *g3, *bsums3, *bsums_ab;
  float *dgamma_a, *dbeta_a, *dgamma_b, *dbeta_b, *dw3, *db3, *dgamma3, *dbeta3, *dwt, *dbt;
  FastDiv fd_w, fd_h, fd_w2, fd_h2;
};

// BatchNorm scale / shift / mean / rstd of channel c from column sums (training) or moving statistics (inference)
__device__ __forceinline__ void bn_coeffs(const GateK& k, const float* sums, const float* gamma, const float* beta, const float* mm, const float* mv,
                                          int c, float* sc, float* sh, float* mu, float* rs) {
  float mean, var;
  if (k.training) {
    mean = sums[c] * k.inv_count;
    var = fmaxf(sums[k.C + c] * k.inv_count - mean * mean, 0.f);
  } else {
    mean = mm[c]; var = mv[c];
  }
  const float r = rsqrtf(var + k.eps);
  *sc = gamma[c] * r; *sh = beta[c] - mean * gamma[c] * r; *mu = mean; *rs = r;
}
__device__ __forceinline__ void update_moving(const GateK& k, float* mm, float* mv, int c, float mean, float var) {
  const float uv = (k.bessel && k.count > 1.f) ? var * k.count / (k.count - 1.f) : var;
  mm[c] = mm[c] * k.momentum + mean * (1.f - k.momentum);
  mv[c] = mv[c] * k.momentum + uv * (1.f - k.momentum);
}

// ------------------------------------------------------------------------------------------ forward, low resolution
// One warp handles 32 / LPP pixels at a time: LPP = min(C / 8, 32) lanes per pixel, each lane VPL = C / 8 / LPP 8-channel vectors.
// Shared memory: [4][C] floats (scale_a, scale_b, shift_a + shift_b, w3).  All 2 * GU * VPL 16-byte loads of a round are issued
// before the first use (clamped addresses instead of branches), so a warp has 4 - 8 KB in flight.
// c = a * scale_a + (b * scale_b + (shift_a + shift_b)) — the same expression in the two backward passes (identical ReLU mask).
template <int VPL>      // 1, 2, 4; 0 = any (run-time trip count)
__global__ void __launch_bounds__(256) gate_mid_fwd_kernel(const GateK k) {
  pdl_prologue();
  extern __shared__ float sm[];
  const int C = k.C;
  float *s_sa = sm, *s_sb = sm + C, *s_t = sm + 2 * C, *s_w3 = sm + 3 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float sca, sha, mua, rsa, scb, shb, mub, rsb;
    bn_coeffs(k, k.sums_a, k.gamma_a, k.beta_a, k.mm_a, k.mv_a, c, &sca, &sha, &mua, &rsa);
    bn_coeffs(k, k.sums_b, k.gamma_b, k.beta_b, k.mm_b, k.mv_b, c, &scb, &shb, &mub, &rsb);
    s_sa[c] = sca; s_sb[c] = scb; s_t[c] = sha + shb;
    if (blockIdx.x == 0) {
      k.vec_a[c] = sca; k.vec_a[C + c] = sha; k.vec_a[2 * C + c] = mua; k.vec_a[3 * C + c] = rsa;
      k.vec_b[c] = scb; k.vec_b[C + c] = shb; k.vec_b[2 * C + c] = mub; k.vec_b[3 * C + c] = rsb;
      if (k.training) {
        update_moving(k, k.mm_a, k.mv_a, c, mua, fmaxf(k.sums_a[C + c] * k.inv_count - mua * mua, 0.f));
        update_moving(k, k.mm_b, k.mv_b, c, mub, fmaxf(k.sums_b[C + c] * k.inv_count - mub * mub, 0.f));
      }
    }
    s_w3[c] = k.w3[c];
  }
  __syncthreads();
  const int vpp = C >> 3;
  const int lpp = vpp < 32 ? vpp : 32;
  const int vpl = VPL ? VPL : vpp / lpp;
  const int ppw = 32 / lpp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / lpp, lv = lane % lpp;
  const float b3 = k.b3[0];
  float acc_s = 0.f, acc_q = 0.f;
  const int groups = (k.npix + ppw - 1) / ppw;
  constexpr int GU = VPL == 4 ? 1 : VPL == 2 ? 2 : 4;     // pixel groups in flight per warp (8 x 2 16-byte loads per lane)
  constexpr int NV = VPL ? VPL : 1;
  for (int g0 = (blockIdx.x * 8 + warp) * GU; g0 < groups; g0 += gridDim.x * 8 * GU) {
    float dot[GU];
    int pixs[GU];
    const __nv_bfloat16 *pa[GU], *pb[GU];
    uint4 ra[GU][NV], rb[GU][NV];
#pragma unroll
    for (int u = 0; u < GU; ++u) {
      const int pix = (g0 + u) * ppw + sub;
      pixs[u] = pix;
      const int pc = pix < k.npix ? pix : k.npix - 1;
      const uint32_t q = fast_div((uint32_t)pc, k.fd_w);
      const int x = pc - (int)q * k.w;
      const uint32_t n = fast_div(q, k.fd_h);
      const int y = (int)q - (int)n * k.h;
      pa[u] = vaddr(k.za, (int)n, y, x, lv * 8);
      pb[u] = vaddr(k.zb, (int)n, y, x, lv * 8);
      if (VPL) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          ra[u][i] = __ldg(reinterpret_cast<const uint4*>(pa[u] + i * lpp * 8));
          rb[u][i] = __ldg(reinterpret_cast<const uint4*>(pb[u] + i * lpp * 8));
        }
      }
    }
#pragma unroll
    for (int u = 0; u < GU; ++u) {
      float d = 0.f;
      for (int i = 0; i < vpl; ++i) {
        const int c0 = (lv + i * lpp) * 8;
        uint4 va, vb;
        if (VPL) { va = ra[u][VPL ? i : 0]; vb = rb[u][VPL ? i : 0]; }
        else {
          va = __ldg(reinterpret_cast<const uint4*>(pa[u] + i * lpp * 8));
          vb = __ldg(reinterpret_cast<const uint4*>(pb[u] + i * lpp * 8));
        }
        const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&va);
        const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&vb);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 fa = __bfloat1622float2(ha[e]), fb = __bfloat1622float2(hb[e]);
          const int c = c0 + 2 * e;
          const float c_lo = fmaxf(fmaf(fa.x, s_sa[c], fmaf(fb.x, s_sb[c], s_t[c])), 0.f);
          const float c_hi = fmaxf(fmaf(fa.y, s_sa[c + 1], fmaf(fb.y, s_sb[c + 1], s_t[c + 1])), 0.f);
          d = fmaf(c_lo, s_w3[c], d);
          d = fmaf(c_hi, s_w3[c + 1], d);
        }
      }
      dot[u] = d;
    }
#pragma unroll
    for (int u = 0; u < GU; ++u) {
      float dsum = dot[u];
      for (int off = lpp >> 1; off > 0; off >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, off);
      if (lv == 0 && pixs[u] < k.npix) {
        const float z = dsum + b3;
        k.z[pixs[u]] = z;
        acc_s += z; acc_q = fmaf(z, z, acc_q);
      }
    }
  }
  if (k.training) {
    __shared__ float red[2][8];
    for (int off = 16; off > 0; off >>= 1) {
      acc_s += __shfl_xor_sync(0xffffffffu, acc_s, off);
      acc_q += __shfl_xor_sync(0xffffffffu, acc_q, off);
    }
    if (lane == 0) { red[0][warp] = acc_s; red[1][warp] = acc_q; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f, q = 0.f;
      for (int i = 0; i < 8; ++i) { s += red[0][i]; q += red[1][i]; }
      atomicAdd(k.sums3, s);
      atomicAdd(k.sums3 + 1, q);
    }
  }
}

// third BatchNorm: (scale, shift, mean, rstd) from the two sums (training) or the moving statistics
__device__ __forceinline__ void bn3_coeffs(const GateK& k, float* sc, float* sh, float* mu, float* rs) {
  float mean, var;
  if (k.training) {
    mean = k.sums3[0] * k.inv_count;
    var = fmaxf(k.sums3[1] * k.inv_count - mean * mean, 0.f);
  } else {
    mean = k.mm3[0]; var = k.mv3[0];
  }
  const float r = rsqrtf(var + k.eps);
  *sc = k.gamma3[0] * r; *sh = k.beta3[0] - mean * k.gamma3[0] * r; *mu = mean; *rs = r;
}

struct Resampler {
  float r;      // bilinear + LeakyReLU(transposed conv)
  float lk;     // LeakyReLU'(pre-activation of the transposed conv): 1 or 0.3
};
// m(y, x) = sigmoid(z * sc + sh) of image n with edge clamp (bilinear) or zero outside (transposed conv) handled by the caller
template <bool FROM_MAP>
__device__ __forceinline__ float gate_m(const GateK& k, int n, int y, int x, float sc, float sh) {
  if (FROM_MAP) return __ldg(k.m + (n * k.h + y) * k.w + x);
  return 1.f / (1.f + __expf(-fmaf(k.z[(n * k.h + y) * k.w + x], sc, sh)));
}
// resampler value at high-resolution pixel (oy, ox) of image n.
//   bilinear x2, half-pixel centres, edge clamp (tf.image.resize): even o = 2i: 0.25 m[i-1] + 0.75 m[i]; odd o = 2i+1: 0.75 m[i] + 0.25 m[i+1]
//   Conv2DTranspose(4x4, s2, 'same') == ConvTranspose2d(k4, s2, p1): o = 2i - 1 + ky: even o: (ky=1, i), (ky=3, i-1); odd o: (ky=2, i), (ky=0, i+1)
template <bool FROM_MAP>
__device__ __forceinline__ Resampler gate_resample(const GateK& k, int n, int oy, int ox, float sc, float sh, const float* wt, float bt) {
  const int iy = oy >> 1, ix = ox >> 1;
  const int py = oy & 1, px = ox & 1;
  const int y2 = py ? iy + 1 : iy - 1, x2 = px ? ix + 1 : ix - 1;      // the second source row / column
  const bool vy2 = y2 >= 0 && y2 < k.h, vx2 = x2 >= 0 && x2 < k.w;
  const int yc = vy2 ? y2 : iy, xc = vx2 ? x2 : ix;                     // clamped (bilinear)
  const float m00 = gate_m<FROM_MAP>(k, n, iy, ix, sc, sh);
  const float m01 = gate_m<FROM_MAP>(k, n, iy, xc, sc, sh);
  const float m10 = gate_m<FROM_MAP>(k, n, yc, ix, sc, sh);
  const float m11 = gate_m<FROM_MAP>(k, n, yc, xc, sc, sh);
  const float bil = 0.75f * (0.75f * m00 + 0.25f * m01) + 0.25f * (0.75f * m10 + 0.25f * m11);
  const int ky1 = py ? 2 : 1, ky2 = py ? 0 : 3, kx1 = px ? 2 : 1, kx2 = px ? 0 : 3;
  float pre = bt + m00 * wt[ky1 * 4 + kx1];
  if (vx2) pre = fmaf(m01, wt[ky1 * 4 + kx2], pre);
  if (vy2) pre = fmaf(m10, wt[ky2 * 4 + kx1], pre);
  if (vy2 && vx2) pre = fmaf(m11, wt[ky2 * 4 + kx2], pre);
  Resampler o;
  o.lk = pre > 0.f ? 1.f : 0.3f;
  o.r = bil + pre * o.lk;
  return o;
}

// ------------------------------------------------------------------------------------------ forward / backward, high resolution
// Block = 256 threads; one block iteration = a chunk of P consecutive high-resolution pixels in (n, y, x) order (P = 32 .. 1024, a
// power of two picked by the host so that small maps still fill the SMs).  r and the pixel coordinates are computed once per pixel
// into shared memory, then the chunk's (pixel, 8-channel vector) items are streamed, kGateU 16-byte loads in flight per thread and no
// barrier until the chunk is done.  BWD: dskip = dout * r and dr = sum_c dout * skip (per-pixel reduction: warp shuffle, then shared).
constexpr int kGateP = 1024;
constexpr int kGateU = 4;      // (pixel, vector) items in flight per thread
template <int MODE>   // 0 forward: out = skip * r;  1 backward: dr = sum_c dout * skip (+ dskip = dout * r when no second pass follows);
                      // 2 backward, second pass: dskip = dout * r + the stride-2 projection's input gradient at the even pixels
__global__ void __launch_bounds__(256) gate_out_kernel(const GateK k, int P, int vshift) {
  constexpr bool BWD = MODE == 1;
  pdl_prologue();
  __shared__ float s_r[kGateP];
  __shared__ float s_dr[BWD ? kGateP : 1];
  __shared__ uint32_t s_ny[kGateP];     // n << 16 | y
  __shared__ uint16_t s_x[kGateP];
  __shared__ float s_wt[16];
  float sc, sh, mu, rs;
  bn3_coeffs(k, &sc, &sh, &mu, &rs);
  if (threadIdx.x < 16) s_wt[threadIdx.x] = k.wt[threadIdx.x * k.wt_stride];
  const float bt = k.bt[0];
  const int W2 = 2 * k.w, H2 = 2 * k.h;
  const int total = k.skip.N * H2 * W2;
  const int vmask = (1 << vshift) - 1;
  const int span = vmask < 31 ? vmask + 1 : 32;
  if (MODE == 0 && blockIdx.x == 0 && threadIdx.x == 0 && k.training) {
    const float var = fmaxf(k.sums3[1] * k.inv_count - mu * mu, 0.f);
    const float uv = (k.bessel && k.count > 1.f) ? var * k.count / (k.count - 1.f) : var;
    k.mm3[0] = k.mm3[0] * k.momentum + mu * (1.f - k.momentum);
    k.mv3[0] = k.mv3[0] * k.momentum + uv * (1.f - k.momentum);
  }
  for (int p0 = blockIdx.x * P; p0 < total; p0 += gridDim.x * P) {
    const int npx = total - p0 < P ? total - p0 : P;
    __syncthreads();
    for (int j = threadIdx.x; j < npx; j += 256) {
      const uint32_t rowi = fast_div((uint32_t)(p0 + j), k.fd_w2);
      const int ox = p0 + j - (int)rowi * W2;
      const uint32_t n = fast_div(rowi, k.fd_h2);
      const int oy = (int)rowi - (int)n * H2;
      s_ny[j] = (n << 16) | (uint32_t)oy;
      s_x[j] = (uint16_t)ox;
      s_r[j] = gate_resample<MODE != 0>(k, (int)n, oy, ox, sc, sh, s_wt, bt).r;      // backward: m from the map the forward left
      if (BWD) s_dr[j] = 0.f;
      else if (MODE == 0 && !(oy & 1) && !(ox & 1)) k.m[((int)n * k.h + (oy >> 1)) * k.w + (ox >> 1)] = gate_m<false>(k, (int)n, oy >> 1, ox >> 1, sc, sh);
    }
    __syncthreads();
    const int work = npx << vshift;
    // (uniform trip count: the warp shuffles below need every lane, also in the ragged last round)
    for (int i0 = 0; i0 < work; i0 += 256 * kGateU) {
      uint4 sv[kGateU], dv4[kGateU];
      int pp[kGateU], cc[kGateU];
      bool live[kGateU];
#pragma unroll
      for (int u = 0; u < kGateU; ++u) {
        const int i = i0 + u * 256 + (int)threadIdx.x;
        live[u] = i < work;
        pp[u] = live[u] ? i >> vshift : 0;
        cc[u] = live[u] ? (i & vmask) * 8 : 0;
        sv[u] = make_uint4(0u, 0u, 0u, 0u);
        dv4[u] = sv[u];
        if (live[u]) {
          const uint32_t ny = s_ny[pp[u]];
          const int n = (int)(ny >> 16), oy = (int)(ny & 0xffffu), ox = (int)s_x[pp[u]];
          if (MODE == 2) {       // sv = the projection's input gradient (low resolution) at even pixels, else 0; dv4 = dout
            if (!(oy & 1) && !(ox & 1)) sv[u] = __ldg(reinterpret_cast<const uint4*>(vaddr(k.da_low, n, oy >> 1, ox >> 1, cc[u])));
            dv4[u] = __ldg(reinterpret_cast<const uint4*>(vaddr(k.dout, n, oy, ox, cc[u])));
          } else {
            sv[u] = __ldg(reinterpret_cast<const uint4*>(vaddr(k.skip, n, oy, ox, cc[u])));
            if (BWD) dv4[u] = __ldg(reinterpret_cast<const uint4*>(vaddr(k.dout, n, oy, ox, cc[u])));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kGateU; ++u) {
        if (i0 + u * 256 >= work) break;          // warp-uniform (whole 256-item rounds)
        float s[8], o[8];
        const __nv_bfloat162* hs = reinterpret_cast<const __nv_bfloat162*>(&sv[u]);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 t = __bfloat1622float2(hs[e]); s[2 * e] = t.x; s[2 * e + 1] = t.y; }
        const float r = s_r[pp[u]];
        const uint32_t ny = s_ny[pp[u]];
        const int n = (int)(ny >> 16), oy = (int)(ny & 0xffffu), ox = (int)s_x[pp[u]];
        if (MODE == 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = s[e] * r;
          if (live[u]) store8(vaddr(k.out, n, oy, ox, cc[u]), o);
        } else if (MODE == 2) {
          float d[8];
          const __nv_bfloat162* hd = reinterpret_cast<const __nv_bfloat162*>(&dv4[u]);
#pragma unroll
          for (int e = 0; e < 4; ++e) { const float2 t = __bfloat1622float2(hd[e]); d[2 * e] = t.x; d[2 * e + 1] = t.y; }
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = fmaf(d[e], r, s[e]);
          if (live[u]) store8(vaddr(k.dskip, n, oy, ox, cc[u]), o);
        } else {
          float d[8];
          const __nv_bfloat162* hd = reinterpret_cast<const __nv_bfloat162*>(&dv4[u]);
#pragma unroll
          for (int e = 0; e < 4; ++e) { const float2 t = __bfloat1622float2(hd[e]); d[2 * e] = t.x; d[2 * e + 1] = t.y; }
          float part = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) { o[e] = d[e] * r; part = fmaf(d[e], s[e], part); }
          if (live[u] && k.dskip.ptr) store8(vaddr(k.dskip, n, oy, ox, cc[u]), o);
          // lanes holding the same pixel are adjacent (the vectors per pixel are a power of two): reduce within the warp first
          for (int off = span >> 1; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
          if (live[u] && (threadIdx.x & (span - 1)) == 0) atomicAdd(&s_dr[pp[u]], part);
        }
      }
    }
    if (BWD) {
      __syncthreads();
      for (int j = threadIdx.x; j < npx; j += 256) k.dr[p0 + j] = s_dr[j];     // dr is dense in (n, y, x) order
    }
  }
}

// ------------------------------------------------------------------------------------------ backward, low-resolution maps
// One thread per low-resolution pixel: dm = adjoint of (bilinear + ConvT with LeakyReLU') applied to dr over the 4x4 high-resolution
// window this pixel feeds; g3 = dm * m * (1 - m); block sums of g3, g3 * zhat3 and of the 16 + 1 transposed-conv gradients.
__global__ void __launch_bounds__(256) gate_low_bwd_kernel(const GateK k) {
  pdl_prologue();
  float sc, sh, mu, rs;
  bn3_coeffs(k, &sc, &sh, &mu, &rs);
  __shared__ float wt[16];
  if (threadIdx.x < 16) wt[threadIdx.x] = k.wt[threadIdx.x * k.wt_stride];
  __syncthreads();
  const float bt = k.bt[0];
  const int W2 = 2 * k.w, H2 = 2 * k.h;
  float acc[19];      // [0..15] dwt, [16] dbt, [17] sum g3, [18] sum g3 * zhat3
#pragma unroll
  for (int i = 0; i < 19; ++i) acc[i] = 0.f;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < k.npix; pix += gridDim.x * blockDim.x) {
    const uint32_t q = fast_div((uint32_t)pix, k.fd_w);
    const int x = pix - (int)q * k.w;
    const uint32_t n = fast_div(q, k.fd_h);
    const int y = (int)q - (int)n * k.h;
    const float zv = k.z[pix];
    const float m = __ldg(k.m + pix);
    float dm = 0.f;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
      const int oy = 2 * y - 1 + ky;
      if (oy < 0 || oy >= H2) continue;
      // bilinear weight of m[y] in output row oy: rows 2y, 2y+1 take 0.75; rows 2y-1, 2y+2 take 0.25; the clamped edge rows
      // (oy = 0 from y = 0, oy = H2-1 from y = h-1) take the clamped source's share as well
      float wy = (ky == 1 || ky == 2) ? 0.75f : 0.25f;
      if ((oy == 0 && y == 0) || (oy == H2 - 1 && y == k.h - 1)) wy = 1.f;
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int ox = 2 * x - 1 + kx;
        if (ox < 0 || ox >= W2) continue;
        float wx = (kx == 1 || kx == 2) ? 0.75f : 0.25f;
        if ((ox == 0 && x == 0) || (ox == W2 - 1 && x == k.w - 1)) wx = 1.f;
        const float d = k.dr[((size_t)n * H2 + oy) * W2 + ox];
        const float lk = gate_resample<true>(k, (int)n, oy, ox, sc, sh, wt, bt).lk;
        dm = fmaf(d, fmaf(lk, wt[ky * 4 + kx], wy * wx), dm);
        acc[ky * 4 + kx] = fmaf(d * lk, m, acc[ky * 4 + kx]);
        if ((ky == 1 || ky == 2) && (kx == 1 || kx == 2)) acc[16] = fmaf(d, lk, acc[16]);   // each high-resolution pixel counted once (by its 2x2 owner)
      }
    }
    const float g = dm * m * (1.f - m);
    k.g3[pix] = g;
    acc[17] += g;
    acc[18] = fmaf(g, (zv - mu) * rs, acc[18]);
  }
  __shared__ float red[19][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 19; ++i) {
    float v = acc[i];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 19) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += red[threadIdx.x][i];
    if (threadIdx.x < 16) atomicAdd(k.dwt + threadIdx.x * k.wt_stride, v);
    else if (threadIdx.x == 16) atomicAdd(k.dbt, v);
    else atomicAdd(k.bsums3 + (threadIdx.x - 17), v);
  }
}

// ------------------------------------------------------------------------------------------ backward, projections
// Thread = one 8-channel vector (tcv) of a strided set of pixels (trow); blockIdx.x = vector group, blockIdx.y strides the pixels.
// PASS 0: per-channel sums {g, g zhat_a, g zhat_b, dz3 c}; PASS 1: dza, dzb.   With c = a sa + b sb + t and g = [c > 0] dz3 w3:
//   PASS 0 accumulates sum g, sum g a, sum g b, sum dz3 relu(c) (constants sa, sb, t, w3 only) and turns the raw products into
//          sum g zhat = rstd (sum g x - mean sum g) once per block, before the atomics;
//   PASS 1 dza = sa (g - mean(g) - zhat_a mean(g zhat_a)) = [c > 0] dz3 (sa w3) - A0 - A1 a with per-channel A0, A1 (same for b).
template <int PASS>
__global__ void __launch_bounds__(256, 2) gate_mid_bwd_kernel(const GateK k, int cvb) {
  pdl_prologue();
  const int C = k.C;
  const int tcv = threadIdx.x % cvb, trow = threadIdx.x / cvb, rows = 256 / cvb;
  const int c0 = (blockIdx.x * cvb + tcv) * 8;
  float sc3, sh3, mu3, rs3;
  bn3_coeffs(k, &sc3, &sh3, &mu3, &rs3);
  const float mg = k.bsums3[0] * k.inv_count, mgz = k.bsums3[1] * k.inv_count;     // mean g3, mean g3 zhat3
  float sa[8], sb[8], tt[8], w3[8];                    // PASS 1: w3 is not kept, saw = sa w3 and sbw = sb w3 are
  float saw[PASS ? 8 : 1], sbw[PASS ? 8 : 1], A0[PASS ? 8 : 1], A1[PASS ? 8 : 1], B0[PASS ? 8 : 1], B1[PASS ? 8 : 1];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c0 + e;
    sa[e] = k.vec_a[c]; sb[e] = k.vec_b[c]; tt[e] = k.vec_a[C + c] + k.vec_b[C + c];
    w3[e] = k.w3[c];
    if (PASS == 1) {
      const float ma = k.vec_a[2 * C + c], ra = k.vec_a[3 * C + c], mb = k.vec_b[2 * C + c], rb = k.vec_b[3 * C + c];
      const float cg = k.bsums_ab[c] * k.inv_count, cga = k.bsums_ab[C + c] * k.inv_count, cgb = k.bsums_ab[2 * C + c] * k.inv_count;
      saw[e] = sa[e] * w3[e]; sbw[e] = sb[e] * w3[e];
      A1[e] = sa[e] * ra * cga; A0[e] = sa[e] * cg - A1[e] * ma;
      B1[e] = sb[e] * rb * cgb; B0[e] = sb[e] * cg - B1[e] * mb;
    }
  }
  float acc[PASS ? 1 : 4][8];
#pragma unroll
  for (int a = 0; a < (PASS ? 1 : 4); ++a)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[a][e] = 0.f;
  float acc_db3 = 0.f;
  constexpr int PU = PASS ? 2 : 4;      // pixels in flight per thread (two blocks per SM: 128 registers)
  const int pstride = gridDim.y * rows;
  for (int pix0 = blockIdx.y * rows + trow; pix0 < k.npix; pix0 += pstride * PU) {
    uint4 rawa[PU], rawb[PU];
    float dz3s[PU];
    const __nv_bfloat16 *oa[PU], *ob[PU];
#pragma unroll
    for (int u = 0; u < PU; ++u) {
      const int pix = pix0 + u * pstride;
      const int pc = pix < k.npix ? pix : k.npix - 1;
      const uint32_t q = fast_div((uint32_t)pc, k.fd_w);
      const int x = pc - (int)q * k.w;
      const uint32_t n = fast_div(q, k.fd_h);
      const int y = (int)q - (int)n * k.h;
      const float zh3 = (k.z[pc] - mu3) * rs3;
      dz3s[u] = pix < k.npix ? sc3 * (k.g3[pc] - mg - zh3 * mgz) : 0.f;
      rawa[u] = __ldg(reinterpret_cast<const uint4*>(vaddr(k.za, (int)n, y, x, c0)));
      rawb[u] = __ldg(reinterpret_cast<const uint4*>(vaddr(k.zb, (int)n, y, x, c0)));
      if (PASS == 1) { oa[u] = vaddr(k.dza, (int)n, y, x, c0); ob[u] = vaddr(k.dzb, (int)n, y, x, c0); }
    }
#pragma unroll
    for (int u = 0; u < PU; ++u) {
      const int pix = pix0 + u * pstride;
      if (pix >= k.npix) break;
      const float dz3 = dz3s[u];
      float a[8], b[8];
      {
        const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&rawa[u]);
        const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&rawb[u]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 fa = __bfloat1622float2(ha[e]), fb = __bfloat1622float2(hb[e]);
          a[2 * e] = fa.x; a[2 * e + 1] = fa.y; b[2 * e] = fb.x; b[2 * e + 1] = fb.y;
        }
      }
      if (PASS == 0) {
        if (blockIdx.x == 0 && tcv == 0 && k.db3) acc_db3 += dz3;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float c = fmaf(a[e], sa[e], fmaf(b[e], sb[e], tt[e]));
          const float g = c > 0.f ? dz3 * w3[e] : 0.f;
          acc[0][e] += g;
          acc[1][e] = fmaf(g, a[e], acc[1][e]);
          acc[2][e] = fmaf(g, b[e], acc[2][e]);
          acc[3][e] = fmaf(dz3, fmaxf(c, 0.f), acc[3][e]);
        }
      } else {
        float da[8], db[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float c = fmaf(a[e], sa[e], fmaf(b[e], sb[e], tt[e]));
          const float G = c > 0.f ? dz3 : 0.f;
          da[e] = fmaf(G, saw[e], -fmaf(A1[e], a[e], A0[e]));
          db[e] = fmaf(G, sbw[e], -fmaf(B1[e], b[e], B0[e]));
        }
        store8(oa[u], da);
        store8(ob[u], db);
      }
    }
  }
  if (PASS == 0) {
    extern __shared__ float red[];      // [256][33]
    float* mine = red + threadIdx.x * 33;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int e = 0; e < 8; ++e) mine[a * 8 + e] = acc[a][e];
    mine[32] = acc_db3;
    __syncthreads();
    // thread t < cvb * 32 sums slot (t % 32) of vector (t / 32) over the rows
    for (int t = threadIdx.x; t < cvb * 32; t += 256) {
      const int v = t >> 5, slot = t & 31;
      const int which = slot >> 3;
      float s = 0.f, sg = 0.f;
      for (int r = 0; r < rows; ++r) {
        s += red[(r * cvb + v) * 33 + slot];
        if (which == 1 || which == 2) sg += red[(r * cvb + v) * 33 + (slot & 7)];     // sum g of the same channel
      }
      const int c = (blockIdx.x * cvb + v) * 8 + (slot & 7);
      if (which == 1) s = k.vec_a[3 * C + c] * (s - k.vec_a[2 * C + c] * sg);         // sum g zhat_a = rstd_a (sum g a - mean_a sum g)
      if (which == 2) s = k.vec_b[3 * C + c] * (s - k.vec_b[2 * C + c] * sg);
      if (which < 3) atomicAdd(k.bsums_ab + which * C + c, s);
      else atomicAdd(k.dw3 + c, s);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      float s = 0.f;
      for (int r = 0; r < rows; ++r) s += red[(r * cvb) * 33 + 32];
      if (k.db3) atomicAdd(k.db3, s);    // (null: the bias feeds a BatchNorm, its gradient is analytically zero and the caller keeps it exactly 0)
    }
  } else if (blockIdx.y == 0 && trow == 0) {
    // BatchNorm parameter gradients of the two branches and of the one-channel map (plain stores: each belongs to this gate only)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = c0 + e;
      k.dbeta_a[c] = k.bsums_ab[c]; k.dgamma_a[c] = k.bsums_ab[C + c];
      k.dbeta_b[c] = k.bsums_ab[c]; k.dgamma_b[c] = k.bsums_ab[2 * C + c];
    }
    if (blockIdx.x == 0 && tcv == 0) { k.dbeta3[0] = k.bsums3[0]; k.dgamma3[0] = k.bsums3[1]; }
  }
}

// ------------------------------------------------------------------------------------------ host side
static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static int fill_gate(const b2seg_gate_desc* d, GateK* k, bool bwd) {
  memset(k, 0, sizeof(*k));
  const b2seg_view& za = d->za;
  if (za.C % 8 || !pow2(za.C / 8) || za.C > 4096) { set_error("gate: projection channels %d (need 8 * 2^k <= 4096)", za.C); return -1; }
  if (d->zb.N != za.N || d->zb.H != za.H || d->zb.W != za.W || d->zb.C != za.C) { set_error("gate: projection shapes differ"); return -1; }
  if (d->skip.N != za.N || d->skip.H != 2 * za.H || d->skip.W != 2 * za.W || d->skip.C % 8 || !pow2(d->skip.C / 8)) {
    set_error("gate: skip (%d,%d,%d,%d) must be (N, 2h, 2w, 8 * 2^k) for projections (%d,%d,%d)", d->skip.N, d->skip.H, d->skip.W, d->skip.C, za.N, za.H, za.W);
    return -1;
  }
  if ((long long)za.N * za.H * za.W * 4 >= (1ll << 31) || za.N >= 65536 || 2 * za.H >= 65536 || 2 * za.W >= 65536) { set_error("gate: map too large"); return -1; }
  if (!d->w3 || !d->b3 || !d->z || !d->m || !d->wt || !d->bt || !d->gamma3 || !d->beta3 || !d->mm3 || !d->mv3 || !d->vec_a || !d->vec_b) { set_error("gate: null parameter"); return -1; }
  if (d->training && (!d->sums_a || !d->sums_b || !d->sums3)) { set_error("gate: training needs the statistics accumulators"); return -1; }
  k->za = dv(d->za); k->zb = dv(d->zb); k->skip = dv(d->skip); k->out = dv(d->out);
  k->dout = dv(d->dout); k->dskip = dv(d->dskip); k->dza = dv(d->dza); k->dzb = dv(d->dzb); k->da_low = dv(d->da_low);
#define P(f) k->f = reinterpret_cast<decltype(k->f)>(d->f)
  P(sums_a); P(sums_b); P(gamma_a); P(beta_a); P(gamma_b); P(beta_b); P(mm_a); P(mv_a); P(mm_b); P(mv_b); P(vec_a); P(vec_b);
  P(w3); P(b3); P(z); P(m); P(sums3); P(gamma3); P(beta3); P(mm3); P(mv3); P(wt); P(bt);
  P(dr); P(g3); P(bsums3); P(bsums_ab); P(dgamma_a); P(dbeta_a); P(dgamma_b); P(dbeta_b); P(dw3); P(db3); P(dgamma3); P(dbeta3); P(dwt); P(dbt);
#undef P
  k->wt_stride = d->wt_stride;
  k->count = (float)d->count; k->inv_count = (float)(1.0 / d->count); k->eps = d->eps; k->momentum = d->momentum;
  k->training = d->training; k->bessel = d->bessel;
  k->C = za.C; k->h = za.H; k->w = za.W; k->npix = za.N * za.H * za.W;
  k->fd_w = make_fastdiv((uint32_t)za.W); k->fd_h = make_fastdiv((uint32_t)za.H);
  k->fd_w2 = make_fastdiv((uint32_t)(2 * za.W)); k->fd_h2 = make_fastdiv((uint32_t)(2 * za.H));
  if (!bwd) {
    if (!d->out.ptr || d->out.C != d->skip.C || d->out.H != d->skip.H || d->out.W != d->skip.W) { set_error("gate: bad output view"); return -1; }
  } else {
    if (!d->training) { set_error("gate: backward of an inference plan"); return -1; }
    if (!d->dout.ptr || !d->dza.ptr || !d->dzb.ptr || !d->dr || !d->g3 || !d->bsums3 || !d->bsums_ab || !d->dgamma_a || !d->dbeta_a ||
        !d->dgamma_b || !d->dbeta_b || !d->dw3 || !d->dgamma3 || !d->dbeta3 || !d->dwt || !d->dbt) { set_error("gate: null backward pointer"); return -1; }
  }
  return 0;
}

struct GateLaunch : PreparedOp {
  GateK k;
  bool bwd;
  int mode = 0;
  int launch(cudaStream_t s) override {
    const int sms = num_sms();
    const int total = k.skip.N * 4 * k.h * k.w;
    int P = kGateP;                                   // pixels per block iteration: smaller when the map would not fill the SMs
    while (P > 32 && total / P < sms * 4) P >>= 1;
    const int chunks = (total + P - 1) / P;
    const int grid_out = chunks < sms * 8 ? chunks : sms * 8;
    int vshift = 0;
    while ((8 << vshift) < k.skip.C) ++vshift;
    const int vpp = k.C / 8;
    if (!bwd) {
      const int lpp = vpp < 32 ? vpp : 32, ppw = 32 / lpp;
      const int groups = (k.npix + ppw - 1) / ppw;
      const int gu = vpp / lpp == 4 ? 1 : vpp / lpp == 2 ? 2 : 4;      // = GU of the kernel
      int grid = (groups + 8 * gu - 1) / (8 * gu);      // 8 warps x GU groups per block iteration
      if (grid < 1) grid = 1;
      if (grid > sms * 4) grid = sms * 4;
      const int smem = 4 * k.C * 4;
      const int vpl = vpp / lpp;
      auto fwd = vpl == 1 ? gate_mid_fwd_kernel<1> : vpl == 2 ? gate_mid_fwd_kernel<2> : vpl == 4 ? gate_mid_fwd_kernel<4> : gate_mid_fwd_kernel<0>;
      if (smem > 48 * 1024) B2_CUDA_OK(cudaFuncSetAttribute(fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      B2_CUDA_OK(launch_k(fwd, dim3(grid), dim3(256), smem, s, k));
      B2_CUDA_OK(launch_k(gate_out_kernel<0>, dim3(grid_out), dim3(256), 0, s, k, P, vshift));
    } else if (mode == 2) {
      B2_CUDA_OK(launch_k(gate_out_kernel<2>, dim3(grid_out), dim3(256), 0, s, k, P, vshift));
    } else {
      B2_CUDA_OK(cudaMemsetAsync(k.bsums3, 0, 8, s));
      B2_CUDA_OK(cudaMemsetAsync(k.bsums_ab, 0, (size_t)3 * k.C * 4, s));
      B2_CUDA_OK(launch_k(gate_out_kernel<1>, dim3(grid_out), dim3(256), 0, s, k, P, vshift));
      int grid_low = (k.npix + 255) / 256;
      if (grid_low > sms * 4) grid_low = sms * 4;
      B2_CUDA_OK(launch_k(gate_low_bwd_kernel, dim3(grid_low), dim3(256), 0, s, k));
      const int cvb = vpp < 32 ? vpp : 32, rows = 256 / cvb;
      const int gx = vpp / cvb;
      int gy = (k.npix + rows - 1) / rows;
      const int cap = (sms * 4 + gx - 1) / gx;
      if (gy > cap) gy = cap;
      static bool attr_set = false;
      if (!attr_set) {
        B2_CUDA_OK(cudaFuncSetAttribute(gate_mid_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 33 * 4));
        attr_set = true;
      }
      B2_CUDA_OK(launch_k(gate_mid_bwd_kernel<0>, dim3(gx, gy), dim3(256), 256 * 33 * 4, s, k, cvb));
      B2_CUDA_OK(launch_k(gate_mid_bwd_kernel<1>, dim3(gx, gy), dim3(256), 0, s, k, cvb));
    }
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
  int num_launches() const override { return mode == 2 ? 1 : (bwd ? 4 : 2); }
};

PreparedOp* prepare_gate_fwd(const b2seg_gate_desc* d) {
  auto* L = new GateLaunch();
  L->bwd = false;
  if (fill_gate(d, &L->k, false) != 0) { delete L; return nullptr; }
  return L;
}
PreparedOp* prepare_gate_bwd(const b2seg_gate_desc* d) {
  auto* L = new GateLaunch();
  L->bwd = true;
  if (fill_gate(d, &L->k, true) != 0) { delete L; return nullptr; }
  if (d->da_low.ptr) {     // second pass only: dskip = dout * r + scatter(da_low)
    const b2seg_view& v = d->da_low;
    if (!d->dskip.ptr || v.N != d->za.N || v.H != d->za.H || v.W != d->za.W || v.C != d->skip.C) { set_error("gate: bad da_low / dskip views"); delete L; return nullptr; }
    L->mode = 2;
  }
  return L;
}

}  // namespace b2
