// HBM-streaming kernels of the hot path: BatchNorm finalize / apply (+activation, +max-pool) / backward,
// fused Adam, the small-Cout pointwise heads, the loss seed, element-wise glue and input cast.
// All activation tensors are NHWC bf16 addressed through b2seg_view; every thread moves 16-byte vectors
// (8 channels) so global accesses are coalesced along the channel axis.
#include "stream_common.cuh"

namespace b2 {

// ------------------------------------------------------------------------------------------ cast input
__global__ void cast_input_kernel(const float* __restrict__ src, int N, int H, int W, int C, DView out) {
  const int cv = out.C / 8;
  const unsigned total = (unsigned)N * H * W * cv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    unsigned pix = i / cv;
    const int w = (int)(pix % W); pix /= W;
    const int h = (int)(pix % H);
    const int n = (int)(pix / H);
    const float* sp = src + (((long long)n * H + h) * W + w) * C;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = (v * 8 + e < C) ? __ldg(sp + v * 8 + e) : 0.f;
    store8(vaddr(out, n, h, w, v * 8), f);
  }
}
// K-packed im2col of the network input: out(n,h,w, (i*kw + j)*C + c) = src(n, h + i - (kh-1)/2, w + j - (kw-1)/2, c), zero outside the
// image and in the padding lanes.  A k x k convolution that reads the (thin: 1-3 channel) input becomes a 1x1 convolution over
// this tensor: ONE 128-byte-row tap on the tensor cores instead of kh*kw taps that each hold 8 real channels (the 3 -> 64 first
// layer of config 2 took 0.22 ms forward + 0.25 ms weight gradient for 0.02 ms of math, profiles/r1_tile_trace.txt).
// One thread per output pixel: the kh rows of its window are kw*C consecutive floats each (neighbouring threads overlap, L1 serves
// the re-reads), written as KP/8 16-byte vectors — consecutive threads write consecutive KP*2-byte records.  (A first version with
// one thread per output vector and per-element index arithmetic ran at 0.6 TB/s: 0.26 ms for the 134 MB of config 2.)
template <int KH, int KW, int C, int KP>
__global__ void __launch_bounds__(256) im2col_input_kernel_t(const float* __restrict__ src, int N, int H, int W, DView out) {
  constexpr int PH = (KH - 1) / 2, PW = (KW - 1) / 2;
  const unsigned total = (unsigned)N * H * W;
  for (unsigned pix = blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += gridDim.x * blockDim.x) {
    const int w = (int)(pix % W);
    const unsigned t = pix / W;
    const int h = (int)(t % H), n = (int)(t / H);
    float f[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) f[k] = 0.f;
#pragma unroll
    for (int i = 0; i < KH; ++i) {
      const int hh = h + i - PH;
      if (hh < 0 || hh >= H) continue;
      const float* row = src + (((long long)n * H + hh) * W + (w - PW)) * C;
#pragma unroll
      for (int j = 0; j < KW; ++j) {
        const int ww = w + j - PW;
        if (ww < 0 || ww >= W) continue;
#pragma unroll
        for (int c = 0; c < C; ++c) f[(i * KW + j) * C + c] = __ldg(row + j * C + c);
      }
    }
#pragma unroll
    for (int v = 0; v < KP / 8; ++v) {
      float g[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = f[v * 8 + e];
      store8(vaddr(out, n, h, w, v * 8), g);
    }
  }
}
__global__ void im2col_input_kernel(const float* __restrict__ src, int N, int H, int W, int C, int kh, int kw, DView out) {
  const int cv = out.C / 8;
  const int K = kh * kw * C;
  const int ph = (kh - 1) / 2, pw = (kw - 1) / 2;
  const unsigned total = (unsigned)N * H * W * cv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    unsigned pix = i / cv;
    const int w = (int)(pix % W); pix /= W;
    const int h = (int)(pix % H);
    const int n = (int)(pix / H);
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = v * 8 + e;
      float val = 0.f;
      if (k < K) {
        const int tap = k / C, c = k - tap * C;
        const int hh = h + tap / kw - ph, ww = w + tap % kw - pw;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) val = __ldg(src + (((long long)n * H + hh) * W + ww) * C + c);
      }
      f[e] = val;
    }
    store8(vaddr(out, n, h, w, v * 8), f);
  }
}
struct CastLaunch : PreparedOp {
  b2seg_cast_desc d;
  int launch(cudaStream_t s) override {
    const long long work = (long long)d.N * d.H * d.W * (d.out.C / 8);
    const float* src = reinterpret_cast<const float*>(d.src);
    const int gp = grid_for((long long)d.N * d.H * d.W, 256);
    if (d.kh == 3 && d.kw == 3 && d.C == 3 && d.out.C == 32)
      im2col_input_kernel_t<3, 3, 3, 32><<<gp, 256, 0, s>>>(src, d.N, d.H, d.W, dv(d.out));
    else if (d.kh == 3 && d.kw == 3 && d.C == 1 && d.out.C == 16)
      im2col_input_kernel_t<3, 3, 1, 16><<<gp, 256, 0, s>>>(src, d.N, d.H, d.W, dv(d.out));
    else if (d.kh == 1 && d.kw == 3 && d.C == 1 && d.out.C == 8)
      im2col_input_kernel_t<1, 3, 1, 8><<<gp, 256, 0, s>>>(src, d.N, d.H, d.W, dv(d.out));
    else if (d.kh * d.kw > 1)
      im2col_input_kernel<<<grid_for(work, 256), 256, 0, s>>>(src, d.N, d.H, d.W, d.C, d.kh, d.kw, dv(d.out));
    else
      cast_input_kernel<<<grid_for(work, 256), 256, 0, s>>>(reinterpret_cast<const float*>(d.src), d.N, d.H, d.W, d.C, dv(d.out));
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
PreparedOp* prepare_cast(const b2seg_cast_desc* d) {
  if (d->out.C % 8) { set_error("cast: out.C %% 8"); return nullptr; }
  if (d->kh < 0 || d->kw < 0 || (d->kh * d->kw > 1 && d->kh * d->kw * d->C > d->out.C)) { set_error("cast: im2col window %dx%dx%d does not fit %d channels", d->kh, d->kw, d->C, d->out.C); return nullptr; }
  CastLaunch* L = new CastLaunch(); L->d = *d; return L;
}

// ------------------------------------------------------------------------------------------ BN finalize
__global__ void bn_finalize_kernel(b2seg_bn_finalize_desc d) {
  pdl_prologue();
  __shared__ double sh_s[32][33], sh_q[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const float* part = reinterpret_cast<const float*>(d.partials);
  double s = 0.0, q = 0.0;
  if (c < d.C && !d.inference) {
    for (int pp = ty; pp < d.n_partials; pp += 32) {
      s += (double)part[(size_t)pp * 2 * d.C + c];
      q += (double)part[(size_t)pp * 2 * d.C + d.C + c];
    }
  }
  sh_s[ty][tx] = s;
  sh_q[ty][tx] = q;
  __syncthreads();
  if (ty == 0 && c < d.C) {
    float* mm = reinterpret_cast<float*>(d.moving_mean);
    float* mv = reinterpret_cast<float*>(d.moving_var);
    const float gamma = d.gamma ? reinterpret_cast<const float*>(d.gamma)[c] : 1.f;
    const float beta = d.beta ? reinterpret_cast<const float*>(d.beta)[c] : 0.f;
    double mean, var;
    if (d.inference) {
      mean = mm[c];
      var = mv[c];
    } else {
      for (int i = 1; i < 32; ++i) { s += sh_s[i][tx]; q += sh_q[i][tx]; }
      mean = s / d.count;
      var = q / d.count - mean * mean;
      if (var < 0.0) var = 0.0;
      if (d.update_moving) {
        const double uv = (d.bessel && d.count > 1.0) ? var * d.count / (d.count - 1.0) : var;
        mm[c] = (float)(mm[c] * (double)d.momentum + mean * (1.0 - (double)d.momentum));
        mv[c] = (float)(mv[c] * (double)d.momentum + uv * (1.0 - (double)d.momentum));
      }
    }
    const double rstd = 1.0 / sqrt(var + (double)d.eps);
    reinterpret_cast<float*>(d.scale)[c] = (float)(gamma * rstd);
    reinterpret_cast<float*>(d.shift)[c] = (float)(beta - mean * gamma * rstd);
    if (d.mean) reinterpret_cast<float*>(d.mean)[c] = (float)mean;
    if (d.rstd) reinterpret_cast<float*>(d.rstd)[c] = (float)rstd;
  }
}
struct BnFinalizeLaunch : PreparedOp {
  b2seg_bn_finalize_desc d;
  int launch(cudaStream_t s) override {
    launch_k(bn_finalize_kernel, dim3((d.C + 31) / 32), dim3(1024), 0, s, d);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
PreparedOp* prepare_bn_finalize(const b2seg_bn_finalize_desc* d) { auto* L = new BnFinalizeLaunch(); L->d = *d; return L; }

// ------------------------------------------------------------------------------------------ BN apply + act (+pool)
struct BnActK {
  DView x, out0, out1, pooled;
  const float* scale; const float* shift;
  int act, n_out, ph, pw, c_valid;
};
// One thread owns 8 channels of U windows (U*WIN pixels in flight: all loads are issued before any use).
template <int WIN, int U>
__global__ void __launch_bounds__(256, 4) bn_act_kernel(BnActK k) {
  const int cv = k.x.C / 8;
  const int Ho = k.x.H / k.ph, Wo = k.x.W / k.pw;
  const unsigned n_win = (unsigned)k.x.N * Ho * Wo;
  const unsigned total = ((n_win + U - 1) / U) * cv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    const unsigned wbase = (i / cv) * U;
    float sc[8], sf[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sc[e] = k.scale ? __ldg(k.scale + v * 8 + e) : 1.f;
      sf[e] = k.shift ? __ldg(k.shift + v * 8 + e) : 0.f;
    }
    float f[U][WIN][8];
    int wn[U], wh[U], ww[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      unsigned t = wbase + u < n_win ? wbase + u : n_win - 1;
      ww[u] = (int)(t % Wo); t /= Wo;
      wh[u] = (int)(t % Ho);
      wn[u] = (int)(t / Ho);
#pragma unroll
      for (int q = 0; q < WIN; ++q) load8(vaddr(k.x, wn[u], wh[u] * k.ph + q / k.pw, ww[u] * k.pw + q % k.pw, v * 8), f[u][q]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (wbase + u >= n_win) break;
      float mx[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) mx[e] = -INFINITY;
#pragma unroll
      for (int q = 0; q < WIN; ++q) {
        const int h = wh[u] * k.ph + q / k.pw, w = ww[u] * k.pw + q % k.pw;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          f[u][q][e] = (k.c_valid && v * 8 + e >= k.c_valid) ? 0.f : act_fwd(fmaf(f[u][q][e], sc[e], sf[e]), k.act);
          mx[e] = fmaxf(mx[e], f[u][q][e]);
        }
        if (k.n_out > 0) store8(vaddr(k.out0, wn[u], h, w, v * 8), f[u][q]);
        if (k.n_out > 1) store8(vaddr(k.out1, wn[u], h, w, v * 8), f[u][q]);
      }
      if (k.pooled.ptr) store8(vaddr(k.pooled, wn[u], wh[u], ww[u], v * 8), mx);
    }
  }
}
// generic window size (UNet3+ pools 4/8/16): sequential over the window
__global__ void bn_act_generic_kernel(BnActK k) {
  const int cv = k.x.C / 8;
  const int Ho = k.x.H / k.ph, Wo = k.x.W / k.pw;
  const unsigned total = (unsigned)k.x.N * Ho * Wo * cv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    unsigned pix = i / cv;
    const int wo = (int)(pix % Wo); pix /= Wo;
    const int ho = (int)(pix % Ho);
    const int n = (int)(pix / Ho);
    float sc[8], sf[8], mx[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sc[e] = k.scale ? __ldg(k.scale + v * 8 + e) : 1.f;
      sf[e] = k.shift ? __ldg(k.shift + v * 8 + e) : 0.f;
      mx[e] = -INFINITY;
    }
    for (int a = 0; a < k.ph; ++a)
      for (int b = 0; b < k.pw; ++b) {
        const int h = ho * k.ph + a, w = wo * k.pw + b;
        float f[8];
        load8(vaddr(k.x, n, h, w, v * 8), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          f[e] = (k.c_valid && v * 8 + e >= k.c_valid) ? 0.f : act_fwd(fmaf(f[e], sc[e], sf[e]), k.act);
          mx[e] = fmaxf(mx[e], f[e]);
        }
        if (k.n_out > 0) store8(vaddr(k.out0, n, h, w, v * 8), f);
        if (k.n_out > 1) store8(vaddr(k.out1, n, h, w, v * 8), f);
      }
    if (k.pooled.ptr) store8(vaddr(k.pooled, n, ho, wo, v * 8), mx);
  }
}
struct BnActLaunch : PreparedOp {
  BnActK k;
  int launch(cudaStream_t s) override {
    const long long n_win = (long long)k.x.N * (k.x.H / k.ph) * (k.x.W / k.pw);
    const int cv = k.x.C / 8;
    const int win = k.ph * k.pw;
    const int cap = num_sms() * 16;
    if (win == 1) {
      int g = grid_for(((n_win + 3) / 4) * cv, 256); if (g > cap) g = cap;
      bn_act_kernel<1, 4><<<g, 256, 0, s>>>(k);
    } else if (win == 2) {
      int g = grid_for(((n_win + 1) / 2) * cv, 256); if (g > cap) g = cap;
      bn_act_kernel<2, 2><<<g, 256, 0, s>>>(k);
    } else if (win == 4) {
      int g = grid_for(n_win * cv, 256); if (g > cap) g = cap;
      bn_act_kernel<4, 1><<<g, 256, 0, s>>>(k);
    } else {
      int g = grid_for(n_win * cv, 256); if (g > cap) g = cap;
      bn_act_generic_kernel<<<g, 256, 0, s>>>(k);
    }
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
PreparedOp* prepare_bn_act(const b2seg_bn_act_desc* d) {
  if (d->x.C % 8) { set_error("bn_act: C %% 8"); return nullptr; }
  if (PreparedOp* fast = prepare_bn_act_fast(d)) return fast;   // instruction-lean row walker (stream_fast.cu) when eligible
  if (d->add.ptr || d->out_stats) { set_error("bn_act: the fused addend / output statistics need the row-walking fast path (no pooling, no channel mask, ReLU / LeakyReLU / none)"); return nullptr; }
  auto* L = new BnActLaunch();
  BnActK& k = L->k;
  memset(&k, 0, sizeof(k));
  k.x = dv(d->x);
  k.scale = reinterpret_cast<const float*>(d->scale);
  k.shift = reinterpret_cast<const float*>(d->shift);
  k.act = d->act;
  k.n_out = d->n_out;
  if (d->n_out > 0) k.out0 = dv(d->out[0]);
  if (d->n_out > 1) k.out1 = dv(d->out[1]);
  k.ph = d->pool_h > 1 ? d->pool_h : 1;
  k.pw = d->pool_w > 1 ? d->pool_w : 1;
  if (k.ph > 1 || k.pw > 1) k.pooled = dv(d->pooled);
  k.c_valid = d->c_valid;
  return L;
}

// ------------------------------------------------------------------------------------------ BN backward
struct GradSrcK { DView g; int kind; };
struct BnBwdK {
  DView x, dx;
  const float* scale; const float* shift; const float* mean; const float* rstd;
  int act, n_src;
  GradSrcK src[B2SEG_MAX_GRADSRC];
  int ph, pw;        // window processed per thread (pool window if a pooled source exists, else 1x1)
  float inv_count;
  float* partials;   // [n_blocks][2][C]
  int n_blocks;
  float* dgamma; float* dbeta;
  int cvb, rp;       // channel vectors per block-row, pixel rows per block
};

// One thread owns 8 channels of U windows per iteration (WIN = ph*pw pixels each; U*WIN = 4 pixels in flight, every load
// issued before the first use).  PASS 0: per-channel partial sums of g and g*xhat (g = summed incoming gradient * act'(y)).
// PASS 1: writes dx.
template <int PASS, int WIN, int U>
__global__ void __launch_bounds__(256) bn_bwd_kernel(BnBwdK k) {
  extern __shared__ float red[];  // PASS 0: [256][16]
  const int cvec_total = k.x.C / 8;
  const int tcv = threadIdx.x % k.cvb, trow = threadIdx.x / k.cvb;
  const int v = blockIdx.x * k.cvb + tcv;
  const bool active = trow < k.rp && v < cvec_total;
  const int Ho = k.x.H / k.ph, Wo = k.x.W / k.pw;
  const unsigned n_win = (unsigned)k.x.N * Ho * Wo;
  const unsigned n_grp = (n_win + U - 1) / U;
  float sc[8], sf[8], mu[8], rs[8], cb[8], cg[8];
  float acc_b[8], acc_g[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { acc_b[e] = 0.f; acc_g[e] = 0.f; }
  if (active) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = v * 8 + e;
      sc[e] = k.scale ? __ldg(k.scale + c) : 1.f;
      sf[e] = k.shift ? __ldg(k.shift + c) : 0.f;
      mu[e] = k.mean ? __ldg(k.mean + c) : 0.f;
      rs[e] = k.rstd ? __ldg(k.rstd + c) : 1.f;
      if (PASS == 1 && k.scale) {
        cb[e] = __ldg(k.dbeta + c) * k.inv_count;
        cg[e] = __ldg(k.dgamma + c) * k.inv_count;
      } else { cb[e] = 0.f; cg[e] = 0.f; }
    }
    for (unsigned grp = blockIdx.y * k.rp + trow; grp < n_grp; grp += gridDim.y * k.rp) {
      float xv[U][WIN][8], g[U][WIN][8], pooled[U][8];
      int wn[U], wh[U], ww[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        unsigned t = grp * U + u < n_win ? grp * U + u : n_win - 1;
        ww[u] = (int)(t % Wo); t /= Wo;
        wh[u] = (int)(t % Ho);
        wn[u] = (int)(t / Ho);
#pragma unroll
        for (int q = 0; q < WIN; ++q) {
          load8(vaddr(k.x, wn[u], wh[u] * k.ph + q / k.pw, ww[u] * k.pw + q % k.pw, v * 8), xv[u][q]);
#pragma unroll
          for (int e = 0; e < 8; ++e) g[u][q][e] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) pooled[u][e] = 0.f;
      }
      for (int s = 0; s < k.n_src; ++s) {
        if (k.src[s].kind == 0) {
          float f[U][WIN][8];
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < WIN; ++q) load8(vaddr(k.src[s].g, wn[u], wh[u] * k.ph + q / k.pw, ww[u] * k.pw + q % k.pw, v * 8), f[u][q]);
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < WIN; ++q)
#pragma unroll
              for (int e = 0; e < 8; ++e) g[u][q][e] += f[u][q][e];
        } else {
          float f[U][8];
#pragma unroll
          for (int u = 0; u < U; ++u) load8(vaddr(k.src[s].g, wn[u], wh[u], ww[u], v * 8), f[u]);
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int e = 0; e < 8; ++e) pooled[u][e] += f[u][e];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (grp * U + u >= n_win) break;
        // forward recompute: y, first arg-max of the window (TF/torch tie rule), activation derivative
        float y[WIN][8];
        int amax[8];
        float ymax[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { amax[e] = 0; ymax[e] = -INFINITY; }
#pragma unroll
        for (int q = 0; q < WIN; ++q)
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            y[q][e] = act_fwd(fmaf(xv[u][q][e], sc[e], sf[e]), k.act);
            if (y[q][e] > ymax[e]) { ymax[e] = y[q][e]; amax[e] = q; }
          }
#pragma unroll
        for (int q = 0; q < WIN; ++q) {
          float xh[8], gg[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            gg[e] = (g[u][q][e] + (amax[e] == q ? pooled[u][e] : 0.f)) * act_bwd_from_y(y[q][e], k.act);
            xh[e] = (xv[u][q][e] - mu[e]) * rs[e];
          }
          if (PASS == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { acc_b[e] += gg[e]; acc_g[e] = fmaf(gg[e], xh[e], acc_g[e]); }
          } else {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = k.scale ? sc[e] * (gg[e] - cb[e] - xh[e] * cg[e]) : gg[e];
            store8(vaddr(k.dx, wn[u], wh[u] * k.ph + q / k.pw, ww[u] * k.pw + q % k.pw, v * 8), o);
          }
        }
      }
    }
  }
  if (PASS == 0) {
    float* mine = red + (size_t)threadIdx.x * 16;
#pragma unroll
    for (int e = 0; e < 8; ++e) { mine[e] = active ? acc_b[e] : 0.f; mine[8 + e] = active ? acc_g[e] : 0.f; }
    __syncthreads();
    if (trow == 0 && v < cvec_total) {
      float sb[8], sg[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { sb[e] = 0.f; sg[e] = 0.f; }
      for (int r = 0; r < k.rp; ++r) {
        const float* o = red + (size_t)(r * k.cvb + tcv) * 16;
#pragma unroll
        for (int e = 0; e < 8; ++e) { sb[e] += o[e]; sg[e] += o[8 + e]; }
      }
      float* pp = k.partials + (size_t)blockIdx.y * 2 * k.x.C;
#pragma unroll
      for (int e = 0; e < 8; ++e) { pp[v * 8 + e] = sb[e]; pp[k.x.C + v * 8 + e] = sg[e]; }
    }
  }
}
// raw_gx = 1: partials hold (sum g, sum g*x) of the lean kernel; convert to dgamma = rstd * (sum g*x - mean * sum g)
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partials, int n_blocks, int C, float* dgamma, float* dbeta,
                                       const float* __restrict__ mean, const float* __restrict__ rstd, int raw_gx) {
  // one warp per 32 channels x 8 slices of the partial list
  __shared__ double sh[2][8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  double b = 0.0, g = 0.0;
  if (c < C)
    for (int i = ty; i < n_blocks; i += 8) {
      b += (double)partials[(size_t)i * 2 * C + c];
      g += (double)partials[(size_t)i * 2 * C + C + c];
    }
  sh[0][ty][tx] = b;
  sh[1][ty][tx] = g;
  __syncthreads();
  if (ty == 0 && c < C) {
    for (int i = 1; i < 8; ++i) { b += sh[0][i][tx]; g += sh[1][i][tx]; }
    if (raw_gx) g = (double)rstd[c] * (g - (double)mean[c] * b);
    dbeta[c] = (float)b;
    dgamma[c] = (float)g;
  }
}

int launch_bn_bwd_finalize(const float* partials, int n_blocks, int C, float* dgamma, float* dbeta, const float* mean, const float* rstd,
                           int raw_gx, cudaStream_t s) {
  bn_bwd_finalize_kernel<<<(C + 31) / 32, 256, 0, s>>>(partials, n_blocks, C, dgamma, dbeta, mean, rstd, raw_gx);
  B2_CUDA_OK(cudaGetLastError());
  return 0;
}

// Lean BN(+ReLU/LeakyReLU/identity) backward.  Per element: mask from the sign of t = x*scale+shift (for a pooled source the
// window arg-max of t, which equals the arg-max of relu(t) wherever the gradient survives the mask),
//   PASS 0:  sum g, sum g*x          PASS 1:  dx = A*g + B*x + D   with per-channel A = scale, B = -scale*cg*rstd,
//   D = scale*(cg*rstd*mean - cb), cb = dbeta/count, cg = dgamma/count  (== scale*(g - cb - xhat*cg)).
template <int PASS, int WIN, int U, int ACT>
__global__ void __launch_bounds__(256, WIN == 4 ? 2 : 4) bn_bwd_lean_kernel(BnBwdK k) {
  extern __shared__ float red[];  // PASS 0: [256][16]
  const int cvec_total = k.x.C / 8;
  const int tcv = threadIdx.x % k.cvb, trow = threadIdx.x / k.cvb;
  const int v = blockIdx.x * k.cvb + tcv;
  const bool active = trow < k.rp && v < cvec_total;
  const unsigned Ho = k.x.H / k.ph, Wo = k.x.W / k.pw;
  const unsigned n_win = (unsigned)k.x.N * Ho * Wo;
  const unsigned n_grp = (n_win + U - 1) / U;
  const bool has_bn = k.scale != nullptr;
  float sc[8], sf[8], cB[8], cD[8];
  float acc_b[8], acc_g[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { acc_b[e] = 0.f; acc_g[e] = 0.f; }
  if (active) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = v * 8 + e;
      sc[e] = has_bn ? __ldg(k.scale + c) : 1.f;
      sf[e] = has_bn ? __ldg(k.shift + c) : 0.f;
      cB[e] = 0.f; cD[e] = 0.f;
      if (PASS == 1 && has_bn) {
        const float cb = __ldg(k.dbeta + c) * k.inv_count, cg = __ldg(k.dgamma + c) * k.inv_count;
        const float rs = __ldg(k.rstd + c), mu = __ldg(k.mean + c);
        cB[e] = -sc[e] * cg * rs;
        cD[e] = sc[e] * (cg * rs * mu - cb);
      }
    }
    for (unsigned grp = blockIdx.y * k.rp + trow; grp < n_grp; grp += gridDim.y * k.rp) {
      uint4 xr[U][WIN];
      float g[U][WIN][8];
      float pooled[U][8];
      int wn[U], wh[U], ww[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        unsigned t = grp * U + u < n_win ? grp * U + u : n_win - 1;
        ww[u] = (int)(t % Wo); t /= Wo;
        wh[u] = (int)(t % Ho);
        wn[u] = (int)(t / Ho);
#pragma unroll
        for (int q = 0; q < WIN; ++q)
          xr[u][q] = __ldg(reinterpret_cast<const uint4*>(vaddr(k.x, wn[u], wh[u] * k.ph + q / k.pw, ww[u] * k.pw + q % k.pw, v * 8)));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int e = 0; e < 8; ++e) pooled[u][e] = 0.f;
#pragma unroll
        for (int q = 0; q < WIN; ++q)
#pragma unroll
          for (int e = 0; e < 8; ++e) g[u][q][e] = 0.f;
      }
      for (int s = 0; s < k.n_src; ++s) {
        if (k.src[s].kind == 0) {
          uint4 r[U][WIN];
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < WIN; ++q)
              r[u][q] = __ldg(reinterpret_cast<const uint4*>(vaddr(k.src[s].g, wn[u], wh[u] * k.ph + q / k.pw, ww[u] * k.pw + q % k.pw, v * 8)));
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int q = 0; q < WIN; ++q) {
              const uint32_t w4[4] = {r[u][q].x, r[u][q].y, r[u][q].z, r[u][q].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                g[u][q][2 * e] += __uint_as_float(w4[e] << 16);
                g[u][q][2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
              }
            }
        } else if (WIN > 1) {
          uint4 r[U];
#pragma unroll
          for (int u = 0; u < U; ++u) r[u] = __ldg(reinterpret_cast<const uint4*>(vaddr(k.src[s].g, wn[u], wh[u], ww[u], v * 8)));
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const uint32_t w4[4] = {r[u].x, r[u].y, r[u].z, r[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              pooled[u][2 * e] += __uint_as_float(w4[e] << 16);
              pooled[u][2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (grp * U + u >= n_win) break;
        float x[WIN][8];
#pragma unroll
        for (int q = 0; q < WIN; ++q) {
          const uint32_t w4[4] = {xr[u][q].x, xr[u][q].y, xr[u][q].z, xr[u][q].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            x[q][2 * e] = __uint_as_float(w4[e] << 16);
            x[q][2 * e + 1] = __uint_as_float(w4[e] & 0xffff0000u);
          }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float t[WIN];
#pragma unroll
          for (int q = 0; q < WIN; ++q) t[q] = fmaf(x[q][e], sc[e], sf[e]);
          if (WIN > 1) {
            float m = t[0];
#pragma unroll
            for (int q = 1; q < WIN; ++q) m = fmaxf(m, t[q]);
            bool taken = false;
#pragma unroll
            for (int q = 0; q < WIN; ++q) {
              const bool hit = !taken && t[q] == m;
              taken = taken || hit;
              if (hit) g[u][q][e] += pooled[u][e];
            }
          }
#pragma unroll
          for (int q = 0; q < WIN; ++q) {
            if (ACT == B2SEG_ACT_RELU) g[u][q][e] = t[q] > 0.f ? g[u][q][e] : 0.f;
            if (ACT == B2SEG_ACT_LEAKY) g[u][q][e] = t[q] > 0.f ? g[u][q][e] : 0.3f * g[u][q][e];
          }
        }
#pragma unroll
        for (int q = 0; q < WIN; ++q) {
          if (PASS == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { acc_b[e] += g[u][q][e]; acc_g[e] = fmaf(g[u][q][e], x[q][e], acc_g[e]); }
          } else {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = has_bn ? fmaf(cB[e], x[q][e], fmaf(sc[e], g[u][q][e], cD[e])) : g[u][q][e];
            store8(vaddr(k.dx, wn[u], wh[u] * k.ph + q / k.pw, ww[u] * k.pw + q % k.pw, v * 8), o);
          }
        }
      }
    }
  }
  if (PASS == 0) {
    float* mine = red + (size_t)threadIdx.x * 16;
#pragma unroll
    for (int e = 0; e < 8; ++e) { mine[e] = active ? acc_b[e] : 0.f; mine[8 + e] = active ? acc_g[e] : 0.f; }
    __syncthreads();
    if (trow == 0 && v < cvec_total) {
      float sb[8], sg[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { sb[e] = 0.f; sg[e] = 0.f; }
      for (int r = 0; r < k.rp; ++r) {
        const float* o = red + (size_t)(r * k.cvb + tcv) * 16;
#pragma unroll
        for (int e = 0; e < 8; ++e) { sb[e] += o[e]; sg[e] += o[8 + e]; }
      }
      float* pp = k.partials + (size_t)blockIdx.y * 2 * k.x.C;
#pragma unroll
      for (int e = 0; e < 8; ++e) { pp[v * 8 + e] = sb[e]; pp[k.x.C + v * 8 + e] = sg[e]; }
    }
  }
}
struct BnBwdLaunch : PreparedOp {
  BnBwdK k;
  bool has_bn, lean;
  dim3 grid0, grid1;
  int smem0, win;
  template <int PASS>
  int go(dim3 grid, int smem, cudaStream_t s) {
    if (lean) {
#define B2_LEAN(ACT)                                                                      \
      switch (win) {                                                                      \
        case 1: bn_bwd_lean_kernel<PASS, 1, 2, ACT><<<grid, 256, smem, s>>>(k); break;    \
        case 2: bn_bwd_lean_kernel<PASS, 2, 1, ACT><<<grid, 256, smem, s>>>(k); break;    \
        default: bn_bwd_lean_kernel<PASS, 4, 1, ACT><<<grid, 256, smem, s>>>(k); break;   \
      }
      if (k.act == B2SEG_ACT_RELU) { B2_LEAN(B2SEG_ACT_RELU) }
      else if (k.act == B2SEG_ACT_LEAKY) { B2_LEAN(B2SEG_ACT_LEAKY) }
      else { B2_LEAN(B2SEG_ACT_NONE) }
#undef B2_LEAN
    } else {
      switch (win) {
        case 1: bn_bwd_kernel<PASS, 1, 4><<<grid, 256, smem, s>>>(k); break;
        case 2: bn_bwd_kernel<PASS, 2, 2><<<grid, 256, smem, s>>>(k); break;
        default: bn_bwd_kernel<PASS, 4, 1><<<grid, 256, smem, s>>>(k); break;
      }
    }
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
  int launch(cudaStream_t s) override {
    if (has_bn) {
      int rc = go<0>(grid0, smem0, s);
      if (rc) return rc;
      bn_bwd_finalize_kernel<<<(k.x.C + 31) / 32, 256, 0, s>>>(k.partials, k.n_blocks, k.x.C, k.dgamma, k.dbeta, k.mean, k.rstd, lean ? 1 : 0);
      B2_CUDA_OK(cudaGetLastError());
    }
    return go<1>(grid1, 0, s);
  }
  int num_launches() const override { return has_bn ? 3 : 1; }
};
PreparedOp* prepare_bn_bwd(const b2seg_bn_bwd_desc* d) {
  if (d->x.C % 8 || d->n_src < 1 || d->n_src > B2SEG_MAX_GRADSRC) { set_error("bn_bwd: bad C or n_src"); return nullptr; }
  if (PreparedOp* fast = prepare_bn_bwd_fast(d)) return fast;   // instruction-lean row walker (stream_fast.cu) when eligible
  if (d->x_relu_mask) { set_error("bn_bwd: x_relu_mask needs the row-walking fast path (BatchNorm present, no head source, views below 2^31 elements)"); return nullptr; }
  for (int i = 0; i < d->n_src; ++i)
    if (d->src[i].kind == 2) {
      set_error("bn_bwd: a pointwise-head source (kind 2) needs BN + ReLU/LeakyReLU, no pooled source, cout <= 2 and 16-byte aligned views below 2^31 elements");
      return nullptr;
    }
  auto* L = new BnBwdLaunch();
  BnBwdK& k = L->k;
  memset(&k, 0, sizeof(k));
  k.x = dv(d->x); k.dx = dv(d->dx);
  k.scale = reinterpret_cast<const float*>(d->scale);
  k.shift = reinterpret_cast<const float*>(d->shift);
  k.mean = reinterpret_cast<const float*>(d->mean);
  k.rstd = reinterpret_cast<const float*>(d->rstd);
  k.act = d->act;
  k.n_src = d->n_src;
  k.ph = 1; k.pw = 1;
  for (int i = 0; i < d->n_src; ++i) {
    k.src[i].g = dv(d->src[i].g);
    k.src[i].kind = d->src[i].kind;
    if (d->src[i].kind == 1) {
      const int ph = d->src[i].pool_h > 1 ? d->src[i].pool_h : 1, pw = d->src[i].pool_w > 1 ? d->src[i].pool_w : 1;
      if ((k.ph != 1 || k.pw != 1) && (k.ph != ph || k.pw != pw)) { set_error("bn_bwd: mixed pool windows"); delete L; return nullptr; }
      k.ph = ph; k.pw = pw;
    }
  }
  L->win = k.ph * k.pw;
  L->lean = (k.act == B2SEG_ACT_NONE || k.act == B2SEG_ACT_RELU || k.act == B2SEG_ACT_LEAKY);
  if (L->win != 1 && L->win != 2 && L->win != 4) { set_error("bn_bwd: pool window must have 1, 2 or 4 elements"); delete L; return nullptr; }
  k.inv_count = (float)(1.0 / d->count);
  k.partials = reinterpret_cast<float*>(d->partials);
  k.n_blocks = d->n_blocks;
  k.dgamma = reinterpret_cast<float*>(d->dgamma);
  k.dbeta = reinterpret_cast<float*>(d->dbeta);
  L->has_bn = d->scale != 0;
  const int cvec = k.x.C / 8;
  k.cvb = cvec < 256 ? cvec : 256;
  k.rp = 256 / k.cvb;
  const int gx = (cvec + k.cvb - 1) / k.cvb;
  const int upt = L->lean ? (L->win == 1 ? 2 : 1) : 4 / L->win;  // windows per thread-iteration
  const long long n_win = ((long long)k.x.N * (k.x.H / k.ph) * (k.x.W / k.pw) + upt - 1) / upt;  // window groups
  L->grid0 = dim3(gx, d->n_blocks > 0 ? d->n_blocks : 1);
  long long gy1 = (n_win + k.rp - 1) / k.rp;
  const long long cap = (long long)num_sms() * 16 / gx + 1;
  if (gy1 > cap) gy1 = cap;
  if (gy1 > 65535) gy1 = 65535;
  if (gy1 < 1) gy1 = 1;
  L->grid1 = dim3(gx, (unsigned)gy1);
  L->smem0 = 256 * 16 * 4;
  if (L->has_bn && (d->n_blocks < 1 || d->n_blocks > 65535 || !d->partials || !d->dgamma || !d->dbeta)) {
    set_error("bn_bwd: partials/dgamma/dbeta/n_blocks required with BN");
    delete L;
    return nullptr;
  }
  return L;
}

// ------------------------------------------------------------------------------------------ Adam (Keras-2 rule)
struct AdamK { float* w; const float* g; float* m; float* v; __nv_bfloat16* wb; long long n; float alpha, b1, b2, eps, gs; };
__global__ void __launch_bounds__(128, 12) adam_kernel(AdamK k) {   // 128 threads x 40 registers: two blocks fit into the register file a persistent convolution CTA (576 x 96) leaves free, so a bucket's Adam runs UNDER the tensor-core kernels of backward
  pdl_prologue();
  const long long n4 = k.n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 w = reinterpret_cast<float4*>(k.w)[i];
    const float4 g4 = reinterpret_cast<const float4*>(k.g)[i];
    float4 m = reinterpret_cast<float4*>(k.m)[i];
    float4 v = reinterpret_cast<float4*>(k.v)[i];
    float* wp = &w.x; const float* gp = &g4.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float g = gp[e] * k.gs;
      mp[e] = k.b1 * mp[e] + (1.f - k.b1) * g;
      vp[e] = k.b2 * vp[e] + (1.f - k.b2) * g * g;
      wp[e] -= k.alpha * mp[e] / (sqrtf(vp[e]) + k.eps);
    }
    reinterpret_cast<float4*>(k.w)[i] = w;
    reinterpret_cast<float4*>(k.m)[i] = m;
    reinterpret_cast<float4*>(k.v)[i] = v;
    uint2 o;
    o.x = pack_bf16x2(w.x, w.y);
    o.y = pack_bf16x2(w.z, w.w);
    reinterpret_cast<uint2*>(k.wb)[i] = o;
  }
}
struct AdamLaunch : PreparedOp {
  b2seg_adam_desc d;
  int launch(cudaStream_t s) override {
    AdamK k;
    k.w = reinterpret_cast<float*>(d.w); k.g = reinterpret_cast<const float*>(d.g);
    k.m = reinterpret_cast<float*>(d.m); k.v = reinterpret_cast<float*>(d.v);
    k.wb = reinterpret_cast<__nv_bfloat16*>(d.w_bf16); k.n = d.n;
    const double t = (double)d.step;
    k.alpha = (float)((double)d.lr * sqrt(1.0 - pow((double)d.beta2, t)) / (1.0 - pow((double)d.beta1, t)));
    k.b1 = d.beta1; k.b2 = d.beta2; k.eps = d.eps; k.gs = d.grad_scale;
    int grid = grid_for(d.n / 4, 128);
    // blocks per SM: 12 fill an idle SM; beside a persistent convolution CTA only two fit, and more than that make the convolution
    // kernel that is launched next wait for Adam blocks to retire (B2SEG_ADAM_BLOCKS_PER_SM, A/B knob)
    static const int bps = getenv("B2SEG_ADAM_BLOCKS_PER_SM") ? atoi(getenv("B2SEG_ADAM_BLOCKS_PER_SM")) : 24;
    const int cap = num_sms() * (bps > 0 ? bps : 24);
    if (grid > cap) grid = cap;
    launch_k(adam_kernel, dim3(grid), dim3(128), 0, s, k);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
PreparedOp* prepare_adam(const b2seg_adam_desc* d) {
  if (d->n % 4) { set_error("adam: n must be a multiple of 4 (pad the flat buffer)"); return nullptr; }
  auto* L = new AdamLaunch(); L->d = *d; return L;
}
bool is_adam(PreparedOp* op) { return dynamic_cast<AdamLaunch*>(op) != nullptr; }
void adam_update(PreparedOp* op, float lr, int64_t step, float grad_scale) {
  if (auto* a = dynamic_cast<AdamLaunch*>(op)) { a->d.lr = lr; a->d.step = step; a->d.grad_scale = grad_scale; }
}

// ------------------------------------------------------------------------------------------ pointwise head
// One group of G lanes (G = 8..32) per output pixel: each lane covers 8 channels per step.  COUT is a template
// parameter so the per-class accumulators stay in registers and the shuffle reduction moves only live values.
template <int COUT>
__global__ void __launch_bounds__(256) head_fwd_kernel(DView x, const float* __restrict__ w, const float* __restrict__ b, int act, int stride,
                                                       float* __restrict__ y, float* __restrict__ logits, int G) {
  const int Ho = (x.H + stride - 1) / stride, Wo = (x.W + stride - 1) / stride;
  const unsigned n_pix = (unsigned)x.N * Ho * Wo;
  const int gl = threadIdx.x % G;
  const unsigned gid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const unsigned gstride = gridDim.x * blockDim.x / G;
  const int cvec = x.C / 8;
  const unsigned n_iter = (n_pix + gstride - 1) / gstride;  // uniform trip count: the shuffles below need whole warps
  for (unsigned it = 0; it < n_iter; ++it) {
    const unsigned pix = gid + it * gstride;
    const bool valid = pix < n_pix;
    unsigned t = valid ? pix : 0;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float acc[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
    for (int v = gl; v < cvec; v += G) {
      float f[8];
      load8(vaddr(x, n, ho * stride, wo * stride, v * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] = fmaf(f[e], __ldg(w + (size_t)(v * 8 + e) * COUT + o), acc[o]);
    }
    for (int off = G / 2; off > 0; off >>= 1)
#pragma unroll
      for (int o = 0; o < COUT; ++o) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off, 32);
    if (gl == 0 && valid) {
      float z[COUT];
      float zmax = -INFINITY;
#pragma unroll
      for (int o = 0; o < COUT; ++o) { z[o] = acc[o] + __ldg(b + o); zmax = fmaxf(zmax, z[o]); }
      if (logits) {
#pragma unroll
        for (int o = 0; o < COUT; ++o) logits[pix * COUT + o] = z[o];
      }
      if (act == B2SEG_ACT_SOFTMAX) {
        float den = 0.f;
#pragma unroll
        for (int o = 0; o < COUT; ++o) { z[o] = __expf(z[o] - zmax); den += z[o]; }
#pragma unroll
        for (int o = 0; o < COUT; ++o) y[pix * COUT + o] = z[o] / den;
      } else {
#pragma unroll
        for (int o = 0; o < COUT; ++o) y[pix * COUT + o] = act_fwd(z[o], act);
      }
    }
  }
}
static int head_group(int C) {
  int g = 8;
  while (g < 32 && g * 8 < C) g <<= 1;
  return g;
}
#define B2_COUT_SWITCH(cout, CALL)        \
  switch (cout) {                         \
    case 1: { constexpr int CO = 1; CALL; } break; \
    case 2: { constexpr int CO = 2; CALL; } break; \
    case 3: { constexpr int CO = 3; CALL; } break; \
    case 4: { constexpr int CO = 4; CALL; } break; \
    case 5: { constexpr int CO = 5; CALL; } break; \
    case 6: { constexpr int CO = 6; CALL; } break; \
    case 7: { constexpr int CO = 7; CALL; } break; \
    default: { constexpr int CO = 8; CALL; } break; \
  }
struct HeadFwdLaunch : PreparedOp {
  b2seg_head_desc d;
  int launch(cudaStream_t s) override {
    const int st = d.stride > 1 ? d.stride : 1;
    const long long n_pix = (long long)d.x.N * ((d.x.H + st - 1) / st) * ((d.x.W + st - 1) / st);
    const int G = head_group(d.x.C);
    int grid = grid_for(n_pix * G, 256);
    const int cap = num_sms() * 32;
    if (grid > cap) grid = cap;
    B2_COUT_SWITCH(d.cout, (head_fwd_kernel<CO><<<grid, 256, 0, s>>>(dv(d.x), reinterpret_cast<const float*>(d.w), reinterpret_cast<const float*>(d.b),
                                                                    d.act, st, reinterpret_cast<float*>(d.y), reinterpret_cast<float*>(d.logits), G)));
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
PreparedOp* prepare_head_fwd(const b2seg_head_desc* d) {
  if (d->cout < 1 || d->cout > 8 || d->x.C % 8) { set_error("head: cout in 1..8, C %% 8 == 0"); return nullptr; }
  if (PreparedOp* fast = prepare_head_fast(d, false)) return fast;
  if (d->bn_scale) { set_error("head: the fused BatchNorm prologue needs the pixel-contiguous fast path (C / 8 a power of two <= 32, stride 1)"); return nullptr; }
  auto* L = new HeadFwdLaunch(); L->d = *d; return L;
}

// backward: dx = dl . W^T (bf16); dW[c][o] += sum_pix x[c]*dl[o]; db[o] += sum_pix dl[o]
template <int COUT>
__global__ void __launch_bounds__(256) head_bwd_dx_kernel(DView x, DView dx, const float* __restrict__ w, int stride, const float* __restrict__ dl) {
  const int Ho = (x.H + stride - 1) / stride, Wo = (x.W + stride - 1) / stride;
  const int cvec = x.C / 8;
  const unsigned total = (unsigned)x.N * x.H * x.W * cvec;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = (int)(i % cvec);
    unsigned t = i / cvec;
    const int wq = (int)(t % x.W); t /= x.W;
    const int h = (int)(t % x.H);
    const int n = (int)(t / x.H);
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
    if (h % stride == 0 && wq % stride == 0) {
      const unsigned pix = ((unsigned)n * Ho + h / stride) * Wo + wq / stride;
#pragma unroll
      for (int q = 0; q < COUT; ++q) {
        const float d = __ldg(dl + pix * COUT + q);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(d, __ldg(w + (size_t)(v * 8 + e) * COUT + q), o[e]);
      }
    }
    store8(vaddr(dx, n, h, wq, v * 8), o);
  }
}
template <int COUT>
__global__ void __launch_bounds__(256) head_bwd_dw_kernel(DView x, int stride, const float* __restrict__ dl, float* dw, float* db, int cvb, int rp) {
  extern __shared__ float red[];  // [256][9] per output channel pass
  const int Ho = (x.H + stride - 1) / stride, Wo = (x.W + stride - 1) / stride;
  const unsigned n_pix = (unsigned)x.N * Ho * Wo;
  const int cvec = x.C / 8;
  const int tcv = threadIdx.x % cvb, trow = threadIdx.x / cvb;
  const int v = blockIdx.x * cvb + tcv;
  const bool active = trow < rp && v < cvec;
  float acc[COUT][8];
  float accb[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) { accb[o] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[o][e] = 0.f; }
  if (active) {
    for (unsigned pix = blockIdx.y * rp + trow; pix < n_pix; pix += gridDim.y * rp) {
      unsigned t = pix;
      const int wo = (int)(t % Wo); t /= Wo;
      const int ho = (int)(t % Ho);
      const int n = (int)(t / Ho);
      float f[8];
      load8(vaddr(x, n, ho * stride, wo * stride, v * 8), f);
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        const float d = __ldg(dl + pix * COUT + o);
        accb[o] += d;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[o][e] = fmaf(d, f[e], acc[o][e]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
    float* mine = red + (size_t)threadIdx.x * 9;
#pragma unroll
    for (int e = 0; e < 8; ++e) mine[e] = active ? acc[o][e] : 0.f;
    mine[8] = active ? accb[o] : 0.f;
    __syncthreads();
    if (trow == 0 && v < cvec) {
      float s[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) s[e] = 0.f;
      for (int r = 0; r < rp; ++r) {
        const float* q = red + (size_t)(r * cvb + tcv) * 9;
#pragma unroll
        for (int e = 0; e < 9; ++e) s[e] += q[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(dw + (size_t)(v * 8 + e) * COUT + o, s[e]);
      if (v == 0) atomicAdd(db + o, s[8]);
    }
    __syncthreads();
  }
}
struct HeadBwdLaunch : PreparedOp {
  b2seg_head_desc d;
  int launch(cudaStream_t s) override {
    const int st = d.stride > 1 ? d.stride : 1;
    if (d.dx.ptr) {
      const long long work = (long long)d.x.N * d.x.H * d.x.W * (d.x.C / 8);
      int g = grid_for(work, 256);
      const int cap = num_sms() * 32;
      if (g > cap) g = cap;
      B2_COUT_SWITCH(d.cout, (head_bwd_dx_kernel<CO><<<g, 256, 0, s>>>(dv(d.x), dv(d.dx), reinterpret_cast<const float*>(d.w), st,
                                                                      reinterpret_cast<const float*>(d.dlogits))));
      B2_CUDA_OK(cudaGetLastError());
    }
    const int cvec = d.x.C / 8;
    const int cvb = cvec < 256 ? cvec : 256, rp = 256 / cvb;
    const long long n_pix = (long long)d.x.N * ((d.x.H + st - 1) / st) * ((d.x.W + st - 1) / st);
    long long gy = (n_pix + rp * 64 - 1) / (rp * 64);
    if (gy > 1024) gy = 1024;
    if (gy < 1) gy = 1;
    dim3 grid((cvec + cvb - 1) / cvb, (unsigned)gy);
    B2_COUT_SWITCH(d.cout, (head_bwd_dw_kernel<CO><<<grid, 256, 256 * 9 * 4, s>>>(dv(d.x), st, reinterpret_cast<const float*>(d.dlogits),
                                                                                 reinterpret_cast<float*>(d.dw), reinterpret_cast<float*>(d.db), cvb, rp)));
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
  int num_launches() const override { return d.dx.ptr ? 2 : 1; }
};
PreparedOp* prepare_head_bwd(const b2seg_head_desc* d) {
  if (d->cout < 1 || d->cout > 8 || d->x.C % 8) { set_error("head: cout in 1..8, C %% 8 == 0"); return nullptr; }
  if (PreparedOp* fast = prepare_head_fast(d, true)) return fast;
  auto* L = new HeadBwdLaunch(); L->d = *d; return L;
}

// ------------------------------------------------------------------------------------------ loss seed
// One thread per pixel, a loop over the (few) output channels.  Every loss is written as  L = scale * sum_pix sum_o l(p, t)  with
// g = dl/dp the derivative with respect to the ACTIVATED output p; the seed dL/dlogits follows from the head's activation
// (none: g; sigmoid: g p (1 - p); softmax: p_j (g_j - sum_k g_k p_k)), except for the cross-entropies on their own activation,
// which Keras evaluates from the cached logits: seed (p - t).  Losses that reduce over the channel axis first (categorical
// hinge, cosine similarity, cross-entropy on un-normalised outputs) get their per-pixel reductions in `LossPix`.
struct LossPix { float a, b, c; int j; };

__device__ __forceinline__ float keras_label(float t) { return (t == 0.f || t == 1.f) ? 2.f * t - 1.f : t; }   // losses.py _maybe_convert_labels

// per-element loss and derivative; `prob` = cross-entropy on probabilities (the head is not the matching activation)
__device__ __forceinline__ void loss_elem(int kind, bool prob, float p, float t, const LossPix& px, int o, float* l, float* g) {
  const float eps = 1e-7f;
  switch (kind) {
    case 0: {   // binary cross-entropy
      if (!prob) {   // from the logits of the sigmoid head: only the value is needed here (seed = p - t)
        const float pc = fminf(fmaxf(p, eps), 1.f - eps);
        *l = -(t * logf(pc) + (1.f - t) * logf(1.f - pc)); *g = 0.f;
      } else {       // backend.binary_crossentropy on probabilities: clip, log(p + eps)
        const float pc = fminf(fmaxf(p, eps), 1.f - eps);
        *l = -(t * logf(pc + eps) + (1.f - t) * logf(1.f - pc + eps));
        *g = (p > eps && p < 1.f - eps) ? -(t / (pc + eps) - (1.f - t) / (1.f - pc + eps)) : 0.f;
      }
      break;
    }
    case 1: {   // categorical cross-entropy
      if (!prob) { *l = -t * logf(fmaxf(p, eps)); *g = 0.f; }
      else {      // output / sum(output), clipped: px.a = sum p
        const float q = p / px.a, qc = fminf(fmaxf(q, eps), 1.f - eps);
        *l = -t * logf(qc);
        *g = -((q > eps && q < 1.f - eps) ? t / p : 0.f) + px.c / px.a;     // px.c = sum of t over the unclipped channels
      }
      break;
    }
    case 2: { const float d = p - t; *l = d * d; *g = 2.f * d; break; }
    case 3: { const float d = p - t; *l = fabsf(d); *g = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); break; }
    case 4: {   // mean squared logarithmic error
      const float a = logf(fmaxf(p, eps) + 1.f), b = logf(fmaxf(t, eps) + 1.f);
      *l = (a - b) * (a - b); *g = p > eps ? 2.f * (a - b) / (p + 1.f) : 0.f; break;
    }
    case 5: {   // Huber, delta = 1
      const float d = p - t, ad = fabsf(d);
      if (ad <= 1.f) { *l = 0.5f * d * d; *g = d; } else { *l = ad - 0.5f; *g = d > 0.f ? 1.f : -1.f; }
      break;
    }
    case 6: {   // log-cosh: x + softplus(-2x) - log 2
      const float d = p - t, m2 = -2.f * d;
      const float sp = m2 > 15.f ? m2 : log1pf(__expf(m2));
      *l = d + sp - 0.69314718056f; *g = tanhf(d); break;
    }
    case 7: {   // binary focal cross-entropy, gamma = 2: (1 - p_t)^2 * bce
      const float pt = t * p + (1.f - t) * (1.f - p), om = 1.f - pt;
      const float pc = fminf(fmaxf(p, eps), 1.f - eps);
      if (!prob) {
        const float bce = -(t * logf(pc) + (1.f - t) * logf(1.f - pc));
        *l = om * om * bce;
        // caller multiplies g by p (1 - p) (sigmoid head): d/dz = -2 om (2t - 1) p(1-p) bce + om^2 (p - t)
        *g = -2.f * om * (2.f * t - 1.f) * bce + om * om * (p - t) / fmaxf(p * (1.f - p), 1e-30f);
      } else {
        const float bce = -(t * logf(pc + eps) + (1.f - t) * logf(1.f - pc + eps));
        const float dbce = (p > eps && p < 1.f - eps) ? -(t / (pc + eps) - (1.f - t) / (1.f - pc + eps)) : 0.f;
        *l = om * om * bce;
        *g = -2.f * om * (2.f * t - 1.f) * bce + om * om * dbce;
      }
      break;
    }
    case 8: { *l = p - t * logf(p + eps); *g = 1.f - t / (p + eps); break; }              // Poisson
    case 9: {   // Kullback-Leibler divergence (sum over channels)
      const float tc = fminf(fmaxf(t, eps), 1.f), pc = fminf(fmaxf(p, eps), 1.f);
      *l = tc * logf(tc / pc); *g = (p > eps && p < 1.f) ? -tc / pc : 0.f; break;
    }
    case 10: { const float y = keras_label(t), m = 1.f - y * p; *l = fmaxf(m, 0.f); *g = m > 0.f ? -y : 0.f; break; }                 // hinge
    case 11: { const float y = keras_label(t), m = fmaxf(1.f - y * p, 0.f); *l = m * m; *g = -2.f * m * y; break; }                    // squared hinge
    case 12: {  // mean absolute percentage error
      const float den = fmaxf(fabsf(t), eps), d = (t - p) / den;
      *l = 100.f * fabsf(d); *g = 100.f * (d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f)) / den; break;
    }
    case 13: {  // categorical hinge: max(neg - pos + 1, 0), pos = sum t p, neg = max (1 - t) p   (px.a = margin, px.j = arg max)
      *l = o == 0 ? fmaxf(px.a, 0.f) : 0.f;
      *g = px.a > 0.f ? ((o == px.j ? 1.f - t : 0.f) - t) : 0.f; break;
    }
    default: {  // 14: cosine similarity: -sum(l2norm(t) l2norm(p)); px.a = rsqrt(max(sum t^2, 1e-12)), px.b = rsqrt(max(sum p^2, 1e-12)), px.c = sum t p
      *l = o == 0 ? -px.c * px.a * px.b : 0.f;
      *g = -px.a * (t * px.b - (px.b > 9.9e5f ? 0.f : px.c * p * px.b * px.b * px.b)); break;
    }
  }
}

__global__ void loss_kernel(b2seg_loss_desc d) {
  pdl_prologue();
  const float* yp = reinterpret_cast<const float*>(d.y_pred);
  const float* yt = reinterpret_cast<const float*>(d.y_true);
  float* dl = reinterpret_cast<float*>(d.dlogits);
  const int co = d.cout, kind = d.kind, act = d.act;
  const bool chan_sum = kind == 1 || kind == 9 || kind == 13 || kind == 14;       // reduce_sum / one value per pixel, then mean over pixels
  const float scale = chan_sum ? 1.f / (float)d.n_pix : 1.f / ((float)d.n_pix * (float)co);
  const bool own_act = (kind == 0 && act == B2SEG_ACT_SIGMOID) || (kind == 1 && act == B2SEG_ACT_SOFTMAX) || (kind == 7 && act == B2SEG_ACT_SIGMOID);
  const bool prob = !own_act;
  float local = 0.f, m_sse = 0.f, m_sae = 0.f, m_bin = 0.f, m_cat = 0.f;
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < d.n_pix; pix += (long long)gridDim.x * blockDim.x) {
    const float* pp = yp + pix * co;
    const float* tt = yt + pix * co;
    LossPix px = {0.f, 0.f, 0.f, 0};
    if (kind == 1 && prob) {
      for (int o = 0; o < co; ++o) px.a += pp[o];
      for (int o = 0; o < co; ++o) { const float q = pp[o] / px.a; if (q > 1e-7f && q < 1.f - 1e-7f) px.c += tt[o]; }   // channels the clip leaves alone
    } else if (kind == 13) {
      float pos = 0.f, neg = -3.0e38f;
      for (int o = 0; o < co; ++o) { pos += tt[o] * pp[o]; const float v = (1.f - tt[o]) * pp[o]; if (v > neg) { neg = v; px.j = o; } }
      px.a = neg - pos + 1.f;
    } else if (kind == 14) {
      float st = 0.f, sp = 0.f;
      for (int o = 0; o < co; ++o) { st += tt[o] * tt[o]; sp += pp[o] * pp[o]; px.c += tt[o] * pp[o]; }
      px.a = rsqrtf(fmaxf(st, 1e-12f)); px.b = rsqrtf(fmaxf(sp, 1e-12f));
    }
    float gp = 0.f;     // softmax head under a generic loss: sum_k g_k p_k
    if (act == B2SEG_ACT_SOFTMAX && !(kind == 1 && own_act))
      for (int o = 0; o < co; ++o) { float l, g; loss_elem(kind, prob, pp[o], tt[o], px, o, &l, &g); gp += g * pp[o]; }
    int arg_p = 0, arg_t = 0;
    for (int o = 0; o < co; ++o) {
      const float p = pp[o], t = tt[o];
      float l, g;
      loss_elem(kind, prob, p, t, px, o, &l, &g);
      local += l * scale;
      float dz;
      if ((kind == 0 || kind == 1) && own_act) dz = p - t;                                 // cross-entropy from the cached logits
      else if (act == B2SEG_ACT_SIGMOID) dz = g * p * (1.f - p);
      else if (act == B2SEG_ACT_SOFTMAX) dz = p * (g - gp);
      else dz = g;
      if (dl) dl[pix * co + o] = d.weight * dz * scale;
      const float e = p - t;
      m_sse += e * e; m_sae += fabsf(e);
      m_bin += ((p > 0.5f) == (t > 0.5f)) ? 1.f : 0.f;
      if (p > pp[arg_p]) arg_p = o;
      if (t > tt[arg_t]) arg_t = o;
    }
    m_cat += arg_p == arg_t ? 1.f : 0.f;
  }
  // block reduction of {loss, sum sq err, sum abs err, binary hits, arg-max hits}
  __shared__ float sh[5][32];
  float v[5] = {local, m_sse, m_sae, m_bin, m_cat};
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    for (int off = 16; off > 0; off >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
    if ((threadIdx.x & 31) == 0) sh[q][threadIdx.x >> 5] = v[q];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      float t = threadIdx.x < (blockDim.x >> 5) ? sh[q][threadIdx.x] : 0.f;
      for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
      if (threadIdx.x == 0) {
        if (q == 0 && d.loss) atomicAdd(reinterpret_cast<float*>(d.loss), d.weight * t);
        if (d.metrics) atomicAdd(reinterpret_cast<float*>(d.metrics) + q, t);
      }
    }
  }
}
struct LossLaunch : PreparedOp {
  b2seg_loss_desc d;
  int launch(cudaStream_t s) override {
    int grid = grid_for(d.n_pix, 256);
    if (grid > 1024) grid = 1024;
    launch_k(loss_kernel, dim3(grid), dim3(256), 0, s, d);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
PreparedOp* prepare_loss(const b2seg_loss_desc* d) {
  if (d->kind < 0 || d->kind > 14) { set_error("loss: kind 0..14"); return nullptr; }
  if (d->act != B2SEG_ACT_NONE && d->act != B2SEG_ACT_SIGMOID && d->act != B2SEG_ACT_SOFTMAX) { set_error("loss: head activation %d", d->act); return nullptr; }
  if (((d->kind == 0 || d->kind == 7) && d->act == B2SEG_ACT_SOFTMAX) || (d->kind == 1 && d->act == B2SEG_ACT_SIGMOID)) {
    set_error("loss: a cross-entropy on the other activation (binary on softmax, categorical on sigmoid) is not lowered"); return nullptr;
  }
  auto* L = new LossLaunch(); L->d = *d; return L;
}

// ------------------------------------------------------------------------------------------ element-wise
struct EltK { int op, act; DView a, b, c, out; };
__global__ void eltwise_kernel(EltK k) {
  pdl_prologue();
  const int cv = k.out.C / 8;
  const unsigned total = (unsigned)k.out.N * k.out.H * k.out.W * cv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    unsigned t = i / cv;
    const int w = (int)(t % k.out.W); t /= k.out.W;
    const int h = (int)(t % k.out.H);
    const int n = (int)(t / k.out.H);
    float a[8], b[8], o[8];
    load8(vaddr(k.a, n, h, w, v * 8), a);
    if (k.op == 0 || k.op == 3) {
      load8(vaddr(k.b, n, h, w, v * 8), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = a[e] + b[e];
      if (k.op == 3) {
        load8(vaddr(k.c, n, h, w, v * 8), b);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] += b[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = act_fwd(o[e], k.act);
    } else if (k.op == 2) {
      load8(vaddr(k.b, n, h, w, v * 8), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = a[e] * (b[e] > 0.f ? 1.f : 0.3f);
    } else if (k.op == 4 || k.op == 5) {
      // operational layers (onn_layers.py:19): 4: a^p ; 5: p * a^(p-1) * b, p = k.act >= 2 (repeated products, exact in fp32)
      const int reps = k.op == 4 ? k.act - 1 : k.act - 2;
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = a[e];
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] *= a[e];
      }
      if (k.op == 5) {
        load8(vaddr(k.b, n, h, w, v * 8), b);
        const float p = (float)k.act;
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = p * o[e] * b[e];
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = a[e];
    }
    store8(vaddr(k.out, n, h, w, v * 8), o);
  }
}
struct EltLaunch : PreparedOp {
  EltK k;
  int launch(cudaStream_t s) override {
    const long long work = (long long)k.out.N * k.out.H * k.out.W * (k.out.C / 8);
    launch_k(eltwise_kernel, dim3(grid_for(work, 256)), dim3(256), 0, s, k);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
PreparedOp* prepare_eltwise(const b2seg_eltwise_desc* d) {
  if (d->out.C % 8) { set_error("eltwise: C %% 8"); return nullptr; }
  if (d->op < 0 || d->op > 5) { set_error("eltwise: unknown op %d", d->op); return nullptr; }
  if ((d->op == 4 || d->op == 5) && (d->act < 2 || d->act > 8)) { set_error("eltwise: power %d outside 2..8", d->act); return nullptr; }
  auto* L = new EltLaunch();
  L->k.op = d->op; L->k.act = d->act; L->k.a = dv(d->a); L->k.b = dv(d->b); L->k.c = dv(d->c); L->k.out = dv(d->out);
  return L;
}

// ------------------------------------------------------------------------------------------ column sum (bias grad)
__global__ void colsum_kernel(DView g, float* out, int cvb, int rp) {
  extern __shared__ float red[];
  const int cvec = g.C / 8;
  const int tcv = threadIdx.x % cvb, trow = threadIdx.x / cvb;
  const int v = blockIdx.x * cvb + tcv;
  const bool active = trow < rp && v < cvec;
  const unsigned n_pix = (unsigned)g.N * g.H * g.W;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (active)
    for (unsigned pix = blockIdx.y * rp + trow; pix < n_pix; pix += gridDim.y * rp) {
      unsigned t = pix;
      const int w = (int)(t % g.W); t /= g.W;
      const int h = (int)(t % g.H);
      const int n = (int)(t / g.H);
      float f[8];
      load8(vaddr(g, n, h, w, v * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
  float* mine = red + (size_t)threadIdx.x * 8;
#pragma unroll
  for (int e = 0; e < 8; ++e) mine[e] = active ? acc[e] : 0.f;
  __syncthreads();
  if (trow == 0 && v < cvec) {
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
    for (int r = 0; r < rp; ++r) {
      const float* q = red + (size_t)(r * cvb + tcv) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] += q[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(out + v * 8 + e, s[e]);
  }
}
struct ColsumLaunch : PreparedOp {
  b2seg_colsum_desc d;
  int launch(cudaStream_t s) override {
    const int cvec = d.g.C / 8;
    const int cvb = cvec < 256 ? cvec : 256, rp = 256 / cvb;
    const long long n_pix = (long long)d.g.N * d.g.H * d.g.W;
    long long gy = (n_pix + rp * 32 - 1) / (rp * 32);
    if (gy > 2048) gy = 2048;
    if (gy < 1) gy = 1;
    colsum_kernel<<<dim3((cvec + cvb - 1) / cvb, (unsigned)gy), 256, 256 * 8 * 4, s>>>(dv(d.g), reinterpret_cast<float*>(d.out), cvb, rp);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
PreparedOp* prepare_colsum(const b2seg_colsum_desc* d) {
  if (d->g.C % 8) { set_error("colsum: C %% 8"); return nullptr; }
  auto* L = new ColsumLaunch(); L->d = *d; return L;
}

// ------------------------------------------------------------------------------------------ memset
struct MemsetLaunch : PreparedOp {
  b2seg_memset_desc d;
  int launch(cudaStream_t s) override {
    B2_CUDA_OK(cudaMemsetAsync(reinterpret_cast<void*>(d.ptr), 0, (size_t)d.bytes, s));
    return 0;
  }
};
PreparedOp* prepare_memset(const b2seg_memset_desc* d) { auto* L = new MemsetLaunch(); L->d = *d; return L; }

}  // namespace b2
