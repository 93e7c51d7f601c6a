// HBM-streaming kernels for the decoder variants beyond plain UNet: up-sampling (nearest / bilinear) forward and adjoint,
// the attention gate's broadcast multiply and its channel-reduction backward, batch statistics of arbitrary tensors,
// the single-step ConvLSTM gate arithmetic, and max-pool backward for arbitrary windows.
// Same conventions as stream_kernels.cu: NHWC bf16 views, one thread = 8 channels (16 bytes), 32-bit index math.
#include "stream_common.cuh"

namespace b2 {

// ------------------------------------------------------------------------------------------ resize (UpSampling)
struct ResizeK { DView x, y, yfwd; int fh, fw, mode, act, c_valid; int n_vseg; int vseg_off[8], vseg_cnt[8]; };

// half-pixel bilinear source coordinates for output index o at integer scale f: src = (o + 0.5)/f - 0.5, edge-clamped taps
__device__ __forceinline__ void bil_taps(int o, int f, int in_size, int& i0, int& i1, float& lam) {
  const float src = (o + 0.5f) / (float)f - 0.5f;
  const float fl = floorf(src);
  lam = src - fl;
  i0 = (int)fl;
  i1 = i0 + 1;
  if (i0 < 0) i0 = 0;
  if (i1 < 0) i1 = 0;
  if (i0 > in_size - 1) i0 = in_size - 1;
  if (i1 > in_size - 1) i1 = in_size - 1;
}

__global__ void __launch_bounds__(256) resize_fwd_kernel(ResizeK k) {
  const int cv = k.y.C / 8;
  const unsigned total = (unsigned)k.y.N * k.y.H * k.y.W * cv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    unsigned t = i / cv;
    const int wo = (int)(t % k.y.W); t /= k.y.W;
    const int ho = (int)(t % k.y.H);
    const int n = (int)(t / k.y.H);
    float o[8];
    if (k.mode == 0) {
      load8(vaddr(k.x, n, ho / k.fh, wo / k.fw, v * 8), o);
    } else {
      int h0, h1, w0, w1;
      float lh, lw;
      bil_taps(ho, k.fh, k.x.H, h0, h1, lh);
      bil_taps(wo, k.fw, k.x.W, w0, w1, lw);
      float a[8], b[8], c[8], d[8];
      load8(vaddr(k.x, n, h0, w0, v * 8), a);
      load8(vaddr(k.x, n, h0, w1, v * 8), b);
      load8(vaddr(k.x, n, h1, w0, v * 8), c);
      load8(vaddr(k.x, n, h1, w1, v * 8), d);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float top = a[e] + (b[e] - a[e]) * lw, bot = c[e] + (d[e] - c[e]) * lw;
        o[e] = top + (bot - top) * lh;
      }
    }
    if (k.n_vseg > 0) {      // gapped channel layout: a lane is real only inside one of the segments
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = v * 8 + e;
        bool real = false;
        for (int sgi = 0; sgi < k.n_vseg; ++sgi) real = real || (c >= k.vseg_off[sgi] && c < k.vseg_off[sgi] + k.vseg_cnt[sgi]);
        o[e] = real ? act_fwd(o[e], k.act) : 0.f;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (k.c_valid && v * 8 + e >= k.c_valid) ? 0.f : act_fwd(o[e], k.act);
    }
    store8(vaddr(k.y, n, ho, wo, v * 8), o);
  }
}

// adjoint as a gather: every low-resolution pixel sums the weighted high-resolution gradients that referenced it
__global__ void __launch_bounds__(256) resize_bwd_kernel(ResizeK k) {
  const int cv = k.x.C / 8;
  const unsigned total = (unsigned)k.x.N * k.x.H * k.x.W * cv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    unsigned t = i / cv;
    const int wi = (int)(t % k.x.W); t /= k.x.W;
    const int hi = (int)(t % k.x.H);
    const int n = (int)(t / k.x.H);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    int oh_lo, oh_hi, ow_lo, ow_hi;
    if (k.mode == 0) {
      oh_lo = hi * k.fh; oh_hi = oh_lo + k.fh; ow_lo = wi * k.fw; ow_hi = ow_lo + k.fw;
    } else {
      oh_lo = max(0, k.fh * (hi - 1)); oh_hi = min(k.y.H, k.fh * (hi + 2));
      ow_lo = max(0, k.fw * (wi - 1)); ow_hi = min(k.y.W, k.fw * (wi + 2));
    }
    for (int oh = oh_lo; oh < oh_hi; ++oh) {
      float wh = 1.f;
      if (k.mode == 1) {
        int h0, h1; float lh;
        bil_taps(oh, k.fh, k.x.H, h0, h1, lh);
        wh = (h0 == hi ? 1.f - lh : 0.f) + (h1 == hi ? lh : 0.f);
        if (wh == 0.f) continue;
      }
      for (int ow = ow_lo; ow < ow_hi; ++ow) {
        float ww = 1.f;
        if (k.mode == 1) {
          int w0, w1; float lw;
          bil_taps(ow, k.fw, k.x.W, w0, w1, lw);
          ww = (w0 == wi ? 1.f - lw : 0.f) + (w1 == wi ? lw : 0.f);
          if (ww == 0.f) continue;
        }
        float g[8];
        load8(vaddr(k.y, n, oh, ow, v * 8), g);
        if (k.act != B2SEG_ACT_NONE) {
          float yv[8];
          load8(vaddr(k.yfwd, n, oh, ow, v * 8), yv);
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] *= act_bwd_from_y(yv[e], k.act);
        }
        const float wgt = wh * ww;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(wgt, g[e], acc[e]);
      }
    }
    store8(vaddr(k.x, n, hi, wi, v * 8), acc);
  }
}
struct ResizeLaunch : PreparedOp {
  ResizeK k;
  bool bwd;
  int launch(cudaStream_t s) override {
    const DView& g = bwd ? k.x : k.y;
    const long long work = (long long)g.N * g.H * g.W * (g.C / 8);
    int grid = grid_for(work, 256);
    const int cap = num_sms() * 32;
    if (grid > cap) grid = cap;
    if (bwd) resize_bwd_kernel<<<grid, 256, 0, s>>>(k);
    else resize_fwd_kernel<<<grid, 256, 0, s>>>(k);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
static PreparedOp* prep_resize(const b2seg_resize_desc* d, bool bwd) {
  if (d->x.C % 8 || d->y.C != d->x.C || d->fh < 1 || d->fw < 1 || d->y.H != d->x.H * d->fh || d->y.W != d->x.W * d->fw ||
      (d->act == B2SEG_ACT_SOFTMAX)) {
    set_error("resize: bad geometry (x %dx%dx%d, y %dx%dx%d, f %dx%d)", d->x.H, d->x.W, d->x.C, d->y.H, d->y.W, d->y.C, d->fh, d->fw);
    return nullptr;
  }
  if (d->n_vseg < 0 || d->n_vseg > 8) { set_error("resize: at most 8 valid-channel segments"); return nullptr; }
  auto* L = new ResizeLaunch();
  L->k = ResizeK{dv(d->x), dv(d->y), dv(d->yfwd), d->fh, d->fw, d->mode, d->act, d->c_valid, d->n_vseg, {0}, {0}};
  for (int i = 0; i < d->n_vseg; ++i) { L->k.vseg_off[i] = d->vseg_off[i]; L->k.vseg_cnt[i] = d->vseg_cnt[i]; }
  L->bwd = bwd;
  return L;
}
PreparedOp* prepare_resize_fwd(const b2seg_resize_desc* d) { return prep_resize(d, false); }
PreparedOp* prepare_resize_bwd(const b2seg_resize_desc* d) { return prep_resize(d, true); }

// ------------------------------------------------------------------------------------------ broadcast multiply
struct MulbcK { DView a, b, out, dout, da, db; };
__global__ void __launch_bounds__(256) mulbc_fwd_kernel(MulbcK k) {
  const int cv = k.out.C / 8;
  const unsigned total = (unsigned)k.out.N * k.out.H * k.out.W * cv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    unsigned t = i / cv;
    const int w = (int)(t % k.out.W); t /= k.out.W;
    const int h = (int)(t % k.out.H);
    const int n = (int)(t / k.out.H);
    float a[8];
    load8(vaddr(k.a, n, h, w, v * 8), a);
    const float m = __bfloat162float(__ldg(vaddr(k.b, n, h, w, 0)));
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] *= m;
    store8(vaddr(k.out, n, h, w, v * 8), a);
  }
}
// one group of G lanes per pixel: da = dout * b0 and db0 = sum_c dout * a (shuffle reduction over the channel vectors)
__global__ void __launch_bounds__(256) mulbc_bwd_kernel(MulbcK k, int G) {
  const unsigned n_pix = (unsigned)k.a.N * k.a.H * k.a.W;
  const int gl = threadIdx.x % G;
  const unsigned gid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
  const unsigned gstride = gridDim.x * blockDim.x / G;
  const int cvec = k.a.C / 8;
  const unsigned n_iter = (n_pix + gstride - 1) / gstride;
  for (unsigned it = 0; it < n_iter; ++it) {
    const unsigned pix = gid + it * gstride;
    const bool valid = pix < n_pix;
    unsigned t = valid ? pix : 0;
    const int w = (int)(t % k.a.W); t /= k.a.W;
    const int h = (int)(t % k.a.H);
    const int n = (int)(t / k.a.H);
    const float m = __bfloat162float(__ldg(vaddr(k.b, n, h, w, 0)));
    float acc = 0.f;
    for (int v = gl; v < cvec; v += G) {
      float a[8], g[8], o[8];
      load8(vaddr(k.a, n, h, w, v * 8), a);
      load8(vaddr(k.dout, n, h, w, v * 8), g);
#pragma unroll
      for (int e = 0; e < 8; ++e) { acc = fmaf(g[e], a[e], acc); o[e] = g[e] * m; }
      if (valid) store8(vaddr(k.da, n, h, w, v * 8), o);
    }
    for (int off = G / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off, 32);
    if (gl == 0 && valid) {
      float o[8];
      o[0] = acc;
#pragma unroll
      for (int e = 1; e < 8; ++e) o[e] = 0.f;
      for (int v = 0; v < k.db.C / 8; ++v) {
        store8(vaddr(k.db, n, h, w, v * 8), o);
        o[0] = 0.f;
      }
    }
  }
}
struct MulbcLaunch : PreparedOp {
  MulbcK k;
  bool bwd;
  int launch(cudaStream_t s) override {
    const int cap = num_sms() * 32;
    if (!bwd) {
      const long long work = (long long)k.out.N * k.out.H * k.out.W * (k.out.C / 8);
      int grid = grid_for(work, 256);
      if (grid > cap) grid = cap;
      mulbc_fwd_kernel<<<grid, 256, 0, s>>>(k);
    } else {
      int G = 8;
      while (G < 32 && G * 8 < k.a.C) G <<= 1;
      const long long n_pix = (long long)k.a.N * k.a.H * k.a.W;
      int grid = grid_for(n_pix * G, 256);
      if (grid > cap) grid = cap;
      mulbc_bwd_kernel<<<grid, 256, 0, s>>>(k, G);
    }
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
static PreparedOp* prep_mulbc(const b2seg_mulbc_desc* d, bool bwd) {
  if (d->a.C % 8 || d->b.C % 8) { set_error("mulbc: C %% 8"); return nullptr; }
  auto* L = new MulbcLaunch();
  L->k = MulbcK{dv(d->a), dv(d->b), dv(d->out), dv(d->dout), dv(d->da), dv(d->db)};
  L->bwd = bwd;
  return L;
}
PreparedOp* prepare_mulbc_fwd(const b2seg_mulbc_desc* d) { return prep_mulbc(d, false); }
PreparedOp* prepare_mulbc_bwd(const b2seg_mulbc_desc* d) { return prep_mulbc(d, true); }

// ------------------------------------------------------------------------------------------ column statistics
__global__ void __launch_bounds__(256) colstats_kernel(DView x, float* partials, int cvb, int rp) {
  extern __shared__ float red[];  // [256][16]
  const int cvec = x.C / 8;
  const int tcv = threadIdx.x % cvb, trow = threadIdx.x / cvb;
  const int v = blockIdx.x * cvb + tcv;
  const bool active = trow < rp && v < cvec;
  const unsigned n_pix = (unsigned)x.N * x.H * x.W;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
  if (active)
    for (unsigned pix = blockIdx.y * rp + trow; pix < n_pix; pix += gridDim.y * rp) {
      unsigned t = pix;
      const int w = (int)(t % x.W); t /= x.W;
      const int h = (int)(t % x.H);
      const int n = (int)(t / x.H);
      float f[8];
      load8(vaddr(x, n, h, w, v * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
    }
  float* mine = red + (size_t)threadIdx.x * 16;
#pragma unroll
  for (int e = 0; e < 8; ++e) { mine[e] = active ? s[e] : 0.f; mine[8 + e] = active ? q[e] : 0.f; }
  __syncthreads();
  if (trow == 0 && v < cvec) {
    float ss[8], qq[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { ss[e] = 0.f; qq[e] = 0.f; }
    for (int r = 0; r < rp; ++r) {
      const float* o = red + (size_t)(r * cvb + tcv) * 16;
#pragma unroll
      for (int e = 0; e < 8; ++e) { ss[e] += o[e]; qq[e] += o[8 + e]; }
    }
    float* pp = partials + (size_t)blockIdx.y * 2 * x.C;
#pragma unroll
    for (int e = 0; e < 8; ++e) { pp[v * 8 + e] = ss[e]; pp[x.C + v * 8 + e] = qq[e]; }
  }
}
struct ColstatsLaunch : PreparedOp {
  b2seg_colstats_desc d;
  int launch(cudaStream_t s) override {
    const int cvec = d.x.C / 8;
    const int cvb = cvec < 256 ? cvec : 256, rp = 256 / cvb;
    colstats_kernel<<<dim3((cvec + cvb - 1) / cvb, d.n_blocks), 256, 256 * 16 * 4, s>>>(dv(d.x), reinterpret_cast<float*>(d.partials), cvb, rp);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
PreparedOp* prepare_colstats(const b2seg_colstats_desc* d) {
  if (d->x.C % 8 || d->n_blocks < 1 || d->n_blocks > 65535) { set_error("colstats: bad C / n_blocks"); return nullptr; }
  auto* L = new ColstatsLaunch(); L->d = *d; return L;
}

// ------------------------------------------------------------------------------------------ ConvLSTM gates (T = 1)
// Keras-2 hard_sigmoid = clip(x * 0.2 + 0.5, 0, 1) as separate float32 ops; clip_by_value passes the gradient on [0, 1] inclusive.
// The multiply and the add must NOT contract into an FMA: gate pre-activations are stored in bf16, so x = -2.5 exactly is common
// (~3e-4 of all elements), and fma(0.2f, -2.5f, 0.5f) = -7.5e-9 < 0 puts it outside the clip range while mul-then-add gives exactly 0
// (inside): the teacher-forced gradient check measured 2-4 % rel-L2 on the input / output gates from this alone.
__device__ __forceinline__ float hard_sigmoid_lin(float x) { return __fadd_rn(__fmul_rn(0.2f, x), 0.5f); }
__device__ __forceinline__ float hard_sigmoid(float x) { return fminf(fmaxf(hard_sigmoid_lin(x), 0.f), 1.f); }
__device__ __forceinline__ float hard_sigmoid_grad(float x) { const float t = hard_sigmoid_lin(x); return (t >= 0.f && t <= 1.f) ? 0.2f : 0.f; }
struct LstmK { DView z, h, dh, dz; int F; };
template <bool BWD>
__global__ void __launch_bounds__(256) lstm_kernel(LstmK k) {
  const DView& ref = BWD ? k.dh : k.h;
  const int cv = k.F / 8;
  const unsigned total = (unsigned)ref.N * ref.H * ref.W * cv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    unsigned t = i / cv;
    const int w = (int)(t % ref.W); t /= ref.W;
    const int h = (int)(t % ref.H);
    const int n = (int)(t / ref.H);
    float zi[8], zg[8], zo[8];
    load8(vaddr(k.z, n, h, w, v * 8), zi);
    load8(vaddr(k.z, n, h, w, k.F + v * 8), zg);
    load8(vaddr(k.z, n, h, w, 2 * k.F + v * 8), zo);
    if (!BWD) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = hard_sigmoid(zo[e]) * tanhf(hard_sigmoid(zi[e]) * tanhf(zg[e]));
      store8(vaddr(k.h, n, h, w, v * 8), o);
    } else {
      float dh[8], di[8], dg[8], dO[8];
      load8(vaddr(k.dh, n, h, w, v * 8), dh);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float gi = hard_sigmoid(zi[e]), gg = tanhf(zg[e]), go = hard_sigmoid(zo[e]);
        const float c = gi * gg, tc = tanhf(c);
        dO[e] = dh[e] * tc * hard_sigmoid_grad(zo[e]);
        const float dc = dh[e] * go * (1.f - tc * tc);
        di[e] = dc * gg * hard_sigmoid_grad(zi[e]);
        dg[e] = dc * gi * (1.f - gg * gg);
      }
      store8(vaddr(k.dz, n, h, w, v * 8), di);
      store8(vaddr(k.dz, n, h, w, k.F + v * 8), dg);
      store8(vaddr(k.dz, n, h, w, 2 * k.F + v * 8), dO);
    }
  }
}
struct LstmLaunch : PreparedOp {
  LstmK k;
  bool bwd;
  int launch(cudaStream_t s) override {
    const DView& ref = bwd ? k.dh : k.h;
    const long long work = (long long)ref.N * ref.H * ref.W * (k.F / 8);
    int grid = grid_for(work, 256);
    const int cap = num_sms() * 32;
    if (grid > cap) grid = cap;
    if (bwd) lstm_kernel<true><<<grid, 256, 0, s>>>(k);
    else lstm_kernel<false><<<grid, 256, 0, s>>>(k);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
static PreparedOp* prep_lstm(const b2seg_lstm_desc* d, bool bwd) {
  if (d->F % 8 || d->F < 8) { set_error("lstm: F must be a multiple of 8"); return nullptr; }
  auto* L = new LstmLaunch();
  L->k = LstmK{dv(d->z), dv(d->h), dv(d->dh), dv(d->dz), d->F};
  L->bwd = bwd;
  return L;
}
PreparedOp* prepare_lstm_fwd(const b2seg_lstm_desc* d) { return prep_lstm(d, false); }
PreparedOp* prepare_lstm_bwd(const b2seg_lstm_desc* d) { return prep_lstm(d, true); }

// ------------------------------------------------------------------------------------------ max-pool backward (any window)
struct PoolBwdK { DView y, dp, dx; int ph, pw; };
__global__ void __launch_bounds__(256) pool_bwd_kernel(PoolBwdK k) {
  const int cv = k.y.C / 8;
  const int Ho = k.y.H / k.ph, Wo = k.y.W / k.pw;
  const unsigned total = (unsigned)k.y.N * Ho * Wo * cv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    unsigned t = i / cv;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float best[8];
    int arg[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; arg[e] = 0; }
    for (int a = 0; a < k.ph; ++a)
      for (int b = 0; b < k.pw; ++b) {
        float f[8];
        load8(vaddr(k.y, n, ho * k.ph + a, wo * k.pw + b, v * 8), f);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (f[e] > best[e]) { best[e] = f[e]; arg[e] = a * k.pw + b; }
      }
    float g[8];
    load8(vaddr(k.dp, n, ho, wo, v * 8), g);
    for (int a = 0; a < k.ph; ++a)
      for (int b = 0; b < k.pw; ++b) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (arg[e] == a * k.pw + b) ? g[e] : 0.f;
        store8(vaddr(k.dx, n, ho * k.ph + a, wo * k.pw + b, v * 8), o);
      }
  }
}
struct PoolBwdLaunch : PreparedOp {
  PoolBwdK k;
  int launch(cudaStream_t s) override {
    const long long work = (long long)k.y.N * (k.y.H / k.ph) * (k.y.W / k.pw) * (k.y.C / 8);
    int grid = grid_for(work, 256);
    const int cap = num_sms() * 32;
    if (grid > cap) grid = cap;
    pool_bwd_kernel<<<grid, 256, 0, s>>>(k);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
PreparedOp* prepare_pool_bwd(const b2seg_poolbwd_desc* d) {
  if (d->y.C % 8 || d->ph < 1 || d->pw < 1) { set_error("pool_bwd: bad args"); return nullptr; }
  auto* L = new PoolBwdLaunch();
  L->k = PoolBwdK{dv(d->y), dv(d->dp), dv(d->dx), d->ph, d->pw};
  return L;
}

// ------------------------------------------------------------------------------------------ activation as a model output
// Oper2D(output_nums, (1,1), activation=final_activation, q) (unet_variants.py:1107-1108): the logits are a bf16 tensor (the sum
// of q pointwise convolutions; also a softmax / sigmoid head with more than 8 classes, whose convolution runs on the tensor-core
// kernels); the loss kernels read fp32 [pixel][cout].  One thread per pixel, 16-byte loads / stores.
struct OutActK { DView x, dx; float* y; const float* dlogits; int cout, act; };
__global__ void __launch_bounds__(256) outact_fwd_kernel(OutActK k) {
  const unsigned total = (unsigned)k.x.N * k.x.H * k.x.W;
  const int nvec = (k.cout + 7) / 8;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned t = i;
    const int w = (int)(t % k.x.W); t /= k.x.W;
    const int h = (int)(t % k.x.H);
    const int n = (int)(t / k.x.H);
    float* yp = k.y + (size_t)i * k.cout;
    float z[8];
    if (k.act == B2SEG_ACT_SOFTMAX) {
      // three short passes over the pixel's logits (the re-reads hit L1): max, sum of exponentials, normalised write
      float m = -INFINITY, sum = 0.f;
      for (int v = 0; v < nvec; ++v) {
        load8(vaddr(k.x, n, h, w, v * 8), z);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (v * 8 + e < k.cout) m = fmaxf(m, z[e]);
      }
      for (int v = 0; v < nvec; ++v) {
        load8(vaddr(k.x, n, h, w, v * 8), z);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (v * 8 + e < k.cout) sum += __expf(z[e] - m);
      }
      const float inv = 1.f / sum;
      for (int v = 0; v < nvec; ++v) {
        load8(vaddr(k.x, n, h, w, v * 8), z);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (v * 8 + e < k.cout) yp[v * 8 + e] = __expf(z[e] - m) * inv;
      }
    } else {
      for (int v = 0; v < nvec; ++v) {
        load8(vaddr(k.x, n, h, w, v * 8), z);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (v * 8 + e < k.cout) yp[v * 8 + e] = act_fwd(z[e], k.act);
      }
    }
  }
}
__global__ void __launch_bounds__(256) outact_bwd_kernel(OutActK k) {
  const unsigned total = (unsigned)k.dx.N * k.dx.H * k.dx.W;
  const int nvec = k.dx.C / 8;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned t = i;
    const int w = (int)(t % k.dx.W); t /= k.dx.W;
    const int h = (int)(t % k.dx.H);
    const int n = (int)(t / k.dx.H);
    const float* gp = k.dlogits + (size_t)i * k.cout;
    for (int v = 0; v < nvec; ++v) {
      float g[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = v * 8 + e < k.cout ? gp[v * 8 + e] : 0.f;
      store8(vaddr(k.dx, n, h, w, v * 8), g);
    }
  }
}
struct OutActLaunch : PreparedOp {
  OutActK k;
  bool bwd;
  int launch(cudaStream_t s) override {
    const DView& v = bwd ? k.dx : k.x;
    int grid = grid_for((long long)v.N * v.H * v.W, 256);
    const int cap = num_sms() * 32;
    if (grid > cap) grid = cap;
    if (bwd) outact_bwd_kernel<<<grid, 256, 0, s>>>(k);
    else outact_fwd_kernel<<<grid, 256, 0, s>>>(k);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
static PreparedOp* prep_outact(const b2seg_outact_desc* d, bool bwd) {
  const b2seg_view& v = bwd ? d->dx : d->x;
  if (v.C < 8 || v.C % 8 || d->cout < 1 || d->cout > v.C) { set_error("outact: needs a view of C %% 8 == 0 channels and 1 <= cout <= C"); return nullptr; }
  if (d->act != B2SEG_ACT_NONE && d->act != B2SEG_ACT_SIGMOID && d->act != B2SEG_ACT_SOFTMAX) { set_error("outact: activation %d", d->act); return nullptr; }
  if ((long long)v.N * v.H * v.W >= (1ll << 31)) { set_error("outact: too many pixels"); return nullptr; }
  if (bwd ? !d->dlogits : !d->y) { set_error("outact: null fp32 buffer"); return nullptr; }
  auto* L = new OutActLaunch();
  L->k = OutActK{dv(d->x), dv(d->dx), reinterpret_cast<float*>(d->y), reinterpret_cast<const float*>(d->dlogits), d->cout, d->act};
  L->bwd = bwd;
  return L;
}
PreparedOp* prepare_outact_fwd(const b2seg_outact_desc* d) { return prep_outact(d, false); }
PreparedOp* prepare_outact_bwd(const b2seg_outact_desc* d) { return prep_outact(d, true); }

// ------------------------------------------------------------------------------------------ deep-supervision target pyramid
// level-k target = MaxPooling2D(2^k) of the mask (helper_functions.py:359-380) / window mean in 1D (notebook cell 31), fp32.
struct TPoolK { const float* src; float* dst; int N, H, W, C, ph, pw, mode; };
__global__ void __launch_bounds__(256) target_pool_kernel(TPoolK k) {
  const int Ho = k.H / k.ph, Wo = k.W / k.pw;
  const unsigned total = (unsigned)k.N * Ho * Wo * k.C;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = (int)(i % k.C);
    unsigned t = i / k.C;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float best = -INFINITY, sum = 0.f;
    for (int a = 0; a < k.ph; ++a)
      for (int b = 0; b < k.pw; ++b) {
        const float v = __ldg(k.src + (((size_t)n * k.H + ho * k.ph + a) * k.W + wo * k.pw + b) * k.C + c);
        best = fmaxf(best, v);
        sum += v;
      }
    k.dst[i] = k.mode == 0 ? best : sum / (float)(k.ph * k.pw);
  }
}
struct TPoolLaunch : PreparedOp {
  TPoolK k;
  int launch(cudaStream_t s) override {
    int grid = grid_for((long long)k.N * (k.H / k.ph) * (k.W / k.pw) * k.C, 256);
    const int cap = num_sms() * 32;
    if (grid > cap) grid = cap;
    target_pool_kernel<<<grid, 256, 0, s>>>(k);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
PreparedOp* prepare_target_pool(const b2seg_tpool_desc* d) {
  if (!d->src || !d->dst || d->N < 1 || d->C < 1 || d->ph < 1 || d->pw < 1 || d->H < d->ph || d->W < d->pw || d->H % d->ph || d->W % d->pw ||
      (d->mode != 0 && d->mode != 1)) {
    set_error("target_pool: bad arguments");
    return nullptr;
  }
  if ((long long)d->N * (d->H / d->ph) * (d->W / d->pw) * d->C >= (1ll << 31)) { set_error("target_pool: too many elements"); return nullptr; }
  auto* L = new TPoolLaunch();
  L->k = TPoolK{reinterpret_cast<const float*>(d->src), reinterpret_cast<float*>(d->dst), d->N, d->H, d->W, d->C, d->ph, d->pw, d->mode};
  return L;
}

// ------------------------------------------------------------------------------------------ BatchNorm folded into the conv weights (inference)
// Inference normalises with the moving statistics, which are known before the convolution runs:
//   act(BN(conv(x, W) + b)) = act(conv(x, W * s) + (b * s + t)),  s = gamma / sqrt(moving_var + eps),  t = beta - moving_mean * s
// One thread per weight element writes the bf16 folded kernel row by row ([cout_p][row] layout, one scale per row); the first `cout_p`
// threads also write the folded bias.  Runs once per weight change, not per predict call.
__global__ void __launch_bounds__(256) fold_bn_kernel(b2seg_fold_desc d) {
  const float* w = reinterpret_cast<const float*>(d.w);
  const float* gamma = reinterpret_cast<const float*>(d.gamma);
  const float* beta = reinterpret_cast<const float*>(d.beta);
  const float* mm = reinterpret_cast<const float*>(d.moving_mean);
  const float* mv = reinterpret_cast<const float*>(d.moving_var);
  __nv_bfloat16* wo = reinterpret_cast<__nv_bfloat16*>(d.w_folded);
  const long long total = (long long)d.cout_p * d.row;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / d.row);
    const float s = gamma[co] * rsqrtf(mv[co] + d.eps);
    wo[i] = __float2bfloat16_rn(w[i] * s);
    if (i < d.cout_p) {
      const int c = (int)i;
      const float sc = gamma[c] * rsqrtf(mv[c] + d.eps);
      const float b = d.bias ? reinterpret_cast<const float*>(d.bias)[c] : 0.f;
      reinterpret_cast<float*>(d.bias_folded)[c] = b * sc + (beta[c] - mm[c] * sc);
    }
  }
}
struct FoldLaunch : PreparedOp {
  b2seg_fold_desc d;
  int launch(cudaStream_t s) override {
    int grid = grid_for((long long)d.cout_p * d.row, 256);
    const int cap = num_sms() * 16;
    if (grid > cap) grid = cap;
    fold_bn_kernel<<<grid, 256, 0, s>>>(d);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
PreparedOp* prepare_fold_bn(const b2seg_fold_desc* d) {
  if (!d->w || !d->gamma || !d->beta || !d->moving_mean || !d->moving_var || !d->w_folded || !d->bias_folded || d->cout_p < 1 || d->row < 1) {
    set_error("fold_bn: bad arguments (cout_p %d, row %d)", d->cout_p, d->row);
    return nullptr;
  }
  auto* L = new FoldLaunch(); L->d = *d; return L;
}

}  // namespace b2
