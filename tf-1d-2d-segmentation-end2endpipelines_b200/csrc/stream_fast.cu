// Instruction-lean fast paths of the two HBM-streaming kernels that dominate a training step: BN apply (+activation
// +2x2 / 1x2 max-pool) and BN / activation / max-pool backward.
//
// ncu on the generic kernels (profiles/r1_prof_stream.txt): 73 % issue-slot utilisation at 56 % of HBM bandwidth and
// ~180 thread instructions per 8-channel vector — they were bound by instruction issue, not by memory: per-element
// runtime activation switches and channel-mask tests, 64-bit strided address arithmetic and div/mod index decoding for
// every vector.  Here the work is walked row by row (one magic-number division per image row), offsets are 32-bit
// element offsets, the activation / window / source configuration are template parameters, and bf16 <-> fp32 uses the
// shift/mask form, which brings a vector down to ~40-60 instructions so the kernels run at memory speed.
//
// Eligibility is checked on the host (prepare_*_fast return nullptr when a view needs 64-bit offsets, the channel count
// needs masking, the window is unusual, ...) and the generic kernels of stream_kernels.cu remain the fallback.
#include <algorithm>

#include "stream_common.cuh"

namespace b2 {

struct FV { unsigned long long ptr; unsigned sn, sh, sw; };   // strides in elements, all offsets < 2^31

static bool fv_make(const b2seg_view& v, FV* o) {
  if (v.ptr == 0 || (v.ptr & 15) || v.sn < 0 || v.sh < 0 || v.sw < 0 || (v.sn % 8) || (v.sh % 8) || (v.sw % 8)) return false;
  const long long span = (long long)(v.N - 1) * v.sn + (long long)(v.H - 1) * v.sh + (long long)(v.W - 1) * v.sw + v.C;
  if (span >= (1ll << 31)) return false;
  o->ptr = v.ptr; o->sn = (unsigned)v.sn; o->sh = (unsigned)v.sh; o->sw = (unsigned)v.sw;
  return true;
}

__device__ __forceinline__ uint4 ld16(const FV& v, unsigned off) {
  return __ldg(reinterpret_cast<const uint4*>(v.ptr + (unsigned long long)off * 2ull));
}
__device__ __forceinline__ void st16(const FV& v, unsigned off, uint4 u) {
  *reinterpret_cast<uint4*>(v.ptr + (unsigned long long)off * 2ull) = u;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
template <int ACT>
__device__ __forceinline__ float act_f(float x) {
  if (ACT == B2SEG_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == B2SEG_ACT_LEAKY) return fmaxf(x, 0.3f * x);
  if (ACT == B2SEG_ACT_SIGMOID) return 1.f / (1.f + __expf(-x));
  return x;
}

// ---------------------------------------------------------------------------------------------------- BN apply
struct BnActF {
  FV x, out0, out1, pooled;
  const float* scale; const float* shift;
  int n_out, has_pool;
  int C, Ho, Wo, rows;        // rows = N * Ho (window rows)
  FastDiv fd_ho;
  int cvb, rp;
};

// One thread owns 8 channels; a block covers cvb channel vectors x rp window columns; blocks walk window rows.
template <int PH, int PW, int ACT, int U>
__global__ void __launch_bounds__(256, PH * PW * U > 4 ? 2 : 4) bn_act_fast_kernel(const BnActF k) {
  pdl_prologue();
  constexpr int WIN = PH * PW;
  const int tcv = threadIdx.x % k.cvb, trow = threadIdx.x / k.cvb;
  const int v = blockIdx.x * k.cvb + tcv;
  if (v * 8 >= k.C || trow >= k.rp) return;
  float sc[8], sf[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = k.scale ? __ldg(k.scale + v * 8 + e) : 1.f;
    sf[e] = k.shift ? __ldg(k.shift + v * 8 + e) : 0.f;
  }
  const unsigned step = (unsigned)k.rp;
  for (int rr = blockIdx.y; rr < k.rows; rr += gridDim.y) {
    // rows are walked from the END of the tensor: the convolution that produced x wrote its last tiles last, so they are
    // still in the 126 MB L2, and the head of the output is what the next (forward-walking) convolution reads first
    const int r = k.rows - 1 - rr;
    const unsigned n = fast_div((unsigned)r, k.fd_ho), ho = (unsigned)r - n * (unsigned)k.Ho;
    const unsigned xrow = n * k.x.sn + ho * PH * k.x.sh + v * 8;
    const unsigned o0row = n * k.out0.sn + ho * PH * k.out0.sh + v * 8;
    const unsigned o1row = n * k.out1.sn + ho * PH * k.out1.sh + v * 8;
    const unsigned prow = n * k.pooled.sn + ho * k.pooled.sh + v * 8;
    for (unsigned wo0 = trow; wo0 < (unsigned)k.Wo; wo0 += step * U) {
      uint4 raw[U][WIN];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned wo = wo0 + u * step;
        if (wo < (unsigned)k.Wo) {
#pragma unroll
          for (int q = 0; q < WIN; ++q) raw[u][q] = ld16(k.x, xrow + (q / PW) * k.x.sh + (wo * PW + q % PW) * k.x.sw);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned wo = wo0 + u * step;
        if (wo >= (unsigned)k.Wo) break;
        float mx[8];
#pragma unroll
        for (int q = 0; q < WIN; ++q) {
          float f[8];
          unpack8(raw[u][q], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            f[e] = act_f<ACT>(fmaf(f[e], sc[e], sf[e]));
            mx[e] = q == 0 ? f[e] : fmaxf(mx[e], f[e]);
          }
          const uint4 o = pack8(f);
          if (k.n_out > 0) st16(k.out0, o0row + (q / PW) * k.out0.sh + (wo * PW + q % PW) * k.out0.sw, o);      // (0: a stand-alone MaxPooling, only the pooled tensor is written)
          if (k.n_out > 1) st16(k.out1, o1row + (q / PW) * k.out1.sh + (wo * PW + q % PW) * k.out1.sw, o);
        }
        if (WIN > 1 && k.has_pool) st16(k.pooled, prow + wo * k.pooled.sw, pack8(mx));
      }
    }
  }
}

struct BnActFastLaunch : PreparedOp {
  BnActF k;
  int ph, pw, act;
  dim3 grid;
  template <int PH, int PW, int U>
  void go(cudaStream_t s) {
    switch (act) {
      case B2SEG_ACT_RELU: launch_k(bn_act_fast_kernel<PH, PW, B2SEG_ACT_RELU, U>, grid, dim3(256), 0, s, k); break;
      case B2SEG_ACT_LEAKY: launch_k(bn_act_fast_kernel<PH, PW, B2SEG_ACT_LEAKY, U>, grid, dim3(256), 0, s, k); break;
      case B2SEG_ACT_SIGMOID: launch_k(bn_act_fast_kernel<PH, PW, B2SEG_ACT_SIGMOID, U>, grid, dim3(256), 0, s, k); break;
      default: launch_k(bn_act_fast_kernel<PH, PW, B2SEG_ACT_NONE, U>, grid, dim3(256), 0, s, k); break;
    }
  }
  int launch(cudaStream_t s) override {
    if (ph == 2 && pw == 2) go<2, 2, 2>(s);
    else if (ph == 1 && pw == 2) go<1, 2, 2>(s);
    else go<1, 1, 4>(s);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};

// y = act(x * scale + shift + add) with the per-channel sums of y and y^2 added into caller-zeroed accumulators: MultiResBlock /
// ResPath glue in one pass (2DCNN/models/unet_variants.py:96-99, 108-112).  `add` fuses Add([shortcut, BatchNormalization(concat)]) +
// Activation('relu') into the BatchNorm apply; the sums are the batch statistics of the BatchNormalization that follows, so that
// layer needs no statistics pass of its own.  The sums are taken over the bf16 values that are stored (what its apply will read).
// The sums are reproducible: the blocks add their partial sums into DOUBLE accumulators (red.add.f64: the order of the adds moves the
// total by 2^-53 at most), the last block to arrive (ticket counter) rounds the totals to fp32 once, adds them into the caller's
// accumulators and clears the scratch for the next launch.  (fp32 red.add from every block made a 48-BatchNorm MultiRes net at
// random init differ by 0.6 % in its output and 10-35 % in its first layers' gradients from one run to the next:
// profiles/r2_diag_multires_run_to_run.txt; with this scheme the forward pass of two runs is bit-identical.)
struct BnAct2F {
  FV x, add, out0, out1;
  const float* scale; const float* shift;
  float* stats;
  double* scratch;         // [2][scr_pitch] double accumulators (owned by the launch object), zero between launches
  unsigned* tickets;       // [gridDim.x], zero between launches
  int scr_pitch;
  int stats_pitch;
  int n_out;
  int C, W, rows;
  int cvb, rp;
};
template <int ACT, bool ADD, bool STATS, int U>
__global__ void __launch_bounds__(256, 3) bn_act2_fast_kernel(const BnAct2F k) {
  pdl_prologue();
  extern __shared__ float red[];   // STATS: [256][16]
  const int tcv = threadIdx.x % k.cvb, trow = threadIdx.x / k.cvb;
  const int v = blockIdx.x * k.cvb + tcv;
  const bool active = v * 8 < k.C && trow < k.rp;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
  if (active) {
    float sc[8], sf[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sc[e] = k.scale ? __ldg(k.scale + v * 8 + e) : 1.f;
      sf[e] = k.shift ? __ldg(k.shift + v * 8 + e) : 0.f;
    }
    const unsigned step = (unsigned)k.rp;
    for (int r = blockIdx.y; r < k.rows; r += gridDim.y) {      // rows = N * H (every view is addressed by (row, w): sn == H * sh is checked on the host)
      const unsigned xrow = (unsigned)r * k.x.sh + v * 8, arow = (unsigned)r * k.add.sh + v * 8;
      const unsigned o0row = (unsigned)r * k.out0.sh + v * 8, o1row = (unsigned)r * k.out1.sh + v * 8;
      for (unsigned w0 = trow; w0 < (unsigned)k.W; w0 += step * U) {
        uint4 xr[U], ar[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned w = w0 + u * step < (unsigned)k.W ? w0 + u * step : (unsigned)k.W - 1;
          xr[u] = ld16(k.x, xrow + w * k.x.sw);
          if (ADD) ar[u] = ld16(k.add, arow + w * k.add.sw);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned w = w0 + u * step;
          if (w >= (unsigned)k.W) break;
          float f[8], a[8];
          unpack8(xr[u], f);
          if (ADD) unpack8(ar[u], a);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = act_f<ACT>(fmaf(f[e], sc[e], sf[e]) + (ADD ? a[e] : 0.f));
          const uint4 o = pack8(f);
          st16(k.out0, o0row + w * k.out0.sw, o);
          if (k.n_out > 1) st16(k.out1, o1row + w * k.out1.sw, o);
          if (STATS) {
            float g[8];
            unpack8(o, g);
#pragma unroll
            for (int e = 0; e < 8; ++e) { s[e] += g[e]; q[e] = fmaf(g[e], g[e], q[e]); }
          }
        }
      }
    }
  }
  if (STATS) {
    float* mine = red + (size_t)threadIdx.x * 16;
#pragma unroll
    for (int e = 0; e < 8; ++e) { mine[e] = s[e]; mine[8 + e] = q[e]; }
    __syncthreads();
    if (trow == 0 && v * 8 < k.C) {
      float ts[8], tq[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { ts[e] = 0.f; tq[e] = 0.f; }
      for (int r = 0; r < k.rp; ++r) {
        const float* o = red + (size_t)(r * k.cvb + tcv) * 16;
#pragma unroll
        for (int e = 0; e < 8; ++e) { ts[e] += o[e]; tq[e] += o[8 + e]; }
      }
      double* acc = k.scratch + v * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) { atomicAdd(acc + e, (double)ts[e]); atomicAdd(acc + k.scr_pitch + e, (double)tq[e]); }
    }
    __threadfence();
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(k.tickets + blockIdx.x, 1u) == gridDim.y - 1;
    __syncthreads();
    if (s_last) {
      __threadfence();
      const int nch = k.cvb * 8;      // channels of this column of blocks
      for (int i = threadIdx.x; i < 2 * nch; i += 256) {
        const int which = i >= nch, c = blockIdx.x * nch + (which ? i - nch : i);
        if (c >= k.C) continue;
        double* col = k.scratch + (size_t)which * k.scr_pitch + c;
        const double total = __ldcg(col);
        __stcg(col, 0.0);
        atomicAdd(k.stats + which * k.stats_pitch + c, (float)total);
      }
      if (threadIdx.x == 0) k.tickets[blockIdx.x] = 0u;
    }
  }
}
struct BnAct2Launch : PreparedOp {
  BnAct2F k;
  int act;
  bool add, stats;
  dim3 grid;
  void* d_scratch = nullptr;
  ~BnAct2Launch() override { if (d_scratch) cudaFree(d_scratch); }
  template <int ACT>
  void go(cudaStream_t s) {
    const int smem = stats ? 256 * 16 * 4 : 0;
    if (add && stats) launch_k(bn_act2_fast_kernel<ACT, true, true, 2>, grid, dim3(256), smem, s, k);
    else if (add) launch_k(bn_act2_fast_kernel<ACT, true, false, 2>, grid, dim3(256), smem, s, k);
    else launch_k(bn_act2_fast_kernel<ACT, false, true, 4>, grid, dim3(256), smem, s, k);
  }
  int launch(cudaStream_t s) override {
    switch (act) {
      case B2SEG_ACT_RELU: go<B2SEG_ACT_RELU>(s); break;
      case B2SEG_ACT_LEAKY: go<B2SEG_ACT_LEAKY>(s); break;
      default: go<B2SEG_ACT_NONE>(s); break;
    }
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};

static void row_grid(int C, int rows, int Wo, int* cvb, int* rp, dim3* grid) {
  const int cv = C / 8;
  *cvb = cv < 256 ? cv : 256;
  // threads along the window-column axis: no more than the row has columns (power of two so that cvb * rp <= 256)
  int r = 256 / *cvb;
  while (r > 1 && r / 2 >= Wo) r /= 2;
  *rp = r;
  const int gx = (cv + *cvb - 1) / *cvb;
  long long gy = (long long)num_sms() * 16 / gx;
  if (gy > rows) gy = rows;
  if (gy < 1) gy = 1;
  *grid = dim3(gx, (unsigned)gy);
}

PreparedOp* prepare_bn_act_fast(const b2seg_bn_act_desc* d) {
  static const bool disabled = getenv("B2SEG_NO_FAST_STREAM") != nullptr;
  if (disabled || d->c_valid != 0 || d->x.C % 8 || d->n_out < 0 || d->n_out > 2) return nullptr;
  const int ph = d->pool_h > 1 ? d->pool_h : 1, pw = d->pool_w > 1 ? d->pool_w : 1;
  if (d->n_out == 0 && (ph * pw == 1 || d->add.ptr || d->out_stats)) return nullptr;      // no un-pooled output: only as a stand-alone pooling
  if (d->add.ptr || d->out_stats) {
    // fused add / output statistics (MultiResBlock, ResPath): un-pooled, ReLU / LeakyReLU / none, views addressable by (row, w)
    if (ph * pw > 1 || (d->act != B2SEG_ACT_NONE && d->act != B2SEG_ACT_RELU && d->act != B2SEG_ACT_LEAKY)) return nullptr;
    auto* L2 = new BnAct2Launch();
    BnAct2F& k2 = L2->k;
    memset(&k2, 0, sizeof(k2));
    auto rowwise = [&](const b2seg_view& v) { return v.N == 1 || v.sn == (long long)v.H * v.sh; };
    bool ok2 = fv_make(d->x, &k2.x) && fv_make(d->out[0], &k2.out0) && rowwise(d->x) && rowwise(d->out[0]);
    k2.out1 = k2.out0; k2.add = k2.x;
    if (ok2 && d->n_out > 1) ok2 = fv_make(d->out[1], &k2.out1) && rowwise(d->out[1]);
    if (ok2 && d->add.ptr) ok2 = fv_make(d->add, &k2.add) && rowwise(d->add) && d->add.C == d->x.C && d->add.H == d->x.H && d->add.W == d->x.W && d->add.N == d->x.N;
    if (!ok2) { delete L2; return nullptr; }
    k2.scale = reinterpret_cast<const float*>(d->scale); k2.shift = reinterpret_cast<const float*>(d->shift);
    k2.stats = reinterpret_cast<float*>(d->out_stats);
    k2.stats_pitch = d->out_stats_pitch > 0 ? d->out_stats_pitch : d->x.C;
    k2.n_out = d->n_out; k2.C = d->x.C; k2.W = d->x.W; k2.rows = d->x.N * d->x.H;
    L2->act = d->act; L2->add = d->add.ptr != 0; L2->stats = d->out_stats != 0;
    row_grid(k2.C, k2.rows, k2.W, &k2.cvb, &k2.rp, &L2->grid);
    if (L2->stats) {   // one resident wave: the kernel ends with 16 * cvb atomics per block
      const unsigned wave = (unsigned)std::max(1, num_sms() * 3 / (int)L2->grid.x);
      if (L2->grid.y > wave) L2->grid.y = wave;
      k2.scr_pitch = (int)L2->grid.x * k2.cvb * 8;
      const size_t scr = (size_t)2 * k2.scr_pitch * sizeof(double), tick = (size_t)L2->grid.x * sizeof(unsigned);
      if (cudaMalloc(&L2->d_scratch, scr + tick) != cudaSuccess || cudaMemset(L2->d_scratch, 0, scr + tick) != cudaSuccess) {
        set_error("bn_act: cannot allocate %zu bytes of statistics scratch", scr + tick);
        delete L2;
        return nullptr;      // (the caller falls back to the generic kernel, which refuses out_stats: the plan fails loudly)
      }
      k2.scratch = reinterpret_cast<double*>(L2->d_scratch);
      k2.tickets = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(L2->d_scratch) + scr);
    }
    return L2;
  }
  if (!((ph == 1 && pw == 1) || (ph == 2 && pw == 2) || (ph == 1 && pw == 2))) return nullptr;
  if (d->x.H % ph || d->x.W % pw) return nullptr;
  if (d->act != B2SEG_ACT_NONE && d->act != B2SEG_ACT_RELU && d->act != B2SEG_ACT_LEAKY && d->act != B2SEG_ACT_SIGMOID) return nullptr;
  auto* L = new BnActFastLaunch();
  BnActF& k = L->k;
  memset(&k, 0, sizeof(k));
  bool ok = fv_make(d->x, &k.x);
  k.out0 = k.x;
  if (ok && d->n_out > 0) ok = fv_make(d->out[0], &k.out0);
  k.out1 = k.out0;
  if (ok && d->n_out > 1) ok = fv_make(d->out[1], &k.out1);
  k.pooled = k.out0;
  k.has_pool = (ph * pw > 1) ? 1 : 0;
  if (ok && k.has_pool) ok = fv_make(d->pooled, &k.pooled);
  if (!ok) { delete L; return nullptr; }
  k.scale = reinterpret_cast<const float*>(d->scale);
  k.shift = reinterpret_cast<const float*>(d->shift);
  k.n_out = d->n_out;
  k.C = d->x.C; k.Ho = d->x.H / ph; k.Wo = d->x.W / pw; k.rows = d->x.N * k.Ho;
  k.fd_ho = make_fastdiv((uint32_t)k.Ho);
  L->ph = ph; L->pw = pw; L->act = d->act;
  row_grid(k.C, k.rows, k.Wo, &k.cvb, &k.rp, &L->grid);
  return L;
}

// ---------------------------------------------------------------------------------------------------- BN backward
struct BnBwdF {
  FV x, dx;
  FV src[B2SEG_MAX_GRADSRC];
  int kind[B2SEG_MAX_GRADSRC];   // 0 direct (same pixel grid as x), 1 pooled (one value per window)
  int n_src;
  const float* scale; const float* shift; const float* mean; const float* rstd;
  const float* dgamma; const float* dbeta;
  float inv_count;
  float* acc_dgamma; float* acc_dbeta;   // PASS 0: red.add targets (the dgamma / dbeta slots, zero before the launch)
  int C, Ho, Wo, rows;
  FastDiv fd_ho;
  int cvb, rp;
  // pointwise-head source (b2seg_gradsrc kind 2): g = dlogits . head_w^T formed on the fly; PASS 0 also accumulates the head's dW, db
  const float* dlogits; const float* head_w;
  float* head_dw; float* head_db;
  int x_relu_mask;   // PASS 1: dx *= (x > 0)
};

// Same arithmetic as bn_bwd_lean_kernel (stream_kernels.cu): mask from the sign of t = x*scale + shift, pooled gradients
// go to the first arg-max of t in the window.  PASS 0 accumulates dbeta = sum g and dgamma = sum g*xhat (xhat = (x-mean)*rstd,
// formed per element so that no cancellation is left for a finalize step) and adds the block's sums straight into the
// dgamma / dbeta slots with red.global.add -- there is no partials buffer and no finalize launch.
// PASS 1 reads them back and writes dx = A*g + B*x + D.
template <int PASS, int PH, int PW, int ACT, int U>
__global__ void __launch_bounds__(256, PH * PW * U >= 4 ? 2 : 3) bn_bwd_fast_kernel(const BnBwdF k) {
  pdl_prologue();
  constexpr int WIN = PH * PW;
  extern __shared__ float red[];   // PASS 0: [256][16]
  const int tcv = threadIdx.x % k.cvb, trow = threadIdx.x / k.cvb;
  const int v = blockIdx.x * k.cvb + tcv;
  const bool active = v * 8 < k.C && trow < k.rp;
  const bool has_bn = k.scale != nullptr;
  float sc[8], sf[8], cB[8], cD[8], acc_b[8], acc_g[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { acc_b[e] = 0.f; acc_g[e] = 0.f; sc[e] = 1.f; sf[e] = 0.f; cB[e] = 0.f; cD[e] = 0.f; }
  if (active) {
    if (has_bn) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = v * 8 + e;
        sc[e] = __ldg(k.scale + c);
        sf[e] = __ldg(k.shift + c);
        const float rs = __ldg(k.rstd + c), mu = __ldg(k.mean + c);
        if (PASS == 1) {
          const float cb = __ldg(k.dbeta + c) * k.inv_count, cg = __ldg(k.dgamma + c) * k.inv_count;
          cB[e] = -sc[e] * cg * rs;
          cD[e] = sc[e] * (cg * rs * mu - cb);
        } else {
          cB[e] = rs;          // PASS 0: xhat = x*cB + cD
          cD[e] = -mu * rs;
        }
      }
    }
    const unsigned step = (unsigned)k.rp;
    const int n_src = k.n_src;
    for (int rr = blockIdx.y; rr < k.rows; rr += gridDim.y) {
      // PASS 0 walks the rows backwards (the tail of the gradient its producer just wrote is still in L2) and leaves the head
      // of x / g in L2 for PASS 1, which walks forwards
      const int r = PASS == 0 ? k.rows - 1 - rr : rr;
      const unsigned n = fast_div((unsigned)r, k.fd_ho), ho = (unsigned)r - n * (unsigned)k.Ho;
      const unsigned xrow = n * k.x.sn + ho * PH * k.x.sh + v * 8;
      const unsigned drow = n * k.dx.sn + ho * PH * k.dx.sh + v * 8;
      for (unsigned wo0 = trow; wo0 < (unsigned)k.Wo; wo0 += step * U) {
        uint4 xr[U][WIN];
        float g[U][WIN][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned wo = wo0 + u * step < (unsigned)k.Wo ? wo0 + u * step : (unsigned)k.Wo - 1;   // clamp: loads stay in range
#pragma unroll
          for (int q = 0; q < WIN; ++q) xr[u][q] = ld16(k.x, xrow + (q / PW) * k.x.sh + (wo * PW + q % PW) * k.x.sw);
        }
        float pooled[U][8];
        bool any_pooled = false;
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int e = 0; e < 8; ++e) pooled[u][e] = 0.f;
#pragma unroll
          for (int q = 0; q < WIN; ++q)
#pragma unroll
            for (int e = 0; e < 8; ++e) g[u][q][e] = 0.f;
        }
        for (int s = 0; s < n_src; ++s) {
          const FV sv = k.src[s];
          if (k.kind[s] == 0) {
            const unsigned srow = n * sv.sn + ho * PH * sv.sh + v * 8;
            uint4 rr[U][WIN];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const unsigned wo = wo0 + u * step < (unsigned)k.Wo ? wo0 + u * step : (unsigned)k.Wo - 1;
#pragma unroll
              for (int q = 0; q < WIN; ++q) rr[u][q] = ld16(sv, srow + (q / PW) * sv.sh + (wo * PW + q % PW) * sv.sw);
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int q = 0; q < WIN; ++q) {
                float f[8];
                unpack8(rr[u][q], f);
#pragma unroll
                for (int e = 0; e < 8; ++e) g[u][q][e] += f[e];
              }
          } else if (WIN > 1) {
            const unsigned srow = n * sv.sn + ho * sv.sh + v * 8;
            any_pooled = true;
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const unsigned wo = wo0 + u * step < (unsigned)k.Wo ? wo0 + u * step : (unsigned)k.Wo - 1;
              float f[8];
              unpack8(ld16(sv, srow + wo * sv.sw), f);
#pragma unroll
              for (int e = 0; e < 8; ++e) pooled[u][e] += f[e];
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned wo = wo0 + u * step;
          if (wo >= (unsigned)k.Wo) break;
          float x[WIN][8];
#pragma unroll
          for (int q = 0; q < WIN; ++q) unpack8(xr[u][q], x[q]);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float t[WIN];
#pragma unroll
            for (int q = 0; q < WIN; ++q) t[q] = fmaf(x[q][e], sc[e], sf[e]);
            if (WIN > 1 && any_pooled) {
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
              for (int e = 0; e < 8; ++e) { acc_b[e] += g[u][q][e]; acc_g[e] = fmaf(g[u][q][e], fmaf(x[q][e], cB[e], cD[e]), acc_g[e]); }
            } else {
              float o[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = has_bn ? fmaf(cB[e], x[q][e], fmaf(sc[e], g[u][q][e], cD[e])) : g[u][q][e];
              if (k.x_relu_mask) {     // the BatchNorm's input is ReLU(.): its mask folded in (dx is then the gradient BEFORE that ReLU)
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = x[q][e] > 0.f ? o[e] : 0.f;
              }
              st16(k.dx, drow + (q / PW) * k.dx.sh + (wo * PW + q % PW) * k.dx.sw, pack8(o));
            }
          }
        }
      }
    }
  }
  if (PASS == 0) {
    float* mine = red + (size_t)threadIdx.x * 16;
#pragma unroll
    for (int e = 0; e < 8; ++e) { mine[e] = acc_b[e]; mine[8 + e] = acc_g[e]; }
    __syncthreads();
    if (trow == 0 && v * 8 < k.C) {
      float sb[8], sg[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { sb[e] = 0.f; sg[e] = 0.f; }
      for (int r = 0; r < k.rp; ++r) {
        const float* o = red + (size_t)(r * k.cvb + tcv) * 16;
#pragma unroll
        for (int e = 0; e < 8; ++e) { sb[e] += o[e]; sg[e] += o[8 + e]; }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { atomicAdd(k.acc_dbeta + v * 8 + e, sb[e]); atomicAdd(k.acc_dgamma + v * 8 + e, sg[e]); }
    }
  }
}

// BN + activation backward of the layer a pointwise head (Conv 1x1, HC <= 2 outputs) reads, with the head's backward folded in:
// the head's input gradient g[pix][c] = sum_o dlogits[pix][o] * W[c][o] is never materialised (the separate head kernel
// wrote it once and this kernel read it twice: 805 MB at the 256x256x64 output layer of config 2), and PASS 0, which has
// a = act(BN(x)) in registers, also accumulates the head's dW[c][o] = sum a * dlogits and db[o] = sum dlogits.
// Un-pooled tensors only; further direct sources (a transposed conv reading the same tensor under deep supervision) are added.
template <int PASS, int ACT, int HC, int U>
__global__ void __launch_bounds__(256, 2) bn_bwd_head_kernel(const BnBwdF k) {
  pdl_prologue();
  extern __shared__ float red[];   // PASS 0: [256][16 + 8 * HC + 1]
  constexpr int kRed = 16 + 8 * HC + 1;
  const int tcv = threadIdx.x % k.cvb, trow = threadIdx.x / k.cvb;
  const int v = blockIdx.x * k.cvb + tcv;
  const bool active = v * 8 < k.C && trow < k.rp;
  float sc[8], sf[8], cB[8], cD[8], acc_b[8], acc_g[8], hw[8][HC], acc_hw[HC][8], acc_hb[HC];
#pragma unroll
  for (int e = 0; e < 8; ++e) { acc_b[e] = 0.f; acc_g[e] = 0.f; sc[e] = 1.f; sf[e] = 0.f; cB[e] = 0.f; cD[e] = 0.f; }
#pragma unroll
  for (int o = 0; o < HC; ++o) {
    acc_hb[o] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) { acc_hw[o][e] = 0.f; hw[e][o] = 0.f; }
  }
  if (active) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = v * 8 + e;
      sc[e] = __ldg(k.scale + c);
      sf[e] = __ldg(k.shift + c);
      const float rs = __ldg(k.rstd + c), mu = __ldg(k.mean + c);
      if (PASS == 1) {
        const float cb = __ldg(k.dbeta + c) * k.inv_count, cg = __ldg(k.dgamma + c) * k.inv_count;
        cB[e] = -sc[e] * cg * rs;
        cD[e] = sc[e] * (cg * rs * mu - cb);
      } else {
        cB[e] = rs;
        cD[e] = -mu * rs;
      }
#pragma unroll
      for (int o = 0; o < HC; ++o) hw[e][o] = __ldg(k.head_w + (size_t)c * HC + o);
    }
    const unsigned step = (unsigned)k.rp;
    const int n_src = k.n_src;
    for (int rr = blockIdx.y; rr < k.rows; rr += gridDim.y) {
      const int r = PASS == 0 ? k.rows - 1 - rr : rr;
      const unsigned n = fast_div((unsigned)r, k.fd_ho), ho = (unsigned)r - n * (unsigned)k.Ho;
      const unsigned xrow = n * k.x.sn + ho * k.x.sh + v * 8;
      const unsigned drow = n * k.dx.sn + ho * k.dx.sh + v * 8;
      const size_t prow = (size_t)r * k.Wo;      // pixel index of the row start: dlogits is dense [N*H*W][HC]
      for (unsigned wo0 = trow; wo0 < (unsigned)k.Wo; wo0 += step * U) {
        uint4 xr[U];
        float g[U][8], dl[U][HC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned wo = wo0 + u * step < (unsigned)k.Wo ? wo0 + u * step : (unsigned)k.Wo - 1;
          xr[u] = ld16(k.x, xrow + wo * k.x.sw);
#pragma unroll
          for (int o = 0; o < HC; ++o) dl[u][o] = wo0 + u * step < (unsigned)k.Wo ? __ldg(k.dlogits + (prow + wo) * HC + o) : 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            g[u][e] = 0.f;
#pragma unroll
            for (int o = 0; o < HC; ++o) g[u][e] = fmaf(dl[u][o], hw[e][o], g[u][e]);
          }
        }
        for (int s = 0; s < n_src; ++s) {
          if (k.kind[s] != 0) continue;
          const FV sv = k.src[s];
          const unsigned srow = n * sv.sn + ho * sv.sh + v * 8;
          uint4 rr4[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const unsigned wo = wo0 + u * step < (unsigned)k.Wo ? wo0 + u * step : (unsigned)k.Wo - 1;
            rr4[u] = ld16(sv, srow + wo * sv.sw);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            float f[8];
            unpack8(rr4[u], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) g[u][e] += f[e];
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const unsigned wo = wo0 + u * step;
          if (wo >= (unsigned)k.Wo) break;
          float x[8];
          unpack8(xr[u], x);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float t = fmaf(x[e], sc[e], sf[e]);
            if (PASS == 0) {
              const float a = act_f<ACT>(t);
#pragma unroll
              for (int o = 0; o < HC; ++o) acc_hw[o][e] = fmaf(a, dl[u][o], acc_hw[o][e]);
            }
            if (ACT == B2SEG_ACT_RELU) g[u][e] = t > 0.f ? g[u][e] : 0.f;
            if (ACT == B2SEG_ACT_LEAKY) g[u][e] = t > 0.f ? g[u][e] : 0.3f * g[u][e];
          }
          if (PASS == 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { acc_b[e] += g[u][e]; acc_g[e] = fmaf(g[u][e], fmaf(x[e], cB[e], cD[e]), acc_g[e]); }
            if (v == 0) {
#pragma unroll
              for (int o = 0; o < HC; ++o) acc_hb[o] += dl[u][o];
            }
          } else {
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o8[e] = fmaf(cB[e], x[e], fmaf(sc[e], g[u][e], cD[e]));
            st16(k.dx, drow + wo * k.dx.sw, pack8(o8));
          }
        }
      }
    }
  }
  if (PASS == 0) {
    float* mine = red + (size_t)threadIdx.x * kRed;
#pragma unroll
    for (int e = 0; e < 8; ++e) { mine[e] = acc_b[e]; mine[8 + e] = acc_g[e]; }
#pragma unroll
    for (int o = 0; o < HC; ++o) {
#pragma unroll
      for (int e = 0; e < 8; ++e) mine[16 + o * 8 + e] = acc_hw[o][e];
    }
    mine[16 + 8 * HC] = 0.f;
    __syncthreads();
    if (trow == 0 && v * 8 < k.C) {
      float sum[16 + 8 * HC];
#pragma unroll
      for (int i = 0; i < 16 + 8 * HC; ++i) sum[i] = 0.f;
      for (int r = 0; r < k.rp; ++r) {
        const float* o = red + (size_t)(r * k.cvb + tcv) * kRed;
#pragma unroll
        for (int i = 0; i < 16 + 8 * HC; ++i) sum[i] += o[i];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(k.acc_dbeta + v * 8 + e, sum[e]);
        atomicAdd(k.acc_dgamma + v * 8 + e, sum[8 + e]);
#pragma unroll
        for (int o = 0; o < HC; ++o) atomicAdd(k.head_dw + (size_t)(v * 8 + e) * HC + o, sum[16 + o * 8 + e]);
      }
    }
    // db: the threads of channel vector 0 hold the per-pixel sums of dlogits
    if (v == 0 && active) {
#pragma unroll
      for (int o = 0; o < HC; ++o) atomicAdd(k.head_db + o, acc_hb[o]);
    }
  }
}

struct BnBwdFastLaunch : PreparedOp {
  BnBwdF k;
  bool has_bn;
  bool accumulate;
  int ph, pw, act;
  dim3 grid0, grid1;
  template <int PASS, int PH, int PW, int U>
  void go_act(dim3 grid, int smem, cudaStream_t s) {
    switch (act) {
      case B2SEG_ACT_RELU: launch_k(bn_bwd_fast_kernel<PASS, PH, PW, B2SEG_ACT_RELU, U>, grid, dim3(256), smem, s, k); break;
      case B2SEG_ACT_LEAKY: launch_k(bn_bwd_fast_kernel<PASS, PH, PW, B2SEG_ACT_LEAKY, U>, grid, dim3(256), smem, s, k); break;
      default: launch_k(bn_bwd_fast_kernel<PASS, PH, PW, B2SEG_ACT_NONE, U>, grid, dim3(256), smem, s, k); break;
    }
  }
  int head_cout = 0;
  template <int PASS, int HC>
  void go_head(dim3 grid, cudaStream_t s) {
    const int smem = PASS == 0 ? 256 * (16 + 8 * HC + 1) * 4 : 0;
    if (act == B2SEG_ACT_LEAKY) launch_k(bn_bwd_head_kernel<PASS, B2SEG_ACT_LEAKY, HC, 2>, grid, dim3(256), smem, s, k);
    else launch_k(bn_bwd_head_kernel<PASS, B2SEG_ACT_RELU, HC, 2>, grid, dim3(256), smem, s, k);
  }
  template <int PASS>
  void go(dim3 grid, int smem, cudaStream_t s) {
    if (head_cout == 1) { go_head<PASS, 1>(grid, s); return; }
    if (head_cout == 2) { go_head<PASS, 2>(grid, s); return; }
    if (ph == 2 && pw == 2) go_act<PASS, 2, 2, 1>(grid, smem, s);
    else if (ph == 1 && pw == 2) go_act<PASS, 1, 2, 2>(grid, smem, s);
    else go_act<PASS, 1, 1, 2>(grid, smem, s);
  }
  int launch(cudaStream_t s) override {
    if (has_bn) {
      if (!accumulate) {
        B2_CUDA_OK(cudaMemsetAsync(k.acc_dgamma, 0, (size_t)k.C * 4, s));
        B2_CUDA_OK(cudaMemsetAsync(k.acc_dbeta, 0, (size_t)k.C * 4, s));
      }
      go<0>(grid0, 256 * 16 * 4, s);
      B2_CUDA_OK(cudaGetLastError());
    }
    go<1>(grid1, 0, s);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
  int num_launches() const override { return has_bn ? 2 : 1; }
};

PreparedOp* prepare_bn_bwd_fast(const b2seg_bn_bwd_desc* d) {
  static const bool disabled = getenv("B2SEG_NO_FAST_STREAM") != nullptr;
  if (disabled || d->x.C % 8 || d->n_src < 1 || d->n_src > B2SEG_MAX_GRADSRC) return nullptr;
  if (d->act != B2SEG_ACT_NONE && d->act != B2SEG_ACT_RELU && d->act != B2SEG_ACT_LEAKY) return nullptr;
  int ph = 1, pw = 1;
  for (int i = 0; i < d->n_src; ++i)
    if (d->src[i].kind == 1) {
      const int a = d->src[i].pool_h > 1 ? d->src[i].pool_h : 1, b = d->src[i].pool_w > 1 ? d->src[i].pool_w : 1;
      if ((ph != 1 || pw != 1) && (ph != a || pw != b)) return nullptr;
      ph = a; pw = b;
    }
  if (!((ph == 1 && pw == 1) || (ph == 2 && pw == 2) || (ph == 1 && pw == 2))) return nullptr;
  if (d->x.H % ph || d->x.W % pw) return nullptr;
  const bool has_bn = d->scale != 0;
  if (has_bn && (!d->dgamma || !d->dbeta || !d->mean || !d->rstd || !d->shift)) return nullptr;
  auto* L = new BnBwdFastLaunch();
  BnBwdF& k = L->k;
  memset(&k, 0, sizeof(k));
  bool ok = fv_make(d->x, &k.x) && fv_make(d->dx, &k.dx);
  int n_head = 0;
  for (int i = 0; ok && i < d->n_src; ++i) {
    k.kind[i] = d->src[i].kind;
    if (d->src[i].kind == 2) {
      const b2seg_gradsrc& hs = d->src[i];
      ++n_head;
      // only what bn_bwd_head_kernel implements: BN + ReLU/LeakyReLU, no pooled source, at most 2 head outputs
      ok = n_head == 1 && has_bn && ph == 1 && pw == 1 && (hs.cout == 1 || hs.cout == 2) && hs.dlogits && hs.head_w && hs.head_dw && hs.head_db &&
           (d->act == B2SEG_ACT_RELU || d->act == B2SEG_ACT_LEAKY);
      k.dlogits = reinterpret_cast<const float*>(hs.dlogits); k.head_w = reinterpret_cast<const float*>(hs.head_w);
      k.head_dw = reinterpret_cast<float*>(hs.head_dw); k.head_db = reinterpret_cast<float*>(hs.head_db);
      L->head_cout = hs.cout;
      continue;
    }
    ok = fv_make(d->src[i].g, &k.src[i]);
    if (d->src[i].kind != 0 && d->src[i].kind != 1) ok = false;
  }
  if (ok && n_head) {
    for (int i = 0; i < d->n_src; ++i)
      if (d->src[i].kind == 1) ok = false;
  }
  if (!ok) { delete L; return nullptr; }
  k.n_src = d->n_src;
  k.x_relu_mask = d->x_relu_mask;
  if (d->x_relu_mask && (n_head || !has_bn)) { delete L; return nullptr; }
  k.scale = reinterpret_cast<const float*>(d->scale);
  k.shift = reinterpret_cast<const float*>(d->shift);
  k.mean = reinterpret_cast<const float*>(d->mean);
  k.rstd = reinterpret_cast<const float*>(d->rstd);
  k.dgamma = reinterpret_cast<const float*>(d->dgamma);
  k.dbeta = reinterpret_cast<const float*>(d->dbeta);
  k.acc_dgamma = reinterpret_cast<float*>(d->dgamma);
  k.acc_dbeta = reinterpret_cast<float*>(d->dbeta);
  k.inv_count = (float)(1.0 / d->count);
  k.C = d->x.C; k.Ho = d->x.H / ph; k.Wo = d->x.W / pw; k.rows = d->x.N * k.Ho;
  k.fd_ho = make_fastdiv((uint32_t)k.Ho);
  L->has_bn = has_bn; L->ph = ph; L->pw = pw; L->act = d->act; L->accumulate = d->accumulate != 0;
  row_grid(k.C, k.rows, k.Wo, &k.cvb, &k.rp, &L->grid1);
  // PASS 0 ends with 2 * 8 * cvb atomics per block: one resident wave of blocks walking the rows keeps them few
  L->grid0 = L->grid1;
  const unsigned wave = (unsigned)std::max(1, num_sms() * ((ph * pw > 1 || L->head_cout) ? 2 : 3) / (int)L->grid0.x);   // = blocks resident at once (launch bounds)
  if (L->grid0.y > wave) L->grid0.y = wave;
  return L;
}

// ---------------------------------------------------------------------------------------------------- pointwise heads
// The `out` / `level{k}` 1x1 convolutions (Cout <= 8) on a pixel-contiguous view: pixel p starts p * pitch elements after
// the first, so no coordinate decoding is needed at all.  G = C/8 lanes share a pixel (one 16-byte vector each), the head
// weights live in registers, U pixels are in flight per thread.  The backward kernel produces dx, dW and db in ONE pass
// over x (the generic path reads dlogits twice and runs two kernels).
struct HeadF {
  unsigned long long x, dx;
  unsigned pitch, dpitch;        // elements between consecutive pixels
  unsigned n_pix;
  int G, act, has_dx;
  const float* w; const float* b;
  float* y; float* logits;
  const float* dl; float* dw; float* db;
  const float* sc; const float* sh; int pre_act;   // forward only: the head consumes act(x * sc + sh) (x = raw conv output of the last layer)
};

static bool pixel_contiguous(const b2seg_view& v, unsigned* pitch) {
  if (v.ptr == 0 || (v.ptr & 15) || v.sw <= 0 || (v.sw % 8)) return false;
  if (v.W > 1 && v.H > 1 && v.sh != (long long)v.W * v.sw) return false;
  if (v.N > 1 && v.sn != (long long)v.H * (v.H > 1 || v.W > 1 ? (v.H > 1 ? v.sh : (long long)v.W * v.sw) : v.sw)) return false;
  if ((long long)v.N * v.H * v.W * v.sw >= (1ll << 31)) return false;
  *pitch = (unsigned)v.sw;
  return true;
}

constexpr int head_fwd_minb(int cout) { return cout <= 2 ? 4 : (cout <= 4 ? 3 : 2); }
constexpr int head_bwd_minb(int cout) { return cout == 1 ? 4 : (cout == 2 ? 3 : (cout <= 4 ? 2 : 1)); }

template <int COUT, int U>
__global__ void __launch_bounds__(256, head_fwd_minb(COUT)) head_fwd_fast_kernel(const HeadF k) {
  pdl_prologue();
  const int G = k.G;
  const int gl = threadIdx.x % G;
  const unsigned gid = (blockIdx.x * 256u + threadIdx.x) / G;
  const unsigned gstride = gridDim.x * 256u / G;
  float w[8][COUT];
#pragma unroll
  for (int e = 0; e < 8; ++e)
#pragma unroll
    for (int o = 0; o < COUT; ++o) w[e][o] = __ldg(k.w + (size_t)(gl * 8 + e) * COUT + o);
  float bias[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) bias[o] = __ldg(k.b + o);
  // BatchNorm + activation of the layer the head reads, applied on the fly: that layer's activated tensor (only the head and its
  // folded backward would ever touch it) is then never written or read
  const bool pre = k.sc != nullptr;
  float psc[8], psh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { psc[e] = pre ? __ldg(k.sc + gl * 8 + e) : 1.f; psh[e] = pre ? __ldg(k.sh + gl * 8 + e) : 0.f; }
  const unsigned n_iter = (k.n_pix + gstride * U - 1) / (gstride * U);   // uniform trip count: the shuffles need whole warps
  for (unsigned it = 0; it < n_iter; ++it) {
    uint4 raw[U];
    unsigned pix[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      pix[u] = gid + (it * U + u) * gstride;
      const unsigned pc = pix[u] < k.n_pix ? pix[u] : k.n_pix - 1;
      raw[u] = __ldg(reinterpret_cast<const uint4*>(k.x + ((unsigned long long)pc * k.pitch + gl * 8) * 2ull));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float f[8], acc[COUT];
      unpack8(raw[u], f);
      if (pre) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float t = fmaf(f[e], psc[e], psh[e]);
          f[e] = k.pre_act == B2SEG_ACT_RELU ? fmaxf(t, 0.f) : (k.pre_act == B2SEG_ACT_LEAKY ? (t > 0.f ? t : 0.3f * t) : t);
        }
      }
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        acc[o] = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[o] = fmaf(f[e], w[e][o], acc[o]);
      }
      for (int off = G / 2; off > 0; off >>= 1)
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], off, 32);
      if (gl == 0 && pix[u] < k.n_pix) {
        float z[COUT];
        float zmax = -INFINITY;
#pragma unroll
        for (int o = 0; o < COUT; ++o) { z[o] = acc[o] + bias[o]; zmax = fmaxf(zmax, z[o]); }
        if (k.logits) {
#pragma unroll
          for (int o = 0; o < COUT; ++o) k.logits[(size_t)pix[u] * COUT + o] = z[o];
        }
        if (k.act == B2SEG_ACT_SOFTMAX) {
          float den = 0.f;
#pragma unroll
          for (int o = 0; o < COUT; ++o) { z[o] = __expf(z[o] - zmax); den += z[o]; }
#pragma unroll
          for (int o = 0; o < COUT; ++o) k.y[(size_t)pix[u] * COUT + o] = z[o] / den;
        } else {
#pragma unroll
          for (int o = 0; o < COUT; ++o) k.y[(size_t)pix[u] * COUT + o] = act_fwd(z[o], k.act);
        }
      }
    }
  }
}

template <int COUT, int U>
__global__ void __launch_bounds__(256, head_bwd_minb(COUT)) head_bwd_fast_kernel(const HeadF k) {
  pdl_prologue();
  extern __shared__ float red[];   // [256][9]
  const int G = k.G;
  const int gl = threadIdx.x % G;
  const unsigned gid = (blockIdx.x * 256u + threadIdx.x) / G;
  const unsigned gstride = gridDim.x * 256u / G;
  float w[8][COUT], acc[COUT][8], accb[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
    accb[o] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) { w[e][o] = __ldg(k.w + (size_t)(gl * 8 + e) * COUT + o); acc[o][e] = 0.f; }
  }
  for (unsigned p0 = gid; p0 < k.n_pix; p0 += gstride * U) {
    uint4 raw[U];
    float d[U][COUT];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned pix = p0 + u * gstride;
      const bool ok = pix < k.n_pix;
      const unsigned pc = ok ? pix : k.n_pix - 1;
      raw[u] = __ldg(reinterpret_cast<const uint4*>(k.x + ((unsigned long long)pc * k.pitch + gl * 8) * 2ull));
#pragma unroll
      for (int o = 0; o < COUT; ++o) d[u][o] = ok ? __ldg(k.dl + (size_t)pc * COUT + o) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned pix = p0 + u * gstride;
      float f[8], o8[8];
      unpack8(raw[u], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) o8[e] = 0.f;
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        accb[o] += d[u][o];
#pragma unroll
        for (int e = 0; e < 8; ++e) { acc[o][e] = fmaf(d[u][o], f[e], acc[o][e]); o8[e] = fmaf(d[u][o], w[e][o], o8[e]); }
      }
      if (k.has_dx && pix < k.n_pix)
        *reinterpret_cast<uint4*>(k.dx + ((unsigned long long)pix * k.dpitch + gl * 8) * 2ull) = pack8(o8);
    }
  }
  const int rows = 256 / G, trow = threadIdx.x / G;
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
    float* mine = red + (size_t)threadIdx.x * 9;
#pragma unroll
    for (int e = 0; e < 8; ++e) mine[e] = acc[o][e];
    mine[8] = accb[o];
    __syncthreads();
    if (trow == 0) {
      float s9[9];
#pragma unroll
      for (int e = 0; e < 9; ++e) s9[e] = 0.f;
      for (int r = 0; r < rows; ++r) {
        const float* q = red + (size_t)(r * G + gl) * 9;
#pragma unroll
        for (int e = 0; e < 9; ++e) s9[e] += q[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(k.dw + (size_t)(gl * 8 + e) * COUT + o, s9[e]);
      if (gl == 0) atomicAdd(k.db + o, s9[8]);
    }
    __syncthreads();
  }
}

#define B2_FAST_COUT_SWITCH(cout, CALL)            \
  switch (cout) {                                  \
    case 1: { constexpr int CO = 1; CALL; } break; \
    case 2: { constexpr int CO = 2; CALL; } break; \
    case 3: { constexpr int CO = 3; CALL; } break; \
    case 4: { constexpr int CO = 4; CALL; } break; \
    case 5: { constexpr int CO = 5; CALL; } break; \
    case 6: { constexpr int CO = 6; CALL; } break; \
    case 7: { constexpr int CO = 7; CALL; } break; \
    default: { constexpr int CO = 8; CALL; } break; \
  }

struct HeadFastLaunch : PreparedOp {
  HeadF k;
  int cout;
  bool bwd;
  int launch(cudaStream_t s) override {
    const long long threads = (long long)k.n_pix * k.G;
    long long grid = (threads / 4 + 255) / 256;
    const int cap = num_sms() * (bwd ? head_bwd_minb(cout) : head_fwd_minb(cout));
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    if (bwd) {
      B2_FAST_COUT_SWITCH(cout, (launch_k(head_bwd_fast_kernel<CO, 4>, dim3((unsigned)grid), dim3(256), 256 * 9 * 4, s, k)));
    } else {
      B2_FAST_COUT_SWITCH(cout, (launch_k(head_fwd_fast_kernel<CO, 4>, dim3((unsigned)grid), dim3(256), 0, s, k)));
    }
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};

PreparedOp* prepare_head_fast(const b2seg_head_desc* d, bool bwd) {
  static const bool disabled = getenv("B2SEG_NO_FAST_STREAM") != nullptr;
  if (disabled || d->stride > 1 || d->cout < 1 || d->cout > 8 || d->x.C % 8) return nullptr;
  const int cvec = d->x.C / 8;
  if (cvec > 32 || (cvec & (cvec - 1))) return nullptr;
  HeadF k;
  memset(&k, 0, sizeof(k));
  if (!pixel_contiguous(d->x, &k.pitch)) return nullptr;
  k.x = d->x.ptr;
  k.n_pix = (unsigned)((long long)d->x.N * d->x.H * d->x.W);
  if (k.n_pix == 0) return nullptr;
  k.G = cvec; k.act = d->act;
  k.w = reinterpret_cast<const float*>(d->w); k.b = reinterpret_cast<const float*>(d->b);
  k.y = reinterpret_cast<float*>(d->y); k.logits = reinterpret_cast<float*>(d->logits);
  if (bwd) {
    k.has_dx = d->dx.ptr != 0;
    if (k.has_dx) {
      if (!pixel_contiguous(d->dx, &k.dpitch) || d->dx.N != d->x.N || d->dx.H != d->x.H || d->dx.W != d->x.W) return nullptr;
      k.dx = d->dx.ptr;
    }
    k.dl = reinterpret_cast<const float*>(d->dlogits); k.dw = reinterpret_cast<float*>(d->dw); k.db = reinterpret_cast<float*>(d->db);
    if (!k.dl || !k.dw || !k.db) return nullptr;
  } else if (!k.w || !k.b || !k.y) {
    return nullptr;
  }
  if (d->bn_scale) {
    if (bwd || !d->bn_shift || (d->bn_act != B2SEG_ACT_NONE && d->bn_act != B2SEG_ACT_RELU && d->bn_act != B2SEG_ACT_LEAKY)) return nullptr;
    k.sc = reinterpret_cast<const float*>(d->bn_scale); k.sh = reinterpret_cast<const float*>(d->bn_shift); k.pre_act = d->bn_act;
  }
  auto* L = new HeadFastLaunch();
  L->k = k; L->cout = d->cout; L->bwd = bwd;
  return L;
}

// ---------------------------------------------------------------------------------------------------- row sum
// out[c] (+)= sum_r partials[r * pitch + c]: 32 channels x 32 row slices per block, fixed summation order.
__global__ void __launch_bounds__(1024) rowsum_kernel(const float* __restrict__ partials, int n_rows, int pitch, int C, float* out, int accumulate) {
  pdl_prologue();
  __shared__ float sh[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float s = 0.f;
  if (c < C)
    for (int r = ty; r < n_rows; r += 32) s += __ldg(partials + (size_t)r * pitch + c);
  sh[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
    for (int i = 1; i < 32; ++i) s += sh[i][tx];
    out[c] = accumulate ? out[c] + s : s;
  }
}
struct RowsumLaunch : PreparedOp {
  b2seg_rowsum_desc d;
  int launch(cudaStream_t s) override {
    launch_k(rowsum_kernel, dim3((d.C + 31) / 32), dim3(1024), 0, s, reinterpret_cast<const float*>(d.partials), d.n_rows, d.pitch, d.C,
             reinterpret_cast<float*>(d.out), d.accumulate);
    B2_CUDA_OK(cudaGetLastError());
    return 0;
  }
};
PreparedOp* prepare_rowsum(const b2seg_rowsum_desc* d) {
  if (!d->partials || !d->out || d->n_rows < 1 || d->C < 1 || d->pitch < d->C) { set_error("rowsum: bad arguments"); return nullptr; }
  auto* L = new RowsumLaunch(); L->d = *d; return L;
}

}  // namespace b2
