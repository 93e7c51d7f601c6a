// Helpers shared by the HBM-streaming kernels: strided NHWC views, 16-byte vector load/store, activations.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace b2 {

struct DView {
  unsigned long long ptr;
  int N, H, W, C;
  long long sn, sh, sw;
};
static inline DView dv(const b2seg_view& v) { return DView{v.ptr, v.N, v.H, v.W, v.C, v.sn, v.sh, v.sw}; }

__device__ __forceinline__ const __nv_bfloat16* vaddr(const DView& v, int n, int h, int w, int c) {
  return reinterpret_cast<const __nv_bfloat16*>(v.ptr) + n * v.sn + h * v.sh + w * v.sw + c;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(const __nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(const_cast<__nv_bfloat16*>(p)) = u;
}
__device__ __forceinline__ float act_fwd(float x, int act) {
  switch (act) {
    case B2SEG_ACT_RELU: return fmaxf(x, 0.f);
    case B2SEG_ACT_LEAKY: return x > 0.f ? x : 0.3f * x;
    case B2SEG_ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
    case B2SEG_ACT_TANH: return tanhf(x);
    default: return x;
  }
}
// derivative of the activation given its output y
__device__ __forceinline__ float act_bwd_from_y(float y, int act) {
  switch (act) {
    case B2SEG_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case B2SEG_ACT_LEAKY: return y > 0.f ? 1.f : 0.3f;
    case B2SEG_ACT_SIGMOID: return y * (1.f - y);
    case B2SEG_ACT_TANH: return 1.f - y * y;
    default: return 1.f;
  }
}

static inline int grid_for(long long work, int block) {
  long long g = (work + block - 1) / block;
  if (g < 1) g = 1;
  if (g > 0x7fffffffll) g = 0x7fffffffll;
  return (int)g;
}


}  // namespace b2
