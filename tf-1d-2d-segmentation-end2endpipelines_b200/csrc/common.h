// Host-side helpers shared by the b2seg translation units: error reporting, tensor-map encoding,
// and the "prepared launch" objects a plan replays.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

#include "../../include/b2seg.h"

namespace b2 {

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define B2_CUDA_OK(expr)                                                                          \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) return b2::fail(B2SEG_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

int num_sms();
int require_sm100();
bool pdl_enabled();   // programmatic dependent launch (opt-in with B2SEG_PDL=1)

#ifdef __CUDACC__
// Launch with the programmatic-stream-serialization attribute: the kernel may become resident before its predecessor in
// the stream has drained; it must call pdl_prologue() / pdl_wait() (ptx.cuh) before its first global-memory access.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// 4-D activation map over a view: dims (C, W, H, N), bf16, SWIZZLE_128B, zero OOB fill.
int encode_act_map(CUtensorMap* m, const b2seg_view& v, int box_c, int box_w, int box_h, int box_n);
// 3-D weight map over bf16 [cout][taps][cin]: dims (cin, taps, cout).
int encode_weight_map(CUtensorMap* m, uint64_t ptr, int cout, int taps, int cin, int box_cin, int box_cout);

// pick (bw, bh, bn) with bw*bh*bn == pixels (a power of two) covering the (N,H,W) grid with little waste
void pick_box(int N, int H, int W, int pixels, int* bw, int* bh, int* bn);

struct PreparedOp {
  virtual ~PreparedOp() {}
  virtual int launch(cudaStream_t s) = 0;
  virtual int num_launches() const { return 1; }
};

PreparedOp* prepare_conv(const b2seg_conv_desc* d);
PreparedOp* prepare_wgrad(const b2seg_wgrad_desc* d);
PreparedOp* prepare_wgrad_halo(const b2seg_wgrad_desc* d, bool* hard_error);
PreparedOp* prepare_bn_finalize(const b2seg_bn_finalize_desc* d);
PreparedOp* prepare_bn_act(const b2seg_bn_act_desc* d);
PreparedOp* prepare_bn_bwd(const b2seg_bn_bwd_desc* d);
PreparedOp* prepare_bn_act_fast(const b2seg_bn_act_desc* d);   // nullptr (no error) when not eligible
PreparedOp* prepare_bn_bwd_fast(const b2seg_bn_bwd_desc* d);
PreparedOp* prepare_rowsum(const b2seg_rowsum_desc* d);
PreparedOp* prepare_head_fast(const b2seg_head_desc* d, bool bwd);   // nullptr (no error) when not eligible
PreparedOp* prepare_adam(const b2seg_adam_desc* d);
PreparedOp* prepare_head_fwd(const b2seg_head_desc* d);
PreparedOp* prepare_head_bwd(const b2seg_head_desc* d);
PreparedOp* prepare_loss(const b2seg_loss_desc* d);
PreparedOp* prepare_eltwise(const b2seg_eltwise_desc* d);
PreparedOp* prepare_cast(const b2seg_cast_desc* d);
PreparedOp* prepare_colsum(const b2seg_colsum_desc* d);
PreparedOp* prepare_memset(const b2seg_memset_desc* d);
PreparedOp* prepare_resize_fwd(const b2seg_resize_desc* d);
PreparedOp* prepare_resize_bwd(const b2seg_resize_desc* d);
PreparedOp* prepare_mulbc_fwd(const b2seg_mulbc_desc* d);
PreparedOp* prepare_mulbc_bwd(const b2seg_mulbc_desc* d);
PreparedOp* prepare_colstats(const b2seg_colstats_desc* d);
PreparedOp* prepare_lstm_fwd(const b2seg_lstm_desc* d);
PreparedOp* prepare_lstm_bwd(const b2seg_lstm_desc* d);
PreparedOp* prepare_pool_bwd(const b2seg_poolbwd_desc* d);
PreparedOp* prepare_outact_fwd(const b2seg_outact_desc* d);
PreparedOp* prepare_outact_bwd(const b2seg_outact_desc* d);
PreparedOp* prepare_target_pool(const b2seg_tpool_desc* d);
PreparedOp* prepare_gate_fwd(const b2seg_gate_desc* d);
PreparedOp* prepare_gate_bwd(const b2seg_gate_desc* d);
PreparedOp* prepare_fold_bn(const b2seg_fold_desc* d);

// adam launches expose their mutable hyper-parameters to the plan
void adam_update(PreparedOp* op, float lr, int64_t step, float grad_scale);
bool is_adam(PreparedOp* op);

int read_halo_trace(unsigned long long* out, int n);   // conv_halo.cu (debug)
int conv_num_mtiles(const b2seg_conv_desc* d);
int conv_num_stat_rows(const b2seg_conv_desc* d);

}  // namespace b2
