// C ABI (include/b2seg.h): error plumbing, tensor-map encoding, op-level entry points and the plan object.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <memory>
#include <vector>

#include "common.h"

namespace b2 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// Data-parallel training: the backward phase leaves `g_bwd_sm_reserve` SMs to the collective's CTAs.  Every tensor-core
// kernel here is persistent with one CTA per SM and a static tile split, so a single SM held by an NCCL CTA turns a
// launch into two waves (measured at 8 GPUs: +1.5 ms per step with NVLS' 24 CTAs); a grid of (SMs - reserve) runs beside them.
static int g_bwd_sm_reserve = 0;
static thread_local int g_sm_override = 0;

int num_sms() {
  if (g_sm_override > 0) return g_sm_override;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

bool pdl_enabled() {
  static const bool on = getenv("B2SEG_PDL") != nullptr;   // measured on cfg2: 2821 (on) vs 2848 (off) images/s -- no gain, so opt-in
  return on;
}

int require_sm100() {
  static int checked = 0;  // 0 unknown, 1 ok, -1 bad
  if (checked == 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      checked = -1;
    } else {
      checked = (major == 10) ? 1 : -1;
    }
  }
  if (checked != 1) return fail(B2SEG_ERR_DEVICE, "b2seg requires an sm_100 (B200) CUDA device; there is no CPU fallback");
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_act_map(CUtensorMap* m, const b2seg_view& v, int box_c, int box_w, int box_h, int box_n) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(B2SEG_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  if ((v.ptr & 15) || (v.sw % 8) || (v.sh % 8) || (v.sn % 8) || v.C < 1)
    return fail(B2SEG_ERR_ARG, "view not 16-byte aligned (ptr=%llx sw=%lld sh=%lld sn=%lld)", (unsigned long long)v.ptr, (long long)v.sw,
                (long long)v.sh, (long long)v.sn);
  cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
  cuuint64_t strides[3] = {(cuuint64_t)v.sw * 2, (cuuint64_t)v.sh * 2, (cuuint64_t)v.sn * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, reinterpret_cast<void*>(v.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(B2SEG_ERR_CUDA, "cuTensorMapEncodeTiled(act) failed: %d (dims %d,%d,%d,%d strides %lld,%lld,%lld box %d,%d,%d,%d)", (int)r,
                v.C, v.W, v.H, v.N, (long long)v.sw, (long long)v.sh, (long long)v.sn, box_c, box_w, box_h, box_n);
  return 0;
}

int encode_weight_map(CUtensorMap* m, uint64_t ptr, int cout, int taps, int cin, int box_cin, int box_cout) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(B2SEG_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  if ((ptr & 15) || (cin % 8)) return fail(B2SEG_ERR_ARG, "weights not 16-byte aligned / cin %% 8");
  cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)taps, (cuuint64_t)cout};
  cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)cin * taps * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cin, 1, (cuuint32_t)box_cout};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, reinterpret_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B2SEG_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  return 0;
}

static int run_once(PreparedOp* op, void* stream) {
  if (!op) return g_err[0] ? B2SEG_ERR_ARG : fail(B2SEG_ERR_ARG, "prepare failed");
  std::unique_ptr<PreparedOp> guard(op);
  return op->launch(reinterpret_cast<cudaStream_t>(stream));
}

static PreparedOp* prepare_any(int op, const void* desc, size_t bytes) {
#define B2_CASE(code, type, fn)                                                                      \
  case code:                                                                                         \
    if (bytes != sizeof(type)) { set_error("op %d: descriptor size %zu != %zu", op, bytes, sizeof(type)); return nullptr; } \
    return fn(reinterpret_cast<const type*>(desc));
  switch (op) {
    B2_CASE(B2SEG_OP_CONV, b2seg_conv_desc, prepare_conv)
    B2_CASE(B2SEG_OP_WGRAD, b2seg_wgrad_desc, prepare_wgrad)
    B2_CASE(B2SEG_OP_BN_FINALIZE, b2seg_bn_finalize_desc, prepare_bn_finalize)
    B2_CASE(B2SEG_OP_BN_ACT, b2seg_bn_act_desc, prepare_bn_act)
    B2_CASE(B2SEG_OP_BN_BWD, b2seg_bn_bwd_desc, prepare_bn_bwd)
    B2_CASE(B2SEG_OP_ADAM, b2seg_adam_desc, prepare_adam)
    B2_CASE(B2SEG_OP_HEAD_FWD, b2seg_head_desc, prepare_head_fwd)
    B2_CASE(B2SEG_OP_HEAD_BWD, b2seg_head_desc, prepare_head_bwd)
    B2_CASE(B2SEG_OP_LOSS, b2seg_loss_desc, prepare_loss)
    B2_CASE(B2SEG_OP_ELTWISE, b2seg_eltwise_desc, prepare_eltwise)
    B2_CASE(B2SEG_OP_CAST, b2seg_cast_desc, prepare_cast)
    B2_CASE(B2SEG_OP_COLSUM, b2seg_colsum_desc, prepare_colsum)
    B2_CASE(B2SEG_OP_MEMSET, b2seg_memset_desc, prepare_memset)
    B2_CASE(B2SEG_OP_RESIZE_FWD, b2seg_resize_desc, prepare_resize_fwd)
    B2_CASE(B2SEG_OP_RESIZE_BWD, b2seg_resize_desc, prepare_resize_bwd)
    B2_CASE(B2SEG_OP_MULBC_FWD, b2seg_mulbc_desc, prepare_mulbc_fwd)
    B2_CASE(B2SEG_OP_MULBC_BWD, b2seg_mulbc_desc, prepare_mulbc_bwd)
    B2_CASE(B2SEG_OP_COLSTATS, b2seg_colstats_desc, prepare_colstats)
    B2_CASE(B2SEG_OP_LSTM_FWD, b2seg_lstm_desc, prepare_lstm_fwd)
    B2_CASE(B2SEG_OP_LSTM_BWD, b2seg_lstm_desc, prepare_lstm_bwd)
    B2_CASE(B2SEG_OP_POOL_BWD, b2seg_poolbwd_desc, prepare_pool_bwd)
    B2_CASE(B2SEG_OP_ROWSUM, b2seg_rowsum_desc, prepare_rowsum)
    B2_CASE(B2SEG_OP_OUTACT_FWD, b2seg_outact_desc, prepare_outact_fwd)
    B2_CASE(B2SEG_OP_OUTACT_BWD, b2seg_outact_desc, prepare_outact_bwd)
    B2_CASE(B2SEG_OP_TARGET_POOL, b2seg_tpool_desc, prepare_target_pool)
    B2_CASE(B2SEG_OP_GATE_FWD, b2seg_gate_desc, prepare_gate_fwd)
    B2_CASE(B2SEG_OP_GATE_BWD, b2seg_gate_desc, prepare_gate_bwd)
    B2_CASE(B2SEG_OP_FOLD_BN, b2seg_fold_desc, prepare_fold_bn)
    default:
      set_error("unknown op code %d", op);
      return nullptr;
  }
#undef B2_CASE
}

}  // namespace b2

struct b2seg_plan {
  std::vector<std::unique_ptr<b2::PreparedOp>> phase[3];
};

extern "C" {

const char* b2seg_last_error(void) { return b2::g_err; }
int b2seg_version(void) { return 103; }   // 103: b2seg_gate_fwd/bwd, b2seg_conv_desc.stats_atomic; 102: b2seg_loss kinds 4..14 + metrics; 101: eltwise ops 4/5, B2SEG_ACT_TANH, b2seg_outact_fwd/bwd, b2seg_target_pool (additive)

// sizeof of each op descriptor: lets a binding verify its struct mirrors without touching the GPU
int b2seg_sizeof_desc(int op) {
  switch (op) {
    case B2SEG_OP_CONV: return (int)sizeof(b2seg_conv_desc);
    case B2SEG_OP_WGRAD: return (int)sizeof(b2seg_wgrad_desc);
    case B2SEG_OP_BN_FINALIZE: return (int)sizeof(b2seg_bn_finalize_desc);
    case B2SEG_OP_BN_ACT: return (int)sizeof(b2seg_bn_act_desc);
    case B2SEG_OP_BN_BWD: return (int)sizeof(b2seg_bn_bwd_desc);
    case B2SEG_OP_ADAM: return (int)sizeof(b2seg_adam_desc);
    case B2SEG_OP_HEAD_FWD:
    case B2SEG_OP_HEAD_BWD: return (int)sizeof(b2seg_head_desc);
    case B2SEG_OP_LOSS: return (int)sizeof(b2seg_loss_desc);
    case B2SEG_OP_ELTWISE: return (int)sizeof(b2seg_eltwise_desc);
    case B2SEG_OP_CAST: return (int)sizeof(b2seg_cast_desc);
    case B2SEG_OP_COLSUM: return (int)sizeof(b2seg_colsum_desc);
    case B2SEG_OP_MEMSET: return (int)sizeof(b2seg_memset_desc);
    case B2SEG_OP_RESIZE_FWD:
    case B2SEG_OP_RESIZE_BWD: return (int)sizeof(b2seg_resize_desc);
    case B2SEG_OP_MULBC_FWD:
    case B2SEG_OP_MULBC_BWD: return (int)sizeof(b2seg_mulbc_desc);
    case B2SEG_OP_COLSTATS: return (int)sizeof(b2seg_colstats_desc);
    case B2SEG_OP_LSTM_FWD:
    case B2SEG_OP_LSTM_BWD: return (int)sizeof(b2seg_lstm_desc);
    case B2SEG_OP_POOL_BWD: return (int)sizeof(b2seg_poolbwd_desc);
    case B2SEG_OP_ROWSUM: return (int)sizeof(b2seg_rowsum_desc);
    case B2SEG_OP_OUTACT_FWD:
    case B2SEG_OP_OUTACT_BWD: return (int)sizeof(b2seg_outact_desc);
    case B2SEG_OP_TARGET_POOL: return (int)sizeof(b2seg_tpool_desc);
    case B2SEG_OP_GATE_FWD:
    case B2SEG_OP_GATE_BWD: return (int)sizeof(b2seg_gate_desc);
    case B2SEG_OP_FOLD_BN: return (int)sizeof(b2seg_fold_desc);
    default: return -1;
  }
}

int b2seg_device_check(int device) {
  if (cudaSetDevice(device) != cudaSuccess) return b2::fail(B2SEG_ERR_DEVICE, "cudaSetDevice(%d) failed", device);
  return b2::require_sm100();
}

#define B2_ENTRY(name, type, prep)                          \
  int name(const type* d, void* stream) {                   \
    b2::g_err[0] = 0;                                       \
    if (!d) return b2::fail(B2SEG_ERR_ARG, #name ": null descriptor"); \
    int rc = b2::require_sm100();                           \
    if (rc) return rc;                                      \
    return b2::run_once(prep(d), stream);                   \
  }

B2_ENTRY(b2seg_conv, b2seg_conv_desc, b2::prepare_conv)
B2_ENTRY(b2seg_wgrad, b2seg_wgrad_desc, b2::prepare_wgrad)
B2_ENTRY(b2seg_bn_finalize, b2seg_bn_finalize_desc, b2::prepare_bn_finalize)
B2_ENTRY(b2seg_bn_act, b2seg_bn_act_desc, b2::prepare_bn_act)
B2_ENTRY(b2seg_bn_bwd, b2seg_bn_bwd_desc, b2::prepare_bn_bwd)
B2_ENTRY(b2seg_adam, b2seg_adam_desc, b2::prepare_adam)
B2_ENTRY(b2seg_head_fwd, b2seg_head_desc, b2::prepare_head_fwd)
B2_ENTRY(b2seg_head_bwd, b2seg_head_desc, b2::prepare_head_bwd)
B2_ENTRY(b2seg_loss, b2seg_loss_desc, b2::prepare_loss)
B2_ENTRY(b2seg_eltwise, b2seg_eltwise_desc, b2::prepare_eltwise)
B2_ENTRY(b2seg_cast_input, b2seg_cast_desc, b2::prepare_cast)
B2_ENTRY(b2seg_colsum, b2seg_colsum_desc, b2::prepare_colsum)
B2_ENTRY(b2seg_resize_fwd, b2seg_resize_desc, b2::prepare_resize_fwd)
B2_ENTRY(b2seg_resize_bwd, b2seg_resize_desc, b2::prepare_resize_bwd)
B2_ENTRY(b2seg_mulbc_fwd, b2seg_mulbc_desc, b2::prepare_mulbc_fwd)
B2_ENTRY(b2seg_mulbc_bwd, b2seg_mulbc_desc, b2::prepare_mulbc_bwd)
B2_ENTRY(b2seg_colstats, b2seg_colstats_desc, b2::prepare_colstats)
B2_ENTRY(b2seg_lstm_fwd, b2seg_lstm_desc, b2::prepare_lstm_fwd)
B2_ENTRY(b2seg_lstm_bwd, b2seg_lstm_desc, b2::prepare_lstm_bwd)
B2_ENTRY(b2seg_pool_bwd, b2seg_poolbwd_desc, b2::prepare_pool_bwd)
B2_ENTRY(b2seg_rowsum, b2seg_rowsum_desc, b2::prepare_rowsum)
B2_ENTRY(b2seg_outact_fwd, b2seg_outact_desc, b2::prepare_outact_fwd)
B2_ENTRY(b2seg_outact_bwd, b2seg_outact_desc, b2::prepare_outact_bwd)
B2_ENTRY(b2seg_target_pool, b2seg_tpool_desc, b2::prepare_target_pool)
B2_ENTRY(b2seg_gate_fwd, b2seg_gate_desc, b2::prepare_gate_fwd)
B2_ENTRY(b2seg_gate_bwd, b2seg_gate_desc, b2::prepare_gate_bwd)
B2_ENTRY(b2seg_fold_bn, b2seg_fold_desc, b2::prepare_fold_bn)

int b2seg_conv_num_mtiles(const b2seg_conv_desc* d) {
  if (!d) return b2::fail(B2SEG_ERR_ARG, "null descriptor");
  return b2::conv_num_mtiles(d);
}

int b2seg_conv_num_stat_rows(const b2seg_conv_desc* d) {
  if (!d) return b2::fail(B2SEG_ERR_ARG, "null descriptor");
  return b2::conv_num_stat_rows(d);
}

// debug aid (B2SEG_TRACE=1): per-tile clock64() stamps of CTA 0 of the last conv_halo launch, [tile][8] =
// {mma: accumulator free, first A tile landed, all MMAs issued; producer: first TMA issued; epilogue: accumulator full,
//  TMEM read done, tile finished, unused}
int b2seg_debug_read_trace(uint64_t* out, int n) { return b2::read_halo_trace(reinterpret_cast<unsigned long long*>(out), n); }

int b2seg_set_backward_sm_reserve(int sms) {
  if (sms < 0 || sms > 64) return b2::fail(B2SEG_ERR_ARG, "backward SM reserve must be in 0..64");
  b2::g_bwd_sm_reserve = sms;
  return 0;
}

int b2seg_plan_create(b2seg_plan** out) {
  b2::g_err[0] = 0;
  if (!out) return b2::fail(B2SEG_ERR_ARG, "null out");
  int rc = b2::require_sm100();
  if (rc) return rc;
  *out = new b2seg_plan();
  return 0;
}

int b2seg_plan_add(b2seg_plan* p, int phase, int op, const void* desc, size_t desc_bytes) {
  b2::g_err[0] = 0;
  if (!p || phase < 0 || phase > 2 || !desc) return b2::fail(B2SEG_ERR_ARG, "plan_add: bad arguments");
  if (phase == 1 && b2::g_bwd_sm_reserve > 0) {
    const int all = b2::num_sms();
    b2::g_sm_override = all - b2::g_bwd_sm_reserve > 16 ? all - b2::g_bwd_sm_reserve : 0;
  }
  b2::PreparedOp* po = b2::prepare_any(op, desc, desc_bytes);
  b2::g_sm_override = 0;
  if (!po) return b2::g_err[0] ? B2SEG_ERR_ARG : b2::fail(B2SEG_ERR_ARG, "plan_add: prepare failed for op %d", op);
  p->phase[phase].emplace_back(po);
  return 0;
}

int b2seg_plan_run(b2seg_plan* p, int phase, void* stream) {
  if (!p || phase < 0 || phase > 2) return b2::fail(B2SEG_ERR_ARG, "plan_run: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  for (auto& op : p->phase[phase]) {
    int rc = op->launch(s);
    if (rc) return rc;
  }
  return 0;
}

int b2seg_plan_run_range(b2seg_plan* p, int phase, int first_op, int n_ops, void* stream) {
  if (!p || phase < 0 || phase > 2) return b2::fail(B2SEG_ERR_ARG, "plan_run_range: bad arguments");
  const int n = (int)p->phase[phase].size();
  if (first_op < 0 || n_ops < 0 || first_op + n_ops > n) return b2::fail(B2SEG_ERR_ARG, "plan_run_range: ops [%d, %d) outside [0, %d)", first_op, first_op + n_ops, n);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  for (int i = first_op; i < first_op + n_ops; ++i) {
    int rc = p->phase[phase][i]->launch(s);
    if (rc) return rc;
  }
  return 0;
}

int b2seg_plan_run_timed(b2seg_plan* p, int phase, void* stream, float* ms_per_op, int n_ops) {
  if (!p || phase < 0 || phase > 2 || !ms_per_op) return b2::fail(B2SEG_ERR_ARG, "plan_run_timed: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int n = (int)p->phase[phase].size();
  if (n_ops < n) return b2::fail(B2SEG_ERR_ARG, "plan_run_timed: need room for %d ops", n);
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) B2_CUDA_OK(cudaEventCreate(&e));
  B2_CUDA_OK(cudaEventRecord(ev[0], s));
  for (int i = 0; i < n; ++i) {
    int rc = p->phase[phase][i]->launch(s);
    if (rc) return rc;
    B2_CUDA_OK(cudaEventRecord(ev[i + 1], s));
  }
  B2_CUDA_OK(cudaEventSynchronize(ev[n]));
  for (int i = 0; i < n; ++i) B2_CUDA_OK(cudaEventElapsedTime(&ms_per_op[i], ev[i], ev[i + 1]));
  for (auto& e : ev) cudaEventDestroy(e);
  return 0;
}

int b2seg_plan_num_ops(const b2seg_plan* p, int phase) {
  if (!p || phase < 0 || phase > 2) return -1;
  return (int)p->phase[phase].size();
}

int b2seg_plan_num_launches(const b2seg_plan* p, int phase) {
  if (!p || phase < 0 || phase > 2) return -1;
  int n = 0;
  for (auto& op : p->phase[phase]) n += op->num_launches();
  return n;
}

int b2seg_plan_set_adam(b2seg_plan* p, float lr, int64_t step, float grad_scale) {
  if (!p) return b2::fail(B2SEG_ERR_ARG, "null plan");
  for (int ph = 0; ph < 3; ++ph)
    for (auto& op : p->phase[ph])
      if (b2::is_adam(op.get())) b2::adam_update(op.get(), lr, step, grad_scale);
  return 0;
}

void b2seg_plan_destroy(b2seg_plan* p) { delete p; }

}  // extern "C"
