"""ctypes binding of libb2seg.so — the C ABI declared in include/b2seg.h.

The structures below mirror include/b2seg.h field for field.  Loading fails loudly when the shared library
has not been built (``python __graft_entry__.py`` / ``make -C tf-1d-2d-segmentation-end2endpipelines_b200``):
there is no CPU or eager fallback for the hot path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(PKG_ROOT, "libb2seg.so")

MAX_SRC, MAX_TAPS, MAX_GROUPS, MAX_GRADSRC = 4, 16, 4, 6
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID, ACT_SOFTMAX, ACT_TANH = 0, 1, 2, 3, 4, 5
ACT_CODES = {None: ACT_NONE, "linear": ACT_NONE, "relu": ACT_RELU, "ReLU": ACT_RELU, "LeakyReLU": ACT_LEAKY,
             "sigmoid": ACT_SIGMOID, "softmax": ACT_SOFTMAX, "tanh": ACT_TANH}
(OP_CONV, OP_WGRAD, OP_BN_FINALIZE, OP_BN_ACT, OP_BN_BWD, OP_ADAM, OP_HEAD_FWD, OP_HEAD_BWD, OP_LOSS, OP_ELTWISE,
 OP_CAST, OP_COLSUM, OP_MEMSET, OP_RESIZE_FWD, OP_RESIZE_BWD, OP_MULBC_FWD, OP_MULBC_BWD, OP_COLSTATS, OP_LSTM_FWD,
 OP_LSTM_BWD, OP_POOL_BWD, OP_ROWSUM, OP_OUTACT_FWD, OP_OUTACT_BWD, OP_TARGET_POOL, OP_GATE_FWD, OP_GATE_BWD, OP_FOLD_BN) = range(1, 29)
PHASE_FWD, PHASE_BWD, PHASE_OPT = 0, 1, 2
ABI_VERSION = 103    # b2seg_version() of the library this binding mirrors (include/b2seg.h)


class View(C.Structure):
    _fields_ = [("ptr", C.c_uint64), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("sn", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64)]


class Tap(C.Structure):
    _fields_ = [("src", C.c_int32), ("dh", C.c_int32), ("dw", C.c_int32), ("widx", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [("n_src", C.c_int32), ("src", View * MAX_SRC), ("weights", C.c_uint64),
                ("w_cout", C.c_int32), ("w_taps", C.c_int32), ("w_cin", C.c_int32), ("b_mn_major", C.c_int32),
                ("n_groups", C.c_int32), ("taps_per_group", C.c_int32), ("taps", Tap * MAX_TAPS),
                ("out", View * MAX_GROUPS), ("bias", C.c_uint64), ("act", C.c_int32), ("stats", C.c_uint64),
                ("mul_src", C.c_uint64), ("mul_view", View), ("mul_mode", C.c_int32), ("block_n", C.c_int32),
                ("stats_atomic", C.c_int32)]


class WgradTap(C.Structure):
    _fields_ = [("pair", C.c_int32), ("dyh", C.c_int32), ("dyw", C.c_int32), ("dh", C.c_int32), ("dw", C.c_int32),
                ("widx", C.c_int32)]


class WgradDesc(C.Structure):
    _fields_ = [("n_pair", C.c_int32), ("dy", View * MAX_SRC), ("x", View * MAX_SRC),
                ("gN", C.c_int32), ("gH", C.c_int32), ("gW", C.c_int32), ("n_taps", C.c_int32),
                ("taps", WgradTap * MAX_TAPS), ("dw", C.c_uint64),
                ("w_cout", C.c_int32), ("w_taps", C.c_int32), ("w_cin", C.c_int32),
                ("ksplit", C.c_int32), ("accumulate", C.c_int32)]


class BnFinalizeDesc(C.Structure):
    _fields_ = [("partials", C.c_uint64), ("n_partials", C.c_int32), ("C", C.c_int32), ("count", C.c_double),
                ("gamma", C.c_uint64), ("beta", C.c_uint64), ("moving_mean", C.c_uint64), ("moving_var", C.c_uint64),
                ("update_moving", C.c_int32), ("bessel", C.c_int32), ("eps", C.c_float), ("momentum", C.c_float),
                ("scale", C.c_uint64), ("shift", C.c_uint64), ("mean", C.c_uint64), ("rstd", C.c_uint64),
                ("inference", C.c_int32)]


class BnActDesc(C.Structure):
    _fields_ = [("x", View), ("scale", C.c_uint64), ("shift", C.c_uint64), ("act", C.c_int32), ("n_out", C.c_int32),
                ("out", View * 2), ("pool_h", C.c_int32), ("pool_w", C.c_int32), ("pooled", View), ("c_valid", C.c_int32),
                ("add", View), ("out_stats", C.c_uint64), ("out_stats_pitch", C.c_int32)]


class GradSrc(C.Structure):
    _fields_ = [("g", View), ("kind", C.c_int32), ("pool_h", C.c_int32), ("pool_w", C.c_int32),
                ("dlogits", C.c_uint64), ("head_w", C.c_uint64), ("head_dw", C.c_uint64), ("head_db", C.c_uint64), ("cout", C.c_int32)]


class BnBwdDesc(C.Structure):
    _fields_ = [("x", View), ("scale", C.c_uint64), ("shift", C.c_uint64), ("mean", C.c_uint64), ("rstd", C.c_uint64),
                ("act", C.c_int32), ("n_src", C.c_int32), ("src", GradSrc * MAX_GRADSRC), ("count", C.c_double),
                ("partials", C.c_uint64), ("n_blocks", C.c_int32), ("dgamma", C.c_uint64), ("dbeta", C.c_uint64),
                ("dx", View), ("accumulate", C.c_int32), ("x_relu_mask", C.c_int32)]


class AdamDesc(C.Structure):
    _fields_ = [("w", C.c_uint64), ("g", C.c_uint64), ("m", C.c_uint64), ("v", C.c_uint64), ("w_bf16", C.c_uint64),
                ("n", C.c_int64), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("grad_scale", C.c_float), ("step", C.c_int64)]


class HeadDesc(C.Structure):
    _fields_ = [("x", View), ("w", C.c_uint64), ("b", C.c_uint64), ("cout", C.c_int32), ("act", C.c_int32),
                ("stride", C.c_int32), ("y", C.c_uint64), ("dlogits", C.c_uint64), ("dx", View),
                ("dw", C.c_uint64), ("db", C.c_uint64), ("logits", C.c_uint64),
                ("bn_scale", C.c_uint64), ("bn_shift", C.c_uint64), ("bn_act", C.c_int32)]


class LossDesc(C.Structure):
    _fields_ = [("y_pred", C.c_uint64), ("y_true", C.c_uint64), ("n_pix", C.c_int64), ("cout", C.c_int32),
                ("kind", C.c_int32), ("act", C.c_int32), ("weight", C.c_float), ("dlogits", C.c_uint64),
                ("loss", C.c_uint64), ("metrics", C.c_uint64)]


class EltwiseDesc(C.Structure):
    _fields_ = [("op", C.c_int32), ("a", View), ("b", View), ("c", View), ("out", View), ("act", C.c_int32)]


class ResizeDesc(C.Structure):
    _fields_ = [("x", View), ("y", View), ("yfwd", View), ("fh", C.c_int32), ("fw", C.c_int32), ("mode", C.c_int32),
                ("act", C.c_int32), ("c_valid", C.c_int32), ("n_vseg", C.c_int32), ("vseg_off", C.c_int32 * 8), ("vseg_cnt", C.c_int32 * 8)]


class MulbcDesc(C.Structure):
    _fields_ = [("a", View), ("b", View), ("out", View), ("dout", View), ("da", View), ("db", View)]


class ColstatsDesc(C.Structure):
    _fields_ = [("x", View), ("partials", C.c_uint64), ("n_blocks", C.c_int32)]


class LstmDesc(C.Structure):
    _fields_ = [("z", View), ("h", View), ("dh", View), ("dz", View), ("F", C.c_int32)]


class PoolBwdDesc(C.Structure):
    _fields_ = [("y", View), ("dp", View), ("dx", View), ("ph", C.c_int32), ("pw", C.c_int32)]


class CastDesc(C.Structure):
    _fields_ = [("src", C.c_uint64), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("out", View), ("kh", C.c_int32), ("kw", C.c_int32)]


class ColsumDesc(C.Structure):
    _fields_ = [("g", View), ("out", C.c_uint64), ("scratch", C.c_uint64), ("n_blocks", C.c_int32)]


class RowsumDesc(C.Structure):
    _fields_ = [("partials", C.c_uint64), ("n_rows", C.c_int32), ("pitch", C.c_int32), ("C", C.c_int32),
                ("out", C.c_uint64), ("accumulate", C.c_int32)]


class OutActDesc(C.Structure):
    _fields_ = [("x", View), ("cout", C.c_int32), ("act", C.c_int32), ("y", C.c_uint64), ("dlogits", C.c_uint64), ("dx", View)]


class TPoolDesc(C.Structure):
    _fields_ = [("src", C.c_uint64), ("dst", C.c_uint64), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("ph", C.c_int32), ("pw", C.c_int32), ("mode", C.c_int32)]


class GateDesc(C.Structure):
    _fields_ = [("za", View), ("zb", View), ("sums_a", C.c_uint64), ("sums_b", C.c_uint64),
                ("gamma_a", C.c_uint64), ("beta_a", C.c_uint64), ("mm_a", C.c_uint64), ("mv_a", C.c_uint64),
                ("gamma_b", C.c_uint64), ("beta_b", C.c_uint64), ("mm_b", C.c_uint64), ("mv_b", C.c_uint64),
                ("vec_a", C.c_uint64), ("vec_b", C.c_uint64), ("w3", C.c_uint64), ("b3", C.c_uint64), ("z", C.c_uint64), ("m", C.c_uint64), ("sums3", C.c_uint64),
                ("gamma3", C.c_uint64), ("beta3", C.c_uint64), ("mm3", C.c_uint64), ("mv3", C.c_uint64), ("wt", C.c_uint64), ("bt", C.c_uint64),
                ("wt_stride", C.c_int32), ("training", C.c_int32), ("bessel", C.c_int32), ("eps", C.c_float), ("momentum", C.c_float),
                ("count", C.c_double), ("skip", View), ("out", View), ("dout", View), ("dskip", View), ("dr", C.c_uint64), ("g3", C.c_uint64),
                ("bsums3", C.c_uint64), ("bsums_ab", C.c_uint64), ("dza", View), ("dzb", View),
                ("dgamma_a", C.c_uint64), ("dbeta_a", C.c_uint64), ("dgamma_b", C.c_uint64), ("dbeta_b", C.c_uint64),
                ("dgamma3", C.c_uint64), ("dbeta3", C.c_uint64), ("dw3", C.c_uint64), ("db3", C.c_uint64), ("dwt", C.c_uint64), ("dbt", C.c_uint64),
                ("da_low", View)]


class FoldDesc(C.Structure):
    _fields_ = [("w", C.c_uint64), ("bias", C.c_uint64), ("gamma", C.c_uint64), ("beta", C.c_uint64), ("moving_mean", C.c_uint64),
                ("moving_var", C.c_uint64), ("eps", C.c_float), ("cout_p", C.c_int32), ("row", C.c_int32), ("w_folded", C.c_uint64),
                ("bias_folded", C.c_uint64)]


class MemsetDesc(C.Structure):
    _fields_ = [("ptr", C.c_uint64), ("bytes", C.c_int64)]


OP_DESC = {OP_CONV: ConvDesc, OP_WGRAD: WgradDesc, OP_BN_FINALIZE: BnFinalizeDesc, OP_BN_ACT: BnActDesc,
           OP_BN_BWD: BnBwdDesc, OP_ADAM: AdamDesc, OP_HEAD_FWD: HeadDesc, OP_HEAD_BWD: HeadDesc, OP_LOSS: LossDesc,
           OP_ELTWISE: EltwiseDesc, OP_CAST: CastDesc, OP_COLSUM: ColsumDesc, OP_MEMSET: MemsetDesc,
           OP_RESIZE_FWD: ResizeDesc, OP_RESIZE_BWD: ResizeDesc, OP_MULBC_FWD: MulbcDesc, OP_MULBC_BWD: MulbcDesc,
           OP_COLSTATS: ColstatsDesc, OP_LSTM_FWD: LstmDesc, OP_LSTM_BWD: LstmDesc, OP_POOL_BWD: PoolBwdDesc,
           OP_ROWSUM: RowsumDesc, OP_OUTACT_FWD: OutActDesc, OP_OUTACT_BWD: OutActDesc,
           OP_TARGET_POOL: TPoolDesc, OP_GATE_FWD: GateDesc, OP_GATE_BWD: GateDesc, OP_FOLD_BN: FoldDesc}

# every symbol include/b2seg.h declares (the CPU test-suite checks the library exports all of them)
EXPORTED = ["b2seg_last_error", "b2seg_version", "b2seg_device_check", "b2seg_sizeof_desc", "b2seg_conv", "b2seg_conv_num_mtiles", "b2seg_conv_num_stat_rows",
            "b2seg_wgrad", "b2seg_bn_finalize", "b2seg_bn_act", "b2seg_bn_bwd", "b2seg_adam", "b2seg_head_fwd",
            "b2seg_head_bwd", "b2seg_loss", "b2seg_eltwise", "b2seg_cast_input", "b2seg_colsum", "b2seg_plan_create",
            "b2seg_resize_fwd", "b2seg_resize_bwd", "b2seg_mulbc_fwd", "b2seg_mulbc_bwd", "b2seg_colstats", "b2seg_lstm_fwd",
            "b2seg_lstm_bwd", "b2seg_pool_bwd", "b2seg_rowsum", "b2seg_outact_fwd", "b2seg_outact_bwd", "b2seg_target_pool", "b2seg_gate_fwd", "b2seg_gate_bwd", "b2seg_fold_bn",
            "b2seg_set_backward_sm_reserve", "b2seg_debug_read_trace", "b2seg_plan_add", "b2seg_plan_run", "b2seg_plan_run_range", "b2seg_plan_num_launches", "b2seg_plan_num_ops", "b2seg_plan_run_timed", "b2seg_plan_set_adam", "b2seg_plan_destroy"]

_lib = None


class B2SegError(RuntimeError):
    pass


def load():
    """Load libb2seg.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2SegError(f"{LIB_PATH} not built: run `python __graft_entry__.py` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    have = lib.b2seg_version() if hasattr(lib, "b2seg_version") else -1
    if have != ABI_VERSION:
        raise B2SegError(f"{LIB_PATH} implements ABI {have}, this binding needs {ABI_VERSION}: rebuild it (`python __graft_entry__.py`)")
    lib.b2seg_last_error.restype = C.c_char_p
    for name, desc in [("b2seg_conv", ConvDesc), ("b2seg_wgrad", WgradDesc), ("b2seg_bn_finalize", BnFinalizeDesc),
                       ("b2seg_bn_act", BnActDesc), ("b2seg_bn_bwd", BnBwdDesc), ("b2seg_adam", AdamDesc),
                       ("b2seg_head_fwd", HeadDesc), ("b2seg_head_bwd", HeadDesc), ("b2seg_loss", LossDesc),
                       ("b2seg_eltwise", EltwiseDesc), ("b2seg_cast_input", CastDesc), ("b2seg_colsum", ColsumDesc),
                       ("b2seg_resize_fwd", ResizeDesc), ("b2seg_resize_bwd", ResizeDesc), ("b2seg_mulbc_fwd", MulbcDesc),
                       ("b2seg_mulbc_bwd", MulbcDesc), ("b2seg_colstats", ColstatsDesc), ("b2seg_lstm_fwd", LstmDesc),
                       ("b2seg_lstm_bwd", LstmDesc), ("b2seg_pool_bwd", PoolBwdDesc), ("b2seg_rowsum", RowsumDesc),
                       ("b2seg_outact_fwd", OutActDesc), ("b2seg_outact_bwd", OutActDesc), ("b2seg_target_pool", TPoolDesc),
                       ("b2seg_gate_fwd", GateDesc), ("b2seg_gate_bwd", GateDesc), ("b2seg_fold_bn", FoldDesc)]:
        fn = getattr(lib, name)
        fn.argtypes = [C.POINTER(desc), C.c_void_p]
        fn.restype = C.c_int
    lib.b2seg_conv_num_mtiles.argtypes = [C.POINTER(ConvDesc)]
    lib.b2seg_conv_num_mtiles.restype = C.c_int
    lib.b2seg_conv_num_stat_rows.argtypes = [C.POINTER(ConvDesc)]
    lib.b2seg_conv_num_stat_rows.restype = C.c_int
    lib.b2seg_device_check.argtypes = [C.c_int]
    lib.b2seg_set_backward_sm_reserve.argtypes = [C.c_int]
    lib.b2seg_plan_create.argtypes = [C.POINTER(C.c_void_p)]
    lib.b2seg_plan_add.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    lib.b2seg_plan_run.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.b2seg_plan_run_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.b2seg_plan_num_launches.argtypes = [C.c_void_p, C.c_int]
    lib.b2seg_plan_num_ops.argtypes = [C.c_void_p, C.c_int]
    lib.b2seg_plan_run_timed.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_int]
    lib.b2seg_plan_set_adam.argtypes = [C.c_void_p, C.c_float, C.c_int64, C.c_float]
    lib.b2seg_plan_destroy.argtypes = [C.c_void_p]
    lib.b2seg_plan_destroy.restype = None
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().b2seg_last_error().decode("utf-8", "replace")
        raise B2SegError(f"b2seg {what} failed (rc={rc}): {msg}")


def call(name, desc, stream=0):
    lib = load()
    check(getattr(lib, name)(C.byref(desc), C.c_void_p(stream)), name)


def view_of(t, c_off=0, C_=None):
    """b2seg view of a torch NHWC bf16 tensor (N,H,W,Ctot), optionally a channel window."""
    N, H, W, Ct = t.shape
    assert t.stride(3) == 1
    return View(t.data_ptr() + 2 * c_off, N, H, W, Ct - c_off if C_ is None else C_, t.stride(0), t.stride(1), t.stride(2))
