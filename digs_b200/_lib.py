"""ctypes binding of libdigs_b200.so (C-ABI declared in include/digs_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  If it has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or `make -C digs_b200/csrc`) loading raises.
"""
import ctypes
import os

MAX_LAYERS = 18
MODE_FP32, MODE_BF16X2, MODE_BF16 = 0, 1, 2
MODES = {"fp32": MODE_FP32, "bf16x2": MODE_BF16X2, "bf16": MODE_BF16}
HEAD_IDENTITY, HEAD_SQRT = 0, 1

_c_float_p = ctypes.c_void_p  # device pointers travel as integers


class DigsNet(ctypes.Structure):
    _fields_ = [
        ("d", ctypes.c_int32),
        ("H", ctypes.c_int32),
        ("Lh", ctypes.c_int32),
        ("head", ctypes.c_int32),
        ("radius", ctypes.c_float),
        ("scale", ctypes.c_float),
        ("w0_stride", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("W", ctypes.c_void_p * MAX_LAYERS),
        ("b", ctypes.c_void_p * MAX_LAYERS),
        ("prepared", ctypes.c_void_p),
    ]


class DigsLossCfg(ctypes.Structure):
    _fields_ = [
        ("weights", ctypes.c_float * 6),
        ("weights_device", ctypes.c_void_p),
        ("add_normal", ctypes.c_int32),
        ("use_div", ctypes.c_int32),
        ("div_l2", ctypes.c_int32),
        ("div_clamp", ctypes.c_float),
        ("Nm_total", ctypes.c_int64),
        ("Nn_total", ctypes.c_int64),
    ]


class DigsGrads(ctypes.Structure):
    _fields_ = [
        ("dW", ctypes.c_void_p * MAX_LAYERS),
        ("db", ctypes.c_void_p * MAX_LAYERS),
    ]


EXPORTS = (
    "digs_abi_version", "digs_last_error", "digs_launch_count", "digs_saved_bytes", "digs_forward_workspace_bytes",
    "digs_backward_workspace_bytes", "digs_value_workspace_bytes", "digs_grid_workspace_bytes",
    "digs_prepared_bytes", "digs_prepare_weights", "digs_step_workspace_bytes", "digs_step_fused",
    "digs_jet_forward", "digs_jet_backward", "digs_loss_fused", "digs_loss_fused_dw", "digs_value_forward", "digs_grid_query",
    "digs_clip_adam", "digs_clip_adam_dev", "digs_sample_batch",
    "digs_mt_classify", "digs_mt_emit", "digs_nn_query",
)

# DIGS_B200_LIB: measurement-only override (diagnostic builds with one piece of a kernel compiled out, csrc/Makefile DIAG=)
LIB_PATH = os.environ.get("DIGS_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdigs_b200.so")
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "digs_b200: CUDA library %s is missing; build it with `make -C digs_b200/csrc` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    i64, i32, vp, sz = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t
    netp = ctypes.POINTER(DigsNet)
    L.digs_abi_version.restype = i32
    L.digs_last_error.restype = ctypes.c_char_p
    L.digs_launch_count.restype = ctypes.c_longlong
    L.digs_saved_bytes.restype = sz
    L.digs_saved_bytes.argtypes = [netp, i64, i32, i32]
    L.digs_forward_workspace_bytes.restype = sz
    L.digs_forward_workspace_bytes.argtypes = [netp, i64, i32, i32]
    L.digs_backward_workspace_bytes.restype = sz
    L.digs_backward_workspace_bytes.argtypes = [netp, i64, i32, i32]
    L.digs_value_workspace_bytes.restype = sz
    L.digs_value_workspace_bytes.argtypes = [netp, i64, i32]
    L.digs_grid_workspace_bytes.restype = sz
    L.digs_grid_workspace_bytes.argtypes = [netp, i32]
    L.digs_jet_forward.restype = i32
    L.digs_jet_forward.argtypes = [netp, vp, i64, i64, vp, i64, i32, vp, vp, vp, vp, sz, vp, sz, i32, vp]
    L.digs_jet_backward.restype = i32
    L.digs_jet_backward.argtypes = [netp, vp, i64, i64, vp, i64, i32, vp, vp, vp, vp, sz, vp, sz,
                                    ctypes.POINTER(DigsGrads), vp, i32, vp]
    L.digs_loss_fused.restype = i32
    L.digs_loss_fused.argtypes = [vp, vp, i64, vp, vp, vp, i64, vp, i32, ctypes.POINTER(ctypes.c_float), i32, i32,
                                  i32, ctypes.c_float, i64, i64, vp, vp, vp, vp, vp, vp, vp]
    L.digs_loss_fused_dw.restype = i32
    L.digs_loss_fused_dw.argtypes = [vp, vp, i64, vp, vp, vp, i64, vp, i32, vp, i32, i32,
                                     i32, ctypes.c_float, i64, i64, vp, vp, vp, vp, vp, vp, vp]
    L.digs_value_forward.restype = i32
    L.digs_value_forward.argtypes = [netp, vp, i64, i64, vp, i64, vp, vp, sz, i32, vp]
    L.digs_grid_query.restype = i32
    L.digs_grid_query.argtypes = [netp, vp, i32, vp, i32, vp, i32, i32, i32, vp, vp, vp, sz, i32, vp]
    L.digs_clip_adam.restype = i32
    L.digs_clip_adam.argtypes = [vp, vp, vp, vp, i64, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                 ctypes.c_float, i32, vp, vp]
    L.digs_clip_adam_dev.restype = i32
    L.digs_clip_adam_dev.argtypes = [vp, vp, vp, vp, i64, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                     ctypes.c_float, vp, vp, vp]
    L.digs_prepared_bytes.restype = sz
    L.digs_prepared_bytes.argtypes = [netp, i32]
    L.digs_prepare_weights.restype = i32
    L.digs_prepare_weights.argtypes = [netp, vp, sz, i32, vp]
    L.digs_step_workspace_bytes.restype = sz
    L.digs_step_workspace_bytes.argtypes = [netp, i64, i64, i32, i32]
    L.digs_step_fused.restype = i32
    L.digs_step_fused.argtypes = [netp, vp, i64, vp, i64, vp, i64, vp, vp, i64, i64, i32, ctypes.POINTER(DigsLossCfg),
                                  vp, vp, vp, vp, vp, vp, ctypes.POINTER(DigsGrads), vp, vp, vp, sz, i32, vp]
    L.digs_sample_batch.restype = i32
    L.digs_sample_batch.argtypes = [vp, vp, i64, i32, i64, ctypes.c_float, ctypes.c_uint64, vp, vp, vp, vp, vp]
    L.digs_mt_classify.restype = i32
    L.digs_mt_classify.argtypes = [vp, i32, i32, i32, ctypes.c_float, vp, vp, vp]
    L.digs_mt_emit.restype = i32
    L.digs_mt_emit.argtypes = [vp, i32, i32, i32, ctypes.c_float, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_float),
                               ctypes.POINTER(ctypes.c_float), vp, vp, vp, vp, vp, vp, vp, vp]
    L.digs_nn_query.restype = i32
    L.digs_nn_query.argtypes = [vp, i64, vp, i64, i32, vp, vp, vp]
    if L.digs_abi_version() != 2:
        raise RuntimeError("digs_b200: ABI version mismatch")
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().digs_last_error().decode("utf-8", "replace")
        raise RuntimeError("digs_b200.%s failed (code %d): %s" % (what, rc, msg))
