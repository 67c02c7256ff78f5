"""ctypes binding of the C-ABI in ``include/gs2m_rasterizer.h`` (libgs2m_rasterizer.so, sm_100a CUDA).

This is the only place the shared library is touched.  There is deliberately no CPU or PyTorch fallback: if the
library is missing or a call fails the error is raised to the caller.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GS2M_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libgs2m_rasterizer.so")

ABI_VERSION = 3
NUM_CHANNELS = 3
NUM_FEATURES = 10
ACC_STRIDE = 24

RESIZE_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)

_fp = C.c_void_p  # device pointers travel as integers


class ForwardArgs(C.Structure):
    _fields_ = [
        ("geometry_buffer", RESIZE_FN), ("geometry_user", C.c_void_p),
        ("binning_buffer", RESIZE_FN), ("binning_user", C.c_void_p),
        ("image_buffer", RESIZE_FN), ("image_user", C.c_void_p),
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int),
        ("background", _fp),
        ("width", C.c_int), ("height", C.c_int),
        ("means3D", _fp), ("shs", _fp), ("colors_precomp", _fp), ("opacities", _fp), ("scales", _fp),
        ("scale_modifier", C.c_float),
        ("rotations", _fp), ("cov3D_precomp", _fp), ("features", _fp),
        ("viewmatrix", _fp), ("projmatrix", _fp), ("cam_pos", _fp),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
        ("prefiltered", C.c_int), ("feature_count", C.c_int),
        ("out_color", _fp), ("out_radii", _fp), ("out_observe", _fp), ("out_buffer", _fp),
        ("stream", C.c_void_p),
        ("R_capacity", C.c_int), ("no_wait", C.c_int), ("no_backward", C.c_int),
    ]


class ParamChain(C.Structure):
    _fields_ = [("scaling_raw", _fp), ("rotation_raw", _fp), ("opacity_raw", _fp), ("albedo_raw", _fp), ("roughness_raw", _fp),
                ("metallic_raw", _fp), ("z_depth", C.c_int), ("blend_metallic", C.c_int),
                ("d_xyz", _fp), ("d_scaling_raw", _fp), ("d_rotation_raw", _fp), ("d_opacity_raw", _fp), ("d_albedo_raw", _fp),
                ("d_roughness_raw", _fp), ("d_metallic_raw", _fp)]


class BackwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("D", C.c_int), ("M", C.c_int), ("R", C.c_int), ("R_capacity", C.c_int),
        ("background", _fp),
        ("width", C.c_int), ("height", C.c_int),
        ("means3D", _fp), ("shs", _fp), ("colors_precomp", _fp), ("scales", _fp),
        ("scale_modifier", C.c_float),
        ("rotations", _fp), ("cov3D_precomp", _fp), ("features", _fp),
        ("viewmatrix", _fp), ("projmatrix", _fp), ("cam_pos", _fp),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
        ("radii", _fp),
        ("geometry_buffer", _fp), ("binning_buffer", _fp), ("image_buffer", _fp),
        ("geometry_bytes", C.c_size_t), ("binning_bytes", C.c_size_t), ("image_bytes", C.c_size_t),
        ("feature_count", C.c_int),
        ("grad_color", _fp), ("grad_buffer", _fp),
        ("dL_dmeans2D", _fp), ("dL_dconic", _fp), ("dL_dopacity", _fp), ("dL_dcolor", _fp), ("dL_dmeans3D", _fp),
        ("dL_dcov3D", _fp), ("dL_dsh", _fp), ("dL_dscale", _fp), ("dL_drot", _fp), ("dL_dfeatures", _fp),
        ("accumulate", C.c_int),
        ("stream", C.c_void_p),
        ("phase", C.c_int), ("row_begin", C.c_int), ("row_end", C.c_int),
        ("grad_acc_dirty", C.c_int),
        ("densify_grad_accum", _fp), ("densify_grad_accum_abs", _fp), ("densify_denom", _fp),
        ("chain", C.POINTER(ParamChain)),
    ]


class AdamGroup(C.Structure):
    _fields_ = [("param", _fp), ("exp_avg", _fp), ("exp_avg_sq", _fp), ("grad", _fp), ("rows", C.c_longlong),
                ("width", C.c_int), ("grad_row_stride", C.c_int), ("grad_col_offset", C.c_int), ("lr", C.c_float)]


class StateView(C.Structure):
    _fields_ = [(n, _fp) for n in (
        "depths", "rec_a", "rec_b", "rgb", "cov3D", "clamped", "tiles_touched", "point_offsets", "grad_acc",
        "keys_sorted", "point_list", "masks", "dense_gid", "dense_pos", "final_T", "n_contrib", "ranges", "bin_info",
        "block_ranges", "n_contrib_dense")]


# every symbol include/gs2m_rasterizer.h declares: (name, restype, argtypes)
EXPORTS = [
    ("gs2m_abi_version", C.c_int, []),
    ("gs2m_last_error", C.c_char_p, []),
    ("gs2m_rasterize_forward", C.c_int, [C.POINTER(ForwardArgs)]),
    ("gs2m_rasterize_backward", C.c_int, [C.POINTER(BackwardArgs)]),
    ("gs2m_rasterize_backward_views", C.c_int, [C.POINTER(BackwardArgs), C.c_int]),
    ("gs2m_last_instance_count", C.c_longlong, []),
    ("gs2m_geometry_bytes", C.c_size_t, [C.c_int]),
    ("gs2m_image_bytes", C.c_size_t, [C.c_int, C.c_int]),
    ("gs2m_binning_bytes", C.c_size_t, [C.c_int]),
    ("gs2m_mark_visible", C.c_int, [C.c_int, _fp, _fp, _fp, _fp, C.c_void_p]),
    ("gs2m_state_view_get", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, C.POINTER(StateView)]),
    ("gs2m_sort_temp_bytes", C.c_size_t, [C.c_int]),
    ("gs2m_sort_pairs_u64", C.c_int, [_fp, _fp, _fp, _fp, C.c_int, C.c_int, _fp, C.c_void_p]),
    ("gs2m_scan_temp_bytes", C.c_size_t, [C.c_int]),
    ("gs2m_inclusive_sum_u32", C.c_int, [_fp, _fp, C.c_int, _fp, C.c_void_p]),
    ("gs2m_pack_forward", C.c_int, [C.c_int] + [_fp] * 9 + [C.c_int, C.c_int] + [_fp] * 4 + [C.c_void_p]),
    ("gs2m_pack_backward", C.c_int, [C.c_int] + [_fp] * 9 + [C.c_int, C.c_int] + [_fp] * 11 + [C.c_void_p]),
    ("gs2m_pack_backward_accumulate", C.c_int, [C.c_int] + [_fp] * 9 + [C.c_int, C.c_int] + [_fp] * 12 + [C.c_void_p]),
    ("gs2m_postblend_forward", C.c_int, [C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_int] + [_fp] * 5 + [C.c_void_p]),
    ("gs2m_postblend_backward", C.c_int, [C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_int] + [_fp] * 5 + [C.c_void_p]),
    ("gs2m_sobel_normal_forward", C.c_int, [C.c_int, C.c_int] + [C.c_float] * 4 + [_fp] * 5 + [C.c_void_p]),
    ("gs2m_sobel_normal_backward", C.c_int, [C.c_int, C.c_int] + [C.c_float] * 4 + [_fp] * 7 + [C.c_void_p]),
    ("gs2m_photometric_loss_forward", C.c_int, [C.c_int] * 3 + [_fp] * 6 + [C.c_void_p]),
    ("gs2m_photometric_loss_backward", C.c_int, [C.c_int] * 3 + [_fp] * 5 + [C.c_float, C.c_float, _fp, _fp, C.c_void_p]),
    ("gs2m_adam_step", C.c_int, [C.POINTER(AdamGroup), C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]),
    ("gs2m_view_stats_update", C.c_int, [C.c_int, _fp, _fp, _fp, _fp, C.c_void_p]),
    ("gs2m_profile_enable", None, [C.c_int]),
    ("gs2m_profile_read", C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    ("gs2m_launch_count", C.c_longlong, []),
]

STAGES = ("preprocess_fwd", "scan", "duplicate", "sort", "ranges", "blend_fwd", "blend_bwd", "preprocess_bwd")


def profile_read():
    """{stage: (total_ms, calls)} accumulated since the previous read (requires gs2m_profile_enable(1))."""
    ms = (C.c_float * len(STAGES))()
    calls = (C.c_int * len(STAGES))()
    load().gs2m_profile_read(ms, calls)
    return {n: (float(ms[i]), int(calls[i])) for i, n in enumerate(STAGES)}

_lib = None


def load():
    """Load the shared library once; raises ImportError when it has not been built (run ``gs-2m_b200/build.py``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libgs2m_rasterizer.so not found at %s — build it with `python gs-2m_b200/build.py` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in EXPORTS:
        fn = getattr(lib, name)  # AttributeError if the header and the library diverge
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.gs2m_abi_version() != ABI_VERSION:
        raise ImportError("libgs2m_rasterizer.so ABI version mismatch")
    _lib = lib
    return lib


ERR_CAPACITY = -5   # GS2M_ERR_CAPACITY: a speculative forward saw more instances than its R_capacity
BIN_OVERFLOW, BIN_PREFILTERED, BIN_TOO_LARGE = 1, 2, 4


class RasterizerError(RuntimeError):
    code = 0


def check(rc, what):
    if rc < 0:
        msg = load().gs2m_last_error().decode("utf-8", "replace")
        err = RasterizerError("%s failed (code %d): %s" % (what, rc, msg))
        err.code = rc
        raise err
    return rc
