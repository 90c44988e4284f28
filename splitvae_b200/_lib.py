"""ctypes binding of libsplitvae.so (include/splitvae.h).  There is no fallback: if the CUDA
library has not been built (``python splitvae_b200/build.py`` / ``__graft_entry__.build()``),
importing the product path fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsplitvae.so")

SV_OK = 0
SV_MODEL = {"lgvae": 0, "lggmvae": 1, "gmvae": 2}
SV_PRECISION = {"bf16": 0, "fp32": 1, "bf16x3": 2}
SV_FLAG_PLAN_ONLY = 1
SV_FLAG_NO_TC = 2

OUT_NAMES = ["dec_x", "dec_x_hat", "z_x", "z_mean_x", "z_sig_x", "z_x_hat", "z_mean_x_hat", "z_sig_x_hat",
             "y", "y_logits", "z_prior_mean", "z_prior_sig", "scalars", "scalar_sums"]
SCALAR_NAMES = ["recon_x", "recon_x_hat", "kl_x", "kl_x_hat", "total_kl_or_y_kl", "total"]

# every symbol include/splitvae.h declares (checked by tests/test_abi.py)
SYMBOLS = ["sv_create", "sv_destroy", "sv_last_error", "sv_version", "sv_param_count", "sv_param_describe",
           "sv_arena_floats", "sv_workspace_bytes", "sv_bind", "sv_params_updated", "sv_forward", "sv_loss_fwd_bwd",
           "sv_num_segments", "sv_segment_num_ranges", "sv_segment_range", "sv_backward_segment", "sv_backward_segment_deferred", "sv_adam_step", "sv_adam_segment", "sv_nvls_adam_segment", "sv_repack_segment", "sv_train_step",
           "sv_capture_graph", "sv_replay", "sv_output_ptr", "sv_decode", "sv_encode_y", "sv_get_iterations", "sv_set_iterations", "sv_launch_count",
           "sv_discretised_logistic_loss", "sv_adam_flat", "sv_stage_scramble", "sv_stage_resize_scramble", "sv_draw_permutations", "sv_debug_layer_count",
           "sv_debug_layer_info", "sv_debug_run_layer", "sv_debug_pixel_loss", "sv_debug_halo_trace"]


class SvConfig(C.Structure):
    _fields_ = [("model", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("batch", C.c_int32),
                ("global_latent_dims", C.c_int32), ("local_latent_dims", C.c_int32), ("y_size", C.c_int32),
                ("tau", C.c_float), ("beta", C.c_float), ("alpha", C.c_float), ("learning_rate", C.c_float),
                ("world_size", C.c_int32), ("precision", C.c_int32), ("flags", C.c_int32)]


class SvParamDesc(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("ndim", C.c_int32), ("shape", C.c_int32 * 4),
                ("offset", C.c_int64), ("count", C.c_int64)]


class SvLayerInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64)] + [(n, C.c_int32) for n in
                ("kh", "kw", "stride", "Hi", "Wi", "Ci", "Ho", "Wo", "Co", "in_ld", "in_coff", "out_ld", "dout_ld", "din_ld",
                 "in_dt", "out_dt", "act_dt", "has_dgrad", "tc_fwd", "tc_dgrad", "tc_wgrad")] + \
               [(n, C.c_void_p) for n in ("in_", "out", "dout", "din")] + \
               [(n, C.c_int64) for n in ("in_elems", "out_elems", "dout_elems", "din_elems")] + \
               [(n, C.c_int32) for n in ("kern_fwd", "kern_dgrad", "kern_wgrad", "split_fwd")] + \
               [(n, C.c_void_p) for n in ("in_lo", "out_lo")] + \
               [(n, C.c_int32) for n in ("wgrad_ctas", "reserved_")]


KERNEL_NAMES = {0: "reference", 1: "igemm_kernel", 2: "halo_conv_kernel", 3: "nsconv_kernel", 4: "wgrad_kernel", 5: "halo_wgrad_kernel", 6: "pconv_kernel"}


class SplitVaeError(RuntimeError):
    pass


_lib = None


def load():
    """Loads libsplitvae.so once and declares the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SplitVaeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "splitvae_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.sv_create.argtypes = [C.POINTER(SvConfig), C.POINTER(vp)]
    lib.sv_destroy.argtypes = [vp]
    lib.sv_last_error.argtypes = [vp]
    lib.sv_last_error.restype = C.c_char_p
    lib.sv_version.restype = C.c_char_p
    lib.sv_param_count.argtypes = [vp]
    lib.sv_param_describe.argtypes = [vp, i32, C.POINTER(SvParamDesc)]
    lib.sv_arena_floats.argtypes = [vp]
    lib.sv_arena_floats.restype = i64
    lib.sv_workspace_bytes.argtypes = [vp]
    lib.sv_workspace_bytes.restype = i64
    lib.sv_bind.argtypes = [vp, vp, vp, vp, vp, vp, i64]
    lib.sv_params_updated.argtypes = [vp, vp]
    lib.sv_forward.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.sv_loss_fwd_bwd.argtypes = [vp, vp, vp]
    lib.sv_num_segments.argtypes = [vp]
    lib.sv_segment_num_ranges.argtypes = [vp, i32]
    lib.sv_segment_range.argtypes = [vp, i32, i32, C.POINTER(i64), C.POINTER(i64)]
    lib.sv_backward_segment.argtypes = [vp, i32, vp]
    lib.sv_backward_segment_deferred.argtypes = [vp, i32, vp, vp]
    lib.sv_adam_step.argtypes = [vp, vp]
    lib.sv_adam_segment.argtypes = [vp, i32, vp]
    lib.sv_nvls_adam_segment.argtypes = [vp, i32, vp, vp, i32, i32, i32, vp]
    lib.sv_repack_segment.argtypes = [vp, i32, vp]
    lib.sv_train_step.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.sv_capture_graph.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.sv_replay.argtypes = [vp, i32, vp]
    lib.sv_output_ptr.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i64)]
    lib.sv_decode.argtypes = [vp, vp, vp, vp]
    lib.sv_encode_y.argtypes = [vp, vp, vp]
    lib.sv_get_iterations.argtypes = [vp, C.POINTER(i64), vp]
    lib.sv_set_iterations.argtypes = [vp, i64, vp]
    lib.sv_launch_count.argtypes = [vp]
    lib.sv_launch_count.restype = i64
    lib.sv_discretised_logistic_loss.argtypes = [vp, vp, vp, vp, i64, vp]
    lib.sv_adam_flat.argtypes = [vp, vp, vp, vp, i64, f32, vp]
    lib.sv_stage_scramble.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    lib.sv_stage_resize_scramble.argtypes = [vp, vp, vp] + [i32] * 10 + [vp]
    lib.sv_draw_permutations.argtypes = [vp, i32, i32, C.c_uint64, C.c_uint64, vp]
    lib.sv_debug_layer_count.argtypes = [vp]
    lib.sv_debug_layer_info.argtypes = [vp, i32, C.POINTER(SvLayerInfo)]
    lib.sv_debug_run_layer.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.sv_debug_pixel_loss.argtypes = [vp, vp, vp]
    lib.sv_debug_halo_trace.argtypes = [vp, i32]
    lib.sv_debug_halo_trace.restype = i32
    _lib = lib
    return lib


def check(status: int, handle=None, what: str = ""):
    if status != SV_OK:
        msg = load().sv_last_error(handle)
        raise SplitVaeError(f"{what} failed (sv_status {status}): {msg.decode() if msg else ''}")
