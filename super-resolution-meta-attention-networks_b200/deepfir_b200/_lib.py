"""ctypes loader + prototypes for libdfir_b200.so (C ABI: include/dfir.h)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
_LIB = None


class DfirError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_PKG, "libdfir_b200.so")


def build_library(verbose=False):
    """Compile csrc/*.cu for sm_100a into the in-tree libdfir_b200.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_PKG, "csrc"), "-j4"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise DfirError("building libdfir_b200.so failed (exit %d)" % res.returncode)
    return lib_path()


class QrcanNet(C.Structure):
    """mirror of `dfir_qrcan_net` (include/dfir.h) — field order must match exactly."""
    _fields_ = [
        ("n_groups", C.c_int), ("n_blocks", C.c_int), ("n_feats", C.c_int),
        ("scale", C.c_int), ("style", C.c_int), ("reduced", C.c_int),
        ("num_metadata", C.c_int), ("attr_size", C.c_int), ("meta_hidden", C.c_int),
        ("in_feats", C.c_int), ("out_feats", C.c_int),
        ("q_enabled", C.c_void_p), ("any_q", C.c_int), ("chunk_images", C.c_int),
        ("schedule", C.c_int), ("no_group_conv", C.c_int), ("meta_relu", C.c_int), ("res_scale", C.c_float),
        ("conv_w_bf16", C.c_void_p), ("tail_w_bf16", C.c_void_p),
        ("conv_w_f32", C.c_void_p), ("up_w_f32", C.c_void_p), ("tail_w_f32", C.c_void_p),
        ("head_w_f32", C.c_void_p),
        ("conv_b", C.c_void_p), ("up_b", C.c_void_p), ("tail_b", C.c_void_p), ("head_b", C.c_void_p),
        ("ca_blob", C.c_void_p), ("ca_stride", C.c_int),
        ("meta_w1", C.c_void_p), ("meta_b1", C.c_void_p), ("meta_w2", C.c_void_p), ("meta_b2", C.c_void_p),
        ("conv_wT_bf16", C.c_void_p), ("conv_wT_f32", C.c_void_p), ("up_wT_f32", C.c_void_p),
        ("tail_wT_f32", C.c_void_p),
        ("pa_blob", C.c_void_p), ("pa_stride", C.c_int),
    ]


class QrcanParams(C.Structure):
    """mirror of `dfir_qrcan_params` (include/dfir.h): device pointer tables of the fp32 parameters / gradients."""
    _fields_ = [
        ("conv_w", C.c_void_p), ("conv_b", C.c_void_p), ("up_w", C.c_void_p), ("up_b", C.c_void_p),
        ("tail_w", C.c_void_p), ("tail_b", C.c_void_p), ("head_w", C.c_void_p), ("head_b", C.c_void_p),
        ("ca", C.c_void_p), ("meta", C.c_void_p), ("pa", C.c_void_p),
    ]


_vp, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/dfir.h declares
PROTOTYPES = {
    "dfir_version": (C.c_char_p, []),
    "dfir_error_string": (C.c_char_p, [_i]),
    "dfir_check_device": (_i, []),
    "dfir_debug_watchdog": (_i, [C.POINTER(C.c_uint * 8), _i]),
    "dfir_debug_trace": (_i, [C.POINTER(C.c_ulonglong * 1024)]),
    "dfir_pack_conv3x3_bf16": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "dfir_pack_conv3x3_f32": (_i, [_vp, _vp, _i, _i, _vp]),
    "dfir_conv3x3_c64": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _ll, _ll, _ll, _vp, _vp, _vp, _i, _vp]),
    "dfir_conv3x3_c64_fused": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _i, _i, _i, _i,
                                    _vp, _vp, _vp, _vp]),
    "dfir_conv3x3_c64_stats": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "dfir_conv3x3_c64_stats_fx": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "dfir_ca_from_stats": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "dfir_conv3x3_c64_scale_skip": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i,
                                         _i, _vp, _vp, _vp]),
    "dfir_conv3x3_c64_scale_skip_hl": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i,
                                            _i, _i, _vp, _vp, _i, _vp]),
    "dfir_conv3x3_c64_scale_skip_hl8": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i,
                                             _i, _i, _vp, _vp, _i, _vp]),
    "dfir_conv3x3_c64_accumulate_hl8": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "dfir_stream_encode_hl8": (_i, [_vp, _vp, _vp, _ll, _vp]),
    "dfir_stream_decode_hl8": (_i, [_vp, _vp, _vp, _ll, _vp]),
    "dfir_conv3x3_c64_accumulate": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "dfir_conv3x3_c64_tail": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp]),
    "dfir_conv3x3_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dfir_head_conv": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "dfir_meta_attention": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _f, _vp]),
    "dfir_ca_scale_residual": (_i, [_vp, _i, _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp, _f, _vp, _vp,
                                    _i, _i, _i, _vp]),
    "dfir_ca_pa_scale_residual": (_i, [_vp, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i,
                                       _vp]),
    "dfir_postprocess_rgb": (_i, [_vp, _vp, _vp, _i, _ll, _f, _f, _vp]),
    "dfir_postprocess_u8_scratch_bytes": (_sz, [_i, _ll]),
    "dfir_postprocess_u8": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _ll, _vp]),
    "dfir_pool_rows_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "dfir_qrcan_workspace_bytes": (_sz, [C.POINTER(QrcanNet), _i, _i, _i, _i]),
    "dfir_qrcan_launch_count": (C.c_longlong, [C.POINTER(QrcanNet), _i, _i, _i, _i]),
    "dfir_qrcan_stages": (_i, [C.POINTER(QrcanNet), _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "dfir_channel_scale": (_i, [_vp, _vp, _vp, _f, _vp, _i, _i, _i, _vp]),
    "dfir_lam_scratch_bytes": (_sz, [_i, _i]),
    "dfir_lam": (_i, [_vp, _ll, _f, _vp, _vp, _i, _i, _i, _i, _vp]),
    "dfir_csam": (_i, [_vp, _vp, _f, _f, _vp, _i, _i, _i, _i, _vp]),
    "dfir_soca_scratch_bytes": (_sz, [_i]),
    "dfir_soca": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp]),
    "dfir_batch_blur": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "dfir_pca_encode": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "dfir_covpool_scratch_bytes": (_sz, [_i]),
    "dfir_covpool": (_i, [_vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "dfir_covpool_backward": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "dfir_sqrtm_scratch_bytes": (_sz, [_i, _i]),
    "dfir_sqrtm": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "dfir_sqrtm_backward": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _vp]),
    "dfir_nonlocal_scratch_bytes": (_sz, [_i, _i, _i]),
    "dfir_nonlocal": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "dfir_qrcan_repack": (_i, [C.POINTER(QrcanNet), C.POINTER(QrcanParams), _i, _i, _vp]),
    "dfir_qrcan_train_workspace_bytes": (_sz, [C.POINTER(QrcanNet), _i, _i, _i, _i]),
    "dfir_qrcan_train_forward": (_i, [C.POINTER(QrcanNet), _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "dfir_qrcan_train_backward": (_i, [C.POINTER(QrcanNet), C.POINTER(QrcanParams), _vp, _vp, _vp, _i, _i, _i, _i, _vp,
                                       _sz, _vp]),
    "dfir_qrcan_train_launch_count": (C.c_longlong, [C.POINTER(QrcanNet), _i, _i, _i, _i]),
    "dfir_adam_step": (_i, [_vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _ll, _vp]),
    "dfir_conv3x3_wgrad_scratch_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "dfir_conv3x3_wgrad_c64": (_i, [_vp, _ll, _ll, _ll, _vp, _i, _i, _i, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "dfir_conv3x3_wgrad_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "dfir_conv3x3_c64_dgrad": (_i, [_vp, _ll, _ll, _ll, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "dfir_pack_conv3x3_bf16_ex": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "dfir_conv3x3_wgrad_small_scratch_bytes": (_sz, [_i, _i, _i]),
    "dfir_conv3x3_wgrad_small": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "dfir_qrcan_forward": (_i, [C.POINTER(QrcanNet), _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "dfir_channel_dot_scratch_bytes": (_sz, [_i, _i]),
    "dfir_channel_dot": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _sz, _i, _ll, _i, _vp]),
    "dfir_soca_mlp": (_i, [_vp, _vp, _i, _vp, _i, _vp]),
    "dfir_soca_mlp_backward_scratch_bytes": (_sz, [_i, _i]),
    "dfir_soca_mlp_backward": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _sz, _i, _vp]),
    "dfir_lam_backward_scratch_bytes": (_sz, [_i, _i]),
    "dfir_lam_backward": (_i, [_vp, _ll, _vp, _f, _vp, _vp, _ll, _vp, _vp, _sz, _i, _i, _i, _i, _vp]),
    "dfir_csam_backward_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "dfir_csam_backward": (_i, [_vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _vp]),
    "dfir_nonlocal_backward_scratch_bytes": (_sz, [_i, _i, _i]),
    "dfir_nonlocal_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _i, _i, _i, _i, _vp]),
    "dfir_pack_conv3x3_f32_ex": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "dfir_qrcan_train_stage_forward": (_i, [C.POINTER(QrcanNet), _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i,
                                            _vp, _sz, _vp]),
    "dfir_qrcan_train_stage_backward": (_i, [C.POINTER(QrcanNet), C.POINTER(QrcanParams), _i, _i, _i, _vp, _vp, _vp, _vp,
                                             _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
}


def load_library():
    """dlopen the in-tree library; raises DfirError (never falls back) when it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.isfile(path):
        raise DfirError("%s not found — build it with `make -C %s` (or __graft_entry__.build()); "
                        "there is no CPU / eager fallback" % (path, os.path.join(_PKG, "csrc")))
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        lib = load_library()
        msg = lib.dfir_error_string(rc).decode()
        extra = ""
        if rc == -2:
            try:
                import torch
                torch.cuda.synchronize()
            except Exception as e:  # surface the sticky CUDA error text
                extra = " [%s]" % e
        raise DfirError("dfir %s failed: %s (code %d)%s" % (what, msg, rc, extra))
