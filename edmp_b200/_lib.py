"""ctypes binding of libedmp_b200.so (include/edmp_b200.h).  There is deliberately no fallback:
if the CUDA library is missing or no GPU is present the product raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libedmp_b200.so")

PRECISIONS = {"fp32": 0, "tf32x3": 1, "tf32": 2, "bf16x3": 3, "bf16": 4, "f16x3": 5, "f16": 6}

c_void_p, c_int, c_size_t, c_char_p = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_char_p
c_double, c_uint64, c_longlong, c_float = ctypes.c_double, ctypes.c_uint64, ctypes.c_longlong, ctypes.c_float
P = ctypes.POINTER

# symbol -> (restype, argtypes); mirrors include/edmp_b200.h one to one
SIGNATURES = {
    "edmp_last_error": (c_char_p, []),
    "edmp_version": (c_int, []),
    "edmp_unet_param_count": (c_size_t, [P(c_int), c_int]),
    "edmp_unet_create": (c_int, [c_void_p, c_size_t, P(c_int), c_int, c_int, c_int, P(c_void_p)]),
    "edmp_unet_destroy": (None, [c_void_p]),
    "edmp_unet_pack": (c_int, [c_void_p, c_size_t, P(c_int), c_int, c_int, c_int, P(c_void_p)]),
    "edmp_unet_blob_bytes": (c_size_t, [c_void_p]),
    "edmp_unet_blob_read": (c_int, [c_void_p, c_void_p, c_size_t]),
    "edmp_unet_create_from_blob": (c_int, [c_void_p, c_size_t, c_int, P(c_void_p)]),
    "edmp_unet_blob_layout_version": (c_int, []),
    "edmp_unet_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "edmp_unet_read_activation": (c_int, [c_void_p, c_char_p, c_int, c_void_p, P(c_int), P(c_int), c_void_p]),
    "edmp_unet_profile": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "edmp_unet_op_name": (c_char_p, [c_void_p, c_int]),
    "edmp_unet_op_kernel": (c_char_p, [c_void_p, c_int]),
    "edmp_unet_tc_trace": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, P(c_int), c_void_p]),
    "edmp_unet_precision": (c_int, [c_void_p]),
    "edmp_unet_range_status": (c_int, [c_void_p, P(c_int), c_void_p]),
    "edmp_unet_launches_per_forward": (c_int, [c_void_p]),
    "edmp_scene_create": (c_int, [c_void_p, c_int, c_void_p, P(c_void_p)]),
    "edmp_scene_destroy": (None, [c_void_p]),
    "edmp_scene_set_guide_tables": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "edmp_guide_gradient": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "edmp_guide_volumes": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "edmp_guide_final_cost": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "edmp_sampler_create": (c_int, [c_int, c_double, c_int, P(c_void_p)]),
    "edmp_sampler_destroy": (None, [c_void_p]),
    "edmp_sample_guided": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_uint64, c_int, c_int, c_int, c_void_p, c_void_p]),
    "edmp_sample_guided_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64,
                                        c_int, c_void_p, c_void_p]),
    "edmp_sampler_schedule": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "edmp_sampler_last_launches": (c_longlong, [c_void_p]),
    "edmp_sampler_set_condition": (c_int, [c_void_p, c_int]),
    "edmp_sdf_scene_create": (c_int, [c_void_p, c_int, c_void_p, c_int, P(c_void_p)]),
    "edmp_sdf_scene_destroy": (None, [c_void_p]),
    "edmp_sdf_guide": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "edmp_sdf_cloud_clearance": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "edmp_metrics_nfft": (c_int, [c_int, c_int]),
    "edmp_ee_transform": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "edmp_trajectory_metrics": (c_int, [c_void_p, c_int, c_int, c_double, c_int, c_double, c_double, c_void_p,
                                        c_void_p, c_void_p, c_void_p]),
    "edmp_sparc": (c_int, [c_void_p, c_int, c_int, c_double, c_int, c_double, c_double, c_void_p, c_void_p, c_void_p,
                           c_void_p]),
}

_lib = None


class EdmpError(RuntimeError):
    pass


def load():
    """Loads the shared library (does not touch the GPU)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EdmpError(
                "libedmp_b200.so is not built (%s). Build it with `make -C edmp_b200/csrc` or "
                "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        raise EdmpError("%s failed (rc=%d): %s" % (what, rc, load().edmp_last_error().decode()))


def require_cuda(device):
    import torch
    dev = torch.device(device)
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise EdmpError("edmp_b200 runs its hot path on a CUDA device only (got device=%r, "
                        "cuda available=%s); there is no CPU fallback." % (device, torch.cuda.is_available()))
    return dev


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def host_f64(a):
    """contiguous float64 numpy array + its pointer"""
    import numpy as np
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return arr, arr.ctypes.data_as(c_void_p)
