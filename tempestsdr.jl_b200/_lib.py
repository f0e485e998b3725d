"""ctypes loader for libtempest_b200.so (the C ABI of include/tempest_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc, and
if that fails -- or a compute call finds no CUDA device -- an exception is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# TEMPEST_B200_LIB (also read by the Julia wrapper) points at another build of the library, e.g. for A/B timing
SO = os.environ.get("TEMPEST_B200_LIB") or os.path.join(HERE, "libtempest_b200.so")

TSDR_CHAIN_PUBLISH_ALL = 1
TSDR_CHAIN_NO_ALIGN = 2
TSDR_CHAIN_SUM = 4
TSDR_CHAIN_NO_OVERLAP = 8
TSDR_CHAIN_FULLRES = 16

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/tempest_b200.h declares
SIGNATURES = {
    "tsdr_version": (C.c_int, []),
    "tsdr_last_error_string": (C.c_char_p, []),
    "tsdr_device_count": (C.c_int, [_ip]),
    "tsdr_set_device": (C.c_int, [C.c_int]),
    "tsdr_am_demod_f32": (C.c_int, [_vp, _vp, C.c_size_t]),
    "tsdr_invert_am_demod_f32": (C.c_int, [_vp, _vp, C.c_size_t]),
    "tsdr_fm_demod_f32": (C.c_int, [_vp, _vp, C.c_size_t]),
    "tsdr_abs2_f32": (C.c_int, [_vp, _vp, C.c_size_t]),
    "tsdr_sig_to_image_f32": (C.c_int, [_vp, C.c_size_t, C.c_int, C.c_int, _vp]),
    "tsdr_downgrade_f32": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "tsdr_naive_resampler_f32": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int]),
    "tsdr_upsampler_create": (C.c_int, [C.c_size_t, C.c_int, C.POINTER(_vp)]),
    "tsdr_upsampler_apply_f32": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t]),
    "tsdr_upsampler_get_filter": (C.c_int, [_vp, _vp]),
    "tsdr_upsampler_destroy": (C.c_int, [_vp]),
    "tsdr_autocorr_f32": (C.c_int, [_vp, C.c_size_t, C.c_double, C.c_double, C.c_double, C.c_int, _vp,
                                    C.POINTER(C.c_size_t)]),
    "tsdr_autocorr_out_len": (C.c_int, [C.c_size_t, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_size_t)]),
    "tsdr_findmax_f32": (C.c_int, [_vp, C.c_size_t, _fp, C.POINTER(C.c_size_t)]),
    "tsdr_findmax_windows_dev_f32": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), _fp,
                                               C.POINTER(C.c_size_t), _vp]),
    "tsdr_findmax_dev_f32": (C.c_int, [_vp, C.c_size_t, _fp, C.POINTER(C.c_size_t), _vp]),
    "tsdr_full_scale_f32": (C.c_int, [_vp, _vp, C.c_size_t]),
    "tsdr_sync_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_vp)]),
    "tsdr_sync_bounds": (C.c_int, [_vp, _ip, _ip, _ip, _ip]),
    "tsdr_vsync_f32": (C.c_int, [_vp, _vp, _ip, _ip]),
    "tsdr_sync_get_beta": (C.c_int, [_vp, _vp, _vp]),
    "tsdr_sync_destroy": (C.c_int, [_vp]),
    "tsdr_chain_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, C.c_float,
                                    C.c_size_t, C.c_uint, _vp]),
    "tsdr_chain_configure": (C.c_int, [_vp, C.c_double, C.c_int, C.c_int, C.c_double]),
    "tsdr_chain_set_alpha": (C.c_int, [_vp, C.c_float]),
    "tsdr_chain_reset": (C.c_int, [_vp]),
    "tsdr_chain_push_host": (C.c_int, [_vp, _vp, C.c_size_t, _ip]),
    "tsdr_chain_push_device": (C.c_int, [_vp, _vp, C.c_size_t, _ip]),
    "tsdr_chain_push_host_i16": (C.c_int, [_vp, _vp, C.c_size_t, _ip]),
    "tsdr_chain_push_host_i16_deliver": (C.c_int, [_vp, _vp, C.c_size_t, _ip, _vp]),
    "tsdr_chain_push_device_i16": (C.c_int, [_vp, _vp, C.c_size_t, _ip]),
    "tsdr_chain_push_host_deliver": (C.c_int, [_vp, _vp, C.c_size_t, _ip, _vp]),
    "tsdr_chain_wait_delivery": (C.c_int, [_vp, C.c_int]),
    "tsdr_chain_prime_host": (C.c_int, [_vp, _vp, C.c_size_t]),
    "tsdr_chain_prime_device": (C.c_int, [_vp, _vp, C.c_size_t]),
    "tsdr_chain_sync": (C.c_int, [_vp]),
    "tsdr_chain_flush": (C.c_int, [_vp]),
    "tsdr_chain_read_image": (C.c_int, [_vp, _vp]),
    "tsdr_chain_image_size": (C.c_int, [_vp, _ip, _ip]),
    "tsdr_chain_read_image_downgraded": (C.c_int, [_vp, _vp]),
    "tsdr_chain_read_offsets": (C.c_int, [_vp, _vp, _vp, C.c_int, _ip]),
    "tsdr_chain_read_scores": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, _ip]),
    "tsdr_chain_read_published": (C.c_int, [_vp, _vp, C.c_int, _ip]),
    "tsdr_chain_accumulator": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(C.c_size_t)]),
    "tsdr_chain_scale_accumulator": (C.c_int, [_vp, C.c_float]),
    "tsdr_chain_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "tsdr_chain_launch_count": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "tsdr_chain_set_profiling": (C.c_int, [_vp, C.c_int]),
    "tsdr_chain_kernel_times": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]),
    "tsdr_chain_destroy": (C.c_int, [_vp]),
    "tsdr_comm_available": (C.c_int, [_ip]),
    "tsdr_comm_get_unique_id": (C.c_int, [_vp]),
    "tsdr_comm_init_rank": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, _vp]),
    "tsdr_comm_init_all": (C.c_int, [C.POINTER(_vp), C.c_int, _ip]),
    "tsdr_comm_info": (C.c_int, [_vp, _ip, _ip, _ip, C.POINTER(C.c_uint64)]),
    "tsdr_comm_group_start": (C.c_int, []),
    "tsdr_comm_group_end": (C.c_int, []),
    "tsdr_chain_allreduce": (C.c_int, [_vp, _vp, C.c_float]),
    "tsdr_chain_integrate_device": (C.c_int, [_vp, _vp, C.c_size_t, C.POINTER(_vp), C.POINTER(C.c_size_t), C.c_int, _vp,
                                              C.c_float, _ip]),
    "tsdr_comm_allreduce_f32": (C.c_int, [_vp, _vp, C.c_size_t, C.c_float, _vp]),
    "tsdr_comm_allgather": (C.c_int, [_vp, _vp, _vp, C.c_size_t, _vp]),
    "tsdr_comm_destroy": (C.c_int, [_vp]),
    "tsdr_get_spectrum_f32": (C.c_int, [_vp, C.c_size_t, C.c_int, _vp]),
    "tsdr_get_welch_f32": (C.c_int, [_vp, C.c_size_t, C.c_int, _vp]),
    "tsdr_get_waterfall_f32": (C.c_int, [_vp, C.c_size_t, C.c_int, _vp]),
    "tsdr_chain_push_ring": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _ip]),
    "tsdr_ring_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t, C.c_int, C.c_int]),
    "tsdr_ring_destroy": (C.c_int, [_vp]),
    "tsdr_ring_put": (C.c_int, [_vp, _vp, C.c_size_t]),
    "tsdr_ring_take": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int]),
    "tsdr_ring_acquire_write": (C.c_int, [_vp, C.POINTER(C.c_void_p)]),
    "tsdr_ring_commit": (C.c_int, [_vp]),
    "tsdr_ring_acquire_read": (C.c_int, [_vp, C.POINTER(C.c_void_p), C.c_int]),
    "tsdr_ring_release_read": (C.c_int, [_vp]),
    "tsdr_ring_stats": (C.c_int, [_vp, _ip, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "tsdr_ring_slot_bytes": (C.c_size_t, [_vp]),
    "tsdr_selftest_hypot": (C.c_int, [C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
    "tsdr_autocorr_plan_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_size_t, _vp]),
    "tsdr_autocorr_plan_exec": (C.c_int, [_vp, _vp, C.c_size_t, C.c_size_t, C.c_int, _vp]),
    "tsdr_autocorr_plan_launch_count": (C.c_int, [_vp, C.POINTER(C.c_uint64)]),
    "tsdr_autocorr_plan_destroy": (C.c_int, [_vp]),
}


class TempestError(RuntimeError):
    """Non-zero status from libtempest_b200 (the Julia wrapper throws ErrorException)."""

    def __init__(self, status, message):
        super().__init__("libtempest_b200 status %d: %s" % (status, message))
        self.status = status


_lib = None


def load():
    """Load (building first if needed) the shared library; raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_tsdr_build", os.path.join(HERE, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    lib = C.CDLL(SO)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise TempestError(status, load().tsdr_last_error_string().decode("utf-8", "replace"))


def device_count():
    n = C.c_int(0)
    check(load().tsdr_device_count(C.byref(n)))
    return n.value
