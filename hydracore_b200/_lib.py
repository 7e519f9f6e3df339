"""ctypes binding of include/hydracore_cuda.h.  Fails loudly when the library is missing: there is no fallback path."""
import ctypes as ct
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
HC_HOST, HC_DEVICE = 0, 1


class HcError(RuntimeError):
    pass


class hc_hit(ct.Structure):
    _fields_ = [("t", ct.c_float), ("primId", ct.c_int32), ("instId", ct.c_int32), ("geomId", ct.c_int32)]


class hc_stats(ct.Structure):
    _fields_ = [("raysClosest", ct.c_uint64), ("raysShadow", ct.c_uint64), ("paths", ct.c_uint64), ("kernelLaunches", ct.c_uint64),
                ("msClosest", ct.c_float), ("msShadow", ct.c_float), ("msShade", ct.c_float), ("msOther", ct.c_float)]


def lib_path():
    # HC_LIB: another build of the same library (kernel experiments under scripts/), never a different implementation
    return os.environ.get("HC_LIB") or os.path.join(_HERE, "libhydracore_b200.so")


_P, _I, _I64, _U64 = ct.c_void_p, ct.c_int, ct.c_int64, ct.c_uint64
_PP = ct.POINTER(ct.c_void_p)

# name -> (restype, argtypes); mirrors include/hydracore_cuda.h one to one (tests/test_abi.py checks the header against this)
SIGNATURES = {
    "hc_abi_version": (_I, []),
    "hc_last_error": (ct.c_char_p, []),
    "hc_device_count": (_I, [ct.POINTER(_I)]),
    "hc_ctx_create": (_I, [_I, _PP]),
    "hc_ctx_destroy": (None, [_P]),
    "hc_device_name": (_I, [_P, ct.c_char_p, _I]),
    "hc_mem_info": (_I, [_P, ct.POINTER(ct.c_size_t), ct.POINTER(ct.c_size_t)]),
    "hc_sync": (_I, [_P]),
    "hc_stream": (_I, [_P, _PP]),
    "hc_storage_reserve": (_I, [_P, _I, _U64]),
    "hc_storage_write": (_I, [_P, _I, _U64, _P, _U64]),
    "hc_storage_capacity": (_I, [_P, _I, ct.POINTER(_U64)]),
    "hc_set_globals": (_I, [_P, _P, _U64]),
    "hc_set_bvh": (_I, [_P, _I, _P, _I, _P, _I, _I]),
    "hc_set_bvh_alpha": (_I, [_P, _I, _P, _I, _P, _I, _P, _I, _I]),
    "hc_set_remap_lists": (_I, [_P, _P, _P, _I, _I]),
    "hc_set_inst_remap_ids": (_I, [_P, _P, _I]),
    "hc_bvh_device_layout": (_I, [_P, _I, _P, _I, _P, _P, _I64, ct.POINTER(_I64), ct.POINTER(_I)]),
    "hc_set_inst_matrices": (_I, [_P, _P, _I]),
    "hc_set_inst_light_ids": (_I, [_P, _P, _I]),
    "hc_resize": (_I, [_P, _I, _I]),
    "hc_make_eye_rays": (_I, [_P, _I, _I, _P, _P, _I]),
    "hc_trace_closest": (_I, [_P, _P, _I64, _P, _I]),
    "hc_trace_shadow": (_I, [_P, _P, _I64, _P, _I]),
    "hc_make_shadow_rays": (_I, [_P, _P, _P, _I64, ct.POINTER(ct.c_float), _P, _I]),
    "hc_raycast_pass": (_I, [_P, ct.POINTER(ct.c_float), _P, _P, _I]),
    "hc_trace_last_ms": (_I, [_P, ct.POINTER(ct.c_float)]),
    "hc_measure_read_bandwidth": (_I, [_P, _U64, _I, ct.POINTER(ct.c_float)]),
    "hc_pt_init": (_I, [_P, _I]),
    "hc_pt_set_tiles": (_I, [_P, _I, _I, _I]),
    "hc_pt_set_material_sort": (_I, [_P, _I, _I]),
    "hc_pt_set_shadow_trees": (_I, [_P, _I]),
    "hc_pt_set_sample_streams": (_I, [_P, _I, ct.c_int64]),
    "hc_pt_group_passes": (_I, [_P, ct.POINTER(ct.c_int)]),
    "hc_pt_pass": (_I, [_P, _I, _I]),
    "hc_fb_clear": (_I, [_P]),
    "hc_fb_device_ptr": (_I, [_P, _PP, ct.POINTER(_I64)]),
    "hc_fb_read_hdr": (_I, [_P, _P, _I, _I]),
    "hc_fb_read_sum": (_I, [_P, _P, _I, _I]),
    "hc_fb_read_ldr": (_I, [_P, _P, _I, _I]),
    "hc_comm_unique_id": (_I, [_P]),
    "hc_comm_init": (_I, [_P, _P, _I, _I]),
    "hc_comm_version": (_I, [ct.POINTER(_I)]),
    "hc_fb_reduce": (_I, [_P, _I, _I, ct.POINTER(ct.c_float)]),
    "hc_get_spp": (_I, [_P, ct.POINTER(ct.c_float)]),
    "hc_get_stats": (_I, [_P, ct.POINTER(hc_stats)]),
    "hc_reset_stats": (_I, [_P]),
    "hc_bvh_create": (_I, [_PP]),
    "hc_bvh_destroy": (None, [_P]),
    "hc_bvh_add_mesh": (_I, [_P, _P, _I, _P, _I, ct.POINTER(_I)]),
    "hc_bvh_add_instance": (_I, [_P, _I, _P, ct.POINTER(_I)]),
    "hc_bvh_add_instance_id": (_I, [_P, _I, _P, _I]),
    "hc_bvh_commit": (_I, [_P]),
    "hc_bvh_result": (_I, [_P, _PP, ct.POINTER(_I), _PP, ct.POINTER(_I), _PP, ct.POINTER(_I), ct.POINTER(_I)]),
    "hc_bvh_bounds": (_I, [_P, ct.POINTER(ct.c_float), ct.POINTER(ct.c_float)]),
}

_lib = None


def load():
    """Load libhydracore_b200.so (built in-tree by __graft_entry__.build()) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise HcError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` — there is no CPU fallback")
    lib = ct.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        if os.environ.get("HC_LIB") and not hasattr(lib, name):
            continue                  # an older build of the library under test (scripts/ only)
        fn = getattr(lib, name)       # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().hc_last_error().decode(errors="replace")
        raise HcError(f"{what} failed with status {rc}: {msg}")
