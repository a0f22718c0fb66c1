"""ctypes binding of libmrcnn_b200.so (include/mrcnn_b200.h).  There is no fallback: if the library
is missing or a call fails, an exception is raised."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MRCNN_B200_DEBUG=1 selects the -DMRCNN_DEBUG build (`make -C maskrcnn_b200/csrc debug`: device-side index asserts)
LIB_PATH = os.path.join(_HERE, "libmrcnn_b200_debug.so" if os.environ.get("MRCNN_B200_DEBUG") == "1" else "libmrcnn_b200.so")

NCHW, NHWC = 0, 1
OK = 0
BWD_AUTO, BWD_GATHER, BWD_SCATTER = 0, 1, 2
E_INVALID_ARG, E_NOT_DEVICE_PTR, E_WORKSPACE, E_CUDA, E_BOX_INDEX, E_CLASS_ID = -1, -2, -3, -4, -5, -6

_vp = ctypes.c_void_p
_i = ctypes.c_int
_f = ctypes.c_float
_sz = ctypes.c_size_t
_i4 = ctypes.c_int * 4
_vp4 = ctypes.c_void_p * 4
_f4 = ctypes.c_float * 4

# name -> (restype, argtypes); must list every symbol include/mrcnn_b200.h declares (tests/test_abi.py)
SIGNATURES = {
    "mrcnn_abi_version": (_i, []),
    "mrcnn_set_deterministic": (_i, [_i]),
    "mrcnn_last_error": (ctypes.c_char_p, []),
    "mrcnn_poll_device_errors": (_i, [_vp]),
    "mrcnn_crop_forward": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _f, _i, _i, _vp, _i, _vp]),
    "mrcnn_crop_backward": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "mrcnn_pyramid_roi_align_forward": (_i, [_vp4, _i4, _i4, _i, _i, _i, _vp, _vp, _i, _i, _f, _vp, _i, _vp, _vp]),
    "mrcnn_pyramid_roi_align_forward_pair": (_i, [_vp4, _i4, _i4, _i, _i, _vp, _vp, _i, _f, _vp, _vp, _vp]),
    "mrcnn_pyramid_roi_align_backward_workspace_bytes": (_sz, [_i4, _i4, _i, _i, _i]),
    "mrcnn_pyramid_roi_align_backward_workspace_bytes_ex": (_sz, [_i4, _i4, _i, _i, _i, _i, _i]),
    "mrcnn_pyramid_roi_align_backward": (_i, [_vp, _i, _i4, _i4, _i, _i, _vp, _vp, _i, _i, _f, _vp4, _i, _i, _vp, _i, _vp, _sz, _vp]),
    "mrcnn_pyramid_roi_align_backward_plan": (_i, [_i4, _i4, _i, _i, _vp, _vp, _i, _i, _f, _vp, _sz, _vp]),
    "mrcnn_pyramid_roi_align_backward_planned": (_i, [_vp, _i4, _i4, _i, _i, _i, _i, _vp4, _i, _vp, _sz, _vp]),
    "mrcnn_pyramid_roi_align_backward_pair_workspace_bytes": (_sz, [_i4, _i4, _i, _i, _i, _i]),
    "mrcnn_pyramid_roi_align_backward_pair": (_i, [_vp, _i, _vp, _i, _i4, _i4, _i, _i, _vp, _vp, _i, _f, _vp4, _i, _vp, _sz, _vp]),
    "mrcnn_rpn_match_workspace_bytes": (_sz, [_i]),
    "mrcnn_rpn_match": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "mrcnn_compact_equal_workspace_bytes": (_sz, [_i]),
    "mrcnn_compact_equal": (_i, [_vp, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "mrcnn_scatter_fill": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "mrcnn_rpn_deltas": (_i, [_vp, _vp, _vp, _vp, _vp, _i, ctypes.c_double * 4, _vp, _vp]),
    "mrcnn_target_classify": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mrcnn_target_select": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "mrcnn_target_emit": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f4, _i, _i, _i, _vp, _vp, _vp,
                               _vp, _vp]),
    "mrcnn_nms_workspace_bytes": (_sz, [_i]),
    "mrcnn_nms": (_i, [_vp, _i, _f, _vp, _vp, _vp, _sz, _vp]),
    "mrcnn_proposal_workspace_bytes": (_sz, [_i, _i, _i]),
    "mrcnn_proposal_layer": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _f4, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    "mrcnn_full_masks_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "mrcnn_full_masks": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "mrcnn_decode_masks_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "mrcnn_decode_masks": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "mrcnn_set_proposal_nms": (_i, [_i]),
    "mrcnn_set_detection_nms": (_i, [_i]),
    "mrcnn_proposal_layer_fg": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _f4, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    "mrcnn_rpn_pack": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "mrcnn_rpn_unpack": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "mrcnn_detection_workspace_bytes": (_sz, [_i, _i]),
    "mrcnn_detection_layer": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _i, _f4, _f, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mrcnn_detection_exchange_bytes": (_sz, [_i, _i, _i]),
    "mrcnn_detection_layer_exchange": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _i, _f4, _f, _f, _vp, _vp, _vp, _vp, _i, _i,
                                            _vp, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "mrcnn_detection_collect": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
}


class MrcnnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmrcnn_b200 error %d: %s" % (code, msg))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "maskrcnn_b200: %s is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C maskrcnn_b200/csrc`.  There is no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.mrcnn_abi_version() != 6:
        raise ImportError("maskrcnn_b200: ABI version mismatch")
    return lib


lib = _load()


def check(rc):
    if rc != OK:
        raise MrcnnError(rc, lib.mrcnn_last_error().decode("utf-8", "replace"))


def i4(vals):
    return _i4(*[int(v) for v in vals])


def vp4(vals):
    return _vp4(*[int(v) for v in vals])


def i4_cached(vals, _cache={}):
    key = tuple(int(v) for v in vals)
    a = _cache.get(key)
    if a is None:
        a = _cache[key] = _i4(*key)
    return a


def i32_array(vals):
    """Host int32 array (e.g. image_offsets_host) or None."""
    if vals is None:
        return None
    arr = (ctypes.c_int32 * len(vals))(*[int(v) for v in vals])
    return arr


def f4(vals):
    return _f4(*[float(v) for v in vals])
