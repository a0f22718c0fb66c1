"""maskrcnn_b200 — B200-native (sm_100a) RoI hot path of Mask R-CNN behind the reference's operator API.

    nms, CropFunction                      drop-ins for c++ext/maskrcnn/__init__.py
    roi_align, rpn_refine, mrn_refine, mrn_samples   drop-ins for the model.py functions that call them
    rpn_samples                            drop-in for data.rpn_samples (RPN anchor matching)
    full_masks, decode_masks               drop-ins for data.full_masks / data.decode_masks (mask paste-back, back to the original frame)
    rpn_detect, rpn_pack                   drop-in for MaskRCNN.rpn_detect: RPN head outputs -> [B,A,2] / [B,A,4] in one launch
    pyramid_roi_align, proposal_layer, detection_layer, detection_targets   batched, sync-free variants
    patch(model_module)                    swaps the fused versions into an unmodified reference model.py

All compute is hand-written CUDA in libmrcnn_b200.so (C ABI: include/mrcnn_b200.h).  No CPU fallback.
"""
import importlib

# Everything below lives in .ops, which loads libmrcnn_b200.so on import.  The names resolve on first use (PEP 562), so that
# the host-only helpers (maskrcnn_b200.synth, maskrcnn_b200.roofline: numpy input synthesis and byte accounting, which the
# CPU reference arm of bench.py also uses) can be imported without mapping the CUDA library into that process.  There is
# still no fallback: touching any operator imports .ops, and a missing library raises ImportError there.
_OPS = ("CropFunction", "check_device_errors", "crop_and_resize", "decode_masks", "detection_layer", "detection_targets",
        "full_masks", "mrn_refine", "mrn_samples", "nms", "proposal_layer", "pyramid_roi_align", "pyramid_roi_align_backward_pair",
        "pyramid_roi_align_pair", "roi_align", "rpn_detect", "rpn_pack", "rpn_refine", "rpn_samples", "set_backward_algorithm",
        "set_backward_planning", "set_detection_nms", "set_proposal_nms", "set_deterministic")
_LIB = ("LIB_PATH", "MrcnnError")
__all__ = list(_OPS) + list(_LIB) + ["patch"]


def __getattr__(name):
    if name in _OPS:
        value = getattr(importlib.import_module(".ops", __name__), name)
    elif name in _LIB:
        value = getattr(importlib.import_module("._lib", __name__), name)
    else:
        raise AttributeError("module %r has no attribute %r" % (__name__, name))
    globals()[name] = value
    return value


def __dir__():
    return sorted(set(globals()) | set(__all__))


__version__ = "0.1.0"


def patch(model_module, data_module=None):
    """Monkey-patches an imported reference `model` module (model.py) so that its RoI hot path runs on
    the fused kernels: model.roi_align, model.mrn_samples, MaskRCNN.rpn_detect, MaskRCNN.rpn_refine, MaskRCNN.mrn_refine; with
    the reference's `data` module given as well, data.rpn_samples (the anchor matching the dataset runs per sample - the
    reference's DataLoader uses num_workers=0, model.py:1528-1532, so it runs in the CUDA process), data.full_masks (the
    mask paste-back model.predict calls as datalib.full_masks, model.py:1190) and data.decode_masks (model.detect,
    model.py:1130).  The `maskrcnn` package the
    module imported (model.py:25) should already be this repo's drop-in (put the repo root on sys.path)."""
    from . import ops
    roi_align, rpn_refine, rpn_detect, mrn_refine, mrn_samples = ops.roi_align, ops.rpn_refine, ops.rpn_detect, ops.mrn_refine, ops.mrn_samples
    rpn_samples, full_masks, decode_masks = ops.rpn_samples, ops.full_masks, ops.decode_masks
    model_module.roi_align = roi_align
    model_module.MaskRCNN.rpn_refine = rpn_refine
    model_module.MaskRCNN.rpn_detect = rpn_detect
    model_module.MaskRCNN.mrn_refine = mrn_refine
    model_module.mrn_samples = mrn_samples
    if data_module is not None:
        data_module.rpn_samples = rpn_samples
        data_module.full_masks = full_masks
        data_module.decode_masks = decode_masks
    return model_module
