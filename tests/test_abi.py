"""The C-ABI library loads on a CPU-only box and exports every symbol include/mrcnn_b200.h declares;
argument validation (no compute) behaves as documented."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "mrcnn_b200.h")).read()
    return sorted(set(re.findall(r"MRCNN_API\s+[\w\s\*]+?\b(mrcnn_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from maskrcnn_b200 import _lib
    names = _declared()
    assert len(names) >= 14
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libmrcnn_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "maskrcnn_b200._lib has no binding for %s" % n
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_workspace_queries():
    from maskrcnn_b200 import _lib
    L = _lib.lib
    assert L.mrcnn_abi_version() == 4
    assert L.mrcnn_nms_workspace_bytes(6000) >= 6000 * 94 * 8
    assert L.mrcnn_proposal_workspace_bytes(8, 261888, 6000) >= 8 * 6016 * 94 * 8
    assert L.mrcnn_detection_workspace_bytes(64, 1000) == 256          # mask lives in shared memory up to 1024 RoIs
    assert L.mrcnn_detection_workspace_bytes(2, 2000) >= 2 * 2048 * 32 * 8
    hw = _lib.i4([256, 128, 64, 32])
    units = 16 * (256 * 32 + 128 * 16 + 64 * 8 + 32 * 4)
    assert L.mrcnn_pyramid_roi_align_backward_workspace_bytes(hw, hw, 16, 8192, 14) >= 8 * units + 4 * 8192 * 196 * 16


def test_bad_arguments_are_rejected_without_a_gpu():
    from maskrcnn_b200 import _lib
    L = _lib.lib
    # invalid sizes -> MRCNN_E_INVALID_ARG before anything touches the device
    rc = L.mrcnn_crop_forward(None, 0, 1, 1, 1, 0, None, None, 1, 0.0, 7, 7, None, 0, None)
    assert rc == _lib.E_INVALID_ARG
    assert b"image dims" in L.mrcnn_last_error()
    rc = L.mrcnn_crop_forward(None, 1, 1, 4, 4, 7, None, None, 1, 0.0, 7, 7, None, 0, None)
    assert rc == _lib.E_INVALID_ARG and b"layout" in L.mrcnn_last_error()
    with pytest.raises(_lib.MrcnnError):
        _lib.check(rc)


def test_python_api_has_no_cpu_fallback():
    import torch
    import maskrcnn
    import maskrcnn_b200 as m
    assert maskrcnn.nms is m.nms and maskrcnn.CropFunction is m.CropFunction
    with pytest.raises(TypeError):
        m.nms(torch.zeros(3, 5), 0.5)
    with pytest.raises(TypeError):
        m.pyramid_roi_align([torch.zeros(1, 4, s, s) for s in (8, 4, 2, 1)], torch.zeros(2, 4), None, 7, (32, 32, 3))
    with pytest.raises(TypeError):
        m.proposal_layer(torch.zeros(1, 8, 2), torch.zeros(1, 8, 4), torch.zeros(8, 4), 4, 2, 0.7)
    with pytest.raises(TypeError):
        m.detection_layer(torch.zeros(1, 4, 4), torch.zeros(1, 4, 3), torch.zeros(1, 4, 3, 4), torch.zeros(1, 4), 0, 0.3, 2)
    with pytest.raises(TypeError):
        m.full_masks(torch.zeros(1, dtype=torch.int64), torch.zeros(1, 4), torch.zeros(1, 2, 28, 28), 32, 32)
    with pytest.raises(TypeError):
        m.decode_masks(torch.zeros(1, 32, 32, dtype=torch.bool), 0.5, (32, 32))


def test_patch_swaps_every_replaced_symbol():
    """patch() on stand-ins for the reference's model / data modules: every function SURVEY 8(a) + 8(f) lists is swapped."""
    import types
    import maskrcnn_b200 as m
    model = types.SimpleNamespace(MaskRCNN=type("MaskRCNN", (), {}), roi_align=None, mrn_samples=None)
    data = types.SimpleNamespace(rpn_samples=None, full_masks=None, decode_masks=None)
    assert m.patch(model, data) is model
    assert model.roi_align is m.roi_align and model.mrn_samples is m.mrn_samples
    assert model.MaskRCNN.rpn_detect is m.rpn_detect and model.MaskRCNN.rpn_refine is m.rpn_refine and model.MaskRCNN.mrn_refine is m.mrn_refine
    assert data.rpn_samples is m.rpn_samples and data.full_masks is m.full_masks and data.decode_masks is m.decode_masks


def test_decode_masks_arguments_without_a_gpu():
    from maskrcnn_b200 import _lib
    L = _lib.lib
    assert L.mrcnn_decode_masks_workspace_bytes(640, 1024, 1200, 1920) >= (1200 + 1920) * (8 + 3 * 4)
    rc = L.mrcnn_decode_masks(None, 1, 1, 64, 64, 10, 0, 60, 64, 120, 128, None, None, 0, None)    # window leaves the mask
    assert rc == _lib.E_INVALID_ARG and b"crop window" in L.mrcnn_last_error()
    rc = L.mrcnn_decode_masks(None, 1, 1, 64, 64, 0, 0, 64, 64, 0, 128, None, None, 0, None)       # PIL's message
    assert rc == _lib.E_INVALID_ARG and b"must be > 0" in L.mrcnn_last_error()
    hw = _lib.i4([64, 32, 16, 8])
    rc = L.mrcnn_pyramid_roi_align_backward_plan(hw, hw, 1, 6, None, None, 4, 7, 1.0, None, 0, None)
    assert rc != 0
