"""bindings/_C: the pybind11 module with the reference's native entry points (c++ext/maskrcnn/csrc/vision.cpp:11-15) over
libmrcnn_b200.so.  CPU part: it builds, loads, exports exactly the reference's three names with the reference's argument
lists, and refuses CPU tensors.  GPU part: called the way c++ext/maskrcnn/__init__.py:21-57 calls `_C` (caller-allocated
`crops` / `grads_image`, resize_ inside), results against the oracle and the reference-generated goldens."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle
from helpers import golden, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIND = os.path.join(ROOT, "bindings")


@pytest.fixture(scope="module")
def C():
    if not glob.glob(os.path.join(BIND, "_C*.so")):
        subprocess.check_call(["bash", os.path.join(BIND, "build.sh")])
    if BIND not in sys.path:
        sys.path.insert(0, BIND)
    import _C
    return _C


def test_binding_exports_the_reference_names(C):
    for name in ("nms", "crop_forward", "crop_backward"):
        assert callable(getattr(C, name))
    assert sorted(n for n in dir(C) if not n.startswith("_")) == ["crop_backward", "crop_forward", "nms"]
    # the argument lists of nms.h:15 / crop.h:14-22 / crop.h:36-41, as pybind prints them
    assert C.nms.__doc__.count("torch.Tensor") == 2
    assert C.crop_forward.__doc__.split("->")[0].count("torch.Tensor") == 4 and "-> None" in C.crop_forward.__doc__
    assert C.crop_backward.__doc__.split("->")[0].count("torch.Tensor") == 4 and "-> None" in C.crop_backward.__doc__


def test_binding_has_no_cpu_path(C):
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        C.nms(torch.zeros(4, 5), 0.5)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        C.crop_forward(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4), torch.zeros(1, dtype=torch.int32), 0.0, 7, 7, torch.zeros(1))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        C.crop_backward(torch.zeros(1, 4, 7, 7), torch.zeros(1, 4), torch.zeros(1, dtype=torch.int32), torch.zeros(1, 4, 8, 8))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_binding_nms_golden(C, tag):
    g = golden()
    keep = C.nms(dev(g[f"nms_{tag}_in_dets"]), float(g[f"nms_{tag}_in_thr"]))
    assert keep.is_cuda and keep.dtype == torch.int64
    np.testing.assert_array_equal(keep.cpu().numpy(), g[f"nms_{tag}_out_keep"])
    empty = C.nms(torch.zeros((0, 5), device="cuda"), 0.5)
    assert empty.numel() == 0 and not empty.is_cuda and empty.dtype == torch.int64       # nms.h:20-21


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_binding_crop_golden_like_the_reference_wrapper(C, tag):
    """c++ext/maskrcnn/__init__.py:32-57 verbatim in its calling convention: crops = zeros_like(image) is handed in and comes
    back resized; grad_image = zeros_like(grad).resize_(im_size)."""
    g = golden()
    image, boxes, ind = dev(g[f"crop_{tag}_in_image"]), dev(g[f"crop_{tag}_in_boxes"]), dev(g[f"crop_{tag}_in_ind"])
    want = g[f"crop_{tag}_out_crops"]
    crops = torch.zeros_like(image)                                                       # __init__.py:36
    C.crop_forward(image, boxes, ind, float(g[f"crop_{tag}_in_ev"]), want.shape[2], want.shape[3], crops)
    assert tuple(crops.shape) == want.shape
    np.testing.assert_array_equal(crops.cpu().numpy(), want)
    grad = dev(g[f"crop_{tag}_in_grads"]).contiguous()                                    # __init__.py:51
    grad_image = torch.zeros_like(grad).resize_(*image.size())                            # __init__.py:52
    grad_image.fill_(7.0)                                                                 # the callee overwrites (crop_cuda.cu:285)
    C.crop_backward(grad, boxes, ind, grad_image)
    assert rel_err(grad_image.cpu().numpy(), g[f"crop_{tag}_out_gimage"]) <= 1e-5


@pytest.mark.gpu
def test_binding_matches_oracle_at_head_size(C):
    """One FPN level at model size: 300 boxes x 256 channels on a 64x64 map, 7x7 and 14x14, NCHW and channels-last."""
    from maskrcnn_b200 import synth
    rng = np.random.default_rng(4)
    img = rng.standard_normal((2, 256, 64, 64), dtype=np.float32)
    boxes = synth.random_rois(300, 9, image=64.0, min_size=4, max_size=60)
    ind = rng.integers(0, 2, 300).astype(np.int32)
    for pool in (7, 14):
        want = oracle.crop_forward(img, boxes, ind, pool, pool, 0.0)
        gnp = rng.standard_normal(want.shape, dtype=np.float32)
        wg = oracle.crop_backward(gnp, boxes, ind, img.shape)
        for cl in (False, True):
            image = dev(img).contiguous(memory_format=torch.channels_last) if cl else dev(img)
            crops = torch.zeros(1, device="cuda")
            C.crop_forward(image, dev(boxes), dev(ind), 0.0, pool, pool, crops)
            np.testing.assert_array_equal(crops.cpu().numpy(), want)
            gi = torch.empty_like(image)
            gr = dev(gnp).contiguous(memory_format=torch.channels_last) if cl else dev(gnp)
            C.crop_backward(gr, dev(boxes), dev(ind), gi)
            assert rel_err(gi.cpu().numpy(), wg) <= 1e-5
    d5 = np.concatenate([boxes * 64, synth.unique_scores(300, 5)[:, None]], 1).astype(np.float32)
    np.testing.assert_array_equal(C.nms(dev(d5), 0.5).cpu().numpy(), oracle.nms(d5, 0.5))
