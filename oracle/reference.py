"""TEST INFRASTRUCTURE ONLY — loads the UNMODIFIED reference Python (model.py / data.py / config.py /
utils.py under /root/reference) on top of the reference's own compiled CPU extension (oracle/_ref).

Works only where /root/reference exists (the build container).  It is used by
tests/golden/make_golden.py to generate committed golden vectors and by tests/test_oracle_vs_ref.py
to pin the C restatement; nothing that runs on the GPU box imports it.

Shims (none touches arithmetic):
  * `skimage`, `matplotlib`, `scipy.misc` are absent/removed in this image -> empty stub modules.
  * c++ext/maskrcnn/__init__.py:25-57 `CropFunction` is a legacy (non-static) autograd Function that
    torch 2.x refuses to run -> the same forward/backward bodies expressed as a static Function.
"""
import contextlib
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("REF_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_SO_DIR = os.path.join(_HERE, "_ref")


def available():
    return os.path.isdir(REF_ROOT) and ref_C_available()


def ref_C_available():
    return os.path.isdir(_REF_SO_DIR) and any(f.startswith("ref_C") and f.endswith(".so")
                                              for f in os.listdir(_REF_SO_DIR))


_ref_C = None


def ref_C():
    """The reference's pybind module (vision.cpp:11-15): nms, crop_forward, crop_backward."""
    global _ref_C
    if _ref_C is None:
        import torch  # noqa: F401  (must be imported before the extension, __init__.py:10-14)
        if _REF_SO_DIR not in sys.path:
            sys.path.insert(0, _REF_SO_DIR)
        import ref_C as m
        _ref_C = m
    return _ref_C


@contextlib.contextmanager
def quiet_stdout():
    """crop_cpu.cpp:163 printf()s on every forward; silence fd 1 while the reference runs."""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        yield
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)


def make_maskrcnn_shim():
    """A module object with the operator API of c++ext/maskrcnn/__init__.py backed by ref_C."""
    import torch
    C = ref_C()
    shim = types.ModuleType("maskrcnn")

    def nms(dets, threshold):  # __init__.py:21-22
        return C.nms(dets, threshold)

    class _Crop(torch.autograd.Function):
        @staticmethod
        def forward(ctx, image, boxes, box_ind, ch, cw, ev):  # __init__.py:32-45
            crops = torch.zeros_like(image)
            with quiet_stdout():
                C.crop_forward(image, boxes, box_ind, ev, ch, cw, crops)
            ctx.im_size = image.size()
            ctx.save_for_backward(boxes, box_ind)
            return crops

        @staticmethod
        def backward(ctx, grad_outputs):  # __init__.py:48-57
            boxes, box_ind = ctx.saved_tensors
            grad_outputs = grad_outputs.contiguous()
            grad_image = torch.zeros_like(grad_outputs).resize_(*ctx.im_size)
            C.crop_backward(grad_outputs, boxes, box_ind, grad_image)
            return grad_image, None, None, None, None, None

    class CropFunction(object):  # __init__.py:25-30 constructor signature
        def __init__(self, crop_height, crop_width, extrapolation_value=0):
            self.crop_height = crop_height
            self.crop_width = crop_width
            self.extrapolation_value = extrapolation_value

        def __call__(self, image, boxes, box_ind):
            return _Crop.apply(image, boxes, box_ind, self.crop_height, self.crop_width,
                               self.extrapolation_value)

    shim.nms = nms
    shim.CropFunction = CropFunction
    shim._C = C
    return shim


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


_loaded = None


def load():
    """Returns a namespace with .model, .data, .config, .utils (the reference modules) and .maskrcnn."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference not available here (needs /root/reference and oracle/_ref)")
    import scipy  # noqa: F401
    stubs = {
        "skimage": _stub("skimage"),
        "skimage.io": _stub("skimage.io"),
        "skimage.color": _stub("skimage.color"),
        "skimage.measure": _stub("skimage.measure", find_contours=None),
        "matplotlib": _stub("matplotlib"),
        "matplotlib.pyplot": _stub("matplotlib.pyplot", switch_backend=lambda *a, **k: None),
        "matplotlib.patches": _stub("matplotlib.patches", Polygon=None),
        "scipy.misc": _stub("scipy.misc"),
    }
    shim = make_maskrcnn_shim()
    saved = {k: sys.modules.get(k) for k in list(stubs) + ["maskrcnn", "utils", "data", "config", "model"]}
    sys.modules.update(stubs)
    sys.modules["maskrcnn"] = shim
    mods = {}
    try:
        for name in ("config", "utils", "data", "model"):
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, name + ".py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[name] = m  # reference modules import each other by bare name
            spec.loader.exec_module(m)
            mods[name] = m
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _loaded = types.SimpleNamespace(maskrcnn=shim, **mods)
    return _loaded
