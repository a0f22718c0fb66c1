"""Drop-in replacement for the reference's `maskrcnn` extension package (c++ext/maskrcnn/__init__.py),
the name model.py:25 imports.  Same two public names, same signatures; backed by libmrcnn_b200.so."""
from maskrcnn_b200.ops import CropFunction, nms  # noqa: F401

__all__ = ["nms", "CropFunction"]
