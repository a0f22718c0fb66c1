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
    assert L.mrcnn_abi_version() == 6
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


def _prototypes():
    """name -> (return type, [parameter declarations]) parsed from include/mrcnn_b200.h."""
    text = open(os.path.join(ROOT, "include", "mrcnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    out = {}
    for ret, name, params in re.findall(r"MRCNN_API\s+([\w\s\*]+?)\b(mrcnn_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        params = " ".join(params.split())
        out[name] = (" ".join(ret.split()), [] if params == "void" else [p.strip() for p in params.split(",")])
    return out


def _c_kind(decl):
    if "*" in decl or "[" in decl or "mrcnn_stream_t" in decl:
        return "pointer"
    for word, kind in (("size_t", "size_t"), ("double", "double"), ("float", "float"), ("int32_t", "int"), ("int", "int")):
        if re.search(r"\b%s\b" % word, decl):
            return kind
    raise AssertionError("unclassified parameter: %r" % decl)


def _ctypes_kind(t):
    if t in (ctypes.c_void_p, ctypes.c_char_p) or issubclass(t, ctypes.Array):
        return "pointer"
    return {ctypes.c_int: "int", ctypes.c_float: "float", ctypes.c_double: "double", ctypes.c_size_t: "size_t"}[t]


def test_ctypes_signatures_match_the_header_prototypes():
    """A ctypes binding is unchecked at load time: a missing or mistyped argument would silently shift every later one.
    Every binding's argument kinds (pointer / int / float / size_t, in order) and return type are the header's."""
    from maskrcnn_b200 import _lib
    protos = _prototypes()
    assert sorted(protos) == sorted(_lib.SIGNATURES) == _declared()
    for name, (ret, params) in protos.items():
        res, args = _lib.SIGNATURES[name]
        assert [_c_kind(p) for p in params] == [_ctypes_kind(a) for a in args], name
        assert _ctypes_kind(res) == _c_kind(ret + " "), name


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """The boundary is a C ABI: the header compiles as strict C99 and a C program with no torch / Python in sight links
    against libmrcnn_b200.so, loads it and gets the documented answers from the calls that need no GPU."""
    import shutil
    import subprocess
    from maskrcnn_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "mrcnn_b200.h"
int main(void) {
    int hw[4] = {256, 128, 64, 32};
    if (mrcnn_abi_version() != MRCNN_ABI_VERSION) return 1;
    if (mrcnn_nms_workspace_bytes(6000) == 0) return 2;
    if (mrcnn_pyramid_roi_align_backward_workspace_bytes(hw, hw, 16, 8192, 14) == 0) return 3;
    /* a host pointer is refused: there is no CPU path behind the ABI */
    {
        float dets[5] = {0, 0, 1, 1, 0.5f};
        int64_t keep[1];
        int32_t count[1];
        int rc = mrcnn_nms(dets, 1, 0.5f, keep, count, NULL, 0, NULL);
        if (rc != MRCNN_E_NOT_DEVICE_PTR && rc != MRCNN_E_INVALID_ARG && rc != MRCNN_E_CUDA) return 4;
        if (strlen(mrcnn_last_error()) == 0) return 5;
    }
    if (mrcnn_crop_forward(NULL, 0, 1, 1, 1, MRCNN_NCHW, NULL, NULL, 1, 0.0f, 7, 7, NULL, MRCNN_NCHW, NULL) != MRCNN_E_INVALID_ARG) return 6;
    printf("abi %d ok\n", mrcnn_abi_version());
    return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:libmrcnn_b200.so", "-Wl,-rpath," + libdir], check=True, capture_output=True, text=True)
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert run.stdout.strip() == "abi 6 ok"


def test_device_code_is_sm_100a_only_and_holds_the_hot_path_kernels():
    """One architecture, no PTX to JIT for another, no multi-backend dispatch: every cubin in the library is sm_100a, and
    the kernels DESIGN.md section 3 names are in it (a library that lost one would fall back to nothing: there is no fallback)."""
    import shutil
    import subprocess
    from maskrcnn_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("no cuobjdump")
    elfs = re.findall(r"ELF file\s+\d+:\s+(\S+)", subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout)
    assert len(elfs) >= 9 and all(e.endswith(".sm_100a.cubin") for e in elfs), elfs
    ptx = subprocess.run(["cuobjdump", "-lptx", _lib.LIB_PATH], capture_output=True, text=True)
    assert "PTX file" not in ptx.stdout
    symbols = subprocess.run(["cuobjdump", "-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    kernels = set(re.findall(r"\.text\.(\w+)", symbols))
    for name in ("roialign_fwd_nhwc_col_kernel", "roialign_fwd_nhwc_kernel", "roialign_bwd_gather_kernel", "roialign_bwd_nhwc", "bwd_items_kernel",
                 "bwd_alloc_kernel", "crop_generic_kernel", "proposal_select_kernel", "proposal_lazy_nms_kernel", "proposal_mask_kernel",
                 "proposal_sweep_kernel", "detection_layer_kernel", "nms_prepare_small_kernel", "target_select_kernel", "rpn_match_kernel",
                 "rpn_pack_kernel", "full_masks_kernel", "paste_prepare_kernel"):
        assert any(name in k for k in kernels), "%s is not in libmrcnn_b200.so" % name
