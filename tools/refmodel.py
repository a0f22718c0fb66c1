"""HARNESS (not product, not oracle): loads the reference's UNMODIFIED model.py / data.py / config.py / utils.py and lets a
caller choose which `maskrcnn` package they import (model.py:25 `import maskrcnn`): this repo's drop-in (the product under
test) or the reference's own compiled CPU extension (oracle/_ref, the checker).

The four Python files and images/car58a54312d.jpg are copied by __graft_entry__.build() from /root/reference into the
git-ignored baseline/_ref/ so that they travel to the GPU box with the snapshot (/root/reference does not exist there).
Nothing is patched in the copies.  Shims (none touches arithmetic):
  * skimage, matplotlib: absent from the image -> stub modules (skimage.io.imread -> PIL);
  * scipy.misc.imresize (utils.py:77) is gone from scipy -> what it was: PIL's bilinear resize of the uint8 image.

Used by tests/test_gpu_reference_model.py and bench.py's `predict_flow` leg."""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ("model.py", "data.py", "config.py", "utils.py")
IMAGE = os.path.join("images", "car58a54312d.jpg")


def ref_root():
    """Where the reference's Python lies: baseline/_ref (travels to the GPU box) or /root/reference (build container)."""
    for root in (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("REF_ROOT", "/root/reference")):
        if all(os.path.exists(os.path.join(root, f)) for f in FILES):
            return root
    return None


def available():
    return ref_root() is not None


def image_path():
    root = ref_root()
    p = os.path.join(root, IMAGE) if root else None
    return p if p and os.path.exists(p) else None


def pil_imresize(image, size, interp="bilinear"):
    from PIL import Image
    return np.asarray(Image.fromarray(image).resize((int(size[1]), int(size[0])), Image.BILINEAR))


def pil_imread(path):
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGB"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def load(maskrcnn_module):
    """Fresh copies of the reference modules, importing `maskrcnn_module` as `maskrcnn`.  Returns a namespace with
    .model, .data, .config, .utils, .maskrcnn."""
    root = ref_root()
    if root is None:
        raise RuntimeError("the reference's model.py is not here (run __graft_entry__.build() where /root/reference exists)")
    import scipy  # noqa: F401
    stubs = {
        "skimage": _stub("skimage"),
        "skimage.io": _stub("skimage.io", imread=pil_imread),
        "skimage.color": _stub("skimage.color"),
        "skimage.measure": _stub("skimage.measure", find_contours=None),
        "matplotlib": _stub("matplotlib"),
        "matplotlib.pyplot": _stub("matplotlib.pyplot", switch_backend=lambda *a, **k: None),
        "matplotlib.patches": _stub("matplotlib.patches", Polygon=None),
        "scipy.misc": _stub("scipy.misc", imresize=pil_imresize),
    }
    names = list(stubs) + ["maskrcnn", "utils", "data", "config", "model"]
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update(stubs)
    sys.modules["maskrcnn"] = maskrcnn_module
    mods = {}
    try:
        for name in ("config", "utils", "data", "model"):
            spec = importlib.util.spec_from_file_location(name, os.path.join(root, name + ".py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[name] = m          # the reference modules import each other by bare name
            spec.loader.exec_module(m)
            mods[name] = m
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    mods["utils"].scipy.misc = stubs["scipy.misc"]      # utils.py:77 resolves scipy.misc.imresize at call time
    return types.SimpleNamespace(maskrcnn=maskrcnn_module, **mods)


def make_config(ref, gpu, image_dim=1024, train=False, **overrides):
    base = ref.config.CocoConfig if train else ref.config.CocoInferenceConfig
    attrs = {"GPU_COUNT": 1 if gpu else 0, "IMAGE_MIN_DIM": image_dim, "IMAGE_MAX_DIM": image_dim}
    if train:
        attrs["IMAGES_PER_GPU"] = 1
    attrs.update(overrides)
    return type("HarnessConfig", (base,), attrs)()


def make_model(ref, cfg, seed):
    """MaskRCNN with random-init weights (the weights file is absent, models/README) under torch.manual_seed(seed), always
    initialised on the CPU so that a CPU and a GPU instance hold the same parameters."""
    import torch
    torch.manual_seed(seed)
    gpu = cfg.GPU_COUNT
    model = ref.model.MaskRCNN(model_dir=tempfile.mkdtemp(prefix="mrcnn_logs_"), config=cfg)
    if gpu:
        model = model.cuda()
    return model


def train_inputs(ref, cfg, seed, n_gt=6, gt_boxes=None):
    """One synthetic training sample in the format CocoMaskRCNNDataset.__getitem__ returns (data.py:710-737), batched by 1
    like the DataLoader does (model.py:1528-1532): [images, rpn_match, rpn_bbox, gt_class_ids, gt_boxes, gt_masks].
    gt_boxes: optional [G,4] pixel boxes (rectangular instances); random ones otherwise."""
    import torch
    rng = np.random.default_rng(seed)
    size = int(cfg.IMAGE_MAX_DIM)
    image = torch.from_numpy(rng.standard_normal((3, size, size)).astype(np.float32))
    if gt_boxes is None:
        boxes = np.zeros((n_gt, 4), np.float32)
        for k in range(n_gt):
            h, w = rng.integers(size // 8, size // 2, 2)
            y, x = rng.integers(0, size - h), rng.integers(0, size - w)
            boxes[k] = [y, x, y + h, x + w]
    else:
        boxes = np.asarray(gt_boxes, np.float32).reshape(-1, 4)
        n_gt = len(boxes)
    masks = np.zeros((n_gt, size, size), np.float32)
    for k in range(n_gt):
        y1, x1, y2, x2 = (int(v) for v in boxes[k])
        masks[k, y1:y2, x1:x2] = 1.0
    class_ids = rng.integers(1, 81, n_gt).astype(np.int32)
    anchors = ref.utils.create_pyramid_anchors(cfg.RPN_ANCHOR_SCALES, cfg.RPN_ANCHOR_RATIOS, cfg.BACKBONE_SHAPES,
                                               cfg.BACKBONE_STRIDES, cfg.RPN_ANCHOR_STRIDE)
    np.random.seed(seed)
    rpn_match, rpn_bbox = ref.data.rpn_samples(anchors, class_ids, boxes.astype(np.int32), cfg)
    return [image.unsqueeze(0), torch.from_numpy(np.asarray(rpn_match))[:, None].unsqueeze(0).int(),
            torch.from_numpy(np.asarray(rpn_bbox)).float().unsqueeze(0), torch.from_numpy(class_ids).unsqueeze(0),
            torch.from_numpy(boxes).unsqueeze(0), torch.from_numpy(masks).unsqueeze(0)]


def proposals_of(model, image):
    """The proposals [K,4] (pixels, numpy) the network makes for `image` [1,3,S,S] - the same ones MaskRCNN.extract will see
    (batch norm runs in eval mode there too, model.py:1218-1224).  The random-init harness takes its ground-truth boxes from
    them: an untrained RPN's proposals overlap no random ground truth, and a training step without a positive RoI never
    reaches the RoIAlign heads (model.py:1272-1282)."""
    import torch
    was = model.training
    dev = next(model.parameters()).device
    with torch.no_grad():
        model.eval()
        p2, p3, p4, p5, p6 = model.fpn(image.to(dev))
        _, rpn_class, rpn_bbox = model.rpn_detect([p2, p3, p4, p5, p6])
        rois = model.rpn_refine(rpn_class, rpn_bbox)
    model.train(was)
    size = float(model.config.IMAGE_SHAPE[0])
    return rois[0].detach().cpu().numpy() * size


def gt_from_proposals(rois_px, n_gt=6, min_side=24):
    """n_gt well separated proposals, rounded outwards to whole pixels, as ground-truth boxes."""
    b = np.asarray(rois_px, np.float32)
    side = np.minimum(b[:, 2] - b[:, 0], b[:, 3] - b[:, 1])
    pick = []
    for i in np.argsort(-side):
        if side[i] < min_side or len(pick) == n_gt:
            break
        if all(_iou(b[i], b[j]) < 0.3 for j in pick):
            pick.append(i)
    assert pick, "no proposal is large enough to serve as ground truth"
    g = b[pick]
    return np.stack([np.floor(g[:, 0]), np.floor(g[:, 1]), np.ceil(g[:, 2]), np.ceil(g[:, 3])], 1).astype(np.float32)


def _iou(a, b):
    y1, x1, y2, x2 = max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3])
    inter = max(0.0, y2 - y1) * max(0.0, x2 - x1)
    ua = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return inter / ua if ua > 0 else 0.0


def tolerate_empty_boxes(ref):
    """Random-init heads give some detections a box that the window clip and the rounding (model.py:1429-1432) flatten to
    an empty rectangle, and the reference's data.full_masks raises on those from PIL ("height and width must be > 0",
    data.py:295).  With trained weights this does not happen; for the random-init harness the reference's function is run
    on the other rows and the empty boxes get the empty mask (what this repo's full_masks defines for them).  Returns the
    original function so that the caller can restore it."""
    import torch
    orig = ref.data.full_masks

    def full_masks(class_ids, boxes, masks, height, width):
        ok = ((boxes[:, 2] - boxes[:, 0]).int() > 0) & ((boxes[:, 3] - boxes[:, 1]).int() > 0)
        if bool(ok.all()):
            return orig(class_ids, boxes, masks, height, width)
        out = torch.zeros((len(boxes), height, width), dtype=torch.uint8, device=boxes.device)
        if bool(ok.any()):
            part = orig(class_ids[ok], boxes[ok], masks[ok], height, width)
            out = out.to(part.dtype)
            out[ok] = part
        return out
    ref.data.full_masks = full_masks
    return orig
