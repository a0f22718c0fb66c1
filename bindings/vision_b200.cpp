// bindings/vision_b200.cpp - the reference's native boundary, `maskrcnn._C`, on top of libmrcnn_b200.so.
//
// The reference binds three functions with pybind11 (c++ext/maskrcnn/csrc/vision.cpp:11-15):
//     nms(const Tensor& dets, float threshold) -> Tensor                                           nms.h:15-30
//     crop_forward(image, boxes, box_index, extrapolation_value, crop_height, crop_width, Tensor& crops)   crop.h:14-34
//     crop_backward(grads, boxes, box_index, Tensor& grads_image)                                  crop.h:36-53
// and dispatches each on `tensor.type().is_cuda()` to cpu/*.cpp or cuda/*.cu.  This file is that module with the CUDA
// branch answered by the C ABI of include/mrcnn_b200.h (what cuda/nms_cuda.cu:77-137 and cuda/crop_cuda.cu:228-297 did):
// same names, same argument lists, the same ownership rules - the CALLER allocates `crops` / `grads_image`, the callee
// resize_()s `crops` to [N, C, crop_height, crop_width] (crop_cuda.cu:250) and fills every element; `grads_image`
// arrives pre-sized (c++ext/maskrcnn/__init__.py:52) and is overwritten (crop_cuda.cu:285).  There is no CPU branch:
// a CPU tensor raises, like the reference built without WITH_CUDA does for CUDA tensors (nms.h:24, crop.h:28,47).
//
// Built by bindings/build.sh into bindings/_C*.so (git-ignored, shipped to the GPU box); tests/test_gpu_binding.py drives
// it with the reference's own Python wrapper semantics (c++ext/maskrcnn/__init__.py:32-57).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include "../include/mrcnn_b200.h"

namespace {

void ck(int rc) {
    if (rc != MRCNN_OK) AT_ERROR("libmrcnn_b200: ", mrcnn_last_error());
}

void require_cuda_f32(const at::Tensor& t, const char* name) {
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor: this build has no CPU path");
    TORCH_CHECK(t.scalar_type() == at::kFloat, name, " must be float32");
}

mrcnn_stream_t stream_of(const at::Tensor& t) {
    return (mrcnn_stream_t)at::cuda::getCurrentCUDAStream(t.get_device()).stream();
}

// nms.h:15-30 -> nms_cuda (nms_cuda.cu:77-137): [N,5] (y1,x1,y2,x2,score) -> ascending int64 indices of the survivors, on
// the device; an empty input returns an empty CPU tensor (nms.h:20-21).  Suppression rule: the CPU implementation's
// `IoU >= threshold` (nms_cpu.cpp:65), the parity target.
at::Tensor nms(const at::Tensor& dets_, const float threshold) {
    require_cuda_f32(dets_, "dets");
    if (dets_.numel() == 0) return at::empty({0}, dets_.options().dtype(at::kLong).device(at::kCPU));
    TORCH_CHECK(dets_.dim() == 2 && dets_.size(1) == 5, "dets must be [N,5]");
    c10::cuda::CUDAGuard guard(dets_.device());
    auto dets = dets_.contiguous();
    const int n = (int)dets.size(0);
    auto keep = at::empty({n}, dets.options().dtype(at::kLong));
    auto count = at::empty({1}, dets.options().dtype(at::kInt));
    const size_t ws_bytes = mrcnn_nms_workspace_bytes(n);
    auto ws = at::empty({(int64_t)ws_bytes}, dets.options().dtype(at::kByte));
    ck(mrcnn_nms(dets.data_ptr<float>(), n, threshold, keep.data_ptr<int64_t>(), count.data_ptr<int>(), ws.data_ptr(), ws_bytes,
                 stream_of(dets)));
    return keep.narrow(0, 0, count.item<int>());   // the one 4-byte read: K is data dependent (nms_cuda.cu:135-136)
}

// crop.h:14-34 -> crop_gpu_forward (crop_cuda.cu:228-262)
void crop_forward(const at::Tensor& image_, const at::Tensor& boxes_, const at::Tensor& box_index_, const float extrapolation_value,
                  const int crop_height, const int crop_width, at::Tensor& crops) {
    require_cuda_f32(image_, "image");
    require_cuda_f32(boxes_, "boxes");
    TORCH_CHECK(box_index_.is_cuda() && box_index_.scalar_type() == at::kInt, "box_index must be a CUDA int32 tensor");
    TORCH_CHECK(image_.dim() == 4 && boxes_.dim() == 2 && boxes_.size(1) == 4 && box_index_.numel() == boxes_.size(0),
                "image [B,C,H,W], boxes [N,4], box_index [N]");
    c10::cuda::CUDAGuard guard(image_.device());
    // a channels-last image keeps its layout (and takes the vectorised kernels); anything else is read as NCHW, the
    // reference's layout (crop_cuda.cu:238 calls .contiguous() and drops the result; here it is honoured)
    const bool cl = image_.is_contiguous(at::MemoryFormat::ChannelsLast) && !image_.is_contiguous();
    auto image = cl ? image_ : image_.contiguous();
    auto boxes = boxes_.contiguous();
    auto box_index = box_index_.contiguous();
    const int64_t n = boxes.size(0);
    if (!crops.is_cuda() || crops.scalar_type() != at::kFloat) crops = at::empty({0}, image.options());
    crops.resize_({n, image.size(1), crop_height, crop_width},
                  cl ? at::MemoryFormat::ChannelsLast : at::MemoryFormat::Contiguous);   // crop_cuda.cu:250; no zero_(): every element is written
    ck(mrcnn_crop_forward(image.data_ptr<float>(), (int)image.size(0), (int)image.size(1), (int)image.size(2), (int)image.size(3),
                          cl ? MRCNN_NHWC : MRCNN_NCHW, boxes.data_ptr<float>(), box_index.data_ptr<int>(), (int)n,
                          extrapolation_value, crop_height, crop_width, crops.data_ptr<float>(), cl ? MRCNN_NHWC : MRCNN_NCHW,
                          stream_of(image)));
}

// crop.h:36-53 -> crop_gpu_backward (crop_cuda.cu:264-297)
void crop_backward(const at::Tensor& grads_, const at::Tensor& boxes_, const at::Tensor& box_index_, at::Tensor& grads_image) {
    require_cuda_f32(grads_, "grads");
    require_cuda_f32(boxes_, "boxes");
    require_cuda_f32(grads_image, "grads_image");
    TORCH_CHECK(box_index_.is_cuda() && box_index_.scalar_type() == at::kInt, "box_index must be a CUDA int32 tensor");
    TORCH_CHECK(grads_.dim() == 4 && grads_image.dim() == 4 && grads_.size(1) == grads_image.size(1),
                "grads [N,C,h,w] and a pre-sized grads_image [B,C,H,W] (c++ext/maskrcnn/__init__.py:52)");
    c10::cuda::CUDAGuard guard(grads_.device());
    const bool gcl = grads_.is_contiguous(at::MemoryFormat::ChannelsLast) && !grads_.is_contiguous();
    auto grads = gcl ? grads_ : grads_.contiguous();
    const bool icl = grads_image.is_contiguous(at::MemoryFormat::ChannelsLast) && !grads_image.is_contiguous();
    TORCH_CHECK(icl || grads_image.is_contiguous(), "grads_image must be dense (NCHW or channels-last)");
    auto boxes = boxes_.contiguous();
    auto box_index = box_index_.contiguous();
    ck(mrcnn_crop_backward(grads.data_ptr<float>(), gcl ? MRCNN_NHWC : MRCNN_NCHW, boxes.data_ptr<float>(), box_index.data_ptr<int>(),
                           (int)boxes.size(0), (int)grads.size(2), (int)grads.size(3), grads_image.data_ptr<float>(),
                           (int)grads_image.size(0), (int)grads_image.size(1), (int)grads_image.size(2), (int)grads_image.size(3),
                           icl ? MRCNN_NHWC : MRCNN_NCHW, /*zero_fill=*/1, stream_of(grads)));   // crop_cuda.cu:285 grads_image.zero_()
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {   // the three names of vision.cpp:11-15
    m.def("nms", &nms, "non-maximum suppression");
    m.def("crop_forward", &crop_forward, "crop forward");
    m.def("crop_backward", &crop_backward, "crop backward");
}
