"""Image-sharded multi-GPU plumbing for the RoI path.

Every op of the path is per-image, so ranks own disjoint contiguous blocks of images and run the whole path
locally; the ONLY exchange is one all-gather of the (padded, fixed-shape) detections so that every rank ends
up with the detections of all images in image order.  One process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch on GPUs; gloo on CPU for the host-logic tests)."""
import torch
import torch.distributed as dist


def shard_range(n_images, rank, world):
    """[begin, end) of the images owned by `rank`: contiguous, balanced (sizes differ by at most one)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n_images), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_sizes(n_images, world):
    return [shard_range(n_images, r, world)[1] - shard_range(n_images, r, world)[0] for r in range(world)]


def gather_detections(dets, counts, n_images=None, group=None):
    """dets [B_local, D, 6] fp32 (zero padded), counts [B_local] int32  ->  (dets [B_total, D, 6], counts [B_total])
    on every rank, in image order (rank-major = image order because shards are contiguous blocks).

    Counts travel inside the same buffer (one extra column) so that this is a single collective of
    ~2.4 KB per image — pure latency on NVSwitch.  Shards may be uneven (padded to the largest)."""
    if not dist.is_available() or not dist.is_initialized():
        return dets, counts
    world = dist.get_world_size(group)
    if world == 1:
        return dets, counts
    b_local, d = dets.shape[0], dets.shape[1]
    sizes = shard_sizes(n_images, world) if n_images is not None else None
    b_max = max(sizes) if sizes is not None else b_local
    if sizes is None:  # agree on the largest shard
        t = torch.tensor([b_local], device=dets.device, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        b_max = int(t.item())
    width = d * 6 + 1
    # pack: one concatenation, no host round trip
    send = torch.cat((dets.reshape(b_local, d * 6), counts.to(torch.float32).unsqueeze(1)), 1)
    if b_local < b_max:
        pad = send.new_zeros((b_max - b_local, width))
        pad[:, d * 6] = -1.0  # padding rows
        send = torch.cat((send, pad), 0)
    recv = torch.empty((world * b_max, width), dtype=torch.float32, device=dets.device)
    try:
        dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send.contiguous(), group=group)
        recv = torch.cat(parts, 0)
    if sizes is not None:
        # shard sizes are known on the host: drop the padding rows without looking at device data (no sync)
        if all(sz == b_max for sz in sizes):
            rows = recv
        else:
            keep = [r * b_max + i for r, sz in enumerate(sizes) for i in range(sz)]
            rows = recv.index_select(0, torch.tensor(keep, dtype=torch.int64).to(recv.device, non_blocking=True))
    else:
        rows = recv[recv[:, d * 6] >= 0]
    return rows[:, :d * 6].reshape(-1, d, 6).contiguous(), rows[:, d * 6].to(torch.int32)


def owner_rank(image, n_images, world):
    """Rank that owns `image` under shard_range() - the formula detection_collect_kernel evaluates on the device."""
    base, rem = divmod(int(n_images), int(world))
    cut = rem * (base + 1)
    return image // (base + 1) if image < cut else rem + (image - cut) // max(base, 1)


class DetectionExchange(object):
    """BASELINE configs[4] as three launches per rank and no library collective: detection layer FUSED with the all-gather of
    its results (mrcnn_detection_layer_exchange: every image's CTA stores its packed detections straight into the receive
    buffer of every rank over NVLink / NVSwitch peer mappings and the rank's last CTA raises a flag everywhere), the mask
    head's RoIAlign on the boxes the same kernel wrote, and mrcnn_detection_collect (per image: wait for the owner's flag,
    copy the row out).  Replaces detection_layer + five small torch ops + gather_detections' cat / all_gather_into_tensor /
    slicing.  Results are bit-identical to detection_layer + gather_detections (image order, zero padded).

    The peer mappings come from torch.distributed's symmetric memory (CUDA IPC / fabric handles exchanged through the
    process group's store) - plumbing; the stores, flags and waits are this library's kernels.  world size 1 (or no process
    group) uses a plain local buffer and the same kernels.  Every rank must call run() / collect() the same number of times."""

    MAX_WORLD = 8

    def __init__(self, n_images, max_instances, device=None, group=None):
        import ctypes
        from . import _lib
        self._lib = _lib
        self.n_images, self.D = int(n_images), int(max_instances)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        live = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if live else 1
        self.rank = dist.get_rank(group) if live else 0
        if self.world > self.MAX_WORLD:
            raise ValueError("DetectionExchange serves one NVSwitch domain: at most %d ranks" % self.MAX_WORLD)
        self.begin, self.end = shard_range(self.n_images, self.rank, self.world)
        nfloat = _lib.lib.mrcnn_detection_exchange_bytes(self.world, self.n_images, self.D) // 4
        self._hdl = None
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm
            g = group if group is not None else dist.group.WORLD
            self.buf = symm.empty(nfloat, dtype=torch.float32, device=self.device)
            self._hdl = symm.rendezvous(self.buf, g)
            ptrs = [int(p) for p in self._hdl.buffer_ptrs]
        else:
            self.buf = torch.empty(nfloat, dtype=torch.float32, device=self.device)
            ptrs = [self.buf.data_ptr()]
        self.buf.zero_()
        self.state = torch.zeros(2, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group)           # nobody stores into a buffer that is still being zeroed
        self._peers = (ctypes.c_void_p * self.world)(*ptrs)
        b = self.end - self.begin
        self.dets = torch.empty((max(b, 1), self.D, 6), dtype=torch.float32, device=self.device)
        self.counts = torch.empty(max(b, 1), dtype=torch.int32, device=self.device)
        self.mask_boxes = torch.empty((max(b, 1) * self.D, 4), dtype=torch.float32, device=self.device)
        self.mask_box_ind = torch.empty(max(b, 1) * self.D, dtype=torch.int32, device=self.device)
        self.dets_all = torch.empty((self.n_images, self.D, 6), dtype=torch.float32, device=self.device)
        self.counts_all = torch.empty(self.n_images, dtype=torch.int32, device=self.device)
        self._ws = None

    def run(self, rois, probs, deltas, windows, min_confidence, nms_threshold, std=(0.1, 0.1, 0.2, 0.2), image_hw=(1024, 1024),
            ind_offset=None, ind_mod=None):
        """Detection layer of this rank's images [begin, end) + mask-head RoIs + the sending half of the exchange: one launch.
        rois [b,N,4], probs [b,N,NC], deltas [b,N,NC,4], windows [b,4] (b = end - begin).  Afterwards self.dets / self.counts
        hold the local detections, self.mask_boxes / self.mask_box_ind the mask head's RoIAlign arguments (image index of local
        image i = (ind_offset + i) % ind_mod; default: the local index)."""
        import numpy as np
        L = self._lib
        b, N = rois.shape[:2]
        if self.end == self.begin:
            return self.dets[:0], self.counts[:0]          # a rank without images sends nothing (nobody waits for it)
        if b != self.end - self.begin:
            raise ValueError("this rank owns images [%d, %d): expected a batch of %d" % (self.begin, self.end, self.end - self.begin))
        NC = probs.size(-1)
        for t, name in ((rois, "rois"), (probs, "probs"), (deltas, "deltas"), (windows, "windows")):
            if not (t.is_cuda and t.dtype is torch.float32 and t.is_contiguous()):
                raise TypeError("%s must be a contiguous float32 CUDA tensor" % name)
        if probs.shape != (b, N, NC) or deltas.shape != (b, N, NC, 4) or windows.shape != (b, 4) or rois.size(2) != 4:
            raise ValueError("rois [b,N,4], probs [b,N,NC], deltas [b,N,NC,4], windows [b,4]")
        ws_bytes = L.lib.mrcnn_detection_workspace_bytes(b, N)
        if self._ws is None or self._ws.numel() < ws_bytes:
            self._ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        L.check(L.lib.mrcnn_detection_layer_exchange(
            rois.data_ptr(), probs.data_ptr(), deltas.data_ptr(), windows.data_ptr(), b, N, NC, float(min_confidence or 0.0),
            float(nms_threshold), self.D, L.f4(np.float32(std)), float(image_hw[0]), float(image_hw[1]), self.dets.data_ptr(),
            self.counts.data_ptr(), self.mask_boxes.data_ptr(), self.mask_box_ind.data_ptr(), 0 if ind_offset is None else int(ind_offset),
            b if ind_mod is None else int(ind_mod), self._peers, self.world, self.rank, self.begin, self.n_images, self.state.data_ptr(),
            self._ws.data_ptr(), self._ws.numel(), torch.cuda.current_stream().cuda_stream))
        return self.dets, self.counts

    def collect(self):
        """The receiving half: (dets [n_images, D, 6], counts [n_images]) of ALL images in image order on this rank.  The
        tensors are re-used by the next collect()."""
        L = self._lib
        L.check(L.lib.mrcnn_detection_collect(self.buf.data_ptr(), self.world, self.n_images, self.D, self.state.data_ptr(),
                                              self.dets_all.data_ptr(), self.counts_all.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return self.dets_all, self.counts_all
