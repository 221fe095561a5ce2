"""Multi-GPU data path: images are sharded over ranks (one process per GPU, weights replicated); the ONLY collective is a
fixed-shape all-gather of the per-image detections over NCCL/NVLink (gloo in the CPU tests).

It replaces the reference's pickled, variable-length ``all_gather`` of whole-dataset prediction dicts
(maskrcnn_benchmark/utils/comm.py:61-102, engine/inference.py:293-312): after the top-k cut every image has at most
``max_out`` rows of (x1, y1, x2, y2, score, label), so the exchange is ``[B_local, max_out, 6]`` fp32 + ``[B_local]``
counts per rank (24.6 KB at B_local = 8) — latency-bound, never bandwidth-bound.
"""
import torch
import torch.distributed as dist


def shard_indices(num_images, rank, world):
    """image i -> rank i % world (DistributedSampler order, data/build.py:196-198)."""
    return list(range(rank, num_images, world))


def all_gather_detections(det, num, group=None):
    """det [B_local, max_out, 6], num [B_local] (device or CPU tensors) -> (det_all [world*B_local, max_out, 6],
    num_all [world*B_local]) in RANK-MAJOR order; identity when no process group is initialised."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return det, num
    world = dist.get_world_size(group)
    det_all = torch.empty((world * det.shape[0],) + tuple(det.shape[1:]), dtype=det.dtype, device=det.device)
    num_all = torch.empty((world * num.shape[0],), dtype=num.dtype, device=num.device)
    dist.all_gather_into_tensor(det_all, det.contiguous(), group=group)
    dist.all_gather_into_tensor(num_all, num.contiguous(), group=group)
    return det_all, num_all


def unshard(det_all, num_all, num_images, world):
    """rank-major gathered rows -> original image order for images sharded with ``shard_indices``."""
    per = det_all.shape[0] // world
    order = []
    for r in range(world):
        order += [(i, r * per + k) for k, i in enumerate(shard_indices(num_images, r, world))]
    order.sort()
    idx = torch.tensor([j for _, j in order], dtype=torch.long, device=det_all.device)
    return det_all.index_select(0, idx), num_all.index_select(0, idx)
