"""ImageList (maskrcnn_benchmark/structures/image_list.py): a zero-padded batch tensor + the original (height, width) of
every image.  `to_image_list` accepts what the reference's accepts: an ImageList, one tensor ([C,H,W] or [B,C,H,W]) or a
sequence of [C,H,W] tensors of different sizes (padded bottom/right to a multiple of `size_divisible`)."""
import torch


class ImageList(object):
    def __init__(self, tensors, image_sizes):
        self.tensors = tensors
        self.image_sizes = image_sizes

    def to(self, *args, **kwargs):
        return ImageList(self.tensors.to(*args, **kwargs), self.image_sizes)


def _round_up(v, m):
    return v if m <= 0 else -(-v // m) * m


def to_image_list(tensors, size_divisible=0):
    if isinstance(tensors, ImageList):
        return tensors
    if torch.is_tensor(tensors):
        batch = tensors if tensors.dim() == 4 else tensors[None]
        return ImageList(batch, [tuple(im.shape[-2:]) for im in batch])
    sizes = [tuple(im.shape[-2:]) for im in tensors]
    C = max(im.shape[0] for im in tensors)
    H = _round_up(max(s[0] for s in sizes), size_divisible)
    W = _round_up(max(s[1] for s in sizes), size_divisible)
    batch = tensors[0].new_zeros((len(tensors), C, H, W))
    for dst, im in zip(batch, tensors):
        dst[: im.shape[0], : im.shape[1], : im.shape[2]] = im
    return ImageList(batch, sizes)
