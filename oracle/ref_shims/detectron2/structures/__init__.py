import torch
from torch.nn import functional as F


class ImageList:
    def __init__(self, tensor, image_sizes):
        self.tensor = tensor
        self.image_sizes = image_sizes

    @staticmethod
    def from_tensors(tensors, size_divisibility=0, pad_value=0.0):
        image_sizes = [(im.shape[-2], im.shape[-1]) for im in tensors]
        max_h = max(s[0] for s in image_sizes)
        max_w = max(s[1] for s in image_sizes)
        if size_divisibility > 1:
            s = size_divisibility
            max_h = (max_h + s - 1) // s * s
            max_w = (max_w + s - 1) // s * s
        batched = tensors[0].new_full((len(tensors), tensors[0].shape[0], max_h, max_w), pad_value)
        for i, im in enumerate(tensors):
            batched[i, :, : im.shape[-2], : im.shape[-1]].copy_(im)
        return ImageList(batched, image_sizes)


class Boxes:
    pass


class Instances:
    pass


class BitMasks:
    pass
