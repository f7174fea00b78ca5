from torch.nn import functional as F


def sem_seg_postprocess(result, img_size, output_height, output_width):
    """Crop the padding away, then bilinear-resize to the requested size
    (detectron2.modeling.postprocessing.sem_seg_postprocess semantics)."""
    result = result[:, : img_size[0], : img_size[1]].expand(1, -1, -1, -1)
    result = F.interpolate(result, size=(output_height, output_width), mode="bilinear", align_corners=False)[0]
    return result
