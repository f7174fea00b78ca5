"""Panoptic and open-panoptic inference on the head outputs: the reference's
`MaskFormer.panoptic_inference` (mask2former/maskformer_model.py:394-481) restated without its per-query host round trips.

The reference walks the kept queries one by one with three `.item()` synchronisations each and `panoptic_seg[mask] = id`
scatter writes; here the per-query areas come from one batched pass (bincount of the arg-max map, thresholded mask sums), the
sequential id assignment (overlap filter, stuff merging) runs on those few numbers on the host, and the panoptic map is ONE
gather through a query -> segment-id table.  The open-panoptic branch (:454-481) thresholds the RbA score -- which the fused
score kernel already produced -- and labels its connected components exactly as the reference does (3x3 morphological open /
close + 4-connected components with OpenCV on the host)."""
import numpy as np
import torch
import torch.nn.functional as F


def panoptic_inference(mask_cls, mask_pred, num_classes, object_mask_threshold, overlap_threshold, thing_ids,
                       open_panoptic=False, ood_threshold=-0.1, pixel_min=300, return_ood_pred=False, ood_mask=None):
    """mask_cls (Q,K+1) class logits, mask_pred (Q,H,W) mask logits at the output resolution.
    Returns (panoptic_seg (H,W) int32, segments_info[, ood_mask]) like the reference.
    ood_mask: the (H,W) RbA score if the caller already has it (engine output), else it is computed here."""
    probs = F.softmax(mask_cls, dim=-1)
    scores, labels = probs.max(-1)
    mask_pred = mask_pred.sigmoid()
    keep = labels.ne(num_classes) & (scores > object_mask_threshold)
    cur_scores, cur_classes, cur_masks = scores[keep], labels[keep], mask_pred[keep]
    h, w = mask_pred.shape[-2:]
    panoptic_seg = torch.zeros((h, w), dtype=torch.int32, device=mask_pred.device)
    segments_info = []
    current_segment_id = 0
    n = int(cur_masks.shape[0])
    if n == 0:                                            # maskformer_model.py:414-416 (returns before the open-panoptic branch)
        return panoptic_seg, segments_info
    cur_prob_masks = cur_scores.view(-1, 1, 1) * cur_masks
    cur_mask_ids = cur_prob_masks.argmax(0)                                                   # (H,W)
    own = cur_masks.gather(0, cur_mask_ids.unsqueeze(0))[0] >= 0.5                            # pixel kept by its arg-max query
    mask_area = torch.bincount(cur_mask_ids.flatten(), minlength=n)                           # (cur_mask_ids == k).sum()
    original_area = (cur_masks >= 0.5).flatten(1).sum(1)                                      # (cur_masks[k] >= 0.5).sum()
    inter = torch.bincount(cur_mask_ids.flatten(), weights=own.flatten().to(torch.float64), minlength=n)
    stats = torch.stack([mask_area.double(), original_area.double(), inter, cur_classes.double()]).cpu().numpy()   # one D2H
    thing = set(int(t) for t in thing_ids)
    lut = np.zeros(n, dtype=np.int32)                     # query -> segment id (0: dropped)
    stuff_memory = {}
    for k in range(n):
        ma, oa, it, pc = stats[0, k], stats[1, k], stats[2, k], int(stats[3, k])
        if ma > 0 and oa > 0 and it > 0:
            if ma / oa < overlap_threshold:
                continue
            isthing = pc in thing
            if not isthing:                               # merge stuff regions
                if pc in stuff_memory:
                    lut[k] = stuff_memory[pc]
                    continue
                stuff_memory[pc] = current_segment_id + 1
            current_segment_id += 1
            lut[k] = current_segment_id
            segments_info.append({"id": current_segment_id, "isthing": bool(isthing), "category_id": pc})
    lut_t = torch.from_numpy(lut).to(mask_pred.device)
    panoptic_seg = torch.where(own, lut_t[cur_mask_ids], panoptic_seg)
    if open_panoptic:
        if ood_mask is None:                              # :456-458, the third copy of get_RbA
            semseg = torch.einsum("qc,qhw->chw", probs[..., :-1], mask_pred)
            ood_mask = -(semseg.tanh()).sum(0)
        import cv2
        binary = (ood_mask > ood_threshold).cpu().numpy().astype(np.uint8)
        binary = cv2.morphologyEx(binary, cv2.MORPH_OPEN, np.ones((3, 3), np.uint8))
        binary = cv2.morphologyEx(binary, cv2.MORPH_CLOSE, np.ones((3, 3), np.uint8))
        num_labels, labels_im = cv2.connectedComponents(binary, connectivity=4)
        labels_im = torch.from_numpy(labels_im).to(mask_pred.device)
        free = panoptic_seg == 0                                                              # evaluated once per component in the
        if num_labels > 1:                                                                    # reference; components are disjoint
            counts = torch.bincount(labels_im.flatten(), weights=free.flatten().to(torch.float64), minlength=num_labels).cpu().numpy()
            comp_lut = np.zeros(num_labels, dtype=np.int32)
            for i in range(1, num_labels):
                if counts[i] < pixel_min:
                    continue
                current_segment_id += 1
                comp_lut[i] = current_segment_id
                segments_info.append({"id": current_segment_id, "isthing": True, "category_id": 255})
            comp_t = torch.from_numpy(comp_lut).to(mask_pred.device)[labels_im.long()]
            panoptic_seg = torch.where(free & (comp_t > 0), comp_t, panoptic_seg)
        if return_ood_pred:
            return panoptic_seg, segments_info, ood_mask
    return panoptic_seg, segments_info
