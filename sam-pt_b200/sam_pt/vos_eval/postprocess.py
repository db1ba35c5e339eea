"""The tail of the reference's VOS harness (sam_pt/vos_eval/eval.py:304-355) as one native call.

    logits = torch.stack([zeros] + pred_logits_list, dim=1)          # background channel            eval.py:304
    logits[:gt_ti, i + 1] = -1e8                                     # before the query frame        :321-322
    logits[gt_ti, i + 1] = where(nearest(gt_mask), 1e8, -1e8)        # GT overwrite                  :324-326
    probs = softmax(logits, dim=1)                                   #                               :327
    per frame: bilinear up-sampling if need_resize, flip, argmax     #                               :343-355

`vos_index_masks` returns the uint8 index masks of every frame, (T, Ho, Wo), ready for `mapper.remap_index_mask` + PNG
(eval.py:362-368, host side, unchanged).  INTEGRATION.md shows the three-line patch of eval.py that calls it.
"""
from __future__ import annotations

from ctypes import c_int
from typing import List, Sequence

import torch

from sampt_b200 import native


def vos_index_masks(pred_logits_list: Sequence[torch.Tensor], gt_mask_list: Sequence[torch.Tensor], gt_ti_list: Sequence[int],
                    target_hw, out_shape=None, need_resize: bool = False, flip: bool = False, device=None) -> torch.Tensor:
    """pred_logits_list: M tensors (T, H, W) float32 (the `logits` entries of SamPt.forward's output, any device);
    gt_mask_list: M tensors (Hg, Wg) {0,1}; gt_ti_list: M query frame indices; target_hw == (H, W);
    out_shape: original frame shape when need_resize.  -> uint8 (T, Ho, Wo) on the GPU."""
    dev = torch.device(device) if device is not None else next((l.device for l in pred_logits_list if l.is_cuda), None)
    if dev is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    logits = torch.stack([l.to(dev, torch.float32) for l in pred_logits_list], dim=0).contiguous()      # (M, T, H, W)
    M, T, H, W = logits.shape
    assert (H, W) == tuple(int(v) for v in target_hw)
    gt = torch.stack([g.to(dev, torch.float32) for g in gt_mask_list], dim=0).contiguous()
    ti = torch.as_tensor([int(t) for t in gt_ti_list], dtype=torch.int32, device=dev)
    Ho, Wo = (int(out_shape[0]), int(out_shape[1])) if need_resize else (H, W)
    out = torch.empty((T, Ho, Wo), dtype=torch.uint8, device=dev)
    ctx = native.get_context(dev)
    with torch.cuda.device(dev):
        native.check(native.lib().sampt_vos_index_masks(
            ctx.handle, native.ptr(logits), c_int(M), c_int(T), c_int(H), c_int(W), native.ptr(gt), c_int(gt.shape[1]), c_int(gt.shape[2]),
            native.ptr(ti), c_int(Ho), c_int(Wo), c_int(1 if need_resize else 0), c_int(1 if flip else 0), native.ptr(out),
            native.stream_ptr()), "vos_index_masks")
    return out
