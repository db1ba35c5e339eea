"""ORACLE (test infrastructure, NOT product code): the tail of the reference's VOS harness, restated with the same torch calls.

Only ``tests/`` may import this.  Follows /root/reference/sam_pt/vos_eval/eval.py:304-355 line by line (the code is inline in
`evaluate()`, which needs datasets / hydra / wandb to run, hence the restatement; every call below is the reference's own
torch call, so there is nothing to pin beyond the line-by-line reading).
"""
import torch
import torch.nn.functional as F


@torch.no_grad()
def vos_index_masks(pred_logits_list, gt_mask_list, gt_ti_list, target_hw, out_shape=None, need_resize=False, flip=False):
    logits = torch.stack([torch.zeros_like(pred_logits_list[0])] + list(pred_logits_list), dim=1)      # eval.py:304
    assert torch.all(logits[:, 0] == 0)
    for i, gt_ti in enumerate(gt_ti_list):                                                            # :321-322
        logits[:gt_ti, i + 1] = -1e8
    for i, (gt_ti, gt_mask) in enumerate(zip(gt_ti_list, gt_mask_list)):                              # :324-326
        g = F.interpolate(gt_mask[None, None, :, :], tuple(target_hw), mode="nearest")[0, 0]
        logits[gt_ti, i + 1] = torch.where(g.bool(), 1e8, -1e8)
    probs = F.softmax(logits, dim=1)                                                                  # :327
    out = []
    for ti in range(probs.shape[0]):                                                                  # :340-355
        prob = probs[ti]
        if need_resize:
            prob = F.interpolate(prob.unsqueeze(1), tuple(out_shape), mode="bilinear", align_corners=False)[:, 0]
        if flip:
            prob = torch.flip(prob, dims=[-1])
        out.append(torch.argmax(prob, dim=0).to(torch.uint8))
    return torch.stack(out)
