"""Checkpoint loader with the reference's directory contract (sam_pt/point_tracker/utils/saverloader.py:30-73):
`<dir>/model-%09d.pth` holding {'model_state_dict': ...}; the highest step wins; loaded strict=False."""
import os

import torch


def load(ckpt_dir, model, device=None, optimizer=None, scheduler=None, model_ema=None, step=0, model_name="model",
         ignore_load=None):
    print("reading ckpt from %s" % ckpt_dir)
    assert os.path.exists(ckpt_dir)
    names = os.listdir(ckpt_dir)
    assert len(names) > 0
    steps = [int((n.split("-")[1]).split(".")[0]) for n in names]
    if step == 0:
        step = max(steps)
    path = os.path.join(ckpt_dir, "%s-%09d.pth" % (model_name, step))
    print("...found checkpoint %s" % path)
    sd = torch.load(path, map_location="cpu")["model_state_dict"]
    if ignore_load is not None:
        sd = {k: v for k, v in sd.items() if not any(ign in k for ign in ignore_load)}
    model.load_state_dict(sd, strict=False)
    return step
