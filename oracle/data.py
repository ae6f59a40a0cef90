"""Oracle: the ground-truth path of a step.  TEST INFRASTRUCTURE (see oracle/__init__.py).

PINNED by construction: it CALLS the function the reference calls -- torch.nn.functional.interpolate(mode="bilinear",
align_corners=False) on the CPU (main_train_dimo.py:307, 312; utils/load_utils.py:79-80) -- on byte / 255 floats
(utils/load_utils.py:25, 70)."""
import torch
import torch.nn.functional as F


def fetch(store_u8_or_f32, slots, resolution):
    """store [F,4,H,W] (uint8 or float32, CPU), slots: list[int] -> (rgb [S,3,r,r], mask [S,1,r,r])."""
    x = store_u8_or_f32[torch.tensor(slots, dtype=torch.long)]
    x = x.float() / 255.0 if x.dtype == torch.uint8 else x.float()
    rgb = F.interpolate(x[:, :3], (resolution, resolution), mode="bilinear", align_corners=False)
    mask = F.interpolate(x[:, 3:], (resolution, resolution), mode="bilinear", align_corners=False)
    return rgb, mask
