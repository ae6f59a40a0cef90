#!/usr/bin/env python
"""Generates tests/golden/smooth.npz by EXECUTING THE REFERENCE'S OWN src/loss.py in the build container
(needs /root/reference; nothing is copied, the fixture holds inputs and outputs only):

  compute_edge_aware_smoothness_loss(depth [B,H,W,1], rgb [B,H,W,3])          src/loss.py:64-84
  compute_bilateral_normal_smoothness_loss(normal [B,H,W,3], rgb [B,H,W,3])  src/loss.py:87-107

called the way main_train_dimo.py:363-372 does (NCHW renders permuted to NHWC), plus their autograd gradients.
    python tests/golden/make_golden_smooth.py
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    spec = importlib.util.spec_from_file_location("ref_loss", os.path.join(REF, "src/loss.py"))
    ls = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ls)
    g = torch.Generator().manual_seed(7)
    B, H, W = 2, 14, 11
    depth = (torch.rand(B, 1, H, W, generator=g) * 3).requires_grad_(True)
    normal = (torch.rand(B, 3, H, W, generator=g) - 0.5).requires_grad_(True)
    rgb = torch.rand(B, 3, H, W, generator=g)
    rgb[0, :, 3, 4] = rgb[0, :, 3, 5]           # an exact tie: |.| has zero gradient there
    rgb = rgb.requires_grad_(True)
    ld = ls.compute_edge_aware_smoothness_loss(depth.permute(0, 2, 3, 1), rgb.permute(0, 2, 3, 1))
    ln = ls.compute_bilateral_normal_smoothness_loss(normal.permute(0, 2, 3, 1), rgb.permute(0, 2, 3, 1))
    gd = torch.autograd.grad(ld, [depth, rgb], retain_graph=True)
    gn = torch.autograd.grad(ln, [normal, rgb])
    np.savez(os.path.join(HERE, "smooth.npz"), depth=depth.detach().numpy(), normal=normal.detach().numpy(),
             rgb=rgb.detach().numpy(), loss_depth=ld.item(), loss_normal=ln.item(),
             ddepth=gd[0].numpy(), drgb_depth=gd[1].numpy(), dnormal=gn[0].numpy(), drgb_normal=gn[1].numpy())
    print("wrote smooth.npz", ld.item(), ln.item())


if __name__ == "__main__":
    main()
