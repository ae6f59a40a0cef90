"""Host-side packing of a step (Renderer.prepare_step): frame list -> camera block + unique (motion, t) pairs.  The
deformation depends on (motion, t) only, so frames that differ in the view share one TimeNet / LBS
evaluation; this is pure host logic and runs without a GPU."""
import math

import torch

from dimo_b200 import raster
from dimo_b200.camera import orbit_minicam
from dimo_b200.renderer import Renderer


def test_pairs_are_deduplicated_and_cameras_packed():
    r = Renderer(sh_degree=0, device="cpu", num_latent_code=4)
    r.gaussians._xyz = torch.zeros(5, 3)
    views = [0, 3, 0, 3, 5, 5]
    cams = [orbit_minicam(v, 8, 64, 48, device="cpu") for v in views]
    times = [0.25, 0.25, 0.5, 0.5, 0.25, 0.75]
    motions = [2, 2, 2, 2, 1, 2]
    prep = r.prepare_step(cams, times, motions)
    assert prep["S"] == 6 and prep["W"] == 64 and prep["H"] == 48
    # unique (motion, t) pairs in first-seen order; frames 0/1 and 2/3 share theirs
    assert prep["U"] == 4 and prep["pair_of_frame"] == [0, 0, 1, 1, 2, 3]
    assert prep["pf"].tolist() == [0, 0, 1, 1, 2, 3] and prep["pf32"].dtype == torch.int32
    assert prep["li"].tolist() == [2, 2, 1, 2] and torch.allclose(prep["t"], torch.tensor([0.25, 0.5, 0.25, 0.75]))
    assert prep["expand"] is True
    for i, c in enumerate(cams):
        want = raster.pack_cameras(c.world_view_transform, c.full_proj_transform, c.camera_center,
                                   math.tan(c.FoVx * 0.5), math.tan(c.FoVy * 0.5), r.bg_color)
        assert torch.allclose(prep["cams"][i], want[0], atol=0, rtol=0)
    assert prep["cams"].shape == (6, raster.CAM_FLOATS)
    # static-buffer refresh (CUDA-graph replay): same shapes, new content, same tensor objects
    keep = {k: prep[k] for k in ("cams", "t", "li", "pf", "pf32")}
    cams2 = [orbit_minicam(v, 8, 64, 48, device="cpu") for v in [1, 2, 1, 2, 7, 7]]
    out = r.prepare_step(cams2, [0.1, 0.1, 0.9, 0.9, 0.1, 0.3], [0, 0, 0, 0, 3, 0], out=prep)
    assert all(out[k] is keep[k] for k in keep)
    assert out["li"].tolist() == [0, 0, 3, 0] and torch.allclose(out["t"], torch.tensor([0.1, 0.9, 0.1, 0.3]))
    assert torch.equal(out["cams"][4, :16], cams2[4].world_view_transform.reshape(-1).float())


def test_one_pair_per_frame_needs_no_expansion():
    r = Renderer(sh_degree=0, device="cpu", num_latent_code=2)
    r.gaussians._xyz = torch.zeros(3, 3)
    cams = [orbit_minicam(v, 4, 32, 32, device="cpu") for v in range(3)]
    prep = r.prepare_step(cams, [0.0, 0.5, 0.75], [0, 0, 1])
    assert prep["U"] == 3 and prep["expand"] is False


def test_clone_prep_owns_its_buffer():
    """the static inputs of a captured graph: a deep copy whose tensors are views of ONE own buffer, refreshed in place
    by prepare_step(out=...)"""
    import torch
    from dimo_b200.camera import orbit_minicam
    from dimo_b200.renderer import Renderer
    r = Renderer(sh_degree=0, num_latent_code=4, device="cpu")
    r.gaussians._xyz = torch.zeros(4, 3)
    cams = [orbit_minicam(v, 3, 32, 32, device="cpu") for v in range(3)]
    prep = r.prepare_step(cams, [0.0, 0.5, 0.5], [0, 1, 1])
    st = Renderer.clone_prep(prep)
    assert st["_buf"].data_ptr() != prep["_buf"].data_ptr()
    for k in ("cams", "t", "li", "pf", "pf32"):
        assert torch.equal(st[k], prep[k])
        lo, hi = st["_buf"].data_ptr(), st["_buf"].data_ptr() + st["_buf"].numel()
        assert lo <= st[k].data_ptr() < hi                       # a view of the clone's own buffer
    keep = st["cams"]
    r.prepare_step(cams[::-1], [0.25, 0.25, 0.75], [2, 2, 3], out=st)
    assert keep.data_ptr() == st["cams"].data_ptr()
    assert st["li"].tolist() == [2, 3] and st["pf"].tolist() == [0, 0, 1]
    assert torch.allclose(st["t"], torch.tensor([0.25, 0.75]))
    assert torch.equal(prep["li"], torch.tensor([0, 1]))         # the original is untouched
