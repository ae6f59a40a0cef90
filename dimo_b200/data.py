"""Ground truth of a training run resident in HBM (data half).

Reference: `GUI.__init__` decodes every (motion, view, frame) PNG into fp32 HOST tensors (main_train_dimo.py:102-126,
utils/load_utils.py:56-83: value = byte / 255, mask = alpha / 255); `train_step` then uploads each frame of the step and
resamples it to the current render resolution with F.interpolate(bilinear, align_corners=False) (:283-284, 305-313).
Here the frames stay on the device as the bytes they were decoded from (4 channels: R, G, B, mask) and a step's ground
truth is one `dimo_gt_fetch` launch over a slot list."""
import numpy as np
import torch

from . import _lib


class GroundTruthCache:
    """store [n_motions * n_views * n_frames, 4, size, size], uint8 (default; exact for 8-bit sources) or float32."""

    def __init__(self, n_motions, n_views, n_frames, size, device="cuda", dtype=torch.uint8):
        if dtype not in (torch.uint8, torch.float32):
            raise ValueError("GroundTruthCache stores uint8 or float32 samples")
        self.shape = (int(n_motions), int(n_views), int(n_frames))
        self.size = int(size)
        self.dtype = dtype
        self.store = torch.zeros(n_motions * n_views * n_frames, 4, size, size, dtype=dtype, device=device)

    def nbytes(self):
        return self.store.numel() * self.store.element_size()

    def slot(self, motion, view, frame):
        m, v, f = self.shape
        if not (0 <= motion < m and 0 <= view < v and 0 <= frame < f):
            raise IndexError(f"(motion, view, frame) = {(motion, view, frame)} outside {self.shape}")
        return (motion * v + view) * f + frame

    def put(self, motion, view, frame, image, mask):
        """image [1,3,H,W] or [3,H,W], mask [1,1,H,W] or [1,H,W]: uint8 samples, or floats in [0,1] as the reference's
        loader returns them (byte / 255).  A uint8 store takes floats only if they ARE byte / 255 (else it would
        silently quantise): checked."""
        img = image.reshape(3, *image.shape[-2:])
        msk = mask.reshape(1, *mask.shape[-2:])
        both = torch.cat((img, msk.to(img.dtype)), dim=0)
        if both.shape[-2:] != (self.size, self.size):
            raise ValueError(f"frame is {tuple(both.shape[-2:])}, cache holds {self.size}x{self.size}")
        if self.dtype == torch.uint8 and both.dtype != torch.uint8:
            scaled = both.float() * 255.0
            q = torch.round(scaled)
            # byte / 255 * 255 is within an ulp of the byte; anything further off is not an 8-bit sample.  (Not tested as
            # q / 255 == frame: torch's CUDA tensor / scalar multiplies by the reciprocal, which is not IEEE division;
            # the kernel's own (float)byte / 255.0f is, like NumPy's in utils/load_utils.py:70.)
            if float((scaled - q).abs().max()) > 1e-3:
                raise ValueError("float frame is not byte / 255: use a float32 cache for it")
            both = q.to(torch.uint8)
        self.store[self.slot(motion, view, frame)].copy_(both.to(self.dtype), non_blocking=True)

    def fetch(self, triples, resolution=None, out=None):
        """triples: list of (motion, view, frame) -> (rgb [S,3,r,r], mask [S,1,r,r]) fp32 at `resolution` (default: the
        stored size).  `out`: a previous result to overwrite (static buffers under a CUDA graph)."""
        r = int(resolution or self.size)
        S = len(triples)
        dev = self.store.device
        slots = torch.tensor([self.slot(*t) for t in triples], dtype=torch.int32).to(dev, non_blocking=True)
        if out is None:
            rgb = torch.empty(S, 3, r, r, dtype=torch.float32, device=dev)
            mask = torch.empty(S, 1, r, r, dtype=torch.float32, device=dev)
        else:
            rgb, mask = out
        _lib.call("dimo_gt_fetch", S, self.size, self.size, r, r, int(self.dtype == torch.uint8), _lib.ptr(self.store),
                  _lib.ptr(slots), _lib.ptr(rgb), _lib.ptr(mask), _lib.stream())
        return rgb, mask


def load_frame_arrays(bgr_or_bgra, alpha=None):
    """The arithmetic of utils/load_utils.py:64-76 on an already decoded uint8 array [H,W,3|4] (BGR(A), cv2 order):
    returns (rgb uint8 [3,H,W], mask uint8 [1,H,W] or None).  Decoding / background removal stay with the driver."""
    a = np.asarray(bgr_or_bgra)
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[-1] not in (3, 4):
        raise ValueError("expected a uint8 [H,W,3|4] array")
    rgb = torch.from_numpy(a[..., :3][..., ::-1].copy()).permute(2, 0, 1)
    if a.shape[-1] == 4:
        alpha = a[..., 3:4]
    mask = None if alpha is None else torch.from_numpy(np.ascontiguousarray(alpha)).reshape(a.shape[0], a.shape[1], 1) \
        .permute(2, 0, 1)
    return rgb, mask
