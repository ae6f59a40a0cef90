"""Drop-in for `chamferdist.ChamferDistance` as the reference uses it (main_train_dimo.py:26,149,298-299):
    self.chamferDist = ChamferDistance();  self.chamferDist(cpts[None], cpts_ori[None])
-> the forward (source -> target) term, squared distances summed over the source points (dimo_chamfer_fwd/_bwd)."""
from dimo_b200.points import ChamferDistance

__all__ = ["ChamferDistance"]
