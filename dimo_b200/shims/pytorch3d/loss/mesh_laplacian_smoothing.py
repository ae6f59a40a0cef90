def cot_laplacian(*_a, **_k):
    """Imported by utils/deform_utils.py:3 and never called on the DIMO path."""
    raise NotImplementedError("dimo_b200 pytorch3d shim: cot_laplacian is not used by DIMO")
