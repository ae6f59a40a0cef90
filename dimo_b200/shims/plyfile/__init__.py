"""Drop-in for the slice of `plyfile` the reference uses (renderer/latent_gs_renderer.py:5, 538-627):
    el = PlyElement.describe(structured_array, 'vertex');  PlyData([el]).write(path)
    plydata = PlyData.read(path);  plydata.elements[0]["x"];  plydata.elements[0].properties[i].name
on top of dimo_b200.ply (binary little endian out; binary LE/BE and ascii in; scalar properties only)."""
import numpy as np

from dimo_b200 import ply as _ply

__all__ = ["PlyData", "PlyElement", "PlyProperty"]


class PlyProperty:
    def __init__(self, name, dtype):
        self.name = name
        self.val_dtype = dtype

    def __repr__(self):
        return f"PlyProperty({self.name!r}, {self.val_dtype!r})"


class PlyElement:
    def __init__(self, name, data):
        self.name = name
        self.data = data

    @staticmethod
    def describe(data, name, **_kw):
        if not isinstance(data, np.ndarray) or data.dtype.names is None:
            raise TypeError("only structured numpy arrays are supported")
        return PlyElement(name, data)

    @property
    def properties(self):
        return tuple(PlyProperty(n, self.data.dtype[n].str) for n in self.data.dtype.names)

    @property
    def count(self):
        return self.data.shape[0]

    def __getitem__(self, key):
        return self.data[key]

    def __len__(self):
        return self.data.shape[0]


class PlyData:
    def __init__(self, elements=(), text=False, byte_order="=", comments=(), obj_info=()):
        if text:
            raise NotImplementedError("dimo_b200 plyfile shim writes binary_little_endian only")
        self.elements = list(elements)
        self.comments = list(comments)

    def __getitem__(self, name):
        for e in self.elements:
            if e.name == name:
                return e
        raise KeyError(name)

    def write(self, stream):
        if len(self.elements) != 1:
            raise NotImplementedError("dimo_b200 plyfile shim writes one element per file (the DIMO files)")
        e = self.elements[0]
        _ply.write_structured_ply(stream, e.data, element=e.name)

    @staticmethod
    def read(stream):
        parsed = _ply.read_ply(stream)
        return PlyData([PlyElement(n, a) for n, a in parsed.items()], comments=parsed.comments)
