"""Minimal PLY codec for the two files the reference writes and reads through `plyfile`
(renderer/latent_gs_renderer.py:538-569 save_ply, :576-627 load_ply; released checkpoints, README.md:59-62):

    point_cloud.ply    element vertex N: x y z nx ny nz f_dc_* f_rest_* opacity scale_* rot_*   (all float)
    point_cloud_c.ply  element vertex M: c_x c_y c_z c_radius                                    (all float)

`plyfile` is a pip dependency that is not part of the hot path (and is absent from this image), so the format is
implemented directly: `write_vertex_ply` emits what `PlyData([PlyElement.describe(arr, 'vertex')]).write(path)`
emits for an all-'f4' structured array (binary_little_endian 1.0, one `property float <name>` line per field, rows
packed without padding); `read_ply` accepts binary little/big endian and ascii files with any scalar property types
(list properties are rejected: neither file has them) and returns one structured numpy array per element.
"""
import numpy as np

_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
          "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
          "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}
_NAMES = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint", "f4": "float",
          "f8": "double"}


def _emit(path, header, payload):
    blob = ("\n".join(header) + "\n").encode("ascii") + payload
    if hasattr(path, "write"):
        path.write(blob)
    else:
        with open(path, "wb") as f:
            f.write(blob)


def write_vertex_ply(path, names, columns, element="vertex"):
    """names: list of property names; columns: [N, len(names)] float array (written as float32)."""
    columns = np.ascontiguousarray(np.asarray(columns, dtype="<f4"))
    if columns.ndim != 2 or columns.shape[1] != len(names):
        raise ValueError(f"columns {columns.shape} do not match {len(names)} property names")
    header = ["ply", "format binary_little_endian 1.0", f"element {element} {columns.shape[0]}"]
    header += [f"property float {n}" for n in names]
    header.append("end_header")
    _emit(path, header, columns.tobytes())


def write_structured_ply(path, array, element="vertex"):
    """Structured numpy array (any scalar field types) -> one element, binary little endian."""
    header = ["ply", "format binary_little_endian 1.0", f"element {element} {array.shape[0]}"]
    fields = []
    for n in array.dtype.names:
        code = array.dtype[n].str.lstrip("<>|=")
        if code not in _NAMES:
            raise ValueError(f"unsupported field type {array.dtype[n]} for property {n}")
        header.append(f"property {_NAMES[code]} {n}")
        fields.append((n, "<" + code if code[1] != "1" else code))
    header.append("end_header")
    packed = np.empty(array.shape[0], dtype=fields)
    for n in array.dtype.names:
        packed[n] = array[n]
    _emit(path, header, packed.tobytes())


class PlyElements(dict):
    """{element name: structured array}; `.first` is the first element in file order (plyfile's elements[0])."""
    first = None
    comments = ()


def read_ply(path):
    """path: file name or a binary file object."""
    if hasattr(path, "read"):
        data = path.read()
        path = getattr(path, "name", "<stream>")
    else:
        with open(path, "rb") as f:
            data = f.read()
    if not data.startswith(b"ply"):
        raise ValueError(f"{path}: not a PLY file")
    end = data.find(b"end_header")
    if end < 0:
        raise ValueError(f"{path}: PLY header has no end_header")
    nl = data.find(b"\n", end)
    body = data[nl + 1:]
    lines = data[:end].decode("ascii", "replace").replace("\r", "").split("\n")
    fmt, elements, comments = None, [], []
    for ln in lines[1:]:
        tok = ln.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] in ("comment", "obj_info"):
            comments.append(ln)
        elif tok[0] == "element":
            elements.append((tok[1], int(tok[2]), []))
        elif tok[0] == "property":
            if not elements:
                raise ValueError(f"{path}: property before any element")
            if tok[1] == "list":
                raise ValueError(f"{path}: list property '{tok[-1]}' is not supported (not used by DIMO files)")
            if tok[1] not in _TYPES:
                raise ValueError(f"{path}: unknown property type '{tok[1]}'")
            elements[-1][2].append((tok[2], _TYPES[tok[1]]))
    if fmt not in ("binary_little_endian", "binary_big_endian", "ascii"):
        raise ValueError(f"{path}: unsupported PLY format '{fmt}'")
    out = PlyElements()
    out.comments = tuple(comments)
    if fmt == "ascii":
        toks = body.split()
        pos = 0
        for name, count, props in elements:
            arr = np.empty(count, dtype=[(n, t) for n, t in props])
            k = len(props)
            vals = toks[pos:pos + count * k]
            if len(vals) != count * k:
                raise ValueError(f"{path}: element '{name}' is truncated")
            pos += count * k
            for j, (n, t) in enumerate(props):
                col = np.array(vals[j::k], dtype="f8" if t[0] == "f" else "i8") if count else np.empty(0)
                arr[n] = col.astype(t)
            out[name] = arr
            if out.first is None:
                out.first = arr
        return out
    order = "<" if fmt == "binary_little_endian" else ">"
    pos = 0
    for name, count, props in elements:
        dt = np.dtype([(n, (order + t) if t[1] != "1" else t) for n, t in props])
        nbytes = dt.itemsize * count
        if pos + nbytes > len(body):
            raise ValueError(f"{path}: element '{name}' is truncated ({len(body) - pos} of {nbytes} bytes)")
        arr = np.frombuffer(body, dtype=dt, count=count, offset=pos)
        pos += nbytes
        native = np.empty(count, dtype=[(n, t) for n, t in props])
        for n, _t in props:
            native[n] = arr[n]
        out[name] = native
        if out.first is None:
            out.first = native
    return out


def property_names(arr):
    return list(arr.dtype.names)
