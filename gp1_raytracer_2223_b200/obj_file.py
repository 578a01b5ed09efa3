"""Wavefront OBJ -> TriangleMesh data, with the semantics of the reference's loader.

Restates ``Utils::ParseOBJ`` (reference source/Utils.h:377-451) as host plumbing (numpy, float32 arithmetic in the
reference's operation order); nothing here renders.  What the reference does, and this does too:

* the file is a whitespace-separated token stream; after each command the rest of the line is dropped;
* ``v x y z`` appends a position (decimal text -> nearest float);
* ``f a b c`` takes the text of each corner up to the first ``/`` (Maya style ``v/vt/vn``), reads it as a FLOAT
  (``std::stof``), truncates to int and subtracts one - only the first three corners of a face are used, negative
  (relative) indices are not resolved;
* ``#`` and every other command (``vn``, ``vt``, ``g``, ``o``, ``s``, ``usemtl`` ...) only drop their line;
* one face normal per triangle, ``Cross(v1 - v0, v2 - v0)`` then ``Normalize`` (source/Vector3.cpp:32-57: the cross
  product is spelled ``UnitX * s0 - UnitY * s1 + UnitZ * s2``, the division is by ``sqrtf(x*x + y*y + z*z)``) - a
  degenerate triangle yields NaNs exactly like the reference.

The result is the mesh BEFORE any ``UpdateTransforms``: what ``rt_upload_mesh_source`` takes.
"""
from __future__ import annotations

import dataclasses

import numpy as np

F = np.float32


@dataclasses.dataclass
class ObjMesh:
    positions: np.ndarray   # (V, 3) float32
    indices: np.ndarray     # (T, 3) int32
    normals: np.ndarray     # (T, 3) float32


def _stof_prefix(text: str) -> float:
    """std::stof: the longest leading decimal float; raises like the reference would throw."""
    import re
    m = re.match(r"\s*[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?|inf(inity)?|nan)", text, re.IGNORECASE)
    if not m:
        raise ValueError(f"face corner {text!r} does not start with a number")
    return float(m.group(0))


def face_normals(positions: np.ndarray, indices: np.ndarray) -> np.ndarray:
    """source/Utils.h:424-446 over all triangles at once, float32 operation by operation."""
    p = np.ascontiguousarray(positions, dtype=F)
    i = np.ascontiguousarray(indices, dtype=np.int64).reshape(-1, 3)
    a = p[i[:, 1]] - p[i[:, 0]]                       # edgeV0V1
    b = p[i[:, 2]] - p[i[:, 0]]                       # edgeV0V2
    s0 = a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1]
    s1 = a[:, 0] * b[:, 2] - a[:, 2] * b[:, 0]
    s2 = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
    one, zero = F(1), F(0)
    with np.errstate(invalid="ignore", divide="ignore"):
        # UnitX * s0 - UnitY * s1 + UnitZ * s2, component by component (keeps the reference's signed zeros)
        x = (one * s0 - zero * s1) + zero * s2
        y = (zero * s0 - one * s1) + zero * s2
        z = (zero * s0 - zero * s1) + one * s2
        m = np.sqrt((x * x + y * y) + z * z)
        n = np.stack([x / m, y / m, z / m], axis=1)
    return np.ascontiguousarray(n, dtype=F)


def parse_obj(path: str) -> ObjMesh:
    positions, indices = [], []
    with open(path, "r", errors="replace") as f:
        for line in f:
            tokens = line.split()
            if not tokens:
                continue
            command = tokens[0]
            if command == "v":
                if len(tokens) < 4:
                    raise ValueError(f"{path}: vertex line with fewer than three numbers: {line!r}")
                positions.append([F(tokens[1]), F(tokens[2]), F(tokens[3])])
            elif command == "f":
                if len(tokens) < 4:
                    continue                                # the reference skips a face it cannot read three corners of
                indices.append([int(_stof_prefix(t.split("/")[0])) - 1 for t in tokens[1:4]])
    pos = np.array(positions, dtype=F).reshape(-1, 3)
    idx = np.array(indices, dtype=np.int32).reshape(-1, 3)
    if idx.size and (idx.min() < 0 or idx.max() >= len(pos)):
        raise ValueError(f"{path}: face index outside the {len(pos)} vertices (the reference would read out of bounds)")
    return ObjMesh(pos, idx, face_normals(pos, idx))
