"""Result output, mirroring the `vtk_grid(...)` block of examples/poisson.jl:268-295 (WriteVTK.jl: an
unstructured grid of VTK_LINE / VTK_TRIANGLE / VTK_TETRA cells with vertex cochains as point data).

Host-side I/O: the mesh tables and cochains are downloaded through the C ABI and written
as a VTK XML UnstructuredGrid (.vtu, ASCII like the reference's `append=false, ascii=true` comment).
Top-rank cochains may be attached as cell data."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .core import Fun, Manifold

_CELLTYPE = {1: 3, 2: 5, 3: 10}          # VTK_LINE, VTK_TRIANGLE, VTK_TETRA


def _arr(a: np.ndarray, fmt: str) -> str:
    return " ".join(fmt % v for v in np.asarray(a).reshape(-1))


def write_vtu(path: str, mfd: Manifold, point_data: Optional[Dict[str, Fun]] = None,
              cell_data: Optional[Dict[str, Fun]] = None) -> str:
    """Write `mfd` with vertex cochains (`point_data`, R = 0) and top-rank cochains (`cell_data`, R = D) to
    `path` (".vtu" appended if missing).  Returns the file name."""
    D = mfd.D
    if D not in _CELLTYPE:
        raise ValueError("write_vtu: D must be 1, 2 or 3 (poisson.jl:268)")
    if not path.endswith(".vtu"):
        path += ".vtu"
    coords = mfd.get_coords()
    V, Cdim = coords.shape
    pts = np.zeros((V, 3))
    pts[:, :Cdim] = coords
    cells = mfd.get_simplices(D)
    n = cells.shape[0]
    point_data = point_data or {}
    cell_data = cell_data or {}
    for name, f in point_data.items():
        if (f.R, len(f)) != (0, V):
            raise ValueError(f"write_vtu: point data {name!r} must be a vertex cochain")
    for name, f in cell_data.items():
        if (f.R, len(f)) != (D, n):
            raise ValueError(f"write_vtu: cell data {name!r} must be a rank-{D} cochain")
    with open(path, "w", encoding="utf-8") as fh:
        fh.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="1.0" byte_order="LittleEndian">\n')
        fh.write(f'<UnstructuredGrid>\n<Piece NumberOfPoints="{V}" NumberOfCells="{n}">\n')
        fh.write('<Points>\n<DataArray type="Float64" NumberOfComponents="3" format="ascii">\n')
        fh.write(_arr(pts, "%.17g") + "\n</DataArray>\n</Points>\n<Cells>\n")
        fh.write('<DataArray type="Int64" Name="connectivity" format="ascii">\n' + _arr(cells, "%d") + "\n</DataArray>\n")
        fh.write('<DataArray type="Int64" Name="offsets" format="ascii">\n'
                 + _arr((D + 1) * np.arange(1, n + 1), "%d") + "\n</DataArray>\n")
        fh.write('<DataArray type="UInt8" Name="types" format="ascii">\n'
                 + _arr(np.full(n, _CELLTYPE[D]), "%d") + "\n</DataArray>\n</Cells>\n")
        for tag, data in (("PointData", point_data), ("CellData", cell_data)):
            fh.write(f"<{tag}>\n")
            for name, f in data.items():
                fh.write(f'<DataArray type="Float64" Name="{name}" format="ascii">\n' + _arr(f.values, "%.17g")
                         + "\n</DataArray>\n")
            fh.write(f"</{tag}>\n")
        fh.write("</Piece>\n</UnstructuredGrid>\n</VTKFile>\n")
    return path
