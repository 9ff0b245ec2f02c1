"""Element-log export: the generated quads of an episode in the
reference's two on-disk formats, built from the device element log (mg_get_elements).

* ``write_2_file``  -- JSON ``{"nodes": {i: {"coordinates": [x, y], "connected": [...]}},
  "elements": {k: [i0, i1, i2, i3]}}`` as ``BoudaryEnv.write_2_file`` (rl/boundary_env.py:648-669):
  node index = position in ``boundary.vertices`` = original vertices followed by the inserted ones
  in insertion order, which is exactly this package's vertex-id convention; ``connected`` lists the
  partners in segment order (the two polygon neighbours first, then element edges as they were
  connected, general/components.py:832-837).
* ``write_inp``     -- Abaqus ``.inp`` as ``write_generated_elements_2_file`` (general/mesh.py:1842-1864).
"""
from __future__ import annotations

import json
from typing import Dict, List

import numpy as np


def connectivity(n0: int, quads: np.ndarray, n_vertices: int) -> List[List[int]]:
    """Per-vertex partner lists in the order the reference's Vertex.segments would hold them:
    deep_copy links i-1 <-> i for the polygon (components.py:213-220), then every element calls
    connect_vertices, which links v[i] <-> v[i-1] unless a segment already exists."""
    adj: List[List[int]] = [[] for _ in range(n_vertices)]
    for i in range(n0):
        a = (i - 1) % n0
        adj[a].append(i)
        adj[i].append(a)
    for q in quads:
        for i in range(4):
            a, b = int(q[i]), int(q[i - 1])
            if b not in adj[a]:
                adj[a].append(b)
                adj[b].append(a)
    return adj


def mesh_dict(n0: int, quads: np.ndarray, vertex_xy: np.ndarray) -> Dict:
    adj = connectivity(n0, quads, len(vertex_xy))
    nodes = {}
    for i, (x, y) in enumerate(vertex_xy):
        conn: List[int] = []
        for p in adj[i]:
            if p not in conn:
                conn.append(p)
        nodes[i] = {"coordinates": [float(x), float(y)], "connected": conn}
    elements = {k: [int(v) for v in q] for k, q in enumerate(quads)}
    return {"nodes": nodes, "elements": elements}


def write_2_file(filename, n0: int, quads: np.ndarray, vertex_xy: np.ndarray) -> None:
    with open(filename, "w") as fw:
        json.dump(mesh_dict(n0, quads, vertex_xy), fw)


def write_inp(filename, n0: int, quads: np.ndarray, vertex_xy: np.ndarray) -> None:
    if len(quads) == 0:
        print("There are no elements generated!")
        return
    order: List[int] = []
    for q in quads:
        for v in q:
            if int(v) not in order:
                order.append(int(v))
    for v in range(n0):            # the reference indexes every original vertex (mesh.py:1856-1857)
        if v not in order:
            order.append(v)
    pos = {v: k + 1 for k, v in enumerate(order)}
    with open(filename, "w") as fw:
        fw.write("*NODE, NSET=ALLNODES\n")
        for v in order:
            fw.write(f"{pos[v]}, {vertex_xy[v][0]}, {vertex_xy[v][1]}\n")
        i = 0
        for i in range(1, n0):
            fw.write(f"*ELEMENT, TYPE=B21, ELSET=EB{i}\n {i+1}, {pos[i-1]}, {pos[i]}\n")
        fw.write(f"*ELEMENT, TYPE=S4R, ELSET=EB{i+1} \n")
        for k, q in enumerate(quads):
            fw.write(f"{k+1}, {pos[int(q[0])]}, {pos[int(q[1])]}, {pos[int(q[2])]}, {pos[int(q[3])]}\n")


def save_meshes_figure(name, meshes, boundary_xy, indexing: bool = False, dpi: int = 300, style: str = "k.-") -> str:
    """``MeshGeneration.save_meshes`` (general/mesh.py:1785-1792 over generate_meshes_canvas :1761-1783): the original
    boundary and every generated element as a closed polyline, element index at the centroid when ``indexing``.
    With matplotlib the figure is written to ``name`` exactly like the reference; without it an SVG with the same
    content is written to ``name`` with the suffix replaced by ``.svg``.  Returns the path written.

    ``meshes``: iterable of (4, 2) arrays or objects with ``.vertices[i].x/.y``."""
    import os

    def xy_of(m):
        if hasattr(m, "vertices"):
            return np.array([[float(v.x), float(v.y)] for v in m.vertices], np.float64)
        return np.asarray(m, np.float64).reshape(-1, 2)

    polys = [xy_of(m) for m in meshes]
    bxy = np.asarray(boundary_xy, np.float64).reshape(-1, 2)
    try:
        import matplotlib
        if not isinstance(getattr(matplotlib, "__version__", None), str):     # an inert stand-in module, not matplotlib
            raise ImportError("matplotlib is not installed")
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:
        plt = None
    if plt is not None:
        plt.clf()
        closed = np.vstack([bxy, bxy[:1]])
        plt.plot(closed[:, 0], closed[:, 1], style, linewidth=1)
        for k, p in enumerate(polys):
            c = np.vstack([p, p[:1]])
            plt.plot(c[:, 0], c[:, 1], style, linewidth=1)
            if indexing:
                plt.text(p[:, 0].mean(), p[:, 1].mean(), k, fontsize=4)
        plt.gca().set_aspect("equal", adjustable="box")
        plt.subplots_adjust(top=1, bottom=0, right=1, left=-0, hspace=0, wspace=0)
        plt.savefig(name, dpi=dpi)
        plt.close("all")
        return str(name)
    pts = np.vstack([bxy] + polys) if polys else bxy
    lo, hi = pts.min(axis=0), pts.max(axis=0)
    span = max(hi[0] - lo[0], hi[1] - lo[1], 1e-9)
    scale, pad = 1000.0 / span, 10.0

    def tr(p):
        return (pad + (p[0] - lo[0]) * scale, pad + (hi[1] - p[1]) * scale)          # y up

    w, h = 2 * pad + (hi[0] - lo[0]) * scale, 2 * pad + (hi[1] - lo[1]) * scale
    out = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{w:.0f}" height="{h:.0f}" viewBox="0 0 {w:.1f} {h:.1f}">',
           '<rect width="100%" height="100%" fill="white"/>']

    def poly(p, width):
        s = " ".join(f"{x:.2f},{y:.2f}" for x, y in (tr(q) for q in p))
        out.append(f'<polygon points="{s}" fill="none" stroke="black" stroke-width="{width}"/>')

    poly(bxy, 1.5)
    for k, p in enumerate(polys):
        poly(p, 0.8)
        if indexing:
            cx, cy = tr(p.mean(axis=0))
            out.append(f'<text x="{cx:.1f}" y="{cy:.1f}" font-size="8" text-anchor="middle">{k}</text>')
    out.append("</svg>")
    path = os.path.splitext(str(name))[0] + ".svg"
    with open(path, "w") as f:
        f.write("\n".join(out))
    return path
