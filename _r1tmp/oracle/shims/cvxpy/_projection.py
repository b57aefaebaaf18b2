"""Exact Euclidean projection of a 2-D point onto {x : G x <= h}.

TEST INFRASTRUCTURE ONLY.  This is the definition of "exact projection" that the
C restatement (oracle/anm_oracle.c: project_polygon) and the CUDA kernel follow:

  1. if the point satisfies every row, it is its own projection (returned bit-for-bit);
  2. otherwise the projection lies on the boundary: it is the closest *feasible* point
     among (a) the orthogonal projections onto each row's boundary line and
     (b) the pairwise intersections of the boundary lines.
Rows with a non-finite right-hand side are ignored (never active).  Axis-aligned
rows are handled without rounding, so that box clipping is reproduced exactly (the
reference's tests use assertEqual there, tests/simulator/test_devices.py:541-549).
"""
import numpy as np

FEAS_TOL = 1e-12  # slack allowed on the rows that do not define a candidate


def _line_projection(a, b, h, p, q):
    if b == 0.0:
        return h / a, q
    if a == 0.0:
        return p, h / b
    t = (a * p + b * q - h) / (a * a + b * b)
    return p - t * a, q - t * b


def project_onto_polygon(G, h, point):
    G = np.asarray(G, dtype=np.float64)
    h = np.asarray(h, dtype=np.float64)
    p, q = float(point[0]), float(point[1])
    rows = [(float(G[i, 0]), float(G[i, 1]), float(h[i])) for i in range(len(h)) if np.isfinite(h[i])]

    def feasible(x, y, skip=()):
        for k, (a, b, c) in enumerate(rows):
            if k in skip:
                continue
            if a * x + b * y - c > FEAS_TOL:
                return False
        return True

    if feasible(p, q):
        return np.array([p, q])

    best, best_d = None, np.inf
    n = len(rows)
    for i in range(n):
        a, b, c = rows[i]
        x, y = _line_projection(a, b, c, p, q)
        if feasible(x, y, skip=(i,)):
            d = (x - p) ** 2 + (y - q) ** 2
            if d < best_d:
                best, best_d = (x, y), d
    for i in range(n):
        a1, b1, c1 = rows[i]
        for j in range(i + 1, n):
            a2, b2, c2 = rows[j]
            det = a1 * b2 - a2 * b1
            if det == 0.0:
                continue
            x = (c1 * b2 - c2 * b1) / det
            y = (a1 * c2 - a2 * c1) / det
            if feasible(x, y, skip=(i, j)):
                d = (x - p) ** 2 + (y - q) ** 2
                if d < best_d:
                    best, best_d = (x, y), d
    if best is None:
        raise ValueError("empty feasible polygon")
    return np.array(best)
