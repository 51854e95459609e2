"""Host-side global planner substitute (stays on the host, per north_star).

The reference seeds the NMPC with an A* path on a visibility graph of inflated polygons,
built with two third-party packages that are absent from this image:
  pyclipper          — polygon offset with miter joins   (src/visibility/visibility.py:90-105)
  extremitypathfinder — visibility graph + shortest path  (src/visibility/visibility.py:47,64-67,81)
This module provides the two operations the reference needs from them, written from the
published algorithms (Clipper's offset join rules; visibility-graph shortest path), so
that realistic NMPC inputs can be generated offline.  It is input tooling for the solver
path, not part of the accelerated path.
"""
import heapq
import math

import numpy as np

EPS = 1e-9


def signed_area(poly):
    a = 0.0
    n = len(poly)
    for i in range(n):
        x0, y0 = poly[i]
        x1, y1 = poly[(i + 1) % n]
        a += x0 * y1 - x1 * y0
    return 0.5 * a


def offset_polygon(poly, delta, miter_limit=2.0):
    """Offset a simple polygon by `delta` (>0 grows, <0 shrinks) with mitered joins.

    Follows Clipper's ClipperOffset join rules for JT_MITER: a join is mitered while
    1 + cos(turn) >= 2 / miter_limit^2 and squared off otherwise; reflex joins keep the
    intersection of the two shifted edges.  Self-intersections of the result are not
    cleaned up (none of the reference's maps produce them at delta = +-vehicle_width).
    Returns a counter-clockwise list of (x, y), like Clipper's output orientation."""
    pts = [(float(x), float(y)) for x, y in poly]
    if signed_area(pts) < 0:
        pts.reverse()
    n = len(pts)
    normals = []
    for i in range(n):
        x0, y0 = pts[i]
        x1, y1 = pts[(i + 1) % n]
        dx, dy = x1 - x0, y1 - y0
        ln = math.hypot(dx, dy)
        normals.append((dy / ln, -dx / ln))          # outward normal of a CCW polygon
    miter_lim = 2.0 / (miter_limit * miter_limit)
    out = []
    for j in range(n):
        k = (j - 1) % n                               # previous edge
        nk, nj = normals[k], normals[j]
        x, y = pts[j]
        sin_a = nk[0] * nj[1] - nj[0] * nk[1]
        cos_a = nk[0] * nj[0] + nk[1] * nj[1]
        r = 1.0 + cos_a
        if sin_a * delta < 0 or r >= miter_lim:
            # reflex join (shifted edges intersect) or an allowed miter: same formula
            q = delta / r
            out.append((x + (nk[0] + nj[0]) * q, y + (nk[1] + nj[1]) * q))
        else:
            dxs = math.tan(math.atan2(sin_a, cos_a) / 4.0)
            out.append((x + delta * (nk[0] - nk[1] * dxs), y + delta * (nk[1] + nk[0] * dxs)))
            out.append((x + delta * (nj[0] + nj[1] * dxs), y + delta * (nj[1] - nj[0] * dxs)))
    return out


# ---------------------------------------------------------------------------------------
def _edges(polys):
    a, b = [], []
    for poly in polys:
        n = len(poly)
        for i in range(n):
            a.append(poly[i])
            b.append(poly[(i + 1) % n])
    return np.asarray(a, dtype=float).reshape(-1, 2), np.asarray(b, dtype=float).reshape(-1, 2)


def _cross(ax, ay, bx, by):
    return ax * by - ay * bx


class PolygonEnvironment:
    """Visibility-graph shortest path in a polygon with polygonal holes.  Same call
    surface as extremitypathfinder's PolygonEnvironment as used by the reference
    (store / prepare / find_shortest_path, src/visibility/visibility.py:64-67,81)."""

    def __init__(self):
        self.prepared = False

    def store(self, boundary_coordinates, list_of_hole_coordinates, validate=False):
        self.boundary = [(float(x), float(y)) for x, y in boundary_coordinates]
        self.holes = [[(float(x), float(y)) for x, y in h] for h in list_of_hole_coordinates]
        self.prepared = False

    # -- geometry -----------------------------------------------------------------------
    def _inside(self, poly_a, poly_b, pts):
        """even-odd rule + on-edge flag for points [n,2] against one polygon's edges."""
        px, py = pts[:, 0:1], pts[:, 1:2]
        ax, ay, bx, by = poly_a[:, 0], poly_a[:, 1], poly_b[:, 0], poly_b[:, 1]
        ex, ey = bx - ax, by - ay
        ln2 = ex * ex + ey * ey
        t = np.clip(((px - ax) * ex + (py - ay) * ey) / ln2, 0.0, 1.0)
        d2 = (ax + t * ex - px) ** 2 + (ay + t * ey - py) ** 2
        on_edge = (d2 < (1e-7) ** 2).any(axis=1)
        cond = (ay > py) != (by > py)
        with np.errstate(divide="ignore", invalid="ignore"):
            xint = ax + (py - ay) * ex / np.where(ey == 0, 1.0, ey)
        crossings = (cond & (px < xint)).sum(axis=1)
        return (crossings % 2 == 1), on_edge

    def _free(self, pts):
        """points in the closed free space: inside/on the boundary, not strictly inside a hole."""
        pts = np.asarray(pts, dtype=float).reshape(-1, 2)
        ins, on = self._inside(self._ba, self._bb, pts)
        ok = ins | on
        for ha, hb in self._hole_edges:
            ins, on = self._inside(ha, hb, pts)
            ok &= ~(ins & ~on)
        return ok

    def _visible(self, A, B):
        """segments A[i]->B[i] stay in the closed free space?  A, B: [n,2]."""
        A = np.asarray(A, dtype=float).reshape(-1, 2)
        B = np.asarray(B, dtype=float).reshape(-1, 2)
        n = A.shape[0]
        ea, eb = self._ea, self._eb
        dx, dy = (B - A)[:, 0:1], (B - A)[:, 1:2]
        ax, ay = A[:, 0:1], A[:, 1:2]
        # proper crossings with any polygon edge
        o1 = _cross(dx, dy, ea[:, 0] - ax, ea[:, 1] - ay)
        o2 = _cross(dx, dy, eb[:, 0] - ax, eb[:, 1] - ay)
        ex, ey = (eb - ea)[:, 0], (eb - ea)[:, 1]
        o3 = _cross(ex, ey, ax - ea[:, 0], ay - ea[:, 1])
        o4 = _cross(ex, ey, B[:, 0:1] - ea[:, 0], B[:, 1:2] - ea[:, 1])
        scale = np.maximum(np.abs(dx) + np.abs(dy), 1e-12)
        tol = 1e-9 * scale * np.maximum(np.abs(ex) + np.abs(ey), 1.0)
        proper = (o1 * o2 < -tol * tol) & (o3 * o4 < -tol * tol) & (np.abs(o1) > tol) & (np.abs(o2) > tol) \
            & (np.abs(o3) > tol) & (np.abs(o4) > tol)
        ok = ~proper.any(axis=1)
        # touching: split at polygon vertices lying on the segment and test the pieces' midpoints
        V = self._verts
        ln2 = np.maximum(dx * dx + dy * dy, 1e-300)
        tv = ((V[:, 0] - ax) * dx + (V[:, 1] - ay) * dy) / ln2
        dist2 = (ax + tv * dx - V[:, 0]) ** 2 + (ay + tv * dy - V[:, 1]) ** 2
        on_seg = (dist2 < (1e-7) ** 2) & (tv > 1e-9) & (tv < 1 - 1e-9)
        ts = np.where(on_seg, tv, np.nan)
        ts = np.concatenate([np.zeros((n, 1)), np.sort(ts, axis=1), np.ones((n, 1))], axis=1)
        # after the sort NaNs sit at the end (before the appended 1): forward-fill them with 1
        ts = np.where(np.isnan(ts), 1.0, ts)
        mids = 0.5 * (ts[:, :-1] + ts[:, 1:])
        mx = ax + mids * dx
        my = ay + mids * dy
        free = self._free(np.stack([mx.ravel(), my.ravel()], axis=1)).reshape(mids.shape)
        degenerate = (ts[:, 1:] - ts[:, :-1]) < 1e-12
        ok &= (free | degenerate).all(axis=1)
        return ok

    # -- graph --------------------------------------------------------------------------
    def prepare(self):
        self._ba, self._bb = _edges([self.boundary])
        self._hole_edges = [_edges([h]) for h in self.holes]
        self._ea, self._eb = _edges([self.boundary] + self.holes)
        self._verts = np.asarray([p for poly in [self.boundary] + self.holes for p in poly], dtype=float).reshape(-1, 2)
        # extremities: boundary vertices with interior angle > 180 deg, hole vertices with < 180 deg
        nodes = []
        polys = [(self.boundary, +1)] + [(h, -1) for h in self.holes]
        for poly, want in polys:
            sgn = 1.0 if signed_area(poly) > 0 else -1.0
            n = len(poly)
            for i in range(n):
                x0, y0 = poly[i - 1]
                x1, y1 = poly[i]
                x2, y2 = poly[(i + 1) % n]
                turn = _cross(x1 - x0, y1 - y0, x2 - x1, y2 - y1) * sgn   # > 0: convex corner of the polygon
                if (want == +1 and turn < -EPS) or (want == -1 and turn > EPS):
                    nodes.append((x1, y1))
        nodes = np.asarray(nodes, dtype=float).reshape(-1, 2)
        if len(nodes):
            nodes = nodes[self._free(nodes)]
        self.nodes = nodes
        m = len(nodes)
        self.adj = np.full((m, m), np.inf)
        if m > 1:
            ii, jj = np.triu_indices(m, 1)
            vis = self._visible(nodes[ii], nodes[jj])
            d = np.hypot(*(nodes[ii] - nodes[jj]).T)
            self.adj[ii[vis], jj[vis]] = d[vis]
            self.adj[jj[vis], ii[vis]] = d[vis]
        self.prepared = True

    def find_shortest_path(self, start_coordinates, goal_coordinates, free_space_after=True, verify=True):
        """-> (path as a list of (x, y) tuples incl. start and goal, length); ([], None) if unreachable."""
        if not self.prepared:
            self.prepare()
        s = np.asarray(start_coordinates, dtype=float)
        g = np.asarray(goal_coordinates, dtype=float)
        m = len(self.nodes)
        if self._visible(s[None], g[None])[0]:
            return [tuple(s), tuple(g)], float(np.hypot(*(g - s)))
        vs = self._visible(np.repeat(s[None], m, 0), self.nodes) if m else np.zeros(0, bool)
        vg = self._visible(np.repeat(g[None], m, 0), self.nodes) if m else np.zeros(0, bool)
        ds = np.hypot(*(self.nodes - s).T) if m else np.zeros(0)
        dg = np.hypot(*(self.nodes - g).T) if m else np.zeros(0)
        # Dijkstra with an admissible straight-line heuristic (A*), nodes 0..m-1, goal = m
        dist = np.full(m + 1, np.inf)
        prev = np.full(m + 1, -2, dtype=int)
        heap = []
        for i in np.nonzero(vs)[0]:
            dist[i] = ds[i]
            prev[i] = -1
            heapq.heappush(heap, (ds[i] + dg[i], ds[i], int(i)))
        done = np.zeros(m + 1, dtype=bool)
        while heap:
            _, d, i = heapq.heappop(heap)
            if done[i]:
                continue
            done[i] = True
            if i == m:
                break
            if vg[i] and d + dg[i] < dist[m]:
                dist[m] = d + dg[i]
                prev[m] = i
                heapq.heappush(heap, (dist[m], dist[m], m))
            row = self.adj[i]
            for j in np.nonzero(np.isfinite(row))[0]:
                nd = d + row[j]
                if nd < dist[j] - 1e-12:
                    dist[j] = nd
                    prev[j] = i
                    heapq.heappush(heap, (nd + dg[j], nd, int(j)))
        if not np.isfinite(dist[m]):
            return [], None
        path = [tuple(g)]
        i = prev[m]
        while i >= 0:
            path.append((float(self.nodes[i, 0]), float(self.nodes[i, 1])))
            i = prev[i]
        path.append(tuple(s))
        path.reverse()
        return [(float(x), float(y)) for x, y in path], float(dist[m])
