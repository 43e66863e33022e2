"""GridRoad: m x n grid of intersections joined by one-way roads without turns.

Host-side table builder with the attribute names of reference gym_traffic/envs/roadgraph.py:25-64
(`len, m, n, train_roads, roads, intersections, phases, dest, nexts, entrypoints`,
`generate_entrypoints`, `get_next`; `locs` - the renderer's segment coordinates, roadgraph.py:5-22 - is out of scope).  Road id = d * V + row * n + col with V = m * n and d in
{0: east-bound, 1: west-bound, 2: south-bound (row + 1), 3: north-bound (row - 1)}; ids >= 4V are
the 2n + 2m exit roads.  The device builds the same tables in te_create (te_api.cu) - tests compare.
"""
import numpy as np


class GridRoad(object):
    def __init__(self, m, n, l):
        self.m, self.n = int(m), int(n)
        self.len = np.float32(l)
        v = self.m * self.n
        self.intersections = v
        self.train_roads = 4 * v
        self.roads = 4 * v + 2 * self.n + 2 * self.m
        ids = np.arange(self.roads)
        self.phases = (ids // v < 2).astype(np.int32)              # E/W-bound roads share phase 1
        self.dest = np.where(ids < 4 * v, ids % v, -1).astype(np.int32)
        self.nexts = np.array([self.get_next(i) for i in range(self.roads)], dtype=np.int32)
        self.generate_entrypoints(0)

    def get_next(self, i):
        """Road a car continues on after road i (straight ahead), an exit road at the boundary, or -1."""
        v, n, m = self.intersections, self.n, self.m
        if i >= 4 * v:
            return -1
        d, cell = divmod(i, v)
        row, col = divmod(cell, n)
        if d == 0:
            return i + 1 if col + 1 < n else 4 * v + n + row
        if d == 1:
            return i - 1 if col > 0 else 4 * v + 2 * n + m + row
        if d == 2:
            return i + n if row + 1 < m else 4 * v + n + m + col
        return i - n if row > 0 else 4 * v + col

    def generate_entrypoints(self, choices):
        """Entry roads of the open sides; bit k of `choices` set closes side k (west, east, north, south)."""
        n, m, v = self.n, self.m, self.intersections
        sides = [n * np.arange(m), v + n * np.arange(1, m + 1) - 1, 2 * v + np.arange(n),
                 3 * v + n * (m - 1) + np.arange(n)]
        keep = [s for k, s in enumerate(sides) if not (int(choices) >> k) & 1]
        self.entrypoints = (np.concatenate(keep) if keep else np.empty(0)).astype(np.int32)
