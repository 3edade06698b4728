"""Host-side graph substrate and message schedules.

The reference takes these from NamedGraphs (`named_grid`, `forest_cover`, `post_order_dfs_edges`;
src/edge_sequences.jl:32-51) — an un-vendored dependency — so they are restated here for the host
mirror.  Vertices are 0..nv-1; the engine receives explicit edge lists, so any consistent order works.
"""
import sys


class NamedGraph:
    def __init__(self, nv, edges, names=None):
        self.nv = int(nv)
        self.edges = [(int(u), int(v)) for u, v in edges]
        self.names = names  # optional vertex names (e.g. grid coordinates)
        self.inc = [[] for _ in range(self.nv)]
        self.eid = {}
        for e, (u, v) in enumerate(self.edges):
            self.inc[u].append(e)
            self.inc[v].append(e)
            self.eid[(u, v)] = e
            self.eid[(v, u)] = e

    @property
    def ne(self):
        return len(self.edges)

    def other(self, e, v):
        u, w = self.edges[e]
        return w if u == v else u

    def neighbors(self, v):
        return [self.other(e, v) for e in self.inc[v]]

    def degree(self, v):
        return len(self.inc[v])

    def has_edge(self, u, v):
        return (u, v) in self.eid

    def components(self):
        seen = [False] * self.nv
        comps = []
        for r in range(self.nv):
            if seen[r]:
                continue
            seen[r] = True
            stack, comp = [r], []
            while stack:
                x = stack.pop()
                comp.append(x)
                for y in self.neighbors(x):
                    if not seen[y]:
                        seen[y] = True
                        stack.append(y)
            comps.append(sorted(comp))
        return comps

    def is_tree(self):
        return self.ne == self.nv - 1 and len(self.components()) == 1


def named_grid(dims):
    dims = tuple(int(d) for d in dims)
    strides = [1] * len(dims)
    for i in range(len(dims) - 2, -1, -1):
        strides[i] = strides[i + 1] * dims[i + 1]
    nv = 1
    for d in dims:
        nv *= d
    names, edges = [], []
    for i in range(nv):
        c, r = [], i
        for s in strides:
            c.append(r // s)
            r %= s
        names.append(tuple(c))
        for ax in range(len(dims)):
            if c[ax] + 1 < dims[ax]:
                edges.append((i, i + strides[ax]))
    return NamedGraph(nv, edges, names)


def named_path_graph(n):
    return NamedGraph(n, [(i, i + 1) for i in range(n - 1)])


def named_comb_tree(dims):
    nx, ny = dims
    idx = lambda i, j: i * ny + j
    edges = [(idx(i, 0), idx(i + 1, 0)) for i in range(nx - 1)]
    for i in range(nx):
        edges += [(idx(i, j), idx(i, j + 1)) for j in range(ny - 1)]
    return NamedGraph(nx * ny, edges)


def heavy_hex_eagle():
    """IBM Eagle 127-qubit heavy-hex coupling graph (BASELINE.json config 3): 127 vertices, 144 edges.

    Seven rows of 14/15 qubits; consecutive rows are joined by four bridge qubits each."""
    starts = [0, 18, 37, 56, 75, 94, 113]
    lens = [14, 15, 15, 15, 15, 15, 14]
    edges = []
    for s, n in zip(starts, lens):
        edges += [(s + i, s + i + 1) for i in range(n - 1)]
    # bridge qubit: (upper-row qubit, lower-row qubit)
    bridges = {14: (0, 18), 15: (4, 22), 16: (8, 26), 17: (12, 30),
               33: (20, 39), 34: (24, 43), 35: (28, 47), 36: (32, 51),
               52: (37, 56), 53: (41, 60), 54: (45, 64), 55: (49, 68),
               71: (58, 77), 72: (62, 81), 73: (66, 85), 74: (70, 89),
               90: (75, 94), 91: (79, 98), 92: (83, 102), 93: (87, 106),
               109: (96, 114), 110: (100, 118), 111: (104, 122), 112: (108, 126)}
    for b, (u, w) in sorted(bridges.items()):
        edges += [(u, b), (b, w)]
    return NamedGraph(127, edges)


def forest_cover(g):
    remaining = set(range(g.ne))
    forests = []
    while remaining:
        seen = [False] * g.nv
        forest = []
        for r in range(g.nv):
            if seen[r]:
                continue
            seen[r] = True
            queue = [r]
            while queue:
                x = queue.pop(0)
                for e in g.inc[x]:
                    if e in remaining:
                        y = g.other(e, x)
                        if not seen[y]:
                            seen[y] = True
                            forest.append(e)
                            queue.append(y)
        remaining -= set(forest)
        forests.append(forest)
    return forests


def default_edge_sequence(g):
    """edge_sequence(::Algorithm"forest_cover") (src/edge_sequences.jl:32-47)."""
    sys.setrecursionlimit(max(10000, 4 * g.nv))
    seq = []
    for forest in forest_cover(g):
        adj = {}
        for e in forest:
            u, v = g.edges[e]
            adj.setdefault(u, []).append(v)
            adj.setdefault(v, []).append(u)
        seen = set()
        for root in sorted(adj):
            if root in seen:
                continue
            tree = []

            def rec(x, parent):
                seen.add(x)
                for y in adj[x]:
                    if y != parent:
                        rec(y, x)
                        tree.append((y, x))

            rec(root, -1)
            seq += tree + [(b, a) for (a, b) in reversed(tree)]
    return seq


def tree_gauge_sequence(g, region):
    """edge_sequence_between_regions(g, vertices(g), region) (src/abstractitensornetwork.jl:399-405): the (child, parent)
    edges, in post-order from region[0], of a spanning tree (the graph itself when it is a tree), without the edges
    inside `region` -- the walk of tree_gauge / tree_orthogonalize."""
    region = [int(region)] if isinstance(region, (int,)) or not hasattr(region, "__iter__") else [int(v) for v in region]
    if set(region) == set(range(g.nv)) or g.ne == 0:
        return []
    sys.setrecursionlimit(max(10000, 4 * g.nv))
    adj = {}
    for e in forest_cover(g)[0]:
        u, v = g.edges[e]
        adj.setdefault(u, []).append(v)
        adj.setdefault(v, []).append(u)
    out = []

    def rec(x, parent):
        for y in adj.get(x, []):
            if y != parent:
                rec(y, x)
                out.append((y, x))

    rec(region[0], -1)
    rs = set(region)
    return [(a, b) for (a, b) in out if not (a in rs and b in rs)]


def parallel_edge_sequence(g):
    """edge_sequence(::Algorithm"parallel") (src/edge_sequences.jl:49-51): one group per directed edge."""
    return [[e] for e in list(g.edges) + [(v, u) for (u, v) in g.edges]]


def edge_coloring(g):
    """Vertex-disjoint layers of edges, so that a whole gate layer runs as one batch."""
    used = [set() for _ in range(g.nv)]
    colors = []
    for e, (u, v) in enumerate(g.edges):
        c = 0
        while c in used[u] or c in used[v]:
            c += 1
        used[u].add(c)
        used[v].add(c)
        while len(colors) <= c:
            colors.append([])
        colors[c].append(e)
    return colors
