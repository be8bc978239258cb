"""Minimal stand-in for graph-tool 2.43 -- ONLY what workflow/scripts/process-by-contig_lowmem_AR.py uses
(Graph.add_edge_list(hashed=True), vertex property 'name', topology.extract_largest_component,
topology.label_components, GraphView(vfilt=...)).  Used by tests/golden/make_golden_py.py to run the
reference's own script in a container where graph-tool cannot be installed.  Documented graph-tool
behaviour that is reproduced: vertices are created in order of first appearance in the edge list (source
before target); label_components numbers components in order of their lowest vertex; the largest component
is the first one of maximal size (numpy argmax over the label histogram)."""
import numpy as np


class _PMap:
    def __init__(self, arr):
        self._a = np.asarray(arr)

    def get_array(self):
        return self._a

    @property
    def a(self):
        return self._a


class _VP(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class Graph:
    def __init__(self, directed=True):
        self.directed = directed
        self.vp = _VP()
        self._edges = []
        self._n = 0

    def add_edge_list(self, edge_list, hashed=False, hash_type="string", eprops=None):
        assert hashed, "stub supports hashed edge lists only"
        idx, names = {}, []
        for row in np.asarray(edge_list):
            pair = []
            for v in (row[0], row[1]):
                v = int(v)
                if v not in idx:
                    idx[v] = len(names)
                    names.append(v)
                pair.append(idx[v])
            self._edges.append(tuple(pair))
        self._n = len(names)
        return _PMap(np.asarray(names, dtype=np.int64))

    def num_vertices(self):
        return self._n

    def get_vertices(self):
        return np.arange(self._n)

    def _labels(self):
        par = list(range(self._n))

        def find(x):
            while par[x] != x:
                par[x] = par[par[x]]
                x = par[x]
            return x
        for a, b in self._edges:
            ra, rb = find(a), find(b)
            if ra != rb:
                par[max(ra, rb)] = min(ra, rb)
        lab, out = {}, np.zeros(self._n, dtype=np.int32)
        for v in range(self._n):  # components numbered by their lowest vertex
            r = find(v)
            if r not in lab:
                lab[r] = len(lab)
            out[v] = lab[r]
        return out


class GraphView:
    def __init__(self, g, vfilt=None):
        self._g = g
        self._mask = np.asarray(vfilt, dtype=bool) if vfilt is not None else np.ones(g.num_vertices(), bool)

    def get_vertices(self):
        return np.nonzero(self._mask)[0]

    @property
    def vp(self):
        return self._g.vp


from . import topology  # noqa: E402,F401
