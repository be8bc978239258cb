import numpy as np

from . import GraphView, _PMap


def label_components(g):
    lab = g._labels()
    hist = np.bincount(lab) if len(lab) else np.zeros(0, dtype=np.int64)
    return _PMap(lab), hist


def extract_largest_component(g, directed=None, prune=False):
    if g.num_vertices() == 0:
        raise ValueError("attempt to get argmax of an empty sequence")  # graph-tool fails on the empty graph as well
    lab, hist = label_components(g)
    return GraphView(g, vfilt=lab.a == int(np.argmax(hist)))
