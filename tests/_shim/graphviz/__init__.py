"""Stub of the optional ``graphviz`` package so the reference (which imports it
unconditionally, nanograd/viz/comp_graph.py:1) can be imported in the build container."""


class Digraph:
    def __init__(self, *args, **kwargs):
        self.nodes, self.edges = [], []

    def node(self, *args, **kwargs):
        self.nodes.append((args, kwargs))

    def edge(self, *args, **kwargs):
        self.edges.append((args, kwargs))
