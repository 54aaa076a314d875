"""Computational-graph plots for ``Tensor.plot_forward`` / ``Tensor.plot_backward``.

Same entry points as the reference (nanograd/viz/comp_graph.py:4-107).  At the reference's HEAD
the plots are broken — ``Tensor.children`` is never filled (tensor.py:84) and the backward plot
reads a ``grad_fn`` attribute that no longer exists (comp_graph.py:101).  Here
``Function.apply`` records ``children`` and ``op`` on every result, and the backward label uses
the node's ctx class.  Device tensors are drawn from ``.shape`` only: nothing is copied off the
GPU.  ``graphviz`` is used when importable; otherwise a small stand-in collects the same
nodes/edges and renders DOT text (``.source``).
"""

try:  # pragma: no cover - depends on the environment
    from graphviz import Digraph as _Digraph
except Exception:  # graphviz is an optional dependency
    _Digraph = None


class DotGraph:
    """Minimal Digraph look-alike: node()/edge() bookkeeping and a DOT ``source`` string."""

    def __init__(self, format='png', graph_attr=None):
        self.format = format
        self.graph_attr = dict(graph_attr or {})
        self.nodes = {}
        self.edges = []

    def node(self, name, label=None, **attrs):
        self.nodes[name] = dict(label=label if label is not None else name, **attrs)

    def edge(self, tail, head, **attrs):
        self.edges.append((tail, head, attrs))

    @property
    def source(self):
        def esc(s):
            return str(s).replace('"', '\\"')
        lines = ["digraph {"]
        for k, v in self.graph_attr.items():
            lines.append(f'\t{k}="{esc(v)}"')
        for name, attrs in self.nodes.items():
            body = " ".join(f'{k}="{esc(v)}"' for k, v in attrs.items())
            lines.append(f'\t"{esc(name)}" [{body}]')
        for tail, head, _ in self.edges:
            lines.append(f'\t"{esc(tail)}" -> "{esc(head)}"')
        lines.append("}")
        return "\n".join(lines)


def _new_graph(rankdir):
    assert rankdir in ['LR', 'TB'], f"Unexpected rankdir argument (TB, LR available). Got {rankdir}."
    cls = _Digraph if _Digraph is not None else DotGraph
    return cls(format='png', graph_attr={'rankdir': rankdir})


def _display_name(n):
    if n.name != "no_name":
        return n.name
    return f"{n.op}_res" if n.op else n.name


class CompGraphVisualizer:
    """Walks ``children`` links from a root tensor and collects nodes and edges."""

    def __init__(self):
        self.nodes, self.edges = set(), set()

    def visualize(self, root, rankdir="LR"):
        self._build_trace(root)
        return self._build_graph(rankdir=rankdir)

    def _build_trace(self, root):
        stack = [root]
        while stack:
            node = stack.pop()
            if node in self.nodes:
                continue
            self.nodes.add(node)
            for child in node.children:
                self.edges.add((child, node))
                stack.append(child)

    def _build_graph(self, rankdir='LR'):
        raise NotImplementedError("build_graph() must be implemented for subclasses!")


class ForwardGraphVisualizer(CompGraphVisualizer):
    def _build_graph(self, rankdir: str = 'LR'):
        graph = _new_graph(rankdir)
        for n in self.nodes:
            uid = str(id(n))
            graph.node(name=uid, label=f"{_display_name(n)} | {n.shape}", shape='record')
            if n.op:
                graph.node(name=uid + n.op, label=n.op)
                graph.edge(uid + n.op, uid)
        for src, dst in self.edges:
            graph.edge(str(id(src)), str(id(dst)) + dst.op if dst.op else str(id(dst)))
        return graph


class BackwardGraphVisualizer(CompGraphVisualizer):
    def _build_graph(self, rankdir: str = 'LR'):
        graph = _new_graph(rankdir)
        for n in self.nodes:
            label = f"{_display_name(n)} | {n.shape}"
            if n.ctx is not None:
                label += f" | {type(n.ctx).__name__}Backward"
            graph.node(name=str(id(n)), label=label, shape="record")
        for src, dst in self.edges:
            graph.edge(str(id(dst)), str(id(src)))  # gradients flow from results to operands
        return graph
