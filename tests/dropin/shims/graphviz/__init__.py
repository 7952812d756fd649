"""Test-only stub: torchopt/visual.py:25 imports graphviz at package import; nothing on the Adam path draws."""


class Digraph:  # noqa: D101
    def __init__(self, *args, **kwargs):
        raise RuntimeError("graphviz is not installed (test stub)")
