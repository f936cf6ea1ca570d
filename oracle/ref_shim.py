"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (bamse_b200/).

Loads the UNMODIFIED reference modules (clustering, tree_tools, tree_io) so that
tests / golden-vector generation / the `bench.py --impl reference` arm can call the
reference's own functions.  Two sources, tried in this order:

  1. /root/reference/*.py            (this container only; read-only mount)
  2. oracle/_ref/*.refbc             (sourceless byte-code compiled from (1) by
                                      oracle/build_ref.py; git-ignored build output
                                      that travels to the GPU box like a built .so)

The reference imports three modules that are absent / removed in this image
(SURVEY.md section 8c).  They are stubbed *before* import, with zero edits to the
reference:
  scipy.misc   (clustering.py:10 imports comb, logsumexp from it)  -> scipy.special
  cvxpy        (clustering.py:11; only used inside solve4 :410-426) -> empty module
  asciitree    (tree_io.py:3; only used by the text writer)         -> tiny stub
"""
import importlib
import importlib.machinery
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_PYC = os.path.join(_HERE, "_ref")

_loaded = None


def _install_stubs():
    import scipy.special as sp

    if "scipy.misc" not in sys.modules or not hasattr(sys.modules["scipy.misc"], "logsumexp"):
        m = types.ModuleType("scipy.misc")
        m.comb = sp.comb
        m.logsumexp = sp.logsumexp
        sys.modules["scipy.misc"] = m
        import scipy

        scipy.misc = m
    if "cvxpy" not in sys.modules:
        sys.modules["cvxpy"] = types.ModuleType("cvxpy")
    if "asciitree" not in sys.modules:
        a = types.ModuleType("asciitree")

        class LeftAligned(object):  # only needs to exist for `from asciitree import LeftAligned`
            def __call__(self, tree):
                raise NotImplementedError("asciitree stub")

        a.LeftAligned = LeftAligned
        sys.modules["asciitree"] = a


def available():
    """Where the reference can be loaded from: 'source', 'pyc' or None."""
    if os.path.isfile(os.path.join(REF_SRC, "clustering.py")):
        return "source"
    if os.path.isfile(os.path.join(REF_PYC, "clustering.refbc")):
        return "pyc"
    return None


def load():
    """Returns a namespace with .clustering, .tree_tools, .tree_io (reference modules)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    kind = available()
    if kind is None:
        raise ImportError("reference not available (neither /root/reference nor oracle/_ref)")
    _install_stubs()
    saved = {k: sys.modules.pop(k, None) for k in ("clustering", "tree_tools", "tree_io")}
    old_dont = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # never write __pycache__ into the read-only mount

    def _load(name):
        # the reference modules import each other by bare name; loading them in dependency
        # order with sys.modules populated satisfies those imports from either source
        if kind == "source":
            loader = importlib.machinery.SourceFileLoader(name, os.path.join(REF_SRC, name + ".py"))
        else:
            loader = importlib.machinery.SourcelessFileLoader(name, os.path.join(REF_PYC, name + ".refbc"))
        spec = importlib.util.spec_from_loader(name, loader)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        loader.exec_module(mod)
        return mod

    try:
        ns = types.SimpleNamespace(tree_tools=_load("tree_tools"), tree_io=_load("tree_io"),
                                   clustering=_load("clustering"), kind=kind)
    finally:
        sys.dont_write_bytecode = old_dont
        # keep the reference modules out of the global namespace of the caller
        for k in ("clustering", "tree_tools", "tree_io"):
            mod = sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]
            # the reference modules cross-import each other by bare name at import
            # time only, so removing them from sys.modules afterwards is safe.
            del mod
    _loaded = ns
    return ns
