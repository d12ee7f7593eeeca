"""Import shim for the UNMODIFIED reference package (test infrastructure only).

This module is part of the oracle tooling: it is used ONLY by
``tests/golden/gen_golden.py`` (run in the build container, where
``/root/reference`` is mounted) to execute the real jcoheur/pybitup code and
record golden vectors.  Nothing in ``pybitup_b200`` imports it, and nothing
that runs on the GPU box needs ``/root/reference``.

Why a shim: ``pybitup/__init__.py:1-10`` imports every sub-module, and several
of them need packages that are absent from this image (matplotlib, seaborn,
tikzplotlib, jsmin, mpi4py, sobol_seq) or a symbol that scipy removed
(``scipy.integrate.simps``, used at ``pybitup/distributions.py:4``).  We
pre-seed ``sys.modules`` with inert stand-ins so the reference imports and its
sampling path runs exactly as written.
"""
import re
import sys
import types

import os

REFERENCE_ROOT = "/root/reference"
# byte-for-byte copy of the package made by oracle/make_ref.py (git-ignored; travels to the GPU box, where
# /root/reference does not exist)
REFERENCE_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


class _Anything(types.ModuleType):
    """A module whose every attribute is another inert callable module."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Anything(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return _Anything(self.__name__ + "()")

    def __iter__(self):
        return iter(())


def _jsmin(text):
    # the reference only needs // and /* */ comments removed (solve_problem.py:35-37)
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = []
    for line in text.splitlines():
        # strip // comments that are not inside a string
        in_str = False
        i = 0
        while i < len(line):
            c = line[i]
            if c == '"' and (i == 0 or line[i - 1] != "\\"):
                in_str = not in_str
            if not in_str and line.startswith("//", i):
                line = line[:i]
                break
            i += 1
        out.append(line)
    return "\n".join(out)


class _Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def send(self, *a, **k):
        raise RuntimeError("single-process shim")

    def recv(self, *a, **k):
        raise RuntimeError("single-process shim")


def install(prefer_copy=False):
    """Make ``import pybitup`` resolve to the unmodified reference: ``/root/reference`` (the build container; what the
    golden-vector generator uses), or the copy under ``oracle/_ref`` when that is asked for or is all there is."""
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm",
                 "matplotlib.patches", "matplotlib.ticker", "matplotlib.lines",
                 "seaborn", "tikzplotlib", "sobol_seq"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    if "jsmin" not in sys.modules:
        m = types.ModuleType("jsmin")
        m.jsmin = _jsmin
        sys.modules["jsmin"] = m
    if "mpi4py" not in sys.modules:
        mpi4py = types.ModuleType("mpi4py")
        MPI = types.ModuleType("mpi4py.MPI")
        MPI.COMM_WORLD = _Comm()
        mpi4py.MPI = MPI
        sys.modules["mpi4py"] = mpi4py
        sys.modules["mpi4py.MPI"] = MPI
    import scipy.integrate
    if not hasattr(scipy.integrate, "simps"):
        scipy.integrate.simps = scipy.integrate.simpson
    root = REFERENCE_ROOT
    if prefer_copy or not os.path.isdir(os.path.join(REFERENCE_ROOT, "pybitup")):
        root = REFERENCE_COPY
        if not os.path.isdir(os.path.join(root, "pybitup")):
            raise ImportError("the reference is not available: neither %s nor %s (python oracle/make_ref.py builds the copy)"
                              % (REFERENCE_ROOT, REFERENCE_COPY))
    if root not in sys.path:
        sys.path.insert(0, root)
    import pybitup  # noqa: F401  (the reference)
    return pybitup
