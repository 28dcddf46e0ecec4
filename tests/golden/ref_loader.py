"""Execute the reference's OWN method bodies under Python 3 (fixture generation only).

The reference (``/root/reference``) is Python 2 and cannot be imported.  This loader reads
``phylo_hmrf.py`` / ``base.py`` / ``utility.py`` *where they lie*, cuts out the named
functions by indentation, turns py2 ``print`` statements into ``pass`` (they are pure
logging), and ``exec``s the result in memory.  No reference source is written to the
repo; the loader only works in the build container (``/root/reference`` does not exist
on the GPU box), which is why its outputs are committed as ``tests/golden/*.npz``.

Third-party callables that are absent here are injected as stubs by the caller
(``make_golden.py``): ``pygco.cut_general_graph`` (records its arguments) and
``log_multivariate_normal_density`` (restated sklearn-0.18 formula).
"""
from __future__ import annotations

import os
import re
import textwrap

REF = os.environ.get("PHMRF_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "phylo_hmrf.py"))


def _indent_of(line: str) -> int:
    return len(line) - len(line.lstrip("\t"))


def _extract(lines, name, level):
    """Return the source lines of ``def name`` at tab-indent ``level``."""
    pat = re.compile(r"^\t{%d}def %s\(" % (level, re.escape(name)))
    start = next(i for i, l in enumerate(lines) if pat.match(l))
    end = start + 1
    while end < len(lines):
        l = lines[end]
        if l.strip() and _indent_of(l) <= level and not l.lstrip().startswith("#"):
            break
        end += 1
    return lines[start:end]


_PRINT = re.compile(r"^(\s*)print(\s|$)(?!\()")


def _py3(src_lines, level):
    out = []
    for l in src_lines:
        m = _PRINT.match(l)
        if m and not l.lstrip().startswith("print("):
            out.append(m.group(1) + "pass\n")
        else:
            out.append(l)
    body = "".join(x[level:] if x.startswith("\t" * level) else x for x in out)
    return textwrap.dedent(body)


def load_functions(filename, names, level, namespace, patches=()):
    """Exec the named functions of ``filename`` (tab-indent ``level``) into ``namespace``.

    ``patches``: (old, new) text replacements applied to the extracted source in memory -- only
    for Python-2 semantics that Python 3 spells differently (classic integer division ``a/b`` on
    integers -> ``a//b``); every caller lists its patches next to the call."""
    with open(os.path.join(REF, filename), encoding="utf-8", errors="replace") as f:
        lines = f.readlines()
    for n in names:
        code = _py3(_extract(lines, n, level), level)
        for old, new in patches:
            code = code.replace(old, new)
        exec(compile(code, "%s:%s" % (filename, n), "exec"), namespace)
    return namespace


def build_reference_class(extra_globals):
    """A class carrying the reference's hot-path methods (phylo_hmrf.py) on top of the
    reference's base-class statistics helpers (base.py)."""
    import numpy as np
    import time

    g_base = {"np": np}
    load_functions("base.py", ["_initialize_sufficient_statistics", "_accumulate_sufficient_statistics_1"],
                   1, g_base)
    Base = type("_BaseGraph", (object,), {k: v for k, v in g_base.items() if callable(v) and k.startswith("_")
                                           and k != "__builtins__"})

    g = {"np": np, "time": time, "small_eps": 1e-16, "_BaseGraph": Base}
    g.update(extra_globals)
    names = ["_compute_log_likelihood", "_predict_posteriors", "_compute_posteriors_graph", "_compute_cost_v1",
             "_pairwise_compare", "_pairwise_compareLocal", "_pairwise_compare_ensemble",
             "_pairwise_compare_single", "predict", "_estimate_state_graphcuts_gco", "_pairwise_potential",
             "_edge_weight_undirected_vec", "_connected_edge", "_initialize_sufficient_statistics"]
    load_functions("phylo_hmrf.py", names, 1, g)
    cls = type("phyloHMRF", (Base,), {n: g[n] for n in names})
    g["phyloHMRF"] = cls  # for super(phyloHMRF, self)
    return cls


def load_utility(names, patches=()):
    import numpy as np
    g = {"np": np}
    load_functions("utility.py", names, 0, g, patches)
    return g
