#!/usr/bin/env python
"""Materialise a runnable copy of the reference under oracle/_ref/ (git-ignored).

TEST INFRASTRUCTURE ONLY.

The reference (/root/reference, Python 2.7) cannot be imported by Python 3.12 /
scipy 1.18 as it stands: print statements, ``xrange``, ``except A, B:`` and the
removed ``scipy.integrate.romberg``.  This script reads the reference sources
*where they lie* and writes mechanically transformed copies to ``oracle/_ref``.
No reference source is committed to this repository; ``oracle/_ref`` is listed
in .gitignore (but not in .gpurunignore, so the generated copy travels to the
GPU box where it serves as the ``cpu_baseline.kind == "reference"`` arm).

Transformations (line numbers are preserved, so reference file:line citations
also hold for the generated copy):

1. ``print X``            -> ``print(X)``           (py2 statement)
2. ``xrange(``            -> ``range(``
3. ``except A, B:``       -> ``except (A, B):``      (covariance.py:79,85)
4. ``(3/4)``              -> ``(3//4)``              (cosmology.py:464: Python-2
                                                     integer division is part of
                                                     the reference's arithmetic;
                                                     its own golden values at
                                                     unit_test.py:185 need it)
5. ``1/b`` (kernel.py:167) -> ``_py2compat.py2_div(1, b)``
6. first line of every module gains ``import _py2compat;`` which installs the
   Romberg restatement as ``scipy.integrate.romberg``.
7. correlation.py:106 ``(k_min < self.halo._k_min or k_max > self.halo._k_max)`` ->
   ``(_py2compat.py2_lt(k_min, ...) or _py2compat.py2_gt(k_max, ...))``: Python 2 orders None
   below every number instead of raising TypeError.
8. covariance.py:60 ``if self.corr_a == self.corr_b:`` -> ``if _py2compat.py2_corr_eq(self.corr_a, self.corr_b):``.
   Correlation.__eq__ (correlation.py:125-133) compares the objects' __dict__s.  CPython 2.7 walks the left
   dictionary in hash order -- 'D_z', 'kernel', ... for a Correlation (restated in _py2compat) -- and stops at
   the first unequal value, so two different correlations compare unequal at 'D_z' / 'kernel'; Python 3 walks in
   insertion order, reaches the theta arrays first and raises ValueError.
9. kernel.py:606,608 debug ``.write('test_window_*')`` calls are left alone
   (they are the reference's behaviour) -- callers chdir to a temp dir.
"""
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_SRC = "/root/reference"
DEFAULT_DST = os.path.join(HERE, "_ref")

_PRINT_RE = re.compile(r"^(\s*)print\s+(?!\()(.*?)\s*$")
_PRINT_PAREN_RE = re.compile(r"^(\s*)print\s+\((.*)$")
_EXCEPT_RE = re.compile(r"^(\s*)except\s+(\w+)\s*,\s*(\w+)\s*:")


def transform(name, text):
    lines = text.split("\n")
    out = []
    i = 0
    injected = False
    while i < len(lines):
        line = lines[i]
        m = _PRINT_RE.match(line)
        if m and not line.lstrip().startswith("#"):
            body = m.group(2)
            if body.endswith("\\"):
                # backslash continuation (kernel.py:847-848): keep both lines,
                # close the call on the second one
                nxt = lines[i + 1]
                out.append("%sprint(%s" % (m.group(1), body))
                out.append(nxt.rstrip() + ")")
                i += 2
                continue
            line = "%sprint(%s)" % (m.group(1), body)
        else:
            m2 = _PRINT_PAREN_RE.match(line)
            if m2:
                line = "%sprint(%s" % (m2.group(1), m2.group(2))
        m = _EXCEPT_RE.match(line)
        if m:
            line = "%sexcept (%s, %s):" % m.groups() + line[m.end():]
        line = line.replace("xrange(", "range(")
        if name == "cosmology.py":
            line = line.replace("(Omb2)**(3/4)", "(Omb2)**(3//4)")
        if name == "kernel.py" and line.strip() == "1/b)*":
            line = line.replace("1/b)*", "_py2compat.py2_div(1, b))*")
        if name == "correlation.py" and "(k_min < self.halo._k_min or k_max > self.halo._k_max)):" in line:
            line = line.replace("(k_min < self.halo._k_min or k_max > self.halo._k_max)",
                                "(_py2compat.py2_lt(k_min, self.halo._k_min) or _py2compat.py2_gt(k_max, self.halo._k_max))")
        if name == "covariance.py" and line.strip() == "if self.corr_a == self.corr_b:":
            line = line.replace("self.corr_a == self.corr_b", "_py2compat.py2_corr_eq(self.corr_a, self.corr_b)")
        if (not injected and (line.startswith("import ") or
                              line.startswith("from ")) and
                "__future__" not in line):
            line = "import _py2compat; " + line
            injected = True
        out.append(line)
        i += 1
    return "\n".join(out)


def build(src=DEFAULT_SRC, dst=DEFAULT_DST, quiet=False):
    if not os.path.isdir(src):
        raise FileNotFoundError(src)
    os.makedirs(dst, exist_ok=True)
    names = sorted(n for n in os.listdir(src) if n.endswith(".py"))
    for name in names:
        with open(os.path.join(src, name)) as f:
            text = f.read()
        with open(os.path.join(dst, name), "w") as f:
            f.write(transform(name, text))
    shutil.copy(os.path.join(HERE, "_py2compat.py"),
                os.path.join(dst, "_py2compat.py"))
    # every generated module must at least compile
    import py_compile
    for name in names:
        py_compile.compile(os.path.join(dst, name), doraise=True)
    if not quiet:
        print("oracle/_ref: %d modules generated from %s" % (len(names), src))
    return dst


if __name__ == "__main__":
    build(*(sys.argv[1:3]))
