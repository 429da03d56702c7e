"""Mechanical Python 2 -> Python 3 source shim for the reference's layer modules.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference's `code/lib` is Python 2:
print statements, xrange, integer '/'.  This shim is everything that is done to its sources
before they are executed here -- by tests/golden/make_layers_golden.py (golden fixtures generated
by the reference's own layer code) and by oracle/build_ref.py:build_callsites (the reference's
call sites compiled to code objects under oracle/_ref/callsites/ for the drop-in test)."""
import re


def py2_to_py3(src, int_div_lines=()):
    """print statements -> calls (following continuation lines), xrange -> range, and `/` ->
    `//` on the named 1-based lines."""
    lines = src.split("\n")
    out = []
    i = 0
    in_triple = False                    # inside a triple-quoted string at the start of the line
    while i < len(lines):
        line = lines[i]
        started_in_triple = in_triple
        if (line.count('"""') + line.count("'''")) % 2 == 1:
            in_triple = not in_triple
        if started_in_triple:
            out.append(line)
            i += 1
            continue
        if (i + 1) in int_div_lines:
            line = line.replace(" / ", " // ")
        m = re.match(r"^(\s*)print\s+(?!\()(.*)$", line) or re.match(r"^(\s*)print\s*$", line)
        if m and not line.lstrip().startswith("#"):
            indent = m.group(1)
            rest = m.group(2) if m.lastindex and m.lastindex >= 2 else ""
            buf = [rest]
            depth = rest.count("(") + rest.count("[") - rest.count(")") - rest.count("]")
            while depth > 0 or buf[-1].rstrip().endswith("\\"):
                i += 1
                buf.append(lines[i])
                depth += lines[i].count("(") + lines[i].count("[") - lines[i].count(")") - lines[i].count("]")
            out.append(indent + "print(" + "\n".join(buf) + ")")
        else:
            out.append(line)
        i += 1
    return "\n".join(out).replace("xrange", "range")
