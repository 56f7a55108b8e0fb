#!/usr/bin/env python3
"""Extract the Butcher tableaus from the reference source into a golden fixture.

Reads /root/reference/src/runge_kutta.jl:66-157 (only available in the build
container, not on the GPU box) and writes tests/golden/tableaus.json with every
coefficient as an exact "num/den" string, plus the reference's `order` tuple.
Run:  python tools/gen_golden_tableaus.py
"""
import json
import os
import re
from fractions import Fraction

SRC = "/root/reference/src/runge_kutta.jl"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "tableaus.json")


def split_top_level_arrays(body):
    """Return the top-level [...] literals of a call body."""
    out, depth, start = [], 0, None
    for i, ch in enumerate(body):
        if ch == "[":
            if depth == 0:
                start = i
            depth += 1
        elif ch == "]":
            depth -= 1
            if depth == 0:
                out.append(body[start + 1:i])
    return out


def parse_matrix(txt):
    rows = []
    for line in txt.strip().splitlines():
        line = line.strip()
        if not line:
            continue
        toks = re.split(r"[\s,]+", line.strip(", "))
        rows.append([str(Fraction(t.replace("//", "/"))) for t in toks if t])
    return rows


def main():
    src = open(SRC).read()
    res = {}
    for m in re.finditer(r"const (bt_\w+) = TableauRKExplicit\(:(\w+),\s*\(([\d,\s]+)\),\s*Rational\{Int64\},", src):
        name = m.group(1)
        order = [int(x) for x in m.group(3).replace(" ", "").strip(",").split(",") if x]
        # body up to the matching close paren
        i = m.end()
        depth = 1
        j = i
        while depth:
            if src[j] == "(":
                depth += 1
            elif src[j] == ")":
                depth -= 1
            j += 1
        body = src[i:j - 1]
        if name == "bt_feuler":  # zeros(Int,1,1), [1][:,:], [0]
            a, b, c = [["0"]], [["1"]], ["0"]
        else:
            arrs = split_top_level_arrays(body)
            assert len(arrs) == 3, (name, len(arrs))
            a, b = parse_matrix(arrs[0]), parse_matrix(arrs[1])
            c = [x for r in parse_matrix(arrs[2]) for x in r]
        S = len(c)
        assert len(a) == S and all(len(r) == S for r in a), name
        assert all(len(r) == S for r in b) and len(b) == len(order), name
        res[name] = dict(order=order, a=a, b=b, c=c)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print("wrote", OUT, sorted(res))


if __name__ == "__main__":
    main()
