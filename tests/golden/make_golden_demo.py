"""Golden values for the demo query-point text format, produced by the REFERENCE's own parser (build container only).

    python tests/golden/make_golden_demo.py

/root/reference/demo/demo.py cannot be imported (hydra, matplotlib, ... are absent), but `load_query_points` (demo.py:225-252)
is a pure function: its source is cut out of the file with `ast` and executed unmodified."""
import ast
import json
import os
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/demo/demo.py"

CASES = [
    # the README example (data/demo_data/README.md:33-37)
    dict(text="1\n0 ;      10,20       30,30   40,40\n4 ; 123.123,456.456  72,72    5,6\n", frame_stride=1, resize_factor=1.0),
    dict(text="2\n0 ; 1,2 3,4 5,6\n\n8 ; 7.5,8.25 9,10 11,12\n16 ; 0,0 1,1 2,2\n", frame_stride=4, resize_factor=0.5),
    dict(text="3\n12 ; 100.0,200.0 101,201 102,202\n", frame_stride=3, resize_factor=1024 / 1920),
]


def main():
    tree = ast.parse(open(SRC).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "load_query_points")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), SRC, "exec"), ns)
    out = []
    for c in CASES:
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            f.write(c["text"])
        qp, npos = ns["load_query_points"](f.name, c["frame_stride"], c["resize_factor"])
        os.unlink(f.name)
        out.append(dict(c, num_positive_points=npos, shape=list(qp.shape), query_points=qp.tolist()))
    json.dump({"source": "reference demo/demo.py:225-252 executed unmodified", "cases": out},
              open(os.path.join(HERE, "demo_query_points_golden.json"), "w"), indent=1)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
