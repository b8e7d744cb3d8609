"""Golden vectors for the array steps either side of the LPSR stage, generated FROM THE UNMODIFIED REFERENCE:
``format_long_plate`` / ``restack_to_square`` of /root/reference/inference/run.py:21-78 (the two function definitions are executed from
the reference's own source text, so run.py's heavy imports -- loguru, the YOLOv5 tree -- are not needed) and OpenCV's
``cvtColor(COLOR_RGB2BGR)`` on the (H, W, 1) SR output (run.py:204).  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_pipeline.py
"""
import ast
import os
from typing import Tuple  # noqa: F401  (used by the reference's annotations)

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/inference/run.py"


def reference_functions():
    tree = ast.parse(open(SRC).read())
    ns = {"cv2": cv2, "np": np, "Tuple": Tuple}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("format_long_plate", "restack_to_square"):
            exec(compile(ast.Module([node], []), SRC, "exec"), ns)
    return ns["format_long_plate"], ns["restack_to_square"]


def main():
    fmt, restack = reference_functions()
    rng = np.random.default_rng(0)
    shapes = [(40, 50), (41, 50), (33, 47), (20, 90), (30, 45), (30, 46), (64, 64), (2, 3), (3, 2), (17, 25), (32, 192), (31, 191), (24, 37)]
    out = {}
    for i, (h, w) in enumerate(shapes):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        f, changed = fmt(img)
        out[f"in_{i}"] = img
        out[f"fmt_{i}"] = np.ascontiguousarray(f)
        out[f"changed_{i}"] = np.asarray(changed)
        out[f"restack_{i}"] = np.ascontiguousarray(restack(f))          # what run.py:196 feeds the OCR when changed
        out[f"restack_raw_{i}"] = np.ascontiguousarray(restack(img))    # restack on an arbitrary image
        g = rng.integers(0, 256, (h, w, 1), dtype=np.uint8)
        out[f"gray_{i}"] = g
        out[f"bgr_{i}"] = cv2.cvtColor(g, cv2.COLOR_RGB2BGR)             # run.py:204 on the 1-channel SR output
    out["n"] = np.asarray(len(shapes))
    np.savez_compressed(os.path.join(HERE, "pipeline_cases.npz"), **out)
    print("wrote", len(shapes), "cases; opencv", cv2.__version__)


if __name__ == "__main__":
    main()
