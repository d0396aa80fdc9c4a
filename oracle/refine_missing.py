"""Oracle (test infrastructure): missing-joint refinement of the bottom-up inferencer.

Restates ``BottomUpHeatMapAEInferencer._refine_missing``
(mindpose/engine/inferencer/bottomup_inferencer.py:189-249) -- SURVEY.md section 8(f), row N3.
Per person: the mean tag of the detected joints (float32 ``np.mean``), then for every joint
``argmax(heatmap - round(||tagging_heatmap - mean_tag||))`` over the map (first occurrence),
``+0.5`` and a ``+-0.25`` shift toward the higher neighbour (minus on ties; neighbours
clamped at the border); the result replaces joints that were not detected
(``val == 0``) when the heat-map value there is positive.  Only x, y, val are written.

PINNED: the function lives in a module that imports MindSpore, so it cannot be imported;
``reference_function()`` below extracts the body of ``_refine_missing`` from the reference
source with ``ast`` and executes it unchanged (it is pure numpy).
tests/golden/refine_missing_ref.npz holds its outputs; the live comparison runs under the
``needs_reference`` marker.
"""
import numpy as np


def refine_missing(heatmap, tagging_heatmap, keypoints):
    """heatmap f32 [K,H,W], tagging_heatmap f32 [K,H,W,L], keypoints f32 [K,4] -> [K,4] (copy)."""
    heatmap = np.asarray(heatmap, np.float32)
    tagging_heatmap = np.asarray(tagging_heatmap, np.float32)
    keypoints = np.array(keypoints, dtype=np.float32, copy=True)
    k, h, w = heatmap.shape
    loc = keypoints[:, :2].astype(np.int32)
    tags = [tagging_heatmap[i, loc[i, 1], loc[i, 0]] for i in range(k) if keypoints[i, 2] > 0]
    mean_tag = np.mean(tags, axis=0)
    dist = np.round(np.linalg.norm(tagging_heatmap - mean_tag[None, None, None, :], axis=3))
    flat = (heatmap - dist).reshape(k, -1)
    best = np.argmax(flat, axis=1)
    ys, xs = np.unravel_index(best, (h, w))
    out = keypoints
    for i in range(k):
        x, y = int(xs[i]), int(ys[i])
        fx = np.float32(x) + np.float32(0.5)
        fy = np.float32(y) + np.float32(0.5)
        if heatmap[i, y, min(x + 1, w - 1)] > heatmap[i, y, max(x - 1, 0)]:
            fx += np.float32(0.25)
        else:
            fx -= np.float32(0.25)
        if heatmap[i, min(y + 1, h - 1), x] > heatmap[i, max(0, y - 1), x]:
            fy += np.float32(0.25)
        else:
            fy -= np.float32(0.25)
        val = heatmap[i, y, x]
        if val > 0 and out[i, 2] == 0:
            out[i, :3] = (fx, fy, val)
    return out


def reference_function(reference_root="/root/reference"):
    """The reference's own ``_refine_missing`` as a plain function (self, heatmap,
    tagging_heatmap, keypoints), extracted from the source file without importing it."""
    import ast
    import os
    import textwrap
    from typing import List  # noqa: F401  (used by the extracted annotations)

    path = os.path.join(reference_root, "mindpose/engine/inferencer/bottomup_inferencer.py")
    src = open(path).read()
    tree = ast.parse(src)
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "_refine_missing":
            code = textwrap.dedent(ast.get_source_segment(src, node))
            ns = {"np": np, "List": List}
            exec(compile(code, path, "exec"), ns)
            return ns["_refine_missing"]
    raise RuntimeError("_refine_missing not found in the reference")
