"""Reads the extrema of the six curves of cell 37 of the reference notebook
examples/notebooks/03_constrained_fwi.ipynb (the last receiver's trace of shots 1, 4, 8 in the
true and the initial model) off the pixels of the embedded PNG.  Runs only where /root/reference
exists (the build container); its output is the table NB_FIG37 of
tests/test_reference_notebook.py.  Needs PIL."""
import base64
import io
import json
import sys

import numpy as np
from PIL import Image

NB = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/examples/notebooks/03_constrained_fwi.ipynb"
cell = json.load(open(NB))["cells"][37]
png = [o["data"]["image/png"] for o in cell["outputs"] if "image/png" in o.get("data", {})][0]
im = np.array(Image.open(io.BytesIO(base64.b64decode(png))).convert("RGB")).astype(int)
dark = im.sum(axis=2) < 200


def mask(c, tol=60):
    return np.abs(im - np.array(c)).sum(axis=2) < tol


curves = {"true": mask((31, 119, 180)), "init": mask((255, 127, 14))}      # matplotlib C0 / C1
frame_rows = np.where((im.sum(axis=2) < 120).sum(axis=1) > 500)[0]          # axes frames
frame_cols = np.where((im.sum(axis=2) < 120).sum(axis=0) > 500)[0]
c0, c1 = frame_cols[0], frame_cols[-1]
labels = {1: [-2, -1, 0, 1, 2, 3], 4: [-2, 0, 2, 4, 6], 8: [-2, -1, 0, 1, 2, 3, 4]}   # y tick labels
for k, shot in enumerate((1, 4, 8)):
    r0, r1 = frame_rows[2 * k], frame_rows[2 * k + 1]
    ticks = [r for r in range(r0 + 1, r1) if dark[r, c0 - 5:c0].sum() >= 3]
    groups = []
    for r in ticks:
        if groups and r - groups[-1][-1] <= 1:
            groups[-1].append(r)
        else:
            groups.append([r])
    rows = [np.mean(g) for g in groups]
    fit = np.polyfit(rows, labels[shot][::-1], 1)
    for name, mk in curves.items():
        sub = mk[r0 + 1:r1, c0 + 421:c1]                      # right of the legend box
        rr, cc = np.where(sub)
        t = lambda c: 4.5 + (c + 421) / (c1 - c0) * (4000 - 4.5)    # xlim([4.5, 4000])
        print("shot %d %-4s  max %.2f at %4.0f ms   min %.2f at %4.0f ms"
              % (shot, name, np.polyval(fit, rr.min() + r0 + 0.5), t(cc[rr.argmin()]),
                 np.polyval(fit, rr.max() + r0 + 1.5), t(cc[rr.argmax()])))
