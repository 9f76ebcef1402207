#!/usr/bin/env python
"""Generate the committed test fixtures under tests/golden/ from the read-only reference tree.

Run in the build container (where /root/reference exists):

    python tools/make_fixtures.py

What it writes (all small, all data -- no reference source code is copied):

* tests/golden/assets/{CornellBox,shapes,scenes}/...  geometry, materials and the 100 scene
  descriptions (golden *inputs*: the loader / renderer parity tests need the exact bytes the
  reference loader would parse; reference: scenes/*.txt, CornellBox/*, shapes/*).
* tests/golden/golden_0_true.json   region means (u8) of the reference's only known-answer
  output preds/0_true.png + the region rectangles used.
* tests/golden/golden_0_true_ds4.npy  the same image box-downsampled 4x (125x125x3 float32 in
  [0,1]) for whole-image RMSE checks.
* tests/golden/golden_0_true.png  the 500x500 target itself (createGraph / inverse-fit target of
  BASELINE config 3; the reference reads it through stbi_load, inv_scene.h:56).

Nothing in tests/, bench.py or smoke() reads /root/reference at run time.
"""
import json
import os
import shutil
import sys

import numpy as np
from PIL import Image

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "..", "tests", "golden")

# (row0,row1,col0,col1) rectangles safely inside one surface each; derived from the scene
# geometry (Cornell box x,y in [-2,2], z in [2,6], camera at origin, 90 deg FOV, x flipped).
REGIONS = {
    "light": (100, 135, 225, 275),
    "floor": (440, 490, 150, 350),
    "back_wall": (180, 260, 180, 320),
    "orange_wall_image_right": (200, 300, 400, 480),
    "cyan_wall_image_left": (200, 300, 20, 100),
    "ceiling": (20, 60, 150, 350),
    "cube_top": (308, 319, 230, 270),
    "cube_front": (330, 385, 222, 278),
}


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures can only be regenerated in the build container")
    assets = os.path.join(GOLD, "assets")
    for sub, names in (
        ("CornellBox", ["CornellBox-Empty-CO.obj", "CornellBox-Empty-CO.mtl"]),
        ("shapes", ["cube.obj", "sphere.obj", "scene.obj"]),
        ("scenes", ["%d.txt" % i for i in range(100)]),
    ):
        os.makedirs(os.path.join(assets, sub), exist_ok=True)
        for n in names:
            shutil.copyfile(os.path.join(REF, sub, n), os.path.join(assets, sub, n))
            os.chmod(os.path.join(assets, sub, n), 0o644)

    im = np.asarray(Image.open(os.path.join(REF, "preds", "0_true.png")).convert("RGB"))
    assert im.shape == (500, 500, 3)
    shutil.copyfile(os.path.join(REF, "preds", "0_true.png"), os.path.join(GOLD, "golden_0_true.png"))
    os.chmod(os.path.join(GOLD, "golden_0_true.png"), 0o644)
    means = {}
    for k, (r0, r1, c0, c1) in REGIONS.items():
        means[k] = [float(x) for x in im[r0:r1, c0:c1].reshape(-1, 3).mean(0)]
    with open(os.path.join(GOLD, "golden_0_true.json"), "w") as f:
        json.dump({"source": "preds/0_true.png", "shape": [500, 500, 3], "regions": REGIONS,
                   "region_means_u8": means}, f, indent=1)
    ds = im.astype(np.float32).reshape(125, 4, 125, 4, 3).mean((1, 3)) / 255.0
    np.save(os.path.join(GOLD, "golden_0_true_ds4.npy"), ds.astype(np.float32))
    for k, v in means.items():
        print("%-26s %6.1f %6.1f %6.1f" % (k, *v))


def loader_tables():
    """Golden loader tables from the REFERENCE's own loader (oracle/_ref/ref_loader_dump, built from
    /root/reference by oracle/Makefile): per asset the vertex floats, triangle vertex indices and
    per-face material ids, and the materials.  Committed as tests/golden/loader_tables.npz so the
    loader parity test also runs where the reference tree is absent."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from oracle import oracle as O
    O.build()
    out = {}
    assets = os.path.join(GOLD, "assets")
    for key, obj, mtl in (("cornell", "CornellBox/CornellBox-Empty-CO.obj", "CornellBox/CornellBox-Empty-CO.mtl"),
                          ("cube", "shapes/cube.obj", "/nonexistent.mtl"), ("sphere", "shapes/sphere.obj", "/nonexistent.mtl"),
                          ("scene", "shapes/scene.obj", "/nonexistent.mtl")):
        d = O.ref_loader_dump(os.path.join(assets, obj), os.path.join(assets, mtl) if not mtl.startswith("/") else mtl)
        out[key + "_V"] = d["V"]
        out[key + "_N"] = d["N"]
        out[key + "_F"] = d["F"]
        out[key + "_M"] = np.array([m["kd"] + m["ks"] + m["ke"] + [m["ns"], m["ni"], m["d"], float(m["illum"])] for m in d["M"]],
                                   np.float32).reshape(-1, 13)
        out[key + "_Mnames"] = np.array([m["name"] for m in d["M"]])
    np.savez_compressed(os.path.join(GOLD, "loader_tables.npz"), **out)
    print("loader tables:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
    loader_tables()
