#!/usr/bin/env python
"""Loader fuzz fixtures: random OBJ/MTL files in the corners of the format (number spellings, index forms,
negative indices, quads, groups, unknown materials, CRLF, comments), each run through the REFERENCE's own
loader (oracle/_ref/ref_loader_dump, built from /root/reference).  Inputs go to tests/golden/assets/fuzz/,
the reference's tables to tests/golden/loader_fuzz.npz; tests/test_oracle.py and tests/test_host.py hold the
oracle's and the product's loaders to them bit for bit.

    python tools/make_loader_fuzz.py        (needs /root/reference; the committed fixtures do not)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
OUT = os.path.join(GOLD, "assets", "fuzz")
N_CASES = 28        # cases 0..27: triangles and quads
N_POLY_CASES = 20   # cases 28..47: also polygons with 5..9 corners (the reference ear-clips them)


def spell(rng, x):
    """one of the spellings tinyobj's number grammar accepts"""
    k = rng.randint(0, 9)
    if k == 0:
        return "%d" % int(round(x)) if abs(x) < 1000 else "%.1f" % x
    if k == 1:
        return "%.3e" % x
    if k == 2:
        return "%.4E" % x
    if k == 3:
        return ("+" if x >= 0 else "") + "%.5f" % x
    if k == 4:
        s = "%.3f" % x
        return s.replace("0.", ".", 1) if s.startswith("0.") or s.startswith("-0.") else s
    if k == 5:
        return "%d." % int(x)
    if k == 6:
        return "%.9g" % x
    if k == 7:
        return "%.2fe%+d" % (x / 10.0, 1)
    return repr(float(np.float32(x)))


def make_case(seed):
    rng = np.random.RandomState(1000 + seed)
    nl = "\r\n" if seed % 5 == 3 else "\n"
    sep = (lambda: " ") if seed % 4 else (lambda: rng.choice([" ", "  ", "\t", " \t "]))
    nv = int(rng.randint(5, 40))
    n_mode = seed % 3            # 0: no normals, 1: one per vertex, 2: a different count
    nn = 0 if n_mode == 0 else (nv if n_mode == 1 else int(rng.randint(1, nv)))
    nt = int(rng.randint(0, 6))
    n_mat = int(rng.randint(0, 4))
    names = ["m%d_%d" % (seed, i) for i in range(n_mat)]
    mtl = ["# fuzz case %d" % seed, ""]
    for nm in names:
        mtl.append("newmtl " + nm)
        fields = [("Kd", 3), ("Ks", 3), ("Ke", 3), ("Ns", 1), ("Ni", 1), ("d", 1), ("illum", 0), ("Ka", 3), ("Tf", 3), ("map_Kd", -1), ("Tr", 1)]
        rng.shuffle(fields)
        for key, n in fields[: int(rng.randint(2, len(fields) + 1))]:
            if n == -1:
                mtl.append("map_Kd tex_%d.png" % seed)
            elif n == 0:
                mtl.append("illum %d" % rng.randint(0, 8))
            elif key == "Ke":
                on = rng.rand() < 0.5
                mtl.append(key + sep() + sep().join(spell(rng, v) for v in (rng.uniform(0, 12, 3) if on else np.zeros(3))))
            elif key == "Ns":
                mtl.append(key + sep() + spell(rng, rng.uniform(0, 50)))
            else:
                mtl.append(key + sep() + sep().join(spell(rng, v) for v in rng.uniform(0, 1, n)))
        if rng.rand() < 0.5:
            mtl.append("")
        if rng.rand() < 0.3:
            mtl.append("   # indented comment")
    obj = ["# fuzz case %d" % seed]
    if n_mat and seed % 7 != 6:
        obj.append("mtllib case_%02d.mtl" % seed)
    verts = rng.uniform(-3, 3, (nv, 3))
    planar = []   # (first vertex, corner count) of planar polygons appended to the vertex list
    if seed >= N_CASES:
        for _ in range(int(rng.randint(1, 4))):
            k = int(rng.randint(5, 10))
            o, a, b = rng.uniform(-2, 2, 3), rng.normal(size=3), rng.normal(size=3)
            ang = np.sort(rng.uniform(0, 2 * np.pi, k))
            rad = rng.uniform(0.3, 1.5, k) if rng.rand() < 0.6 else np.full(k, 1.0)   # star-shaped (concave) or convex
            if rng.rand() < 0.5:
                ang = ang[::-1]
            ring = o + rad[:, None] * (np.cos(ang)[:, None] * a + np.sin(ang)[:, None] * b)
            planar.append((verts.shape[0], k))
            verts = np.concatenate([verts, ring])
        nv = verts.shape[0]
        if n_mode == 1:
            nn = nv
    lines_v = ["v" + sep() + sep().join(spell(rng, c) for c in v) + ("  # c" if rng.rand() < 0.1 else "") for v in verts]
    lines_n = ["vn" + sep() + sep().join(spell(rng, c) for c in n) for n in rng.normal(size=(nn, 3))]
    lines_t = ["vt" + sep() + sep().join(spell(rng, c) for c in t) for t in rng.uniform(0, 1, (nt, 2))]
    interleave = seed % 2 == 0
    if interleave:
        obj += lines_v + lines_t + lines_n
    else:
        pool = lines_v + lines_n + lines_t
        # keep relative order inside each kind (indices stay meaningful), shuffle the kinds' interleaving
        tags = ["v"] * len(lines_v) + ["n"] * len(lines_n) + ["t"] * len(lines_t)
        rng.shuffle(tags)
        it = {"v": iter(lines_v), "n": iter(lines_n), "t": iter(lines_t)}
        obj += [next(it[t]) for t in tags]
        del pool
    nf = int(rng.randint(1, 30))
    for f in range(nf):
        r = rng.rand()
        if r < 0.12:
            obj.append(rng.choice(["g grp_%d" % f, "o obj_%d" % f, "s 1", "s off", "g", "# mid comment", ""]))
        if n_mat and rng.rand() < 0.25:
            obj.append("usemtl " + (rng.choice(names) if rng.rand() < 0.85 else "no_such_material"))
        k = 3 if rng.rand() < 0.6 else 4
        if seed >= N_CASES and rng.rand() < 0.35:
            if planar and rng.rand() < 0.6:
                first, cnt = planar[int(rng.randint(0, len(planar)))]
                idx = np.arange(first, first + cnt)
            else:
                idx = rng.choice(nv, int(rng.randint(5, 9)), replace=False)     # arbitrary, usually non-planar
        elif k == 4:
            # a planar-ish convex quad from a random parallelogram so both diagonals are plausible
            idx = rng.choice(nv, 4, replace=False)
        else:
            idx = rng.choice(nv, 3, replace=False)
        form = rng.randint(0, 5)
        toks = []
        for i in idx:
            vi = int(i) + 1 if rng.rand() < 0.7 else int(i) - nv          # absolute or relative (negative) index
            if form == 0 or (form in (1, 3) and nt == 0) or (form in (2, 3) and nn == 0):
                toks.append("%d" % vi)
            elif form == 1:
                toks.append("%d/%d" % (vi, rng.randint(1, nt + 1)))
            elif form == 2:
                toks.append("%d//%d" % (vi, rng.randint(1, nn + 1)))
            elif form == 3:
                toks.append("%d/%d/%d" % (vi, rng.randint(1, nt + 1), rng.randint(1, nn + 1)))
            else:
                toks.append("%d" % vi)
        obj.append("f" + sep() + sep().join(toks))
    return nl.join(obj) + nl, nl.join(mtl) + nl


def main():
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    O.build()
    if not os.path.exists(O.REF_DUMP):
        raise SystemExit("oracle/_ref/ref_loader_dump is missing (needs /root/reference)")
    os.makedirs(OUT, exist_ok=True)
    out, kept = {}, []
    for seed in range(N_CASES + N_POLY_CASES):
        obj, mtl = make_case(seed)
        po, pm = os.path.join(OUT, "case_%02d.obj" % seed), os.path.join(OUT, "case_%02d.mtl" % seed)
        open(po, "w", newline="").write(obj)
        open(pm, "w", newline="").write(mtl)
        try:
            d = O.ref_loader_dump(po, pm)
        except Exception as e:   # the reference loader rejected the file: not a fixture
            print("case %02d rejected by the reference loader: %s" % (seed, e))
            os.remove(po)
            os.remove(pm)
            continue
        key = "case_%02d" % seed
        kept.append(key)
        out[key + "_V"] = d["V"]
        out[key + "_N"] = d["N"]
        out[key + "_F"] = d["F"]
        out[key + "_M"] = np.array([m["kd"] + m["ks"] + m["ke"] + [m["ns"], m["ni"], m["d"], float(m["illum"])] for m in d["M"]],
                                   np.float32).reshape(-1, 13)
        print(key, "V", d["V"].shape[0], "N", d["N"].shape[0], "F", d["F"].shape[0], "M", len(d["M"]), "unassigned faces", int((d["F"][:, 3] < 0).sum()))
    out["cases"] = np.array(kept)
    np.savez_compressed(os.path.join(GOLD, "loader_fuzz.npz"), **out)


if __name__ == "__main__":
    main()
