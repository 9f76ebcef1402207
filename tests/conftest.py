import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
ASSETS = os.path.join(GOLDEN, "assets")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def assets():
    return ASSETS


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand with `make -C oracle`."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def ipt():
    """The product package; the shared library must already be built (python -m inverse_path_tracer_b200.build)."""
    import inverse_path_tracer_b200 as pkg
    from inverse_path_tracer_b200 import native
    if not os.path.exists(native.library_path()):
        from inverse_path_tracer_b200 import build
        build.build()
    native.lib()
    return pkg


def scene_objects(name):
    """Object lists used across tests (paths relative to ASSETS)."""
    cornell = {"pos": [0, 0, 4], "ori": [0, 0, 0], "scl": [2, 2, 2], "obj": "./CornellBox/CornellBox-Empty-CO.obj",
               "mtl": "./CornellBox/CornellBox-Empty-CO.mtl"}
    if name == "cornell":
        return [cornell]
    if name == "cornell_sphere":   # per-vertex normals (Triangle::getNormal) + 1298 triangles -> BVH mode
        return [cornell, {"pos": [0.3, -1.2, 4.2], "ori": [0.2, 0.5, 0.1], "scl": [1.4, 1.4, 1.4], "obj": "./shapes/sphere.obj",
                          "mtl": "*Kd 0.8 0.3 0.6*"}]
    if name == "rotated_cubes":    # rotation / non-uniform scale / five objects
        objs = [cornell]
        rng = np.random.RandomState(3)
        for i in range(4):
            objs.append({"pos": [float(rng.uniform(-1.2, 1.2)), float(rng.uniform(-1.5, 0.5)), float(rng.uniform(3, 5))],
                         "ori": [float(x) for x in rng.uniform(-1, 1, 3)], "scl": [float(x) for x in rng.uniform(0.3, 0.9, 3)],
                         "obj": "./shapes/cube.obj", "mtl": "*Kd %f %f %f*" % tuple(rng.uniform(0, 1, 3))})
        return objs
    raise KeyError(name)


def random_rays(n, seed, origin_box=((-1.9, -1.9, 0.0), (1.9, 1.9, 5.9))):
    rng = np.random.RandomState(seed)
    lo, hi = np.array(origin_box[0]), np.array(origin_box[1])
    p = rng.uniform(lo, hi, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([p, d], 1).astype(np.float32)
