"""The header-only C++ adapter (include/slide_pr/place_recognition.hpp) compiles against the
C-ABI with plain g++ (CPU check) and reproduces the golden result on the GPU."""
import json
import os
import subprocess

import numpy as np
import pytest

import spr_helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "adapter_test")


def build():
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    libdir = os.path.join(ROOT, "slide_slam_b200")
    subprocess.run([gxx, "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "adapter_test.cpp"), "-o", EXE, "-L", libdir,
                    "-l:libslide_pr.so", f"-Wl,-rpath,{libdir}", "-ldl", "-lpthread"], check=True)


def write_maps(tmp_path):
    maps, cases = H.golden_maps(), H.golden_cases()
    c = cases["indoor01_forest_yaml_nodim"]
    paths = []
    for key in (c["ref"], c["qry"]):
        p = tmp_path / f"{key}.txt"
        with open(p, "w") as f:
            for row in maps[key]:
                f.write(" ".join(repr(float(v)) for v in row) + "\n")
        paths.append(str(p))
    return c, paths


def test_adapter_compiles_and_fails_loudly_without_gpu(tmp_path):
    build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    c, paths = write_maps(tmp_path)
    r = subprocess.run([EXE] + paths, capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_adapter_reproduces_golden(tmp_path):
    build()
    c, paths = write_maps(tmp_path)
    r = subprocess.run([EXE] + paths, capture_output=True, text=True, check=True)
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["found"] == c["found"] and out["best"] == c["best_num_inliers"] and out["hyp"] == c["best_hyp_index"]
    assert out["n_matched"] == c["best_num_inliers"]
    assert out["R_t"] == c["R_t"][:6]
    # findInterLoopClosure rebuilds a yaw + xyz transform from xyzYaw (PR.cpp:523-536)
    np.testing.assert_allclose(out["xyz_yaw"], c["xyz_yaw"], rtol=1e-5, atol=1e-9)
    # the SlideGraph entry points ran through the same adapter (indoor maps: 32 / 35 objects, gate 20)
    sg = json.loads(r.stdout.strip().splitlines()[-2])
    assert sg["sg_triangles"][0] > 40 and sg["sg_triangles"][1] > 40 and sg["sg_matches"] >= 0
    assert isinstance(sg["sg_found"], bool) and sg["sg_found"] == sg["sc_found"]
