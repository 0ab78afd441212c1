"""The index maps of the pre-arranged weight tables and of the split rows are bijections (checked on the host: the maps
are __host__ __device__ functions of tilingnn_b200/csrc/layouts.cuh and hsplit.cuh)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which(os.environ.get("NVCC", "nvcc")) is None, reason="needs nvcc")
def test_table_index_maps_are_bijections(tmp_path):
    exe = str(tmp_path / "layout_check")
    src = os.path.join(ROOT, "tests", "native", "layout_check.cu")
    r = subprocess.run([os.environ.get("NVCC", "nvcc"), "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                        "-I" + os.path.join(ROOT, "tilingnn_b200", "csrc"), src, "-o", exe],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    print(r.stdout)
    assert r.returncode == 0 and "FAIL" not in r.stdout, r.stdout
