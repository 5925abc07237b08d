"""The C ABI: the library loads, exports every symbol include/genlib_cuda.h declares,
and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "genlib_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(genlib_[a-z0-9_]+)\s*\(", text)))


def test_header_and_library_agree(gen):
    from genlib_jl_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and header drifted apart"
    L = C.CDLL(gen.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} not exported"
    assert gen.lib().genlib_version() == 2


def test_layer_info_struct_layout(gen):
    from genlib_jl_b200 import _lib
    assert C.sizeof(_lib.LayerInfo) == 88 and C.sizeof(_lib.Stats) == 96


def test_no_cpu_fallback(gen):
    if gen.lib().genlib_device_count() > 0:
        pytest.skip("a CUDA device is present")
    ped = gen.genealogy(gen.geneaJi)
    with pytest.raises(gen.GenlibError) as e:
        gen.phi(ped)
    assert e.value.status == 4            # GENLIB_ECUDA
    with pytest.raises(gen.GenlibError):
        gen.phi_arrays(ped.father, ped.mother, ped.rank_of([1, 2, 29]))


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "genlib.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".jl", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle"


def _build_c_client(tmp_path):
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "phi_cabi")
    libdir = os.path.join(root, "genlib.jl_b200")
    subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "examples", "phi_cabi.c"), "-o", exe, "-L" + libdir, "-lgenlib_cuda",
                    "-Wl,-rpath," + libdir], check=True)
    return exe, os.path.join(root, "tests", "data", "geneaJi.csv")


def test_plain_c_client_links_and_fails_loudly_without_a_gpu(gen, tmp_path):
    """examples/phi_cabi.c drives the whole path through the C ABI (no Python, no torch): it must
    compile against include/genlib_cuda.h, and without a CUDA device the engine must refuse
    (GENLIB_ECUDA) instead of falling back to anything."""
    import subprocess
    exe, csv = _build_c_client(tmp_path)
    res = subprocess.run([exe, csv], capture_output=True, text=True)
    if gen.lib().genlib_device_count() > 0:
        assert res.returncode == 0 and len(res.stdout.splitlines()) == 3
    else:
        assert res.returncode == 4 and "no CPU fallback" in res.stderr and res.stdout == ""
    res = subprocess.run([exe, "/nonexistent.csv"], capture_output=True, text=True)
    assert res.returncode != 0 and res.stderr


@pytest.mark.gpu
def test_plain_c_client_on_gpu(tmp_path):
    import subprocess
    import numpy as np
    exe, csv = _build_c_client(tmp_path)
    want = np.array([[0.591796875, 0.37109375, 0.072265625], [0.37109375, 0.591796875, 0.072265625],
                     [0.072265625, 0.072265625, 0.53515625]])                  # test/runtests.jl:51-52
    import genlib_b200 as gen
    runs = [[exe, csv], [exe, csv, "sparse"], [exe, csv, "devices=0"]]
    if gen.lib().genlib_device_count() >= 2:                                   # one process, two GPUs (genlib_phi_multi)
        runs.append([exe, csv, "devices=0,1"])
    for args in runs:
        res = subprocess.run(args, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        got = np.array([[float(v) for v in line.split()] for line in res.stdout.splitlines()])
        assert np.array_equal(got, want)
