"""The CMake package (CMakeLists.txt, cmake/TMACConfig.cmake.in): `find_package(TMAC)` + target t_mac_no_tvm + kcfg.ini in
TMAC_LIB_DIR, consumed the way the vendored llama.cpp does (ref:3rdparty/llama.cpp/ggml/src/CMakeLists.txt:851-885).
The library itself is NOT rebuilt here (nvcc takes a minute): the project is only configured, and a prefix is assembled
from the in-tree libtmac_b200.so + the generated TMACConfig.cmake; the consumer project is then built and run (host-only
entry points)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "t-mac_b200", "libtmac_b200.so")


@pytest.mark.skipif(shutil.which("cmake") is None or shutil.which("ninja") is None, reason="cmake / ninja not installed")
@pytest.mark.skipif(not os.path.exists(LIB), reason="libtmac_b200.so not built")
def test_find_package_tmac_consumer(tmp_path):
    build, prefix, cons = tmp_path / "build", tmp_path / "prefix", tmp_path / "consumer"
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    run = lambda *a, **k: subprocess.run(a, check=True, capture_output=True, text=True, **k)
    run("cmake", "-G", "Ninja", "-S", ROOT, "-B", str(build), "-DCMAKE_CUDA_COMPILER=" + nvcc, "-DCMAKE_INSTALL_PREFIX=" + str(prefix),
        "-DTMAC_PRESET=llama-2-7b-2bit")
    (prefix / "lib" / "cmake" / "TMAC").mkdir(parents=True)
    (prefix / "include").mkdir()
    shutil.copy(LIB, prefix / "lib" / "libtmac_b200.so")
    shutil.copy(build / "kcfg.ini", prefix / "lib" / "kcfg.ini")          # generated at configure time (tools/make_kcfg.py)
    shutil.copy(build / "TMACConfig.cmake", prefix / "lib" / "cmake" / "TMAC" / "TMACConfig.cmake")
    shutil.copy(os.path.join(ROOT, "include", "tmac_b200.h"), prefix / "include" / "tmac_b200.h")
    shutil.copytree(os.path.join(ROOT, "t-mac_b200", "include", "t-mac"), prefix / "include" / "t-mac")
    run("cmake", "-G", "Ninja", "-S", os.path.join(ROOT, "tests", "cmake_consumer"), "-B", str(cons), "-DCMAKE_PREFIX_PATH=" + str(prefix))
    run("cmake", "--build", str(cons))
    env = dict(os.environ, LD_LIBRARY_PATH=str(prefix / "lib"))
    out = subprocess.run([str(cons / "consumer")], env=env, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "consumer ok" in out.stdout
    # examples/gguf_gemv: parses the golden file; the upload then fails loudly without a GPU (no CPU fallback) or succeeds on one
    ex = subprocess.run([str(cons / "gguf_gemv"), os.path.join(ROOT, "tests", "golden", "tiny_tmac.gguf"), str(prefix / "lib" / "kcfg.ini"),
                         "blk.0.ffn_up.weight"], env=env, capture_output=True, text=True)
    assert "5 tensors, architecture llama" in ex.stdout and "ggml type 2 (4 bits), 64 x 256" in ex.stdout, ex.stdout + ex.stderr
    assert ex.returncode == 0 or "no CUDA device" in ex.stderr or "upload failed" in ex.stderr
