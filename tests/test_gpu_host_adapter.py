"""The C++ host side above the C ABI (csrc/host/imrcd_host.hpp): built with g++ against include/imrcd.h, linked to
libimrcd.so and run on the GPU; and, where the reference checkout is present, the engine drop-in
(csrc/host/CollisionDetection_drop_in.hpp) is compiled against the reference's own headers."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "inmyroom_vulkan_b200", "csrc", "host")


@pytest.mark.gpu
def test_cpp_host_adapter(tmp_path, gpu_ctx):
    exe = str(tmp_path / "test_host_adapter")
    lib = os.path.join(ROOT, "inmyroom_vulkan_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", HOST, os.path.join(ROOT, "tests", "cpp", "test_host_adapter.cpp"), "-o", exe,
                    "-L", lib, "-limrcd", f"-Wl,-rpath,{lib}", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"], check=True)
    import torch
    n_gpus = min(torch.cuda.device_count(), 2)          # with two GPUs also the group constructor (one process, N GPUs, in-library NCCL merge)
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "gltf_scene.glb"), str(n_gpus)], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0 and "host adapter ok" in r.stdout, r.stdout + r.stderr
    if n_gpus > 1:
        assert f"host adapter ok on {n_gpus} gpus" in r.stdout


def test_drop_in_compiles_against_the_reference_headers():
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "inMyRoom_vulkan", "include")):
        pytest.skip("reference checkout not present")
    r = subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-w", "-DENABLE_CPP_INTERFACE", f"-I{ref}/inMyRoom_vulkan/include", f"-I{ref}/inMyRoom_vulkan/shaders",
                        f"-I{ref}/glm", f"-I{ref}/eig3", "-I", HOST, "-x", "c++", os.path.join(HOST, "CollisionDetection_drop_in.hpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
