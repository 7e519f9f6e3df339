"""CPU: the hook-up of the CUDA layer into the reference's driver and command line is a checked artefact:
patches/hydracore_cuda_layer.patch must apply to the reference tree as it is (git apply --check), and the C++ layer must compile against the
PATCHED IHWLayer.h (the factory declaration and the GPU_RT_HW_LAYER_CUDA flag it adds).  A full build of hydra needs HydraAPI, Embree and
FreeImage, which are not vendored in the reference (CMakeLists.txt:11-15), so RenderDriverRTE.cpp / input.cpp / main.cpp are only checked to apply."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PATCH = os.path.join(ROOT, "patches", "hydracore_cuda_layer.patch")
FILES = ["hydra_drv/IHWLayer.h", "hydra_drv/RenderDriverRTE.cpp", "hydra_drv/CMakeLists.txt", "hydra_app/input.cpp", "hydra_app/input.h", "hydra_app/main.cpp"]


def test_patch_touches_the_documented_files():
    txt = open(PATCH).read()
    for f in FILES:
        assert ("+++ b/" + f) in txt, f
    assert "CreateCudaImpl" in txt and "GPU_RT_HW_LAYER_CUDA" in txt and "-cuda_device_id" in txt


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "hydra_drv")), reason="the reference tree exists in the build container only")
def test_patch_applies_to_the_reference_and_the_layer_compiles_against_it(tmp_path):
    for f in FILES:
        dst = tmp_path / f
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copy(os.path.join(REF, f), dst)
    r = subprocess.run(["git", "apply", "--check", "-p1", PATCH], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    subprocess.run(["git", "apply", "-p1", PATCH], cwd=tmp_path, check=True)
    hdr = (tmp_path / "hydra_drv" / "IHWLayer.h").read_text()
    assert "IHWLayer* CreateCudaImpl(int w, int h, int a_flags, int a_deviceId);" in hdr and "GPU_RT_HW_LAYER_CUDA" in hdr
    drv = (tmp_path / "hydra_drv" / "RenderDriverRTE.cpp").read_text()
    assert "m_pHWLayer = CreateCudaImpl(m_width, m_height, m_initFlags, m_devId);" in drv
    # GPUCUDALayer.cpp against the PATCHED header (searched first), the rest of the reference headers in place
    shim = os.path.join(ROOT, "hydracore_b200", "cpp", "compat", "ref_shim")
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-w", "-msse4.2", "-I" + str(tmp_path / "hydra_drv"), "-I" + os.path.join(REF, "hydra_drv"),
           "-I" + os.path.join(shim, "x", "y"), "-I" + os.path.join(shim, "HydraAPI"), "-I" + os.path.join(shim, "HydraAPI", "hydra_api"),
           "-I" + os.path.join(ROOT, "hydracore_b200", "cpp", "compat"), "-I" + os.path.join(ROOT, "include"),
           "-DHC_CHECK_PATCHED_HEADER", os.path.join(ROOT, "hydracore_b200", "cpp", "GPUCUDALayer.cpp")]
    env = dict(os.environ)
    env.pop("CXX", None)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    assert r.returncode == 0, r.stdout[-3000:]
