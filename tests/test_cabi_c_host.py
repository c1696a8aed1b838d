"""The boundary is a C ABI: the header must compile as plain C and a C program with no Python / PyTorch in it must be able
to drive the library (tests/cabi/host_example.c: create, load 34 float32 tensors, one model call, one short sampling chain
through the host-buffer entry).  CPU: it builds, links and fails loudly without a device.  GPU: it runs."""
import os
import shutil
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "s1-to-s2_super-resolution_project-code_b200", "s1s2_b200")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def _build(tmp_path):
    import __graft_entry__ as ge
    ge.build()
    exe = str(tmp_path / "host_example")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           os.path.join(ROOT, "tests", "cabi", "host_example.c"), "-o", exe, "-L", LIBDIR, "-ls1s2_b200",
           "-L", os.path.join(CUDA, "lib64"), "-lcudart", "-lm", f"-Wl,-rpath,{LIBDIR}", f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


needs_gcc = pytest.mark.skipif(shutil.which("gcc") is None or not os.path.exists(os.path.join(CUDA, "include", "cuda_runtime_api.h")),
                               reason="needs gcc and the CUDA toolkit headers")


@needs_gcc
def test_header_is_plain_c(tmp_path):
    src = tmp_path / "hdr.c"
    src.write_text('#include "s1s2_b200.h"\nint main(void) { return S1S2_OK; }\n')
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


@needs_gcc
@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_c_host_builds_and_fails_loudly_without_a_device(tmp_path):
    res = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 77, (res.returncode, res.stdout, res.stderr)
    assert "no CPU fallback" in res.stdout


@needs_gcc
@pytest.mark.gpu
def test_c_host_runs_on_the_gpu(tmp_path):
    res = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, (res.returncode, res.stdout, res.stderr)
    assert res.stdout.strip().endswith("ok") and "forward: sum" in res.stdout and "sample_host: sum" in res.stdout
