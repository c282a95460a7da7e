"""CPU: the C-ABI library loads and exports every entry point include/crb200.h declares; without
a GPU it refuses to create a context (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "crb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crb_[a-z_0-9]+)\s*\(", src)) - {"crb_stage_fn"})


def test_header_functions_are_exported(crb):
    lib = crb.load_library()
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "libcrb200.so does not export %s" % n
    assert lib.crb_abi_version() == 3


def test_builtin_pipes_resolve_by_name(crb):
    lib = crb.load_library()
    for base, s, f, blend in [("passthrough", 0, 1, "BlendReplace"), ("gouraud", 0, 3, "BlendReplace"), ("gouraud", 2, 3, "BlendSrcOver"),
                              ("texPhong", 2, 3, "BlendReplace"), ("gouraud", 3, 3, "BlendReplace")]:
        name = crb.pipe_name(base, s, f, blend)
        for suffix in ("_triangleSetup", "_binRaster", "_coarseRaster", "_fineRaster", "_spec", "_orderIndependent"):
            assert hasattr(lib, name + suffix), name + suffix
    for suffix in ("_triangleSetup", "_binRaster", "_coarseRaster", "_fineRaster", "_spec"):
        assert hasattr(lib, "PixelPipe_passthrough" + suffix)


def test_pipe_spec_layout(crb):
    from cudaraster_linux_b200.binding import _PipeSpec
    lib = crb.load_library()
    spec = _PipeSpec.in_dll(lib, crb.pipe_name("gouraud", 2, 3, "BlendSrcOver") + "_spec")
    assert (spec.samplesLog2, spec.vertexStructSize, spec.renderModeFlags, spec.profilingMode) == (2, 32, 3, 0)
    assert spec.blendShaderName == b"BlendSrcOver"
    assert ctypes.sizeof(_PipeSpec) == 144  # == sizeof(PixelPipeSpec), cuda/PrivateDefs.hpp:147-154


def test_no_cpu_fallback(crb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = crb.load_library()
    ctx = ctypes.c_void_p()
    rc = lib.crb_create(0, ctypes.byref(ctx))
    assert rc == 3 and not ctx.value  # CRB_ERR_NO_DEVICE
    with pytest.raises(crb.CrbError):
        crb.CudaRaster(0)


def test_pack_helpers_match_reference_known_answers(crb):
    lib = crb.load_library()
    # known answers computed with the reference's own host code
    assert lib.crb_pack_abgr(0.2, 0.4, 0.8, 1.0) == 0xFFCC6633
    assert lib.crb_encode_clear_depth(1.0) == 0xFFFFBB3F
    assert lib.crb_encode_clear_depth(0.5) == 0x7FFFFFFF
    assert lib.crb_encode_clear_depth(0.0) == 0x000044C0


def test_runtime_pipe_compiler_builds_and_caches(tmp_path):
    """FW::CudaCompiler (include/cudaraster/CudaCompiler.hpp): compiles a user pixel-pipe source with
    nvcc for sm_100a into a cached shared object that exports the five by-name symbols of the pipe
    (CudaRaster.cpp:190-201); a different -D define gives a different cache entry, the same one a hit."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "cpp", "compile_pipe")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", os.path.join(root, "examples", "cpp")], check=True, capture_output=True)
    src = os.path.join(root, "examples", "cpp", "UserPipes.cu")
    cache = str(tmp_path / "cudacache")
    r1 = subprocess.run([exe, src, cache, "USER_STRIPE_SHIFT=4"], capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0 and "Compiling" in r1.stdout, r1.stdout + r1.stderr
    so = r1.stdout.strip().splitlines()[-1]
    r2 = subprocess.run([exe, src, cache, "USER_STRIPE_SHIFT=4"], capture_output=True, text=True, timeout=600)
    assert r2.returncode == 0 and "cached" in r2.stdout and r2.stdout.strip().splitlines()[-1] == so
    r3 = subprocess.run([exe, src, cache, "USER_STRIPE_SHIFT=5"], capture_output=True, text=True, timeout=600)
    assert r3.returncode == 0 and r3.stdout.strip().splitlines()[-1] != so
    syms = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
    for suffix in ("_triangleSetup", "_binRaster", "_coarseRaster", "_fineRaster", "_spec"):
        assert "PixelPipe_user" + suffix in syms


def test_split_frame_c_equals_python():
    """crb_split_frame (the C++ host layer's sort-first partition) and multigpu.split_frame (Python host layer) are the same function."""
    import ctypes
    import cudaraster_linux_b200 as crb
    from cudaraster_linux_b200 import multigpu
    lib = crb.load_library()
    for fw, fh in ((3840, 2160), (1920, 1080), (7680, 4320), (5120, 2880), (3010, 2000), (2049, 17), (640, 384)):
        for parts in (1, 2, 4, 8, 16):
            n = lib.crb_split_frame(fw, fh, parts, None, 0)
            buf = (ctypes.c_int * (4 * n))()
            assert lib.crb_split_frame(fw, fh, parts, buf, n) == n
            got = [tuple(buf[4 * i:4 * i + 4]) for i in range(n)]
            assert got == multigpu.split_frame(fw, fh, parts), (fw, fh, parts)
