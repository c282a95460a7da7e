"""The built library really contains the sm_100 mechanisms DESIGN.md claims (CPU test: disassembles libcrb200.so with cuobjdump).
Guards against silent regressions of code generation, e.g. the visibility write of the micro raster falling back to a generic
atomic with a shared-memory CAS path (ATOM.E.MIN.64 + QSPC), which cost the setup kernel 4 us on C2."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cudaraster-linux_b200", "libcrb200.so")


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    if not os.path.exists(LIB):
        import cudaraster_linux_b200 as crb
        crb.build_library()
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    fns, cur = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            fns[cur] = []
        elif cur is not None:
            fns[cur].append(ln)
    assert fns, "no SASS in the library"
    return fns


def _count(fns, name_part, mnemonic):
    return {k: sum(mnemonic in ln for ln in v) for k, v in fns.items() if name_part in k}


def test_library_is_sm100a_only(sass):
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_micro_raster_uses_fire_and_forget_reductions(sass):
    setup = _count(sass, "triangleSetupKernel", "REDG.E.MIN.64")
    single = {k: v for k, v in setup.items() if "ELi0ELj3E" in k or "ELi0ELj1E" in k}   # single-sample, depth-tested pipes
    assert single and all(v >= 4 for v in single.values()), single
    assert sum(_count(sass, "triangleSetupKernel", "ATOM.E.MIN.64").values()) == 0, "generic 64-bit atomic min is back"
    assert sum(_count(sass, "triangleSetupKernel", "ATOMS.CAST.SPIN").values()) == 0, "shared-memory CAS fallback of a generic atomic is back"


def test_wide_record_accesses(sass):
    assert sum(_count(sass, "fineRaster", "ENL2.256.CONSTANT").values()) > 0   # LDG.E[.NA].ENL2.256.CONSTANT (NA = no L1 allocation)
    st = _count(sass, "triangleSetupKernel", "STG.E.ENL2.256")
    assert sum(st.values()) > 0
    # ptxas 12.9 miscompiles st.global.v8.b32 inside non-inlined functions: the wide store must only exist on the inlined fast
    # path, i.e. exactly two per kernel instance that has a micro path, none in the kernels without one (MSAA)
    assert all(v in (0, 2) for v in st.values()), st


def test_cluster_scan_and_bulk_store(sass):
    assert sum(_count(sass, "binScanKernel", "UCGABAR_ARV").values()) > 0, "bin scan no longer synchronises a thread-block cluster"
    assert sum(_count(sass, "fineRasterSingleKernel", "UBLKCP").values()) > 0, "tile-major colour tiles no longer leave through cp.async.bulk"


def test_programmatic_dependent_launch_everywhere(sass):
    # the kernels of a frame (setup, bin / coarse or direct alloc / scatter, fine raster); helpers outside the frame chain
    # (resolve, IPC marks, probes, vertex shaders) are launched in plain stream order
    kernels = [k for k in sass if re.search(r"triangleSetupKernel|binScanKernel|binScatterKernel|coarseScanKernel|coarseScatterKernel|directAllocKernel|directScatterKernel|fineRaster(Single|Multi)Kernel", k)]
    assert len(kernels) > 20
    missing = [k for k in kernels if not any("ACQBULK" in ln for ln in sass[k])]
    assert not missing, missing[:3]
