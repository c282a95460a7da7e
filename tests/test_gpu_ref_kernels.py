"""GPU: the REFERENCE's own CUDA kernels, rebuilt for sm_100a from the sources under /root/reference
(oracle/_ref/libcrref_cuda.so, built in the container by oracle/Makefile and shipped to the box),
as a second oracle.  Run in a subprocess with a timeout: the Fermi code is implicitly
warp-synchronous and is not guaranteed to terminate on Blackwell.

What is asserted: the reference's TRIANGLE SETUP kernel (per-thread code) produces bit-identical
triSubtris / triHeader / triData to the CPU oracle -- this pins the oracle's snap / cull / clip /
plane-equation arithmetic to the reference's real device code.  The UNMODIFIED bin / coarse / fine kernels mis-execute on
sm_100a (implicitly warp-synchronous Fermi code, profiles/r1_ref_kernels.md); with the synchronisation patch
(oracle/ref_kernels/b200_sync_patch.py -> oracle/_ref/libcrref_cuda_sync.so: lock-step assumptions made explicit, nothing else
changed) the WHOLE reference pipeline runs on the B200 and its frames are asserted equal to the oracle's and to the product's:
queue order, LESS depth test and ROP order are thereby pinned to the reference's own kernels, not only to the oracle's reading."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(workload, libname="libcrref_cuda.so", extra=("--check", "--check-setup")):
    lib = os.path.join(ROOT, "oracle", "_ref", libname)
    if not os.path.exists(lib):
        pytest.skip("oracle/_ref/%s not built (needs /root/reference at build time)" % libname)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_ref_kernels.py"), "--workload", workload, "--frames", "2"] + list(extra),
                       capture_output=True, text=True, timeout=300, env=dict(os.environ, CRREF_LIBRARY=lib))
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, "reference kernels failed: rc=%d %s" % (r.returncode, (r.stderr or r.stdout)[-500:])
    return json.loads(lines[-1])


@pytest.mark.parametrize("workload", ["soup", "soup_pass", "soup_msaa"])
def test_reference_setup_kernel_matches_oracle(workload):
    out = _run(workload)
    print(json.dumps(out))
    assert out["setup"]["status"] == "bit-exact", out["setup"]
    assert out["setup"]["single"] > 1000 and out["setup"]["clipped"] > 100


@pytest.mark.parametrize("size", [(2048, 2048), (1920, 1080), (640, 360)])
def test_reference_device_functions_match_oracle(size):
    """The oracle's coverage rule, MSAA sample coverage, barycentrics / Gouraud colour and blend
    arithmetic against the reference's OWN device functions (the LUT coverage path of its fine raster
    included) executed per thread on the GPU: bit-exact (colour within 1 LSB)."""
    lib = os.path.join(ROOT, "oracle", "_ref", "libcrref_cuda.so")
    if not os.path.exists(lib):
        pytest.skip("oracle/_ref/libcrref_cuda.so not built (needs /root/reference at build time)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_ref_devfuncs.py"), "--width", str(size[0]), "--height", str(size[1]), "--tris", "24000"],
                       capture_output=True, text=True, timeout=600)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, "harness failed: rc=%d %s" % (r.returncode, (r.stderr or r.stdout)[-800:])
    out = json.loads(lines[-1])
    print(json.dumps(out))
    assert out["status"] == "ok", out
    assert out["cover8x8_exact_fast"]["mismatch"] == 0 and out["cover8x8_exact_fast"]["nonempty"] > 5000
    for s in (1, 2, 3):
        assert out["coverMSAA_fast_s%d" % s]["mismatch"] == 0 and out["coverMSAA_fast_s%d" % s]["partial"] > 100
    for s in (0, 1, 2, 3):
        assert out["shade_gouraud_s%d" % s]["bary_bit_mismatch"] == 0 and out["shade_gouraud_s%d" % s]["color_max_lsb"] <= 1


@pytest.mark.parametrize("workload", ["c1", "soup", "soup_pass", "soup_blend", "soup_msaa_front", "soup_msaa", "ties", "c2", "c4"])
def test_reference_pipeline_sync_patched_equals_oracle_and_product(workload):
    """The reference's four kernels (synchronisation patch only) on the B200: C1 cube, clipped / w <= 0 soups, ordered SrcOver
    blending through its ROP, 4x MSAA (per-sample ROP on the surfaces), a duplicated mesh whose fragments all tie in depth, and BASELINE configs 2 and 4 at FULL size.
    Depth bit-exact and colour exact against the oracle AND against the product's frame."""
    out = _run(workload, "libcrref_cuda_sync.so", ("--check", "--check-product") + (("--check-setup",) if workload.startswith("soup") else ()))
    print(json.dumps(out))
    if workload == "soup_msaa":
        # The one place where the reference's kernels and the oracle may part: a sub-triangle left by clipping at w ~ 0 whose
        # fixed-point depth plane yields depths BELOW the triangle's own zmin; the reference's zmin-based culls
        # (FineRaster.inl:229-241, :1048-1062) then drop fragments the plane would have let through (DESIGN.md "known divergences":
        # 1 sample of 921 600 here; none in soup_msaa_front, the same soup without vertices at w <= 0).  Product == oracle there.
        assert out["depth_mismatch_texels"] <= 4 and out["product_depth_mismatch_texels"] == out["depth_mismatch_texels"]
        for _, _, ref_depth, oracle_depth in out.get("depth_mismatches", []):
            assert ref_depth > oracle_depth      # the reference CULLED a fragment the plane equation puts in front
        return
    assert out["status"] == "ok", out
    assert out["depth_mismatch_texels"] == 0 and out["color_mismatch_texels"] == 0
    assert out["product_depth_mismatch_texels"] == 0 and out["product_color_max_lsb"] <= 1
