"""Generates tests/golden/ref_host_vectors.json by calling the REFERENCE's own host-compilable
helpers (oracle/_ref/libcrref_host.so, compiled by oracle/Makefile from the sources where they lie
under /root/reference: cuda/Util.hpp, cuda/PrivateDefs.hpp, base/Math.cpp).  TEST INFRASTRUCTURE.

Run in the build container (the reference tree does not exist on the GPU box):
    make -C oracle ref && python oracle/gen_golden_vectors.py
The committed fixture pins oracle/golden.hpp (tests/test_oracle_golden_vectors.py)."""
import ctypes
import json
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_host_vectors.json")


def f2u(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


def main():
    L = ctypes.CDLL(os.path.join(HERE, "_ref", "libcrref_host.so"))
    vp = ctypes.c_void_p
    for n in ("ref_select_flips", "ref_encode_depth", "ref_decode_depth", "ref_to_abgr", "ref_clear_depth", "ref_depth_min", "ref_depth_max"):
        getattr(L, n).restype = ctypes.c_uint32
    L.ref_encode_depth.argtypes = [ctypes.c_uint32]
    L.ref_decode_depth.argtypes = [ctypes.c_uint32]
    L.ref_to_abgr.argtypes = [ctypes.c_float] * 4
    L.ref_clear_depth.argtypes = [ctypes.c_float]
    L.ref_clip_triangle.argtypes = [vp, vp, vp, vp]
    rng = np.random.Generator(np.random.PCG64(0xC0DE))
    out = {"source": "reference host helpers: cuda/Util.hpp:182-298, base/Math.cpp:41-48, cuda/PrivateDefs.hpp, cuda/Constants.hpp:73-84"}

    out["constants"] = {"sizeof_header": L.ref_sizeof_header(), "sizeof_data": L.ref_sizeof_data(), "sizeof_params": L.ref_sizeof_params(),
                        "depth_min": L.ref_depth_min(), "depth_max": L.ref_depth_max(), "bary_max": L.ref_bary_max()}
    out["msaa_x"] = [[L.ref_msaa_x(s, i) for i in range(1 << s)] for s in range(4)]

    # cover8x8_selectFlips: every small direction + random large ones
    dirs = [(dx, dy) for dx in range(-3, 4) for dy in range(-3, 4)]
    dirs += [tuple(int(v) for v in rng.integers(-32767, 32768, 2)) for _ in range(400)]
    dirs += [(0, 5), (5, 0), (0, -5), (-5, 0), (7, 7), (-7, 7), (7, -7), (-7, -7), (32767, 1), (1, 32767), (-32767, -1)]
    out["select_flips"] = [[dx, dy, L.ref_select_flips(dx, dy)] for dx, dy in dirs]

    # selectMSAACentroid: all masks of all sample counts
    out["msaa_centroid"] = [[s, m, L.ref_msaa_centroid(s, m)] for s in range(4) for m in range(1 << (1 << s))]

    # encodeDepth / decodeDepth / clear depth
    ds = [0, 1, 2, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFE, 0xFFFFFFFF, 17600, 0xFFFFBB3F] + [int(v) for v in rng.integers(0, 2**32, 200)]
    out["encode_depth"] = [[d, L.ref_encode_depth(d)] for d in ds]
    out["decode_depth"] = [[d, L.ref_decode_depth(d)] for d in ds]
    fs = [0.0, 1.0, 0.5, 0.25, 0.999999, 1e-9, 2.0]  # negative depths are UB in the reference (double -> U64 cast) + [float(np.float32(v)) for v in rng.uniform(0, 1, 100)]
    out["clear_depth"] = [[f2u(np.float32(f)), L.ref_clear_depth(f)] for f in fs]

    # Vec4f::toABGR: edge values around the rounding points + randoms + out-of-range
    cs = [0.0, 1.0, 0.5, 0.2, 0.4, 0.8, 1.0 / 255, 0.5 / 255, 0.49999 / 255, 1.5 / 255, 254.5 / 255, 254.49 / 255, -0.1, 1.1, 2.0, 1e-8]
    cs += [float(np.float32(v)) for v in rng.uniform(-0.2, 1.2, 300)]
    vec = []
    for k in range(0, len(cs) - 3):
        r, g, b, a = cs[k], cs[k + 1], cs[k + 2], cs[k + 3]
        vec.append([f2u(np.float32(r)), f2u(np.float32(g)), f2u(np.float32(b)), f2u(np.float32(a)), L.ref_to_abgr(r, g, b, a)])
    out["to_abgr"] = vec

    # clipTriangleWithFrustum: random triangles crossing the frustum (inputs and outputs as F32 bit patterns).
    # NOTE: the reference helper is compiled by g++ for the HOST (no FMA contraction, -O2); the device
    # build contracts a*b+c into FMA.  The vertex COUNT and topology are compared exactly, barycentrics
    # to 1e-5 (tests/test_oracle_golden_vectors.py); bit-exactness of the device arithmetic is pinned on
    # the GPU against the reference kernels (tests/test_gpu_ref_kernels.py).
    clips = []
    for _ in range(300):
        c = rng.uniform(-1.2, 1.2, 3)
        p = c + rng.uniform(-1, 1, (3, 3)) * rng.choice([0.2, 1.0, 3.0])
        w = rng.uniform(0.3, 2.0, (3, 1))
        if rng.uniform() < 0.2:
            w[0, 0] = rng.uniform(-0.5, 0.05)
        v = np.concatenate([p * w, w], axis=1).astype(np.float32)
        bary = np.zeros(18, np.float32)
        n = L.ref_clip_triangle(v[0].ctypes.data, v[1].ctypes.data, v[2].ctypes.data, bary.ctypes.data)
        clips.append({"v": [int(x) for x in v.view(np.uint32).reshape(-1)], "n": n, "bary": [int(x) for x in bary[:2 * n].view(np.uint32)]})
    out["clip_triangle"] = clips

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
