"""Synthetic scenes for the BASELINE.json configurations (definitions: SURVEY.md 8d).

Every generator returns ``(vertices, indices)``: ``vertices`` is a float32 array
[numVerts, vertexStructSize/4] laid out as the pixel-pipe vertex struct (clipPos first, then one
Vec4f per varying; reference: cuda/PixelPipe.hpp:41-47) and ``indices`` an int32 array
[numTris, 3].  All data is synthetic and seeded (numpy PCG64)."""
import numpy as np


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


# ---- C1: the reference demo's cube (test/App.cpp:208-241) -----------------------------------------
CUBE_POSITIONS = np.array([[-1, -1, -1], [-1, -1, 1], [-1, 1, -1], [-1, 1, 1], [1, -1, -1], [1, -1, 1], [1, 1, -1], [1, 1, 1]], dtype=np.float32)
CUBE_INDICES = np.array([[7, 3, 1], [7, 1, 5], [7, 5, 6], [6, 5, 4], [6, 4, 2], [2, 4, 0], [2, 0, 3], [3, 0, 1], [3, 7, 6], [3, 6, 2], [5, 1, 0], [5, 0, 4]], dtype=np.int32)


def _perspective(fovy_deg, aspect, z_near, z_far):
    """glm::perspective of GLM 0.9.3 (fov in degrees; thirdparty/glm/gtc/matrix_transform.inl:223-244), FP32."""
    f32 = np.float32
    rng = f32(np.tan(np.radians(f32(fovy_deg) / f32(2)))) * f32(z_near)
    left, right, bottom, top = -rng * f32(aspect), rng * f32(aspect), -rng, rng
    m = np.zeros((4, 4), dtype=np.float32)  # m[col][row]
    m[0][0] = (f32(2) * f32(z_near)) / (right - left)
    m[1][1] = (f32(2) * f32(z_near)) / (top - bottom)
    m[2][2] = -(f32(z_far) + f32(z_near)) / (f32(z_far) - f32(z_near))
    m[2][3] = -f32(1)
    m[3][2] = -(f32(2) * f32(z_far) * f32(z_near)) / (f32(z_far) - f32(z_near))
    return m


def _look_at(eye, center, up):
    eye, center, up = [np.asarray(v, dtype=np.float32) for v in (eye, center, up)]
    f = center - eye
    f = f / np.float32(np.linalg.norm(f))
    u = up / np.float32(np.linalg.norm(up))
    s = np.cross(f, u)
    s = s / np.float32(np.linalg.norm(s))
    u = np.cross(s, f)
    m = np.eye(4, dtype=np.float32)  # m[col][row]
    m[0][0], m[1][0], m[2][0] = s
    m[0][1], m[1][1], m[2][1] = u
    m[0][2], m[1][2], m[2][2] = -f
    m[3][0], m[3][1], m[3][2] = -np.dot(s, eye), -np.dot(u, eye), np.dot(f, eye)
    return m


def cube_mvp(width=1024, height=768):
    """The demo camera's posToClip, column-major m[col][row] (test/App.cpp:83-103)."""
    eye, target = np.array([0, 2, 4], np.float32), np.zeros(3, np.float32)
    d = target - eye
    d = d / np.float32(np.linalg.norm(d))
    left = np.cross(np.array([0, 1, 0], np.float32), d)
    left = left / np.float32(np.linalg.norm(left))
    up = np.cross(d, left)
    up = up / np.float32(np.linalg.norm(up))
    proj = _perspective(60.0, np.float32(width) / np.float32(height), 0.1, 100.0)
    view = _look_at(eye, target, up)
    # column-major m[col][row]: (P*V)[c][r] = sum_k P[k][r] * V[c][k]
    mvp = np.zeros((4, 4), np.float32)
    for c in range(4):
        for r in range(4):
            acc = np.float32(0)
            for k in range(4):
                acc = np.float32(acc + proj[k][r] * view[c][k])
            mvp[c][r] = acc
    return mvp


def cube(width=1024, height=768):
    """BASELINE config 1: 12 triangles; MVP = perspective(60 deg, w/h, 0.1, 100) * lookAt((0,2,4),(0,0,0),up)
    (test/App.cpp:83-103); vertex shader = clipPos = MVP * (pos, 1) (test/shader/PassThrough.cu:31-34)."""
    mvp = cube_mvp(width, height)
    verts = np.zeros((8, 4), np.float32)
    for i, p in enumerate(CUBE_POSITIONS):
        for r in range(4):
            verts[i, r] = np.float32(np.float32(np.float32(mvp[0][r] * p[0] + mvp[1][r] * p[1]) + mvp[2][r] * p[2]) + mvp[3][r])
    return verts, CUBE_INDICES.copy()


# ---- grids (C2, C3, C5) -----------------------------------------------------------------------------
def _hash_colors(n, seed):
    v = (np.arange(n, dtype=np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
    v ^= v >> np.uint64(29)
    v *= np.uint64(0xBF58476D1CE4E5B9)
    v ^= v >> np.uint64(32)
    c = np.zeros((n, 4), np.float32)
    c[:, 0] = ((v >> np.uint64(0)) & np.uint64(0xFF)).astype(np.float32) / 255.0
    c[:, 1] = ((v >> np.uint64(8)) & np.uint64(0xFF)).astype(np.float32) / 255.0
    c[:, 2] = ((v >> np.uint64(16)) & np.uint64(0xFF)).astype(np.float32) / 255.0
    c[:, 3] = 1.0
    return c


def _grid_positions(nx, ny, rng, span=1.02, jitter=0.25, z_amp=0.2, z_ofs=0.0, perspective=0.5):
    """(nx+1)*(ny+1) clip-space positions: xy in [-span,span], jittered, z/w = z_amp*sin(7x)cos(5y)+z_ofs, w = 1+perspective*y."""
    gx = np.linspace(-span, span, nx + 1, dtype=np.float64)
    gy = np.linspace(-span, span, ny + 1, dtype=np.float64)
    x, y = np.meshgrid(gx, gy)
    x = x + rng.uniform(-jitter, jitter, x.shape) * (2 * span / nx)
    y = y + rng.uniform(-jitter, jitter, y.shape) * (2 * span / ny)
    z = z_amp * np.sin(7 * x) * np.cos(5 * y) + z_ofs
    w = 1.0 + perspective * y
    pos = np.stack([x * w, y * w, z * w, w], axis=-1).reshape(-1, 4).astype(np.float32)
    return pos


def _grid_indices(nx, ny, base=0):
    """Two CCW triangles per quad, row-major quads."""
    j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    v00 = (j * (nx + 1) + i).reshape(-1)
    v10, v01, v11 = v00 + 1, v00 + (nx + 1), v00 + (nx + 2)
    tris = np.empty((nx * ny * 2, 3), np.int32)
    tris[0::2] = np.stack([v00, v10, v11], axis=1)
    tris[1::2] = np.stack([v00, v11, v01], axis=1)
    return tris + np.int32(base)


def grid_gouraud(nx=1000, ny=500, seed=0xC0DE0002, **kw):
    """BASELINE config 2: nx x ny quads -> 2*nx*ny CCW triangles, GouraudVertex (32 B)."""
    rng = _rng(seed)
    pos = _grid_positions(nx, ny, rng, **kw)
    verts = np.concatenate([pos, _hash_colors(pos.shape[0], seed)], axis=1).astype(np.float32)
    return verts, _grid_indices(nx, ny)


def layered_phong(nx=1000, ny=500, layers=5, seed=0xC0DE0003):
    """BASELINE config 3: `layers` grids at different depths, submission order shuffled,
    ShadedVertex_texPhong (64 B: clipPos, cameraPos, cameraNormal, texCoord)."""
    rng = _rng(seed)
    order = rng.permutation(layers)
    all_v, all_i = [], []
    base = 0
    for layer in order:
        pos = _grid_positions(nx, ny, rng, z_amp=0.08, z_ofs=-0.6 + 1.2 * (layer + 0.5) / layers)
        n = pos.shape[0]
        ndc = pos[:, :3] / pos[:, 3:4]
        cam = np.concatenate([ndc[:, :2] * 2.0, (ndc[:, 2:3] * 2.0 - 5.0), np.ones((n, 1), np.float32)], axis=1)
        nrm = np.concatenate([0.3 * np.cos(9 * ndc[:, 0:1]), 0.3 * np.sin(11 * ndc[:, 1:2]), np.ones((n, 1)), np.zeros((n, 1))], axis=1)
        tex = np.concatenate([ndc[:, :2] * 0.5 + 0.5 + 0.1 * layer, np.zeros((n, 1)), np.ones((n, 1))], axis=1)
        all_v.append(np.concatenate([pos, cam, nrm, tex], axis=1).astype(np.float32))
        all_i.append(_grid_indices(nx, ny, base))
        base += n
    return np.concatenate(all_v, axis=0), np.concatenate(all_i, axis=0)


def subpixel_soup(num_tris=10_000_000, width=1920, height=1080, seed=0xC0DE0004):
    """BASELINE config 4: independent triangles with 0.1..0.9 px edges, unindexed, ShadedVertexBase (16 B)."""
    rng = _rng(seed)
    cx = rng.uniform(-1, 1, num_tris)
    cy = rng.uniform(-1, 1, num_tris)
    edge = rng.uniform(0.1, 0.9, num_tris)
    rot = rng.uniform(0, 2 * np.pi, num_tris)
    z = rng.uniform(-0.9, 0.9, num_tris)
    verts = np.empty((num_tris, 3, 4), np.float32)
    for k in range(3):
        a = rot + k * (2 * np.pi / 3)  # increasing angle = CCW
        r = edge / np.sqrt(3.0)
        verts[:, k, 0] = cx + r * np.cos(a) * (2.0 / width)
        verts[:, k, 1] = cy + r * np.sin(a) * (2.0 / height)
        verts[:, k, 2] = z
        verts[:, k, 3] = 1.0
    idx = np.arange(num_tris * 3, dtype=np.int32).reshape(-1, 3)
    return verts.reshape(-1, 4), idx


def random_soup(num_tris, seed, stride_floats=8, size=0.2, w_range=(0.6, 1.8), clip_fraction=0.15, behind_fraction=0.05):
    """Parity-test scene: random triangles of mixed size and winding, some crossing the frustum
    planes (x, y, near/far), some with w <= 0 vertices, plus random varyings."""
    rng = _rng(seed)
    c = rng.uniform(-1.1, 1.1, (num_tris, 1, 3))
    s = size * rng.uniform(0.02, 1.0, (num_tris, 1, 1)) ** 2
    p = c + rng.uniform(-1, 1, (num_tris, 3, 3)) * s
    big = rng.uniform(0, 1, num_tris) < clip_fraction
    p[big] = c[big] + rng.uniform(-1, 1, (int(big.sum()), 3, 3)) * 2.5
    w = rng.uniform(w_range[0], w_range[1], (num_tris, 3, 1))
    behind = rng.uniform(0, 1, num_tris) < behind_fraction
    w[behind, 0, 0] = rng.uniform(-0.5, 0.05, int(behind.sum()))
    pos = np.concatenate([p * w, w], axis=2).astype(np.float32)
    verts = np.zeros((num_tris * 3, stride_floats), np.float32)
    verts[:, :4] = pos.reshape(-1, 4)
    if stride_floats > 4:
        verts[:, 4:] = rng.uniform(0, 1, (num_tris * 3, stride_floats - 4)).astype(np.float32)
    idx = np.arange(num_tris * 3, dtype=np.int32).reshape(-1, 3)
    # share some vertices / reverse some windings
    flip = rng.uniform(0, 1, num_tris) < 0.3
    idx[flip] = idx[flip][:, ::-1]
    return verts, idx


def view_matrix_variants(num_views, seed=0xC0DE0006):
    """48 view transforms (6 cube faces x 8 positions) as 2x3 affine maps applied to NDC xy of the
    C2-style mesh (scale, rotation, offset) -- stand-in for per-view MVPs of config 5(ii)."""
    rng = _rng(seed)
    out = []
    for v in range(num_views):
        ang = (v % 6) * (np.pi / 3) + rng.uniform(-0.1, 0.1)
        sc = 0.7 + 0.05 * (v // 6)
        out.append((sc * np.cos(ang), -sc * np.sin(ang), rng.uniform(-0.1, 0.1), sc * np.sin(ang), sc * np.cos(ang), rng.uniform(-0.1, 0.1)))
    return np.array(out, np.float32)


def apply_view(verts, m):
    """Returns a copy of `verts` with clip xy transformed by the affine map m (NDC space, w preserved)."""
    out = verts.copy()
    x, y, w = verts[:, 0], verts[:, 1], verts[:, 3]
    out[:, 0] = m[0] * x + m[1] * y + m[2] * w
    out[:, 1] = m[3] * x + m[4] * y + m[5] * w
    return out


def grid_gouraud_view(nx=1000, ny=500, view=0, num_views=48, **kw):
    """BASELINE config 5(ii): view `view` of the 48 (6 cube faces x 8 positions) views of the C2 mesh."""
    v, i = grid_gouraud(nx, ny, **kw)
    return apply_view(v, view_matrix_variants(num_views)[view]), i


def grid_gouraud_4k(nx=2000, ny=1000, seed=0xC0DE0005):
    """BASELINE config 5(i): C2-style 4 M-triangle grid for the 3840x2160 sort-first frame."""
    return grid_gouraud(nx, ny, seed=seed)
