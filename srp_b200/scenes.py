"""Synthetic scenes and the BASELINE configs, described once and replayed
through any implementation of the srp C API (srp_b200.host.SrpLibrary).

A Scene is data: framebuffer size, textures, and a list of draws, each with the context
state to set, the built-in program to use, its uniform bytes and its buffers.  `render`
issues exactly the call sequence a C program would: srpNewContext, state setters,
srp*BufferCopyData, srpFramebufferClear, srpDraw*Buffer.

Matrices are built with numpy float32/float64 in the operation order of the library's own
constructors (reference src/math/mat.c:77-215); both implementations get the same bytes,
so the comparison isolates the draw path.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from pathlib import Path
from typing import Callable

import numpy as np

from . import host as H

ROOT = Path(__file__).resolve().parent.parent
f32 = np.float32


# ------------------------------------------------------------------------------------ math
def _mm(a, b):
    """mat4MultiplyMat4: every element ((p0+p1)+p2)+p3 with float32 rounding at each step"""
    a = a.astype(f32); b = b.astype(f32)
    r = np.zeros((4, 4), f32)
    for i in range(4):
        for j in range(4):
            acc = f32(a[i, 0] * b[0, j])
            for k in (1, 2, 3):
                acc = f32(acc + f32(a[i, k] * b[k, j]))
            r[i, j] = acc
    return r


def rotate(x, y, z):
    x, y, z = (float(f32(v)) for v in (x, y, z))
    sx, cx, sy, cy, sz, cz = np.sin(x), np.cos(x), np.sin(y), np.cos(y), np.sin(z), np.cos(z)
    r = np.zeros((4, 4), np.float64)
    r[0, 0] = cy * cz; r[0, 1] = sx * sy * cz - cx * sz; r[0, 2] = cx * sy * cz + sx * sz
    r[1, 0] = cy * sz; r[1, 1] = sx * sy * sz + cx * cz; r[1, 2] = cx * sy * sz - sx * cz
    r[2, 0] = -sy; r[2, 1] = sx * cy; r[2, 2] = cx * cy
    r[3, 3] = 1
    return r.astype(f32)


def translate(x, y, z):
    r = np.eye(4, dtype=f32); r[0, 3] = x; r[1, 3] = y; r[2, 3] = z
    return r


def scale(x, y, z):
    return np.diag(np.array([x, y, z, 1], f32))


def trs(t, r, s):
    return _mm(translate(*t), _mm(rotate(*r), scale(*s)))


def view(camera, rotation=(0, 0, 0), zoom=(1, 1, 1)):
    return trs([-c for c in camera], [-a for a in rotation], zoom)


def orthogonal(x0, x1, y0, y1, z0, z1):
    x0, x1, y0, y1, z0, z1 = (f32(v) for v in (x0, x1, y0, y1, z0, z1))
    r = np.zeros((4, 4), f32)
    r[0, 0] = f32(2) / (x1 - x0); r[0, 3] = -(x1 + x0) / (x1 - x0)
    r[1, 1] = f32(2) / (y1 - y0); r[1, 3] = -(y1 + y0) / (y1 - y0)
    r[2, 2] = f32(2) / (z1 - z0); r[2, 3] = -(z1 + z0) / (z1 - z0)
    r[3, 3] = 1
    return r


def perspective(x0, x1, y0, y1, near, far):
    near, far = f32(near), f32(far)
    p = np.zeros((4, 4), f32)
    p[0, 0] = near; p[1, 1] = near; p[2, 2] = near + far; p[2, 3] = -near * far; p[3, 2] = 1
    return _mm(orthogonal(x0, x1, y0, y1, near, far), p)


def transform_bytes(model, view_m, proj):
    return model.astype(f32).tobytes() + view_m.astype(f32).tobytes() + proj.astype(f32).tobytes()


# ------------------------------------------------------------------------------------ scene description
@dataclass
class Draw:
    program: str
    primitive: int
    vertices: np.ndarray            # any dtype; raw bytes are uploaded
    stride: int
    uniform: bytes | None | Callable = None      # bytes, None, or f(lib, resources) -> bytes
    indices: np.ndarray | None = None
    start: int = 0
    count: int | None = None
    varyings: list = field(default_factory=list)  # [(nItems, SRPType, interpolation mode)]
    varyings_size: int = 0
    may_overwrite_depth: bool = False
    state: list = field(default_factory=list)     # [("srpDepthTest", True), ...] applied before the draw
    clear_before: bool = False                    # srpFramebufferClear right before this draw


@dataclass
class Scene:
    name: str
    width: int
    height: int
    draws: list
    textures: dict = field(default_factory=dict)  # name -> (rgb uint8 [H,W,3], wrap_x, wrap_y)
    clear: bool = True                            # srpFramebufferClear before the first draw


class Prepared:
    """A scene bound to one library: buffers uploaded, programs and uniforms built."""

    def __init__(self, lib: H.SrpLibrary, scene: Scene):
        self.lib, self.scene = lib, scene
        lib.new_context()
        self.resources = {k: lib.texture_from_memory(*v) for k, v in scene.textures.items()}
        self.fb = lib.framebuffer(scene.width, scene.height)
        self.items = []
        cache = {}
        for d in scene.draws:
            key = (id(d.vertices), id(d.indices))
            if key not in cache:
                vb = lib.vertex_buffer(d.vertices, d.stride)
                ib = lib.index_buffer(d.indices) if d.indices is not None else None
                cache[key] = (vb, ib)
            vb, ib = cache[key]
            prog = lib.program(d.program, d.varyings, d.varyings_size, d.may_overwrite_depth)
            uni = d.uniform(lib, self.resources) if callable(d.uniform) else d.uniform
            prog.set_uniform(uni)
            count = d.count
            if count is None:
                count = len(d.indices) if d.indices is not None else (d.vertices.nbytes // d.stride)
            self.items.append((d, prog, vb, ib, count))
        self._buffers = cache

    def upload(self):
        """re-issue srp*BufferCopyData for every buffer (the end-to-end step includes it)"""
        seen = set()
        for d, _, vb, ib, _ in self.items:
            if id(d.vertices) in seen:
                continue
            seen.add(id(d.vertices))
            self.lib.vertex_buffer_copy(vb, d.vertices, d.stride)
            if ib is not None:
                self.lib.index_buffer_copy(ib, d.indices)

    def draw_all(self):
        lib = self.lib
        if self.scene.clear:
            self.fb.clear()
        for d, prog, vb, ib, count in self.items:
            for fn, *args in d.state:
                getattr(lib.dll, fn)(*args)
            if d.clear_before:
                self.fb.clear()
            lib.draw(self.fb, prog, d.primitive, d.start, count, vb, ib)

    def planes(self):
        return self.fb.planes()

    def free(self):
        for vb, ib in self._buffers.values():
            self.lib.dll.srpFreeVertexBuffer(vb)
            if ib is not None:
                self.lib.dll.srpFreeIndexBuffer(ib)
        for t in self.resources.values():
            self.lib.dll.srpFreeTexture(t)
        self.fb.free()


def render(lib: H.SrpLibrary, scene: Scene):
    p = Prepared(lib, scene)
    try:
        p.draw_all()
        return p.planes()
    finally:
        p.free()


# ------------------------------------------------------------------------------------ meshes
def cube_mesh():
    """24 vertices {vec3 position, vec2 uv} and 36 u8 indices: an axis-aligned cube of
    half-size 1, CCW faces seen from outside (the geometry class of cfg1)."""
    faces = [  # (origin, u axis, v axis) -> 4 corners (0,0) (1,0) (1,1) (0,1)
        ((-1, -1, -1), (2, 0, 0), (0, 2, 0)),   # back  (z = -1)
        ((-1, 1, -1), (2, 0, 0), (0, 0, 2)),    # top
        ((-1, 1, 1), (2, 0, 0), (0, -2, 0)),    # front (z = +1)
        ((-1, -1, 1), (2, 0, 0), (0, 0, -2)),   # bottom
        ((1, -1, -1), (0, 0, 2), (0, 2, 0)),    # right
        ((-1, -1, 1), (0, 0, -2), (0, 2, 0)),   # left
    ]
    verts, idx = [], []
    for o, u, v in faces:
        base = len(verts)
        for (a, b) in ((0, 0), (1, 0), (1, 1), (0, 1)):
            p = [o[k] + a * u[k] + b * v[k] for k in range(3)]
            verts.append(p + [a, b])
        idx += [base, base + 1, base + 2, base, base + 2, base + 3]
    return np.array(verts, f32), np.array(idx, np.uint8)


ASSETS = ROOT / "tests" / "_build" / "res"      # the reference's examples/res, copied there by srp_b200.build


def asset_path(rel):
    """A resource of the reference's examples (teapot OBJ, stone-wall texture).  The build copies
    examples/res next to the scene executables; the snapshot that travels to the GPU box carries
    it.  A missing asset is an error: no config is ever rendered with a stand-in."""
    p = ASSETS / rel
    if not p.exists():
        raise FileNotFoundError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference exists")
    return p


def stone_wall_texture():
    """examples/res/textures/stoneWall.png (480x480) as RGB8, the texture of cfg1"""
    from PIL import Image
    with Image.open(asset_path("textures/stoneWall.png")) as im:
        return np.ascontiguousarray(np.asarray(im.convert("RGB"), dtype=np.uint8))


def load_obj(path):
    """de-indexing OBJ reader with the semantics of the reference's
    examples/utility/objparser.c:8-79: every face corner becomes its own vertex
    {position, uv, normal}, indices are 0..n-1."""
    pos, uvs, nrm, out = [], [], [], []
    for line in Path(path).read_text().splitlines():
        if line.startswith("v "):
            pos.append([float(t) for t in line.split()[1:4]])
        elif line.startswith("vt"):
            uvs.append([float(t) for t in line.split()[1:3]])
        elif line.startswith("vn"):
            nrm.append([float(t) for t in line.split()[1:4]])
        elif line.startswith("f"):
            corners = line.split()[1:4]
            for c in corners:
                vi, ti, ni = (int(t) for t in c.split("/"))
                out.append(pos[vi - 1] + uvs[ti - 1] + nrm[ni - 1])
    verts = np.array(out, f32)
    return verts, np.arange(len(verts), dtype=np.uint32)


def teapot_mesh():
    """(vertices [n, 8] float32, indices u32, source): the Utah teapot of the reference's examples/res"""
    v, i = load_obj(asset_path("objects/utah_teapot.obj"))
    return v, i, "utah_teapot.obj"


def sphere_shell(n=708, radius=3.0):
    """n x n quad grid on a sphere shell seen from inside: 2*n*n triangles (n = 708 ->
    1 002 528), (n+1)^2 vertices x 32 B, u32 indices, normals pointing inwards."""
    u = np.linspace(0.0, 2.0 * np.pi, n + 1)
    v = np.linspace(0.02 * np.pi, 0.98 * np.pi, n + 1)
    uu, vv = np.meshgrid(u, v)
    d = np.stack([np.sin(vv) * np.cos(uu), np.cos(vv), np.sin(vv) * np.sin(uu)], -1)
    verts = np.concatenate([radius * d, np.stack([uu / (2 * np.pi), vv / np.pi], -1), -d], -1).reshape(-1, 8).astype(f32)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    a = (i * (n + 1) + j).reshape(-1)
    quads = np.stack([a, a + 1, a + n + 2, a, a + n + 2, a + n + 1], -1)
    return verts, quads.reshape(-1).astype(np.uint32)


def lcg(seed, count):
    """numerical-recipes LCG, vectorised: returns `count` uint32 values"""
    out = np.empty(count, np.uint32)
    a, c = np.uint64(1664525), np.uint64(1013904223)
    state = np.uint64(seed)
    # block-wise to stay vectorised: x_{k+1} = a x_k + c mod 2^32
    mask = np.uint64(0xFFFFFFFF)
    block = 1 << 16
    mult = np.empty(block, np.uint64); add = np.empty(block, np.uint64)
    m, ad = np.uint64(1), np.uint64(0)
    for k in range(block):
        m = (m * a) & mask; ad = (ad * a + c) & mask
        mult[k] = m; add[k] = ad
    done = 0
    while done < count:
        n = min(block, count - done)
        vals = (mult[:n] * state + add[:n]) & mask
        out[done:done + n] = vals.astype(np.uint32)
        state = vals[n - 1]
        done += n
    return out


def jitter_grid(n=2237, seed=12345):
    """(n+1)^2 tagged vertices on a jittered NDC grid, 2*n*n sub-pixel triangles at 4K
    (n = 2237 -> 10 008 338).  Vertex = {vec3 position, u32 tag} (16 B), z uniform in (-1, 1)."""
    m = n + 1
    rnd = lcg(seed, 3 * m * m).astype(np.float64) / 4294967296.0
    gy, gx = np.mgrid[0:m, 0:m]
    cell = 2.0 / n
    x = -1.0 + gx * cell + (rnd[0::3].reshape(m, m) - 0.5) * cell * 0.8
    y = -1.0 + gy * cell + (rnd[1::3].reshape(m, m) - 0.5) * cell * 0.8
    z = (rnd[2::3].reshape(m, m) * 2.0 - 1.0) * 0.999
    verts = np.zeros((m * m, 4), f32)
    verts[:, 0] = np.clip(x, -1, 1).reshape(-1); verts[:, 1] = np.clip(y, -1, 1).reshape(-1); verts[:, 2] = z.reshape(-1)
    tags = (np.arange(m * m, dtype=np.uint32) * np.uint32(2654435761)) >> np.uint32(24)
    verts.view(np.uint32)[:, 3] = tags
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    a = (i * m + j).reshape(-1)
    quads = np.stack([a, a + 1, a + m + 1, a, a + m + 1, a + m], -1)
    return verts, quads.reshape(-1).astype(np.uint32)


# ------------------------------------------------------------------------------------ uniforms
GOURAUD_VARYINGS = [(3, H.SRP_FLOAT, H.SRP_INTERPOLATION_MODE_PERSPECTIVE)]


def gouraud_uniform(model, view_m, proj):
    material_ambient = material_diffuse = (1.0, 0.5, 0.31)
    light_ambient, light_diffuse, light_dir = (0.1, 0.1, 0.1), (0.5, 0.5, 0.5), (-1.0, -1.0, 1.0)
    tail = struct.pack("<15f", *material_ambient, *material_diffuse, *light_ambient, *light_diffuse, *light_dir)
    return transform_bytes(model, view_m, proj) + tail


def texcube_uniform(model, view_m, proj):
    def build(lib, resources):
        return transform_bytes(model, view_m, proj) + struct.pack("<Q", resources["wall"] or 0)
    return build


# ------------------------------------------------------------------------------------ BASELINE configs
def cfg1_textured_cube(width=800, height=600, frame=70, texture=None):
    """cfg1: textured cube, cull BACK / front CCW, depth test, perspective-correct uv"""
    verts, idx = cube_mesh()
    tex = texture if texture is not None else stone_wall_texture()
    model = rotate(frame / 100.0, frame / 200.0, frame / 500.0)
    d = Draw("texcube", H.SRP_PRIM_TRIANGLES, verts, 20, indices=idx,
             uniform=texcube_uniform(model, view((0, 0, -3)), perspective(-1, 1, -1, 1, 1, 50)),
             varyings=[(2, H.SRP_FLOAT, H.SRP_INTERPOLATION_MODE_PERSPECTIVE)], varyings_size=8,
             state=[("srpRasterFrontFace", H.SRP_WINDING_CCW), ("srpRasterCullFace", H.SRP_FACE_BACK), ("srpDepthTest", True)])
    return Scene("cfg1_textured_cube", width, height, [d], textures={"wall": (tex, H.TW_REPEAT, H.TW_REPEAT)})


def teapot_draw(frame=0, mesh=None):
    verts, idx, _ = mesh if mesh is not None else teapot_mesh()
    model = rotate(f32(-90.0 / 180.0 * np.pi), frame / 200.0, 0)   # RAD(-90), examples/utility/rad.h
    uni = gouraud_uniform(model, view((0, 1.75, -7)), perspective(-1, 1, -1, 1, 1, 10))
    return Draw("gouraud", H.SRP_PRIM_TRIANGLES, verts, 32, indices=idx, uniform=uni,
                varyings=GOURAUD_VARYINGS, varyings_size=12,
                state=[("srpRasterFrontFace", H.SRP_WINDING_CW), ("srpRasterCullFace", H.SRP_FACE_BACK), ("srpDepthTest", True)])


def cfg2_teapot(width=1920, height=1080, frame=40, mesh=None):
    """cfg2: Utah teapot, Gouraud shading, depth test + back-face culling"""
    return Scene("cfg2_teapot", width, height, [teapot_draw(frame, mesh)])


def cfg3_shell(width=3840, height=2160, n=708, radius=3.0):
    """cfg3: ~1M-triangle sphere shell, camera at the origin inside the mesh"""
    verts, idx = sphere_shell(n, radius)
    model = rotate(0.3, 0.2, 0.1)
    uni = gouraud_uniform(model, view((0, 0, 0)), perspective(-1, 1, -1, 1, 1, 10))
    d = Draw("gouraud", H.SRP_PRIM_TRIANGLES, verts, 32, indices=idx, uniform=uni,
             varyings=GOURAUD_VARYINGS, varyings_size=12,
             state=[("srpRasterCullFace", H.SRP_FACE_NONE), ("srpDepthTest", True)])
    return Scene(f"cfg3_shell_n{n}_r{radius}", width, height, [d])


def cfg4_subpixel(width=3840, height=2160, n=2237, n_lines=1_000_000, n_points=1_000_000, seed=12345):
    """cfg4: sub-pixel triangles + lines + points with stencil, scissor and depth"""
    verts, idx = jitter_grid(n, seed)
    ident = np.eye(4, dtype=f32)
    xf = transform_bytes(ident, ident, ident)
    tag_vary = [(1, H.SRP_UINT8, H.SRP_INTERPOLATION_MODE_FLAT)]
    sc = (width // 4, height // 4, width // 2, height // 2)
    common = [("srpScissorTest", True), ("srpScissorOptions", *sc), ("srpDepthTest", True),
              ("srpDepthCompareOp", H.SRP_COMPARE_GREATER), ("srpStencilTest", True)]
    draws = [Draw("tagged", H.SRP_PRIM_TRIANGLES, verts, 16, indices=idx, uniform=xf, varyings=tag_vary, varyings_size=1,
                  state=common + [("srpStencilFunc", H.SRP_COMPARE_ALWAYS, 0, 0xFF),
                                  ("srpStencilOp", H.SRP_STENCIL_KEEP, H.SRP_STENCIL_KEEP, H.SRP_STENCIL_INCR_WRAP)])]
    rnd = lcg(seed + 1, 2 * n_lines + n_points)
    nv = len(verts)
    if n_lines:
        a = (rnd[:n_lines] % np.uint32(nv)).astype(np.int64)
        off = (rnd[n_lines:2 * n_lines] % np.uint32(9)).astype(np.int64) - 4
        b = np.clip(a + off * (n + 1) + (off * 3) % 7 - 3, 0, nv - 1)
        lidx = np.stack([a, b], -1).reshape(-1).astype(np.uint32)
        draws.append(Draw("tagged", H.SRP_PRIM_LINES, verts, 16, indices=lidx, uniform=xf, varyings=tag_vary, varyings_size=1))
    if n_points:
        pidx = (rnd[2 * n_lines:] % np.uint32(nv)).astype(np.uint32)
        half = n_points // 2
        draws.append(Draw("tagged", H.SRP_PRIM_POINTS, verts, 16, indices=pidx[:half], uniform=xf, varyings=tag_vary,
                          varyings_size=1, state=[("srpRasterPointSize", 1.0)]))
        draws.append(Draw("tagged", H.SRP_PRIM_POINTS, verts, 16, indices=pidx[half:], uniform=xf, varyings=tag_vary,
                          varyings_size=1, state=[("srpRasterPointSize", 3.0)]))
    # second pass over the triangles: only where the stencil counter equals 1
    draws.append(Draw("tagged", H.SRP_PRIM_TRIANGLES, verts, 16, indices=idx, uniform=xf, varyings=tag_vary, varyings_size=1,
                      state=[("srpStencilFunc", H.SRP_COMPARE_EQUAL, 1, 0xFF),
                             ("srpStencilOp", H.SRP_STENCIL_KEEP, H.SRP_STENCIL_KEEP, H.SRP_STENCIL_KEEP),
                             ("srpDepthCompareOp", H.SRP_COMPARE_GEQUAL)]))
    return Scene(f"cfg4_subpixel_n{n}", width, height, draws)


def cfg5_frame(frame, size=1024, mesh=None):
    """cfg5: one frame of the 1024-frame teapot batch"""
    return Scene(f"cfg5_frame{frame}", size, size, [teapot_draw(frame, mesh)])
