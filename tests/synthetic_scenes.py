"""Small synthetic scenes for the differential parity tests.

The reference's own suite pins 18 scenes at 512x512 (tests/golden/scenes).  SURVEY.md
section 4 lists what those do NOT pin; every item there has a scene here: index types
u8/u16/u32/u64 and startIndex != 0, strips / fans / loops with culling and clipping,
double / integer varyings, every depth compare op and depth-write off, stencil ops / masks /
separate face state, the late depth-test path (fragment shader writes depth), CLAMP texture
wrap, point sizes and points straddling the edge, odd and non-square resolutions, cull
FRONT, CW front faces, triangles crossing several clip planes, degenerate triangles, vertex
counts that are not a multiple of the primitive size, polygon modes on clipped geometry.

`all_scenes()` returns {name: Scene}; oracle/gen_synthetic_golden.py renders them with the
reference into tests/golden/synthetic/*.npz, the GPU tests render them with the product.
Everything is deterministic (numpy RandomState with fixed seeds).
"""
from __future__ import annotations

import numpy as np

from srp_b200 import host as H
from srp_b200 import scenes as S

f32 = np.float32
IDENT = np.eye(4, dtype=f32)
COLOR_VARY = [(3, H.SRP_FLOAT, H.SRP_INTERPOLATION_MODE_PERSPECTIVE)]


def procedural_texture(size=480, seed=7):
    """deterministic RGB8 test texture (bricks + noise)"""
    y, x = np.mgrid[0:size, 0:size]
    rng = np.random.RandomState(seed)
    noise = rng.randint(0, 48, (size, size))
    brick = (((x // 60 + (y // 30) % 2 * 30 // 30) % 2) * 40 + ((y % 30) < 2) * 60 + ((x + (y // 30 % 2) * 30) % 60 < 2) * 60)
    r = np.clip(96 + brick + noise, 0, 255)
    g = np.clip(80 + brick // 2 + noise, 0, 255)
    b = np.clip(64 + noise * 2, 0, 255)
    return np.stack([r, g, b], -1).astype(np.uint8)


def _xf(model=IDENT, view=IDENT, proj=IDENT):
    return S.transform_bytes(model, view, proj)


def _persp_xf(model=IDENT, cam=(0, 0, -3), near=1, far=20):
    return _xf(model, S.view(cam), S.perspective(-1, 1, -1, 1, near, far))


def random_color_tris(n, seed, spread=1.6, z_range=(-0.9, 0.9), size=0.5):
    """n random triangles {vec3 position, vec3 color} in NDC-ish coordinates"""
    rng = np.random.RandomState(seed)
    c = rng.uniform(-spread, spread, (n, 1, 2))
    xy = c + rng.uniform(-size, size, (n, 3, 2))
    z = rng.uniform(*z_range, (n, 3, 1))
    col = rng.uniform(0, 1, (n, 3, 3))
    return np.concatenate([xy, z, col], -1).reshape(-1, 6).astype(f32)


def clip_safe(verts6, model, cam, near, far, max_planes=3):
    """Drop the triangles of a {position, colour} list that straddle more than `max_planes`
    frustum planes.  The reference sizes its clip buffers for 6 vertices / 4 triangles
    (src/pipeline/clipping.c:83-84, primitive_assembly.c:59-60) and corrupts memory beyond
    that, so parity is only defined for such inputs."""
    m = (S.perspective(-1, 1, -1, 1, near, far).astype(np.float64) @ S.view(cam).astype(np.float64)
         @ model.astype(np.float64))
    p = np.concatenate([verts6[:, :3].astype(np.float64), np.ones((len(verts6), 1))], -1) @ m.T
    x, y, z, w = p[:, 0], p[:, 1], p[:, 2], p[:, 3]
    dist = np.stack([x + w, w - x, y + w, w - y, z + w, w - z], -1).reshape(-1, 3, 6)
    margin = 1e-4 * np.abs(dist).max()
    straddle = ((dist < margin).any(1) & (dist > -margin).any(1)).sum(-1)
    keep = np.repeat(straddle <= max_planes, 3)
    return verts6[keep]


def grid_mesh(nx, ny, z_fn=None, extent=1.2):
    """(nx+1)x(ny+1) colour-vertex grid in the xy plane, indexed triangles"""
    gx, gy = np.meshgrid(np.linspace(-extent, extent, nx + 1), np.linspace(-extent, extent, ny + 1))
    z = z_fn(gx, gy) if z_fn else np.zeros_like(gx)
    col = np.stack([(gx + extent) / (2 * extent), (gy + extent) / (2 * extent), 0.5 + 0.4 * np.sin(3 * gx)], -1)
    verts = np.concatenate([np.stack([gx, gy, z], -1), col], -1).reshape(-1, 6).astype(f32)
    i, j = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    a = (i * (nx + 1) + j).reshape(-1)
    idx = np.stack([a, a + 1, a + nx + 2, a, a + nx + 2, a + nx + 1], -1).reshape(-1)
    return verts, idx


def vcolor(verts, prim=H.SRP_PRIM_TRIANGLES, uniform=None, indices=None, state=(), mode=None, **kw):
    vary = COLOR_VARY if mode is None else [(3, H.SRP_FLOAT, mode)]
    return S.Draw("vcolor", prim, verts, 24, uniform=uniform if uniform is not None else _xf(), indices=indices,
                  varyings=vary, varyings_size=12, state=list(state), **kw)


def all_scenes() -> dict:
    out = {}

    def add(scene):
        assert scene.name not in out
        out[scene.name] = scene

    tris = random_color_tris(60, 1)
    model = S.rotate(0.4, 0.7, 0.2)

    # ---- interpolation modes at odd resolutions, no depth test (painter's order) ----
    for mode, tag in ((H.SRP_INTERPOLATION_MODE_PERSPECTIVE, "persp"), (H.SRP_INTERPOLATION_MODE_AFFINE, "affine"),
                      (H.SRP_INTERPOLATION_MODE_FLAT, "flat")):
        add(S.Scene(f"interp_{tag}_333x217", 333, 217,
                    [vcolor(tris, uniform=_persp_xf(model), mode=mode)]))
    add(S.Scene("flat_provoking_first_97x61", 97, 61,
                [vcolor(tris, uniform=_persp_xf(model), mode=H.SRP_INTERPOLATION_MODE_FLAT,
                        state=[("srpProvokingVertexMode", H.SRP_PROVOKING_VERTEX_FIRST)])]))

    # ---- every depth compare op, with and without depth writes ----
    for op, tag in enumerate(("never", "always", "less", "lequal", "greater", "gequal", "equal", "notequal")):
        st = [("srpDepthTest", True), ("srpDepthCompareOp", op)]
        draws = [vcolor(random_color_tris(40, 2), state=[("srpDepthTest", True), ("srpDepthCompareOp", H.SRP_COMPARE_ALWAYS)]),
                 vcolor(random_color_tris(40, 3), state=st),
                 vcolor(random_color_tris(40, 2), state=st)]          # same geometry again: EQUAL / NOTEQUAL matter
        add(S.Scene(f"depth_{tag}_256x128", 256, 128, draws))
    add(S.Scene("depth_write_off_200x200", 200, 200, [
        vcolor(random_color_tris(30, 4), state=[("srpDepthTest", True)]),
        vcolor(random_color_tris(30, 5), state=[("srpDepthWrite", False)]),
        vcolor(random_color_tris(30, 6), state=[("srpDepthWrite", True)])]))

    # ---- culling / winding on lists, strips and fans ----
    strip = grid_mesh(1, 1)[0]
    rng = np.random.RandomState(7)
    wavy = np.zeros((40, 6), f32)
    wavy[:, 0] = np.linspace(-1.1, 1.1, 40); wavy[:, 1] = np.where(np.arange(40) % 2, 0.5, -0.5) + rng.uniform(-0.2, 0.2, 40)
    wavy[:, 2] = rng.uniform(-0.5, 0.5, 40); wavy[:, 3:] = rng.uniform(0, 1, (40, 3))
    fan = np.zeros((24, 6), f32)
    ang = np.linspace(0, 2 * np.pi, 23)
    fan[1:, 0] = 0.9 * np.cos(ang); fan[1:, 1] = 0.9 * np.sin(ang) * (1 + 0.3 * np.sin(5 * ang)); fan[:, 3:] = rng.uniform(0, 1, (24, 3))
    for cull, ctag in ((H.SRP_FACE_NONE, "none"), (H.SRP_FACE_FRONT, "front"), (H.SRP_FACE_BACK, "back")):
        for wind, wtag in ((H.SRP_WINDING_CCW, "ccw"), (H.SRP_WINDING_CW, "cw")):
            st = [("srpRasterCullFace", cull), ("srpRasterFrontFace", wind)]
            add(S.Scene(f"cull_{ctag}_{wtag}_160x120", 160, 120, [
                vcolor(tris, uniform=_persp_xf(model), state=st),
                vcolor(wavy, prim=H.SRP_PRIM_TRIANGLE_STRIP, uniform=_xf(S.scale(0.8, 0.8, 1))),
                vcolor(fan, prim=H.SRP_PRIM_TRIANGLE_FAN, uniform=_xf(S.trs((0.2, -0.1, 0), (0, 0, 0.3), (0.6, 0.6, 1))))]))
    add(S.Scene("cull_front_and_back_64x64", 64, 64, [
        vcolor(tris, state=[("srpRasterCullFace", H.SRP_FACE_FRONT_AND_BACK)]),
        vcolor(wavy, prim=H.SRP_PRIM_LINE_STRIP)]))      # lines are not affected by culling

    # ---- clipping: camera inside a field of triangles, several planes at once ----
    clip_cam, clip_near, clip_far = (0, 0, -0.5), 0.3, 6
    clip_model = S.rotate(0.3, 0.5, 0.1)
    big = clip_safe(random_color_tris(160, 8, spread=2.5, z_range=(-3, 3), size=1.4), clip_model, clip_cam, clip_near, clip_far)
    big_ident = clip_safe(big, IDENT, clip_cam, clip_near, clip_far)
    big_mixed = clip_safe(big, S.rotate(0.2, 0.1, 0.0), (0, 0, -0.8), 0.4, 8)
    for pm, tag in ((H.SRP_POLYGON_MODE_FILL, "fill"), (H.SRP_POLYGON_MODE_LINE, "line"), (H.SRP_POLYGON_MODE_POINT, "point")):
        add(S.Scene(f"clip_multi_plane_{tag}_320x200", 320, 200, [
            vcolor(big, uniform=_persp_xf(clip_model, cam=clip_cam, near=clip_near, far=clip_far),
                   state=[("srpDepthTest", True), ("srpRasterPolygonMode", pm), ("srpRasterPointSize", 3.0)])]))
    add(S.Scene("clip_lines_and_points_211x157", 211, 157, [
        S.Draw("primid", H.SRP_PRIM_LINES, big_ident, 24, uniform=_persp_xf(IDENT, cam=(0, 0, -0.5), near=0.3, far=6)),
        S.Draw("primid", H.SRP_PRIM_LINE_LOOP, big_ident[:31], 24, uniform=_persp_xf(IDENT, cam=(0, 0.3, -0.7), near=0.3, far=6)),
        S.Draw("primid", H.SRP_PRIM_POINTS, big_ident, 24, uniform=_persp_xf(IDENT, cam=(0, 0, -0.5), near=0.3, far=6),
               state=[("srpRasterPointSize", 2.5)])]))
    # flat and integer varyings through the clipper (clip vertices blend FLAT floats, App. B-15)
    mixed_vary = [(1, H.SRP_DOUBLE, H.SRP_INTERPOLATION_MODE_PERSPECTIVE), (3, H.SRP_FLOAT, H.SRP_INTERPOLATION_MODE_FLAT),
                  (2, H.SRP_FLOAT, H.SRP_INTERPOLATION_MODE_AFFINE), (1, H.SRP_INT32, H.SRP_INTERPOLATION_MODE_FLAT),
                  (2, H.SRP_UINT16, H.SRP_INTERPOLATION_MODE_FLAT)]
    for prov, tag in ((H.SRP_PROVOKING_VERTEX_LAST, "last"), (H.SRP_PROVOKING_VERTEX_FIRST, "first")):
        add(S.Scene(f"mixed_varyings_clip_{tag}_300x180", 300, 180, [
            S.Draw("mixed", H.SRP_PRIM_TRIANGLES, big_mixed, 24, varyings=mixed_vary, varyings_size=40,
                   uniform=_persp_xf(S.rotate(0.2, 0.1, 0.0), cam=(0, 0, -0.8), near=0.4, far=8),
                   state=[("srpDepthTest", True), ("srpProvokingVertexMode", prov)]),
            S.Draw("mixed", H.SRP_PRIM_LINE_STRIP, big_mixed[:50], 24, varyings=mixed_vary, varyings_size=40,
                   uniform=_persp_xf(IDENT, cam=(0, 0, -0.8), near=0.4, far=8))]))

    # ---- index buffers: every index type, startIndex != 0, counts that are not multiples ----
    gverts, gidx = grid_mesh(12, 9, z_fn=lambda x, y: 0.3 * np.sin(2 * x) * np.cos(2 * y))
    for dt, tag in ((np.uint8, "u8"), (np.uint16, "u16"), (np.uint32, "u32"), (np.uint64, "u64")):
        idx = gidx.astype(dt)
        add(S.Scene(f"index_{tag}_start_offset_192x144", 192, 144, [
            vcolor(gverts, indices=idx, uniform=_persp_xf(S.rotate(0.9, 0.1, 0.0)), start=6, count=len(idx) - 6 - 2,
                   state=[("srpDepthTest", True)]),
            S.Draw("primid", H.SRP_PRIM_LINES, gverts, 24, indices=idx, uniform=_persp_xf(S.rotate(0.9, 0.1, 0.0)), start=3, count=101),
            S.Draw("primid", H.SRP_PRIM_POINTS, gverts, 24, indices=idx, uniform=_persp_xf(S.rotate(0.9, 0.1, 0.0)), start=100, count=77,
                   state=[("srpRasterPointSize", 2.0)])]))
    add(S.Scene("vertex_buffer_start_offset_128x128", 128, 128, [
        vcolor(tris, start=7, count=100), vcolor(wavy, prim=H.SRP_PRIM_TRIANGLE_STRIP, start=3, count=30),
        vcolor(fan, prim=H.SRP_PRIM_TRIANGLE_FAN, start=1, count=12)]))

    # ---- stencil: ops, masks, separate faces, depth-fail op ----
    ops = [H.SRP_STENCIL_KEEP, H.SRP_STENCIL_ZERO, H.SRP_STENCIL_REPLACE, H.SRP_STENCIL_INCR, H.SRP_STENCIL_INCR_WRAP,
           H.SRP_STENCIL_DECR, H.SRP_STENCIL_DECR_WRAP, H.SRP_STENCIL_INVERT]
    many = random_color_tris(150, 9, spread=1.0, size=0.8)
    for k, op in enumerate(ops):
        add(S.Scene(f"stencil_op{k}_144x96", 144, 96, [
            vcolor(many, state=[("srpStencilTest", True), ("srpDepthTest", True),
                                ("srpStencilFunc", H.SRP_COMPARE_ALWAYS, 0x35, 0xFF),
                                ("srpStencilOp", H.SRP_STENCIL_INVERT, op, ops[(k + 3) % 8])]),
            vcolor(random_color_tris(60, 10), state=[("srpStencilFunc", H.SRP_COMPARE_LEQUAL, 3, 0x0F),
                                                     ("srpStencilOp", op, H.SRP_STENCIL_KEEP, H.SRP_STENCIL_DECR_WRAP),
                                                     ("srpStencilWriteMask", 0x3C)])]))
    add(S.Scene("stencil_separate_faces_180x180", 180, 180, [
        vcolor(many, uniform=_persp_xf(model), state=[
            ("srpStencilTest", True),
            ("srpStencilFuncSeparate", H.SRP_FACE_FRONT, H.SRP_COMPARE_ALWAYS, 1, 0xFF),
            ("srpStencilFuncSeparate", H.SRP_FACE_BACK, H.SRP_COMPARE_NOTEQUAL, 2, 0x03),
            ("srpStencilOpSeparate", H.SRP_FACE_FRONT, H.SRP_STENCIL_KEEP, H.SRP_STENCIL_KEEP, H.SRP_STENCIL_INCR),
            ("srpStencilOpSeparate", H.SRP_FACE_BACK, H.SRP_STENCIL_INCR_WRAP, H.SRP_STENCIL_KEEP, H.SRP_STENCIL_REPLACE),
            ("srpStencilWriteMaskSeparate", H.SRP_FACE_BACK, 0x0F)]),
        S.Draw("primid", H.SRP_PRIM_TRIANGLES, many, 24, uniform=_persp_xf(S.rotate(0.1, 0.2, 0.3)), state=[
            ("srpStencilFunc", H.SRP_COMPARE_GEQUAL, 2, 0xFF), ("srpStencilOp", H.SRP_STENCIL_KEEP, H.SRP_STENCIL_KEEP, H.SRP_STENCIL_KEEP)])]))

    # ---- late depth test: the fragment shader replaces depth on even columns ----
    add(S.Scene("late_depth_test_150x110", 150, 110, [
        S.Draw("depthout", H.SRP_PRIM_TRIANGLES, random_color_tris(50, 11), 24, uniform=_xf(), varyings=COLOR_VARY,
               varyings_size=12, may_overwrite_depth=True,
               state=[("srpDepthTest", True), ("srpStencilTest", True), ("srpStencilOp", H.SRP_STENCIL_KEEP, H.SRP_STENCIL_INCR, H.SRP_STENCIL_KEEP)]),
        S.Draw("depthout", H.SRP_PRIM_TRIANGLES, random_color_tris(50, 12), 24, uniform=_xf(), varyings=COLOR_VARY,
               varyings_size=12, may_overwrite_depth=True)]))

    # ---- scissor, including a box that leaves the framebuffer ----
    add(S.Scene("scissor_partial_171x99", 171, 99, [
        vcolor(tris, state=[("srpScissorTest", True), ("srpScissorOptions", 20, 10, 90, 60)]),
        S.Draw("primid", H.SRP_PRIM_LINES, tris, 24, uniform=_xf(), state=[("srpScissorOptions", 100, 50, 500, 500)]),
        S.Draw("primid", H.SRP_PRIM_POINTS, tris, 24, uniform=_xf(), state=[("srpRasterPointSize", 4.0), ("srpScissorOptions", 0, 0, 60, 99)])]))

    # ---- points: sizes, edges ----
    rng = np.random.RandomState(13)
    pts = np.zeros((300, 6), f32)
    pts[:, :2] = rng.uniform(-1.0, 1.0, (300, 2)); pts[:20, 0] = np.where(np.arange(20) % 2, 1.0, -1.0)
    pts[20:40, 1] = np.where(np.arange(20) % 2, 1.0, -1.0); pts[:, 2] = rng.uniform(-1, 1, 300); pts[:, 3:] = rng.uniform(0, 1, (300, 3))
    for size in (0.5, 1.0, 1.5, 2.0, 3.0, 7.25):
        add(S.Scene(f"points_size{size}_101x77", 101, 77, [
            vcolor(pts, prim=H.SRP_PRIM_POINTS, mode=H.SRP_INTERPOLATION_MODE_FLAT,
                   state=[("srpRasterPointSize", size), ("srpDepthTest", True)])]))
    add(S.Scene("points_size_zero_32x32", 32, 32, [vcolor(pts, prim=H.SRP_PRIM_POINTS, state=[("srpRasterPointSize", 0.0)])]))

    # ---- lines touching every framebuffer edge (the reference's index wrap, App. B-1) ----
    edge = np.zeros((64, 6), f32)
    a = np.linspace(0, 2 * np.pi, 32, endpoint=False)
    edge[0::2, :2] = 0.2 * np.stack([np.cos(a), np.sin(a)], -1)
    edge[1::2, :2] = np.stack([np.clip(1.6 * np.cos(a), -1, 1), np.clip(1.6 * np.sin(a), -1, 1)], -1)
    # fragments in the row below the framebuffer (ndc y = -1 -> py = height) are writes past
    # the end of the planes in the reference; they are dropped here and land in the oracle's
    # guard band there, so they are allowed in the scene
    edge[:, 3:] = 1
    for w, h in ((64, 64), (129, 65), (50, 201)):
        add(S.Scene(f"lines_to_edges_{w}x{h}", w, h, [S.Draw("primid", H.SRP_PRIM_LINES, edge, 24, uniform=_xf())]))
    add(S.Scene("polygon_line_to_edges_96x96", 96, 96, [
        vcolor(np.array([[-1, -1, 0, 1, 0, 0], [1, -1, 0, 0, 1, 0], [1, 1, 0, 0, 0, 1],
                         [-1, -1, 0.5, 1, 1, 0], [1, 1, 0.5, 0, 1, 1], [-1, 1, 0.5, 1, 0, 1]], f32),
               state=[("srpRasterPolygonMode", H.SRP_POLYGON_MODE_LINE)])]))

    # ---- degenerate input ----
    degen = np.array([[0, 0, 0, 1, 0, 0], [0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 1],           # zero area
                      [-0.5, -0.5, 0, 1, 0, 0], [0.5, 0.5, 0, 0, 1, 0], [0, 0, 0, 0, 0, 1],  # collinear
                      [-0.9, 0.2, 0, 1, 1, 0], [0.9, 0.2, 0, 0, 1, 1], [0, 0.9, 0, 1, 0, 1],  # fine
                      [0.3, 0.3, 0, 1, 1, 1], [0.3, 0.3, 0, 1, 1, 1]], f32)                 # 2 extra vertices
    add(S.Scene("degenerate_and_excess_80x80", 80, 80, [vcolor(degen), vcolor(degen, prim=H.SRP_PRIM_LINES),
                                                         vcolor(degen[:1], prim=H.SRP_PRIM_LINE_STRIP),
                                                         vcolor(degen[:2], prim=H.SRP_PRIM_TRIANGLE_STRIP)]))

    # ---- texture wrap modes ----
    tverts, tidx = S.cube_mesh()
    tverts = tverts.copy(); tverts[:, 3:] = tverts[:, 3:] * 2.5 - 0.75        # uv outside [0, 1]
    tex = procedural_texture(64, seed=3)
    for wx, wy, tag in ((H.TW_REPEAT, H.TW_REPEAT, "repeat"), (H.TW_CLAMP_TO_EDGE, H.TW_CLAMP_TO_EDGE, "clamp"),
                        (H.TW_REPEAT, H.TW_CLAMP_TO_EDGE, "mixed")):
        add(S.Scene(f"texture_wrap_{tag}_222x222", 222, 222, [
            S.Draw("texcube", H.SRP_PRIM_TRIANGLES, tverts, 20, indices=tidx,
                   uniform=S.texcube_uniform(S.rotate(0.5, 0.8, 0.1), S.view((0, 0, -3)), S.perspective(-1, 1, -1, 1, 1, 50)),
                   varyings=[(2, H.SRP_FLOAT, H.SRP_INTERPOLATION_MODE_PERSPECTIVE)], varyings_size=8,
                   state=[("srpRasterCullFace", H.SRP_FACE_BACK), ("srpDepthTest", True)])],
            textures={"wall": (tex, wx, wy)}))

    # ---- several draws without clearing in between, uniform changed between draws ----
    add(S.Scene("multi_draw_no_clear_140x140", 140, 140, [
        S.Draw("solid", H.SRP_PRIM_TRIANGLES, tris, 24, uniform=_xf() + np.array([1, 0, 0, 1], f32).tobytes()),
        S.Draw("solid", H.SRP_PRIM_TRIANGLES, tris, 24, uniform=_xf(S.scale(0.5, 0.5, 1)) + np.array([0, 1, 0.5, 1], f32).tobytes(),
               clear_before=True),
        S.Draw("solid", H.SRP_PRIM_LINE_LOOP, tris[:17], 24, uniform=_xf() + np.array([0.2, 0.3, 1, 0.5], f32).tobytes())]))

    # ---- large triangles: the barycentric-checkpoint path (boxes far wider than a tile) ----
    quad = np.array([[-1, -1, 0.2, 1, 0, 0], [1, -1, 0.2, 0, 1, 0], [1, 1, -0.3, 0, 0, 1],
                     [-1, -1, 0.2, 1, 0, 0], [1, 1, -0.3, 0, 0, 1], [-1, 1, 0.1, 1, 1, 0],
                     [-0.95, 0.9, 0.5, 1, 1, 1], [0.1, -0.97, -0.5, 0.2, 0.2, 0.2], [0.93, 0.55, 0.4, 0, 1, 1],
                     [-2.5, -0.3, 0.0, 1, 0, 1], [2.5, -0.2, 0.0, 0, 1, 0], [0.0, 3.0, 0.0, 0, 0, 1]], f32)
    add(S.Scene("large_triangles_700x500", 700, 500, [
        vcolor(quad, state=[("srpDepthTest", True)]),
        vcolor(quad, uniform=_persp_xf(S.rotate(0.5, 0.9, 1.3), cam=(0, 0, -1.6), near=0.5, far=10), mode=H.SRP_INTERPOLATION_MODE_AFFINE),
        S.Draw("primid", H.SRP_PRIM_TRIANGLES, quad, 24, uniform=_xf(S.rotate(0, 0, 2.2)),
               state=[("srpDepthCompareOp", H.SRP_COMPARE_GEQUAL), ("srpRasterCullFace", H.SRP_FACE_BACK)])]))

    # ---- a mid-size mesh that takes the binned path (thousands of primitives) ----
    mverts, midx = grid_mesh(70, 50, z_fn=lambda x, y: 0.4 * np.sin(3 * x + y) * np.cos(2 * y - x))
    add(S.Scene("binned_mesh_640x360", 640, 360, [
        vcolor(mverts, indices=midx.astype(np.uint32), uniform=_persp_xf(S.rotate(1.0, 0.2, 0.1), cam=(0, 0, -1.4), near=0.5, far=10),
               state=[("srpDepthTest", True)]),
        S.Draw("primid", H.SRP_PRIM_TRIANGLES, mverts, 24, indices=midx.astype(np.uint32),
               uniform=_persp_xf(S.rotate(1.1, 0.25, 0.1), cam=(0, 0, -1.4), near=0.5, far=10),
               state=[("srpRasterPolygonMode", H.SRP_POLYGON_MODE_LINE)])]))
    return out
