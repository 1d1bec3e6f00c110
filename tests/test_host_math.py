"""CPU tests: the host-side vec / mat API (host half) against the
unmodified reference library (oracle/_ref, built from /root/reference by oracle/Makefile) on
seeded random and edge-case inputs, bit for bit.  Reference: src/math/vec.c:16-189,
src/math/mat.c:16-178 -- left-associative sums of products, no FMA contraction, sqrtf, 1.0f/len.
The device twins of the same functions are exercised by the GPU parity tests (every shader
goes through them)."""
import ctypes as C

import numpy as np
import pytest

from srp_b200 import host


def vec_type(n):
    class V(C.Structure):
        _pack_ = 1
        _fields_ = [("v", C.c_float * n)]
    V.__name__ = f"vec{n}"
    return V


VEC = {n: vec_type(n) for n in (2, 3, 4)}
EDGE = np.array([0.0, -0.0, 1.0, -1.0, 1e-38, -1e-38, 1e-45, 3.4e38, -3.4e38, 1e19, 1e-19, 0.1, 255.0, 1 / 255.0], np.float32)


def inputs(n, count, seed):
    """`count` vectors of n floats: wide-range random values mixed with edge cases"""
    rng = np.random.default_rng(seed)
    a = (rng.standard_normal((count, n)) * 10.0 ** rng.integers(-6, 7, (count, 1))).astype(np.float32)
    pick = rng.random((count, n)) < 0.15
    a[pick] = rng.choice(EDGE, int(pick.sum()))
    a[0] = 0.0      # the zero vector (Normalize leaves it alone)
    return a


def bind(dll, name, res, args):
    fn = getattr(dll, name)
    fn.restype, fn.argtypes = res, args
    return fn


def bits(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32)).view(np.uint32)


@pytest.mark.parametrize("n", [2, 3, 4])
def test_vec_api_matches_reference(reference, n):
    prod = host.load_product()
    V = VEC[n]
    ops = {f"vec{n}Add": (V, [V, V]), f"vec{n}Subtract": (V, [V, V]), f"vec{n}DotProduct": (C.c_float, [V, V]),
           f"vec{n}MultiplyScalar": (V, [V, C.c_float]), f"vec{n}Normalize": (V, [V]), f"vec{n}Reflect": (V, [V, V]),
           f"vec{n}MultiplyVec{n}": (V, [V, V]), f"vec{n}Negate": (V, [V])}
    a, b = inputs(n, 400, 100 + n), inputs(n, 400, 200 + n)
    for name, (res, args) in ops.items():
        fp, fr = bind(prod.dll, name, res, args), bind(reference.dll, name, res, args)
        for x, y in zip(a, b):
            call = [V((C.c_float * n)(*x))]
            if len(args) == 2:
                call.append(C.c_float(float(y[0])) if args[1] is C.c_float else V((C.c_float * n)(*y)))
            with np.errstate(all="ignore"):
                got, want = fp(*call), fr(*call)
            g = bits(got if res is C.c_float else list(got.v))
            w = bits(want if res is C.c_float else list(want.v))
            assert np.array_equal(g, w), f"{name}({x}, {y}): {g} != {w}"


def test_mat_products_match_reference(reference):
    prod = host.load_product()
    M, V4 = host.Mat4, VEC[4]
    mv_p = bind(prod.dll, "mat4MultiplyVec4", V4, [C.POINTER(M), V4])
    mv_r = bind(reference.dll, "mat4MultiplyVec4", V4, [C.POINTER(M), V4])
    mm_p = bind(prod.dll, "mat4MultiplyMat4", M, [C.POINTER(M), C.POINTER(M)])
    mm_r = bind(reference.dll, "mat4MultiplyMat4", M, [C.POINTER(M), C.POINTER(M)])
    mats, vecs = inputs(16, 120, 7), inputs(4, 120, 8)
    for i in range(len(mats)):
        a, b = M((C.c_float * 16)(*mats[i])), M((C.c_float * 16)(*mats[(i + 1) % len(mats)]))
        v = V4((C.c_float * 4)(*vecs[i]))
        assert np.array_equal(bits(list(mv_p(C.byref(a), v).v)), bits(list(mv_r(C.byref(a), v).v))), i
        assert np.array_equal(bits(list(mm_p(C.byref(a), C.byref(b)).data)), bits(list(mm_r(C.byref(a), C.byref(b)).data))), i


def test_mat_constructors_match_reference(reference):
    prod = host.load_product()
    rng = np.random.default_rng(11)
    table = {"mat4ConstructScale": 3, "mat4ConstructTranslate": 3, "mat4ConstructRotate": 3, "mat4ConstructTRS": 9,
             "mat4ConstructView": 9, "mat4ConstructOrthogonalProjection": 6, "mat4ConstructPerspectiveProjection": 6}
    assert np.array_equal(bits(prod.mat4("mat4ConstructIdentity")), bits(reference.mat4("mat4ConstructIdentity")))
    for fn, k in table.items():
        for _ in range(200):
            args = (rng.standard_normal(k) * 10.0 ** rng.integers(-2, 3)).astype(np.float32).tolist()
            with np.errstate(all="ignore"):
                assert np.array_equal(bits(prod.mat4(fn, *args)), bits(reference.mat4(fn, *args))), (fn, args)
