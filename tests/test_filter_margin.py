"""Host-side check of the packed line filter's one-sided margin (atx_device.cuh: filter_sphere).

The filter decides "the reference skips this sphere" from a chain that cancels |c|^2 + |o|^2 - 2 o.c
instead of forming o - c first (8 packed FMAs per two tests instead of 12 FP ops); it is only allowed to
reject a sphere when the reference's own computed discriminant is negative (Renderer.cu:263-268), and the
margins 2^-17 (|c|^2 + r^2) per sphere and 2^-17 a |o|^2 per ray are what guarantee that.
Both float32 chains are restated here in numpy (fma = one rounding of the exact product-sum, formed in
float64) and searched for false negatives on the config-3 / config-4 scenes with bounce rays, camera rays
and rays grazing a sphere within 1e-6 of its radius; the candidate inflation is reported. The same chain
with the margin cut to 2^-23 must show false negatives, or this test would prove nothing.
"""
import numpy as np
import pytest

f32, f64 = np.float32, np.float64
MARGIN = f32(2.0 ** -17)   # kFilterMargin


def fma(a, b, c):
    return (np.asarray(a, f64) * np.asarray(b, f64) + np.asarray(c, f64)).astype(f32)


def mul(a, b):
    return (np.asarray(a, f32) * np.asarray(b, f32)).astype(f32)


def add(a, b):
    return (np.asarray(a, f32) + np.asarray(b, f32)).astype(f32)


def dot3(ax, ay, az, bx, by, bz):
    return fma(az, bz, fma(ax, bx, mul(ay, by)))   # fdot3 (atx_exact.cuh)


def keeps(centers, radii, o, d, margin):
    """(reference keeps, filter keeps) for every ray x sphere pair."""
    nc, r = (-centers).astype(f32), radii.astype(f32)
    c2 = (centers.astype(f64) ** 2).sum(1)
    r2 = radii.astype(f64) ** 2
    kk = (c2 - r2 - f64(margin) * (c2 + r2)).astype(f32)                     # pack_scene_kernel
    o, d = o.astype(f32), d.astype(f32)
    ox, oy, oz = (o[:, i:i + 1] for i in range(3))
    dx, dy, dz = (d[:, i:i + 1] for i in range(3))
    ncx, ncy, ncz = (nc[None, :, i] for i in range(3))
    one = np.ones((len(o), len(nc)), f32)
    a = dot3(dx, dy, dz, dx, dy, dz)
    # the reference's sequence (exact_test / Renderer.cu:258-267 as compiled)
    ocx, ocy, ocz = add(ox, ncx), add(oy, ncy), add(oz, ncz)
    hb = dot3(ocx, ocy, ocz, dx * one, dy * one, dz * one)
    cc = fma(-r[None, :] * one, r[None, :] * one, dot3(ocx, ocy, ocz, ocx, ocy, ocz))
    b = add(hb, hb)
    disc = fma(b, b, -mul(mul(a, f32(4.0)) * one, cc))
    ref = ~(disc < 0)
    # filter_sphere
    od, oo = dot3(ox, oy, oz, dx, dy, dz), dot3(ox, oy, oz, ox, oy, oz)
    g = add(fma(-a, oo, mul(a, mul(oo, margin))), f32(2.0 ** -100))          # ray_pair_lane
    o2x, o2y, o2z = add(ox, ox), add(oy, oy), add(oz, oz)
    hbp = fma(ncz * one, dz * one, fma(ncx * one, dx * one, fma(ncy * one, dy * one, od * one)))
    t = fma(ncz * one, o2z * one, fma(ncx * one, o2x * one, fma(ncy * one, o2y * one, kk[None, :] * one)))
    pre = fma(hbp, hbp, fma(t, -a * one, g * one))
    return ref, ~np.signbit(pre)


def ray_sets(scene, atx, rng, n):
    s = atx.pack_spheres(atx.traverseSceneGraph(scene.rootNode))
    C, R = s["center"], s["radius"]
    S = len(C)
    unit = lambda v: v / np.linalg.norm(v, axis=1, keepdims=True)  # noqa: E731
    idx = rng.integers(0, S, n)
    nrm = unit(rng.normal(size=(n, 3)))
    yield "bounce", C, R, C[idx] + nrm * R[idx, None] * (1 + 1e-4), unit(rng.normal(size=(n, 3)))
    cam = np.asarray(scene.camera.getPosition(), f64)
    tgt = C[rng.integers(0, S, n)] + rng.normal(size=(n, 3)) * 0.5
    yield "camera", C, R, np.repeat(cam[None], n, 0), unit(tgt - cam)
    j = rng.integers(0, S - 1, n)                      # not the ground sphere
    nrm = unit(rng.normal(size=(n, 3)))
    tan = unit(np.cross(nrm, rng.normal(size=(n, 3))))
    pt = C[j] + nrm * R[j, None] * (1 + rng.normal(size=(n, 1)) * 1e-6)
    yield "grazing", C, R, pt - tan * rng.uniform(1, 40, (n, 1)), tan


@pytest.mark.parametrize("name,n", [("config3", 1500), ("config4", 200)])
def test_filter_never_rejects_what_the_reference_keeps(built, name, n):
    import ataraxia_b200 as atx
    scene = getattr(atx.synthetic, name)()
    rng = np.random.default_rng(20261018)
    missed_with_thin_margin = 0
    for label, C, R, o, d in ray_sets(scene, atx, rng, n):
        ref, mine = keeps(C, R, o, d, MARGIN)
        assert not (ref & ~mine).any(), (name, label, int((ref & ~mine).sum()))
        assert mine.sum() <= 1.5 * ref.sum(), (name, label)       # the margin costs few extra candidates
        print(f"{name} {label}: {ref.sum()} kept by the reference, {mine.sum()} by the filter of {ref.size} pairs")
        _, thin = keeps(C, R, o, d, f32(2.0 ** -23))
        missed_with_thin_margin += int((ref & ~thin).sum())
    if name == "config3":
        assert missed_with_thin_margin > 0, "the search does not reach the rounding error it is meant to bound"
