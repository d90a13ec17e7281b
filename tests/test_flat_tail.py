"""Host-side check of the branch-free sphere test of the small-scene kernels (atx_device.cuh: flat_tail).

The reference (Renderer.cu:263-278, built -use_fast_math: .ftz arithmetic, approximate sqrt and reciprocal) does, per sphere,

    if (disc < 0) continue;  t0 = (-b - sqrt(disc)) / 2a;  t1 = (-b + sqrt(disc)) / 2a;  t = t0 < t1 ? t0 : t1;
    if (t > 0 && t < tmin) { tmin = t; closest = i; }

flat_tail runs no branch and forms no second root: t = (-b - sqrt.approx(disc)) * rcp.approx(2a), hit = t > 0 && t < tmin.
The header argues that (hit, t) are the reference's for EVERY input: a negative disc gives a NaN t, and t0 <= t1 whenever both
are numbers because sq >= 0 and the reciprocal of 2a = 2 dot(d,d) is never negative. Both procedures are restated here in numpy
float32 with flush-to-zero, and compared on a grid of special values (zeros of both signs, denormals, infinities, NaN, the
largest and smallest normals) crossed with each other, and on random inputs. The argument must not depend on how good the
approximate sqrt and reciprocal are - only on their sign (the monotonic roundings it uses are those of the exact add and multiply)
- so the restatement perturbs both by up to +-8 ulp. A deliberately broken variant (reciprocal allowed to be negative) must produce mismatches, or the comparison
would prove nothing.
"""
import numpy as np

f32 = np.float32
TINY = np.finfo(f32).tiny          # smallest normal
FLT_MAX = np.finfo(f32).max


def ftz(x):
    x = np.asarray(x, f32).copy()
    x[np.abs(x) < TINY] *= f32(0.0)          # keeps the sign: -denormal -> -0
    return x


def add(a, b):
    with np.errstate(all="ignore"):
        return ftz(ftz(a) + ftz(b))


def mul(a, b):
    with np.errstate(all="ignore"):
        return ftz(ftz(a) * ftz(b))


def ulps(x, k):
    """x moved by k float32 steps (k an integer array), sign and specials preserved."""
    x = np.asarray(x, f32)
    i = x.view(np.int32).copy()
    ok = np.isfinite(x) & (x != 0)
    j = i.copy()
    j[ok] = np.where(i[ok] >= 0, np.maximum(i[ok] + k[ok], 1), np.minimum(i[ok] - k[ok], -2147483647))
    out = j.view(f32).copy()
    bad = ~np.isfinite(out) & ok
    out[bad] = x[bad]
    return out


def sqrt_approx(x, k):
    """sqrt.approx.ftz: NaN below zero, -0 -> -0, otherwise a non-negative number near the root."""
    x = ftz(x)
    with np.errstate(all="ignore"):
        r = np.sqrt(x).astype(f32)
    r = np.where(np.isfinite(r) & (r > 0), np.abs(ulps(r, k)), r).astype(f32)
    return ftz(r)


def rcp_approx(x, k, broken=False):
    """rcp.approx.ftz of a non-negative number: +inf at +0, never negative."""
    x = ftz(x)
    with np.errstate(all="ignore"):
        r = (f32(1.0) / x).astype(f32)
    r = np.where(np.isfinite(r) & (r > 0), np.abs(ulps(r, k)), r).astype(f32)
    if broken:
        r = -r
    return ftz(r)


def lt(a, b):
    with np.errstate(all="ignore"):
        return ftz(a) < ftz(b)


def both(b, disc, a2, tmin, k1, k2, broken=False):
    """(hit, t) of the reference's sequence and of flat_tail on the same b, disc, 2a, tmin."""
    sq = sqrt_approx(disc, k1)
    r = rcp_approx(a2, k2, broken)
    nb = -np.asarray(b, f32)
    t0 = mul(add(nb, -sq), r)                 # fsub(fneg(b), sq) * rcp
    t1 = mul(add(sq, nb), r)                  # fsub(sq, b) * rcp
    t = np.where(lt(t0, t1), t0, t1)
    ref_hit = ~lt(disc, f32(0.0)) & lt(f32(0.0), t) & lt(t, tmin)
    flat_hit = lt(f32(0.0), t0) & lt(t0, tmin)
    return ref_hit, t, flat_hit, t0


def check(b, disc, a2, tmin, rng, broken=False):
    k1 = rng.integers(-8, 9, b.shape).astype(np.int32)
    k2 = rng.integers(-8, 9, b.shape).astype(np.int32)
    ref_hit, t, flat_hit, t0 = both(b, disc, a2, tmin, k1, k2, broken)
    same = (ref_hit == flat_hit) & (~ref_hit | (t.view(np.uint32) == t0.view(np.uint32)))
    return same


SPECIALS = np.array([0.0, -0.0, 1e-45, -1e-45, 1e-39, -1e-39, TINY, -TINY, 1.5 * TINY, 1e-20, -1e-20, 1e-4, -1e-4, 0.5, -0.5, 1.0,
                     -1.0, 2.0, 1000.0, -1000.0, 1e20, -1e20, FLT_MAX, -FLT_MAX, np.inf, -np.inf, np.nan], dtype=f32)


def test_special_values_grid():
    rng = np.random.default_rng(7)
    a2 = SPECIALS[(SPECIALS >= 0) | np.isnan(SPECIALS)]              # 2 dot(d,d) is never negative
    tmin = np.array([FLT_MAX, 1.0, 1e-30, 0.0], dtype=f32)
    B, D, A, T = np.meshgrid(SPECIALS, SPECIALS, a2, tmin, indexing="ij")
    same = check(B.ravel(), D.ravel(), A.ravel(), T.ravel(), rng)
    assert same.all(), list(zip(B.ravel()[~same][:5], D.ravel()[~same][:5], A.ravel()[~same][:5], T.ravel()[~same][:5]))


def test_random_inputs_from_real_geometry():
    """b, disc, 2a as the kernels form them (oc.d, its discriminant, 2 d.d) for random rays and spheres, including rays that
    start inside, on and far outside a sphere, near-tangent rays and near-zero directions."""
    rng = np.random.default_rng(11)
    n = 400000
    scale = 10.0 ** rng.uniform(-3, 3, n)
    oc = rng.normal(size=(n, 3)) * scale[:, None]
    d = rng.normal(size=(n, 3))
    aimed = rng.uniform(size=n) < 0.6                                  # most rays point at the sphere, give or take
    d[aimed] = -oc[aimed] / np.linalg.norm(oc[aimed], axis=1, keepdims=True) + rng.normal(size=(int(aimed.sum()), 3)) * 0.3
    d *= (10.0 ** rng.uniform(-22, 1, n))[:, None]
    radius = np.abs(rng.normal(size=n)) * scale * rng.choice([0.0, 1e-3, 0.999, 1.0, 1.001, 3.0], n)
    oc, d, radius = oc.astype(f32), d.astype(f32), radius.astype(f32)
    dot = lambda p, q: add(mul(p[:, 2], q[:, 2]), add(mul(p[:, 0], q[:, 0]), mul(p[:, 1], q[:, 1])))   # noqa: E731 (rounding differs from fma: irrelevant here)
    hb, a = dot(oc, d), dot(d, d)
    cc = add(dot(oc, oc), -mul(radius, radius))
    b = add(hb, hb)
    disc = add(mul(b, b), -mul(mul(a, f32(4.0)), cc))
    tmin = np.where(rng.uniform(size=n) < 0.5, FLT_MAX, np.abs(rng.normal(size=n)) * scale).astype(f32)
    same = check(b, disc, add(a, a), tmin, rng)
    assert same.all(), int((~same).sum())
    ref_hit = both(b, disc, add(a, a), tmin, np.zeros(n, np.int32), np.zeros(n, np.int32))[0]
    assert 0.05 < ref_hit.mean() < 0.95          # the sample exercises both outcomes


def test_the_comparison_can_fail():
    """With a reciprocal of the wrong sign t0 > t1 and the two procedures part: the check is not vacuous."""
    rng = np.random.default_rng(3)
    n = 20000
    b = -np.abs(rng.normal(size=n)).astype(f32)
    disc = np.abs(rng.normal(size=n)).astype(f32)
    a2 = np.abs(rng.normal(size=n)).astype(f32) + f32(0.1)
    tmin = np.full(n, FLT_MAX, f32)
    assert not check(b, disc, a2, tmin, rng, broken=True).all()
