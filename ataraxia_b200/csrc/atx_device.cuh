// atx_device.cuh — device data layout and the per-path building blocks shared by
// the megakernel and the wavefront kernels. Every function cites the reference
// code it restates (file:line under /root/reference) and follows the op sequence of
// the reference's own sm_100a SASS, using only atx_exact.cuh ops.
#pragma once
#include "atx_exact.cuh"

namespace atxk
{

// ---------------------------------------------------------------------------
// Scene records in HBM, written once per upload by pack_scene_kernel.
//   spheres : float4 (-cx, -cy, -cz, radius) the exact test and the hit record. The centre is
//                                           stored NEGATED: o - c == o + (-c) bit for bit
//   sphFilter: float4 (-cx, -cy, -cz, kk)   one LDS.128 per packed filter step (two rays);
//                                           kk = |c|^2 - r^2 - 2^-17 (|c|^2 + r^2), formed in double
//   sphMat  : int32 material index          read only for the winning sphere
//   mats    : 6 x float4 per material       everything the shading needs, including the
//                                           per-material subexpressions the reference
//                                           recomputes every bounce (same device ops,
//                                           so same bits)
//   lights  : 2 x float4 per light          position, intensity*color
// ---------------------------------------------------------------------------
constexpr int kMatStride = 6;
//  mats[6m+0] = albedo.xyz, roughness
//  mats[6m+1] = Fb.xyz (= mix(F0, albedo, metallic), Renderer.cu:335), metallic
//  mats[6m+2] = (1 - Fb).xyz, 1 - metallic
//  mats[6m+3] = a2, a2 - 1, k, 1 - k              (BRDF.cu:42-63)
//  mats[6m+4] = emissionColor*emissionIntensity .xyz, emissionIntensity   (Scene.h:41)
//  mats[6m+5] = fma(a, a, -1) for sampleGGX (BRDF.cu:102), 0, 0, 0
constexpr int kLightStride = 2;
//  lights[2l+0] = position.xyz, 0
//  lights[2l+1] = intensity*color .xyz, 0          (Renderer.cu:362)

struct CameraParams
{
    // inverse projection: columns 0,1 and the per-launch constant (col2*1 + col3*1)
    float ip0[4], ip1[4], ipA1[4];
    // inverse view: columns 0..2 (xyz) and col3*0 (xyz) — kept so signed zeros/NaNs match
    float iv0[3], iv1[3], iv2[3], iv3z[3];
    float pos[3];
};

struct RenderParams
{
    uint32_t width, height;
    uint32_t firstFrame, nFrames, frameStride;
    int32_t maxBounces;
    int32_t skyLight;
    int32_t zeroFirst;      // accumulation starts from 0 instead of the stored value
    int32_t emitRgba;       // pack RGBA8 at the end of the launch
    uint32_t rgbaDivisor;   // frameIndex used for the display divide (Renderer.cu:166)
    uint32_t nSpheres, nMaterials, nLights;
    uint32_t chunkSpheres;  // spheres per shared-memory chunk (>= nSpheres: staged once)
    uint32_t parkThreshold; // while-while kernel: parked hits (of 32 lanes) that trigger the bounce phase
    uint32_t claimThreshold; // idle lanes (of 32) that trigger a batched pixel claim
    uint32_t poolSize;      // pixel ids in the pool: 8x4 tiles x 32, padded tiles included
    uint32_t* pool;         // next unclaimed id (zeroed before the launch)
    // image-tile split across GPUs (atx_render_tiles): this rank renders the 8x4 tiles tileOffset, tileOffset +
    // tileStride, ... (the pool holds only those) and stores every finished pixel into the image of EVERY rank:
    // its own (accum) and the peers' buffers mapped over NVLink (push[0 .. nPush))
    uint32_t tileStride, tileOffset, nTiles;
    uint32_t nPush;
    float4* push[7];
    // per-pixel constants of the launch for the warp-queue form with at most one light (pixel_prologue_kernel):
    // kPrologueStride float4 per pixel
    float4* pixelCache;
    const float4* spheres;    // (-cx, -cy, -cz, r): the exact tests, hit records
    const float4* sphFilter;  // (-cx, -cy, -cz, |c|^2 - r^2 - margin): the packed line filter (filter_sphere)
    const int32_t* sphMat;
    const float4* mats;
    const float4* lights;
    float4* accum;
    uint32_t* rgba;
    unsigned long long* counters; // [0]=paths [1]=rays (reference traceRay calls) [2]=rays whose sphere loop ran
    CameraParams cam;
};

struct V3 { float x, y, z; };

// pixelCache record (6 float4 per pixel, written by pixel_prologue_kernel, read once by the lane that claims the pixel):
//   [0] tPrimary, cPrimary (int bits; < 0: the primary ray misses everything), pr0, ggxT0
//   [1] c0.rgb (color after the frame-independent first bounce), raysPerStart (uint bits) | ggx0 << 8
//   [2] o0.xyz (origin of every frame's bounce ray), tq0.x      [3] N0.xyz, tq0.y      [4] T0.xyz, tq0.z      [5] B0.xyz, -
constexpr uint32_t kPrologueStride = 6;

// ---------------------------------------------------------------------------
// Primary ray, Camera::UpdateRayDirection (Camera.cpp:161-195) — host IEEE
// arithmetic in glm's order: coord = (x/W, y/H)*2 - 1; target = invProj*(cx,cy,1,1)
// as (m0*cx + m1*cy) + (m2*1 + m3*1) (type_mat4x4.inl:562-573);
// n = normalize(target.xyz / target.w) = v * (1/sqrt((x*x + y*y) + z*z));
// dir = normalize(vec3(invView * vec4(n, 0))).
// ---------------------------------------------------------------------------
ATX_DEV V3 ieee_normalize(float x, float y, float z)
{
    const float d = ieee_add(ieee_add(ieee_mul(x, x), ieee_mul(y, y)), ieee_mul(z, z));
    const float inv = ieee_div(1.0f, ieee_sqrt(d));
    return { ieee_mul(x, inv), ieee_mul(y, inv), ieee_mul(z, inv) };
}

ATX_DEV V3 primary_direction(const CameraParams& c, uint32_t x, uint32_t y, uint32_t width, uint32_t height)
{
    const float cx = ieee_sub(ieee_mul(ieee_div(static_cast<float>(x), static_cast<float>(width)), 2.0f), 1.0f);
    const float cy = ieee_sub(ieee_mul(ieee_div(static_cast<float>(y), static_cast<float>(height)), 2.0f), 1.0f);
    float t[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
        t[i] = ieee_add(ieee_add(ieee_mul(c.ip0[i], cx), ieee_mul(c.ip1[i], cy)), c.ipA1[i]);
    const V3 n = ieee_normalize(ieee_div(t[0], t[3]), ieee_div(t[1], t[3]), ieee_div(t[2], t[3]));
    float r[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
        r[i] = ieee_add(ieee_add(ieee_mul(c.iv0[i], n.x), ieee_mul(c.iv1[i], n.y)),
                        ieee_add(ieee_mul(c.iv2[i], n.z), c.iv3z[i]));
    return ieee_normalize(r[0], r[1], r[2]);
}

// ---------------------------------------------------------------------------
// Ray / sphere loop, Renderer::traceRay (Renderer.cu:251-285).
//
// Ground truth is the reference's sm_100a SASS, not its PTX: the PTX carries
// mul/add/sub without ".rn", which ptxas is allowed to (and does) contract. Per
// sphere the reference executes
//     oc = o - c                      3 FADD
//     hb = fma(oc.z,d.z, fma(oc.x,d.x, oc.y*d.y))          FMUL + 2 FFMA
//     q  = fma(oc.z,oc.z, fma(oc.x,oc.x, oc.y*oc.y))       FMUL + 2 FFMA
//     c' = fma(-r, r, q)              FFMA   (r*r is NOT rounded separately)
//     b  = hb + hb                    FADD
//     disc = fma(b, b, -(a4*c'))      FMUL + FFMA
//     if disc < 0 skip
// The miss path here tests sign(fma(hb,hb, -(a*c'))) instead: b*b = 4*hb*hb exactly
// and (4a)*c' = 4*(a*c') exactly (power-of-two scalings commute with rounding), so
// disc = 4 * fma(hb,hb,-(a*c')) and the signs agree; the only exceptions (one value
// flushed to zero, the other not) land on the not-less-than-zero side here and are
// re-decided below with the reference's literal sequence. 12 FP instructions per
// missed sphere instead of the reference's 13.
// ---------------------------------------------------------------------------
struct RayConst
{
    float a;   // dot(d,d)
    float a4;  // a * 4
    float a2;  // a + a
};

ATX_DEV RayConst ray_constants(float dx, float dy, float dz)
{
    RayConst k;
    k.a = fdot3(dx, dy, dz, dx, dy, dz);
    k.a4 = fmul(k.a, 4.0f);
    k.a2 = fadd(k.a, k.a);
    return k;
}

// The reference's literal sequence for one sphere whose line test did not rule it out
// (kernelRender SASS, Renderer.cu:263-278). sp = (-cx, -cy, -cz, radius).
ATX_DEV void exact_tail(float hb, float cc, int index, const RayConst& k, float& tmin, int& closest)
{
    const float b = fadd(hb, hb);
    const float disc = ffma(b, b, fneg(fmul(k.a4, cc)));
    if (!(disc < 0.0f))
    {
        const float sq = fsqrt_approx(disc);
        const float t0 = fdiv_approx(fsub(fneg(b), sq), k.a2);
        const float t1 = fdiv_approx(fsub(sq, b), k.a2);
        const float t = t0 < t1 ? t0 : t1;
        if (t > 0.0f && t < tmin)
        {
            tmin = t;
            closest = index;
        }
    }
}

// full test of one sphere from scratch (candidate resolution of the packed loop)
ATX_DEV void exact_test(const float4 sp, int index, float ox, float oy, float oz, float dx, float dy, float dz,
                        const RayConst& k, float& tmin, int& closest)
{
    const float ocx = fadd(ox, sp.x);
    const float ocy = fadd(oy, sp.y);
    const float ocz = fadd(oz, sp.z);
    const float hb = fdot3(ocx, ocy, ocz, dx, dy, dz);
    const float cc = ffma(fneg(sp.w), sp.w, fdot3(ocx, ocy, ocz, ocx, ocy, ocz));
    exact_tail(hb, cc, index, k, tmin, closest);
}

// The reference's sequence for one sphere WITHOUT its branches (Renderer.cu:263-278), 11 instructions behind hb and cc:
//  * "if (disc < 0) continue" is dropped: sqrt.approx of a negative disc is NaN, so is t, and "t > 0 && t < tmin" is
//    false - the outcome of the skip; a NaN or -0 disc takes the same instructions in both;
//  * t = min(t0, t1) is t0: with sq = sqrt.approx(disc) >= 0 (or -0) the exact values satisfy -b - sq <= sq - b,
//    rounding to nearest is monotonic, and so is the multiplication by rcp.approx(a2) >= 0 (a2 = 2 dot(d,d) >= +0, so
//    the reciprocal is never negative): t0 <= t1 whenever both are numbers, and where they are equal they are the same
//    number (a zero of either sign fails "t > 0"). t0 is NaN with t1 a number only for b = -inf (t1 = +inf, no hit:
//    "t < tmin") or a2 in {0, inf} with t1 in {NaN, 0, inf} (no hit) - and a NaN t0 reports no hit as well. So
//    (hit, t) are those of "t0 < t1 ? t0 : t1" for every input, and t1 never needs to exist.
// With a handful of spheres this is fewer issue slots than a filter in front plus a hit branch that runs a few lanes
// wide almost every time (3-sphere scene: 23 + 18 instructions per sphere), and the tests of one ray interleave.
ATX_DEV void flat_tail(float hb, float cc, int index, const RayConst& k, float& tmin, int& closest)
{
    const float b = fadd(hb, hb);
    const float disc = ffma(b, b, fneg(fmul(k.a4, cc)));
    const float sq = fsqrt_approx(disc);
    const float t = fdiv_approx(fsub(fneg(b), sq), k.a2);
    const bool hit = t > 0.0f && t < tmin;
    tmin = hit ? t : tmin;
    closest = hit ? index : closest;
}

ATX_DEV void exact_flat(const float4 sp, int index, float ox, float oy, float oz, float dx, float dy, float dz,
                        const RayConst& k, float& tmin, int& closest)
{
    const float ocx = fadd(ox, sp.x);
    const float ocy = fadd(oy, sp.y);
    const float ocz = fadd(oz, sp.z);
    const float hb = fdot3(ocx, ocy, ocz, dx, dy, dz);
    const float cc = ffma(fneg(sp.w), sp.w, fdot3(ocx, ocy, ocz, ocx, ocy, ocz));
    flat_tail(hb, cc, index, k, tmin, closest);
}

// scalar test with the cheap line filter in front (small scenes, debug kernels)
ATX_DEV void intersect_sphere(const float4 sp, int index, float ox, float oy, float oz, float dx, float dy, float dz,
                              const RayConst& k, float& tmin, int& closest)
{
    const float ocx = fadd(ox, sp.x);
    const float ocy = fadd(oy, sp.y);
    const float ocz = fadd(oz, sp.z);
    const float hb = fdot3(ocx, ocy, ocz, dx, dy, dz);
    const float cc = ffma(fneg(sp.w), sp.w, fdot3(ocx, ocy, ocz, ocx, ocy, ocz));
    const float pre = ffma(hb, hb, fneg(fmul(k.a, cc)));
    // hb >= 0 (or NaN): b = hb+hb >= 0, so t0 = (-b - sqrt(disc))/2a <= 0 or NaN and min(t0,t1) fails "t > 0"
    // (Renderer.cu:269-272): the sphere is behind the origin, or the origin sits on it pointing away (the
    // 1e-4 offset rays leaving a surface) - no hit either way, so the literal sequence is skipped
    if (!(pre < 0.0f) && hb < 0.0f)
        exact_tail(hb, cc, index, k, tmin, closest);
}

// ---------------------------------------------------------------------------
// Packed filter: TWO rays (the two path slots of a thread) against one sphere per step,
// every op an f32x2 instruction with the sphere as the broadcast operand.
//
// The filter only has to be CONSERVATIVE: a sphere it rejects must be one the reference
// skips ("disc < 0", Renderer.cu:267); everything it keeps is re-decided by exact_test with
// the reference's literal sequence, in ascending index order, so (tmin, closest) is exactly
// Renderer::traceRay's whatever the filter lets through. That freedom is used to take the
// ray origin out of the per-sphere work. With o.d, o.o and 2o formed once per ray,
//
//   disc/4 = (oc.d)^2 - a (oc.oc - r^2),  oc = o - c
//          = (o.d - c.d)^2 - a (|c|^2 - r^2 - 2 o.c) - a |o|^2
//
//   hb' = fma(-cz,dz, fma(-cx,dx, fma(-cy,dy, o.d)))          3 FFMA2
//   t   = fma(-cz,2oz, fma(-cx,2ox, fma(-cy,2oy, kk)))        3 FFMA2   kk = |c|^2 - r^2 - margin_sphere (per sphere, from the pack kernel)
//   w   = fma(t, -a, g)                                       1 FFMA2   g  = -a |o|^2 + margin_ray        (per ray)
//   pre = fma(hb', hb', w)                                    1 FFMA2
//   mask = (mask << 1) | signbit(pre)                         2 SHF     (ALU pipe)
//
// = 1 LDS.128 + 8 packed FP + 2 ALU = 11 instructions for two tests, 19 issue cycles (a packed op
// holds the issue port of its sub-partition for two cycles, tools/ubench_fp32.cu) against 27 for
// the form that follows the reference op by op (3 FADD2 for oc, 6 for the two dot products, 2 FFMA
// for -r*r, 2 FFMA2): the reference's 19 algorithmic flop per test are decided with 16.
//
// Rounding. This chain cancels |c|^2 + |o|^2 - 2 o.c instead of forming o - c first, so its
// absolute error grows with |o|^2 + |c|^2 (times a) instead of |o - c|^2. With u = 2^-24, worst case over
// both chains (the reference's own rounding included, since it is ITS computed disc that decides):
//   | pre - disc_ref/4 |  <=  u a (78 |o|^2 + 80 |c|^2 + 15 r^2)        (derivation: DESIGN.md section 3.2)
// The margins make the test one-sided by more than that: margin_sphere = 2^-17 (|c|^2 + r^2) and
// margin_ray = 2^-17 a |o|^2 (+ 2^-100 for flush-to-zero corner cases): 128 u each. A set sign bit
// therefore proves disc_ref < 0. The price is a sphere radius that looks larger by
// 2^-18 (|o|^2 + |c|^2) / r^2 relative: +0.6 % candidates on the config-3 scene, +2.5 % on config 4
// (measured, tests/test_filter_margin.py, which also searches for false negatives with grazing rays:
// none at 2^-20, the first at 2^-22).
// A NaN pre has a clear sign bit (the device's canonical NaN), so it stays a candidate; the pack kernel
// replaces a non-finite kk by -FLT_MAX (always a candidate).
// ---------------------------------------------------------------------------
constexpr float kFilterMargin = 7.62939453125e-06f; // 2^-17

struct RayPair
{
    f32x2 dx, dy, dz;     // direction
    f32x2 o2x, o2y, o2z;  // 2 * origin
    f32x2 od;             // o . d
    f32x2 na;             // -a, a = d . d
    f32x2 g;              // -a |o|^2 + 2^-17 a |o|^2 + 2^-100
};

ATX_DEV void ray_pair_lane(float ox, float oy, float oz, float dx, float dy, float dz, float a, float& od, float& g)
{
    od = fdot3(ox, oy, oz, dx, dy, dz);
    const float oo = fdot3(ox, oy, oz, ox, oy, oz);
    const float mray = fmul(a, fmul(oo, kFilterMargin));
    g = fadd(ffma(fneg(a), oo, mray), 7.888609052210118e-31f); // + 2^-100
}

// sp = (-cx, -cy, -cz, kk)
ATX_DEV void filter_sphere(const float4 sp, const RayPair& r, uint32_t& m0, uint32_t& m1)
{
    const f32x2 cx = pk2(sp.x, sp.x), cy = pk2(sp.y, sp.y), cz = pk2(sp.z, sp.z);
    const f32x2 hb = ffma2(cz, r.dz, ffma2(cx, r.dx, ffma2(cy, r.dy, r.od)));
    const f32x2 t = ffma2(cz, r.o2z, ffma2(cx, r.o2x, ffma2(cy, r.o2y, pk2(sp.w, sp.w))));
    const f32x2 w = ffma2(t, r.na, r.g);
    const f32x2 pre = ffma2(hb, hb, w);
    m0 = shift_in_sign(m0, lo2(pre));
    m1 = shift_in_sign(m1, hi2(pre));
}

// Renderer::rayHit (Renderer.cu:396-409): p = (o - c) + d*t (fma); n = p * rsqrt(dot(p,p)); wp = p + c
// with sp = (-cx, -cy, -cz, r): o - c == o + (-c) and p + c == p - (-c) exactly.
ATX_DEV void hit_record(const float4 sp, float ox, float oy, float oz, float dx, float dy, float dz, float t,
                        V3& wp, V3& n)
{
    const float px = ffma(t, dx, fadd(ox, sp.x));
    const float py = ffma(t, dy, fadd(oy, sp.y));
    const float pz = ffma(t, dz, fadd(oz, sp.z));
    const float inv = frsqrt_approx(fdot3(px, py, pz, px, py, pz));
    n = { fmul(px, inv), fmul(py, inv), fmul(pz, inv) };
    wp = { fsub(px, sp.x), fsub(py, sp.y), fsub(pz, sp.z) };
}

// ---------------------------------------------------------------------------
// BRDF::cookTorrance (BRDF.cu:9-34) with fresnelSchlick (:36-40), distributionGGX
// (:42-53), geometrySchlickGGX/geometrySmith (:55-70), as compiled: returns
// (kD*diffuse + specular) * NdotL. m0..m3 are the packed material rows.
// ---------------------------------------------------------------------------
ATX_DEV V3 cook_torrance(const float4 m0, const float4 m1, const float4 m2, const float4 m3,
                         const V3 N, const V3 V, const V3 L)
{
    // H = normalize(V + L)
    const float hx0 = fadd(V.x, L.x), hy0 = fadd(V.y, L.y), hz0 = fadd(V.z, L.z);
    const float hinv = frsqrt_approx(fdot3(hx0, hy0, hz0, hx0, hy0, hz0));
    const float Hx = fmul(hx0, hinv), Hy = fmul(hy0, hinv), Hz = fmul(hz0, hinv);

    float NdotL = fdot3(N.x, N.y, N.z, L.x, L.y, L.z);
    NdotL = NdotL < 1e-7f ? 1e-7f : NdotL;
    float NdotV = fdot3(N.x, N.y, N.z, V.x, V.y, V.z);
    NdotV = NdotV < 1e-7f ? 1e-7f : NdotV;
    float NdotH = fdot3(Hx, Hy, Hz, N.x, N.y, N.z);
    NdotH = NdotH < 0.0f ? 0.0f : NdotH;
    float VdotH = fdot3(Hx, Hy, Hz, V.x, V.y, V.z);
    VdotH = VdotH < 0.0f ? 0.0f : VdotH;
    VdotH = VdotH > 1.0f ? 1.0f : VdotH;

    // F = F0 + (1 - F0) * powf(1 - VdotH, 5)   ->  ex2(5 * lg2(x))
    const float pw = fex2_approx(fmul(flg2_approx(fsub(1.0f, VdotH)), 5.0f));
    const float Fx = ffma(m2.x, pw, m1.x);
    const float Fy = ffma(m2.y, pw, m1.y);
    const float Fz = ffma(m2.z, pw, m1.z);

    // D = a2 / (pi * d * d), d = NdotH^2 * (a2 - 1) + 1
    const float dd = ffma(m3.y, fmul(NdotH, NdotH), 1.0f);
    const float D = fdiv_approx(m3.x, fmul(dd, fmul(dd, 3.14159274f)));

    // G = G1(NdotV) * G1(NdotL), G1(x) = x / (x*(1-k) + k)
    const float gV = fdiv_approx(NdotV, ffma(NdotV, m3.w, m3.z));
    const float gL = fdiv_approx(NdotL, ffma(NdotL, m3.w, m3.z));
    const float G = fmul(gV, gL);

    // kD = (1 - F) * (1 - metallic)
    const float omFx = fsub(1.0f, Fx), omFy = fsub(1.0f, Fy), omFz = fsub(1.0f, Fz);
    const float kDx = fmul(m2.w, omFx), kDy = fmul(m2.w, omFy), kDz = fmul(m2.w, omFz);

    // specular = D*G*F / (4*NdotL*NdotV + 0.001)
    const float DG = fmul(D, G);
    const float den = ffma(fmul(NdotL, 4.0f), NdotV, 0.001f);
    const float sx = fdiv_approx(fmul(Fx, DG), den);
    const float sy = fdiv_approx(fmul(Fy, DG), den);
    const float sz = fdiv_approx(fmul(Fz, DG), den);

    // diffuse = (1 - F) * albedo / pi
    const float dfx = fdiv_approx(fmul(omFx, m0.x), 3.14159274f);
    const float dfy = fdiv_approx(fmul(omFy, m0.y), 3.14159274f);
    const float dfz = fdiv_approx(fmul(omFz, m0.z), 3.14159274f);

    return { fmul(NdotL, ffma(kDx, dfx, sx)), fmul(NdotL, ffma(kDy, dfy, sy)), fmul(NdotL, ffma(kDz, dfz, sz)) };
}

// Tangent frame + combination shared by both samplers (BRDF.cu:83-92 / :107-116):
// T from the larger of |N.x|,|N.y|; B = cross(N,T) (fma(a,b,-(c*d)) as ptxas contracts it); result = x*T + y*B + z*N
// as fma(z, N, fma(x, T, y*B)).
ATX_DEV void tangent_frame(const V3 N, V3& T, V3& B)
{
    // the two branches of the reference differ only in which component of N pairs with N.z;
    // selecting the operands instead of branching runs the same instructions on the same values
    const bool xMajor = fabs_(N.x) > fabs_(N.y);
    const float major = xMajor ? N.x : N.y;
    const float s = fsqrt_approx(ffma(N.z, N.z, fmul(major, major)));
    const float nz = fneg(N.z);
    T.x = fdiv_approx(xMajor ? nz : 0.0f, s);
    T.y = fdiv_approx(xMajor ? 0.0f : nz, s);
    T.z = fdiv_approx(major, s);
    // cross(N, T): ptxas fuses the first product of each component (sampler SASS 0x380-0x3d0)
    B.x = ffma(N.y, T.z, fneg(fmul(T.y, N.z)));
    B.y = ffma(T.x, N.z, fneg(fmul(N.x, T.z)));
    B.z = ffma(N.x, T.y, fneg(fmul(N.y, T.x)));
}

// result = x*T + y*B + z*N as fma(z, N, fma(x, T, y*B))
ATX_DEV V3 frame_combine(const V3 N, const V3 T, const V3 B, float x, float y, float z)
{
    return { ffma(z, N.x, ffma(x, T.x, fmul(y, B.x))),
             ffma(z, N.y, ffma(x, T.y, fmul(y, B.y))),
             ffma(z, N.z, ffma(x, T.z, fmul(y, B.z))) };
}

ATX_DEV V3 to_world(const V3 N, float x, float y, float z)
{
    V3 T, B;
    tangent_frame(N, T, B);
    return frame_combine(N, T, B, x, y, z);
}

// BRDF::sampleHemisphereCosineWeighted (BRDF.cu:72-93) and BRDF::sampleGGX (BRDF.cu:95-117) in one
// body: both draw u1, u2, build (r cos phi, r sin phi, z) and go through the same tangent frame; only
// r and z differ, so a warp with both kinds of material shares everything but two short selects.
//   cosine: r = sqrt(u1), z = sqrt(1 - u1)
//   GGX   : z = sqrt((1 - u1) / fma(ggxT, u1, 1)), r = sqrt(fma(-z, z, 1)), ggxT = fma(a, a, -1), a = roughness^2
//           (1 - cos^2 is contracted by ptxas, sampler SASS 0x2d0). The half-vector itself is returned as
//           the new direction (reference quirk Q-ggx).
ATX_DEV void sample_local(bool ggx, float ggxT, uint32_t& seed, float& x, float& y, float& z)
{
    const float u1 = pcg_float(seed);
    const float u2 = pcg_float(seed);
    const float omu = fsub(1.0f, u1);
    float r;
    if (ggx)
    {
        z = fsqrt_approx(fdiv_approx(omu, ffma(ggxT, u1, 1.0f)));
        r = fsqrt_approx(ffma(fneg(z), z, 1.0f));
    }
    else
    {
        r = fsqrt_approx(u1);
        z = fsqrt_approx(omu);
    }
    const float phi = fmul(u2, 6.28318548f);
    x = fmul(r, fcos_approx(phi));
    y = fmul(r, fsin_approx(phi));
}

ATX_DEV V3 sample_direction(const V3 N, bool ggx, float ggxT, uint32_t& seed)
{
    float x, y, z;
    sample_local(ggx, ggxT, seed, x, y, z);
    return to_world(N, x, y, z);
}

// kernelRender's display pack (Renderer.cu:166-168; colorUtils::vec4ToRGBA, Renderer.h:70-78)
ATX_DEV uint32_t pack_rgba8(const float4 acc, float divisor)
{
    float c[4] = { fdiv_approx(acc.x, divisor), fdiv_approx(acc.y, divisor), fdiv_approx(acc.z, divisor),
                   fdiv_approx(acc.w, divisor) };
    uint32_t u[4];
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        float v = c[i] < 0.0f ? 0.0f : c[i];
        v = v > 1.0f ? 1.0f : v;
        u[i] = f32_to_u32_rz_ftz(fmul(v, 255.0f)) & 0xFFu;
    }
    return (u[3] << 24) | (u[2] << 16) | (u[1] << 8) | u[0];
}

// ---------------------------------------------------------------------------
// One path of Renderer::perPixel (Renderer.cu:287-387), cut at its two traceRay calls
// so that every kernel variant (while-while, two-slot packed, wavefront) runs the same
// code between traces: path_begin -> [trace] -> path_hit -> [shadow trace] ->
// path_shadow -> path_bounce -> [trace] ... The functions return what the caller must
// do next; none of them loops.
// ---------------------------------------------------------------------------
struct PathState
{
    float ox, oy, oz, dx, dy, dz; // ray in flight
    float cr, cg, cb;             // color
    float tx, ty, tz;             // throughput
    uint32_t seed;
    int bounce;
    // carried from the closest-hit half to the shadow half of a bounce
    V3 N, V;
    float dist2;
    int matIndex;
    uint32_t lightIndex;
};

ATX_DEV void path_begin(PathState& s, const float camPos[3], const V3 d0, uint32_t pixel, uint32_t frame)
{
    s.ox = camPos[0]; s.oy = camPos[1]; s.oz = camPos[2];
    s.dx = d0.x; s.dy = d0.y; s.dz = d0.z;
    s.cr = s.cg = s.cb = 0.0f;
    s.tx = s.ty = s.tz = 1.0f;
    s.seed = pixel * frame; // Renderer.cu:300-301 (bounce 0 adds 0, :306)
    s.bounce = 0;
}

// closest-hit ray missed everything (Renderer.cu:309-318): the path ends
ATX_DEV void path_miss(const RenderParams& p, PathState& s)
{
    if (p.skyLight)
    {
        s.cr = ffma(s.tx, 0.6f, s.cr);
        s.cg = ffma(s.ty, 0.7f, s.cg);
        s.cb = ffma(s.tz, 0.9f, s.cb);
    }
}

// closest-hit ray hit sphere `closest` at `t` (Renderer.cu:320-349). sp = that sphere's record.
// Returns true when a shadow ray is now in flight (s.o/s.d), false when there are no lights
// (the caller goes straight to path_bounce).
ATX_DEV bool path_hit(const RenderParams& p, PathState& s, const float4 sp, int closest, float t)
{
    V3 wp;
    hit_record(sp, s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, t, wp, s.N);
    s.matIndex = __ldg(p.sphMat + closest);
    const float4 m4 = __ldg(p.mats + kMatStride * s.matIndex + 4);
    if (m4.w > 0.0f) // emission (Renderer.cu:329-333)
    {
        s.cr = ffma(s.tx, m4.x, s.cr);
        s.cg = ffma(s.ty, m4.y, s.cg);
        s.cb = ffma(s.tz, m4.z, s.cb);
    }
    // next origin == shadow origin: pos + N*1e-4 (Renderer.cu:348, :372), an fma
    const float nox = ffma(s.N.x, 0.0001f, wp.x);
    const float noy = ffma(s.N.y, 0.0001f, wp.y);
    const float noz = ffma(s.N.z, 0.0001f, wp.z);
    bool shadow = false;
    if (p.nLights > 0)
    {
        // light pick reuses the un-advanced seed (Renderer.cu:340)
        s.lightIndex = p.nLights == 1u ? 0u : pcg_hash(s.seed) % p.nLights; // x % 1 == 0: skip the hash
        const float4 lp = __ldg(p.lights + kLightStride * s.lightIndex);
        const float lx = fsub(lp.x, wp.x), ly = fsub(lp.y, wp.y), lz = fsub(lp.z, wp.z);
        s.dist2 = fdot3(lx, ly, lz, lx, ly, lz);
        const float inv = frsqrt_approx(s.dist2);
        s.V = { fsub(0.0f, s.dx), fsub(0.0f, s.dy), fsub(0.0f, s.dz) }; // V = -ray.direction (Renderer.cu:359)
        s.dx = fmul(lx, inv); s.dy = fmul(inv, ly); s.dz = fmul(inv, lz);
        shadow = true;
    }
    s.ox = nox; s.oy = noy; s.oz = noz;
    return shadow;
}

// shadow ray result (Renderer.cu:351-368): occluded iff t > 0 && t*t < dist2
ATX_DEV void path_shadow(const RenderParams& p, PathState& s, int closest, float tmin)
{
    const float ts = closest < 0 ? -1.0f : tmin;
    if (!(ts > 0.0f && fmul(ts, ts) < s.dist2))
    {
        const float4 m0 = __ldg(p.mats + kMatStride * s.matIndex + 0);
        const float4 m1 = __ldg(p.mats + kMatStride * s.matIndex + 1);
        const float4 m2 = __ldg(p.mats + kMatStride * s.matIndex + 2);
        const float4 m3 = __ldg(p.mats + kMatStride * s.matIndex + 3);
        const V3 L = { s.dx, s.dy, s.dz };
        const V3 c = cook_torrance(m0, m1, m2, m3, s.N, s.V, L);
        const float4 le = __ldg(p.lights + kLightStride * s.lightIndex + 1);
        // color += emission * specular * throughput / pdf(=1)   (Renderer.cu:362-367): ptxas folds
        // the division by 1 and contracts the last product into the sum (kernelRender SASS 0x32d0-0x3350)
        s.cr = ffma(fmul(le.x, c.x), s.tx, s.cr);
        s.cg = ffma(fmul(le.y, c.y), s.ty, s.cg);
        s.cb = ffma(fmul(le.z, c.z), s.tz, s.cb);
    }
}

// throughput, Russian roulette, next direction (Renderer.cu:371-384). Returns true when the
// path ends (roulette or bounce limit); otherwise the next closest-hit ray is in s.o/s.d.
ATX_DEV bool path_bounce(const RenderParams& p, PathState& s)
{
    const float4 m0 = __ldg(p.mats + kMatStride * s.matIndex + 0);
    const float4 m1 = __ldg(p.mats + kMatStride * s.matIndex + 1);
    s.tx = fmul(s.tx, m0.x); s.ty = fmul(s.ty, m0.y); s.tz = fmul(s.tz, m0.z);
    const float len = fsqrt_approx(fdot3(s.tx, s.ty, s.tz, s.tx, s.ty, s.tz));
    const float pr = fmax_(fmin_(len, 1.0f), 0.1f);
    if (pcg_float(s.seed) > pr)
        return true;
    s.tx = fdiv_approx(s.tx, pr); s.ty = fdiv_approx(s.ty, pr); s.tz = fdiv_approx(s.tz, pr);
    const bool ggx = m1.w > 0.0f; // metallic > 0 (Renderer.cu:381-384)
    const float ggxT = ggx ? __ldg(p.mats + kMatStride * s.matIndex + 5).x : 0.0f;
    const V3 nd = sample_direction(s.N, ggx, ggxT, s.seed);
    s.dx = nd.x; s.dy = nd.y; s.dz = nd.z;
    s.bounce++;
    if (s.bounce >= p.maxBounces)
        return true;
    s.seed += static_cast<uint32_t>(s.bounce); // Renderer.cu:306
    return false;
}

// the finished sums of a pixel: st.global.v4.f32 into the local image and, with an image-tile split, into every
// peer's image over NVLink (the transfer rides along with the rendering, pixel by pixel)
ATX_DEV void store_pixel(const RenderParams& p, uint32_t pixel, const float4 acc)
{
    p.accum[pixel] = acc;
    // (not unrolled: seven predicated peer stores at each of the store sites cost the warp-queue form 5 % on one GPU, where nPush is 0)
#pragma unroll 1
    for (uint32_t r = 0; r < p.nPush; r++)
        p.push[r][pixel] = acc;
}

// accumulation[p] += vec4(color, 1)   (Renderer.cu:165, :386)
ATX_DEV void accumulate_sample(float4& acc, const PathState& s)
{
    acc.x = fadd(s.cr, acc.x); acc.y = fadd(s.cg, acc.y); acc.z = fadd(s.cb, acc.z);
    acc.w = fadd(acc.w, 1.0f);
}

} // namespace atxk
