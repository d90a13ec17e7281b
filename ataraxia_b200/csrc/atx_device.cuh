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
//   spheres : float4 (cx, cy, cz, radius)   SoA-of-float4, one LDS.128 / LDG.128 per test
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
    const float4* spheres;
    const int32_t* sphMat;
    const float4* mats;
    const float4* lights;
    float4* accum;
    uint32_t* rgba;
    unsigned long long* counters; // [0]=paths [1]=rays
    CameraParams cam;
};

struct V3 { float x, y, z; };

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

// sp = (cx, cy, cz, radius)
ATX_DEV void intersect_sphere(const float4 sp, int index, float ox, float oy, float oz, float dx, float dy, float dz,
                              const RayConst& k, float& tmin, int& closest)
{
    const float ocx = fsub(ox, sp.x);
    const float ocy = fsub(oy, sp.y);
    const float ocz = fsub(oz, sp.z);
    const float hb = fdot3(ocx, ocy, ocz, dx, dy, dz);
    const float cc = ffma(fneg(sp.w), sp.w, fdot3(ocx, ocy, ocz, ocx, ocy, ocz));
    const float pre = ffma(hb, hb, fneg(fmul(k.a, cc)));
    if (!(pre < 0.0f))
    {
        // literal reference sequence (kernelRender SASS, Renderer.cu:263-278)
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
}

// Renderer::rayHit (Renderer.cu:396-409): p = (o - c) + d*t (fma); n = p * rsqrt(dot(p,p)); wp = p + c
ATX_DEV void hit_record(const float4 sp, float ox, float oy, float oz, float dx, float dy, float dz, float t,
                        V3& wp, V3& n)
{
    const float px = ffma(t, dx, fsub(ox, sp.x));
    const float py = ffma(t, dy, fsub(oy, sp.y));
    const float pz = ffma(t, dz, fsub(oz, sp.z));
    const float inv = frsqrt_approx(fdot3(px, py, pz, px, py, pz));
    n = { fmul(px, inv), fmul(py, inv), fmul(pz, inv) };
    wp = { fadd(px, sp.x), fadd(py, sp.y), fadd(pz, sp.z) };
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
ATX_DEV V3 to_world(const V3 N, float x, float y, float z)
{
    float Tx, Ty, Tz;
    if (fabs_(N.x) > fabs_(N.y))
    {
        const float s = fsqrt_approx(ffma(N.z, N.z, fmul(N.x, N.x)));
        Tx = fdiv_approx(fneg(N.z), s);
        Ty = fdiv_approx(0.0f, s);
        Tz = fdiv_approx(N.x, s);
    }
    else
    {
        const float s = fsqrt_approx(ffma(N.z, N.z, fmul(N.y, N.y)));
        Tx = fdiv_approx(0.0f, s);
        Ty = fdiv_approx(fneg(N.z), s);
        Tz = fdiv_approx(N.y, s);
    }
    // cross(N, T): ptxas fuses the first product of each component (sampler SASS 0x380-0x3d0)
    const float Bx = ffma(N.y, Tz, fneg(fmul(Ty, N.z)));
    const float By = ffma(Tx, N.z, fneg(fmul(N.x, Tz)));
    const float Bz = ffma(N.x, Ty, fneg(fmul(N.y, Tx)));
    return { ffma(z, N.x, ffma(x, Tx, fmul(y, Bx))),
             ffma(z, N.y, ffma(x, Ty, fmul(y, By))),
             ffma(z, N.z, ffma(x, Tz, fmul(y, Bz))) };
}

// BRDF::sampleHemisphereCosineWeighted (BRDF.cu:72-93)
ATX_DEV V3 sample_cosine(const V3 N, uint32_t& seed)
{
    const float u1 = pcg_float(seed);
    const float u2 = pcg_float(seed);
    const float r = fsqrt_approx(u1);
    const float theta = fmul(u2, 6.28318548f);
    const float x = fmul(r, fcos_approx(theta));
    const float y = fmul(r, fsin_approx(theta));
    const float z = fsqrt_approx(fsub(1.0f, u1));
    return to_world(N, x, y, z);
}

// BRDF::sampleGGX (BRDF.cu:95-117); ggxT = fma(a, a, -1) with a = roughness^2. The
// half-vector itself is returned as the new direction (reference quirk Q-ggx).
ATX_DEV V3 sample_ggx(const V3 N, float ggxT, uint32_t& seed)
{
    const float u1 = pcg_float(seed);
    const float u2 = pcg_float(seed);
    const float cosT = fsqrt_approx(fdiv_approx(fsub(1.0f, u1), ffma(ggxT, u1, 1.0f)));
    const float sinT = fsqrt_approx(ffma(fneg(cosT), cosT, 1.0f)); // 1 - cos^2, contracted by ptxas (SASS 0x2d0)
    const float phi = fmul(u2, 6.28318548f);
    const float x = fmul(sinT, fcos_approx(phi));
    const float y = fmul(sinT, fsin_approx(phi));
    return to_world(N, x, y, cosT);
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

} // namespace atxk
