// oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement ("port") of the reference's per-pixel path-tracing hot path
// (1neskk/Ataraxia @ /root/reference). Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; the
// product (ataraxia_b200/) never links, imports or calls it.
//
// What it restates, function by function (reference file:line in each comment):
//   Random::PcgHash / PcgFloat          Core/include/Random.h:59-70
//   Camera matrices + ray table         Engine/src/Camera.cpp:134-195 (glm 1.0.2 order)
//   Camera::onUpdate (scripted input)   Engine/src/Camera.cpp:30-108
//   SceneNode transform / flatten       Engine/src/SceneNode.cpp:42-59, Renderer.cu:67-96
//   Renderer::traceRay / rayHit         Engine/src/Renderer.cu:251-285, :396-409
//   Renderer::perPixel                  Engine/src/Renderer.cu:287-387
//   BRDF::*                             Engine/src/BRDF.cu:9-117
//   kernelRender accumulate + pack      Engine/src/Renderer.cu:165-168, Renderer.h:70-78
//
// Arithmetic model: the reference's device code compiled for the HOST (the
// north_star's "per-pixel shading compiled host-side" baseline): IEEE float,
// no FMA contraction, libm sqrtf/sinf/cosf/powf/tanf, glm's evaluation order
// ((x*x' + y*y') + z*z' for dot, v * (1/sqrt(d)) for normalize, ...).
//
// PINNING: this port is checked bit for bit against oracle/_ref/libref_cpu.so —
// the reference's own unmodified Renderer.cu/BRDF.cu/Camera.cpp/SceneNode.cpp
// compiled for the host (oracle/ref/Makefile) — and against the golden vectors in
// tests/golden/ that were generated from it (tests/golden/make_golden.py). The
// reference's device build uses approximate MUFU intrinsics (-use_fast_math), so
// GPU bit-exactness is pinned separately against oracle/_ref/ref_headless (the
// reference's CUDA renderer) and its committed golden vectors; against this CPU
// port the GPU path agrees to the tolerance stated in tests/.
//
// Build: g++ -O2 -ffp-contract=off -fno-fast-math -shared -fPIC oracle.cpp
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace
{
struct v3 { float x, y, z; };
struct v4 { float x, y, z, w; };
struct m4 { v4 c[4]; };

inline v3 operator+(v3 a, v3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline v3 operator-(v3 a, v3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline v3 operator*(v3 a, v3 b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
inline v3 operator*(v3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline v3 operator*(float s, v3 a) { return { s * a.x, s * a.y, s * a.z }; }
inline v3 operator/(v3 a, float s) { return { a.x / s, a.y / s, a.z / s }; }
inline v3 operator-(v3 a) { return { -a.x, -a.y, -a.z }; }
inline v4 operator+(v4 a, v4 b) { return { a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w }; }
inline v4 operator-(v4 a, v4 b) { return { a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w }; }
inline v4 operator*(v4 a, v4 b) { return { a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w }; }
inline v4 operator*(v4 a, float s) { return { a.x * s, a.y * s, a.z * s, a.w * s }; }

// glm/detail/func_geometric.inl:48-54
inline float dot(v3 a, v3 b) { v3 t = a * b; return t.x + t.y + t.z; }
// glm/detail/func_geometric.inl:98-105 with func_exponential.inl:134-139
inline v3 normalize(v3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }
inline float length(v3 v) { return std::sqrt(dot(v, v)); }
// glm/detail/func_geometric.inl:73-83
inline v3 cross(v3 x, v3 y) { return { x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y }; }
inline float gmax(float x, float y) { return (x < y) ? y : x; }  // glm::max, func_common.inl
inline float gmin(float x, float y) { return (y < x) ? y : x; }  // glm::min
inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }

// glm/detail/type_mat4x4.inl:562-573
inline v4 mul(const m4& m, v4 v)
{
    v4 Mul0 = m.c[0] * v4{ v.x, v.x, v.x, v.x };
    v4 Mul1 = m.c[1] * v4{ v.y, v.y, v.y, v.y };
    v4 Add0 = Mul0 + Mul1;
    v4 Mul2 = m.c[2] * v4{ v.z, v.z, v.z, v.z };
    v4 Mul3 = m.c[3] * v4{ v.w, v.w, v.w, v.w };
    v4 Add1 = Mul2 + Mul3;
    return Add0 + Add1;
}

// glm/detail/type_mat4x4.inl mul4x4<.., false>
inline m4 mul(const m4& a, const m4& b)
{
    m4 r;
    for (int j = 0; j < 4; j++)
    {
        v4 t = a.c[0] * b.c[j].x;
        t = t + a.c[1] * b.c[j].y;
        t = t + a.c[2] * b.c[j].z;
        t = t + a.c[3] * b.c[j].w;
        r.c[j] = t;
    }
    return r;
}

inline float& el(m4& m, int c, int r) { return (&m.c[c].x)[r]; }
inline float el(const m4& m, int c, int r) { return (&m.c[c].x)[r]; }

inline m4 identity()
{
    m4 m;
    std::memset(&m, 0, sizeof(m));
    m.c[0].x = m.c[1].y = m.c[2].z = m.c[3].w = 1.0f;
    return m;
}

// glm/detail/func_matrix.inl:388-446
m4 inverse(const m4& m)
{
#define M(c, r) el(m, c, r)
    float Coef00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
    float Coef02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
    float Coef03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
    float Coef04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    float Coef06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3);
    float Coef07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
    float Coef08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
    float Coef10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
    float Coef11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
    float Coef12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
    float Coef14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3);
    float Coef15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
    float Coef16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2);
    float Coef18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
    float Coef19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
    float Coef20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
    float Coef22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1);
    float Coef23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
    v4 Fac0{ Coef00, Coef00, Coef02, Coef03 }, Fac1{ Coef04, Coef04, Coef06, Coef07 };
    v4 Fac2{ Coef08, Coef08, Coef10, Coef11 }, Fac3{ Coef12, Coef12, Coef14, Coef15 };
    v4 Fac4{ Coef16, Coef16, Coef18, Coef19 }, Fac5{ Coef20, Coef20, Coef22, Coef23 };
    v4 Vec0{ M(1, 0), M(0, 0), M(0, 0), M(0, 0) }, Vec1{ M(1, 1), M(0, 1), M(0, 1), M(0, 1) };
    v4 Vec2{ M(1, 2), M(0, 2), M(0, 2), M(0, 2) }, Vec3{ M(1, 3), M(0, 3), M(0, 3), M(0, 3) };
#undef M
    v4 Inv0 = Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2;
    v4 Inv1 = Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4;
    v4 Inv2 = Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5;
    v4 Inv3 = Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5;
    v4 SignA{ +1, -1, +1, -1 }, SignB{ -1, +1, -1, +1 };
    m4 Inverse;
    Inverse.c[0] = Inv0 * SignA;
    Inverse.c[1] = Inv1 * SignB;
    Inverse.c[2] = Inv2 * SignA;
    Inverse.c[3] = Inv3 * SignB;
    v4 Row0{ Inverse.c[0].x, Inverse.c[1].x, Inverse.c[2].x, Inverse.c[3].x };
    v4 Dot0 = m.c[0] * Row0;
    float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
    float OneOverDeterminant = 1.0f / Dot1;
    m4 r;
    for (int j = 0; j < 4; j++) r.c[j] = Inverse.c[j] * OneOverDeterminant;
    return r;
}

// Random.h:59-64
inline uint32_t pcg_hash(uint32_t seed)
{
    uint32_t state = seed * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}
// Random.h:66-70
inline float pcg_float(uint32_t& seed)
{
    seed = pcg_hash(seed);
    return static_cast<float>(seed) / static_cast<float>(UINT32_MAX);
}

struct Sphere { float cx, cy, cz, radius; int32_t id; };                         // SceneNode.h:11-21
struct Material { v3 albedo; float roughness, metallic; v3 F0; v3 emissionColor; float emissionIntensity; int32_t id; }; // Scene.h:28-47
struct Light { v3 position, color; float intensity; };                             // Scene.h:17-26
static_assert(sizeof(Sphere) == 20 && sizeof(Material) == 52 && sizeof(Light) == 28, "layout");

struct Hit { float t; v3 pos, normal; int id; };

// Renderer.cu:251-285 + :389-409
Hit trace_ray(v3 o, v3 d, const Sphere* s, size_t n)
{
    int closest = -1;
    float tmin = FLT_MAX;
    for (size_t i = 0; i < n; i++)
    {
        v3 center{ s[i].cx, s[i].cy, s[i].cz };
        v3 oc = o - center;
        const float a = dot(d, d);
        const float b = 2.0f * dot(oc, d);
        const float c = dot(oc, oc) - s[i].radius * s[i].radius;
        const float disc = b * b - 4 * a * c;
        if (disc < 0.0f)
            continue;
        float t0 = (-b - std::sqrt(disc)) / (2.0f * a);
        float t1 = (-b + std::sqrt(disc)) / (2.0f * a);
        const float t = t0 < t1 ? t0 : t1;
        if (t > 0.0f && t < tmin)
        {
            tmin = t;
            closest = static_cast<int>(i);
        }
    }
    Hit h;
    h.id = closest;
    if (closest < 0)
    {
        h.t = -1.0f;
        h.pos = h.normal = v3{ 0, 0, 0 };
        return h;
    }
    h.t = tmin;
    v3 center{ s[closest].cx, s[closest].cy, s[closest].cz };
    v3 origin = o - center;
    h.pos = origin + d * tmin;
    h.normal = normalize(h.pos);
    h.pos = h.pos + center;
    return h;
}

// BRDF.cu:36-40
v3 fresnel_schlick(v3 F0, float cosTheta)
{
    cosTheta = gclamp(cosTheta, 0.0f, 1.0f);
    return F0 + (v3{ 1.0f - F0.x, 1.0f - F0.y, 1.0f - F0.z }) * powf(1.0f - cosTheta, 5.0f);
}
// BRDF.cu:42-53
float distribution_ggx(float NdotH, float roughness)
{
    const float a = roughness * roughness;
    const float a2 = a * a;
    float NdotH2 = NdotH * NdotH;
    float denom = (NdotH2 * (a2 - 1.0f) + 1.0f);
    denom = 3.14159265358979323846264338327950288f * denom * denom;
    return a2 / denom;
}
// BRDF.cu:55-63
float geometry_schlick_ggx(float NdotV, float roughness)
{
    const float r = (roughness + 1.0f);
    const float k = (r * r) / 8.0f;
    return NdotV / (NdotV * (1.0f - k) + k);
}
// BRDF.cu:9-34
v3 cook_torrance(v3 albedo, v3 F0, float metallic, float roughness, v3 N, v3 V, v3 L)
{
    v3 H = normalize(V + L);
    float NdotL = gmax(dot(N, L), 0.0000001f);
    float NdotV = gmax(dot(N, V), 0.0000001f);
    float NdotH = gmax(dot(N, H), 0.0f);
    float VdotH = gmax(dot(V, H), 0.0f);
    v3 F = fresnel_schlick(F0, VdotH);
    float D = distribution_ggx(NdotH, roughness);
    float G = geometry_schlick_ggx(NdotV, roughness) * geometry_schlick_ggx(NdotL, roughness);
    v3 kD = v3{ 1.0f, 1.0f, 1.0f } - F;
    kD = kD * (1.0f - metallic);
    v3 num = (D * G) * F;
    float denom = 4.0f * NdotL * NdotV + 0.001f;
    v3 specular = num / denom;
    const float pi = 3.14159265358979323846264338327950288f;
    v3 diffuse = ((v3{ 1.0f - F.x, 1.0f - F.y, 1.0f - F.z }) * albedo) / pi;
    return (kD * diffuse + specular) * NdotL;
}

// tangent frame of BRDF.cu:83-90 / :107-114
void frame(v3 N, v3& T, v3& B)
{
    if (std::fabs(N.x) > std::fabs(N.y))
        T = v3{ -N.z, 0, N.x } / std::sqrt(N.x * N.x + N.z * N.z);
    else
        T = v3{ 0, -N.z, N.y } / std::sqrt(N.y * N.y + N.z * N.z);
    B = cross(N, T);
}
// BRDF.cu:72-93
v3 sample_cosine(v3 N, uint32_t& seed)
{
    const float u1 = pcg_float(seed);
    const float u2 = pcg_float(seed);
    const float r = std::sqrt(u1);
    const float theta = 2.0f * 3.14159265358979323846264338327950288f * u2;
    const float x = r * cosf(theta);
    const float y = r * sinf(theta);
    const float z = std::sqrt(1 - u1);
    v3 T, B;
    frame(N, T, B);
    return x * T + y * B + z * N;
}
// BRDF.cu:95-117
v3 sample_ggx(v3 N, float roughness, uint32_t& seed)
{
    const float u1 = pcg_float(seed);
    const float u2 = pcg_float(seed);
    const float a = roughness * roughness;
    const float cosTheta = std::sqrt((1.0f - u1) / (1.0f + (a * a - 1.0f) * u1));
    const float sinTheta = std::sqrt(1 - cosTheta * cosTheta);
    const float phi = 2.0f * 3.14159265358979323846264338327950288f * u2;
    v3 H{ sinTheta * cosf(phi), sinTheta * sinf(phi), cosTheta };
    v3 T, B;
    frame(N, T, B);
    return H.x * T + H.y * B + H.z * N;
}

// Renderer.cu:287-387
v4 per_pixel(uint32_t pixelIndex, v3 origin, v3 direction, const Sphere* spheres, size_t nS, const Material* mats,
             size_t nM, uint32_t frameIndex, const Light* lights, size_t nL, int maxBounces, bool skyLight,
             uint64_t* rays)
{
    v3 o = origin, d = direction;
    v3 color{ 0, 0, 0 }, throughput{ 1, 1, 1 };
    uint32_t seed = pixelIndex;
    seed *= frameIndex;
    for (int i = 0; i < maxBounces; i++)
    {
        seed += i;
        Hit ht = trace_ray(o, d, spheres, nS);
        ++*rays;
        if (ht.t < 0.0f)
        {
            if (skyLight)
                color = color + v3{ 0.6f, 0.7f, 0.9f } * throughput;
            break;
        }
        const Material* mat = &mats[spheres[ht.id].id];
        if (mat->emissionIntensity > 0.0f)
            color = color + (mat->emissionColor * mat->emissionIntensity) * throughput;
        // glm::mix(F0, albedo, metallic) = F0*(1-a) + albedo*a
        v3 baseReflectivity = mat->F0 * (1.0f - mat->metallic) + mat->albedo * mat->metallic;
        if (nL > 0)
        {
            uint32_t lightIndex = pcg_hash(seed) % nL;
            const Light& sl = lights[lightIndex];
            v3 L = sl.position - ht.pos;
            float distanceSquared = dot(L, L);
            L = normalize(L);
            v3 so = ht.pos + ht.normal * 0.0001f;
            Hit sh = trace_ray(so, L, spheres, nS);
            ++*rays;
            if (sh.t > 0.0f && sh.t * sh.t < distanceSquared)
            {
            }
            else
            {
                v3 V = -d;
                v3 specular = cook_torrance(mat->albedo, baseReflectivity, mat->metallic, mat->roughness, ht.normal, V, L);
                v3 emission = sl.color * sl.intensity;
                float pdf = 1.0f;
                color = color + emission * specular * throughput / pdf;
            }
        }
        throughput = throughput * mat->albedo;
        o = ht.pos + ht.normal * 0.0001f;
        float p = gmax(0.1f, gmin(1.0f, length(throughput)));
        if (pcg_float(seed) > p)
            break;
        throughput = throughput / p;
        if (mat->metallic > 0.0f)
            d = sample_ggx(ht.normal, mat->roughness, seed);
        else
            d = sample_cosine(ht.normal, seed);
    }
    return { color.x, color.y, color.z, 1.0f };
}

template <typename F>
void parallel_rows(uint32_t height, int threads, F&& fn)
{
    int n = threads > 0 ? threads : static_cast<int>(std::thread::hardware_concurrency());
    if (n < 1) n = 1;
    if (static_cast<uint32_t>(n) > height) n = static_cast<int>(height);
    if (n <= 1) { fn(0u, height, 0); return; }
    std::vector<std::thread> pool;
    // interleaved row blocks would balance better; the reference's own threading is
    // contiguous row bands (Camera.cpp:166-194), restated here
    const uint32_t rows = height / n;
    for (int t = 0; t < n; t++)
    {
        const uint32_t y0 = t * rows, y1 = (t == n - 1) ? height : y0 + rows;
        pool.emplace_back([=, &fn]() { fn(y0, y1, t); });
    }
    for (auto& th : pool) th.join();
}
} // namespace

extern "C" {

uint32_t orc_pcg_hash(uint32_t seed) { return pcg_hash(seed); }
float orc_pcg_float(uint32_t* seed) { return pcg_float(*seed); }

// Camera::onUpdate, Camera.cpp:30-108, with the window queries replaced by arguments: keys W=1 S=2 A=4 D=8 Q=16 E=32,
// right = MouseButton::Right held, (mouseX, mouseY) = Input::GetMousePosition(). glm pieces in their own order:
// angleAxis (ext/quaternion_trigonometric.inl:30-36), cross(quat, quat) and normalize(quat)
// (ext/quaternion_geometric.inl:11-34, dot as (w*w + x*x) + (y*y + z*z), detail/type_quat.inl:17-24), quat * vec3
// (detail/type_quat.inl:359-366). Returns what onUpdate returns.
int orc_camera_update(float pos[3], float dir[3], float lastMouse[2], uint32_t keys, int right, float mouseX, float mouseY, float dt)
{
    const float dx = (mouseX - lastMouse[0]) * 0.002f, dy = (mouseY - lastMouse[1]) * 0.002f; // :33-35
    lastMouse[0] = mouseX;
    lastMouse[1] = mouseY;
    if (!right)
        return 0; // :37-41
    bool moved = false;
    v3 p{ pos[0], pos[1], pos[2] }, d{ dir[0], dir[1], dir[2] };
    const v3 up{ 0.0f, 1.0f, 0.0f };
    const v3 rightV = cross(d, up); // :47
    const float speed = 5.0f;
    if (keys & 1u) { p = p + d * speed * dt; moved = true; }        // W :51-56
    else if (keys & 2u) { p = p - d * speed * dt; moved = true; }   // S
    if (keys & 4u) { p = p - rightV * speed * dt; moved = true; }   // A :64-69
    else if (keys & 8u) { p = p + rightV * speed * dt; moved = true; }
    if (keys & 16u) { p = p - up * speed * dt; moved = true; }      // Q :77-82
    else if (keys & 32u) { p = p + up * speed * dt; moved = true; }
    if (dx != 0.0f || dy != 0.0f) // :90-100
    {
        const float yaw = dx * 0.3f, pitch = dy * 0.3f;
        struct Q { float w, x, y, z; };
        const auto angleAxis = [](float angle, v3 axis) {
            const float sn = std::sin(angle * 0.5f);
            const v3 vs = axis * sn;
            return Q{ std::cos(angle * 0.5f), vs.x, vs.y, vs.z };
        };
        const Q q1 = angleAxis(-pitch, rightV), q2 = angleAxis(-yaw, v3{ 0.0f, 1.0f, 0.0f });
        Q c{ q1.w * q2.w - q1.x * q2.x - q1.y * q2.y - q1.z * q2.z,
             q1.w * q2.x + q1.x * q2.w + q1.y * q2.z - q1.z * q2.y,
             q1.w * q2.y + q1.y * q2.w + q1.z * q2.x - q1.x * q2.z,
             q1.w * q2.z + q1.z * q2.w + q1.x * q2.y - q1.y * q2.x };
        const float len = std::sqrt((c.w * c.w + c.x * c.x) + (c.y * c.y + c.z * c.z));
        if (len <= 0.0f)
            c = Q{ 1.0f, 0.0f, 0.0f, 0.0f };
        else
        {
            const float inv = 1.0f / len;
            c = Q{ c.w * inv, c.x * inv, c.y * inv, c.z * inv };
        }
        const v3 qv{ c.x, c.y, c.z };
        const v3 uv = cross(qv, d);
        const v3 uuv = cross(qv, uv);
        d = d + ((uv * c.w) + uuv) * 2.0f;
        moved = true;
    }
    pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
    dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
    return moved ? 1 : 0;
}

// Camera.cpp:134-159 (perspectiveRH_NO: glm/ext/matrix_clip_space.inl:249-262; lookAtRH: ext/matrix_transform.inl:153-173)
void orc_camera_matrices(const float pos[3], const float dir[3], float fov, float nearClip, float farClip,
                         uint32_t width, uint32_t height, float invProj[16], float invView[16])
{
    const float aspect = static_cast<float>(width) / static_cast<float>(height);
    const float fovy = fov * static_cast<float>(0.01745329251994329576923690768489);
    const float tanHalfFovy = std::tan(fovy / 2.0f);
    m4 P;
    std::memset(&P, 0, sizeof(P));
    el(P, 0, 0) = 1.0f / (aspect * tanHalfFovy);
    el(P, 1, 1) = 1.0f / (tanHalfFovy);
    el(P, 2, 2) = -(farClip + nearClip) / (farClip - nearClip);
    el(P, 2, 3) = -1.0f;
    el(P, 3, 2) = -(2.0f * farClip * nearClip) / (farClip - nearClip);
    m4 iP = inverse(P);

    v3 eye{ pos[0], pos[1], pos[2] };
    v3 center = eye + v3{ dir[0], dir[1], dir[2] };
    v3 up{ 0.0f, 1.0f, 0.0f };
    v3 f = normalize(center - eye);
    v3 s = normalize(cross(f, up));
    v3 u = cross(s, f);
    m4 Vw = identity();
    el(Vw, 0, 0) = s.x; el(Vw, 1, 0) = s.y; el(Vw, 2, 0) = s.z;
    el(Vw, 0, 1) = u.x; el(Vw, 1, 1) = u.y; el(Vw, 2, 1) = u.z;
    el(Vw, 0, 2) = -f.x; el(Vw, 1, 2) = -f.y; el(Vw, 2, 2) = -f.z;
    el(Vw, 3, 0) = -dot(s, eye);
    el(Vw, 3, 1) = -dot(u, eye);
    el(Vw, 3, 2) = dot(f, eye);
    m4 iV = inverse(Vw);
    std::memcpy(invProj, &iP, 64);
    std::memcpy(invView, &iV, 64);
}

// Camera.cpp:161-195
void orc_ray_directions(const float invProj[16], const float invView[16], uint32_t width, uint32_t height, float* out,
                        int threads)
{
    m4 iP, iV;
    std::memcpy(&iP, invProj, 64);
    std::memcpy(&iV, invView, 64);
    parallel_rows(height, threads, [&](uint32_t y0, uint32_t y1, int) {
        for (uint32_t y = y0; y < y1; y++)
            for (uint32_t x = 0; x < width; x++)
            {
                float cx = static_cast<float>(x) / static_cast<float>(width);
                float cy = static_cast<float>(y) / static_cast<float>(height);
                cx = cx * 2.0f - 1.0f;
                cy = cy * 2.0f - 1.0f;
                v4 target = mul(iP, v4{ cx, cy, 1.0f, 1.0f });
                v3 n = normalize(v3{ target.x, target.y, target.z } / target.w);
                v4 r = mul(iV, v4{ n.x, n.y, n.z, 0.0f });
                v3 dd = normalize(v3{ r.x, r.y, r.z });
                float* o = out + 3ull * (x + static_cast<size_t>(y) * width);
                o[0] = dd.x; o[1] = dd.y; o[2] = dd.z;
            }
    });
}

// SceneNode.cpp:42-59: local = translate(I, p) * mat4_cast(q) * scale(I, s); global = parent * local
void orc_node_transform(const float parent[16], const float position[3], const float rot_xyzw[4], const float scale[3],
                        float global[16])
{
    m4 par;
    std::memcpy(&par, parent, 64);
    m4 I = identity();
    // glm/ext/matrix_transform.inl:10-15
    m4 T = I;
    T.c[3] = I.c[0] * position[0] + I.c[1] * position[1] + I.c[2] * position[2] + I.c[3];
    // glm/gtc/quaternion.inl:47-72
    const float qx = rot_xyzw[0], qy = rot_xyzw[1], qz = rot_xyzw[2], qw = rot_xyzw[3];
    float qxx(qx * qx), qyy(qy * qy), qzz(qz * qz), qxz(qx * qz), qxy(qx * qy), qyz(qy * qz), qwx(qw * qx), qwy(qw * qy), qwz(qw * qz);
    m4 R = identity();
    el(R, 0, 0) = 1.0f - 2.0f * (qyy + qzz); el(R, 0, 1) = 2.0f * (qxy + qwz); el(R, 0, 2) = 2.0f * (qxz - qwy);
    el(R, 1, 0) = 2.0f * (qxy - qwz); el(R, 1, 1) = 1.0f - 2.0f * (qxx + qzz); el(R, 1, 2) = 2.0f * (qyz + qwx);
    el(R, 2, 0) = 2.0f * (qxz + qwy); el(R, 2, 1) = 2.0f * (qyz - qwx); el(R, 2, 2) = 1.0f - 2.0f * (qxx + qyy);
    // glm/ext/matrix_transform.inl:78-86
    m4 S;
    S.c[0] = I.c[0] * scale[0]; S.c[1] = I.c[1] * scale[1]; S.c[2] = I.c[2] * scale[2]; S.c[3] = I.c[3];
    m4 local = mul(mul(T, R), S);
    m4 g = mul(par, local);
    std::memcpy(global, &g, 64);
}

// Renderer.cu:77-88
void orc_transform_sphere(const float global[16], const float in[5], float out[5])
{
    m4 g;
    std::memcpy(&g, global, 64);
    v4 c = mul(g, v4{ in[0], in[1], in[2], 1.0f });
    v3 c3 = v3{ c.x, c.y, c.z } / c.w;
    float sx = length(v3{ g.c[0].x, g.c[0].y, g.c[0].z });
    float sy = length(v3{ g.c[1].x, g.c[1].y, g.c[1].z });
    float sz = length(v3{ g.c[2].x, g.c[2].y, g.c[2].z });
    float uniformScale = (sx + sy + sz) / 3.0f;
    out[0] = c3.x; out[1] = c3.y; out[2] = c3.z;
    out[3] = in[3] * uniformScale;
    out[4] = in[4];
}

// primary visibility through trace_ray
void orc_primary_hits(const void* spheres, uint32_t nS, const float origin[3], const float* dirs, uint32_t width,
                      uint32_t height, int32_t* out, int threads)
{
    const Sphere* s = static_cast<const Sphere*>(spheres);
    v3 o{ origin[0], origin[1], origin[2] };
    parallel_rows(height, threads, [&](uint32_t y0, uint32_t y1, int) {
        for (uint32_t y = y0; y < y1; y++)
            for (uint32_t x = 0; x < width; x++)
            {
                const float* d = dirs + 3ull * (x + static_cast<size_t>(y) * width);
                out[x + static_cast<size_t>(y) * width] = trace_ray(o, v3{ d[0], d[1], d[2] }, s, nS).id;
            }
    });
}

// n_frames x kernelRender's accumulate (Renderer.cu:162-165): acc[p] += perPixel(...) for frame indices
// first, first+stride, ... Rows [y_begin, y_end) only (bounded samples for the CPU baseline). Returns rays traced.
uint64_t orc_render(const void* spheres, uint32_t nS, const void* materials, uint32_t nM, const void* lights,
                    uint32_t nL, const float origin[3], const float* dirs, uint32_t width, uint32_t height,
                    uint32_t y_begin, uint32_t y_end, uint32_t first_frame, uint32_t n_frames, uint32_t frame_stride,
                    int max_bounces, int sky_light, float* accum, int threads)
{
    const Sphere* s = static_cast<const Sphere*>(spheres);
    const Material* m = static_cast<const Material*>(materials);
    const Light* l = static_cast<const Light*>(lights);
    v3 o{ origin[0], origin[1], origin[2] };
    if (y_end > height) y_end = height;
    if (y_begin >= y_end) return 0;
    std::vector<uint64_t> rayCounts(1024, 0);
    const uint32_t rowsTotal = y_end - y_begin;
    parallel_rows(rowsTotal, threads, [&](uint32_t r0, uint32_t r1, int t) {
        uint64_t rays = 0;
        for (uint32_t y = y_begin + r0; y < y_begin + r1; y++)
            for (uint32_t x = 0; x < width; x++)
            {
                const uint32_t p = x + y * width;
                const float* d = dirs + 3ull * p;
                float* a = accum + 4ull * p;
                for (uint32_t j = 0; j < n_frames; j++)
                {
                    v4 c = per_pixel(p, o, v3{ d[0], d[1], d[2] }, s, nS, m, nM, first_frame + j * frame_stride, l, nL,
                                     max_bounces, sky_light != 0, &rays);
                    a[0] += c.x; a[1] += c.y; a[2] += c.z; a[3] += c.w;
                }
            }
        rayCounts[t & 1023] += rays;
    });
    uint64_t total = 0;
    for (uint64_t r : rayCounts) total += r;
    return total;
}

// Renderer.cu:166-168 + Renderer.h:70-78
void orc_pack_rgba8(const float* accum, uint32_t n, float divisor, uint32_t* out)
{
    for (uint32_t i = 0; i < n; i++)
    {
        uint32_t u[4];
        for (int k = 0; k < 4; k++)
        {
            float v = accum[4ull * i + k] / divisor;
            v = gclamp(v, 0.0f, 1.0f);
            u[k] = static_cast<uint8_t>(v * 255.0f);
        }
        out[i] = (u[3] << 24) | (u[2] << 16) | (u[1] << 8) | u[0];
    }
}

int orc_hardware_threads(void) { return static_cast<int>(std::thread::hardware_concurrency()); }

} // extern "C"
