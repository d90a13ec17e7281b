// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// C interface over the UNMODIFIED reference sources compiled for the HOST: with
// shim/host_device_shim.h force-included, every __device__ function of
// Renderer.cu / BRDF.cu / Random.h is also emitted as host code, so
// Renderer::perPixel, Renderer::traceRay, BRDF::* and colorUtils::vec4ToRGBA run on
// the CPU exactly as written ("per-pixel shading compiled host-side", north_star).
// Camera, SceneNode, Utils are the reference's own host classes.
//
// Built by oracle/ref/Makefile into oracle/_ref/libref_cpu.so. Used (a) to pin the
// CPU restatement oracle/oracle.cpp bit for bit, (b) to generate tests/golden/*,
// (c) as the "reference" CPU baseline timed by bench.py.
#include <thread>
#include <fstream>
#include <nlohmann/json.hpp>
#include "Random.h"
#define private public
#include "Renderer.h"
#undef private
#include "Utils.h"

#define REF_API extern "C" __attribute__((visibility("default")))

namespace
{
template <typename F>
void parallel_rows(uint32_t rows, int threads, F&& fn)
{
    int n = threads > 0 ? threads : static_cast<int>(std::thread::hardware_concurrency());
    if (n < 1) n = 1;
    if (static_cast<uint32_t>(n) > rows) n = static_cast<int>(rows);
    if (n <= 1) { fn(0u, rows); return; }
    std::vector<std::thread> pool;
    const uint32_t per = rows / n; // contiguous row bands, as Camera.cpp:166-194
    for (int t = 0; t < n; t++)
    {
        const uint32_t r0 = t * per, r1 = (t == n - 1) ? rows : r0 + per;
        pool.emplace_back([=, &fn]() { fn(r0, r1); });
    }
    for (auto& th : pool) th.join();
}
}

REF_API uint32_t refcpu_pcg_hash(uint32_t seed) { return Random::Random::PcgHash(seed); }
REF_API float refcpu_pcg_float(uint32_t* seed) { return Random::Random::PcgFloat(*seed); }

REF_API void* refcpu_scene_load(const char* path)
{
    Scene* s = new Scene(Utils::importScene(path));
    return s;
}
REF_API void refcpu_scene_free(void* scene) { delete static_cast<Scene*>(scene); }

REF_API void refcpu_scene_export(void* scene, const char* path)
{
    Utils::exportScene(*static_cast<Scene*>(scene), path);
}

REF_API void refcpu_scene_info(void* scene, uint32_t* nSpheres, uint32_t* nMaterials, uint32_t* nLights, float camPos[3],
                               float camDir[3], float* fov, int* maxBounces, int* skyLight, int* accumulation)
{
    Scene& s = *static_cast<Scene*>(scene);
    std::vector<Sphere> flat;
    Renderer::traverseSceneGraph(s.rootNode, glm::mat4(1.0f), flat);
    *nSpheres = static_cast<uint32_t>(flat.size());
    *nMaterials = static_cast<uint32_t>(s.materials.size());
    *nLights = static_cast<uint32_t>(s.lights.size());
    for (int i = 0; i < 3; i++) { camPos[i] = s.camera.getPosition()[i]; camDir[i] = s.camera.getDirection()[i]; }
    *fov = s.camera.getFov();
    *maxBounces = s.settings.maxBounces;
    *skyLight = s.settings.skyLight;
    *accumulation = s.settings.accumulation;
}

// flattened world-space spheres with the material clamp of Renderer.cu:30-37, + materials + lights (reference PODs)
REF_API void refcpu_scene_arrays(void* scene, void* spheres20, void* materials52, void* lights28)
{
    Scene& s = *static_cast<Scene*>(scene);
    std::vector<Sphere> flat;
    Renderer::traverseSceneGraph(s.rootNode, glm::mat4(1.0f), flat);
    for (auto& sp : flat)
        if (static_cast<uint32_t>(sp.id) >= s.materials.size())
            sp.id = 0;
    if (spheres20 && !flat.empty()) std::memcpy(spheres20, flat.data(), flat.size() * sizeof(Sphere));
    if (materials52 && !s.materials.empty()) std::memcpy(materials52, s.materials.data(), s.materials.size() * sizeof(Material));
    if (lights28 && !s.lights.empty()) std::memcpy(lights28, s.lights.data(), s.lights.size() * sizeof(Light));
}

// canonical camera protocol (SURVEY.md Q-cam): Camera(fov, near, far, pos, dir) then Resize(W, H)
REF_API int refcpu_camera(const float pos[3], const float dir[3], float fov, float nearClip, float farClip, uint32_t W,
                          uint32_t H, float* rays, float invProj[16], float invView[16])
{
    Camera cam(fov, nearClip, farClip, glm::vec3(pos[0], pos[1], pos[2]), glm::vec3(dir[0], dir[1], dir[2]));
    cam.Resize(W, H);
    const auto& table = cam.getRayDirection();
    if (table.size() != static_cast<size_t>(W) * H)
        return -1; // (1600, 900) early-return quirk: no table was built
    if (rays) std::memcpy(rays, table.data(), table.size() * sizeof(glm::vec3));
    if (invProj) std::memcpy(invProj, &cam.getInverseProjectionMatrix(), 64);
    if (invView) std::memcpy(invView, &cam.getInverseViewMatrix(), 64);
    return 0;
}

// Camera::onUpdate (Camera.cpp:30-108) driven by a scripted input record (input_stub.cpp): one step per
// entry of `steps` = (dt, mouseX, mouseY, keyBits, rightButton) with keyBits W=1 S=2 A=4 D=8 Q=16 E=32.
// Outputs per step: position, direction, inverse view matrix, and what onUpdate returned.
extern "C" void refinput_set(const uint16_t* keyCodes, int nKeys, int rightButton, float mouseX, float mouseY);
struct RefCameraStep { float dt, mouseX, mouseY; uint32_t keys, right; };
REF_API int refcpu_camera_walk(const float pos[3], const float dir[3], float fov, float nearClip, float farClip, uint32_t W,
                               uint32_t H, const RefCameraStep* steps, int nSteps, float* outPos, float* outDir,
                               float* outInvView, int* outMoved, float* finalRays)
{
    Camera cam(fov, nearClip, farClip, glm::vec3(pos[0], pos[1], pos[2]), glm::vec3(dir[0], dir[1], dir[2]));
    cam.Resize(W, H);
    static const uint16_t codes[6] = { 87, 83, 65, 68, 81, 69 }; // W S A D Q E (KeyCodes.h)
    for (int i = 0; i < nSteps; i++)
    {
        uint16_t held[6];
        int n = 0;
        for (int k = 0; k < 6; k++)
            if (steps[i].keys & (1u << k)) held[n++] = codes[k];
        refinput_set(held, n, static_cast<int>(steps[i].right), steps[i].mouseX, steps[i].mouseY);
        outMoved[i] = cam.onUpdate(steps[i].dt) ? 1 : 0;
        std::memcpy(outPos + 3 * i, &cam.getPosition(), 12);
        std::memcpy(outDir + 3 * i, &cam.getDirection(), 12);
        std::memcpy(outInvView + 16 * i, &cam.getInverseViewMatrix(), 64);
    }
    refinput_set(nullptr, 0, 0, 0.0f, 0.0f);
    const auto& table = cam.getRayDirection();
    if (finalRays && table.size() == static_cast<size_t>(W) * H)
        std::memcpy(finalRays, table.data(), table.size() * sizeof(glm::vec3));
    return 0;
}

REF_API void refcpu_primary_hits(const void* spheres, uint32_t nS, const float origin[3], const float* dirs, uint32_t W,
                                 uint32_t H, int32_t* out, int threads)
{
    const Sphere* s = static_cast<const Sphere*>(spheres);
    parallel_rows(H, threads, [&](uint32_t y0, uint32_t y1) {
        for (uint32_t y = y0; y < y1; y++)
            for (uint32_t x = 0; x < W; x++)
            {
                Ray ray;
                ray.origin = glm::vec3(origin[0], origin[1], origin[2]);
                const float* d = dirs + 3ull * (x + static_cast<size_t>(y) * W);
                ray.direction = glm::vec3(d[0], d[1], d[2]);
                auto ht = Renderer::traceRay(ray, s, nS);
                out[x + static_cast<size_t>(y) * W] = ht.t < 0.0f ? -1 : static_cast<int>(ht.id);
            }
    });
}

// n_frames x { accumulation[p] += Renderer::perPixel(...) } (Renderer.cu:162-165) on rows [y_begin, y_end)
REF_API void refcpu_render(const void* spheres, uint32_t nS, const void* materials, uint32_t nM, const void* lights,
                           uint32_t nL, const float origin[3], const float* dirs, uint32_t W, uint32_t H,
                           uint32_t y_begin, uint32_t y_end, uint32_t first_frame, uint32_t n_frames,
                           uint32_t frame_stride, int max_bounces, int sky_light, float* accum, int threads)
{
    DeviceCamera cam;
    cam.position = glm::vec3(origin[0], origin[1], origin[2]);
    cam.direction = glm::vec3(0.0f);
    cam.width = W;
    cam.height = H;
    cam.rayDirection = reinterpret_cast<glm::vec3*>(const_cast<float*>(dirs));
    Settings st;
    st.accumulation = true;
    st.skyLight = sky_light != 0;
    st.maxBounces = max_bounces;
    if (y_end > H) y_end = H;
    if (y_begin >= y_end) return;
    glm::vec4* acc = reinterpret_cast<glm::vec4*>(accum);
    parallel_rows(y_end - y_begin, threads, [&](uint32_t r0, uint32_t r1) {
        for (uint32_t y = y_begin + r0; y < y_begin + r1; y++)
            for (uint32_t x = 0; x < W; x++)
                for (uint32_t j = 0; j < n_frames; j++)
                {
                    const glm::vec4 c = Renderer::perPixel(x, y, W, static_cast<const Sphere*>(spheres), nS, cam,
                        static_cast<const Material*>(materials), nM, first_frame + j * frame_stride,
                        static_cast<const Light*>(lights), nL, st);
                    acc[x + y * W] += c;
                }
    });
}

// finalColor = clamp(acc / frameIndex, 0, 1); vec4ToRGBA   (Renderer.cu:166-168)
REF_API void refcpu_pack_rgba8(const float* accum, uint32_t n, float divisor, uint32_t* out)
{
    const glm::vec4* acc = reinterpret_cast<const glm::vec4*>(accum);
    for (uint32_t i = 0; i < n; i++)
    {
        glm::vec4 finalColor = acc[i] / divisor;
        finalColor = glm::clamp(finalColor, 0.0f, 1.0f);
        out[i] = colorUtils::vec4ToRGBA(finalColor);
    }
}

REF_API int refcpu_hardware_threads(void) { return static_cast<int>(std::thread::hardware_concurrency()); }
