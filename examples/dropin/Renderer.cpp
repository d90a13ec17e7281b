// examples/dropin/Renderer.cpp — the drop-in: this ONE translation unit replaces Engine/src/Renderer.cu and
// Engine/src/BRDF.cu of 1neskk/Ataraxia. It is compiled against the reference's UNMODIFIED headers
// (Engine/include/Renderer.h:19-61 and everything it includes) by a plain C++ compiler and linked with
// -lataraxia_b200; the application's three calls per frame (Engine/src/main.cpp:211-220: onResize,
// Camera::Resize, Render) then run on the B200-native kernels through the C-ABI (include/ataraxia_b200.h).
//
// Renderer.h is not edited: the class still declares its CudaBuffer members (they stay empty) and the backend
// handle lives in a side table keyed by the object's address. oracle/ref/Makefile builds this file together with
// the reference's own Camera.cpp / SceneNode.cpp / Utils.cpp into oracle/_ref/ref_headless_shim, and
// tests/test_dropin.py compares what that binary renders with the unmodified reference (ref_headless), bit for bit.
#include "Renderer.h"

#include <ataraxia_b200.h>

#include <iostream>
#include <mutex>
#include <unordered_map>

namespace
{
std::mutex g_tableLock;
std::unordered_map<const Renderer*, atx_handle> g_backends;

atx_handle backend(const Renderer* r)
{
    std::lock_guard<std::mutex> guard(g_tableLock);
    auto it = g_backends.find(r);
    return it == g_backends.end() ? nullptr : it->second;
}

bool ok(atx_status s, const char* what)
{
    if (s == ATX_OK)
        return true;
    // the reference prints and drops the frame on kernel errors (Renderer.cu:226-238); it exits on allocation
    // errors (DeviceMemory.h:7-16) - this layer never exits
    std::cerr << "ataraxia_b200: " << what << ": " << atx_last_error() << "\n";
    return false;
}
}

// headless addition for tools and tests: the C-ABI handle behind a Renderer (accumulation read-back, counters, ...)
atx_handle ataraxia_b200_backend(const Renderer* r) { return backend(r); }

Renderer::Renderer() : h_imageData_(nullptr), m_frameIndex(1)
{
    atx_handle h = nullptr;
    if (ok(atx_create(0, &h), "atx_create")) // the reference implies device 0 (Renderer.cu:190)
    {
        std::lock_guard<std::mutex> guard(g_tableLock);
        g_backends[this] = h;
    }
}

Renderer::~Renderer()
{
    atx_handle h = nullptr;
    {
        std::lock_guard<std::mutex> guard(g_tableLock);
        auto it = g_backends.find(this);
        if (it != g_backends.end())
        {
            h = it->second;
            g_backends.erase(it);
        }
    }
    atx_destroy(h);
    atx_host_free(h_imageData_);
}

// Renderer.cu:98-146
void Renderer::onResize(uint32_t width, uint32_t height)
{
    if (m_image && m_image->getWidth() == width && m_image->getHeight() == height)
        return;
    if (!ok(atx_resize(backend(this), width, height), "onResize"))
        return;
    m_image = std::make_shared<Image>(width, height, ImageType::RGBA);
    atx_host_free(h_imageData_);
    h_imageData_ = nullptr;
    void* pixels = nullptr; // page-locked: the per-frame read-back is one DMA
    if (!ok(atx_host_alloc(static_cast<size_t>(width) * height * sizeof(uint32_t), &pixels), "onResize (host image)"))
        return;
    h_imageData_ = static_cast<uint32_t*>(pixels);
    m_width = width;
    m_height = height;
    m_frameIndex = 1;
}

// Renderer.cu:67-96: world-space spheres in pre-order (a node's own spheres, then its children in order). The node
// transform is the reference's own SceneNode::updateGlobalTransform; the per-sphere arithmetic is the library's
// (atx_host_transform_sphere), so that this file's compiler flags cannot change a bit of it.
void Renderer::traverseSceneGraph(const std::shared_ptr<SceneNode>& node, const glm::mat4& parentTransform, std::vector<Sphere>& spheres)
{
    if (!node)
        return;
    node->updateGlobalTransform(parentTransform);
    const glm::mat4 global = node->getGlobalTransform();
    static_assert(sizeof(Sphere) == sizeof(atx_sphere), "Sphere is the C-ABI's sphere record");
    for (const Sphere& local : node->getSpheres())
    {
        Sphere world = local;
        atx_host_transform_sphere(&global[0].x, reinterpret_cast<const atx_sphere*>(&local), reinterpret_cast<atx_sphere*>(&world));
        spheres.push_back(world);
    }
    for (const std::shared_ptr<SceneNode>& child : node->getChildren())
        traverseSceneGraph(child, global, spheres);
}

// Renderer.cu:25-57; the id clamp of :30-37 happens inside atx_upload_scene, the message is kept here
void Renderer::allocateDeviceMemory(const Scene& scene)
{
    std::vector<Sphere> world;
    traverseSceneGraph(scene.rootNode, glm::mat4(1.0f), world);
    for (const Sphere& s : world)
        if (static_cast<uint32_t>(s.id) >= scene.materials.size())
            std::cerr << "Sphere ID out of bounds: " << s.id << "\n";
    static_assert(sizeof(Material) == sizeof(atx_material) && sizeof(Light) == sizeof(atx_light), "scene records are the C-ABI's");
    m_numSpheres = world.size();
    m_numMaterials = scene.materials.size();
    m_numLights = scene.lights.size();
    ok(atx_upload_scene(backend(this), reinterpret_cast<const atx_sphere*>(world.data()), world.size(),
                        reinterpret_cast<const atx_material*>(scene.materials.data()), scene.materials.size(),
                        reinterpret_cast<const atx_light*>(scene.lights.data()), scene.lights.size()),
       "scene upload");
}

void Renderer::freeDeviceMemory() {} // the handle owns every device buffer

// Renderer.cu:173-249
void Renderer::Render(Camera& camera, const Scene& scene)
{
    atx_handle h = backend(this);
    if (m_scene != &scene || m_frameIndex == 1) // :175-179
    {
        m_scene = &scene;
        allocateDeviceMemory(scene);
    }
    if (m_frameIndex == 1) // :181-182: the next frame starts from a cleared buffer
        atx_reset(h);
    if (!m_image || !h_imageData_)
        return;
    // Camera::allocateDevice's W*H*12 B table (Camera.cpp:197-210) is never uploaded: rays are generated in the
    // kernel from the two inverse matrices, bit-identical to Camera::UpdateRayDirection
    if (!ok(atx_set_settings(h, m_settings.accumulation, m_settings.skyLight, m_settings.maxBounces), "settings") ||
        !ok(atx_set_camera_matrices(h, &camera.getPosition().x, &camera.getInverseProjectionMatrix()[0].x,
                                    &camera.getInverseViewMatrix()[0].x), "camera") ||
        !ok(atx_render(h, 1, ATX_VARIANT_AUTO), "Render") ||     // kernelRender<<<>>> + sync, :223-238
        !ok(atx_read_rgba8(h, h_imageData_, 0), "read-back"))    // :240
        return;                                                   // the frame is dropped, frameIndex stays
    m_image->setData(h_imageData_);                               // :242
    m_frameIndex = m_settings.accumulation ? m_frameIndex + 1 : 1; // :245-248
}
