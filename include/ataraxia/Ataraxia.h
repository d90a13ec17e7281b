// Ataraxia.h — header-only C++ mirror of the reference's host API for the path-tracing path
// (Engine/include/{Scene,SceneNode,Camera,Renderer,Utils}.h of 1neskk/Ataraxia), over the C-ABI of
// libataraxia_b200.so (include/ataraxia_b200.h). Same class names, method names, argument meaning
// and error behaviour; no Vulkan, GLFW, ImGui, glm or CUDA headers needed by the caller.
//
//   reference                                        here
//   glm::vec3 / quat / mat4                          atx::vec3 / quat / mat4 (layout-compatible, Math.h)
//   Sphere, Material, Light, Settings, Ray           same PODs, same sizes (20 / 52 / 28 / 8 / 24 B)
//   SceneNode (SceneNode.h:23-66, SceneNode.cpp)     same; transforms evaluated inside the library in glm's order
//   Camera (Camera.h:14-88, Camera.cpp)              same minus onUpdate (GLFW input) and allocateDevice/freeDevice
//                                                    (the ray table is never uploaded: rays are generated in-kernel)
//   Image (Core/include/Image.h)                     headless: host RGBA8 pixels, getWidth/getHeight/setData
//   Renderer (Renderer.h:19-31, Renderer.cu)         same + headless additions (frames per launch, accumulation
//                                                    read-back, counters)
//   Utils::importScene/exportScene/... (Utils.h)     same schema and file format (Json.h instead of nlohmann)
//
// All numerics that parity depends on (camera matrices, node transforms, sphere flattening) run inside
// the library (atx_host_*), so the caller's compiler flags cannot perturb them.
#pragma once

#include "../ataraxia_b200.h"
#include "Json.h"
#include "Math.h"

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace ataraxia
{
using atx::mat4;
using atx::quat;
using atx::vec2;
using atx::vec3;
using atx::vec4;

// ---- PODs: Scene.h:11-56, SceneNode.h:11-21 -----------------------------------------------------
struct Ray
{
    vec3 origin;
    vec3 direction;
};

struct Light
{
    vec3 position;
    vec3 color;
    float intensity = 0.0f;
    Light() = default;
    Light(const vec3& pos, const vec3& col, float i) : position(pos), color(col), intensity(i) {}
};

struct Material
{
    vec3 albedo{ 1.0f };
    float roughness = 0.0f;
    float metallic = 0.0f;
    vec3 F0{ 0.04f };
    vec3 emissionColor{ 0.0f };
    float emissionIntensity = 0.0f;
    int id = 0;
    vec3 getEmission() const { return emissionColor * emissionIntensity; }
    Material() = default;
    Material(const vec3& albedo_, float roughness_, float metallic_, const vec3& emissionColor_, float emissionIntensity_, int id_)
        : albedo(albedo_), roughness(roughness_), metallic(metallic_), emissionColor(emissionColor_),
          emissionIntensity(emissionIntensity_), id(id_) {}
};

struct Settings
{
    bool accumulation = true;
    bool skyLight = false;
    int maxBounces = 15;
};

struct Sphere
{
    vec3 center;
    float radius = 0.0f;
    int id = 0;
    Sphere() = default;
    Sphere(const vec3& c, float r, int materialId) : center(c), radius(r), id(materialId) {}
};

static_assert(sizeof(Sphere) == 20 && sizeof(Material) == 52 && sizeof(Light) == 28 && sizeof(Settings) == 8 && sizeof(Ray) == 24,
              "POD layouts must match the reference (SURVEY.md §8c)");
static_assert(sizeof(Sphere) == sizeof(atx_sphere) && sizeof(Material) == sizeof(atx_material) && sizeof(Light) == sizeof(atx_light),
              "PODs are passed to the C-ABI as they are");

// ---- SceneNode: SceneNode.h:23-66, SceneNode.cpp -------------------------------------------------
class SceneNode
{
public:
    SceneNode() : SceneNode("Untitled") {}                       // SceneNode.cpp:3-7
    explicit SceneNode(const std::string& name)                  // SceneNode.cpp:9-13
        : m_name(name), m_position(0.0f), m_rotation(), m_scale(1.0f), m_localTransform(1.0f), m_globalTransform(1.0f),
          m_transformDirty(true) {}

    void setPosition(const vec3& position) { m_position = position; m_transformDirty = true; }
    void setRotation(const quat& rotation) { m_rotation = rotation; m_transformDirty = true; }
    void setScale(const vec3& scale) { m_scale = scale; m_transformDirty = true; }
    const vec3& getPosition() const { return m_position; }
    const quat& getRotation() const { return m_rotation; }
    const vec3& getScale() const { return m_scale; }

    void addChild(std::shared_ptr<SceneNode> child) { m_children.push_back(std::move(child)); }            // :15-18
    void removeChild(std::shared_ptr<SceneNode> child)                                                      // :20-28
    {
        m_children.erase(std::remove(m_children.begin(), m_children.end(), child), m_children.end());
    }
    const std::vector<std::shared_ptr<SceneNode>>& getChildren() const { return m_children; }

    void addSphere(const Sphere& sphere) { m_spheres.push_back(sphere); }                                   // :30-33
    void removeSphere(int sphereIndex)                                                                      // :35-40
    {
        if (sphereIndex >= 0 && sphereIndex < static_cast<int>(m_spheres.size()))
            m_spheres.erase(m_spheres.begin() + sphereIndex);
    }
    const std::vector<Sphere>& getSpheres() const { return m_spheres; }

    // SceneNode.cpp:42-59: local = T * R * S when dirty; global = parent * local; recurse
    void updateGlobalTransform(const mat4& parentTransform = mat4(1.0f))
    {
        if (m_transformDirty)
        {
            const float rot[4] = { m_rotation.x, m_rotation.y, m_rotation.z, m_rotation.w };
            atx_host_node_transform(&parentTransform[0].x, &m_position.x, rot, &m_scale.x, &m_localTransform[0].x,
                                    &m_globalTransform[0].x);
            m_transformDirty = false;
        }
        else
            atx_host_mat4_mul(&parentTransform[0].x, &m_localTransform[0].x, &m_globalTransform[0].x);
        for (auto& child : m_children)
            child->updateGlobalTransform(m_globalTransform);
    }
    const mat4& getGlobalTransform() const { return m_globalTransform; }

    const std::string& getName() const { return m_name; }
    void setName(const std::string& name) { m_name = name; }

private:
    std::string m_name;
    vec3 m_position;
    quat m_rotation;
    vec3 m_scale;
    mat4 m_localTransform;
    mat4 m_globalTransform;
    std::vector<std::shared_ptr<SceneNode>> m_children;
    std::vector<Sphere> m_spheres;
    bool m_transformDirty;
};

// ---- Camera: Camera.h:14-88, Camera.cpp ----------------------------------------------------------
// What Camera::onUpdate asks the window for (Core/include/input/Input.h:11-16) as plain data, so camera motion
// can be scripted without GLFW: the six keys the camera reads, the cursor position, the right mouse button.
struct InputState
{
    bool W = false, A = false, S = false, D = false, Q = false, E = false;
    bool rightButton = false;
    vec2 mouse{ 0.0f, 0.0f };
};

class Camera
{
public:
    Camera() = default;
    // Camera.cpp:14-29: both value constructors preset 1600x900 and build view and projection
    Camera(float fov, float nearClip, float farClip) : Camera(fov, nearClip, farClip, vec3(0.0f, 0.0f, 3.0f), vec3(0.0f, 0.0f, -1.0f)) {}
    Camera(float fov, float nearClip, float farClip, vec3 position, vec3 direction)
        : m_position(position), m_direction(direction), m_fov(fov), m_nearClip(nearClip), m_farClip(farClip), m_width(1600), m_height(900)
    {
        UpdateViewMatrix();
        UpdateProjectionMatrix();
    }

    // Camera::onUpdate (Camera.cpp:30-108): W/S, A/D, Q/E at speed 5 and mouse look at rotation speed 0.3 while the
    // right button is held; returns whether the camera moved. Arithmetic in the library, in glm's order.
    bool onUpdate(float dt, const InputState& input = InputState())
    {
        atx_camera_input in{};
        const auto bit = [](bool held, uint32_t key) { return held ? key : 0u; };
        in.keys = bit(input.W, ATX_KEY_W) | bit(input.S, ATX_KEY_S) | bit(input.A, ATX_KEY_A) | bit(input.D, ATX_KEY_D) |
                  bit(input.Q, ATX_KEY_Q) | bit(input.E, ATX_KEY_E);
        in.right_button = input.rightButton ? 1u : 0u;
        in.mouse_x = input.mouse.x;
        in.mouse_y = input.mouse.y;
        int moved = 0;
        atx_host_camera_update(&m_position.x, &m_direction.x, &m_lastMousePos.x, &in, dt, &moved);
        if (moved)
            m_viewDirty = true;
        if (m_viewDirty && input.rightButton) // :101-105 (not reached when the button is up, :36-40)
            UpdateViewMatrix();
        return moved != 0;
    }

    // Camera.cpp:110-127, including the early return that leaves a 1600x900 camera without a ray table (quirk Q-cam)
    void Resize(uint32_t width, uint32_t height)
    {
        if (width == 0 || height == 0)
        {
            std::cerr << "Error: Width or height cannot be zero." << std::endl;
            return;
        }
        if (width == m_width && height == m_height)
            return;
        m_width = width;
        m_height = height;
        m_projectionDirty = true;
        UpdateProjectionMatrix();
        m_rayDirection.clear(); // rebuilt on demand: the renderer generates rays in-kernel
        m_raysValid = false;
    }

    const mat4& getViewMatrix() const { return m_viewMatrix; }
    const mat4& getProjectionMatrix() const { return m_projectionMatrix; }
    const mat4& getInverseViewMatrix() const { return m_inverseViewMatrix; }
    const mat4& getInverseProjectionMatrix() const { return m_inverseProjectionMatrix; }
    const vec3& getPosition() const { return m_position; }
    const vec3& getDirection() const { return m_direction; }
    const float& getFov() const { return m_fov; }
    uint32_t getWidth() const { return m_width; }
    uint32_t getHeight() const { return m_height; }

    // the setters only mark dirty (Camera.h:56-58): matrices change on the next Update* call
    void setPosition(const vec3& position) { m_position = position; m_viewDirty = true; }
    void setDirection(const vec3& direction) { m_direction = direction; m_viewDirty = true; }
    void setFov(float fov) { m_fov = fov; m_projectionDirty = true; }

    // Camera::UpdateRayDirection (Camera.cpp:161-195): the host table, multithreaded like the reference
    const std::vector<vec3>& getRayDirection() const
    {
        if (!m_raysValid && m_width && m_height)
        {
            m_rayDirection.resize(static_cast<size_t>(m_width) * m_height);
            atx_host_ray_directions(&m_inverseProjectionMatrix[0].x, &m_inverseViewMatrix[0].x, m_width, m_height,
                                    &m_rayDirection[0].x);
            m_raysValid = true;
        }
        return m_rayDirection;
    }
    static float getRotationSpeed() { return 0.3f; }

private:
    void UpdateProjectionMatrix() // Camera.cpp:134-149
    {
        if (!m_projectionDirty)
            return;
        atx_host_camera_matrices(&m_position.x, &m_direction.x, m_fov, m_nearClip, m_farClip, std::max(m_width, 1u), std::max(m_height, 1u),
                                 &m_projectionMatrix[0].x, nullptr, &m_inverseProjectionMatrix[0].x, nullptr);
        m_projectionDirty = false;
        m_raysValid = false;
    }
    void UpdateViewMatrix() // Camera.cpp:151-159
    {
        if (!m_viewDirty)
            return;
        atx_host_camera_matrices(&m_position.x, &m_direction.x, m_fov, m_nearClip, m_farClip, std::max(m_width, 1u), std::max(m_height, 1u),
                                 nullptr, &m_viewMatrix[0].x, nullptr, &m_inverseViewMatrix[0].x);
        m_viewDirty = false;
        m_raysValid = false;
    }

    mat4 m_projectionMatrix{ 1.0f };
    mat4 m_viewMatrix{ 1.0f };
    mat4 m_inverseProjectionMatrix{ 1.0f };
    mat4 m_inverseViewMatrix{ 1.0f };
    vec3 m_position{ 0.0f };
    vec3 m_direction{ 0.0f };
    vec2 m_lastMousePos{ 0.0f }; // Camera.h:83
    mutable std::vector<vec3> m_rayDirection;
    mutable bool m_raysValid = false;
    float m_fov = 45.0f;
    float m_nearClip = 0.1f;
    float m_farClip = 100.0f;
    uint32_t m_width = 0, m_height = 0;
    bool m_viewDirty = true;
    bool m_projectionDirty = true;
};

// ---- Scene: Scene.h:58-80 ------------------------------------------------------------------------
struct Scene
{
    std::shared_ptr<SceneNode> rootNode;
    std::vector<Material> materials;
    std::vector<Light> lights;
    Settings settings;
    Camera camera;
    Scene() : rootNode(std::make_shared<SceneNode>("Scene")) {}
};

// ---- Image: the members Renderer touches (Core/include/Image.h; Renderer.cu:106, :121, :100, :242) ----
enum class ImageType { None = 0, RGBA, RGBA32F };
class Image
{
public:
    Image(uint32_t width, uint32_t height, ImageType type = ImageType::RGBA, const void* data = nullptr)
        : m_width(width), m_height(height), m_type(type), m_pixels(static_cast<size_t>(width) * height, 0u)
    {
        if (data)
            setData(data);
    }
    void setData(const void* data) { std::memcpy(m_pixels.data(), data, m_pixels.size() * sizeof(uint32_t)); }
    uint32_t getWidth() const { return m_width; }
    uint32_t getHeight() const { return m_height; }
    const uint32_t* getPixels() const { return m_pixels.data(); } // RGBA8, row 0 = bottom of the image (main.cpp:186-187 flips V)

    // Output sink that replaces the Vulkan texture upload (Core/src/Image.cpp:183-271): binary PPM (P6), rows
    // written top to bottom, i.e. with the V flip the UI applies when it shows the texture (main.cpp:185-187).
    bool savePPM(const std::string& path) const
    {
        std::ofstream f(path, std::ios::binary);
        if (!f.is_open())
            return false;
        f << "P6\n" << m_width << " " << m_height << "\n255\n";
        std::vector<unsigned char> row(static_cast<size_t>(m_width) * 3);
        for (uint32_t y = 0; y < m_height; y++)
        {
            const uint32_t* src = m_pixels.data() + static_cast<size_t>(m_height - 1 - y) * m_width;
            for (uint32_t x = 0; x < m_width; x++)
            {
                row[3 * x + 0] = static_cast<unsigned char>(src[x] & 0xFFu);         // vec4ToRGBA: r in the low byte (Renderer.h:70-78)
                row[3 * x + 1] = static_cast<unsigned char>((src[x] >> 8) & 0xFFu);
                row[3 * x + 2] = static_cast<unsigned char>((src[x] >> 16) & 0xFFu);
            }
            f.write(reinterpret_cast<const char*>(row.data()), static_cast<std::streamsize>(row.size()));
        }
        return f.good();
    }
    // The same image as a PNG (8-bit RGBA, filter 0, zlib "stored" blocks: no compression library needed).
    bool savePNG(const std::string& path) const
    {
        std::ofstream f(path, std::ios::binary);
        if (!f.is_open())
            return false;
        const auto be32 = [](uint32_t v) { return std::string{ static_cast<char>(v >> 24), static_cast<char>(v >> 16), static_cast<char>(v >> 8), static_cast<char>(v) }; };
        const auto crc32 = [](const std::string& d) {
            uint32_t c = 0xFFFFFFFFu;
            for (unsigned char ch : d)
            {
                c ^= ch;
                for (int k = 0; k < 8; k++)
                    c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
            }
            return c ^ 0xFFFFFFFFu;
        };
        const auto chunk = [&](const char* kind, const std::string& body) {
            const std::string tagged = std::string(kind, 4) + body;
            f << be32(static_cast<uint32_t>(body.size())) << tagged << be32(crc32(tagged));
        };
        // scanlines top to bottom (V flip), each prefixed with filter type 0; bytes r, g, b, a as packed (Renderer.h:70-78)
        std::string raw;
        raw.reserve((static_cast<size_t>(m_width) * 4 + 1) * m_height);
        for (uint32_t y = 0; y < m_height; y++)
        {
            raw.push_back('\0');
            const uint32_t* src = m_pixels.data() + static_cast<size_t>(m_height - 1 - y) * m_width;
            for (uint32_t x = 0; x < m_width; x++)
                for (int k = 0; k < 4; k++)
                    raw.push_back(static_cast<char>((src[x] >> (8 * k)) & 0xFFu));
        }
        uint32_t a = 1u, b = 0u; // Adler-32 of the uncompressed stream
        for (unsigned char ch : raw)
        {
            a = (a + ch) % 65521u;
            b = (b + a) % 65521u;
        }
        std::string z = "\x78\x01";
        for (size_t off = 0; off < raw.size() || off == 0; off += 65535)
        {
            const size_t n = std::min<size_t>(65535, raw.size() - off);
            const bool last = off + n >= raw.size();
            z.push_back(last ? '\x01' : '\x00');
            z.push_back(static_cast<char>(n & 0xFF)); z.push_back(static_cast<char>(n >> 8));
            z.push_back(static_cast<char>(~n & 0xFF)); z.push_back(static_cast<char>((~n >> 8) & 0xFF));
            z.append(raw, off, n);
            if (last)
                break;
        }
        z += be32((b << 16) | a);
        f << "\x89PNG\r\n\x1a\n";
        std::string ihdr = be32(m_width) + be32(m_height);
        ihdr += std::string{ '\x08', '\x06', '\0', '\0', '\0' }; // 8 bits, RGBA, deflate, adaptive filtering, no interlace
        chunk("IHDR", ihdr);
        chunk("IDAT", z);
        chunk("IEND", std::string());
        return f.good();
    }
private:
    uint32_t m_width, m_height;
    ImageType m_type;
    std::vector<uint32_t> m_pixels;
};

// ---- Renderer: Renderer.h:19-31, Renderer.cu ------------------------------------------------------
class Renderer
{
public:
    explicit Renderer(int device = 0)
    {
        if (atx_create(device, &m_handle) != ATX_OK)
            report("atx_create");
    }
    ~Renderer()
    {
        freeImageData();
        atx_destroy(m_handle);
    }
    Renderer(const Renderer&) = delete; // the reference's shallow copy double-frees (DeviceMemory.h:33)
    Renderer& operator=(const Renderer&) = delete;

    // Renderer.cu:98-146
    void onResize(uint32_t width, uint32_t height)
    {
        if (m_image && m_image->getWidth() == width && m_image->getHeight() == height)
            return;
        if (atx_resize(m_handle, width, height) != ATX_OK)
        {
            report("onResize");
            return;
        }
        m_image = std::make_shared<Image>(width, height, ImageType::RGBA);
        freeImageData();
        // page-locked where possible (the per-frame read-back is then one DMA); the reference uses new[] (Renderer.cu:124-129)
        void* pinned = nullptr;
        if (atx_host_alloc(static_cast<size_t>(width) * height * sizeof(uint32_t), &pinned) == ATX_OK)
        {
            h_imageData_ = static_cast<uint32_t*>(pinned);
            m_imagePinned = true;
        }
        else
            h_imageData_ = new uint32_t[static_cast<size_t>(width) * height];
        m_width = width;
        m_height = height;
    }

    // Renderer.cu:173-249. `frames` > 1 renders that many frames in ONE launch (bit-identical to that many calls).
    void Render(Camera& camera, const Scene& scene, uint32_t frames = 1)
    {
        uint32_t frameIndex = 1;
        atx_frame_index(m_handle, &frameIndex);
        if (m_scene != &scene || frameIndex == 1) // :175-179
        {
            m_scene = &scene;
            std::vector<Sphere> spheres;
            traverseSceneGraph(scene.rootNode, mat4(1.0f), spheres);
            for (const Sphere& s : spheres) // the message of Renderer.cu:32-36; the clamp itself happens in the library
                if (static_cast<uint32_t>(s.id) >= scene.materials.size() && !scene.materials.empty())
                    std::cerr << "Warning: Sphere has invalid material ID (" << s.id << "). Setting to 0." << std::endl;
            if (atx_upload_scene(m_handle, reinterpret_cast<const atx_sphere*>(spheres.data()), spheres.size(),
                                 reinterpret_cast<const atx_material*>(scene.materials.data()), scene.materials.size(),
                                 reinterpret_cast<const atx_light*>(scene.lights.data()), scene.lights.size()) != ATX_OK)
                return report("Render: scene upload");
        }
        if (!m_image)
            return; // :184
        if (atx_set_settings(m_handle, m_settings.accumulation, m_settings.skyLight, m_settings.maxBounces) != ATX_OK ||
            atx_set_camera_matrices(m_handle, &camera.getPosition().x, &camera.getInverseProjectionMatrix()[0].x,
                                    &camera.getInverseViewMatrix()[0].x) != ATX_OK)
            return report("Render: state");
        // a launch or execution error drops the frame and leaves frameIndex alone (Renderer.cu:226-238)
        if (atx_render(m_handle, frames, m_variant) != ATX_OK || atx_read_rgba8(m_handle, h_imageData_, 0) != ATX_OK)
            return report("Render");
        m_image->setData(h_imageData_); // :242
    }

    std::shared_ptr<Image> getImage() const { return m_image; }
    const Settings& getSettings() const { return m_settings; }
    void setSettings(const Settings& settings) { m_settings = settings; }
    void resetFrameIndex() { atx_reset(m_handle); }

    // Renderer::traverseSceneGraph (Renderer.cu:67-96): pre-order; world-space centres, mean-scale radii
    static void traverseSceneGraph(const std::shared_ptr<SceneNode>& node, const mat4& parentTransform, std::vector<Sphere>& spheres)
    {
        if (!node)
            return;
        node->updateGlobalTransform(parentTransform);
        const mat4& g = node->getGlobalTransform();
        for (const Sphere& s : node->getSpheres())
        {
            Sphere out;
            atx_host_transform_sphere(&g[0].x, reinterpret_cast<const atx_sphere*>(&s), reinterpret_cast<atx_sphere*>(&out));
            spheres.push_back(out);
        }
        for (const auto& child : node->getChildren())
            traverseSceneGraph(child, g, spheres);
    }

    // ---- headless additions ----
    atx_handle handle() const { return m_handle; }
    void setVariant(int variant) { m_variant = variant; }
    uint32_t frameIndex() const { uint32_t f = 1; atx_frame_index(m_handle, &f); return f; }
    // float4 accumulation buffer (Renderer::d_accumulation_, Renderer.h:55): width*height*4 floats
    std::vector<float> getAccumulation() const
    {
        std::vector<float> acc(static_cast<size_t>(m_width) * m_height * 4);
        if (!acc.empty() && atx_read_accum(m_handle, acc.data()) != ATX_OK)
            report("getAccumulation");
        return acc;
    }
    std::vector<int32_t> getHitIds() const
    {
        std::vector<int32_t> hits(static_cast<size_t>(m_width) * m_height);
        if (!hits.empty() && atx_read_hit_ids(m_handle, hits.data()) != ATX_OK)
            report("getHitIds");
        return hits;
    }
    atx_counters counters() const { atx_counters c{}; atx_get_counters(m_handle, &c); return c; }
    // resumable render on disk: accumulation + next frame index, bound to scene and camera by a hash.
    // loadCheckpoint refuses (false + message) a file rendered with another size, scene, camera or settings.
    bool saveCheckpoint(const std::string& path) const
    {
        return atx_save_checkpoint(m_handle, path.c_str(), 0, 0) == ATX_OK || (report("saveCheckpoint"), false);
    }
    bool loadCheckpoint(const std::string& path)
    {
        return atx_load_checkpoint(m_handle, path.c_str(), nullptr, nullptr) == ATX_OK || (report("loadCheckpoint"), false);
    }
    // float radiance (accumulation / samples, unclamped) as a little-endian PFM, bottom row first as PFM defines it
    bool saveAccumulationPFM(const std::string& path) const
    {
        const std::vector<float> acc = getAccumulation();
        std::ofstream f(path, std::ios::binary);
        if (!f.is_open() || acc.empty())
            return false;
        f << "PF\n" << m_width << " " << m_height << "\n-1.0\n";
        std::vector<float> row(static_cast<size_t>(m_width) * 3);
        for (uint32_t y = 0; y < m_height; y++)
        {
            for (uint32_t x = 0; x < m_width; x++)
            {
                const float* a = &acc[(static_cast<size_t>(y) * m_width + x) * 4];
                const float n = a[3] > 0.0f ? a[3] : 1.0f;
                row[3 * x + 0] = a[0] / n; row[3 * x + 1] = a[1] / n; row[3 * x + 2] = a[2] / n;
            }
            f.write(reinterpret_cast<const char*>(row.data()), static_cast<std::streamsize>(row.size() * sizeof(float)));
        }
        return f.good();
    }
    float lastRenderMs() const { float ms = 0.0f; atx_last_render_ms(m_handle, &ms); return ms; }

private:
    static void report(const char* where) { std::cerr << "ataraxia_b200 (" << where << "): " << atx_last_error() << std::endl; }
    void freeImageData()
    {
        if (m_imagePinned)
            atx_host_free(h_imageData_);
        else
            delete[] h_imageData_;
        h_imageData_ = nullptr;
        m_imagePinned = false;
    }

    atx_handle m_handle = nullptr;
    std::shared_ptr<Image> m_image;
    uint32_t* h_imageData_ = nullptr;
    bool m_imagePinned = false;
    uint32_t m_width = 0, m_height = 0;
    Settings m_settings;
    const Scene* m_scene = nullptr;
    int m_variant = ATX_VARIANT_AUTO;
};

// ---- Utils: Utils.h:11-23, Utils.cpp ---------------------------------------------------------------
namespace Utils
{
using atx::Json;

inline Json vec3ToJson(const vec3& v) { Json a = Json::array(); a.push_back(v.x); a.push_back(v.y); a.push_back(v.z); return a; }
inline vec3 vec3FromJson(const Json& j) { return vec3(j[0].getFloat(), j[1].getFloat(), j[2].getFloat()); }

inline Json serializeSceneNode(const std::shared_ptr<SceneNode>& node) // Utils.cpp:64-95
{
    Json j = Json::object();
    j["name"] = node->getName();
    Json t = Json::object();
    t["position"] = vec3ToJson(node->getPosition());
    const quat& q = node->getRotation();
    Json r = Json::array();
    r.push_back(q.x); r.push_back(q.y); r.push_back(q.z); r.push_back(q.w);
    t["rotation"] = r;
    t["scale"] = vec3ToJson(node->getScale());
    j["transformation"] = t;
    if (!node->getSpheres().empty())
    {
        Json spheres = Json::array();
        for (const Sphere& s : node->getSpheres())
        {
            Json js = Json::object();
            js["center"] = vec3ToJson(s.center);
            js["radius"] = s.radius;
            js["materialIndex"] = s.id;
            spheres.push_back(js);
        }
        j["spheres"] = spheres;
    }
    if (!node->getChildren().empty())
    {
        Json children = Json::array();
        for (const auto& c : node->getChildren())
            children.push_back(serializeSceneNode(c));
        j["children"] = children;
    }
    return j;
}

inline Json serializeScene(const Scene& scene) // Utils.cpp:3-51
{
    Json j = Json::object();
    Json cam = Json::object();
    cam["position"] = vec3ToJson(scene.camera.getPosition());
    cam["direction"] = vec3ToJson(scene.camera.getDirection());
    cam["fov"] = scene.camera.getFov();
    j["camera"] = cam;
    if (scene.rootNode)
    {
        scene.rootNode->updateGlobalTransform();
        j["sceneGraph"] = serializeSceneNode(scene.rootNode);
    }
    if (!scene.materials.empty())
    {
        Json mats = Json::array();
        for (const Material& m : scene.materials)
        {
            Json jm = Json::object();
            jm["albedo"] = vec3ToJson(m.albedo);
            jm["roughness"] = m.roughness;
            jm["metallic"] = m.metallic;
            jm["F0"] = vec3ToJson(m.F0);
            jm["emissionIntensity"] = m.emissionIntensity;
            jm["emissionColor"] = vec3ToJson(m.emissionColor);
            mats.push_back(jm);
        }
        j["materials"] = mats;
    }
    if (!scene.lights.empty())
    {
        Json lights = Json::array();
        for (const Light& l : scene.lights)
        {
            Json jl = Json::object();
            jl["position"] = vec3ToJson(l.position);
            jl["intensity"] = l.intensity;
            jl["color"] = vec3ToJson(l.color);
            lights.push_back(jl);
        }
        j["lights"] = lights;
    }
    Json s = Json::object();
    s["maxBounces"] = scene.settings.maxBounces;
    s["skyLight"] = scene.settings.skyLight;
    s["accumulation"] = scene.settings.accumulation;
    j["settings"] = s;
    return j;
}

inline void exportScene(const Scene& scene, const std::string& filename) // Utils.cpp:53-62
{
    std::ofstream file(filename);
    if (!file.is_open())
    {
        std::cerr << "Failed to open file for writing: " << filename << std::endl;
        return;
    }
    file << serializeScene(scene).dump(4);
}

inline void deserializeSceneNode(const Json& j, std::shared_ptr<SceneNode>& node) // Utils.cpp:139-173
{
    if (!node) // an existing node (the scene root) keeps its own name
        node = std::make_shared<SceneNode>(j["name"].getString());
    const Json& t = j["transformation"];
    node->setPosition(vec3FromJson(t["position"]));
    const Json& r = t["rotation"]; // file order x, y, z, w; glm::quat(w, x, y, z) (Utils.cpp:145)
    node->setRotation(quat(r[3].getFloat(), r[0].getFloat(), r[1].getFloat(), r[2].getFloat()));
    node->setScale(vec3FromJson(t["scale"]));
    if (j.contains("spheres"))
        for (const Json& s : j["spheres"].items())
            node->addSphere(Sphere(vec3FromJson(s["center"]), s["radius"].getFloat(), s["materialIndex"].getInt()));
    if (j.contains("children"))
        for (const Json& c : j["children"].items())
        {
            std::shared_ptr<SceneNode> child;
            deserializeSceneNode(c, child);
            node->addChild(child);
        }
}

inline Scene deserializeScene(const Json& j) // Utils.cpp:97-137
{
    Scene scene; // default camera: setters only, no matrix update (quirk Q-cam ii)
    scene.camera.setPosition(vec3FromJson(j["camera"]["position"]));
    scene.camera.setDirection(vec3FromJson(j["camera"]["direction"]));
    scene.camera.setFov(j["camera"]["fov"].getFloat());
    if (j.contains("sceneGraph"))
    {
        deserializeSceneNode(j["sceneGraph"], scene.rootNode);
        scene.rootNode->updateGlobalTransform();
    }
    for (const Json& m : j["materials"].items())
    {
        Material mat;
        mat.albedo = vec3FromJson(m["albedo"]);
        mat.roughness = m["roughness"].getFloat();
        mat.metallic = m["metallic"].getFloat();
        mat.F0 = vec3FromJson(m["F0"]);
        mat.emissionIntensity = m["emissionIntensity"].getFloat();
        mat.emissionColor = vec3FromJson(m["emissionColor"]);
        scene.materials.push_back(mat); // Material::id is not serialised: stays 0
    }
    for (const Json& l : j["lights"].items())
    {
        Light light;
        light.position = vec3FromJson(l["position"]);
        light.intensity = l["intensity"].getFloat();
        light.color = vec3FromJson(l["color"]);
        scene.lights.push_back(light);
    }
    const Json& s = j["settings"];
    scene.settings.maxBounces = s["maxBounces"].getInt();
    scene.settings.skyLight = s["skyLight"].getBool();
    scene.settings.accumulation = s["accumulation"].getBool();
    return scene;
}

inline Scene importScene(const std::string& filename) // Utils.cpp:175-187: a missing file gives an empty Scene()
{
    std::ifstream file(filename);
    if (!file.is_open())
        return Scene();
    std::stringstream ss;
    ss << file.rdbuf();
    return deserializeScene(Json::parse(ss.str()));
}
} // namespace Utils

// ---- Ataraxia: the application layer of Engine/src/main.cpp:8-283, without the window ------------------
// What the application does AROUND Renderer::Render: camera motion resets the accumulation (main.cpp:22-32),
// Render() = onResize + camera.Resize + Renderer::Render with the wall-clock "Last Render Time" (:211-220),
// scene import/export (:196-209), and the UI widgets of onGuiRender (:34-176) as methods with the widget's own
// reset behaviour: node, sphere and camera edits call resetFrameIndex(); material, light, "Sky Light" and
// "Ray Depth" edits do not, and since the scene is only re-uploaded when frameIndex == 1 (Renderer.cu:175-179)
// material and light edits stay invisible until the next reset. eagerEdits = true resets on those too.
class Ataraxia
{
public:
    explicit Ataraxia(int device = 0, bool eagerEdits = false) : m_renderer(device), m_camera(45.0f, 0.1f, 100.0f), m_eagerEdits(eagerEdits)
    {
        m_scene.camera = m_camera;
        m_scene.settings = m_renderer.getSettings();
        initializeScene();
    }

    void onUpdate(float ts, const InputState& input = InputState()) // main.cpp:22-32
    {
        if (m_camera.onUpdate(ts, input))
        {
            m_renderer.resetFrameIndex();
            m_scene.camera = m_camera;
            m_scene.settings = m_renderer.getSettings();
        }
        m_scene.rootNode->updateGlobalTransform();
    }
    void setViewport(uint32_t width, uint32_t height) { m_viewportWidth = width; m_viewportHeight = height; } // main.cpp:181-182
    void Render(uint32_t frames = 1) // main.cpp:211-220
    {
        const auto t0 = std::chrono::steady_clock::now();
        m_renderer.onResize(m_viewportWidth, m_viewportHeight);
        m_camera.Resize(m_viewportWidth, m_viewportHeight);
        m_renderer.Render(m_camera, m_scene, frames);
        m_lastRenderTime = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    void ImportScene(const std::string& path = "scene.json") // main.cpp:196-202
    {
        m_scene = Utils::importScene(path);
        m_camera = m_scene.camera;
        m_renderer.setSettings(m_scene.settings);
        m_renderer.resetFrameIndex();
    }
    void ExportScene(const std::string& path = "scene.json") // main.cpp:204-209
    {
        m_scene.camera = m_camera;
        m_scene.settings = m_renderer.getSettings();
        Utils::exportScene(m_scene, path);
    }
    Renderer& GetRenderer() { return m_renderer; }
    Scene& GetScene() { return m_scene; }
    Camera& GetCamera() { return m_camera; }
    void SetScene(const Scene& scene) { m_scene = scene; }
    float lastRenderTimeMs() const { return m_lastRenderTime; }

    // "Settings" window (main.cpp:38-66)
    void setAccumulation(bool on) { Settings s = m_renderer.getSettings(); s.accumulation = on; m_renderer.setSettings(s); }
    void resetFrameIndex() { m_renderer.resetFrameIndex(); }
    void setSkyLight(bool on) { Settings s = m_renderer.getSettings(); s.skyLight = on; m_renderer.setSettings(s); edited(); }
    void setMaxBounces(int n) { Settings s = m_renderer.getSettings(); s.maxBounces = std::max(1, std::min(500, n)); m_renderer.setSettings(s); edited(); }
    void setFov(float fov)
    {
        m_camera = Camera(fov, 0.1f, 100.0f, m_camera.getPosition(), m_camera.getDirection());
        m_scene.camera = m_camera;
        m_renderer.resetFrameIndex();
    }
    void resetCamera()
    {
        m_camera = Camera(45.0f, 0.1f, 100.0f);
        m_scene.camera = m_camera;
        m_renderer.resetFrameIndex();
    }
    // "Hierarchy" window (main.cpp:75-143) and the "Add" menu (:292-300)
    void setNodePosition(SceneNode& node, const vec3& p) { node.setPosition(p); m_renderer.resetFrameIndex(); }
    void setNodeRotation(SceneNode& node, const quat& q) { node.setRotation(q); m_renderer.resetFrameIndex(); }
    void setNodeScale(SceneNode& node, const vec3& s) { node.setScale(s); m_renderer.resetFrameIndex(); }
    void removeNode(const std::shared_ptr<SceneNode>& node) { m_scene.rootNode->removeChild(node); m_renderer.resetFrameIndex(); }
    void setSphereCenter(SceneNode& node, size_t i, const vec3& c) { const_cast<Sphere&>(node.getSpheres()[i]).center = c; m_renderer.resetFrameIndex(); }
    void setSphereRadius(SceneNode& node, size_t i, float r) { const_cast<Sphere&>(node.getSpheres()[i]).radius = r; m_renderer.resetFrameIndex(); }
    void setSphereMaterial(SceneNode& node, size_t i, int m) { const_cast<Sphere&>(node.getSpheres()[i]).id = m; m_renderer.resetFrameIndex(); }
    void addSphere() { m_scene.rootNode->addSphere(Sphere(vec3(0.0f), 1.0f, 0)); edited(); }
    // "Material settings" / "Light settings" (main.cpp:145-176): edit through GetScene(), then
    void materialOrLightEdited() { edited(); }

private:
    void edited() { if (m_eagerEdits) m_renderer.resetFrameIndex(); }
    void initializeScene() // main.cpp:234-265
    {
        std::shared_ptr<SceneNode> root = m_scene.rootNode;
        root->addSphere(Sphere(vec3(0.0f, 0.0f, 0.0f), 1.0f, 0));
        auto childNode1 = std::make_shared<SceneNode>("ChildNode1");
        childNode1->setPosition(vec3(2.0f, 0.0f, 0.0f));
        childNode1->addSphere(Sphere(vec3(0.0f, 0.0f, 0.0f), 1.0f, 1));
        root->addChild(childNode1);
        auto grandChildNode = std::make_shared<SceneNode>("GrandChildNode");
        grandChildNode->setPosition(vec3(0.0f, 2.0f, 0.0f));
        grandChildNode->addSphere(Sphere(vec3(0.0f, 0.0f, 0.0f), 1.0f, 2));
        childNode1->addChild(grandChildNode);
        m_scene.materials.push_back(Material(vec3(1.022f, 0.782f, 0.344f), 1.0f, 0.0f, vec3(0.0f), 0.0f, 0));
        m_scene.materials.push_back(Material(vec3(1.0f, 0.0f, 0.0f), 0.3f, 0.0f, vec3(0.0f), 0.0f, 1));
        m_scene.materials.push_back(Material(vec3(0.972f, 0.960f, 0.915f), 0.25f, 1.0f, vec3(0.0f), 0.0f, 2));
        m_scene.lights.push_back(Light(vec3(10.0f, 10.0f, 0.0f), vec3(1.0f), 1.0f));
    }

    Scene m_scene;
    Renderer m_renderer;
    Camera m_camera;
    uint32_t m_viewportWidth = 0, m_viewportHeight = 0;
    float m_lastRenderTime = 0.0f;
    bool m_eagerEdits = false;
};
} // namespace ataraxia
