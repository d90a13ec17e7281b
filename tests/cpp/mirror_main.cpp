// Test driver for the header-only C++ mirror (include/ataraxia/Ataraxia.h). Reads like reference user code:
// the same calls Ataraxia::Render() makes (Engine/src/main.cpp:211-220) with the canonical headless
// camera protocol (SURVEY.md §8 Q-cam). Driven by tests/test_cpp_mirror.py.
//   mirror_main cpu <scene.json> <export.json>        -> prints flattened spheres, matrices, rays as JSON
//   mirror_main gpu <scene.json> W H bounces sky frames <out.bin>  -> hits(int32) acc1(f32) accK(f32) rgbaK(u32)
#include <ataraxia/Ataraxia.h>
#include <cstdio>
#include <cstdlib>

using namespace ataraxia;

static void printMat(const char* name, const mat4& m, bool last = false)
{
    std::printf("\"%s\": [", name);
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++)
            std::printf("%.9e%s", m[c][r], (c == 3 && r == 3) ? "" : ", ");
    std::printf("]%s\n", last ? "" : ",");
}

int main(int argc, char** argv)
{
    if (argc < 3)
        return 2;
    const std::string mode = argv[1];
    Scene scene = Utils::importScene(argv[2]);
    if (mode == "cpu")
    {
        std::vector<Sphere> spheres;
        Renderer::traverseSceneGraph(scene.rootNode, mat4(1.0f), spheres);
        Camera cam(scene.camera.getFov(), 0.1f, 100.0f, scene.camera.getPosition(), scene.camera.getDirection());
        cam.Resize(160, 90);
        std::printf("{\n\"spheres\": [");
        for (size_t i = 0; i < spheres.size(); i++)
            std::printf("[%.9e, %.9e, %.9e, %.9e, %d]%s", spheres[i].center.x, spheres[i].center.y, spheres[i].center.z, spheres[i].radius,
                        spheres[i].id, i + 1 < spheres.size() ? ", " : "");
        std::printf("],\n\"materials\": %zu, \"lights\": %zu, \"maxBounces\": %d, \"rootName\": \"%s\",\n", scene.materials.size(),
                    scene.lights.size(), scene.settings.maxBounces, scene.rootNode->getName().c_str());
        printMat("projection", cam.getProjectionMatrix());
        printMat("view", cam.getViewMatrix());
        printMat("inverseProjection", cam.getInverseProjectionMatrix());
        printMat("inverseView", cam.getInverseViewMatrix());
        const auto& rays = cam.getRayDirection();
        std::printf("\"ray0\": [%.9e, %.9e, %.9e], \"rayLast\": [%.9e, %.9e, %.9e], \"nRays\": %zu\n}\n", rays[0].x, rays[0].y, rays[0].z,
                    rays.back().x, rays.back().y, rays.back().z, rays.size());
        if (argc > 3)
            Utils::exportScene(scene, argv[3]);
        return 0;
    }
    if (mode == "gpu" && argc >= 9)
    {
        const uint32_t W = std::atoi(argv[3]), H = std::atoi(argv[4]);
        const int bounces = std::atoi(argv[5]), sky = std::atoi(argv[6]), frames = std::atoi(argv[7]);
        Camera cam(scene.camera.getFov(), 0.1f, 100.0f, scene.camera.getPosition(), scene.camera.getDirection());
        Renderer renderer;
        Settings st;
        st.accumulation = true; st.skyLight = sky != 0; st.maxBounces = bounces;
        renderer.setSettings(st);
        renderer.onResize(W, H);              // main.cpp:215
        cam.Resize(W, H);                     // main.cpp:216
        renderer.Render(cam, scene);          // main.cpp:217
        const std::vector<int32_t> hits = renderer.getHitIds();
        const std::vector<float> acc1 = renderer.getAccumulation();
        for (int k = 1; k < frames; k++)
            renderer.Render(cam, scene);
        const std::vector<float> accK = renderer.getAccumulation();
        FILE* f = std::fopen(argv[8], "wb");
        if (!f)
            return 3;
        std::fwrite(hits.data(), sizeof(int32_t), hits.size(), f);
        std::fwrite(acc1.data(), sizeof(float), acc1.size(), f);
        std::fwrite(accK.data(), sizeof(float), accK.size(), f);
        std::fwrite(renderer.getImage()->getPixels(), sizeof(uint32_t), static_cast<size_t>(W) * H, f);
        std::fclose(f);
        if (argc > 9)
        {
            renderer.getImage()->savePPM(std::string(argv[9]) + ".ppm");
            renderer.saveAccumulationPFM(std::string(argv[9]) + ".pfm");
        }
        std::printf("frameIndex %u paths %llu\n", renderer.frameIndex(), static_cast<unsigned long long>(renderer.counters().paths));
        return renderer.frameIndex() == static_cast<uint32_t>(frames) + 1 ? 0 : 4;
    }
    return 2;
}
