// render_scene — headless use of the C++ mirror (include/ataraxia/Ataraxia.h): what the reference application does
// per UI frame (Engine/src/main.cpp:211-220: onResize, Camera::Resize, Renderer::Render), without the window.
//
//   g++ -std=c++17 -O2 -Iinclude examples/render_scene.cpp -o render_scene -Lataraxia_b200/lib -lataraxia_b200 -Wl,-rpath,$PWD/ataraxia_b200/lib
//   ./render_scene scene.json out.png [--width 1280] [--height 720] [--spp 64] [--bounces N] [--sky 0|1] [--pfm out.pfm]
//
// The scene's own settings (maxBounces, skyLight) apply unless overridden, as after "Import Scene" (main.cpp:196-202).
#include <ataraxia/Ataraxia.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace ataraxia;

int main(int argc, char** argv)
{
    if (argc < 3)
    {
        std::fprintf(stderr, "usage: %s scene.json out.png [--width W] [--height H] [--spp N] [--bounces N] [--sky 0|1] [--pfm out.pfm]\n", argv[0]);
        return 2;
    }
    uint32_t width = 1280, height = 720, spp = 64;
    int bounces = -1, sky = -1;
    std::string pfm;
    for (int i = 3; i + 1 < argc; i += 2)
    {
        if (!std::strcmp(argv[i], "--width")) width = static_cast<uint32_t>(std::atoi(argv[i + 1]));
        else if (!std::strcmp(argv[i], "--height")) height = static_cast<uint32_t>(std::atoi(argv[i + 1]));
        else if (!std::strcmp(argv[i], "--spp")) spp = static_cast<uint32_t>(std::atoi(argv[i + 1]));
        else if (!std::strcmp(argv[i], "--bounces")) bounces = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--sky")) sky = std::atoi(argv[i + 1]);
        else if (!std::strcmp(argv[i], "--pfm")) pfm = argv[i + 1];
        else { std::fprintf(stderr, "unknown option %s\n", argv[i]); return 2; }
    }

    Scene scene = Utils::importScene(argv[1]);
    if (scene.rootNode->getSpheres().empty() && scene.rootNode->getChildren().empty())
        std::fprintf(stderr, "warning: %s holds no spheres (a missing file imports as an empty scene, Utils.cpp:178-179)\n", argv[1]);

    // the camera the application builds from an imported scene (main.cpp:52)
    Camera camera(scene.camera.getFov(), 0.1f, 100.0f, scene.camera.getPosition(), scene.camera.getDirection());
    Renderer renderer;
    Settings settings = scene.settings;
    if (bounces >= 0) settings.maxBounces = bounces;
    if (sky >= 0) settings.skyLight = sky != 0;
    settings.accumulation = true;
    renderer.setSettings(settings);

    renderer.onResize(width, height);
    camera.Resize(width, height);
    renderer.Render(camera, scene, spp);           // spp frames in one launch; identical to spp Render() calls
    if (renderer.frameIndex() != spp + 1)
    {
        std::fprintf(stderr, "render failed: %s\n", atx_last_error());
        return 1;
    }
    const atx_counters c = renderer.counters();
    std::printf("%ux%u, %u spp, %d bounces: %.3f ms on the device, %.1f Mpaths/s, %.2f Grays/s\n", width, height, spp, settings.maxBounces,
                renderer.lastRenderMs(), c.paths / (renderer.lastRenderMs() * 1e3), c.rays / (renderer.lastRenderMs() * 1e6));
    if (!renderer.getImage()->savePNG(argv[2]))
    {
        std::fprintf(stderr, "cannot write %s\n", argv[2]);
        return 1;
    }
    if (!pfm.empty() && !renderer.saveAccumulationPFM(pfm))
    {
        std::fprintf(stderr, "cannot write %s\n", pfm.c_str());
        return 1;
    }
    return 0;
}
