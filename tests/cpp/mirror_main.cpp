// Test driver for the header-only C++ mirror (include/ataraxia/Ataraxia.h). Reads like reference user code:
// the same calls Ataraxia::Render() makes (Engine/src/main.cpp:211-220) with the canonical headless
// camera protocol (SURVEY.md §8 Q-cam). Driven by tests/test_cpp_mirror.py.
//   mirror_main cpu <scene.json> <export.json>        -> prints flattened spheres, matrices, rays as JSON
//   mirror_main gpu <scene.json> W H bounces sky frames <out.bin>  -> hits(int32) acc1(f32) accK(f32) rgbaK(u32)
//   mirror_main walk <steps.bin> <out.bin> px py pz dx dy dz fov   -> scripted Camera::onUpdate, per-step camera state
//   mirror_main app <scene.json> <out.bin>                          -> the Ataraxia layer: frame-index behaviour of edits
//   mirror_main image <prefix>                                      -> a test pattern through Image::savePPM / savePNG
#include <ataraxia/Ataraxia.h>
#include <clocale>
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

// scripted Camera::onUpdate: steps file = records of (dt, mouseX, mouseY: f32; keys [W=1 S=2 A=4 D=8 Q=16 E=32], right: u32);
// output file = per step position(3) direction(3) inverseView(16) moved(1, as float)
static int cameraWalk(const char* stepsPath, const char* outPath, const float* start)
{
    struct Step { float dt, mx, my; uint32_t keys, right; };
    FILE* f = std::fopen(stepsPath, "rb");
    if (!f)
        return 3;
    std::vector<Step> steps;
    Step st;
    while (std::fread(&st, sizeof(st), 1, f) == 1)
        steps.push_back(st);
    std::fclose(f);
    Camera cam(start[6], 0.1f, 100.0f, vec3(start[0], start[1], start[2]), vec3(start[3], start[4], start[5]));
    cam.Resize(48, 27);
    FILE* o = std::fopen(outPath, "wb");
    if (!o)
        return 3;
    for (const Step& s : steps)
    {
        InputState in;
        in.W = s.keys & 1u; in.S = s.keys & 2u; in.A = s.keys & 4u; in.D = s.keys & 8u; in.Q = s.keys & 16u; in.E = s.keys & 32u;
        in.rightButton = s.right != 0;
        in.mouse = vec2(s.mx, s.my);
        const float moved = cam.onUpdate(s.dt, in) ? 1.0f : 0.0f;
        std::fwrite(&cam.getPosition().x, 4, 3, o);
        std::fwrite(&cam.getDirection().x, 4, 3, o);
        std::fwrite(&cam.getInverseViewMatrix()[0].x, 4, 16, o);
        std::fwrite(&moved, 4, 1, o);
    }
    const auto& rays = cam.getRayDirection();
    std::fwrite(&rays[0].x, 4, rays.size() * 3, o);
    std::fclose(o);
    return 0;
}

// the application layer without the window (main.cpp:8-283): frame-index behaviour of camera motion and UI edits
static int appSession(const char* scenePath, const char* outPath)
{
    Ataraxia app;
    app.setViewport(64, 36);
    app.Render();                                   // default scene of initializeScene()
    const uint32_t f0 = app.GetRenderer().frameIndex();       // 2
    app.onUpdate(0.016f);                           // no input: nothing moves
    app.Render(3);
    const uint32_t f1 = app.GetRenderer().frameIndex();       // 5
    InputState in;
    in.rightButton = true; in.W = true; in.mouse = vec2(12.0f, -7.0f);
    app.onUpdate(0.016f, in);                       // camera moved: accumulation restarts
    const uint32_t f2 = app.GetRenderer().frameIndex();       // 1
    app.Render(2);
    app.GetScene().materials[0].albedo = vec3(0.1f, 0.9f, 0.1f);
    app.materialOrLightEdited();                    // reference behaviour: no reset
    const uint32_t f3 = app.GetRenderer().frameIndex();       // 3
    app.setNodePosition(*app.GetScene().rootNode->getChildren()[0], vec3(2.5f, 0.0f, 0.0f));
    const uint32_t f4 = app.GetRenderer().frameIndex();       // 1
    app.ImportScene(scenePath);
    app.Render(2);
    const std::vector<float> acc = app.GetRenderer().getAccumulation();
    FILE* o = std::fopen(outPath, "wb");
    if (!o)
        return 3;
    std::fwrite(acc.data(), 4, acc.size(), o);
    std::fclose(o);
    std::printf("%u %u %u %u %u %u %.3f\n", f0, f1, f2, f3, f4, app.GetRenderer().frameIndex(), app.lastRenderTimeMs() >= 0.0f ? 1.0 : 0.0);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 3)
        return 2;
    const std::string mode = argv[1];
    if (mode == "walk" && argc >= 11)
    {
        float start[7];
        for (int i = 0; i < 7; i++)
            start[i] = static_cast<float>(std::atof(argv[4 + i]));
        return cameraWalk(argv[2], argv[3], start);
    }
    if (mode == "app" && argc >= 4)
        return appSession(argv[2], argv[3]);
    if (mode == "image")
    {
        // a 37x11 test pattern (RGBA8 as the renderer packs it) through both image sinks
        const uint32_t W = 37, H = 11;
        std::vector<uint32_t> px(W * H);
        for (uint32_t y = 0; y < H; y++)
            for (uint32_t x = 0; x < W; x++)
                px[y * W + x] = 0xFF000000u | ((x * 7u) & 0xFFu) | (((y * 23u) & 0xFFu) << 8) | ((((x + y) * 5u) & 0xFFu) << 16);
        Image img(W, H, ImageType::RGBA, px.data());
        return img.savePPM(std::string(argv[2]) + ".ppm") && img.savePNG(std::string(argv[2]) + ".png") ? 0 : 3;
    }
    if (mode == "json")
    {
        // the JSON reader/writer on its own, under a comma-decimal locale if the host has one (argv[2] = locale name):
        // numbers are read and written with '.', non-JSON numbers are refused, control characters are escaped,
        // nesting is capped
        const bool localeSet = std::setlocale(LC_ALL, argv[2]) != nullptr;
        int bad = 0;
        const atx::Json j = atx::Json::parse("{\"a\": [0.5, -1.25e-3, 100, 1e+20, 3.0e-7], \"s\": \"x\\u0001\\b\\fy\"}");
        bad += j["a"][0].number() != 0.5 || j["a"][1].number() != -1.25e-3 || j["a"][2].getInt() != 100 || j["a"][4].number() != 3.0e-7;
        const std::string out = j.dump();
        bad += out != "{\"a\":[0.5,-0.00125,100,1e+20,3e-07],\"s\":\"x\\u0001\\b\\fy\"}";
        bad += atx::Json::parse(out).dump() != out;
        for (const char* text : { "inf", "nan", "+1", "0x10", ".5", "1.", "01", "1e", "-", "[1,]" })
        {
            try { atx::Json::parse(text); bad++; std::fprintf(stderr, "accepted %s\n", text); }
            catch (const std::exception&) {}
        }
        std::string deep(100000, '[');
        try { atx::Json::parse(deep); bad++; }
        catch (const std::exception&) {}
        std::printf("%d %d %s\n", bad, localeSet ? 1 : 0, out.c_str());
        return bad ? 4 : 0;
    }
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
