// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Drives examples/dropin/Renderer.cpp (the drop-in Renderer translation unit) exactly as ref_harness.cu drives
// the unmodified reference renderer: the same scene import (the reference's Utils.cpp), the same camera (its
// Camera.cpp), the application's three calls per frame (Engine/src/main.cpp:215-217), the same dump files
//   <prefix>.acc<k>.f32  W*H*4   accumulation buffer after frame k (through the C-ABI handle behind the Renderer)
//   <prefix>.rgba<k>.u32 W*H     the Image the Renderer hands to the application after frame k
// so that tests/test_dropin.py can diff the two binaries' outputs byte for byte.
//
// usage: ref_headless_shim scene.json W H maxBounces skyLight frames prefix [k1,k2,...]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <string>
#include <vector>
#include <algorithm>
#include <glm/gtc/quaternion.hpp>
#include <nlohmann/json.hpp>
#include "Renderer.h"
#include "Utils.h"
#include <ataraxia_b200.h>

atx_handle ataraxia_b200_backend(const Renderer* r); // examples/dropin/Renderer.cpp

template <typename T>
static void dump(const std::string& path, const T* data, size_t count)
{
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) { std::fprintf(stderr, "cannot write %s\n", path.c_str()); std::exit(2); }
    std::fwrite(data, sizeof(T), count, f);
    std::fclose(f);
}

int main(int argc, char** argv)
{
    if (argc < 8)
    {
        std::fprintf(stderr, "usage: %s scene.json W H maxBounces skyLight frames prefix [k1,k2,...]\n", argv[0]);
        return 2;
    }
    const uint32_t W = std::atoi(argv[2]), H = std::atoi(argv[3]);
    const int frames = std::atoi(argv[6]);
    const std::string prefix = argv[7];
    std::set<int> dumpAt;
    if (argc > 8)
        for (char* tok = std::strtok(argv[8], ","); tok; tok = std::strtok(nullptr, ","))
            dumpAt.insert(std::atoi(tok));

    Scene scene = Utils::importScene(argv[1]);
    Camera cam(scene.camera.getFov(), 0.1f, 100.0f, scene.camera.getPosition(), scene.camera.getDirection()); // main.cpp:52
    Settings st;
    st.accumulation = true;
    st.skyLight = std::atoi(argv[5]) != 0;
    st.maxBounces = std::atoi(argv[4]);

    Renderer r;
    r.setSettings(st);
    const size_t P = static_cast<size_t>(W) * H;
    std::vector<float> acc(P * 4);
    std::vector<double> ms;
    for (int k = 1; k <= frames; k++)
    {
        const auto t0 = std::chrono::steady_clock::now();
        r.onResize(W, H);           // main.cpp:215
        cam.Resize(W, H);           // :216
        r.Render(cam, scene);       // :217
        ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        if (prefix != "-" && dumpAt.count(k))
        {
            if (atx_read_accum(ataraxia_b200_backend(&r), acc.data()) != ATX_OK)
            {
                std::fprintf(stderr, "%s\n", atx_last_error());
                return 1;
            }
            dump(prefix + ".acc" + std::to_string(k) + ".f32", acc.data(), P * 4);
            dump(prefix + ".rgba" + std::to_string(k) + ".u32", static_cast<const uint32_t*>(r.getImage()->lastData()), P);
        }
    }
    std::sort(ms.begin(), ms.end());
    atx_counters c{};
    atx_get_counters(ataraxia_b200_backend(&r), &c);
    std::printf("{\"impl\": \"dropin-shim\", \"width\": %u, \"height\": %u, \"frames\": %d, \"median_frame_ms\": %.4f, \"min_frame_ms\": %.4f, "
                "\"paths\": %llu, \"kernel_launches\": %llu}\n", W, H, frames, ms.empty() ? 0.0 : ms[ms.size() / 2], ms.empty() ? 0.0 : ms[0],
                static_cast<unsigned long long>(c.paths), static_cast<unsigned long long>(c.launches));
    return c.paths == static_cast<unsigned long long>(P) * frames ? 0 : 1;
}
