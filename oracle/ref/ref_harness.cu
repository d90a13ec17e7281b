// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Headless driver for the UNMODIFIED reference renderer (1neskk/Ataraxia),
// compiled from the sources where they lie under /root/reference by
// oracle/ref/Makefile into oracle/_ref/ref_headless (git-ignored). It follows
// the canonical headless protocol of SURVEY.md §8 (Q-cam): the same three
// calls the application makes per frame (Engine/src/main.cpp:215-217), with a
// camera built the way main.cpp:52 builds it.
//
// It dumps what the parity tests compare against:
//   <prefix>.rays.f32    W*H*3  host ray table   (Camera.cpp:161-195)
//   <prefix>.hit.i32     W*H    primary closest-hit sphere index (-1 = miss),
//                               by calling the reference's own public
//                               Renderer::traceRay (Renderer.cu:251) on the table
//   <prefix>.spheres.f32 N*5    flattened world-space spheres (cx,cy,cz,r,id)
//   <prefix>.acc<k>.f32  W*H*4  accumulation buffer after frame k
//   <prefix>.rgba<k>.u32 W*H    packed image after frame k
// and prints one JSON line with wall-clock timings of Renderer::Render.
//
// Kernel-only time of the reference's kernelRender (BASELINE.md §4 Baseline A (ii)) WITHOUT touching
// its source: Renderer::Render launches through the runtime entry cudaLaunchKernel and then blocks on
// cudaDeviceSynchronize (Renderer.cu:223-240). This binary links the SHARED CUDA runtime and defines
// cudaLaunchKernel itself: the definition below brackets every launch made while a Render() call is in
// progress with a pair of CUDA events on the launch stream and forwards to the real entry
// (dlsym(RTLD_NEXT)). Render() launches exactly one kernel per frame, so the event pairs are the
// per-frame kernelRender times.
//
// usage: ref_headless scene.json W H maxBounces skyLight frames prefix [k1,k2,...]
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>
#include <set>
#include <string>
#include <vector>
#include "Random.h"
#include <glm/gtc/quaternion.hpp>
#include <nlohmann/json.hpp>
#define private public
#include "Renderer.h"
#undef private
#include "Utils.h"

namespace
{
bool g_timeLaunches = false;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_launchEvents;
}

extern "C" cudaError_t cudaLaunchKernel(const void* func, dim3 grid, dim3 block, void** args, size_t sharedMem, cudaStream_t stream)
{
    using Fn = cudaError_t (*)(const void*, dim3, dim3, void**, size_t, cudaStream_t);
    static Fn real = reinterpret_cast<Fn>(dlsym(RTLD_NEXT, "cudaLaunchKernel"));
    if (!real)
    {
        std::fprintf(stderr, "ref_headless: the shared CUDA runtime's cudaLaunchKernel was not found\n");
        std::exit(3);
    }
    if (!g_timeLaunches)
        return real(func, grid, block, args, sharedMem, stream);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, stream);
    const cudaError_t e = real(func, grid, block, args, sharedMem, stream);
    cudaEventRecord(b, stream);
    g_launchEvents.emplace_back(a, b);
    return e;
}

__global__ void primaryHitKernel(uint32_t width, uint32_t height, glm::vec3 origin, const glm::vec3* dirs,
    const Sphere* spheres, size_t numSpheres, int* out)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || y >= height)
        return;
    Ray ray;
    ray.origin = origin;
    ray.direction = dirs[x + y * width];
    auto ht = Renderer::traceRay(ray, spheres, numSpheres);
    out[x + y * width] = ht.t < 0.0f ? -1 : static_cast<int>(ht.id);
}

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
    const std::string scenePath = argv[1];
    const uint32_t W = std::atoi(argv[2]), H = std::atoi(argv[3]);
    const int bounces = std::atoi(argv[4]);
    const bool sky = std::atoi(argv[5]) != 0;
    const int frames = std::atoi(argv[6]);
    const std::string prefix = argv[7];
    std::set<int> dumpAt;
    if (argc > 8)
    {
        std::string s = argv[8];
        size_t p = 0;
        while (p < s.size())
        {
            size_t q = s.find(',', p);
            if (q == std::string::npos) q = s.size();
            dumpAt.insert(std::atoi(s.substr(p, q - p).c_str()));
            p = q + 1;
        }
    }
    const bool quiet = prefix == "-";

    Scene scene = Utils::importScene(scenePath);
    Camera cam(scene.camera.getFov(), 0.1f, 100.0f, scene.camera.getPosition(), scene.camera.getDirection());

    Settings st;
    st.accumulation = true;
    st.skyLight = sky;
    st.maxBounces = bounces;

    Renderer r;
    r.setSettings(st);
    r.onResize(W, H);
    cam.Resize(W, H);

    const size_t P = static_cast<size_t>(W) * H;
    std::vector<float> acc(P * 4);
    std::vector<uint32_t> rgba(P);
    std::vector<double> ms;

    const auto t_all0 = std::chrono::steady_clock::now();
    for (int k = 1; k <= frames; k++)
    {
        const auto t0 = std::chrono::steady_clock::now();
        g_timeLaunches = true;
        r.Render(cam, scene);
        g_timeLaunches = false;
        const auto t1 = std::chrono::steady_clock::now();
        ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
        if (!quiet && dumpAt.count(k))
        {
            cudaMemcpy(acc.data(), r.d_accumulation_.m_data, P * 16, cudaMemcpyDeviceToHost);
            cudaMemcpy(rgba.data(), r.d_imageData_.m_data, P * 4, cudaMemcpyDeviceToHost);
            dump(prefix + ".acc" + std::to_string(k) + ".f32", acc.data(), P * 4);
            dump(prefix + ".rgba" + std::to_string(k) + ".u32", rgba.data(), P);
        }
    }
    cudaDeviceSynchronize();
    const auto t_all1 = std::chrono::steady_clock::now();

    if (!quiet)
    {
        // primary visibility through the reference's own traceRay
        const auto& rays = cam.getRayDirection();
        dump(prefix + ".rays.f32", reinterpret_cast<const float*>(rays.data()), P * 3);

        std::vector<Sphere> flat;
        Renderer::traverseSceneGraph(scene.rootNode, glm::mat4(1.0f), flat);
        std::vector<float> sph;
        for (auto& s : flat)
        {
            int id = static_cast<uint32_t>(s.id) >= scene.materials.size() ? 0 : s.id;
            sph.insert(sph.end(), { s.center.x, s.center.y, s.center.z, s.radius, static_cast<float>(id) });
        }
        dump(prefix + ".spheres.f32", sph.data(), sph.size());

        glm::vec3* d_dirs = nullptr;
        int* d_out = nullptr;
        cudaMalloc(&d_dirs, P * sizeof(glm::vec3));
        cudaMalloc(&d_out, P * sizeof(int));
        cudaMemcpy(d_dirs, rays.data(), P * sizeof(glm::vec3), cudaMemcpyHostToDevice);
        dim3 block(16, 16), grid((W + 15) / 16, (H + 15) / 16);
        primaryHitKernel<<<grid, block>>>(W, H, cam.getPosition(), d_dirs, r.d_spheres_.m_data, r.m_numSpheres, d_out);
        std::vector<int> hit(P);
        cudaMemcpy(hit.data(), d_out, P * sizeof(int), cudaMemcpyDeviceToHost);
        dump(prefix + ".hit.i32", hit.data(), P);
        cudaFree(d_dirs);
        cudaFree(d_out);
    }

    // kernel-only: one event pair per launch made inside Render()
    std::vector<double> kms;
    for (auto& ev : g_launchEvents)
    {
        float v = 0.0f;
        if (cudaEventElapsedTime(&v, ev.first, ev.second) == cudaSuccess)
            kms.push_back(v);
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    std::vector<double> ksorted = kms;
    std::sort(ksorted.begin(), ksorted.end());
    double ksum = 0.0;
    for (double v : kms)
        ksum += v;

    std::vector<double> sorted = ms;
    std::sort(sorted.begin(), sorted.end());
    const double total = std::chrono::duration<double, std::milli>(t_all1 - t_all0).count();
    cudaError_t err = cudaGetLastError();
    std::printf("{\"impl\": \"reference-cuda\", \"width\": %u, \"height\": %u, \"frames\": %d, \"max_bounces\": %d, "
        "\"num_spheres\": %zu, \"total_ms\": %.4f, \"first_frame_ms\": %.4f, \"median_frame_ms\": %.4f, "
        "\"min_frame_ms\": %.4f, \"kernel_launches\": %zu, \"median_kernel_ms\": %.4f, \"min_kernel_ms\": %.4f, "
        "\"mean_kernel_ms\": %.4f, \"cuda_error\": \"%s\"}\n",
        W, H, frames, bounces, r.m_numSpheres, total, ms.empty() ? 0.0 : ms[0],
        sorted.empty() ? 0.0 : sorted[sorted.size() / 2], sorted.empty() ? 0.0 : sorted[0],
        kms.size(), ksorted.empty() ? 0.0 : ksorted[ksorted.size() / 2], ksorted.empty() ? 0.0 : ksorted[0],
        kms.empty() ? 0.0 : ksum / kms.size(), cudaGetErrorString(err));
    return err == cudaSuccess ? 0 : 1;
}
