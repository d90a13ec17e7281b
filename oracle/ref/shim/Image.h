// TEST INFRASTRUCTURE (oracle build only). Headless stand-in that shadows the
// reference's Core/include/Image.h (which pulls in vulkan/vulkan.h). Only the
// four members Renderer touches are provided: ctor (Renderer.cu:106/121),
// getWidth/getHeight (:100) and setData (:242). No Vulkan, no display.
#pragma once
#include <cstdint>
#include <string_view>

enum class ImageType { None = 0, RGBA, RGBA32F };

class Image
{
public:
    Image(uint32_t width, uint32_t height, ImageType type, const void* data = nullptr)
        : m_width(width), m_height(height), m_type(type), m_last(data) {}
    void setData(const void* data) { m_last = data; }
    uint32_t getWidth() const { return m_width; }
    uint32_t getHeight() const { return m_height; }
    const void* lastData() const { return m_last; }
private:
    uint32_t m_width = 0, m_height = 0;
    ImageType m_type = ImageType::None;
    const void* m_last = nullptr;
};
