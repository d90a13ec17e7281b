// TEST INFRASTRUCTURE (oracle build only). The reference's Input class is
// GLFW-backed (Core/src/input/Input.cpp:9-35); headless there is no window, so the
// four queries answer from a record the harness sets (refinput_set): by default
// "nothing pressed", and a scripted key/mouse state when a test drives
// Camera::onUpdate (Camera.cpp:30-108) through the unmodified reference code.
#include "input/Input.h"

namespace
{
struct ScriptedInput
{
    bool keys[512] = {};
    bool rightButton = false;
    float mouseX = 0.0f, mouseY = 0.0f;
} g_input;
}

// keyCodes: GLFW key codes (Core/include/input/KeyCodes.h) that are held down
extern "C" __attribute__((visibility("default"))) void refinput_set(const uint16_t* keyCodes, int nKeys, int rightButton,
                                                                      float mouseX, float mouseY)
{
    for (bool& k : g_input.keys) k = false;
    for (int i = 0; i < nKeys; i++)
        if (keyCodes[i] < 512) g_input.keys[keyCodes[i]] = true;
    g_input.rightButton = rightButton != 0;
    g_input.mouseX = mouseX;
    g_input.mouseY = mouseY;
}

namespace Input
{
    bool Input::IsKeyPressed(KeyCode key) { return static_cast<uint16_t>(key) < 512 && g_input.keys[static_cast<uint16_t>(key)]; }
    bool Input::IsMouseButtonPressed(MouseButton button) { return button == MouseButton::Right && g_input.rightButton; }
    glm::vec2 Input::GetMousePosition() { return { g_input.mouseX, g_input.mouseY }; }
    void Input::SetCursorMode(CursorMode) {}
}
