// TEST INFRASTRUCTURE (oracle build only). The reference's Input class is
// GLFW-backed (Core/src/input/Input.cpp:9-35); headless there is no window, so
// every query answers "nothing pressed". Camera::onUpdate is never called by
// the harness; these exist only to satisfy the linker.
#include "input/Input.h"

namespace Input
{
    bool Input::IsKeyPressed(KeyCode) { return false; }
    bool Input::IsMouseButtonPressed(MouseButton) { return false; }
    glm::vec2 Input::GetMousePosition() { return { 0.0f, 0.0f }; }
    void Input::SetCursorMode(CursorMode) {}
}
