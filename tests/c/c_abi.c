/* The C-ABI header is plain C (C11, -pedantic): what a cgo / JNI / ctypes binding sees. Compiled and linked by
 * tests/test_host_api.py; run, it exercises only calls that need no GPU. */
#include <ataraxia_b200.h>
#include <stdio.h>
#include <string.h>

int main(void)
{
    atx_camera_input in = { ATX_KEY_W | ATX_KEY_D, 1u, 3.0f, -2.0f };
    float pos[3] = { 0.0f, 0.0f, 3.0f }, dir[3] = { 0.0f, 0.0f, -1.0f }, last[2] = { 0.0f, 0.0f };
    int moved = 0;
    float proj[16], view[16], iproj[16], iview[16];
    if (atx_host_camera_update(pos, dir, last, &in, 0.1f, &moved) != ATX_OK || !moved)
        return 1;
    if (atx_host_camera_matrices(pos, dir, 45.0f, 0.1f, 100.0f, 64, 36, proj, view, iproj, iview) != ATX_OK)
        return 2;
    if (atx_host_camera_update(NULL, dir, last, &in, 0.1f, &moved) != ATX_ERR_INVALID || strlen(atx_last_error()) == 0)
        return 3;
    printf("%s %.6f %.6f %.6f\n", atx_version(), pos[0], pos[1], pos[2]);
    return 0;
}
