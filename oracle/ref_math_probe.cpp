// ref_math_probe — runs the REFERENCE's own Math library (compiled from /root/reference/Math where it
// lies; see oracle/Makefile target `_ref/ref_math_probe`) to produce the constant blocks the hot path
// consumes, exactly as the reference's host code builds them:
//
//   view   : Foreground/SceneGraph/SceneView.cpp:27-30
//              ViewMat = node.GetWorldTransform().Inverse().ToMatrix4().Transpose()
//            node world transform = tc::Matrix3x4(Translation, Rotation, Scale)
//              (Foreground/SceneGraph/SceneNode.cpp:245-248), rotation from Euler degrees
//              (Math/Quaternion.cpp:49-66), as App/MainBehaviour.cpp:26-64 sets the nodes up.
//            InvModelView = ViewMat.Inverse()          (Foreground/Renderer/MegaPipeline.cpp:242)
//            light direction = WorldTransform * (0,0,-1,0)   (MegaPipeline.cpp:110)
//   persp  : Foreground/SceneGraph/Camera.cpp:85-102, ProjMat = M.Transpose(), InvProj = M.Inverse().Transpose()
//   ortho  : Foreground/SceneGraph/Camera.cpp:104-111
//
// Output: one line per block, `name: f0 f1 ... f15` with C99 hex floats in MEMORY order (what the
// reference uploads, i.e. what GLSL reads as a column-major mat4).  This is test infrastructure: it
// pins final184_b200/scene.py's numpy mirror of the same arithmetic (tests/test_constants.py via
// tests/golden/ref_constants.json).  It is never linked into the product.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "Matrix3x4.h"
#include "Matrix4.h"
#include "Quaternion.h"
#include "Vector3.h"
#include "Vector4.h"

static void dump(const char* name, const tc::Matrix4& m)
{
    const float* f = m.Data();
    printf("%s:", name);
    for (int i = 0; i < 16; i++) printf(" %a", f[i]);
    printf("\n");
}

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    if (!strcmp(argv[1], "view") && argc == 11)
    {
        float a[9];
        for (int i = 0; i < 9; i++) a[i] = (float)atof(argv[2 + i]);
        tc::Quaternion rot(a[0], a[1], a[2]);
        tc::Matrix3x4 world(tc::Vector3(a[3], a[4], a[5]), rot, tc::Vector3(a[6], a[7], a[8]));
        tc::Matrix4 view = world.Inverse().ToMatrix4().Transpose();
        dump("WorldMat", world.ToMatrix4().Transpose());
        dump("ViewMat", view);
        dump("InvView", view.Inverse());
        tc::Vector3 dir = world * tc::Vector4(0.0f, 0.0f, -1.0f, 0.0f);
        printf("Forward: %a %a %a\n", dir.x, dir.y, dir.z);
        return 0;
    }
    if (!strcmp(argv[1], "persp") && argc == 6)
    {
        float FovY = (float)atof(argv[2]), AspectRatio = (float)atof(argv[3]);
        float NearClip = (float)atof(argv[4]), FarClip = (float)atof(argv[5]);
        // Arithmetic of CCamera::CalcPerspective, evaluated with the reference's tc::M_PI and Matrix4.
        float rad_fovy = FovY / 180.f * tc::M_PI;
        float y_slope = tan(rad_fovy / 2);
        float x_slope = y_slope * AspectRatio;
        float right = x_slope * NearClip;
        float top = y_slope * NearClip;
        float pa = NearClip / right;
        float pb = NearClip / top;
        float pc = (FarClip + NearClip) / (NearClip - FarClip);
        float pd = -2 * FarClip * NearClip / (FarClip - NearClip);
        tc::Matrix4 m(pa, 0, 0, 0, 0, -pb, 0, 0, 0, 0, pc, pd, 0, 0, -1, 0);
        dump("ProjMat", m.Transpose());
        dump("InvProj", m.Inverse().Transpose());
        return 0;
    }
    if (!strcmp(argv[1], "ortho") && argc == 6)
    {
        float MagX = (float)atof(argv[2]), MagY = (float)atof(argv[3]);
        float NearClip = (float)atof(argv[4]), FarClip = (float)atof(argv[5]);
        float oa = 1.0f / MagX;
        float ob = 1.0f / MagY;
        float oc = 1.0f / (NearClip - FarClip);
        float od = NearClip / (NearClip - FarClip);
        tc::Matrix4 m(oa, 0, 0, 0, 0, ob, 0, 0, 0, 0, oc, od, 0, 0, 0, 1);
        dump("ProjMat", m.Transpose());
        dump("InvProj", m.Inverse().Transpose());
        return 0;
    }
    fprintf(stderr, "usage: %s view ex ey ez tx ty tz sx sy sz | persp fovy aspect near far | ortho magx magy near far\n", argv[0]);
    return 2;
}
