// frame_driver.cpp — the voxel/indirect section of CMegaPipeline::Render() (Foreground/Renderer/MegaPipeline.cpp:195-319) written the
// way the reference writes it, against include/f184_renderer.hpp: RenderList for the voxel pass, then every full-screen pass through
// beginRender / setSampler / setImageView / setStruct / blit2d / endRender with the reference's resource names.  Test program:
// reads one binary blob of inputs (tests/test_cpp_host.py writes it), runs two frames, writes the second frame's images.
// Built twice by the tests: against the CPU oracle (-DF184_ORACLE, runs anywhere) and against libf184.so (GPU box).
#ifdef F184_ORACLE
#include "oracle_symbols.h"
#endif
#include <cstdio>
#include <cstdlib>

#include "f184_renderer.hpp"

using namespace f184;

struct Blob
{
    std::vector<char> d; size_t p = 0;
    explicit Blob(const char* path)
    {
        FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(2); }
        fseek(f, 0, SEEK_END); d.resize((size_t)ftell(f)); fseek(f, 0, SEEK_SET);
        if (fread(d.data(), 1, d.size(), f) != d.size()) exit(2);
        fclose(f);
    }
    template <class T> T get() { T v; memcpy(&v, &d[p], sizeof v); p += sizeof v; return v; }
    template <class T> std::vector<T> vec(size_t n) { std::vector<T> v(n); memcpy(v.data(), &d[p], n * sizeof(T)); p += n * sizeof(T); return v; }
};

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: frame_driver in.bin out.bin\n"); return 2; }
    try
    {
        Blob in(argv[1]);
        const uint32_t N = in.get<uint32_t>(), W = in.get<uint32_t>(), H = in.get<uint32_t>(), S = in.get<uint32_t>();
        f184_config cfg{};
        cfg.mode = F184_MODE_REFERENCE; cfg.grid_n = N; cfg.width = W; cfg.height = H; cfg.shadow_res = S;
        cfg.march_steps = 60; cfg.step_size = 0.2f; cfg.cone_max_distance = 32.0f; cfg.nranks = 1;
        CVoxelGI gi(cfg);

        // textures + materials: CBasicMaterial::Bind (Material/BasicMaterial.cpp:9-39)
        const uint32_t n_tex = in.get<uint32_t>();
        for (uint32_t t = 0; t < n_tex; t++)
        {
            const uint32_t tw = in.get<uint32_t>(), th = in.get<uint32_t>();
            auto px = in.vec<uint8_t>((size_t)tw * th * 4);
            gi.Check(f184_texture_upload(gi.Handle(), t, px.data(), tw, th), "f184_texture_upload");
        }
        const uint32_t n_mat = in.get<uint32_t>();
        for (uint32_t m = 0; m < n_mat; m++)
        {
            const int32_t tex = in.get<int32_t>();
            auto factor = in.vec<float>(4);
            gi.Check(f184_material_set(gi.Handle(), m, factor.data(), tex, 1), "f184_material_set");
        }
        // primitives + their model matrices (CSceneView::PrepareToRender, SceneView.cpp:57-64)
        const uint32_t n_prims = in.get<uint32_t>();
        std::vector<std::shared_ptr<CPrimitive>> prims;
        std::vector<Matrix3x4> modelMats;
        for (uint32_t p = 0; p < n_prims; p++)
        {
            auto P = std::make_shared<CPrimitive>();
            P->Material = in.get<uint16_t>(); in.get<uint16_t>();
            const uint32_t nv = in.get<uint32_t>(), ni = in.get<uint32_t>();
            P->Positions = in.vec<float>(3 * (size_t)nv); P->Normals = in.vec<float>(3 * (size_t)nv); P->TexCoords = in.vec<float>(2 * (size_t)nv);
            P->Indices = in.vec<uint32_t>(ni);
            Matrix3x4 M; auto m = in.vec<float>(12); memcpy(M.m, m.data(), 48);
            prims.push_back(P); modelMats.push_back(M);
        }
        // the frame's images
        auto GBufferDepth = MakeImageView(F184_SLOT_DEPTH), GBuffer0 = MakeImageView(F184_SLOT_ALBEDO), GBuffer1 = MakeImageView(F184_SLOT_NORMALS);
        auto GBuffer2 = MakeImageView(F184_SLOT_MATERIAL), ShadowDepth = MakeImageView(F184_SLOT_SHADOW);
        auto indirectTemporal = MakeImageView(F184_SLOT_INDIRECT_HISTORY), taaImageB = MakeImageView(F184_SLOT_TAA_HISTORY);
        const f184_view_constants voxelView = in.get<f184_view_constants>();
        const f184_light_list pointLights = in.get<f184_light_list>(), directionalLights = in.get<f184_light_list>();

        CVoxelizeRenderer VoxelizeRenderer(&gi);
        for (auto& P : prims) VoxelizeRenderer.PreparePrimitiveResources(P);
        std::vector<CPrimitive*> list;
        for (auto& P : prims) list.push_back(P.get());
        CScreenPass gtao_visibility(&gi, "gtao_visibility"), gtao_blur(&gi, "gtao_blur"), lighting_indirect(&gi, "lighting_indirect");
        CScreenPass indirect_blurX(&gi, "indirect_blurX"), indirect_blurY(&gi, "indirect_blurY"), lighting_deferred(&gi, "lighting_deferred"), gtao_color(&gi, "gtao_color");

        const uint32_t frames = in.get<uint32_t>();
        for (uint32_t frame = 0; frame < frames; frame++)
        {
            f184_trace_constants k = in.get<f184_trace_constants>();
            for (auto& v : {GBufferDepth, GBuffer1, ShadowDepth, GBuffer0, GBuffer2})
            {
                auto px = in.vec<char>(gi.ImageBytes(v));
                gi.Upload(v, px.data(), px.size());
            }
            const bool initial = frame == 0;
            if (!initial)
            {   // MegaPipeline.cpp:205-214
                gi.CopyImage(gtao_color.getRTViews()[1], taaImageB);
                gi.CopyImage(lighting_indirect.getRTViews()[0], indirectTemporal);
            }
            VoxelizeRenderer.RenderList(voxelView, modelMats, list);                       // :196, :218-223

            gtao_visibility.beginRender();                                                // :225-233
            gtao_visibility.setSampler("s");
            gtao_visibility.setImageView("t_albedo", GBuffer0);
            gtao_visibility.setImageView("t_normals", GBuffer1);
            gtao_visibility.setImageView("t_depth", GBufferDepth);
            gtao_visibility.setStruct("GlobalConstants", sizeof(f184_view_constants), &k.view);
            gtao_visibility.blit2d();
            gtao_visibility.endRender();

            gtao_blur.beginRender();                                                      // :235-239
            gtao_blur.setSampler("s");
            gtao_blur.setImageView("t_ao", gtao_visibility.getRTViews()[0]);
            gtao_blur.blit2d();
            gtao_blur.endRender();

            lighting_indirect.beginRender();                                              // :252-268
            lighting_indirect.setSampler("s");
            lighting_indirect.setImageView("t_depth", GBufferDepth);
            lighting_indirect.setImageView("t_shadow", ShadowDepth);
            lighting_indirect.setImageView("t_normals", GBuffer1);
            lighting_indirect.setImageView("temporal", indirectTemporal);
            lighting_indirect.setImageView("voxels", gi.getVoxelsImageView());
            lighting_indirect.setStruct("GlobalConstants", sizeof(f184_view_constants), &k.view);
            lighting_indirect.setStruct("ExtendedMatrices", sizeof(f184_extended_matrices), &k.ext);
            lighting_indirect.setStruct("Sun", sizeof(f184_sun), &directionalLights.lights[0]);
            lighting_indirect.setStruct("prevProj", sizeof(f184_prev_proj), &k.prev);
            lighting_indirect.setStruct("EngineCommonMiscs", sizeof(f184_engine_miscs), &k.miscs);
            lighting_indirect.setResetHistory(initial);
            lighting_indirect.blit2d();
            lighting_indirect.endRender();

            indirect_blurX.beginRender();                                                 // :270-276
            indirect_blurX.setSampler("s");
            indirect_blurX.setImageView("t_depth", GBufferDepth);
            indirect_blurX.setImageView("t_indirect", lighting_indirect.getRTViews()[0]);
            indirect_blurX.setStruct("EngineCommonMiscs", sizeof(f184_engine_miscs), &k.miscs);
            indirect_blurX.blit2d();
            indirect_blurX.endRender();

            indirect_blurY.beginRender();                                                 // :278-284
            indirect_blurY.setSampler("s");
            indirect_blurY.setImageView("t_depth", GBufferDepth);
            indirect_blurY.setImageView("t_indirect", indirect_blurX.getRTViews()[0]);
            indirect_blurY.setStruct("EngineCommonMiscs", sizeof(f184_engine_miscs), &k.miscs);
            indirect_blurY.blit2d();
            indirect_blurY.endRender();

            lighting_deferred.beginRender();                                              // :286-300
            lighting_deferred.setSampler("s");
            lighting_deferred.setImageView("t_albedo", GBuffer0);
            lighting_deferred.setImageView("t_normals", GBuffer1);
            lighting_deferred.setImageView("t_material", GBuffer2);
            lighting_deferred.setImageView("t_depth", GBufferDepth);
            lighting_deferred.setImageView("t_shadow", ShadowDepth);
            lighting_deferred.setStruct("GlobalConstants", sizeof(f184_view_constants), &k.view);
            lighting_deferred.setStruct("pointLights", sizeof(f184_light_list), &pointLights);
            lighting_deferred.setStruct("directionalLights", sizeof(f184_light_list), &directionalLights);
            lighting_deferred.setStruct("ExtendedMatrices", sizeof(f184_extended_matrices), &k.ext);
            lighting_deferred.blit2d();
            lighting_deferred.endRender();

            gtao_color.beginRender();                                                     // :302-319
            gtao_color.setSampler("s");
            gtao_color.setImageView("t_albedo", GBuffer0);
            gtao_color.setImageView("t_ao", gtao_blur.getRTViews()[0]);
            gtao_color.setImageView("t_depth", GBufferDepth);
            gtao_color.setImageView("t_lighting", lighting_deferred.getRTViews()[0]);
            gtao_color.setImageView("t_shadow", ShadowDepth);
            gtao_color.setImageView("t_indirect", indirect_blurY.getRTViews()[0]);
            gtao_color.setImageView("taaBuffer", taaImageB);
            gtao_color.setStruct("GlobalConstants", sizeof(f184_view_constants), &k.view);
            gtao_color.setStruct("ExtendedMatrices", sizeof(f184_extended_matrices), &k.ext);
            gtao_color.setStruct("prevProj", sizeof(f184_prev_proj), &k.prev);
            gtao_color.setStruct("Sun", sizeof(f184_sun), &directionalLights.lights[0]);
            gtao_color.setStruct("EngineCommonMiscs", sizeof(f184_engine_miscs), &k.miscs);
            gtao_color.setResetHistory(initial);
            gtao_color.blit2d();
            gtao_color.endRender();
        }

        // protocol behaviour the reference has and the tests pin: calls outside begin/endRender are ignored, unknown names throw
        gtao_color.setStruct("NoSuchBlock", 4, &N);
        bool threw = false;
        gtao_color.beginRender();
        try { gtao_color.setImageView("t_nonexistent", GBuffer0); } catch (const std::out_of_range&) { threw = true; }
        gtao_color.endRender();
        if (!threw) { fprintf(stderr, "unknown resource name did not throw\n"); return 3; }

        FILE* out = fopen(argv[2], "wb");
        for (auto v : {gi.getVoxelsImageView(), lighting_indirect.getRTViews()[0], gtao_blur.getRTViews()[0], indirect_blurY.getRTViews()[0],
                       lighting_deferred.getRTViews()[0], gtao_color.getRTViews()[0], gtao_color.getRTViews()[1]})
        {
            std::vector<char> px(gi.ImageBytes(v));
            gi.Readback(v, px.data(), px.size());
            fwrite(px.data(), 1, px.size(), out);
        }
        fclose(out);
    }
    catch (const std::exception& e)
    {
        fprintf(stderr, "frame_driver: %s\n", e.what());
        return 1;
    }
    return 0;
}
