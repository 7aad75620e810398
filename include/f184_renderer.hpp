// f184_renderer.hpp — the host side above the C-ABI, in the reference's own language and call style.
//
// The reference is C++ and has no plugin boundary: CMegaPipeline::Render() (Foreground/Renderer/MegaPipeline.cpp:148-342)
// drives the path through two small protocols,
//   * CVoxelizeRenderer  (Foreground/Renderer/VoxelizeRenderer.h:17-25):  PreparePrimitiveResources / RenderList /
//                        ClearResourceCache / GarbageCollectResourceCache / SetRenderPass
//   * CMaterial          (Foreground/Components/Material.h:74-99), the full-screen-pass protocol:
//                        beginRender -> setSampler / setImageView / setStruct(name, size, ptr) -> blit2d -> endRender, getRTViews
// This header mirrors both over libf184 (include/f184.h) — same method names, argument meaning and error behaviour — so the
// patched Render() keeps its shape and tests/cpp/frame_driver.cpp reads like the reference's frame:
//   * resource names are the GLSL names the reference binds by reflection ("t_depth", "GlobalConstants", ...); an unknown name
//     throws std::out_of_range exactly as CMaterial's `resources.at(id)` does (Material.cpp:172-196);
//   * set* calls outside beginRender/endRender are ignored, as in the reference (`if (ctx)`, ibid.);
//   * RenderList returns silently when there is nothing to draw (VoxelizeRenderer.cpp:49-50, 89-90);
//   * every libf184 failure becomes f184::CRuntimeError (the reference throws RHI::CRHIRuntimeError).
// Header-only C++17, no dependency beyond f184.h.  F184_FN lets the same code bind another library with the same surface
// (the CPU oracle exports f184o_*: tests build this header against both).
#pragma once
#include <array>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "f184.h"

#ifndef F184_FN
#define F184_FN(name) f184_##name
#endif

namespace f184 {

struct CRuntimeError : std::runtime_error { using std::runtime_error::runtime_error; };

// tc::Matrix3x4 (Math/Matrix3x4.h): three rows of four, row-major
struct Matrix3x4 { float m[12]; };

// What CPrimitive holds for this path (Foreground/SceneGraph/Primitive.h; filled by glTFSceneImporter.cpp:210-326)
struct CPrimitive
{
    std::vector<float> Positions, Normals, TexCoords;      // 3, 3, 2 floats per vertex
    std::vector<uint32_t> Indices;                         // triangle list
    uint16_t Material = 0;
};

// An image of the frame: what a CImageView::Ref is to the reference.  Identified by the slot it lives in.
struct CImageView
{
    uint32_t Slot = F184_SLOT_COUNT;
    using Ref = std::shared_ptr<CImageView>;
};
inline CImageView::Ref MakeImageView(uint32_t slot) { auto v = std::make_shared<CImageView>(); v->Slot = slot; return v; }

// The voxel/indirect section of CMegaPipeline as a context (MegaPipeline.cpp:27-60, 470-590)
class CVoxelGI
{
public:
    explicit CVoxelGI(const f184_config& config)
    {
        f184_config c = config;
        c.struct_size = sizeof(f184_config);
        if (F184_FN(create)(&c, &Ctx) != F184_OK) throw CRuntimeError(std::string("f184_create: ") + F184_FN(last_error)(nullptr));
    }
    ~CVoxelGI() { if (Ctx) F184_FN(destroy)(Ctx); }
    CVoxelGI(const CVoxelGI&) = delete;
    CVoxelGI& operator=(const CVoxelGI&) = delete;
    f184_ctx* Handle() const { return Ctx; }
    void Check(int rc, const char* what) const
    {
        if (rc != F184_OK) throw CRuntimeError(std::string(what) + ": " + F184_FN(last_error)(Ctx));
    }
    // test / tool convenience: host <-> image (in production the images are imported Vulkan memory, f184.h)
    void Upload(const CImageView::Ref& v, const void* host, size_t bytes) { Check(F184_FN(upload_image)(Ctx, v->Slot, host, bytes), "f184_upload_image"); }
    void Readback(const CImageView::Ref& v, void* host, size_t bytes) { Check(F184_FN(readback)(Ctx, v->Slot, host, bytes), "f184_readback"); }
    size_t ImageBytes(const CImageView::Ref& v)
    {
        f184_image_desc d{};
        Check(F184_FN(image_info)(Ctx, v->Slot, &d), "f184_image_info");
        return (size_t)d.size_bytes;
    }
    // copyCtx->CopyImage(src, dst) of the history images, MegaPipeline.cpp:205-214
    void CopyImage(const CImageView::Ref& src, const CImageView::Ref& dst)
    {
        if (src->Slot == F184_SLOT_INDIRECT_OUT && dst->Slot == F184_SLOT_INDIRECT_HISTORY) Check(F184_FN(copy_indirect_to_history)(Ctx), "f184_copy_indirect_to_history");
        else if (src->Slot == F184_SLOT_TAA_OUT && dst->Slot == F184_SLOT_TAA_HISTORY) Check(F184_FN(copy_taa_to_history)(Ctx), "f184_copy_taa_to_history");
        else throw CRuntimeError("CopyImage: only the two history copies of the section are supported");
    }
    // CMegaPipeline::getVoxelsImageView(), MegaPipeline.h:34
    CImageView::Ref getVoxelsImageView() const { return MakeImageView(F184_SLOT_VOXELS); }

private:
    f184_ctx* Ctx = nullptr;
};

// ---------------------------------------------------------------------------------------------------------------------
class CVoxelizeRenderer
{
public:
    explicit CVoxelizeRenderer(CVoxelGI* p) : Parent(p) {}

    void SetRenderPass(const void* /*renderPass*/, uint32_t /*subpass*/ = 0) {}       // the pass is libf184's own

    // The reference builds one pipeline + descriptor set per primitive here (VoxelizeRenderer.cpp:32-67); libf184 keeps all
    // prepared geometry resident as one flat scene, (re)uploaded by the next RenderList.
    void PreparePrimitiveResources(std::shared_ptr<CPrimitive> primitive)
    {
        if (!primitive) return;
        for (auto& w : Cached)
            if (w.lock() == primitive) return;
        Cached.push_back(primitive);
        Dirty = true;
    }
    void ClearResourceCache() { Cached.clear(); Dirty = true; }
    void GarbageCollectResourceCache()
    {
        const size_t n = Cached.size();
        for (size_t i = Cached.size(); i-- > 0;)
            if (Cached[i].expired()) Cached.erase(Cached.begin() + (long)i);
        if (Cached.size() != n) Dirty = true;
    }

    // ClearImage(VoxelImage) + the voxelization pass (MegaPipeline.cpp:196, 218-223).  `context` of the reference is the voxelizer
    // camera's view constants here: the one thing the render context contributes to the shaders (BindEngineCommonForView).
    void RenderList(const f184_view_constants& context, const std::vector<Matrix3x4>& modelMats, const std::vector<CPrimitive*>& primitives)
    {
        if (primitives.empty() || modelMats.size() != primitives.size()) return;          // silent, as the reference
        if (Dirty || primitives != Drawn || !SameMats(modelMats)) UploadScene(modelMats, primitives);
        Parent->Check(F184_FN(voxelize)(Parent->Handle(), &context), "f184_voxelize");
    }

private:
    bool SameMats(const std::vector<Matrix3x4>& m) const
    {
        return m.size() == Mats.size() && (m.empty() || std::memcmp(m.data(), Mats.data(), m.size() * sizeof(Matrix3x4)) == 0);
    }
    void UploadScene(const std::vector<Matrix3x4>& modelMats, const std::vector<CPrimitive*>& prims)
    {
        std::vector<float> pos, nrm, uv, mats;
        std::vector<uint32_t> idx;
        std::vector<uint16_t> tmat, tmodel;
        for (size_t p = 0; p < prims.size(); p++)
        {
            const CPrimitive& P = *prims[p];
            const uint32_t base = (uint32_t)(pos.size() / 3);
            pos.insert(pos.end(), P.Positions.begin(), P.Positions.end());
            nrm.insert(nrm.end(), P.Normals.begin(), P.Normals.end());
            uv.insert(uv.end(), P.TexCoords.begin(), P.TexCoords.end());
            for (uint32_t i : P.Indices) idx.push_back(base + i);
            tmat.insert(tmat.end(), P.Indices.size() / 3, P.Material);
            tmodel.insert(tmodel.end(), P.Indices.size() / 3, (uint16_t)p);
            // tc::Matrix3x4 -> mat4 in upload order (ToMatrix4().Transpose(), VoxelizeRenderer.cpp:100-106): column-major
            const float* r = modelMats[p].m;
            const float cm[16] = {r[0], r[4], r[8], 0.f, r[1], r[5], r[9], 0.f, r[2], r[6], r[10], 0.f, r[3], r[7], r[11], 1.f};
            mats.insert(mats.end(), cm, cm + 16);
        }
        f184_scene_desc d{};
        d.positions = pos.data(); d.normals = nrm.data(); d.uvs = uv.data(); d.indices = idx.data();
        d.tri_material = tmat.data(); d.tri_model = tmodel.data(); d.model_mats = mats.data();
        d.n_verts = (uint32_t)(pos.size() / 3); d.n_tris = (uint32_t)(idx.size() / 3); d.n_models = (uint32_t)prims.size();
        Parent->Check(F184_FN(scene_upload)(Parent->Handle(), &d), "f184_scene_upload");
        Drawn = prims; Mats = modelMats; Dirty = false;
    }

    CVoxelGI* Parent;
    std::vector<std::weak_ptr<CPrimitive>> Cached;
    std::vector<CPrimitive*> Drawn;
    std::vector<Matrix3x4> Mats;
    bool Dirty = true;
};

// ---------------------------------------------------------------------------------------------------------------------
// One full-screen pass of the section, named by its fragment shader as the reference names them (MegaPipeline.cpp:543-590):
//   "gtao_visibility", "gtao_blur", "lighting_indirect", "indirect_blurX", "indirect_blurY", "lighting_deferred", "gtao_color"
class CScreenPass
{
public:
    CScreenPass(CVoxelGI* p, const std::string& fragmentShader) : Parent(p), Name(fragmentShader)
    {
        struct Def { const char* name; std::vector<std::pair<const char*, uint32_t>> images; std::vector<const char*> structs; std::vector<uint32_t> targets; };
        static const Def defs[] = {
            {"gtao_visibility", {{"t_albedo", F184_SLOT_ALBEDO}, {"t_normals", F184_SLOT_NORMALS}, {"t_depth", F184_SLOT_DEPTH}}, {"GlobalConstants"}, {F184_SLOT_AO_RAW}},
            {"gtao_blur", {{"t_ao", F184_SLOT_AO_RAW}}, {}, {F184_SLOT_AO_OUT}},
            {"lighting_indirect", {{"t_depth", F184_SLOT_DEPTH}, {"t_shadow", F184_SLOT_SHADOW}, {"t_normals", F184_SLOT_NORMALS}, {"temporal", F184_SLOT_INDIRECT_HISTORY},
                                   {"voxels", F184_SLOT_VOXELS}}, {"GlobalConstants", "ExtendedMatrices", "Sun", "prevProj", "EngineCommonMiscs"}, {F184_SLOT_INDIRECT_OUT}},
            {"indirect_blurX", {{"t_depth", F184_SLOT_DEPTH}, {"t_indirect", F184_SLOT_INDIRECT_OUT}}, {"EngineCommonMiscs"}, {F184_SLOT_INDIRECT_BLUR_X}},
            {"indirect_blurY", {{"t_depth", F184_SLOT_DEPTH}, {"t_indirect", F184_SLOT_INDIRECT_BLUR_X}}, {"EngineCommonMiscs"}, {F184_SLOT_INDIRECT_FINAL}},
            {"lighting_deferred", {{"t_albedo", F184_SLOT_ALBEDO}, {"t_normals", F184_SLOT_NORMALS}, {"t_material", F184_SLOT_MATERIAL}, {"t_depth", F184_SLOT_DEPTH},
                                   {"t_shadow", F184_SLOT_SHADOW}}, {"GlobalConstants", "pointLights", "directionalLights", "ExtendedMatrices"}, {F184_SLOT_LIGHTING}},
            {"gtao_color", {{"t_albedo", F184_SLOT_ALBEDO}, {"t_ao", F184_SLOT_AO_OUT}, {"t_depth", F184_SLOT_DEPTH}, {"t_lighting", F184_SLOT_LIGHTING}, {"t_shadow", F184_SLOT_SHADOW},
                            {"t_indirect", F184_SLOT_INDIRECT_FINAL}, {"taaBuffer", F184_SLOT_TAA_HISTORY}},
             {"GlobalConstants", "ExtendedMatrices", "prevProj", "Sun", "EngineCommonMiscs"}, {F184_SLOT_COLOR_OUT, F184_SLOT_TAA_OUT}},
        };
        for (const Def& d : defs)
            if (Name == d.name)
            {
                for (auto& i : d.images) Resources[i.first] = i.second;
                for (auto* s : d.structs) Resources[s] = F184_SLOT_COUNT;
                Resources["s"] = F184_SLOT_COUNT;
                for (uint32_t t : d.targets) Targets.push_back(MakeImageView(t));
                return;
            }
        throw CRuntimeError("no such pass: " + fragmentShader);            // the reference fails to load the SPIR-V file
    }

    void createPipeline(int /*w*/, int /*h*/) {}                           // sizes belong to the context (f184_config)

    void beginRender(const void* /*cmdList*/ = nullptr) { Recording = true; }
    void setSampler(const std::string& id, const void* /*sampler*/ = nullptr) { if (Recording) Resources.at(id); }
    void setImageView(const std::string& id, const CImageView::Ref& obj)
    {
        if (!Recording) return;
        const uint32_t want = Resources.at(id);                            // std::out_of_range on an unknown name, as the reference
        if (!obj || obj->Slot != want) throw CRuntimeError(Name + ": " + id + " is wired to a fixed image of the section");
    }
    void setStruct(const std::string& id, size_t size, const void* obj)
    {
        if (!Recording) return;
        Resources.at(id);
        auto put = [&](void* dst, size_t n) { if (size != n) throw CRuntimeError(Name + ": " + id + " has the wrong size"); std::memcpy(dst, obj, n); };
        if (id == "GlobalConstants") put(&K.view, sizeof K.view);
        else if (id == "ExtendedMatrices") put(&K.ext, sizeof K.ext);
        else if (id == "prevProj") put(&K.prev, sizeof K.prev);
        else if (id == "Sun") put(&K.sun, sizeof K.sun);
        else if (id == "EngineCommonMiscs") put(&K.miscs, sizeof K.miscs);
        else if (id == "pointLights") put(&Point, sizeof Point);
        else if (id == "directionalLights") put(&Directional, sizeof Directional);
    }
    // first frame / after a resize: the history images are cleared instead of copied (MegaPipeline.cpp:197-204)
    void setResetHistory(bool reset) { K.reset_history = reset ? 1u : 0u; }

    void blit2d()
    {
        if (!Recording) return;
        f184_ctx* c = Parent->Handle();
        if (Name == "gtao_visibility") Parent->Check(F184_FN(gtao)(c, &K.view), "f184_gtao");             // visibility + its blur are one entry point
        else if (Name == "gtao_blur") {}
        else if (Name == "lighting_indirect") Parent->Check(F184_FN(trace_indirect)(c, &K), "f184_trace_indirect");
        else if (Name == "indirect_blurX") Parent->Check(F184_FN(blur_indirect)(c, &K.miscs), "f184_blur_indirect");   // X and Y are one entry point
        else if (Name == "indirect_blurY") {}
        else if (Name == "lighting_deferred") Parent->Check(F184_FN(lighting_deferred)(c, &K.view, &K.ext, &Point, &Directional), "f184_lighting_deferred");
        else if (Name == "gtao_color") Parent->Check(F184_FN(composite)(c, &K), "f184_composite");
    }
    void endRender() { Recording = false; }
    const std::vector<CImageView::Ref>& getRTViews() const { return Targets; }

private:
    CVoxelGI* Parent;
    std::string Name;
    std::unordered_map<std::string, uint32_t> Resources;
    std::vector<CImageView::Ref> Targets;
    f184_trace_constants K{};
    f184_light_list Point{}, Directional{};
    bool Recording = false;
};

}  // namespace f184
