// lumen_b200_adapter.hpp — the reference-side binding of liblumen_b200.so: a header-only C++17 implementation of the reference's
// abstract renderer interface over the C ABI of lumen_b200.h.
//
// It is compiled INSIDE the reference tree (it includes the reference's own interface headers, nothing of theirs is restated here):
//     class B200::Renderer  : LumenRenderer            Lumen/src/Lumen/Renderer/LumenRenderer.h:37-218
//     class B200::Scene     : Lumen::ILumenScene       Lumen/src/Lumen/ModelLoading/ILumenScene.h:11-71
//     class B200::Material  : Lumen::ILumenMaterial    Lumen/src/Lumen/Renderer/ILumenResources.h:18-87
//     B200::Texture / Primitive / Mesh / Volume        ILumenResources.h:11-16,89-136
//     B200::MeshInstance / VolumeInstance              ModelLoading/MeshInstance.h:22-112, VolumeInstance.h:9-28
// so that the application switches renderers by changing one line (Sandbox/src/Application.cpp:83):
//     std::make_shared<WaveFront::WaveFrontRenderer>()   ->   std::make_shared<B200::Renderer>(settings)
// Include paths needed besides the reference's own: <repo>/include. Link: -llumen_b200.
//
// Behaviour mirrored from WaveFront::WaveFrontRenderer (LumenPT/src/Framework/WaveFrontRenderer.cpp):
//   * resources are uploaded synchronously by the Create* calls (:1148-1330);
//   * the scene graph stays in the reference's own classes (Transform, MeshInstance, Camera); before every frame the adapter pushes
//     what changed — world matrices (glm::transpose, as PTMeshInstance.cpp:143-147), emissiveness, override materials, camera
//     matrix — exactly the data TraceFrame reads at :576-577 and PTScene.cpp:89-154;
//   * StartRendering() runs frames on the renderer's own thread (:1109-1117); TraceFrame() renders one frame on the caller's.
// Out of scope of the path: GL/D3D interop (GetOutputTexture returns 0), DLSS (InitNGX is a no-op), frame snapshots.
#pragma once

#include "Lumen/Renderer/LumenRenderer.h"
#include "Lumen/Renderer/ILumenResources.h"
#include "Lumen/Renderer/Camera.h"
#include "Lumen/ModelLoading/ILumenScene.h"
#include "Lumen/ModelLoading/MeshInstance.h"
#include "Lumen/ModelLoading/VolumeInstance.h"
#include "Framework/CudaGLTexture.h"      // LumenPT/src: FrameSnapshot.h needs the complete type for its inline constructor
#include "Tools/FrameSnapshot.h"          // LumenPT/src/Tools/FrameSnapshot.h: EndSnapshot() returns a unique_ptr to it

#include <lumen_b200.h>

#include <glm/glm.hpp>
#include <glm/gtc/type_ptr.hpp>

#include <atomic>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

namespace B200 {

// the reference aborts on CUDA / OptiX errors (CudaUtilities.h:24-57); here every failed call throws with the library's message
inline void Check(int a_ReturnCode)
{
    if (a_ReturnCode != LB_OK) throw std::runtime_error(std::string("lumen_b200: ") + lb_last_error());
}

// glm matrices are column-major; the C ABI takes row-major 4x4 (what PTMeshInstance.cpp:143-147 uploads after glm::transpose)
inline void RowMajor(const glm::mat4& a_Matrix, float a_Out[16])
{
    const glm::mat4 t = glm::transpose(a_Matrix);
    std::memcpy(a_Out, glm::value_ptr(t), 64);
}

struct Texture final : Lumen::ILumenTexture
{
    explicit Texture(LbHandle a_Handle) : m_Handle(a_Handle) {}
    LbHandle m_Handle;
};

inline LbHandle HandleOf(const std::shared_ptr<Lumen::ILumenTexture>& a_Texture)
{
    return a_Texture ? static_cast<const Texture&>(*a_Texture).m_Handle : LB_NO_HANDLE;
}

// PTMaterial (LumenPT/src/Framework/PTMaterial.cpp): every setter re-uploads the device material
class Material final : public Lumen::ILumenMaterial
{
public:
    Material(LbRenderer a_Renderer, const LumenRenderer::MaterialData& a_Data) : m_Renderer(a_Renderer)
    {
        std::memset(&m_Desc, 0, sizeof m_Desc);
        std::memcpy(m_Desc.diffuse_color, &a_Data.m_DiffuseColor, 16);
        std::memcpy(m_Desc.emission, &a_Data.m_EmissionVal, 12);
        m_Desc.transmission_factor = a_Data.m_TransmissionFactor;
        m_Desc.clear_coat_factor = a_Data.m_ClearCoatFactor;
        m_Desc.clear_coat_roughness_factor = a_Data.m_ClearCoatRoughnessFactor;
        m_Desc.index_of_refraction = a_Data.m_IndexOfRefraction;
        m_Desc.specular_factor = a_Data.m_SpecularFactor;
        m_Desc.specular_tint_factor = a_Data.m_SpecularTintFactor;
        m_Desc.subsurface_factor = a_Data.m_SubSurfaceFactor;
        m_Desc.luminance = a_Data.m_Luminance;
        m_Desc.anisotropic = a_Data.m_Anisotropic;
        m_Desc.sheen_factor = a_Data.m_SheenFactor;
        m_Desc.sheen_tint_factor = a_Data.m_SheenTintFactor;
        m_Desc.metallic_factor = a_Data.m_MetallicFactor;
        m_Desc.roughness_factor = a_Data.m_RoughnessFactor;
        std::memcpy(m_Desc.tint_factor, &a_Data.m_TintFactor, 12);
        std::memcpy(m_Desc.transmittance, &a_Data.m_Transmittance, 12);
        m_Textures[0] = a_Data.m_DiffuseTexture;       m_Textures[1] = a_Data.m_NormalMap;
        m_Textures[2] = a_Data.m_MetallicRoughnessTexture; m_Textures[3] = a_Data.m_EmissiveTexture;
        m_Textures[4] = a_Data.m_TransmissionTexture;  m_Textures[5] = a_Data.m_ClearCoatTexture;
        m_Textures[6] = a_Data.m_ClearCoatRoughnessTexture; m_Textures[7] = a_Data.m_TintTexture;
        SyncTextureHandles();
        Check(lb_material_create(m_Renderer, &m_Desc, &m_Handle));
    }

    LbHandle GetHandle() const { return m_Handle; }

    void SetDiffuseColor(const glm::vec4& a_Color) override { std::memcpy(m_Desc.diffuse_color, &a_Color, 16); Upload(); }
    void SetDiffuseTexture(std::shared_ptr<Lumen::ILumenTexture> a_Texture) override { SetTexture(0, a_Texture); }
    void SetEmission(const glm::vec3& a_Emission = glm::vec3(0.0f, 0.0f, 0.0f)) override { std::memcpy(m_Desc.emission, &a_Emission, 12); Upload(); }
    void SetEmissiveTexture(std::shared_ptr<Lumen::ILumenTexture> a_Texture) override { SetTexture(3, a_Texture); }
    void SetMetalRoughnessTexture(std::shared_ptr<Lumen::ILumenTexture> a_Texture) override { SetTexture(2, a_Texture); }
    void SetNormalTexture(std::shared_ptr<Lumen::ILumenTexture> a_Texture) override { SetTexture(1, a_Texture); }
    void SetClearCoatTexture(std::shared_ptr<Lumen::ILumenTexture> a_Texture) override { SetTexture(5, a_Texture); }
    void SetClearCoatRoughnessTexture(std::shared_ptr<Lumen::ILumenTexture> a_Texture) override { SetTexture(6, a_Texture); }
    void SetClearCoatFactor(float a_Factor) override { m_Desc.clear_coat_factor = a_Factor; Upload(); }
    void SetClearCoatRoughnessFactor(float a_Factor) override { m_Desc.clear_coat_roughness_factor = a_Factor; Upload(); }
    void SetLuminance(float a_Factor) override { m_Desc.luminance = a_Factor; Upload(); }
    void SetSheenFactor(float a_Factor) override { m_Desc.sheen_factor = a_Factor; Upload(); }
    void SetSheenTintFactor(float a_Factor) override { m_Desc.sheen_tint_factor = a_Factor; Upload(); }
    void SetAnisotropic(float a_Factor) override { m_Desc.anisotropic = a_Factor; Upload(); }
    void SetTintTexture(std::shared_ptr<Lumen::ILumenTexture> a_Texture) override { SetTexture(7, a_Texture); }
    void SetTintFactor(const glm::vec3& a_Factor) override { std::memcpy(m_Desc.tint_factor, &a_Factor, 12); Upload(); }
    void SetTransmissionTexture(std::shared_ptr<Lumen::ILumenTexture> a_Texture) override { SetTexture(4, a_Texture); }
    void SetTransmissionFactor(float a_Factor) override { m_Desc.transmission_factor = a_Factor; Upload(); }
    void SetTransmittanceFactor(const glm::vec3& a_Factor) override { std::memcpy(m_Desc.transmittance, &a_Factor, 12); Upload(); }
    void SetIndexOfRefraction(float a_Factor) override { m_Desc.index_of_refraction = a_Factor; Upload(); }
    void SetSpecularFactor(float a_Factor) override { m_Desc.specular_factor = a_Factor; Upload(); }
    void SetSpecularTintFactor(float a_Factor) override { m_Desc.specular_tint_factor = a_Factor; Upload(); }
    void SetSubSurfaceFactor(float a_Factor) override { m_Desc.subsurface_factor = a_Factor; Upload(); }
    void SetMetallicFactor(float a_Factor) override { m_Desc.metallic_factor = a_Factor; Upload(); }
    void SetRoughnessFactor(float a_Factor) override { m_Desc.roughness_factor = a_Factor; Upload(); }

    float GetClearCoatFactor() override { return m_Desc.clear_coat_factor; }
    float GetClearCoatRoughnessFactor() override { return m_Desc.clear_coat_roughness_factor; }
    float GetLuminance() override { return m_Desc.luminance; }
    float GetSheenFactor() override { return m_Desc.sheen_factor; }
    float GetSheenTintFactor() override { return m_Desc.sheen_tint_factor; }
    float GetAnisotropic() override { return m_Desc.anisotropic; }
    glm::vec3 GetTintFactor() override { return glm::make_vec3(m_Desc.tint_factor); }
    float GetTransmissionFactor() override { return m_Desc.transmission_factor; }
    glm::vec3 GetTransmittanceFactor() override { return glm::make_vec3(m_Desc.transmittance); }
    float GetIndexOfRefraction() override { return m_Desc.index_of_refraction; }
    float GetSpecularFactor() override { return m_Desc.specular_factor; }
    float GetSpecularTintFactor() override { return m_Desc.specular_tint_factor; }
    float GetSubSurfaceFactor() override { return m_Desc.subsurface_factor; }
    float GetMetallicFactor() override { return m_Desc.metallic_factor; }
    float GetRoughnessFactor() override { return m_Desc.roughness_factor; }
    glm::vec4 GetDiffuseColor() const override { return glm::make_vec4(m_Desc.diffuse_color); }
    glm::vec3 GetEmissiveColor() const override { return glm::make_vec3(m_Desc.emission); }
    Lumen::ILumenTexture& GetDiffuseTexture() const override { return TextureRef(0); }
    Lumen::ILumenTexture& GetEmissiveTexture() const override { return TextureRef(3); }

private:
    void SyncTextureHandles()
    {
        m_Desc.diffuse_texture = HandleOf(m_Textures[0]);            m_Desc.normal_texture = HandleOf(m_Textures[1]);
        m_Desc.metallic_roughness_texture = HandleOf(m_Textures[2]); m_Desc.emissive_texture = HandleOf(m_Textures[3]);
        m_Desc.transmission_texture = HandleOf(m_Textures[4]);       m_Desc.clear_coat_texture = HandleOf(m_Textures[5]);
        m_Desc.clear_coat_roughness_texture = HandleOf(m_Textures[6]); m_Desc.tint_texture = HandleOf(m_Textures[7]);
    }
    void SetTexture(int a_Slot, std::shared_ptr<Lumen::ILumenTexture>& a_Texture) { m_Textures[a_Slot] = a_Texture; SyncTextureHandles(); Upload(); }
    void Upload() { Check(lb_material_update(m_Renderer, m_Handle, &m_Desc)); }
    Lumen::ILumenTexture& TextureRef(int a_Slot) const
    {
        if (!m_Textures[a_Slot]) throw std::runtime_error("lumen_b200: the material uses the renderer's default texture in this slot");
        return *m_Textures[a_Slot];
    }

    LbRenderer m_Renderer;
    LbHandle m_Handle = LB_NO_HANDLE;
    LbMaterialDesc m_Desc;
    std::shared_ptr<Lumen::ILumenTexture> m_Textures[8];    // diffuse, normal, metal-roughness, emissive, transmission, clear coat, clear-coat roughness, tint
};

struct Primitive final : Lumen::ILumenPrimitive { LbHandle m_Handle = LB_NO_HANDLE; };

struct Mesh final : Lumen::ILumenMesh
{
    Mesh(std::vector<std::shared_ptr<Lumen::ILumenPrimitive>>& a_Primitives, LbHandle a_Handle) : ILumenMesh(a_Primitives), m_Handle(a_Handle) {}
    LbHandle m_Handle;
};

struct Volume final : Lumen::ILumenVolume
{
    explicit Volume(LbHandle a_Handle) : m_Handle(a_Handle) {}
    LbHandle m_Handle;
};

// PTMeshInstance (LumenPT/src/Framework/PTMeshInstance.cpp): the instance exists in the renderer once it has a mesh; afterwards
// the adapter compares what the renderer last saw with the instance's current state before every frame.
class MeshInstance final : public Lumen::MeshInstance
{
public:
    void Synchronise(LbRenderer a_Renderer)
    {
        if (!m_MeshRef) return;
        float world[16]; RowMajor(m_Transform.GetWorldTransformationMatrix(), world);
        LbEmissiveness em;
        em.mode = static_cast<int32_t>(m_EmissiveProperties.m_EmissionMode);       // ENABLED / DISABLED / OVERRIDE = 0 / 1 / 2 in both enums
        std::memcpy(em.override_radiance, &m_EmissiveProperties.m_OverrideRadiance, 12);
        em.scale = m_EmissiveProperties.m_Scale;
        const LbHandle overrideMaterial = m_OverrideMaterial ? static_cast<const Material&>(*m_OverrideMaterial).GetHandle() : LB_NO_HANDLE;
        const LbHandle mesh = static_cast<const Mesh&>(*m_MeshRef).m_Handle;
        if (m_Handle == LB_NO_HANDLE || mesh != m_LastMesh)
        {
            // a changed mesh becomes a new renderer instance; the old one keeps rendering nothing (zero scale), as the reference
            // has no per-instance removal either (ILumenScene only offers Clear())
            if (m_Handle != LB_NO_HANDLE) { const float zero[16] = {0}; Check(lb_instance_set_transform(a_Renderer, m_Handle, zero)); }
            Check(lb_scene_add_mesh_instance(a_Renderer, mesh, world, &em, overrideMaterial, &m_Handle));
            m_LastMesh = mesh;
        }
        else
        {
            if (std::memcmp(world, m_LastWorld, 64) != 0) Check(lb_instance_set_transform(a_Renderer, m_Handle, world));
            if (std::memcmp(&em, &m_LastEmissiveness, sizeof em) != 0) Check(lb_instance_set_emissiveness(a_Renderer, m_Handle, &em));
            if (overrideMaterial != m_LastOverride) Check(lb_instance_set_override_material(a_Renderer, m_Handle, overrideMaterial));
        }
        std::memcpy(m_LastWorld, world, 64); m_LastEmissiveness = em; m_LastOverride = overrideMaterial;
    }
    void Forget() { m_Handle = LB_NO_HANDLE; }

private:
    LbHandle m_Handle = LB_NO_HANDLE, m_LastMesh = LB_NO_HANDLE, m_LastOverride = LB_NO_HANDLE;
    float m_LastWorld[16] = {0};
    LbEmissiveness m_LastEmissiveness{};
};

class VolumeInstance final : public Lumen::VolumeInstance
{
public:
    void Synchronise(LbRenderer a_Renderer)
    {
        if (!m_VolumeRef || m_Handle != LB_NO_HANDLE) return;
        float world[16]; RowMajor(m_Transform.GetWorldTransformationMatrix(), world);
        Check(lb_scene_add_volume_instance(a_Renderer, static_cast<const Volume&>(*m_VolumeRef).m_Handle, world, m_Density, &m_Handle));
    }
    void Forget() { m_Handle = LB_NO_HANDLE; }

private:
    LbHandle m_Handle = LB_NO_HANDLE;
};

// PTScene (LumenPT/src/Framework/PTScene.cpp:24-66): hands out instances of the renderer's own type
class Scene final : public Lumen::ILumenScene
{
public:
    explicit Scene(LbRenderer a_Renderer) : m_Renderer(a_Renderer) {}

    Lumen::MeshInstance* AddMesh() override
    {
        m_MeshInstances.push_back(std::make_unique<B200::MeshInstance>());
        return m_MeshInstances.back().get();
    }
    Lumen::VolumeInstance* AddVolume() override
    {
        m_VolumeInstances.push_back(std::make_unique<B200::VolumeInstance>());
        return m_VolumeInstances.back().get();
    }
    void Clear() override
    {
        Check(lb_scene_clear(m_Renderer));
        Lumen::ILumenScene::Clear();
    }
    // everything TraceFrame reads from the scene (WaveFrontRenderer.cpp:576-577, PTScene.cpp:89-154)
    void Synchronise()
    {
        for (auto& instance : m_MeshInstances) static_cast<B200::MeshInstance&>(*instance).Synchronise(m_Renderer);
        for (auto& instance : m_VolumeInstances) static_cast<B200::VolumeInstance&>(*instance).Synchronise(m_Renderer);
        glm::mat4 previous, current;
        m_Camera->GetMatrixData(previous, current);
        float world[16]; RowMajor(current, world);
        Check(lb_camera_set_matrix(m_Renderer, world));
        const glm::vec2& range = m_Camera->GetMinMaxRenderDistance();
        Check(lb_camera_set_min_max_distance(m_Renderer, range.x, range.y));
    }

private:
    LbRenderer m_Renderer;
};

// mirrors WaveFront::WaveFrontSettings (LumenPT/src/Framework/WaveFrontRenderer.h:31-48) for the fields the path uses
struct Settings
{
    glm::uvec2 renderResolution = {1280, 720};
    glm::uvec2 outputResolution = {1280, 720};
    uint32_t depth = 5;
    bool blendOutput = false;
    bool restir = true, restirTemporal = true, restirSpatial = true;
    int device = 0;
    uint32_t volumeMode = LB_VOLUME_COMPAT;
};

class Renderer final : public LumenRenderer
{
public:
    Renderer() = default;
    explicit Renderer(const Settings& a_Settings) { Init(a_Settings); }
    // a member of a RendererGroup: the handle belongs to the group (lb_group_member), this object only speaks LumenRenderer for it
    Renderer(LbRenderer a_Borrowed, glm::uvec2 a_OutputResolution) : m_Renderer(a_Borrowed), m_Owned(false), m_OutputResolution(a_OutputResolution)
    {
        CreateDefaultResources();
        m_Scene = CreateScene(SceneData());
    }
    ~Renderer() override
    {
        StopRendering();
        m_Scene.reset();
        if (m_Renderer && m_Owned) lb_destroy(m_Renderer);
    }

    // WaveFrontRenderer::Init, WaveFrontRenderer.cpp:70-322 (also creates the default textures and the first scene, :318-321)
    void Init(const Settings& a_Settings)
    {
        LbSettings s;
        std::memset(&s, 0, sizeof s);
        s.width = a_Settings.renderResolution.x; s.height = a_Settings.renderResolution.y; s.depth = a_Settings.depth;
        s.blend_output = a_Settings.blendOutput; s.restir = a_Settings.restir; s.restir_temporal = a_Settings.restirTemporal;
        s.restir_spatial = a_Settings.restirSpatial; s.device = a_Settings.device; s.volume_mode = a_Settings.volumeMode;
        Check(lb_create(&s, &m_Renderer));
        m_OutputResolution = a_Settings.outputResolution;
        CreateDefaultResources();
        m_Scene = CreateScene(SceneData());
    }

    void StartRendering() override
    {
        if (m_Thread.joinable()) return;
        m_Stop = false;
        m_Thread = std::thread([this]() { while (!m_Stop.load()) TraceFrame(); });
    }
    void StopRendering()
    {
        m_Stop = true;
        if (m_Thread.joinable()) m_Thread.join();
    }

    // one frame: WaveFrontRenderer::TraceFrame, WaveFrontRenderer.cpp:435-1089
    void TraceFrame()
    {
        std::lock_guard<std::mutex> lock(m_FrameMutex);
        if (m_Scene) static_cast<B200::Scene&>(*m_Scene).Synchronise();
        Check(lb_render_frames(m_Renderer, 1));
        if (m_Scene) m_Scene->m_Camera->UpdatePreviousFrameMatrix();          // WaveFrontRenderer.cpp:1040
        const char* names = nullptr; float micros[64]; uint32_t count = 0;          // names: one ';'-separated list, stages in launch order
        Check(lb_frame_stats(m_Renderer, &names, micros, 64, &count));
        std::lock_guard<std::mutex> statsLock(m_FrameStatsMutex);
        m_LastFrameStats.m_Id = ++m_FrameId;
        m_LastFrameStats.m_Times.clear();
        const char* at = names ? names : "";
        for (uint32_t k = 0; k < count && k < 64; ++k)
        {
            const char* end = std::strchr(at, ';');
            const std::string stage = end ? std::string(at, end) : std::string(at);
            m_LastFrameStats.m_Times[stage] += static_cast<uint64_t>(micros[k]);         // a stage that runs once per wave adds up
            at = end ? end + 1 : at + stage.size();
        }
    }

    std::unique_ptr<Lumen::ILumenPrimitive> CreatePrimitive(PrimitiveData& a_Data) override
    {
        LbPrimitiveDesc d;
        std::memset(&d, 0, sizeof d);
        if (a_Data.m_Interleaved)
        {
            // `Vertex` (LumenPT/src/Shaders/CppCommon/ModelStructs.h:21-28): float3 position @0, float2 uv @16, float3 normal @24,
            // float4 tangent @48, 64 bytes (CUDA vector alignment)
            const uint8_t* v = a_Data.m_VertexBinary.data();
            d.positions = v; d.uvs = v + 16; d.normals = v + 24; d.tangents = v + 48;
            d.position_stride = d.uv_stride = d.normal_stride = d.tangent_stride = 64;
            d.vertex_count = static_cast<uint32_t>(a_Data.m_VertexBinary.size() / 64);
        }
        else
        {
            // InterleaveVertexData, WaveFrontRenderer.cpp:1096-1146: missing streams are zero-filled there; NULL means the same here
            d.vertex_count = static_cast<uint32_t>(a_Data.m_Positions.Size());
            d.positions = &a_Data.m_Positions[0]; d.position_stride = 12;
            if (!a_Data.m_TexCoords.Empty()) { d.uvs = &a_Data.m_TexCoords[0]; d.uv_stride = 8; }
            if (!a_Data.m_Normals.Empty()) { d.normals = &a_Data.m_Normals[0]; d.normal_stride = 12; }
            if (!a_Data.m_Tangents.Empty()) { d.tangents = &a_Data.m_Tangents[0]; d.tangent_stride = 16; }
        }
        d.indices = a_Data.m_IndexBinary.data(); d.index_size = static_cast<uint32_t>(a_Data.m_IndexSize);
        d.index_count = static_cast<uint32_t>(a_Data.m_IndexBinary.size() / a_Data.m_IndexSize);
        d.material = static_cast<const Material&>(*a_Data.m_Material).GetHandle();
        auto prim = std::make_unique<Primitive>();
        Check(lb_primitive_create(m_Renderer, &d, &prim->m_Handle));
        prim->m_Material = a_Data.m_Material;
        // the light list itself is built by the library at scene commit (GPUDataBufferKernels.cu:66-186); the two bookkeeping
        // members the interface exposes are filled from the material like WaveFrontRenderer.cpp:1188-1204
        prim->m_ContainEmissive = a_Data.m_Material->GetEmissiveColor() != glm::vec3(0.f);
        prim->m_NumLights = prim->m_ContainEmissive ? d.index_count / 3 : 0;
        return prim;
    }

    std::shared_ptr<Lumen::ILumenMesh> CreateMesh(std::vector<std::shared_ptr<Lumen::ILumenPrimitive>>& a_Primitives) override
    {
        std::vector<LbHandle> handles;
        for (auto& p : a_Primitives) handles.push_back(static_cast<const Primitive&>(*p).m_Handle);
        LbHandle h = LB_NO_HANDLE;
        Check(lb_mesh_create(m_Renderer, handles.data(), static_cast<uint32_t>(handles.size()), &h));
        return std::make_shared<Mesh>(a_Primitives, h);
    }

    std::shared_ptr<Lumen::ILumenTexture> CreateTexture(void* a_PixelData, uint32_t a_Width, uint32_t a_Height, bool a_Normalize) override
    {
        // a_Normalize = "sRGB texel, decode to linear on read" (PTTexture.cpp:57-70)
        LbHandle h = LB_NO_HANDLE;
        Check(lb_texture_create(m_Renderer, static_cast<const uint8_t*>(a_PixelData), a_Width, a_Height, a_Normalize ? 1 : 0, &h));
        return std::make_shared<Texture>(h);
    }

    using LumenRenderer::CreateMaterial;
    std::shared_ptr<Lumen::ILumenMaterial> CreateMaterial(const MaterialData& a_Data) override { return std::make_shared<Material>(m_Renderer, a_Data); }

    std::shared_ptr<Lumen::ILumenScene> CreateScene(SceneData a_Data) override
    {
        auto scene = std::make_shared<B200::Scene>(m_Renderer);
        for (auto& instance : a_Data.m_InstancedMeshes)
        {
            auto* added = scene->AddMesh();
            added->SetMesh(instance.GetMesh());
            added->m_Transform.CopyTransform(instance.m_Transform);
        }
        return scene;
    }

    // PTVolume::Load (PT/Framework/PTVolume.cpp:47-108): ".vndb" / ".nvdb" are read by the library (lb_volume_create_file); ".vdb"
    // needs the OpenVDB library and is reported as unsupported — decode it in the application and use CreateVolume(const LbVolumeDesc&)
    std::shared_ptr<Lumen::ILumenVolume> CreateVolume(const std::string& a_FilePath) override
    {
        LbHandle h = LB_NO_HANDLE;
        if (lb_volume_create_file(m_Renderer, a_FilePath.c_str(), &h) != LB_OK)
            throw std::runtime_error("lumen_b200: CreateVolume(" + a_FilePath + "): " + lb_nanovdb_last_error());
        return std::make_shared<Volume>(h);
    }
    std::shared_ptr<Lumen::ILumenVolume> CreateVolume(const LbVolumeDesc& a_Grid)
    {
        LbHandle h = LB_NO_HANDLE;
        Check(lb_volume_create(m_Renderer, &a_Grid, &h));
        return std::make_shared<Volume>(h);
    }

    void InitNGX() override {}
    unsigned int GetOutputTexture() override { return 0; }          // no GL interop: read pixels with GetOutputTexturePixels / ReadHdr

    std::vector<uint8_t> GetOutputTexturePixels(uint32_t& a_Width, uint32_t& a_Height) override
    {
        std::lock_guard<std::mutex> lock(m_FrameMutex);
        Check(lb_get_render_resolution(m_Renderer, &a_Width, &a_Height));
        std::vector<uint8_t> pixels(static_cast<size_t>(a_Width) * a_Height * 4);
        Check(lb_read_ldr(m_Renderer, pixels.data(), pixels.size()));
        return pixels;
    }
    // fp32 RGBA of the merged frame (the reference keeps it in a D3D11 surface only)
    std::vector<float> ReadHdr(uint32_t& a_Width, uint32_t& a_Height)
    {
        std::lock_guard<std::mutex> lock(m_FrameMutex);
        Check(lb_get_render_resolution(m_Renderer, &a_Width, &a_Height));
        std::vector<float> pixels(static_cast<size_t>(a_Width) * a_Height * 4);
        Check(lb_read_hdr(m_Renderer, pixels.data(), pixels.size() * sizeof(float)));
        return pixels;
    }

    void SetRenderResolution(glm::uvec2 a_Resolution) override
    {
        std::lock_guard<std::mutex> lock(m_FrameMutex);
        Check(lb_set_render_resolution(m_Renderer, a_Resolution.x, a_Resolution.y));
    }
    void SetOutputResolution(glm::uvec2 a_Resolution) override { m_OutputResolution = a_Resolution; }      // upscaling (DLSS) is out of scope
    glm::uvec2 GetRenderResolution() override
    {
        uint32_t w = 0, h = 0;
        Check(lb_get_render_resolution(m_Renderer, &w, &h));
        return glm::uvec2(w, h);
    }
    glm::uvec2 GetOutputResolution() override { return m_OutputResolution; }
    void SetBlendMode(bool a_Blend) override
    {
        std::lock_guard<std::mutex> lock(m_FrameMutex);
        Check(lb_set_blend_mode(m_Renderer, a_Blend ? 1 : 0));
    }
    bool GetBlendMode() const override
    {
        int blend = 0;
        Check(lb_get_blend_mode(m_Renderer, &blend));
        return blend != 0;
    }

    void BeginSnapshot() override {}
    std::unique_ptr<FrameSnapshot> EndSnapshot() override { return nullptr; }

    LbRenderer GetHandle() const { return m_Renderer; }

private:
    LbRenderer m_Renderer = nullptr;
    bool m_Owned = true;
    glm::uvec2 m_OutputResolution = {0, 0};
    std::thread m_Thread;
    std::atomic<bool> m_Stop{false};
    std::mutex m_FrameMutex;
    uint64_t m_FrameId = 0;
};

// One process driving n GPUs (lb_group_*, csrc/lb_multigpu.cpp): n member renderers behind the LumenRenderer interface + one NCCL
// communicator. The application loads the same assets into every Member(i) with the calls it already makes on a single renderer
// (SceneManager::LoadGLTF etc.), then Render / Reduce / ReadHdr replace the single renderer's frame loop:
//     B200::RendererGroup group({0, 1, 2, 3}, settings, LB_GROUP_SAMPLES);
//     for (uint32_t i = 0; i < group.Size(); ++i) LoadScene(group.Member(i));
//     group.Render(16); group.Reduce(); auto hdr = group.ReadHdr();          // 64 samples per pixel, one ncclReduce
class RendererGroup
{
public:
    RendererGroup(const std::vector<int>& a_Devices, const Settings& a_Settings, int a_Mode)
    {
        LbSettings s;
        std::memset(&s, 0, sizeof s);
        s.width = a_Settings.renderResolution.x; s.height = a_Settings.renderResolution.y; s.depth = a_Settings.depth;
        s.restir = a_Settings.restir; s.restir_temporal = a_Settings.restirTemporal; s.restir_spatial = a_Settings.restirSpatial; s.volume_mode = a_Settings.volumeMode;
        if (lb_group_create(a_Devices.data(), static_cast<uint32_t>(a_Devices.size()), &s, a_Mode, &m_Group) != LB_OK) throw std::runtime_error(lb_multigpu_last_error());
        m_Width = s.width; m_Height = s.height;
        for (uint32_t i = 0; i < a_Devices.size(); ++i)
        {
            LbRenderer member = nullptr;
            if (lb_group_member(m_Group, i, &member) != LB_OK) throw std::runtime_error(lb_multigpu_last_error());
            m_Members.emplace_back(new Renderer(member, a_Settings.outputResolution));
        }
    }
    ~RendererGroup() { m_Members.clear(); if (m_Group) lb_group_destroy(m_Group); }
    RendererGroup(const RendererGroup&) = delete;
    RendererGroup& operator=(const RendererGroup&) = delete;

    uint32_t Size() const { return static_cast<uint32_t>(m_Members.size()); }
    Renderer& Member(uint32_t a_Index) { return *m_Members.at(a_Index); }
    // the scene graph of every member is flushed to its renderer (MeshInstance / camera changes), then the frames are enqueued on all GPUs
    void Render(uint32_t a_Frames)
    {
        for (auto& m : m_Members) if (m->m_Scene) static_cast<B200::Scene&>(*m->m_Scene).Synchronise();
        if (lb_group_render(m_Group, a_Frames) != LB_OK) throw std::runtime_error(lb_multigpu_last_error());
    }
    void Reduce() { if (lb_group_reduce(m_Group) != LB_OK) throw std::runtime_error(lb_multigpu_last_error()); }
    void Reset() { if (lb_group_reset(m_Group) != LB_OK) throw std::runtime_error(lb_multigpu_last_error()); }
    std::vector<float> ReadHdr()
    {
        std::vector<float> pixels(static_cast<size_t>(m_Width) * m_Height * 4);
        if (lb_group_read_hdr(m_Group, pixels.data(), pixels.size() * sizeof(float)) != LB_OK) throw std::runtime_error(lb_multigpu_last_error());
        return pixels;
    }

private:
    LbGroup m_Group = nullptr;
    uint32_t m_Width = 0, m_Height = 0;
    std::vector<std::unique_ptr<Renderer>> m_Members;
};

} // namespace B200
