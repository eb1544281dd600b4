// adapter_driver.cpp — exercises include/lumen_b200_adapter.hpp the way the reference's Sandbox uses its renderer: every call
// below goes through the reference's OWN interface types (LumenRenderer, ILumenScene, MeshInstance, Transform, Camera — compiled
// from /root/reference in place, see tests/adapter/build.py), never through lumen_b200.h directly.
//
//   adapter_driver <scene.bin> <out-prefix>
// reads a scene dump written by tests/test_adapter.py, builds it with CreateTexture / CreateMaterial / CreatePrimitive /
// CreateMesh / m_Scene->AddMesh(), renders `frames` frames and writes <out>.ldr (GetOutputTexturePixels), <out>.hdr and
// <out>.worlds (the row-major world matrices of the instances and of the camera as the reference's Transform / Camera computed
// them, so that the Python side can render the identical scene through the plain C ABI and compare bit for bit).
#include <lumen_b200_adapter.hpp>

#include <cstdio>
#include <fstream>
#include <iostream>

namespace {
struct Reader {
    std::ifstream in;
    explicit Reader(const char* path) : in(path, std::ios::binary) { if (!in) throw std::runtime_error(std::string("cannot open ") + path); }
    template <class T> T get() { T v; in.read(reinterpret_cast<char*>(&v), sizeof v); if (!in) throw std::runtime_error("scene dump truncated"); return v; }
    std::vector<uint8_t> bytes(size_t n) { std::vector<uint8_t> v(n); in.read(reinterpret_cast<char*>(v.data()), (std::streamsize)n); if (!in) throw std::runtime_error("scene dump truncated"); return v; }
};
void write_file(const std::string& path, const void* data, size_t n) { std::ofstream o(path, std::ios::binary); o.write(static_cast<const char*>(data), (std::streamsize)n); }
}

int main(int argc, char** argv) try {
    if (argc < 3) { std::fprintf(stderr, "usage: adapter_driver scene.bin out-prefix\n"); return 2; }
    Reader rd(argv[1]);
    const std::string out = argv[2];
    if (rd.get<uint32_t>() != 0x4353424Cu) throw std::runtime_error("not a scene dump");
    B200::Settings settings;
    settings.renderResolution.x = rd.get<uint32_t>(); settings.renderResolution.y = rd.get<uint32_t>();
    settings.outputResolution = settings.renderResolution;
    settings.depth = rd.get<uint32_t>();
    settings.restir = rd.get<uint32_t>() != 0;
    const uint32_t frames = rd.get<uint32_t>();
    const uint32_t interleaved = rd.get<uint32_t>();

    // Sandbox/src/Application.cpp:83 — the one line that changes
    std::shared_ptr<LumenRenderer> renderer = std::make_shared<B200::Renderer>(settings);

    const glm::vec3 camPosition = rd.get<glm::vec3>();
    const float qw = rd.get<float>(), qx = rd.get<float>(), qy = rd.get<float>(), qz = rd.get<float>();

    std::vector<std::shared_ptr<Lumen::ILumenTexture>> textures;
    for (uint32_t n = rd.get<uint32_t>(), i = 0; i < n; ++i) {
        const uint32_t w = rd.get<uint32_t>(), h = rd.get<uint32_t>(), srgb = rd.get<uint32_t>();
        auto px = rd.bytes(size_t(w) * h * 4);
        textures.push_back(renderer->CreateTexture(px.data(), w, h, srgb != 0));
    }
    auto tex = [&](int32_t i) { return i >= 0 ? textures.at(size_t(i)) : std::shared_ptr<Lumen::ILumenTexture>(); };

    std::vector<std::shared_ptr<Lumen::ILumenMaterial>> materials;
    for (uint32_t n = rd.get<uint32_t>(), i = 0; i < n; ++i) {
        LumenRenderer::MaterialData m;
        m.m_DiffuseColor = rd.get<glm::vec4>(); m.m_EmissionVal = rd.get<glm::vec3>();
        m.m_TransmissionFactor = rd.get<float>(); m.m_ClearCoatFactor = rd.get<float>(); m.m_ClearCoatRoughnessFactor = rd.get<float>();
        m.m_IndexOfRefraction = rd.get<float>(); m.m_SpecularFactor = rd.get<float>(); m.m_SpecularTintFactor = rd.get<float>();
        m.m_SubSurfaceFactor = rd.get<float>(); m.m_Luminance = rd.get<float>(); m.m_Anisotropic = rd.get<float>();
        m.m_SheenFactor = rd.get<float>(); m.m_SheenTintFactor = rd.get<float>(); m.m_MetallicFactor = rd.get<float>(); m.m_RoughnessFactor = rd.get<float>();
        m.m_TintFactor = rd.get<glm::vec3>(); m.m_Transmittance = rd.get<glm::vec3>();
        m.m_DiffuseTexture = tex(rd.get<int32_t>()); m.m_NormalMap = tex(rd.get<int32_t>()); m.m_MetallicRoughnessTexture = tex(rd.get<int32_t>());
        m.m_EmissiveTexture = tex(rd.get<int32_t>()); m.m_TransmissionTexture = tex(rd.get<int32_t>()); m.m_ClearCoatTexture = tex(rd.get<int32_t>());
        m.m_ClearCoatRoughnessTexture = tex(rd.get<int32_t>()); m.m_TintTexture = tex(rd.get<int32_t>());
        auto material = renderer->CreateMaterial(m);
        // the setter path must give the same device material as the constructor path: re-apply two factors through ILumenMaterial
        material->SetRoughnessFactor(m.m_RoughnessFactor);
        material->SetDiffuseColor(m.m_DiffuseColor);
        materials.push_back(material);
    }

    std::vector<std::shared_ptr<Lumen::ILumenMesh>> meshes;
    for (uint32_t n = rd.get<uint32_t>(), i = 0; i < n; ++i) {
        std::vector<std::shared_ptr<Lumen::ILumenPrimitive>> prims;
        for (uint32_t np = rd.get<uint32_t>(), p = 0; p < np; ++p) {
            const uint32_t nv = rd.get<uint32_t>(), ni = rd.get<uint32_t>(), material = rd.get<uint32_t>(), indexSize = rd.get<uint32_t>();
            std::vector<uint8_t> pos = rd.bytes(size_t(nv) * 12), uv = rd.bytes(size_t(nv) * 8), nrm = rd.bytes(size_t(nv) * 12), tan = rd.bytes(size_t(nv) * 16);
            LumenRenderer::PrimitiveData data;
            data.m_IndexBinary = rd.bytes(size_t(ni) * indexSize);
            data.m_IndexSize = indexSize;
            data.m_Material = materials.at(material);
            if (interleaved) {
                // the 64-byte `Vertex` of LumenPT/src/Shaders/CppCommon/ModelStructs.h:21-28 (what LumenPTModelConverter emits)
                data.m_Interleaved = true;
                data.m_VertexBinary.assign(size_t(nv) * 64, 0);
                for (uint32_t v = 0; v < nv; ++v) {
                    uint8_t* dst = data.m_VertexBinary.data() + size_t(v) * 64;
                    std::memcpy(dst, &pos[size_t(v) * 12], 12); std::memcpy(dst + 16, &uv[size_t(v) * 8], 8);
                    std::memcpy(dst + 24, &nrm[size_t(v) * 12], 12); std::memcpy(dst + 48, &tan[size_t(v) * 16], 16);
                }
            } else {
                data.m_Positions = pos; data.m_TexCoords = uv; data.m_Normals = nrm; data.m_Tangents = tan;
            }
            prims.push_back(renderer->CreatePrimitive(data));
        }
        meshes.push_back(renderer->CreateMesh(prims));
    }

    auto& scene = *renderer->m_Scene;
    std::vector<Lumen::MeshInstance*> instances;
    for (uint32_t n = rd.get<uint32_t>(), i = 0; i < n; ++i) {
        const uint32_t mesh = rd.get<uint32_t>();
        const glm::mat4 rowMajor = rd.get<glm::mat4>();
        const int32_t mode = rd.get<int32_t>(); const glm::vec3 radiance = rd.get<glm::vec3>(); const float scale = rd.get<float>();
        const int32_t overrideMaterial = rd.get<int32_t>();
        Lumen::MeshInstance* inst = scene.AddMesh();
        inst->SetMesh(meshes.at(mesh));
        inst->m_Transform = glm::transpose(rowMajor);                   // Transform::operator=(const glm::mat4&) decomposes into T, R, S
        if (mode != 0) inst->SetEmissiveness(Lumen::MeshInstance::Emissiveness(static_cast<Lumen::EmissionMode>(mode), radiance, scale));
        if (overrideMaterial >= 0) inst->SetOverrideMaterial(materials.at(size_t(overrideMaterial)));
        instances.push_back(inst);
    }
    // optional trailing section: volumes loaded from files through LumenRenderer::CreateVolume(path) + ILumenScene::AddVolume()
    std::vector<Lumen::VolumeInstance*> volumeInstances;
    if (rd.in.peek() != std::char_traits<char>::eof()) {
        for (uint32_t n = rd.get<uint32_t>(), i = 0; i < n; ++i) {
            const std::vector<uint8_t> path = rd.bytes(rd.get<uint32_t>());
            const glm::mat4 rowMajor = rd.get<glm::mat4>();
            const float density = rd.get<float>();
            Lumen::VolumeInstance* inst = scene.AddVolume();
            inst->SetVolume(renderer->CreateVolume(std::string(path.begin(), path.end())));
            inst->m_Transform = glm::transpose(rowMajor);
            inst->m_Density = density;
            volumeInstances.push_back(inst);
        }
    }
    scene.m_Camera->SetRotation(glm::quat(qw, qx, qy, qz));
    scene.m_Camera->SetPosition(camPosition);                           // SetPosition raises the dirty flag, SetRotation does not (Camera.cpp:32-45)
    scene.m_Camera->SetAspectRatio(float(settings.renderResolution.x) / float(settings.renderResolution.y));

    auto& b200 = static_cast<B200::Renderer&>(*renderer);
    for (uint32_t f = 0; f < frames; ++f) b200.TraceFrame();

    uint32_t w = 0, h = 0;
    const std::vector<uint8_t> ldr = renderer->GetOutputTexturePixels(w, h);
    const std::vector<float> hdr = b200.ReadHdr(w, h);
    std::vector<float> worlds;
    for (auto* inst : instances) { float m[16]; B200::RowMajor(inst->m_Transform.GetWorldTransformationMatrix(), m); worlds.insert(worlds.end(), m, m + 16); }
    { glm::mat4 prev, cur; scene.m_Camera->GetMatrixData(prev, cur); float m[16]; B200::RowMajor(cur, m); worlds.insert(worlds.end(), m, m + 16); }
    for (auto* inst : volumeInstances) { float m[16]; B200::RowMajor(inst->m_Transform.GetWorldTransformationMatrix(), m); worlds.insert(worlds.end(), m, m + 16); }
    write_file(out + ".ldr", ldr.data(), ldr.size());
    write_file(out + ".hdr", hdr.data(), hdr.size() * sizeof(float));
    write_file(out + ".worlds", worlds.data(), worlds.size() * sizeof(float));
    const FrameStats stats = renderer->GetLastFrameStats();
    std::printf("frames %u resolution %ux%u instances %zu frame-id %llu stages %zu\n", frames, w, h, instances.size(), (unsigned long long)stats.m_Id, stats.m_Times.size());
    return 0;
} catch (const std::exception& e) {
    std::fprintf(stderr, "adapter_driver: %s\n", e.what());
    return 1;
}
