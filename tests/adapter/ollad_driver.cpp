// ollad_driver.cpp — TEST INFRASTRUCTURE. The reference's OWN cache loader (LumenPTModelConverter::LoadFile + LoadNode,
// LumenPT/src/Tools/LumenPTModelConverter.cpp:70-317, compiled in place from /root/reference together with its stb_image) reads an
// `.ollad` file and feeds it to include/lumen_b200_adapter.hpp through the reference interface alone: SetRendererRef (default
// textures), CreateTexture (stb-decoded pixels), CreateMaterial, CreatePrimitive (interleaved 64-byte vertices), CreateMesh,
// CreateScene, scene->AddMesh(), Transform::AddChild. The scene it returns is rendered.
//
//   ollad_driver <file.ollad> <out-prefix> <width> <height> <depth> <restir> <frames> <px py pz> <qw qx qy qz>
// writes <out>.hdr, <out>.ldr and <out>.worlds: the row-major world matrix of every mesh instance in AddMesh order — what the
// reference's Transform hierarchy makes of the node table — then the camera's. tests/test_adapter.py compares the matrices with
// lb_gltf_instance and the image with the same file uploaded by lb_gltf_upload through the plain C ABI.
#include <lumen_b200_adapter.hpp>

#include <sstream>
#include <fstream>
#include <iostream>
#include <filesystem>
#include <map>
#include <unordered_map>
#include <functional>
#include <thread>
#include <mutex>
#include <future>
#include <regex>
#include <nlohmann/json.hpp>
#include "Tools/LumenPTModelConverter.cpp"

// LoadFile calls CreateScene() with the default argument g++ cannot accept in the class (see build.py); the overlay header declares this overload instead
std::shared_ptr<Lumen::ILumenScene> LumenRenderer::CreateScene() { return CreateScene(SceneData{}); }

int main(int argc, char** argv) try {
    if (argc < 15) { std::fprintf(stderr, "usage: ollad_driver file.ollad out-prefix width height depth restir frames px py pz qw qx qy qz\n"); return 2; }
    const std::string out = argv[2];
    B200::Settings settings;
    settings.renderResolution.x = (uint32_t)std::atoi(argv[3]); settings.renderResolution.y = (uint32_t)std::atoi(argv[4]);
    settings.outputResolution = settings.renderResolution;
    settings.depth = (uint32_t)std::atoi(argv[5]);
    settings.restir = std::atoi(argv[6]) != 0;
    const int frames = std::atoi(argv[7]);
    float v[7]; for (int k = 0; k < 7; ++k) v[k] = std::strtof(argv[8 + k], nullptr);

    std::shared_ptr<LumenRenderer> renderer = std::make_shared<B200::Renderer>(settings);
    LumenPTModelConverter converter;
    converter.SetRendererRef(*renderer);
    Lumen::SceneManager::GLTFResource resource = converter.LoadFile(argv[1]);
    if (resource.m_Path.empty() || resource.m_Scenes.empty()) throw std::runtime_error("LoadFile returned no scene");
    renderer->m_Scene = resource.m_Scenes[0];
    auto& scene = *renderer->m_Scene;
    scene.m_Camera->SetRotation(glm::quat(v[3], v[4], v[5], v[6]));
    scene.m_Camera->SetPosition(glm::vec3(v[0], v[1], v[2]));
    scene.m_Camera->SetAspectRatio(float(settings.renderResolution.x) / float(settings.renderResolution.y));

    auto& b200 = static_cast<B200::Renderer&>(*renderer);
    for (int f = 0; f < frames; ++f) b200.TraceFrame();

    uint32_t w = 0, h = 0;
    const std::vector<uint8_t> ldr = renderer->GetOutputTexturePixels(w, h);
    const std::vector<float> hdr = b200.ReadHdr(w, h);
    std::vector<float> worlds;
    for (auto& inst : scene.m_MeshInstances) { float m[16]; B200::RowMajor(inst->m_Transform.GetWorldTransformationMatrix(), m); worlds.insert(worlds.end(), m, m + 16); }
    { glm::mat4 prev, cur; scene.m_Camera->GetMatrixData(prev, cur); float m[16]; B200::RowMajor(cur, m); worlds.insert(worlds.end(), m, m + 16); }
    auto write_file = [](const std::string& path, const void* data, size_t n) { std::ofstream o(path, std::ios::binary); o.write(static_cast<const char*>(data), (std::streamsize)n); };
    write_file(out + ".ldr", ldr.data(), ldr.size());
    write_file(out + ".hdr", hdr.data(), hdr.size() * sizeof(float));
    write_file(out + ".worlds", worlds.data(), worlds.size() * sizeof(float));
    std::printf("loaded %zu materials %zu meshes %zu instances, resolution %ux%u\n", resource.m_MaterialPool.size(), resource.m_MeshPool.size(), scene.m_MeshInstances.size(), w, h);
    return 0;
} catch (const std::exception& e) {
    std::fprintf(stderr, "ollad_driver: %s\n", e.what());
    return 1;
}
