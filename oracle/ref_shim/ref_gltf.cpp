// ref_gltf.cpp — TEST INFRASTRUCTURE. The reference's OWN model converter (LumenPT/src/Tools/LumenPTModelConverter.cpp: GenerateContent —
// material mapping :347-531, accessor extraction :1027-1059, tangent generation :734-900, interleaving :902-928, node table :930-1025)
// compiled for the host IN PLACE together with the reference's Transform / ILumenScene / Camera sources and its vendored fx-gltf,
// nlohmann-json, glm and stb_image, by oracle/Makefile into oracle/_ref/ref_gltf. Nothing of the reference is copied into this repository.
// `#define private public` reaches the converter's private GenerateContent; an overlay copy of LumenRenderer.h (written by the Makefile
// into oracle/_ref/ov, a sed of the original) replaces the MSVC-only default argument `SceneData a_SceneData = {}`.
//   ref_gltf <in.gltf|in.glb> <out.bin> [out.ollad]      (out.ollad: GenerateHeader + OutputToFile :533-594, the reference's cache file itself)
// out.bin: u32 nMaterials, u32 sizeof(HeaderMaterial), the HeaderMaterial records; u32 nTextures, per texture u64 offset, size, type;
// u32 nMeshes, per mesh u32 nPrimitives, per primitive u64 vertexBytes, indexBytes, indexSize, materialId + the interleaved 64-byte
// vertices + the indices; u32 nScenes, per scene u32 nRoots, per node (depth first) f32 transform[16], i32 meshId, u32 nChildren.
#include <sstream>
#include <fstream>
#include <iostream>
#include <filesystem>
#include <string>
#include <vector>
#include <map>
#include <memory>
#include <unordered_map>
#include <functional>
#include <thread>
#include <mutex>
#include <future>
#include <regex>
#include <nlohmann/json.hpp>
#define private public
#include "Tools/LumenPTModelConverter.cpp"
#undef private
static void dump_node(FILE* f, const LumenPTModelConverter::HeaderNode& n) {
    fwrite(n.m_Header.m_Transform, 4, 16, f); fwrite(&n.m_Header.m_MeshId, 4, 1, f); uint32_t c = (uint32_t)n.m_ChildNodes.size(); fwrite(&c, 4, 1, f);
    for (auto& ch : n.m_ChildNodes) dump_node(f, ch);
}
int main(int argc, char** argv) {
    LumenPTModelConverter c;
    const std::string path = argv[1];
    fx::gltf::Document doc = path.size() > 4 && path.substr(path.size() - 4) == ".glb" ? fx::gltf::LoadFromBinary(path) : fx::gltf::LoadFromText(path);
    auto content = c.GenerateContent(doc, path);
    FILE* f = fopen(argv[2], "wb");
    uint32_t n = (uint32_t)content.m_Materials.size(), sz = (uint32_t)sizeof(LumenPTModelConverter::HeaderMaterial); fwrite(&n, 4, 1, f); fwrite(&sz, 4, 1, f);
    for (auto& m : content.m_Materials) fwrite(&m, sz, 1, f);
    n = (uint32_t)content.m_Textures.size(); fwrite(&n, 4, 1, f);
    for (auto& t : content.m_Textures) { uint64_t v[3] = {t.m_Offset, t.m_Size, t.m_TextureType}; fwrite(v, 8, 3, f); }
    n = (uint32_t)content.m_Meshes.size(); fwrite(&n, 4, 1, f);
    for (auto& mesh : content.m_Meshes) {
        uint32_t np = (uint32_t)mesh.m_Primitives.size(); fwrite(&np, 4, 1, f);
        for (auto& p : mesh.m_Primitives) {
            uint64_t h[4] = {p.m_VertexBufferSize, p.m_IndexBufferSize, p.m_IndexSize, p.m_MaterialId}; fwrite(h, 8, 4, f);
            fwrite(content.m_Blob.m_Data.data() + p.m_VertexBufferOffset, 1, p.m_VertexBufferSize, f);
            fwrite(content.m_Blob.m_Data.data() + p.m_IndexBufferOffset, 1, p.m_IndexBufferSize, f);
        }
    }
    n = (uint32_t)content.m_Scenes.size(); fwrite(&n, 4, 1, f);
    for (auto& s : content.m_Scenes) { uint32_t r = (uint32_t)s.m_RootNodes.size(); fwrite(&r, 4, 1, f); for (auto& rn : s.m_RootNodes) dump_node(f, rn); }
    fclose(f);
    if (argc > 3) { auto header = LumenPTModelConverter::GenerateHeader(content); LumenPTModelConverter::OutputToFile(header, content.m_Blob, argv[3]); }
    fprintf(stderr, "materials %zu meshes %zu textures %zu scenes %zu blob %llu\n", content.m_Materials.size(), content.m_Meshes.size(), content.m_Textures.size(), content.m_Scenes.size(), (unsigned long long)content.m_Blob.m_Offset);
    return 0;
}
std::shared_ptr<Lumen::ILumenScene> LumenRenderer::CreateScene() { return nullptr; }      // the overlay overload; LoadFile (not used here) calls it
