// lb_gltf.cpp — glTF 2.0 ingest in front of the path: file -> vertex streams, Disney material parameters, textures
// and mesh instances, handed to the renderer through the same C ABI an application would use. Host code only.
//
// Restates what the reference does between a .gltf/.glb file and LumenRenderer::CreateTexture/CreateMaterial/CreatePrimitive/
// CreateMesh/AddMesh (paths under /root/reference/Lumen_Engine/):
//   material mapping        LumenPT/src/Tools/LumenPTModelConverter.cpp:347-531  (pbrMetallicRoughness + KHR_materials_transmission /
//                           sheen / ior / clearcoat / specular -> Disney parameters; roughness >= 0.01; luminance 1, transmittance 0)
//   texture typing          :358-398 (image = textures[i].source), :121-135 (RGBA8, green >= 1 for metal-roughness images, sRGB only
//                           for diffuse and emissive images)
//   accessor extraction     :1027-1059 LoadBinary
//   tangent generation      :734-900 GenerateTangentBinary (default UVs, per-vertex Gram-Schmidt against the vertex normal, w = 1,
//                           the LAST triangle that references a vertex wins)
//   node hierarchy          :953-1025 LoadNode/LoadNodeTransform (matrix, else T*R*S), :275-317 (mesh nodes become mesh instances whose
//                           parent is the parent node's transform), Lumen/src/Lumen/ModelLoading/Transform.cpp:264-307 (world = parent * local)
// The intermediate .ollad cache file (ConvertGLTF/OutputToFile) is a serialisation of exactly this content and is not reproduced.
//
// Canonical choices where the reference is undefined or lossy (DESIGN.md "glTF ingest"):
//   * node transforms are used as written (the reference round-trips every matrix through glm::decompose);
//   * a primitive without NORMAL gets flat face normals (the reference reads an empty buffer);
//   * tangents are stored per VERTEX (the reference sizes the buffer by the index count and indexes it by vertex);
//   * 8-bit indices are widened (CreatePrimitive takes 2- or 4-byte indices); strided accessors are read with their stride
//     (LoadBinary advances source and destination by the same stride, which is only right for tightly packed views);
//   * the quirk that the children of a MESH node do not inherit that node's ancestors (:296-306) is reproduced.
// Images: PNG is decoded here (lb_png.h); any other format goes through the caller's decoder callback (the reference links
// stb_image); without one the image becomes the 1x1 default and is counted in LbGltfInfo::undecoded_images.
#include "../../include/lumen_b200.h"
#include "lb_json.h"
#include <unistd.h>
#include "lb_png.h"
#include "lb_jpeg.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {

thread_local std::string g_error;
int gfail(int code, const std::string& msg) { g_error = msg; return code; }

struct Vec3 { float x, y, z; };
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator*(float s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator/(Vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }                 // glm: component-wise division
inline float dot3(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }               // glm::dot: tmp.x + tmp.y + tmp.z
inline Vec3 normalize3(Vec3 v) { const float inv = 1.0f / sqrtf(dot3(v, v)); return v * inv; }  // glm::normalize = v * inversesqrt(dot)
inline Vec3 cross3(Vec3 a, Vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }

// column-major 4x4 (glm::mat4 layout): m[col * 4 + row]
struct Mat4 { float m[16]; };
Mat4 mat_identity() { Mat4 r{}; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.f; return r; }
bool mat_is_identity(const Mat4& a) { const Mat4 i = mat_identity(); for (int k = 0; k < 16; ++k) if (!(a.m[k] == i.m[k])) return false; return true; }   // glm's ==: by value (-0 == 0)
// glm operator*(mat4, mat4): Result[j] = ((A[0]*B[j][0] + A[1]*B[j][1]) + A[2]*B[j][2]) + A[3]*B[j][3]
Mat4 mat_mul(const Mat4& a, const Mat4& b) {
    Mat4 r;
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i)
        r.m[j * 4 + i] = ((a.m[0 * 4 + i] * b.m[j * 4 + 0] + a.m[1 * 4 + i] * b.m[j * 4 + 1]) + a.m[2 * 4 + i] * b.m[j * 4 + 2]) + a.m[3 * 4 + i] * b.m[j * 4 + 3];
    return r;
}
// Transform::UpdateLocalMatrix (Transform.cpp:264-280) in glm's own operation order, so that signed zeros come out as the reference's do:
// glm::translate(I, t) -> operator*(mat4, mat4_cast(q)) -> glm::scale (every column, w included, times its factor)
Mat4 mat_trs(const float t[3], const float q[4] /* x y z w */, const float s[3]) {
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float qxx = x * x, qyy = y * y, qzz = z * z, qxz = x * z, qxy = x * y, qyz = y * z, qwx = w * x, qwy = w * y, qwz = w * z;
    Mat4 rot{};
    rot.m[0] = 1.f - 2.f * (qyy + qzz); rot.m[1] = 2.f * (qxy + qwz); rot.m[2] = 2.f * (qxz - qwy);
    rot.m[4] = 2.f * (qxy - qwz); rot.m[5] = 1.f - 2.f * (qxx + qzz); rot.m[6] = 2.f * (qyz + qwx);
    rot.m[8] = 2.f * (qxz + qwy); rot.m[9] = 2.f * (qyz - qwx); rot.m[10] = 1.f - 2.f * (qxx + qyy);
    rot.m[15] = 1.f;
    const Mat4 id = mat_identity();
    Mat4 tr = id;
    for (int i = 0; i < 4; ++i) tr.m[12 + i] = ((id.m[i] * t[0] + id.m[4 + i] * t[1]) + id.m[8 + i] * t[2]) + id.m[12 + i];
    Mat4 r = mat_mul(tr, rot);
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 4; ++i) r.m[j * 4 + i] = r.m[j * 4 + i] * s[j];
    return r;
}

// LumenPTModelConverter::TextureType (LumenPTModelConverter.h:66-77); the LAST role a material assigns to an image wins (:364-521)
enum TextureType : uint64_t { T_UNSPECIFIED = 0, T_DIFFUSE = 1, T_NORMAL, T_EMISSIVE, T_METAL_ROUGHNESS, T_TRANSMISSIVE, T_CLEAR_COAT, T_CLEAR_COAT_ROUGHNESS, T_TINT };
struct Image {
    std::vector<uint8_t> px; uint32_t w = 1, h = 1; bool decoded = false;
    std::vector<uint8_t> raw;                       // the encoded file as the document holds it (what the .ollad cache stores)
    uint64_t type = T_UNSPECIFIED;
    bool srgb() const { return type == T_DIFFUSE || type == T_EMISSIVE; }      // LoadFile :131
    bool metal_rough() const { return type == T_METAL_ROUGHNESS; }             // LoadFile :122
};
struct Primitive { std::vector<float> pos, uv, nrm, tan; std::vector<uint32_t> idx; int32_t material = -1; uint32_t index_size = 4; /* bytes per index in the source */ };
struct Mesh { std::vector<Primitive> prims; };
struct Instance { uint32_t mesh; float m[16]; /* row-major */ };
// HeaderNode / HeaderScene (LumenPTModelConverter.h:163-186): the node table with LOCAL matrices, kept for the .ollad cache
struct Node { std::string name; Mat4 local; int32_t mesh = -1; std::vector<Node> children; };
struct Scene { std::string name; std::vector<Node> roots; };

bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const bool ok = out.empty() || fread(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    return ok;
}
bool base64_decode(const char* s, size_t n, std::vector<uint8_t>& out) {
    uint32_t acc = 0; int bits = 0;
    for (size_t i = 0; i < n; ++i) {
        const char c = s[i]; int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A'; else if (c >= 'a' && c <= 'z') v = c - 'a' + 26; else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+' || c == '-') v = 62; else if (c == '/' || c == '_') v = 63; else if (c == '=' || c == '\n' || c == '\r') continue; else return false;
        acc = (acc << 6) | (uint32_t)v; bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back((uint8_t)(acc >> bits)); }
    }
    return true;
}
std::string uri_unescape(const std::string& s) {
    std::string o;
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] == '%' && i + 2 < s.size() + 0 && isxdigit((unsigned char)s[i + 1]) && isxdigit((unsigned char)s[i + 2])) { o += (char)strtol(s.substr(i + 1, 2).c_str(), nullptr, 16); i += 2; }
        else o += s[i];
    }
    return o;
}

} // namespace

struct LbGltfOpaque {
    std::vector<Image> images; std::vector<LbMaterialDesc> materials; std::vector<Mesh> meshes; std::vector<Instance> instances;
    std::vector<Scene> scenes;
    LbGltfInfo info{};
};

namespace {
// stb_image in the reference (LoadFile :121); here PNG and JPEG natively (lb_png.h, lb_jpeg.h: both pinned on the reference's stb_image), anything
// else (BMP, TGA, ...) through the caller's decoder, else the 1x1 default
void decode_image(LbGltfOpaque& g, Image& im, LbImageDecodeFn decoder, void* user) {
    const bool have = !im.raw.empty();
    if (have && lb::png::decode_rgba8(im.raw.data(), im.raw.size(), im.px, im.w, im.h)) im.decoded = true;
    else if (have && lb::jpeg::decode_rgba8(im.raw.data(), im.raw.size(), im.px, im.w, im.h)) im.decoded = true;
    else if (have && decoder) {
        uint8_t* px = nullptr; uint32_t w = 0, h = 0;
        if (decoder(im.raw.data(), im.raw.size(), &px, &w, &h, user) == 0 && px && w && h) { im.px.assign(px, px + (size_t)w * h * 4); im.w = w; im.h = h; im.decoded = true; }
        free(px);
    }
    if (!im.decoded) { im.px = {255, 255, 255, 255}; im.w = im.h = 1; ++g.info.undecoded_images; }
}
// :127-134 metal-roughness images: green (roughness) is at least 1/255
void floor_roughness(LbGltfOpaque& g) {
    for (Image& im : g.images) if (im.metal_rough()) for (size_t p = 0; p < im.px.size(); p += 4) if (im.px[p + 1] < 1) im.px[p + 1] = 1;
}
// LoadNode (:275-317): a mesh node's instance transform is parented to the enclosing node's transform, but the children of a MESH node
// are parented to that node's own (unparented) transform. `parent_world`: world matrix of the parent node's transform object.
// A ROOT mesh node's instance keeps an IDENTITY world matrix: `m->m_Transform = node->m_Transform` (:293) is Transform::operator=
// (LM/ModelLoading/Transform.cpp:58-75), which copies the local matrix and a clean dirty flag but not the world matrix, and only
// AddChild (:298-299, parented nodes) raises the flag again. This is what the reference renders (its Sponza, one root mesh node with
// scale 0.008, appears unscaled; Sandbox/src/Application.cpp:146 places the camera accordingly) and what running the reference's own
// LoadFile shows (tests/test_adapter.py::test_reference_ollad_loader_over_the_adapter_cpu).
void flatten_node(LbGltfOpaque& g, const Node& n, const Mat4* parent_world) {
    const Mat4 with_parent = parent_world ? mat_mul(*parent_world, n.local) : n.local;
    Mat4 own_world;
    if (n.mesh >= 0) {
        Instance in; in.mesh = (uint32_t)n.mesh;
        const Mat4 shown = parent_world ? with_parent : mat_identity();
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) in.m[r * 4 + c] = shown.m[c * 4 + r];
        g.instances.push_back(in);
        own_world = n.local;
    } else own_world = with_parent;
    for (const Node& c : n.children) flatten_node(g, c, &own_world);
}
void flatten_scenes(LbGltfOpaque& g) {
    g.instances.clear();
    for (const Scene& s : g.scenes) for (const Node& r : s.roots) flatten_node(g, r, nullptr);
    g.info.instances = (uint32_t)g.instances.size();
}
} // namespace

namespace {

using lb::json::Value;

struct Loader {
    LbGltfOpaque& g; const Value& doc; std::string dir; std::vector<std::vector<uint8_t>> buffers; std::vector<uint8_t> glb_bin;
    LbImageDecodeFn decoder; void* user;

    [[noreturn]] void bad(const std::string& what) const { throw std::runtime_error("glTF: " + what); }

    bool load_uri(const std::string& uri, std::vector<uint8_t>& out) const {
        if (uri.compare(0, 5, "data:") == 0) {
            const size_t comma = uri.find(',');
            if (comma == std::string::npos || uri.find(";base64") == std::string::npos) return false;
            return base64_decode(uri.c_str() + comma + 1, uri.size() - comma - 1, out);
        }
        return read_file(dir + uri_unescape(uri), out);
    }
    void load_buffers() {
        const Value& bs = doc["buffers"];
        for (size_t i = 0; i < bs.size(); ++i) {
            std::vector<uint8_t> data;
            if (bs[i].has("uri")) { if (!load_uri(bs[i]["uri"].string(), data)) bad("cannot read buffer " + std::to_string(i)); }
            else if (i == 0 && !glb_bin.empty()) data = glb_bin;
            else bad("buffer " + std::to_string(i) + " has no data");
            if ((int64_t)data.size() < bs[i]["byteLength"].integer(0)) bad("buffer " + std::to_string(i) + " is shorter than its byteLength");
            buffers.push_back(std::move(data));
        }
    }
    static uint32_t component_size(int64_t t) { return (t == 5120 || t == 5121) ? 1u : (t == 5122 || t == 5123) ? 2u : (t == 5125 || t == 5126) ? 4u : 0u; }
    static uint32_t component_count(const std::string& t) { return t == "SCALAR" ? 1u : t == "VEC2" ? 2u : t == "VEC3" ? 3u : t == "VEC4" ? 4u : t == "MAT2" ? 4u : t == "MAT3" ? 9u : t == "MAT4" ? 16u : 0u; }
    // LoadBinary (:1027-1059): tightly packed copy of the accessor's elements
    std::vector<uint8_t> accessor_bytes(int64_t index, uint32_t& comp_size, uint32_t& comp_count, int64_t& comp_type) const {
        const Value& acc = doc["accessors"][(size_t)index];
        if (index < 0 || acc.is_null()) bad("accessor index out of range");
        if (acc.has("sparse")) bad("sparse accessors are not supported");
        comp_type = acc["componentType"].integer(0); comp_size = component_size(comp_type); comp_count = component_count(acc["type"].string());
        if (!comp_size || !comp_count) bad("accessor with unknown component type");
        const int64_t count_field = acc["count"].integer(0);
        if (count_field < 0 || count_field > ((int64_t)1 << 31)) bad("accessor count out of range");
        const size_t count = (size_t)count_field, elem = (size_t)comp_size * comp_count;
        if (!acc.has("bufferView")) return std::vector<uint8_t>(count * elem, 0);
        const Value& view = doc["bufferViews"][(size_t)acc["bufferView"].integer(-1)];
        if (view.is_null()) bad("bufferView index out of range");
        const size_t buf = (size_t)view["buffer"].integer(0);
        if (buf >= buffers.size()) bad("buffer index out of range");
        const int64_t stride_field = view["byteStride"].integer(0), off_view = view["byteOffset"].integer(0), off_acc = acc["byteOffset"].integer(0);
        if (stride_field < 0 || stride_field > 65536 || off_view < 0 || off_acc < 0) bad("negative or implausible bufferView / accessor offsets");
        const size_t stride = std::max<size_t>(elem, (size_t)stride_field);
        const size_t base = (size_t)off_view + (size_t)off_acc, size = buffers[buf].size();
        // overflow-safe form of: base + (count - 1) * stride + elem <= size
        if (count && (base > size || elem > size - base || count - 1 > (size - base - elem) / stride)) bad("accessor reads past the end of its buffer");
        std::vector<uint8_t> out(count * elem, 0);
        for (size_t i = 0; i < count; ++i) memcpy(out.data() + i * elem, buffers[buf].data() + base + i * stride, elem);
        return out;
    }
    std::vector<float> float_attribute(int64_t index, uint32_t want_count, const char* name) const {
        uint32_t cs, cc; int64_t ct;
        const std::vector<uint8_t> raw = accessor_bytes(index, cs, cc, ct);
        if (ct != 5126 || cc != want_count) bad(std::string(name) + " must be a float accessor of the expected width");
        std::vector<float> out(raw.size() / 4);
        if (!out.empty()) memcpy(out.data(), raw.data(), out.size() * 4);
        return out;
    }

    void load_images() {
        const Value& imgs = doc["images"];
        g.images.resize(imgs.size());
        for (size_t i = 0; i < imgs.size(); ++i) {
            std::vector<uint8_t> bytes; bool have = false;
            if (imgs[i].has("uri")) have = load_uri(imgs[i]["uri"].string(), bytes);
            else if (imgs[i].has("bufferView")) {
                const Value& view = doc["bufferViews"][(size_t)imgs[i]["bufferView"].integer(-1)];
                // signed reads + the overflow-safe range test of accessor_bytes: a negative or huge byteOffset / byteLength must not wrap
                const int64_t buf = view["buffer"].integer(0), off = view["byteOffset"].integer(0), len = view["byteLength"].integer(0);
                if (!view.is_null() && buf >= 0 && (uint64_t)buf < buffers.size() && off >= 0 && len >= 0) {
                    const std::vector<uint8_t>& src = buffers[(size_t)buf];
                    if ((uint64_t)off <= src.size() && (uint64_t)len <= src.size() - (uint64_t)off) { bytes.assign(src.begin() + off, src.begin() + off + len); have = true; }
                }
            }
            Image& im = g.images[i];
            if (have) im.raw = std::move(bytes);
            decode_image(g, im, decoder, user);
        }
        g.info.images = (uint32_t)g.images.size();
    }
    // glTF texture index -> image index (textures[i].source), -1 when absent
    int32_t texture_image(const Value& tex_info) const {
        if (!tex_info.is_object() || !tex_info.has("index")) return LB_NO_HANDLE;
        const Value& t = doc["textures"][(size_t)tex_info["index"].integer(-1)];
        const int64_t src = t.is_null() ? -1 : t["source"].integer(-1);
        return (src >= 0 && src < (int64_t)g.images.size()) ? (int32_t)src : LB_NO_HANDLE;
    }
    // JsonGetOrDefault<uint32_t>(json, "xTexture", -1) reads the member as a plain number (:428); glTF stores a textureInfo
    // object there. Both spellings are accepted.
    int32_t extension_texture(const Value& ext, const char* key) const {
        const Value& v = ext[key];
        if (v.kind == Value::Number) { const Value& t = doc["textures"][(size_t)v.integer(-1)]; const int64_t src = t.is_null() ? -1 : t["source"].integer(-1); return (src >= 0 && src < (int64_t)g.images.size()) ? (int32_t)src : LB_NO_HANDLE; }
        return texture_image(v);
    }
    void load_materials() {
        const Value& mats = doc["materials"];
        for (size_t i = 0; i < mats.size(); ++i) {
            const Value& m = mats[i]; const Value& pbr = m["pbrMetallicRoughness"];
            LbMaterialDesc d{};
            for (int k = 0; k < 4; ++k) d.diffuse_color[k] = (float)pbr["baseColorFactor"][(size_t)k].number(1.0);
            for (int k = 0; k < 3; ++k) d.emission[k] = (float)m["emissiveFactor"][(size_t)k].number(0.0);
            d.diffuse_texture = texture_image(pbr["baseColorTexture"]);
            d.normal_texture = texture_image(m["normalTexture"]);
            d.metallic_roughness_texture = texture_image(pbr["metallicRoughnessTexture"]);
            d.emissive_texture = texture_image(m["emissiveTexture"]);
            auto role = [&](int32_t image, TextureType t) { if (image >= 0) g.images[image].type = t; };      // assignment order of :364-521
            role(d.diffuse_texture, T_DIFFUSE); role(d.normal_texture, T_NORMAL); role(d.metallic_roughness_texture, T_METAL_ROUGHNESS); role(d.emissive_texture, T_EMISSIVE);
            d.metallic_factor = (float)pbr["metallicFactor"].number(1.0);
            d.roughness_factor = fmaxf(0.01f, (float)pbr["roughnessFactor"].number(1.0));
            d.luminance = 1.f; d.subsurface_factor = 0.f; d.anisotropic = 0.f;
            d.tint_factor[0] = d.tint_factor[1] = d.tint_factor[2] = 0.f;      // HeaderMaterial::m_TintFactor is never written by the converter (zero-initialised)
            const Value& ext = m["extensions"];
            const Value& tr = ext["KHR_materials_transmission"];
            d.transmission_factor = tr.is_object() ? (float)tr["transmissionFactor"].number(0.0) : 0.f;
            d.transmission_texture = tr.is_object() ? extension_texture(tr, "transmissionTexture") : LB_NO_HANDLE;
            const Value& sh = ext["KHR_materials_sheen"];
            d.sheen_factor = sh.is_object() ? (float)sh["sheenRoughnessFactor"].number(0.0) : 0.f;
            d.sheen_tint_factor = sh.is_object() ? 1.f : 0.f;
            const Value& ior = ext["KHR_materials_ior"];
            d.index_of_refraction = ior.is_object() ? (float)ior["ior"].number(1.0) : 1.f;
            const Value& cc = ext["KHR_materials_clearcoat"];
            d.clear_coat_factor = cc.is_object() ? (float)cc["clearcoatFactor"].number(0.0) : 0.f;
            d.clear_coat_roughness_factor = cc.is_object() ? (float)cc["clearcoatRoughnessFactor"].number(0.0) : 0.f;
            d.clear_coat_texture = cc.is_object() ? extension_texture(cc, "clearcoatTexture") : LB_NO_HANDLE;
            d.clear_coat_roughness_texture = cc.is_object() ? extension_texture(cc, "clearcoatRoughnessTexture") : LB_NO_HANDLE;
            const Value& sp = ext["KHR_materials_specular"];
            d.specular_factor = sp.is_object() ? (float)sp["specularFactor"].number(0.0) : 0.f;
            d.specular_tint_factor = sp.is_object() ? 1.f : 0.f;
            d.tint_texture = sp.is_object() ? extension_texture(sp, "specularColorTexture") : LB_NO_HANDLE;
            role(d.transmission_texture, T_TRANSMISSIVE); role(d.clear_coat_roughness_texture, T_CLEAR_COAT_ROUGHNESS); role(d.clear_coat_texture, T_CLEAR_COAT); role(d.tint_texture, T_TINT);
            g.materials.push_back(d);
        }
        floor_roughness(g);
        g.info.materials = (uint32_t)g.materials.size();
    }

    // GenerateTangentBinary, :734-900
    static void generate_tangents(Primitive& p) {
        const size_t nv = p.pos.size() / 3;
        p.tan.assign(nv * 4, 0.f);
        const Vec3* pos = reinterpret_cast<const Vec3*>(p.pos.data()); const Vec3* nrm = reinterpret_cast<const Vec3*>(p.nrm.data());
        const bool have_uv = !p.uv.empty();
        const float default_uv[3][2] = {{1.f, 1.f}, {0.f, 1.f}, {1.f, 0.f}};
        for (size_t t = 0; t + 2 < p.idx.size(); t += 3) {
            const uint32_t ix[3] = {p.idx[t], p.idx[t + 1], p.idx[t + 2]};
            float uv[3][2];
            for (int k = 0; k < 3; ++k) { uv[k][0] = have_uv ? p.uv[2 * ix[k]] : default_uv[k][0]; uv[k][1] = have_uv ? p.uv[2 * ix[k] + 1] : default_uv[k][1]; }
            auto len2 = [](const float* a, const float* b) { const float dx = a[0] - b[0], dy = a[1] - b[1]; return sqrtf(dx * dx + dy * dy); };
            const float eps = 1.1920928955078125e-07f;
            if (len2(uv[0], uv[1]) < eps || len2(uv[0], uv[2]) < eps || len2(uv[2], uv[1]) < eps) memcpy(uv, default_uv, sizeof uv);
            const Vec3 dp1 = pos[ix[1]] - pos[ix[0]], dp2 = pos[ix[2]] - pos[ix[0]];
            float du1[2] = {uv[1][0] - uv[0][0], uv[1][1] - uv[0][1]}, du2[2] = {uv[2][0] - uv[0][0], uv[2][1] - uv[0][1]};
            const float cross = du1[0] * du2[1] - du1[1] * du2[0];
            if (cross == 0.f) { du1[0] = default_uv[1][0] - default_uv[0][0]; du1[1] = default_uv[1][1] - default_uv[0][1]; du2[0] = default_uv[2][0] - default_uv[0][0]; du2[1] = default_uv[2][1] - default_uv[0][1]; }
            const Vec3 tg = (du2[1] * dp1 - du1[1] * dp2) / (du1[0] * du2[1] - du2[0] * du1[1]);
            for (int k = 0; k < 3; ++k) {
                const Vec3 ng = normalize3(nrm[ix[k]]);
                const Vec3 tangent = normalize3(tg - ng * dot3(ng, tg));
                float* o = &p.tan[4 * (size_t)ix[k]]; o[0] = tangent.x; o[1] = tangent.y; o[2] = tangent.z; o[3] = 1.f;
            }
        }
    }
    static void flat_normals(Primitive& p) {
        const size_t nv = p.pos.size() / 3;
        p.nrm.assign(nv * 3, 0.f);
        const Vec3* pos = reinterpret_cast<const Vec3*>(p.pos.data()); Vec3* nrm = reinterpret_cast<Vec3*>(p.nrm.data());
        for (size_t t = 0; t + 2 < p.idx.size(); t += 3) {
            const Vec3 n = cross3(pos[p.idx[t + 1]] - pos[p.idx[t]], pos[p.idx[t + 2]] - pos[p.idx[t]]);
            for (int k = 0; k < 3; ++k) { Vec3& d = nrm[p.idx[t + k]]; d = {d.x + n.x, d.y + n.y, d.z + n.z}; }
        }
        for (size_t v = 0; v < nv; ++v) { const float l = sqrtf(dot3(nrm[v], nrm[v])); nrm[v] = l > 0.f ? nrm[v] * (1.0f / l) : Vec3{0.f, 0.f, 1.f}; }
    }
    void load_meshes() {
        const Value& meshes = doc["meshes"];
        for (size_t mi = 0; mi < meshes.size(); ++mi) {
            Mesh mesh;
            const Value& prims = meshes[mi]["primitives"];
            for (size_t pi = 0; pi < prims.size(); ++pi) {
                const Value& fp = prims[pi]; const Value& at = fp["attributes"];
                if (fp["mode"].integer(4) != 4) bad("only triangle lists (mode 4) are supported");
                if (!at.has("POSITION")) bad("primitive without POSITION");
                Primitive p;
                p.pos = float_attribute(at["POSITION"].integer(-1), 3, "POSITION");
                const size_t nv = p.pos.size() / 3;
                if (at.has("TEXCOORD_0")) p.uv = float_attribute(at["TEXCOORD_0"].integer(-1), 2, "TEXCOORD_0");
                if (at.has("NORMAL")) p.nrm = float_attribute(at["NORMAL"].integer(-1), 3, "NORMAL");
                if (at.has("TANGENT")) p.tan = float_attribute(at["TANGENT"].integer(-1), 4, "TANGENT");
                if ((!p.uv.empty() && p.uv.size() != nv * 2) || (!p.nrm.empty() && p.nrm.size() != nv * 3) || (!p.tan.empty() && p.tan.size() != nv * 4)) bad("attribute counts differ within a primitive");
                if (fp.has("indices")) {
                    uint32_t cs, cc; int64_t ct;
                    const std::vector<uint8_t> raw = accessor_bytes(fp["indices"].integer(-1), cs, cc, ct);
                    if (cc != 1 || ct == 5126) bad("indices must be scalar integers");
                    p.idx.resize(raw.size() / cs); p.index_size = cs;
                    for (size_t k = 0; k < p.idx.size(); ++k) p.idx[k] = cs == 1 ? raw[k] : cs == 2 ? (uint32_t)(raw[2 * k] | (raw[2 * k + 1] << 8)) : (uint32_t)(raw[4 * k] | (raw[4 * k + 1] << 8) | (raw[4 * k + 2] << 16) | ((uint32_t)raw[4 * k + 3] << 24));
                } else { p.idx.resize(nv); for (size_t k = 0; k < nv; ++k) p.idx[k] = (uint32_t)k; }
                p.idx.resize(p.idx.size() / 3 * 3);
                for (uint32_t v : p.idx) if (v >= nv) bad("index out of range");
                if (p.nrm.empty()) flat_normals(p);
                if (p.tan.empty()) generate_tangents(p);
                if (p.uv.empty()) p.uv.assign(nv * 2, 0.f);                    // InterleaveVertexBuffers leaves a missing stream zero (:902-928)
                p.material = (int32_t)fp["material"].integer(-1);
                if (p.material >= (int32_t)g.materials.size()) bad("material index out of range");
                g.info.triangles += (uint32_t)(p.idx.size() / 3); g.info.vertices += (uint32_t)nv; ++g.info.primitives;
                mesh.prims.push_back(std::move(p));
            }
            g.meshes.push_back(std::move(mesh));
        }
        g.info.meshes = (uint32_t)g.meshes.size();
    }

    Mat4 node_local(const Value& n) const {
        Mat4 m = mat_identity();
        if (n["matrix"].size() == 16) for (int k = 0; k < 16; ++k) m.m[k] = (float)n["matrix"][(size_t)k].number(0.0);
        if (!mat_is_identity(m)) return m;                                     // LoadNodeTransform :994-1025
        float t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
        for (int k = 0; k < 3; ++k) { t[k] = (float)n["translation"][(size_t)k].number(0.0); s[k] = (float)n["scale"][(size_t)k].number(1.0); }
        for (int k = 0; k < 4; ++k) q[k] = (float)n["rotation"][(size_t)k].number(k == 3 ? 1.0 : 0.0);
        return mat_trs(t, q, s);
    }
    // glTF requires a strict tree: a node reached twice (a DAG expands exponentially: 26 nodes listing each other twice are 2^26 copies)
    // or a cycle is rejected
    std::vector<uint8_t> node_seen;
    Node load_node(int64_t index, int depth) {
        const Value& n = doc["nodes"][(size_t)index];
        if (index < 0 || n.is_null()) bad("node index out of range");
        if (depth > 512) bad("node hierarchy too deep (cycle?)");
        if (node_seen.size() < doc["nodes"].size()) node_seen.resize(doc["nodes"].size(), 0);
        if (node_seen[(size_t)index]) bad("node referenced more than once (glTF node hierarchies are strict trees)");
        node_seen[(size_t)index] = 1;
        Node out; out.name = n["name"].is_null() ? std::string() : n["name"].string(); out.local = node_local(n);
        const int64_t mesh = n["mesh"].integer(-1);
        if (mesh >= (int64_t)g.meshes.size()) bad("mesh index out of range");
        out.mesh = mesh >= 0 ? (int32_t)mesh : -1;
        const Value& ch = n["children"];
        for (size_t k = 0; k < ch.size(); ++k) out.children.push_back(load_node(ch[k].integer(-1), depth + 1));
        return out;
    }
    void load_scenes() {
        const Value& scenes = doc["scenes"];
        for (size_t s = 0; s < scenes.size(); ++s) {
            Scene sc; sc.name = scenes[s]["name"].is_null() ? std::string() : scenes[s]["name"].string();
            const Value& roots = scenes[s]["nodes"];
            node_seen.assign(doc["nodes"].size(), 0);          // per scene: two scenes may share nodes, one scene may not visit a node twice
            for (size_t k = 0; k < roots.size(); ++k) sc.roots.push_back(load_node(roots[k].integer(-1), 0));
            g.scenes.push_back(std::move(sc));
        }
        flatten_scenes(g);
    }
};


// ---- the `.ollad` cache (LumenPTModelConverter: GenerateHeader :533-576, OutputToFile :588-598, LoadFile :70-273) ----
// file = u64 headerSize | header | blob.   header = u64 nTex, HeaderTexture[] | u64 nMat, HeaderMaterial[] | u64 nMesh, per mesh { u32 nPrim,
// HeaderPrimitive[] } | u64 nScenes, per scene { u32 nRoots, u32 nameLength, name, nodes depth first: { u32 nameLength, u32 nChildren,
// f32 local[16] (column-major), i32 mesh, name } }.   blob = the encoded image files, then per primitive the interleaved 64-byte Vertex
// records (ModelStructs.h:21-28: position @0, uv @16, normal @24, tangent @48, padding zero) and the indices at their source width.
// Offsets in the header are relative to the blob.
struct OlladTexture { uint64_t offset, size, type; };
struct OlladMaterial {
    float color[4], emission[3];
    int32_t diffuse, normal, metal_rough, emissive, transmission, clear_coat, clear_coat_roughness, tint;
    float transmission_factor, clear_coat_factor, clear_coat_roughness_factor, ior, specular, specular_tint, subsurface, luminance, anisotropic, sheen, sheen_tint, metallic, roughness;
    float tint_factor[3], transmittance[3];
};
struct OlladPrimitive { uint64_t vb_offset, vb_size, ib_offset, ib_size; uint32_t index_size, material; };
struct OlladNode { uint32_t name_length, num_children; float local[16]; int32_t mesh; };
struct OlladScene { uint32_t num_roots, name_length; };
static_assert(sizeof(OlladTexture) == 24 && sizeof(OlladMaterial) == 136 && sizeof(OlladPrimitive) == 40 && sizeof(OlladNode) == 76 && sizeof(OlladScene) == 8, "ollad records are packed");

void put(std::vector<uint8_t>& o, const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); o.insert(o.end(), b, b + n); }
void put_node(std::vector<uint8_t>& o, const Node& n) {
    OlladNode h{}; h.name_length = (uint32_t)n.name.size(); h.num_children = (uint32_t)n.children.size(); memcpy(h.local, n.local.m, 64); h.mesh = n.mesh;
    put(o, &h, sizeof h); put(o, n.name.data(), n.name.size());
    for (const Node& c : n.children) put_node(o, c);
}
OlladMaterial to_ollad(const LbMaterialDesc& d) {
    OlladMaterial m{};
    memcpy(m.color, d.diffuse_color, 16); memcpy(m.emission, d.emission, 12);
    m.diffuse = d.diffuse_texture; m.normal = d.normal_texture; m.metal_rough = d.metallic_roughness_texture; m.emissive = d.emissive_texture;
    m.transmission = d.transmission_texture; m.clear_coat = d.clear_coat_texture; m.clear_coat_roughness = d.clear_coat_roughness_texture; m.tint = d.tint_texture;
    m.transmission_factor = d.transmission_factor; m.clear_coat_factor = d.clear_coat_factor; m.clear_coat_roughness_factor = d.clear_coat_roughness_factor;
    m.ior = d.index_of_refraction; m.specular = d.specular_factor; m.specular_tint = d.specular_tint_factor; m.subsurface = d.subsurface_factor; m.luminance = d.luminance;
    m.anisotropic = d.anisotropic; m.sheen = d.sheen_factor; m.sheen_tint = d.sheen_tint_factor; m.metallic = d.metallic_factor; m.roughness = d.roughness_factor;
    memcpy(m.tint_factor, d.tint_factor, 12); memcpy(m.transmittance, d.transmittance, 12);
    return m;
}
LbMaterialDesc from_ollad(const OlladMaterial& m) {
    LbMaterialDesc d{};
    memcpy(d.diffuse_color, m.color, 16); memcpy(d.emission, m.emission, 12);
    d.diffuse_texture = m.diffuse; d.normal_texture = m.normal; d.metallic_roughness_texture = m.metal_rough; d.emissive_texture = m.emissive;
    d.transmission_texture = m.transmission; d.clear_coat_texture = m.clear_coat; d.clear_coat_roughness_texture = m.clear_coat_roughness; d.tint_texture = m.tint;
    d.transmission_factor = m.transmission_factor; d.clear_coat_factor = m.clear_coat_factor; d.clear_coat_roughness_factor = m.clear_coat_roughness_factor;
    d.index_of_refraction = m.ior; d.specular_factor = m.specular; d.specular_tint_factor = m.specular_tint; d.subsurface_factor = m.subsurface; d.luminance = m.luminance;
    d.anisotropic = m.anisotropic; d.sheen_factor = m.sheen; d.sheen_tint_factor = m.sheen_tint; d.metallic_factor = m.metallic; d.roughness_factor = m.roughness;
    memcpy(d.tint_factor, m.tint_factor, 12); memcpy(d.transmittance, m.transmittance, 12);
    return d;
}

struct OlladReader {
    LbGltfOpaque& g; const std::vector<uint8_t>& file; size_t at = 8, header_end = 0; const uint8_t* blob = nullptr; size_t blob_size = 0;
    [[noreturn]] void bad(const std::string& what) const { throw std::runtime_error("ollad: " + what); }
    void get(void* out, size_t n) { if (n > header_end - at) bad("header record runs past the end of the header"); memcpy(out, file.data() + at, n); at += n; }
    uint64_t count(size_t record) { uint64_t n; get(&n, 8); if (n > (header_end - at) / (record ? record : 1)) bad("implausible record count"); return n; }
    const uint8_t* span(uint64_t off, uint64_t size) const { if (off > blob_size || size > blob_size - off) bad("payload range outside the file"); return blob + off; }
    std::string name(uint32_t n) { if (n > header_end - at) bad("name runs past the end of the header"); std::string s(reinterpret_cast<const char*>(file.data() + at), n); at += n; return s; }
    Node node(int depth) {
        if (depth > 512) bad("node hierarchy too deep");
        OlladNode h; get(&h, sizeof h);
        Node n; n.name = name(h.name_length); memcpy(n.local.m, h.local, 64); n.mesh = h.mesh;
        if (n.mesh < -1 || n.mesh >= (int64_t)g.meshes.size()) bad("mesh index out of range");
        if (h.num_children > (header_end - at) / sizeof(OlladNode)) bad("implausible child count");
        for (uint32_t k = 0; k < h.num_children; ++k) n.children.push_back(node(depth + 1));
        return n;
    }
    void read(LbImageDecodeFn decoder, void* user) {
        if (file.size() < 8) bad("file shorter than its size field");
        uint64_t header_size; memcpy(&header_size, file.data(), 8);
        if (header_size > file.size() - 8) bad("header size exceeds the file");
        header_end = 8 + (size_t)header_size; blob = file.data() + header_end; blob_size = file.size() - header_end;
        const uint64_t ntex = count(sizeof(OlladTexture));
        g.images.resize((size_t)ntex);
        for (Image& im : g.images) {
            OlladTexture t; get(&t, sizeof t);
            const uint8_t* p = span(t.offset, t.size);
            im.raw.assign(p, p + t.size); im.type = t.type;
            decode_image(g, im, decoder, user);
        }
        floor_roughness(g);
        g.info.images = (uint32_t)g.images.size();
        const uint64_t nmat = count(sizeof(OlladMaterial));
        for (uint64_t i = 0; i < nmat; ++i) {
            OlladMaterial m; get(&m, sizeof m);
            LbMaterialDesc d = from_ollad(m);
            for (LbHandle* h : {&d.diffuse_texture, &d.normal_texture, &d.metallic_roughness_texture, &d.emissive_texture, &d.transmission_texture, &d.clear_coat_texture, &d.clear_coat_roughness_texture, &d.tint_texture})
                if (*h < -1 || *h >= (int64_t)g.images.size()) bad("texture index out of range");
            g.materials.push_back(d);
        }
        g.info.materials = (uint32_t)g.materials.size();
        const uint64_t nmesh = count(4);
        for (uint64_t mi = 0; mi < nmesh; ++mi) {
            uint32_t nprim; get(&nprim, 4);
            if (nprim > (header_end - at) / sizeof(OlladPrimitive)) bad("implausible primitive count");
            Mesh mesh;
            for (uint32_t pi = 0; pi < nprim; ++pi) {
                OlladPrimitive h; get(&h, sizeof h);
                if (h.vb_size % 64 || (h.index_size != 1 && h.index_size != 2 && h.index_size != 4) || h.ib_size % h.index_size) bad("primitive with a malformed vertex or index buffer");
                const uint8_t* vb = span(h.vb_offset, h.vb_size); const uint8_t* ib = span(h.ib_offset, h.ib_size);
                Primitive p; const size_t nv = (size_t)(h.vb_size / 64), ni = (size_t)(h.ib_size / h.index_size) / 3 * 3;
                if (nv > ((size_t)1 << 31) || ni > ((size_t)1 << 32)) bad("primitive too large");
                p.pos.resize(nv * 3); p.uv.resize(nv * 2); p.nrm.resize(nv * 3); p.tan.resize(nv * 4); p.idx.resize(ni); p.index_size = h.index_size;
                for (size_t v = 0; v < nv; ++v) {
                    const uint8_t* r = vb + v * 64;
                    memcpy(&p.pos[v * 3], r, 12); memcpy(&p.uv[v * 2], r + 16, 8); memcpy(&p.nrm[v * 3], r + 24, 12); memcpy(&p.tan[v * 4], r + 48, 16);
                }
                for (size_t k = 0; k < ni; ++k) {
                    const uint8_t* r = ib + k * h.index_size;
                    p.idx[k] = h.index_size == 1 ? r[0] : h.index_size == 2 ? (uint32_t)(r[0] | (r[1] << 8)) : (uint32_t)(r[0] | (r[1] << 8) | (r[2] << 16) | ((uint32_t)r[3] << 24));
                    if (p.idx[k] >= nv) bad("index out of range");
                }
                p.material = (int32_t)h.material;                                  // 0xFFFFFFFF = the glTF default material
                if (p.material < -1 || p.material >= (int32_t)g.materials.size()) bad("material index out of range");
                g.info.triangles += (uint32_t)(ni / 3); g.info.vertices += (uint32_t)nv; ++g.info.primitives;
                mesh.prims.push_back(std::move(p));
            }
            g.meshes.push_back(std::move(mesh));
        }
        g.info.meshes = (uint32_t)g.meshes.size();
        const uint64_t nscenes = count(sizeof(OlladScene));
        for (uint64_t si = 0; si < nscenes; ++si) {
            OlladScene h; get(&h, sizeof h);
            Scene sc; sc.name = name(h.name_length);
            if (h.num_roots > (header_end - at) / sizeof(OlladNode)) bad("implausible root count");
            for (uint32_t k = 0; k < h.num_roots; ++k) sc.roots.push_back(node(0));
            g.scenes.push_back(std::move(sc));
        }
        flatten_scenes(g);
    }
};
bool has_extension(const std::string& p, const char* ext) { const size_t n = strlen(ext); return p.size() >= n && p.compare(p.size() - n, n, ext) == 0; }

} // namespace

extern "C" {

LB_API const char* lb_gltf_last_error(void) { return g_error.c_str(); }

LB_API int lb_gltf_open(const char* path, LbImageDecodeFn decoder, void* user, LbGltf* out) {
    if (!path || !out) return gfail(LB_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    try {
        std::vector<uint8_t> file;
        if (!read_file(path, file)) return gfail(LB_ERR_INVALID_ARGUMENT, std::string("cannot read ") + path);
        const std::string p(path); const size_t slash = p.find_last_of("/\\");
        if (has_extension(p, ".ollad")) {                                            // LumenPTModelConverter::ms_ExtensionName: the cache instead of the source
            std::unique_ptr<LbGltfOpaque> g(new LbGltfOpaque());
            OlladReader R{*g, file};
            R.read(decoder, user);
            *out = g.release();
            return LB_OK;
        }
        std::vector<uint8_t> bin; const char* text = reinterpret_cast<const char*>(file.data()); size_t text_len = file.size();
        if (file.size() >= 12 && memcmp(file.data(), "glTF", 4) == 0) {             // GLB container: header, JSON chunk, optional BIN chunk
            auto le = [&](size_t at) { return (uint32_t)file[at] | ((uint32_t)file[at + 1] << 8) | ((uint32_t)file[at + 2] << 16) | ((uint32_t)file[at + 3] << 24); };
            size_t pos = 12; text = nullptr;
            while (pos + 8 <= file.size()) {
                const uint32_t len = le(pos), type = le(pos + 4);
                if (pos + 8 + (size_t)len > file.size()) return gfail(LB_ERR_INVALID_ARGUMENT, "GLB chunk runs past the end of the file");
                if (type == 0x4E4F534Au && !text) { text = reinterpret_cast<const char*>(file.data() + pos + 8); text_len = len; }
                else if (type == 0x004E4942u && bin.empty()) bin.assign(file.begin() + pos + 8, file.begin() + pos + 8 + len);
                pos += 8 + (size_t)len;
            }
            if (!text) return gfail(LB_ERR_INVALID_ARGUMENT, "GLB without a JSON chunk");
        }
        const Value doc = lb::json::parse(text, text_len);
        if (!doc.is_object() || !doc.has("asset")) return gfail(LB_ERR_INVALID_ARGUMENT, "not a glTF document");
        std::unique_ptr<LbGltfOpaque> g(new LbGltfOpaque());
        Loader L{*g, doc, slash == std::string::npos ? std::string() : p.substr(0, slash + 1), {}, std::move(bin), decoder, user};
        L.load_buffers(); L.load_images(); L.load_materials(); L.load_meshes(); L.load_scenes();
        *out = g.release();
        return LB_OK;
    } catch (const std::exception& e) { return gfail(LB_ERR_INVALID_ARGUMENT, e.what()); }
}
LB_API int lb_gltf_save_ollad(LbGltf g, const char* path) {
    if (!g || !path) return gfail(LB_ERR_INVALID_ARGUMENT, "null argument");
    try {
        std::vector<uint8_t> header, blob;
        const uint64_t ntex = g->images.size(), nmat = g->materials.size(), nmesh = g->meshes.size(), nscenes = g->scenes.size();
        put(header, &ntex, 8);
        for (const Image& im : g->images) { const OlladTexture t{blob.size(), im.raw.size(), im.type}; put(header, &t, sizeof t); put(blob, im.raw.data(), im.raw.size()); }
        put(header, &nmat, 8);
        for (const LbMaterialDesc& d : g->materials) { const OlladMaterial m = to_ollad(d); put(header, &m, sizeof m); }
        put(header, &nmesh, 8);
        for (const Mesh& mesh : g->meshes) {
            const uint32_t nprim = (uint32_t)mesh.prims.size(); put(header, &nprim, 4);
            for (const Primitive& p : mesh.prims) {
                const size_t nv = p.pos.size() / 3;
                OlladPrimitive h{}; h.vb_offset = blob.size(); h.vb_size = nv * 64;
                blob.resize(blob.size() + nv * 64, 0);
                for (size_t v = 0; v < nv; ++v) {
                    uint8_t* r = blob.data() + h.vb_offset + v * 64;
                    memcpy(r, &p.pos[v * 3], 12); memcpy(r + 16, &p.uv[v * 2], 8); memcpy(r + 24, &p.nrm[v * 3], 12); memcpy(r + 48, &p.tan[v * 4], 16);
                }
                h.ib_offset = blob.size(); h.index_size = p.index_size; h.ib_size = p.idx.size() * p.index_size;
                for (uint32_t i : p.idx) put(blob, &i, p.index_size);          // little endian: the low bytes
                h.material = (uint32_t)p.material;
                put(header, &h, sizeof h);
            }
        }
        put(header, &nscenes, 8);
        for (const Scene& sc : g->scenes) {
            const OlladScene h{(uint32_t)sc.roots.size(), (uint32_t)sc.name.size()};
            put(header, &h, sizeof h); put(header, sc.name.data(), sc.name.size());
            for (const Node& n : sc.roots) put_node(header, n);
        }
        // Written to a temporary file in the same directory and renamed over the target: every rank of a multi-GPU job converts the same
        // asset, and a reader must never see a cache whose header is complete while its blob is still being written (such a file passes
        // every span check and loads zeroed geometry). rename() within one directory is atomic.
        const std::string tmp = std::string(path) + ".tmp." + std::to_string((unsigned long long)getpid()) + "." + std::to_string((unsigned long long)(uintptr_t)g);
        FILE* f = fopen(tmp.c_str(), "wb");
        if (!f) return gfail(LB_ERR_INVALID_ARGUMENT, std::string("cannot write ") + path);
        const uint64_t header_size = header.size();
        const bool ok = fwrite(&header_size, 8, 1, f) == 1 && (header.empty() || fwrite(header.data(), 1, header.size(), f) == header.size()) && (blob.empty() || fwrite(blob.data(), 1, blob.size(), f) == blob.size());
        if (fclose(f) != 0 || !ok) { remove(tmp.c_str()); return gfail(LB_ERR_INVALID_ARGUMENT, std::string("short write to ") + path); }
        if (rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return gfail(LB_ERR_INVALID_ARGUMENT, std::string("cannot replace ") + path); }
        return LB_OK;
    } catch (const std::exception& e) { return gfail(LB_ERR_INVALID_ARGUMENT, e.what()); }
}
// SceneManager::LoadGLTF's first two steps (LM/ModelLoading/SceneManager.cpp:55-76) with WaveFrontRenderer's hooks behind them
// (WaveFrontRenderer.cpp:1135-1146): OpenCustomFileFormat = LoadFile(<path with the extension replaced by .ollad>), else
// CreateCustomFileFormat = ConvertGLTF (convert, write the cache next to the source, :27-68).
LB_API int lb_gltf_open_cached(const char* path, LbImageDecodeFn decoder, void* user, LbGltf* out) {
    if (!path || !out) return gfail(LB_ERR_INVALID_ARGUMENT, "null argument");
    std::string cache(path);
    const size_t slash = cache.find_last_of("/\\"), dot = cache.find_last_of('.');
    if (dot != std::string::npos && (slash == std::string::npos || dot > slash + 1)) cache.erase(dot);      // std::filesystem::path::replace_extension
    cache += ".ollad";
    if (FILE* f = fopen(cache.c_str(), "rb")) {
        fclose(f);
        if (lb_gltf_open(cache.c_str(), decoder, user, out) == LB_OK) return LB_OK;      // an unreadable cache is rebuilt (the reference would crash on it)
    }
    const int rc = lb_gltf_open(path, decoder, user, out);
    if (rc == LB_OK && !has_extension(std::string(path), ".ollad")) lb_gltf_save_ollad(*out, cache.c_str());   // a read-only asset directory is not an error
    return rc;
}
LB_API int lb_gltf_close(LbGltf g) { delete g; return LB_OK; }
LB_API int lb_gltf_info(LbGltf g, LbGltfInfo* out) { if (!g || !out) return gfail(LB_ERR_INVALID_ARGUMENT, "null argument"); *out = g->info; return LB_OK; }
LB_API int lb_gltf_image(LbGltf g, uint32_t i, const uint8_t** rgba8, uint32_t* w, uint32_t* h, int* srgb, int* decoded) {
    if (!g || i >= g->images.size()) return gfail(LB_ERR_INVALID_HANDLE, "image");
    const Image& im = g->images[i];
    if (rgba8) *rgba8 = im.px.data(); if (w) *w = im.w; if (h) *h = im.h; if (srgb) *srgb = im.srgb() ? 1 : 0; if (decoded) *decoded = im.decoded ? 1 : 0;
    return LB_OK;
}
LB_API int lb_gltf_material(LbGltf g, uint32_t i, LbMaterialDesc* out) {
    if (!g || !out || i >= g->materials.size()) return gfail(LB_ERR_INVALID_HANDLE, "material");
    *out = g->materials[i]; return LB_OK;
}
LB_API int lb_gltf_mesh_primitive_count(LbGltf g, uint32_t mesh, uint32_t* count) {
    if (!g || !count || mesh >= g->meshes.size()) return gfail(LB_ERR_INVALID_HANDLE, "mesh");
    *count = (uint32_t)g->meshes[mesh].prims.size(); return LB_OK;
}
LB_API int lb_gltf_primitive(LbGltf g, uint32_t mesh, uint32_t prim, LbPrimitiveDesc* out) {
    if (!g || !out || mesh >= g->meshes.size() || prim >= g->meshes[mesh].prims.size()) return gfail(LB_ERR_INVALID_HANDLE, "primitive");
    const Primitive& p = g->meshes[mesh].prims[prim];
    LbPrimitiveDesc d{};
    d.positions = p.pos.data(); d.position_stride = 12; d.uvs = p.uv.data(); d.uv_stride = 8; d.normals = p.nrm.data(); d.normal_stride = 12;
    d.tangents = p.tan.data(); d.tangent_stride = 16; d.vertex_count = (uint32_t)(p.pos.size() / 3);
    d.indices = p.idx.data(); d.index_size = 4; d.index_count = (uint32_t)p.idx.size(); d.material = p.material;
    *out = d; return LB_OK;
}
LB_API int lb_gltf_instance(LbGltf g, uint32_t i, uint32_t* mesh, float* transform16) {
    if (!g || i >= g->instances.size()) return gfail(LB_ERR_INVALID_HANDLE, "instance");
    if (mesh) *mesh = g->instances[i].mesh; if (transform16) memcpy(transform16, g->instances[i].m, 64);
    return LB_OK;
}

// CreateTexture / CreateMaterial / CreatePrimitive / CreateMesh / AddMesh in the order LumenPTModelConverter::LoadFile issues them (:105-268)
LB_API int lb_gltf_upload(LbRenderer r, LbGltf g, const float* root16, LbHandle* first_instance, uint32_t* instance_count) {
    if (!r || !g) return gfail(LB_ERR_INVALID_ARGUMENT, "null argument");
    auto check = [&](int rc) { if (rc != LB_OK) { g_error = lb_last_error(); throw rc; } };
    try {
        std::vector<LbHandle> tex(g->images.size(), LB_NO_HANDLE), mats(g->materials.size(), LB_NO_HANDLE), meshes(g->meshes.size(), LB_NO_HANDLE);
        for (size_t i = 0; i < g->images.size(); ++i) { const Image& im = g->images[i]; if (im.decoded) check(lb_texture_create(r, im.px.data(), im.w, im.h, im.srgb() ? 1 : 0, &tex[i])); }
        auto map_tex = [&](LbHandle& h) { h = h >= 0 ? tex[(size_t)h] : LB_NO_HANDLE; };
        for (size_t i = 0; i < g->materials.size(); ++i) {
            LbMaterialDesc d = g->materials[i];
            map_tex(d.diffuse_texture); map_tex(d.normal_texture); map_tex(d.metallic_roughness_texture); map_tex(d.emissive_texture);
            map_tex(d.transmission_texture); map_tex(d.clear_coat_texture); map_tex(d.clear_coat_roughness_texture); map_tex(d.tint_texture);
            check(lb_material_create(r, &d, &mats[i]));
        }
        LbHandle default_material = LB_NO_HANDLE;
        for (size_t m = 0; m < g->meshes.size(); ++m) {
            std::vector<LbHandle> prims;
            for (uint32_t p = 0; p < g->meshes[m].prims.size(); ++p) {
                LbPrimitiveDesc d; lb_gltf_primitive(g, (uint32_t)m, p, &d);
                if (d.material >= 0) d.material = mats[(size_t)d.material];
                else {                                                           // glTF default material (the reference indexes the pool with -1)
                    if (default_material == LB_NO_HANDLE) {
                        LbMaterialDesc dm{}; dm.diffuse_color[0] = dm.diffuse_color[1] = dm.diffuse_color[2] = dm.diffuse_color[3] = 1.f; dm.metallic_factor = 1.f; dm.roughness_factor = 1.f; dm.luminance = 1.f; dm.index_of_refraction = 1.f;
                        dm.diffuse_texture = dm.normal_texture = dm.metallic_roughness_texture = dm.emissive_texture = dm.transmission_texture = dm.clear_coat_texture = dm.clear_coat_roughness_texture = dm.tint_texture = LB_NO_HANDLE;
                        check(lb_material_create(r, &dm, &default_material));
                    }
                    d.material = default_material;
                }
                LbHandle h; check(lb_primitive_create(r, &d, &h)); prims.push_back(h);
            }
            check(lb_mesh_create(r, prims.data(), (uint32_t)prims.size(), &meshes[m]));
        }
        Mat4 root = mat_identity();
        if (root16) for (int rr = 0; rr < 4; ++rr) for (int c = 0; c < 4; ++c) root.m[c * 4 + rr] = root16[rr * 4 + c];
        LbHandle first = LB_NO_HANDLE;
        for (const Instance& in : g->instances) {
            float m16[16]; memcpy(m16, in.m, 64);
            if (root16) { Mat4 w; for (int rr = 0; rr < 4; ++rr) for (int c = 0; c < 4; ++c) w.m[c * 4 + rr] = in.m[rr * 4 + c]; const Mat4 t = mat_mul(root, w); for (int rr = 0; rr < 4; ++rr) for (int c = 0; c < 4; ++c) m16[rr * 4 + c] = t.m[c * 4 + rr]; }
            LbHandle h; check(lb_scene_add_mesh_instance(r, meshes[in.mesh], m16, nullptr, LB_NO_HANDLE, &h));
            if (first == LB_NO_HANDLE) first = h;
        }
        if (first_instance) *first_instance = first;
        if (instance_count) *instance_count = (uint32_t)g->instances.size();
        return LB_OK;
    } catch (int rc) { return rc; }
}

} // extern "C"
