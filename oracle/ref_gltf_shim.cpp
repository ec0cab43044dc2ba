// oracle/ref_gltf_shim.cpp -- TEST INFRASTRUCTURE ONLY (the checker of SURVEY 8f F4, never on the product path).
//
// The engine's route from a .gltf / .glb file to the arrays Triangle::CreateTriangleList gets:
//   tinygltf (the reference's own pinned reader, compiled in place from /root/reference/tinygltf, images switched off)
//   -> MeshesOfNodes::AddMeshesOfModel's per-mesh primitive order: the SAME std::sort call on the same element type with the same
//      predicate (IMR/src/Graphics/Meshes/MeshesOfNodes.cpp:41-43), so the order is whatever this libstdc++ makes of it, as in the engine
//   -> PrimitiveInitializationData (IMR/src/Graphics/Meshes/PrimitivesOfMeshes.cpp:44-175) + GetPrimitiveOBBtreeData (:637-671).
// PrimitivesOfMeshes.cpp itself cannot be compiled here (it needs the Vulkan and VMA headers, which this image does not have), so the
// extraction of indices / POSITION / NORMAL is restated below, each step citing its line; everything after it (CreateTriangleList,
// the OBB tree) is the unmodified reference in libimr_ref.so.
#define TINYGLTF_IMPLEMENTATION
#define TINYGLTF_NO_STB_IMAGE
#define TINYGLTF_NO_STB_IMAGE_WRITE
#define TINYGLTF_NO_EXTERNAL_IMAGE
#include "tiny_gltf.h"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct RefPrimitive {
    std::vector<float> points, normals;      // n * 3
    std::vector<uint32_t> indices;
    bool has_normals = false, has_indices = false, skin_or_morph = false;
    uint32_t mode = 4, source_index = 0;
};
struct RefFile { tinygltf::Model model; std::vector<std::vector<RefPrimitive>> meshes; };

// GetAccessorBeginEndPtrs (PrimitivesOfMeshes.cpp:700-752): tightly packed, accessor.byteOffset + bufferView.byteOffset
const unsigned char* accessor_begin(const tinygltf::Model& m, const tinygltf::Accessor& a) {
    const tinygltf::BufferView& v = m.bufferViews[a.bufferView];
    return &m.buffers[v.buffer].data[a.byteOffset + v.byteOffset];
}

void vec3_flipped(const tinygltf::Model& m, int accessor, std::vector<float>& out) {     // :81-87, :133-139
    const tinygltf::Accessor& a = m.accessors[accessor];
    const unsigned char* p = accessor_begin(m, a);
    out.resize(a.count * 3);
    for (size_t i = 0; i < a.count; ++i) {
        float in[3];
        std::memcpy(in, p + 12 * i, 12);
        out[3 * i] = in[0]; out[3 * i + 1] = -in[1]; out[3 * i + 2] = -in[2];
    }
}

RefPrimitive extract(const tinygltf::Model& m, const tinygltf::Primitive& pr) {
    RefPrimitive r;
    if (pr.mode != -1) r.mode = pr.mode == TINYGLTF_MODE_LINE_LOOP ? TINYGLTF_MODE_LINE_STRIP : (uint32_t)pr.mode;         // :49-55
    const bool morphed = !pr.targets.empty() && pr.targets[0].find("POSITION") != pr.targets[0].end();                    // :94, :108
    const bool skinned = pr.attributes.find("JOINTS_0") != pr.attributes.end();                                             // jointsCount, :641
    r.skin_or_morph = morphed || skinned;
    if (r.skin_or_morph) return r;
    if (pr.indices != -1) {                                                                                                 // :58-70
        const tinygltf::Accessor& a = m.accessors[pr.indices];
        const unsigned char* p = accessor_begin(m, a);
        r.has_indices = true; r.indices.resize(a.count);
        for (size_t i = 0; i < a.count; ++i) {
            if (a.componentType == TINYGLTF_COMPONENT_TYPE_UNSIGNED_SHORT) { uint16_t x; std::memcpy(&x, p + 2 * i, 2); r.indices[i] = x; }
            else if (a.componentType == TINYGLTF_COMPONENT_TYPE_UNSIGNED_INT) { uint32_t x; std::memcpy(&x, p + 4 * i, 4); r.indices[i] = x; }
            else { r.indices[i] = p[i]; }                                   // u8: the reference asserts; kept so that the product's superset can be checked
        }
    }
    vec3_flipped(m, pr.attributes.at("POSITION"), r.points);
    auto n = pr.attributes.find("NORMAL");
    if (n != pr.attributes.end()) { r.has_normals = true; vec3_flipped(m, n->second, r.normals); }
    return r;
}

}  // namespace

extern "C" {

struct imr_refgltf_view {
    const float* points; const float* normals; const uint32_t* indices;
    uint64_t n_points, n_indices;
    uint32_t mode, skipped, source_index, has_indices;
};

void* imr_refgltf_open(const char* path, char* err, uint64_t cap) {
    RefFile* f = new RefFile;
    tinygltf::TinyGLTF loader;
    std::string e, w;
    const std::string p(path);
    const bool glb = p.size() > 4 && p.compare(p.size() - 4, 4, ".glb") == 0;
    const bool ok = glb ? loader.LoadBinaryFromFile(&f->model, &e, &w, p) : loader.LoadASCIIFromFile(&f->model, &e, &w, p);
    if (!ok) {
        if (err && cap) { std::strncpy(err, e.c_str(), cap - 1); err[cap - 1] = 0; }
        delete f; return nullptr;
    }
    for (const tinygltf::Mesh& mesh : f->model.meshes) {
        std::vector<tinygltf::Primitive> primitives = mesh.primitives;
        for (size_t i = 0; i < primitives.size(); ++i) primitives[i].extras_json_string = std::to_string(i);       // remember where it came from
        // Triangles first: MeshesOfNodes.cpp:41-43, the same call (glTFmode::triangles == 4)
        std::sort(primitives.begin(), primitives.end(),
                  [](const tinygltf::Primitive& lhs, const tinygltf::Primitive& rhs) { return lhs.mode == 4 || lhs.mode == -1; });
        f->meshes.emplace_back();
        for (const tinygltf::Primitive& pr : primitives) {
            f->meshes.back().push_back(extract(f->model, pr));
            f->meshes.back().back().source_index = (uint32_t)std::stoul(pr.extras_json_string);
        }
    }
    return f;
}
void imr_refgltf_close(void* h) { delete static_cast<RefFile*>(h); }
uint32_t imr_refgltf_mesh_count(void* h) { return (uint32_t)static_cast<RefFile*>(h)->meshes.size(); }
uint32_t imr_refgltf_primitive_count(void* h, uint32_t mesh) { return (uint32_t)static_cast<RefFile*>(h)->meshes[mesh].size(); }
void imr_refgltf_primitive(void* h, uint32_t mesh, uint32_t k, imr_refgltf_view* out) {
    const RefPrimitive& p = static_cast<RefFile*>(h)->meshes[mesh][k];
    out->points = p.points.data(); out->normals = p.has_normals ? p.normals.data() : nullptr; out->indices = p.has_indices ? p.indices.data() : nullptr;
    out->n_points = p.points.size() / 3; out->n_indices = p.indices.size();
    out->mode = p.mode; out->skipped = p.skin_or_morph; out->source_index = p.source_index; out->has_indices = p.has_indices;
}

}  // extern "C"
