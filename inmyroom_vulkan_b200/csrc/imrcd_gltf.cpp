// imrcd_gltf.cpp -- SURVEY 8f F4: .gltf / .glb / .bin -> one OBB tree per glTF mesh, without the engine's extractor in the loop.
//
// What the engine does at load time (and what this file replaces, host side only -- the triangles and the tree are made on the device):
//   * tinygltf reads the file (the engine's pinned dependency, /tinygltf); here a small JSON reader + base64 + the GLB container.
//   * per glTF mesh: StartRecordOBBtree, the mesh's primitives in the order MeshesOfNodes.cpp:41-43 leaves them ("triangles first"),
//     AddPrimitive -> PrimitiveInitializationData (IMR/src/Graphics/Meshes/PrimitivesOfMeshes.cpp:44-175): draw mode (line loop falls back
//     to line strip, :49-55), u16 / u32 indices (:58-70), float POSITION and NORMAL with y and z negated (:83-87, :134-139); skinned or
//     morphed primitives are left out of the tree (:641, :669), GetOBBtreeAndReset (:840-863).
//   * accessors are read as the reference reads them (:700-752): accessor.byteOffset + bufferView.byteOffset, `count` elements.  The
//     reference ignores bufferView.byteStride (tightly packed only); a stride is honoured here, which is the same thing on every asset the
//     reference reads correctly.
// Everything below the extraction goes through the public ABI (imrcd_mesh_begin / add_primitive / end).
#include "../../include/imrcd.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <string>
#include <utility>
#include <vector>

void imr_ctx_set_error(imrcd_ctx* ctx, const char* msg);      // imrcd_api.cu

namespace {

// ---------------------------------------------------------------- JSON (RFC 8259 subset: everything glTF uses)
struct JVal {
    enum Type { NUL, BOOL, NUM, STR, ARR, OBJ } type = NUL;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;

    const JVal* get(const char* key) const {
        if (type != OBJ) return nullptr;
        for (const auto& kv : obj) if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool is_index() const { return type == NUM && num >= 0.0 && num == std::floor(num) && num < 4294967296.0; }
    int64_t index_or(const char* key, int64_t dflt) const {
        const JVal* v = get(key);
        return v && v->is_index() ? (int64_t)v->num : dflt;
    }
};

struct JParser {
    const char* p; const char* end; std::string err; int depth = 0;

    void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
    bool fail(const char* what) { if (err.empty()) err = std::string("json: ") + what; return false; }

    static void utf8(std::string& s, uint32_t c) {
        if (c < 0x80) s += (char)c;
        else if (c < 0x800) { s += (char)(0xC0 | (c >> 6)); s += (char)(0x80 | (c & 63)); }
        else if (c < 0x10000) { s += (char)(0xE0 | (c >> 12)); s += (char)(0x80 | ((c >> 6) & 63)); s += (char)(0x80 | (c & 63)); }
        else { s += (char)(0xF0 | (c >> 18)); s += (char)(0x80 | ((c >> 12) & 63)); s += (char)(0x80 | ((c >> 6) & 63)); s += (char)(0x80 | (c & 63)); }
    }
    bool hex4(uint32_t& out) {
        if (end - p < 4) return fail("short \\u escape");
        out = 0;
        for (int i = 0; i < 4; ++i) {
            const char c = *p++;
            out <<= 4;
            if (c >= '0' && c <= '9') out |= (uint32_t)(c - '0');
            else if (c >= 'a' && c <= 'f') out |= (uint32_t)(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') out |= (uint32_t)(c - 'A' + 10);
            else return fail("bad \\u escape");
        }
        return true;
    }
    bool string(std::string& out) {
        if (p >= end || *p != '"') return fail("expected string");
        ++p;
        while (p < end && *p != '"') {
            if (*p != '\\') { out += *p++; continue; }
            if (++p >= end) break;
            const char c = *p++;
            switch (c) {
                case '"': out += '"'; break;   case '\\': out += '\\'; break; case '/': out += '/'; break;
                case 'b': out += '\b'; break;  case 'f': out += '\f'; break;  case 'n': out += '\n'; break;
                case 'r': out += '\r'; break;  case 't': out += '\t'; break;
                case 'u': {
                    uint32_t c1 = 0;
                    if (!hex4(c1)) return false;
                    if (c1 >= 0xD800 && c1 < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                        p += 2; uint32_t c2 = 0;
                        if (!hex4(c2)) return false;
                        c1 = 0x10000 + ((c1 - 0xD800) << 10) + (c2 - 0xDC00);
                    }
                    utf8(out, c1);
                } break;
                default: return fail("bad escape");
            }
        }
        if (p >= end) return fail("unterminated string");
        ++p;
        return true;
    }
    bool value(JVal& v) {
        if (++depth > 256) return fail("nesting too deep");
        ws();
        if (p >= end) return fail("unexpected end");
        bool ok = true;
        if (*p == '{') {
            v.type = JVal::OBJ; ++p; ws();
            if (p < end && *p == '}') ++p;
            else for (;;) {
                ws();
                std::string key;
                if (!string(key)) { ok = false; break; }
                ws();
                if (p >= end || *p != ':') { ok = fail("expected ':'"); break; }
                ++p;
                v.obj.emplace_back(std::move(key), JVal());
                if (!value(v.obj.back().second)) { ok = false; break; }
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; break; }
                ok = fail("expected ',' or '}'"); break;
            }
        } else if (*p == '[') {
            v.type = JVal::ARR; ++p; ws();
            if (p < end && *p == ']') ++p;
            else for (;;) {
                v.arr.emplace_back();
                if (!value(v.arr.back())) { ok = false; break; }
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; break; }
                ok = fail("expected ',' or ']'"); break;
            }
        } else if (*p == '"') {
            v.type = JVal::STR; ok = string(v.str);
        } else if (end - p >= 4 && !std::memcmp(p, "true", 4))  { v.type = JVal::BOOL; v.b = true; p += 4; }
        else if (end - p >= 5 && !std::memcmp(p, "false", 5)) { v.type = JVal::BOOL; v.b = false; p += 5; }
        else if (end - p >= 4 && !std::memcmp(p, "null", 4))  { v.type = JVal::NUL; p += 4; }
        else {
            const char* q = p;
            if (q < end && *q == '-') ++q;
            while (q < end && ((*q >= '0' && *q <= '9') || *q == '.' || *q == 'e' || *q == 'E' || *q == '+' || *q == '-')) ++q;
            if (q == p) ok = fail("unexpected character");
            else {
                const std::string t(p, q);
                char* stop = nullptr;
                v.type = JVal::NUM; v.num = std::strtod(t.c_str(), &stop);
                if (!stop || *stop) ok = fail("bad number");
                p = q;
            }
        }
        --depth;
        return ok;
    }
};

// ---------------------------------------------------------------- bytes: files, base64 data URIs, percent-encoded relative URIs
bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    bool ok = n >= 0;
    if (ok) { out.resize((size_t)n); ok = n == 0 || std::fread(out.data(), 1, (size_t)n, f) == (size_t)n; }
    std::fclose(f);
    return ok;
}
bool base64(const char* s, size_t n, std::vector<uint8_t>& out) {
    uint32_t acc = 0; int bits = 0;
    out.reserve(n / 4 * 3);
    for (size_t i = 0; i < n; ++i) {
        const char c = s[i];
        int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A';
        else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
        else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+' || c == '-') v = 62;
        else if (c == '/' || c == '_') v = 63;
        else if (c == '=' || c == '\n' || c == '\r') continue;
        else return false;
        acc = (acc << 6) | (uint32_t)v; bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back((uint8_t)(acc >> bits)); }
    }
    return true;
}
std::string percent_decode(const std::string& s) {
    std::string o;
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] == '%' && i + 2 < s.size() && std::isxdigit((unsigned char)s[i + 1]) && std::isxdigit((unsigned char)s[i + 2])) {
            o += (char)std::strtol(s.substr(i + 1, 2).c_str(), nullptr, 16); i += 2;
        } else o += s[i];
    }
    return o;
}

// glTF enums (IMR/include/glTFenum.h:5-36)
enum : int64_t { CT_BYTE = 5120, CT_UBYTE = 5121, CT_SHORT = 5122, CT_USHORT = 5123, CT_UINT = 5125, CT_FLOAT = 5126 };
enum : uint32_t { MODE_LINE_LOOP = 2, MODE_LINE_STRIP = 3, MODE_TRIANGLES = 4 };

struct Primitive {
    std::vector<float> points, normals;           // n*3, y and z negated
    std::vector<uint32_t> indices;
    bool has_normals = false, has_indices = false, skipped = false;
    uint32_t mode = MODE_TRIANGLES, source_index = 0;
};
struct Mesh { std::string name; std::vector<Primitive> primitives; };

}  // namespace

struct imrcd_gltf {
    std::vector<Mesh> meshes;
};

namespace {

struct Loader {
    std::string dir, err;
    JVal root;
    std::vector<std::vector<uint8_t>> buffers;
    std::vector<uint8_t> glb_bin; bool has_glb_bin = false;

    bool fail(const std::string& m) { if (err.empty()) err = m; return false; }

    bool parse(const std::string& path) {
        std::vector<uint8_t> file;
        if (!read_file(path, file)) return fail("cannot read " + path);
        const size_t slash = path.find_last_of("/\\");
        dir = slash == std::string::npos ? std::string() : path.substr(0, slash + 1);
        const char* js = (const char*)file.data(); size_t jn = file.size();
        if (file.size() >= 12 && !std::memcmp(file.data(), "glTF", 4)) {          // binary container: 12-byte header, then chunks (length, type, data)
            uint32_t version, total;
            std::memcpy(&version, file.data() + 4, 4); std::memcpy(&total, file.data() + 8, 4);
            if (version != 2 || total > file.size()) return fail("glb: bad header");
            size_t at = 12; js = nullptr;
            while (at + 8 <= total) {
                uint32_t len, type;
                std::memcpy(&len, file.data() + at, 4); std::memcpy(&type, file.data() + at + 4, 4);
                at += 8;
                if (at + len > total) return fail("glb: chunk runs past the end of the file");
                if (type == 0x4E4F534Au && !js) { js = (const char*)file.data() + at; jn = len; }
                else if (type == 0x004E4942u && !has_glb_bin) { glb_bin.assign(file.data() + at, file.data() + at + len); has_glb_bin = true; }
                at += (len + 3u) & ~3ull;
            }
            if (!js) return fail("glb: no JSON chunk");
        }
        JParser jp{js, js + jn, std::string()};
        if (!jp.value(root)) return fail(jp.err);
        jp.ws();
        if (jp.p != jp.end) return fail("json: trailing characters");
        if (root.type != JVal::OBJ) return fail("gltf: top level is not an object");
        return load_buffers() && validate();
    }

    bool load_buffers() {
        const JVal* bs = root.get("buffers");
        if (!bs || bs->type != JVal::ARR) return true;
        buffers.resize(bs->arr.size());
        for (size_t i = 0; i < bs->arr.size(); ++i) {
            const JVal& b = bs->arr[i];
            const JVal* uri = b.get("uri");
            if (!uri || uri->type != JVal::STR) {
                if (i == 0 && has_glb_bin) buffers[i] = glb_bin;
                else return fail("gltf: buffer " + std::to_string(i) + " has no uri");
            } else if (uri->str.compare(0, 5, "data:") == 0) {
                const size_t comma = uri->str.find(',');
                if (comma == std::string::npos || uri->str.compare(comma >= 7 ? comma - 7 : 0, 7, ";base64") != 0) return fail("gltf: data uri is not base64");
                if (!base64(uri->str.data() + comma + 1, uri->str.size() - comma - 1, buffers[i])) return fail("gltf: bad base64 in buffer " + std::to_string(i));
            } else if (!read_file(dir + percent_decode(uri->str), buffers[i])) return fail("gltf: cannot read buffer " + uri->str);
            const int64_t want = b.index_or("byteLength", -1);
            if (want < 0 || (uint64_t)want > buffers[i].size()) return fail("gltf: buffer " + std::to_string(i) + " is shorter than its byteLength");
        }
        return true;
    }

    // What the reference's reader refuses before any primitive is looked at (tinygltf: bufferView / accessor parsing): indices that
    // point nowhere, lengths that are not integers, views that run past their buffer.
    bool validate() {
        const JVal* views = root.get("bufferViews");
        if (views && views->type == JVal::ARR) for (size_t i = 0; i < views->arr.size(); ++i) {
            const JVal& v = views->arr[i];
            const int64_t buf = v.index_or("buffer", -1), len = v.index_or("byteLength", -1);
            const JVal* off = v.get("byteOffset");
            if (buf < 0 || (size_t)buf >= buffers.size()) return fail("gltf: bufferView " + std::to_string(i) + ": bad buffer index");
            if (len < 0 || (off && !off->is_index())) return fail("gltf: bufferView " + std::to_string(i) + ": bad byteLength / byteOffset");
            if ((uint64_t)v.index_or("byteOffset", 0) + (uint64_t)len > buffers[(size_t)buf].size()) return fail("gltf: bufferView " + std::to_string(i) + " runs past the end of its buffer");
        }
        const JVal* accs = root.get("accessors");
        if (accs && accs->type == JVal::ARR) for (size_t i = 0; i < accs->arr.size(); ++i) {
            const JVal* bv = accs->arr[i].get("bufferView");
            if (bv && (!bv->is_index() || !views || views->type != JVal::ARR || (size_t)bv->num >= views->arr.size()))
                return fail("gltf: accessor " + std::to_string(i) + ": bad bufferView index");
        }
        return true;
    }

    // One accessor, element by element: `comps` components of `ctype` each, out[k * comps + c] as T.
    template <typename T>
    bool read_accessor(int64_t id, int comps_wanted, const char* what, std::vector<T>& out, int64_t* ctype_out = nullptr) {
        const JVal* accs = root.get("accessors");
        if (!accs || accs->type != JVal::ARR || id < 0 || (size_t)id >= accs->arr.size()) return fail(std::string("gltf: ") + what + ": bad accessor index");
        const JVal& a = accs->arr[(size_t)id];
        if (a.get("sparse")) return fail(std::string("gltf: ") + what + ": sparse accessors are not supported");
        const int64_t ctype = a.index_or("componentType", -1), count = a.index_or("count", -1), view = a.index_or("bufferView", -1);
        const JVal* type = a.get("type");
        static const struct { const char* name; int n; } kTypes[] = {{"SCALAR", 1}, {"VEC2", 2}, {"VEC3", 3}, {"VEC4", 4}};
        int comps = 0;
        if (type && type->type == JVal::STR) for (const auto& t : kTypes) if (type->str == t.name) comps = t.n;
        if (comps != comps_wanted || count < 0) return fail(std::string("gltf: ") + what + ": unexpected accessor type");
        size_t csize;
        switch (ctype) {
            case CT_BYTE: case CT_UBYTE: csize = 1; break;
            case CT_SHORT: case CT_USHORT: csize = 2; break;
            case CT_UINT: case CT_FLOAT: csize = 4; break;
            default: return fail(std::string("gltf: ") + what + ": bad componentType");
        }
        if (ctype_out) *ctype_out = ctype;
        if (view < 0) {                                                       // no bufferView: zeros (glTF 2.0, 5.1.1)
            if (count > (int64_t(1) << 28)) return fail(std::string("gltf: ") + what + ": accessor without a bufferView is too large");
            out.assign((size_t)count * comps, T(0));
            return true;
        }
        const JVal* views = root.get("bufferViews");
        if (!views || views->type != JVal::ARR || (size_t)view >= views->arr.size()) return fail(std::string("gltf: ") + what + ": bad bufferView index");
        const JVal& v = views->arr[(size_t)view];
        const int64_t buf = v.index_or("buffer", -1);
        if (buf < 0 || (size_t)buf >= buffers.size()) return fail(std::string("gltf: ") + what + ": bad buffer index");
        const std::vector<uint8_t>& data = buffers[(size_t)buf];
        const uint64_t offset = (uint64_t)a.index_or("byteOffset", 0) + (uint64_t)v.index_or("byteOffset", 0);     // PrimitivesOfMeshes.cpp:738-742
        const uint64_t elem = csize * comps;
        uint64_t stride = (uint64_t)v.index_or("byteStride", 0);
        if (stride == 0) stride = elem;
        const uint64_t view_end = (uint64_t)v.index_or("byteOffset", 0) + (uint64_t)v.index_or("byteLength", 0);
        if ((uint64_t)count > data.size() ||                                  // (every element takes at least one byte: no overflow below)
            (count && (offset + stride * (uint64_t)(count - 1) + elem > view_end || view_end > data.size()))) return fail(std::string("gltf: ") + what + ": accessor runs past the end of its bufferView");
        out.assign((size_t)count * comps, T(0));
        for (int64_t k = 0; k < count; ++k) {
            const uint8_t* src = data.data() + offset + stride * (uint64_t)k;
            for (int c = 0; c < comps; ++c) {
                T val;
                switch (ctype) {
                    case CT_BYTE:   { int8_t x;   std::memcpy(&x, src + c, 1);      val = (T)x; } break;
                    case CT_UBYTE:  { uint8_t x;  std::memcpy(&x, src + c, 1);      val = (T)x; } break;
                    case CT_SHORT:  { int16_t x;  std::memcpy(&x, src + 2 * c, 2);  val = (T)x; } break;
                    case CT_USHORT: { uint16_t x; std::memcpy(&x, src + 2 * c, 2);  val = (T)x; } break;
                    case CT_UINT:   { uint32_t x; std::memcpy(&x, src + 4 * c, 4);  val = (T)x; } break;
                    default:        { float x;    std::memcpy(&x, src + 4 * c, 4);  val = (T)x; } break;
                }
                out[(size_t)k * comps + c] = val;
            }
        }
        return true;
    }

    bool primitive(const JVal& jp, Primitive& out) {
        const JVal* attrs = jp.get("attributes");
        if (!attrs || attrs->type != JVal::OBJ) return fail("gltf: primitive without attributes");
        const int64_t mode = jp.index_or("mode", MODE_TRIANGLES);
        if (mode > 6) return fail("gltf: bad draw mode");
        out.mode = mode == MODE_LINE_LOOP ? (uint32_t)MODE_LINE_STRIP : (uint32_t)mode;                           // PrimitivesOfMeshes.cpp:49-55
        // skinned (any JOINTS_n) or morphed (targets[0] moves POSITION): no collision triangles (PrimitivesOfMeshes.cpp:641,669)
        const JVal* targets = jp.get("targets");
        const bool morphed = targets && targets->type == JVal::ARR && !targets->arr.empty() && targets->arr[0].get("POSITION");
        if (morphed || attrs->get("JOINTS_0")) { out.skipped = true; return true; }
        const int64_t pos = attrs->index_or("POSITION", -1);
        if (pos < 0) return fail("gltf: primitive without POSITION");
        int64_t ctype = 0;
        if (!read_accessor<float>(pos, 3, "POSITION", out.points, &ctype)) return false;
        if (ctype != CT_FLOAT) return fail("gltf: POSITION is not float (the reference accepts float only, PrimitivesOfMeshes.cpp:81)");
        for (size_t i = 0; i < out.points.size(); i += 3) { out.points[i + 1] = -out.points[i + 1]; out.points[i + 2] = -out.points[i + 2]; }      // :83-87
        const int64_t nrm = attrs->index_or("NORMAL", -1);
        if (nrm >= 0) {
            if (!read_accessor<float>(nrm, 3, "NORMAL", out.normals, &ctype)) return false;
            if (ctype != CT_FLOAT) return fail("gltf: NORMAL is not float (PrimitivesOfMeshes.cpp:133)");
            if (out.normals.size() != out.points.size()) return fail("gltf: NORMAL and POSITION counts differ");
            for (size_t i = 0; i < out.normals.size(); i += 3) { out.normals[i + 1] = -out.normals[i + 1]; out.normals[i + 2] = -out.normals[i + 2]; }  // :134-139
            out.has_normals = true;
        }
        const int64_t idx = jp.index_or("indices", -1);
        if (idx >= 0) {
            if (!read_accessor<uint32_t>(idx, 1, "indices", out.indices, &ctype)) return false;
            if (ctype != CT_UBYTE && ctype != CT_USHORT && ctype != CT_UINT) return fail("gltf: indices are not unsigned integers");
            const uint64_t np = out.points.size() / 3;
            for (uint32_t i : out.indices) if (i >= np) return fail("gltf: index out of range");
            out.has_indices = true;
        }
        return true;
    }

    bool meshes(std::vector<Mesh>& out) {
        const JVal* ms = root.get("meshes");
        if (!ms) return true;
        if (ms->type != JVal::ARR) return fail("gltf: meshes is not an array");
        for (const JVal& jm : ms->arr) {
            out.emplace_back();
            Mesh& m = out.back();
            if (const JVal* nm = jm.get("name")) if (nm->type == JVal::STR) m.name = nm->str;
            const JVal* prims = jm.get("primitives");
            if (!prims || prims->type != JVal::ARR) return fail("gltf: mesh without primitives");
            // "Triangles first" (MeshesOfNodes.cpp:41-43): std::sort with a predicate that only looks at its left argument.  On the insertion
            // sort libstdc++ runs for up to 16 elements that leaves: the triangle-list primitives behind the first one in reverse order, then
            // everything else (the first primitive included) in file order.  Longer lists keep the same rule here (the reference's own order is
            // then whatever introsort makes of a predicate that is not an ordering).
            std::vector<size_t> order;
            auto is_tri = [&](size_t k) { return prims->arr[k].index_or("mode", MODE_TRIANGLES) == MODE_TRIANGLES; };
            for (size_t k = prims->arr.size(); k-- > 1;) if (is_tri(k)) order.push_back(k);
            for (size_t k = 0; k < prims->arr.size(); ++k) if (k == 0 || !is_tri(k)) order.push_back(k);
            for (size_t k : order) {
                m.primitives.emplace_back();
                m.primitives.back().source_index = (uint32_t)k;
                if (!primitive(prims->arr[k], m.primitives.back())) return false;
            }
        }
        return true;
    }
};

void put_err(char* err, uint64_t cap, const std::string& m) {
    if (!err || !cap) return;
    const size_t n = m.size() < cap - 1 ? m.size() : (size_t)cap - 1;
    std::memcpy(err, m.data(), n); err[n] = 0;
}

}  // namespace

extern "C" int imrcd_gltf_open(const char* path, imrcd_gltf** out, char* err, uint64_t err_cap) {
    if (!path || !out) { put_err(err, err_cap, "imrcd_gltf_open: null argument"); return IMRCD_E_ARG; }
    *out = nullptr;
    try {                                                                     // no exception crosses the ABI
        Loader ld;
        std::unique_ptr<imrcd_gltf> g(new imrcd_gltf);
        if (!ld.parse(path) || !ld.meshes(g->meshes)) { put_err(err, err_cap, ld.err); return IMRCD_E_ARG; }
        *out = g.release();
        return IMRCD_OK;
    } catch (const std::exception& e) {
        put_err(err, err_cap, std::string("imrcd_gltf_open: ") + e.what());
        return IMRCD_E_CAPACITY;
    }
}
extern "C" void imrcd_gltf_close(imrcd_gltf* g) { delete g; }
extern "C" int imrcd_gltf_mesh_count(const imrcd_gltf* g, uint32_t* n) {
    if (!g || !n) return IMRCD_E_ARG;
    *n = (uint32_t)g->meshes.size();
    return IMRCD_OK;
}
extern "C" int imrcd_gltf_primitive_count(const imrcd_gltf* g, uint32_t mesh, uint32_t* n) {
    if (!g || !n || mesh >= g->meshes.size()) return IMRCD_E_ARG;
    *n = (uint32_t)g->meshes[mesh].primitives.size();
    return IMRCD_OK;
}
extern "C" int imrcd_gltf_primitive(const imrcd_gltf* g, uint32_t mesh, uint32_t k, imrcd_gltf_primitive_view* out) {
    if (!g || !out || mesh >= g->meshes.size() || k >= g->meshes[mesh].primitives.size()) return IMRCD_E_ARG;
    const Primitive& p = g->meshes[mesh].primitives[k];
    out->points = p.points.empty() ? nullptr : p.points.data();
    out->normals = p.has_normals && !p.normals.empty() ? p.normals.data() : nullptr;
    out->indices = p.has_indices && !p.indices.empty() ? p.indices.data() : nullptr;
    out->n_points = p.points.size() / 3;
    out->n_indices = p.has_indices ? p.indices.size() : 0;
    out->mode = p.mode; out->skipped = p.skipped ? 1u : 0u; out->source_index = p.source_index; out->has_indices = p.has_indices ? 1u : 0u;
    return IMRCD_OK;
}
extern "C" int imrcd_gltf_build_mesh(imrcd_ctx* ctx, const imrcd_gltf* g, uint32_t mesh, uint32_t build_mode, uint32_t* mesh_id) {
    if (!ctx || !g || !mesh_id || mesh >= g->meshes.size()) return IMRCD_E_ARG;
    int rc = imrcd_mesh_begin(ctx);                                                                               // StartRecordOBBtree
    if (rc) return rc;
    static const uint32_t kNone = 0;
    for (const Primitive& p : g->meshes[mesh].primitives) {
        if (p.skipped) continue;
        const bool empty_indexed = p.has_indices && p.indices.empty();                                            // an indexed primitive that draws nothing
        rc = imrcd_mesh_add_primitive(ctx, p.points.data(), p.points.size() / 3, 3, p.has_normals ? p.normals.data() : nullptr,
                                      p.has_indices ? (empty_indexed ? &kNone : p.indices.data()) : nullptr, p.has_indices ? p.indices.size() : 0, p.mode);
        if (rc) return rc;
    }
    return imrcd_mesh_end(ctx, build_mode, mesh_id);                                                              // GetOBBtreeAndReset
}
extern "C" int imrcd_gltf_load(imrcd_ctx* ctx, const char* path, uint32_t build_mode, uint32_t* mesh_ids, uint32_t capacity, uint32_t* n_meshes) {
    if (!ctx || !path || !n_meshes) return IMRCD_E_ARG;
    imrcd_gltf* g = nullptr;
    char err[256] = {0};
    int rc = imrcd_gltf_open(path, &g, err, sizeof err);
    if (rc) { imr_ctx_set_error(ctx, err); return rc; }
    *n_meshes = (uint32_t)g->meshes.size();
    if (mesh_ids) {
        if (capacity < g->meshes.size()) { imrcd_gltf_close(g); imr_ctx_set_error(ctx, "imrcd_gltf_load: mesh_ids is too small"); return IMRCD_E_CAPACITY; }
        for (uint32_t m = 0; m < g->meshes.size() && !rc; ++m) rc = imrcd_gltf_build_mesh(ctx, g, m, build_mode, &mesh_ids[m]);
    }
    imrcd_gltf_close(g);
    return rc;
}
