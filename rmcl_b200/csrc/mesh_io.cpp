// mesh_io.cpp -- triangle-mesh import for b2_mesh_create_from_file.
//
// The reference loads its map with rm::import_embree_map(file) (rmcl_ros/src/nodes/micp_localization.cpp:188,
// rmcl_ros/src/rmcl/PCDSensorUpdaterEmbree.cpp:158), i.e. assimp -> all meshes of the scene, node transforms applied, triangles only.
// assimp is not available here; this reader covers the two self-contained formats the reference's example maps ship in besides
// COLLADA: Stanford PLY (ascii, binary_little_endian; float/double vertices, any list index type, polygons fan-triangulated like
// assimp's aiProcess_Triangulate for convex faces) and Wavefront OBJ (v / f with v, v/vt, v/vt/vn, v//vn and negative indices).
// Host-side set-up code: runs once per map.
#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

struct Reader {
    FILE* fp = nullptr;
    ~Reader() { if (fp) fclose(fp); }
};

int type_size(const std::string& t)
{
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
    if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
    if (t == "double" || t == "float64") return 8;
    return 0;
}
bool is_float_type(const std::string& t) { return t == "float" || t == "float32" || t == "double" || t == "float64"; }
bool is_signed_type(const std::string& t) { return t == "char" || t == "short" || t == "int" || t == "int8" || t == "int16" || t == "int32"; }

double read_scalar_bin(FILE* fp, const std::string& t, bool& ok)
{
    unsigned char b[8] = {0};
    const int n = type_size(t);
    if (n == 0 || fread(b, 1, (size_t)n, fp) != (size_t)n) { ok = false; return 0.0; }
    if (t == "float" || t == "float32") { float v; memcpy(&v, b, 4); return v; }
    if (t == "double" || t == "float64") { double v; memcpy(&v, b, 8); return v; }
    if (n == 1) return is_signed_type(t) ? (double)(int8_t)b[0] : (double)b[0];
    if (n == 2) { uint16_t v; memcpy(&v, b, 2); return is_signed_type(t) ? (double)(int16_t)v : (double)v; }
    uint32_t v; memcpy(&v, b, 4); return is_signed_type(t) ? (double)(int32_t)v : (double)v;
}

struct PlyProp { std::string name, type, list_count_type; bool is_list = false; };
struct PlyElem { std::string name; size_t count = 0; std::vector<PlyProp> props; };

int load_ply(const char* path, std::vector<float>& V, std::vector<uint32_t>& F, std::string& err)
{
    Reader r; r.fp = fopen(path, "rb");
    if (!r.fp) { err = "cannot open file"; return -1; }
    char line[1024];
    if (!fgets(line, sizeof(line), r.fp) || strncmp(line, "ply", 3) != 0) { err = "not a PLY file"; return -1; }
    bool ascii = true; std::vector<PlyElem> elems;
    while (true) {
        if (!fgets(line, sizeof(line), r.fp)) { err = "truncated PLY header"; return -1; }
        char a[64] = "", b[64] = "", c[64] = "", d[64] = "", e[64] = "";
        const int k = sscanf(line, "%63s %63s %63s %63s %63s", a, b, c, d, e);
        if (k < 1) continue;
        if (!strcmp(a, "end_header")) break;
        if (!strcmp(a, "format")) {
            if (!strcmp(b, "ascii")) ascii = true;
            else if (!strcmp(b, "binary_little_endian")) ascii = false;
            else { err = std::string("unsupported PLY format ") + b; return -1; }
        } else if (!strcmp(a, "element") && k >= 3) {
            PlyElem el; el.name = b; el.count = (size_t)strtoull(c, nullptr, 10); elems.push_back(el);
        } else if (!strcmp(a, "property") && !elems.empty()) {
            PlyProp p;
            if (!strcmp(b, "list") && k >= 5) { p.is_list = true; p.list_count_type = c; p.type = d; p.name = e; }
            else if (k >= 3) { p.type = b; p.name = c; }
            else { err = "malformed PLY property"; return -1; }
            if (type_size(p.type) == 0 || (p.is_list && type_size(p.list_count_type) == 0)) { err = "unknown PLY property type " + p.type; return -1; }
            elems.back().props.push_back(p);
        }
    }
    auto next_ascii = [&](double& v) -> bool { return fscanf(r.fp, "%lf", &v) == 1; };
    for (const PlyElem& el : elems) {
        int ix = -1, iy = -1, iz = -1;
        if (el.name == "vertex") {
            for (size_t p = 0; p < el.props.size(); p++) {
                if (el.props[p].name == "x") ix = (int)p;
                if (el.props[p].name == "y") iy = (int)p;
                if (el.props[p].name == "z") iz = (int)p;
            }
            if (ix < 0 || iy < 0 || iz < 0) { err = "PLY vertex element without x/y/z"; return -1; }
            V.reserve(3 * el.count);
        }
        for (size_t i = 0; i < el.count; i++) {
            float xyz[3] = {0, 0, 0};
            for (size_t p = 0; p < el.props.size(); p++) {
                const PlyProp& pr = el.props[p];
                bool ok = true;
                if (!pr.is_list) {
                    double v = 0.0;
                    if (ascii) ok = next_ascii(v); else v = read_scalar_bin(r.fp, pr.type, ok);
                    if (!ok) { err = "truncated PLY body"; return -1; }
                    if ((int)p == ix) xyz[0] = (float)v;
                    if ((int)p == iy) xyz[1] = (float)v;
                    if ((int)p == iz) xyz[2] = (float)v;
                } else {
                    double cnt = 0.0;
                    if (ascii) ok = next_ascii(cnt); else cnt = read_scalar_bin(r.fp, pr.list_count_type, ok);
                    if (!ok || cnt < 0 || cnt > 1e6) { err = "bad PLY list count"; return -1; }
                    std::vector<uint32_t> idx((size_t)cnt);
                    for (size_t j = 0; j < idx.size(); j++) {
                        double v = 0.0;
                        if (ascii) ok = next_ascii(v); else v = read_scalar_bin(r.fp, pr.type, ok);
                        if (!ok || v < 0) { err = "bad PLY list entry"; return -1; }
                        idx[j] = (uint32_t)v;
                    }
                    const bool is_face_list = el.name == "face" && (pr.name == "vertex_indices" || pr.name == "vertex_index") && !is_float_type(pr.type);
                    if (is_face_list)
                        for (size_t j = 2; j < idx.size(); j++) { F.push_back(idx[0]); F.push_back(idx[j - 1]); F.push_back(idx[j]); }   // triangle fan
                }
            }
            if (el.name == "vertex") { V.push_back(xyz[0]); V.push_back(xyz[1]); V.push_back(xyz[2]); }
        }
    }
    return 0;
}

int load_obj(const char* path, std::vector<float>& V, std::vector<uint32_t>& F, std::string& err)
{
    Reader r; r.fp = fopen(path, "rb");
    if (!r.fp) { err = "cannot open file"; return -1; }
    std::vector<char> buf(1 << 16);
    while (fgets(buf.data(), (int)buf.size(), r.fp)) {
        const char* s = buf.data();
        while (*s == ' ' || *s == '\t') s++;
        if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
            float x, y, z;
            if (sscanf(s + 1, "%f %f %f", &x, &y, &z) != 3) { err = "malformed OBJ vertex"; return -1; }
            V.push_back(x); V.push_back(y); V.push_back(z);
        } else if (s[0] == 'f' && (s[1] == ' ' || s[1] == '\t')) {
            std::vector<uint32_t> idx;
            const char* p = s + 1;
            while (*p) {
                while (*p == ' ' || *p == '\t') p++;
                if (!*p || *p == '\n' || *p == '\r' || *p == '#') break;
                char* end = nullptr;
                const long v = strtol(p, &end, 10);
                if (end == p) { err = "malformed OBJ face"; return -1; }
                const long nv = (long)(V.size() / 3);
                const long k = v > 0 ? v - 1 : nv + v;                    // 1-based, negative = relative to the end
                if (k < 0 || k >= nv) { err = "OBJ face index out of range"; return -1; }
                idx.push_back((uint32_t)k);
                p = end;
                while (*p && !isspace((unsigned char)*p)) p++;           // skip /vt/vn
            }
            for (size_t j = 2; j < idx.size(); j++) { F.push_back(idx[0]); F.push_back(idx[j - 1]); F.push_back(idx[j]); }
        }
    }
    return 0;
}

}  // namespace

// returns 0 on success; on failure a negative code and *err_out points to a static thread-local message
int b2_load_mesh_file(const char* path, std::vector<float>& V, std::vector<uint32_t>& F, const char** err_out)
{
    static thread_local std::string err;
    err.clear(); V.clear(); F.clear();
    const char* dot = strrchr(path, '.');
    std::string ext = dot ? dot + 1 : "";
    for (char& c : ext) c = (char)tolower((unsigned char)c);
    int rc;
    if (ext == "ply") rc = load_ply(path, V, F, err);
    else if (ext == "obj") rc = load_obj(path, V, F, err);
    else { err = "unsupported mesh format '." + ext + "' (supported: .ply, .obj)"; rc = -2; }
    if (rc == 0 && (V.empty() || F.empty())) { err = "mesh file holds no triangles"; rc = -3; }
    if (rc == 0) for (uint32_t i : F) if (i >= V.size() / 3) { err = "face index out of range"; rc = -1; break; }
    *err_out = err.c_str();
    return rc;
}
