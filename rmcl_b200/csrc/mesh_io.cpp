// mesh_io.cpp -- triangle-mesh import for b2_mesh_create_from_file.
//
// The reference loads its map with rm::import_embree_map(file) (rmcl_ros/src/nodes/micp_localization.cpp:188,
// rmcl_ros/src/rmcl/PCDSensorUpdaterEmbree.cpp:158), i.e. assimp -> all meshes of the scene, node transforms applied, triangles only.
// assimp is not available here; this reader covers the two self-contained formats the reference's example maps ship in besides
// assimp's generic path: Stanford PLY (ascii, binary_little_endian; float/double vertices, any list index type, polygons fan-triangulated like
// assimp's aiProcess_Triangulate for convex faces) and Wavefront OBJ (v / f with v, v/vt, v/vt/vn, v//vn and negative indices).
// Host-side set-up code: runs once per map.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <map>
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
    // an element count comes straight from the header: bound it by what the rest of the file can hold (>= 1 byte per property value in
    // binary, >= 2 characters in ascii) before anything is reserved -- a truncated or hostile header must not turn into a huge allocation
    {
        const long body = ftell(r.fp);
        fseek(r.fp, 0, SEEK_END);
        const long end = ftell(r.fp);
        fseek(r.fp, body, SEEK_SET);
        const size_t remaining = end > body ? (size_t)(end - body) : 0;
        for (const PlyElem& el : elems) {
            const size_t min_bytes = std::max<size_t>(1, el.props.size()) * (ascii ? 2 : 1);
            if (el.count > remaining / min_bytes + 1) { err = "PLY header announces more '" + el.name + "' elements than the file can hold"; return -1; }
        }
    }
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

// ---------------------------------------------------------------------------------------------------------------------
// COLLADA (.dae): the format of the reference's example maps (docs/MICPL.md:46-49 "tray.dae"; loaded through assimp by
// rm::import_embree_map, micp_localization.cpp:188).  Self-contained reader for what those files use: <library_geometries> meshes with
// <triangles> / <polylist> / <polygons> primitives (polygons fan-triangulated), POSITION sources through <vertices>, and the
// <visual_scene> node hierarchy with <matrix> / <translate> / <rotate> / <scale> applied to every <instance_geometry> (assimp applies node
// transforms the same way; geometries no node instantiates are taken untransformed).  <up_axis> is NOT applied, like rmagine's AssimpIO which
// switches assimp's up-axis conversion off [RM-recalled]; B2_DAE_APPLY_UP_AXIS=1 applies assimp's default conversion instead.  <unit> is
// ignored (assimp ignores it too).  No materials, cameras, skinning, <lines>, <tristrips>.
// ---------------------------------------------------------------------------------------------------------------------
struct XmlNode {
    std::string name, text;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<XmlNode> kids;
    const std::string& attr(const char* k) const { static const std::string none; for (const auto& a : attrs) if (a.first == k) return a.second; return none; }
    const XmlNode* child(const char* n) const { for (const XmlNode& k : kids) if (k.name == n) return &k; return nullptr; }
};

// minimal XML: elements, attributes, character data; comments, processing instructions, CDATA and DOCTYPE are skipped; no entity decoding
// beyond what numeric mesh data needs (none)
struct XmlParser {
    const char* p; const char* end; std::string& err; int depth = 0;
    void skip_ws() { while (p < end && isspace((unsigned char)*p)) p++; }
    bool starts(const char* s) const { const size_t n = strlen(s); return (size_t)(end - p) >= n && !strncmp(p, s, n); }
    bool skip_misc()
    {
        while (true) {
            skip_ws();
            if (starts("<?")) { const char* q = strstr(p, "?>"); if (!q) { err = "unterminated <? ?>"; return false; } p = q + 2; }
            else if (starts("<!--")) { const char* q = strstr(p, "-->"); if (!q) { err = "unterminated comment"; return false; } p = q + 3; }
            else if (starts("<!")) { while (p < end && *p != '>') p++; if (p < end) p++; }
            else return true;
        }
    }
    bool parse_element(XmlNode& out)
    {
        if (++depth > 64) { err = "XML nesting too deep"; return false; }
        if (p >= end || *p != '<') { err = "expected '<'"; return false; }
        p++;
        const char* s = p;
        while (p < end && !isspace((unsigned char)*p) && *p != '>' && *p != '/') p++;
        out.name.assign(s, p);
        while (true) {
            skip_ws();
            if (p >= end) { err = "unterminated tag"; return false; }
            if (*p == '/') { if (p + 1 < end && p[1] == '>') { p += 2; depth--; return true; } err = "malformed tag"; return false; }
            if (*p == '>') { p++; break; }
            const char* k = p;
            while (p < end && *p != '=' && !isspace((unsigned char)*p) && *p != '>') p++;
            std::string key(k, p);
            skip_ws();
            if (p >= end || *p != '=') { err = "attribute without value"; return false; }
            p++; skip_ws();
            if (p >= end || (*p != '"' && *p != '\'')) { err = "attribute value not quoted"; return false; }
            const char qc = *p++;
            const char* v = p;
            while (p < end && *p != qc) p++;
            if (p >= end) { err = "unterminated attribute value"; return false; }
            out.attrs.emplace_back(key, std::string(v, p));
            p++;
        }
        while (true) {
            const char* t = p;
            while (p < end && *p != '<') p++;
            out.text.append(t, p);
            if (p >= end) { err = "unterminated element <" + out.name + ">"; return false; }
            if (starts("</")) {
                while (p < end && *p != '>') p++;
                if (p < end) p++;
                depth--;
                return true;
            }
            if (starts("<!--")) { const char* q = strstr(p, "-->"); if (!q) { err = "unterminated comment"; return false; } p = q + 3; continue; }
            if (starts("<![CDATA[")) { const char* q = strstr(p, "]]>"); if (!q) { err = "unterminated CDATA"; return false; } out.text.append(p + 9, q); p = q + 3; continue; }
            if (starts("<?")) { const char* q = strstr(p, "?>"); if (!q) { err = "unterminated <? ?>"; return false; } p = q + 2; continue; }
            out.kids.emplace_back();
            if (!parse_element(out.kids.back())) return false;
        }
    }
};

struct M4 { double m[16]; };      // row-major
M4 m4_identity() { M4 r{}; for (int i = 0; i < 4; i++) r.m[i * 5] = 1.0; return r; }
M4 m4_mul(const M4& a, const M4& b)
{
    M4 r{};
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double acc = 0; for (int k = 0; k < 4; k++) acc += a.m[i * 4 + k] * b.m[k * 4 + j]; r.m[i * 4 + j] = acc; }
    return r;
}
bool parse_doubles(const std::string& t, std::vector<double>& out, size_t expect = 0)
{
    const char* s = t.c_str();
    while (true) {
        while (*s && isspace((unsigned char)*s)) s++;
        if (!*s) break;
        char* e = nullptr;
        const double v = strtod(s, &e);
        if (e == s) return false;
        out.push_back(v); s = e;
    }
    return expect == 0 || out.size() == expect;
}
// transform elements of a <node>, in document order, post-multiplied (COLLADA 1.4.1 spec, "node")
bool node_transform(const XmlNode& n, M4& T, std::string& err)
{
    T = m4_identity();
    for (const XmlNode& k : n.kids) {
        std::vector<double> v;
        M4 L = m4_identity();
        if (k.name == "matrix") { if (!parse_doubles(k.text, v, 16)) { err = "malformed <matrix>"; return false; } for (int i = 0; i < 16; i++) L.m[i] = v[i]; }
        else if (k.name == "translate") { if (!parse_doubles(k.text, v, 3)) { err = "malformed <translate>"; return false; } L.m[3] = v[0]; L.m[7] = v[1]; L.m[11] = v[2]; }
        else if (k.name == "scale") { if (!parse_doubles(k.text, v, 3)) { err = "malformed <scale>"; return false; } L.m[0] = v[0]; L.m[5] = v[1]; L.m[10] = v[2]; }
        else if (k.name == "rotate") {
            if (!parse_doubles(k.text, v, 4)) { err = "malformed <rotate>"; return false; }
            const double n2 = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            if (n2 > 0) {
                const double x = v[0] / n2, y = v[1] / n2, z = v[2] / n2, a = v[3] * 3.14159265358979323846 / 180.0, c = std::cos(a), s = std::sin(a), t = 1 - c;
                L.m[0] = t * x * x + c; L.m[1] = t * x * y - s * z; L.m[2] = t * x * z + s * y;
                L.m[4] = t * x * y + s * z; L.m[5] = t * y * y + c; L.m[6] = t * y * z - s * x;
                L.m[8] = t * x * z - s * y; L.m[9] = t * y * z + s * x; L.m[10] = t * z * z + c;
            }
        } else continue;
        T = m4_mul(T, L);
    }
    return true;
}

struct DaeGeom { std::vector<float> V; std::vector<uint32_t> F; };

bool dae_geometry(const XmlNode& geom, DaeGeom& out, std::string& err)
{
    const XmlNode* mesh = geom.child("mesh");
    if (!mesh) return true;                                   // <convex_mesh>, <spline>: nothing to trace
    std::map<std::string, const XmlNode*> sources;
    for (const XmlNode& k : mesh->kids) if (k.name == "source") sources[k.attr("id")] = &k;
    // <vertices>: POSITION -> source
    std::map<std::string, std::string> vertices_pos;
    for (const XmlNode& k : mesh->kids) if (k.name == "vertices")
        for (const XmlNode& in : k.kids) if (in.name == "input" && in.attr("semantic") == "POSITION") vertices_pos[k.attr("id")] = in.attr("source");
    for (const XmlNode& prim : mesh->kids) {
        const bool tri = prim.name == "triangles", plist = prim.name == "polylist", pgons = prim.name == "polygons";
        if (!tri && !plist && !pgons) continue;
        size_t stride = 0, voff = (size_t)-1; std::string vsrc;
        for (const XmlNode& in : prim.kids) if (in.name == "input") {
            const size_t off = (size_t)strtoul(in.attr("offset").c_str(), nullptr, 10);
            stride = std::max(stride, off + 1);
            if (in.attr("semantic") == "VERTEX") { voff = off; vsrc = in.attr("source"); }
        }
        if (voff == (size_t)-1 || vsrc.size() < 2) { err = "<" + prim.name + "> without a VERTEX input"; return false; }
        const auto vp = vertices_pos.find(vsrc.substr(1));
        if (vp == vertices_pos.end() || vp->second.size() < 2) { err = "VERTEX input does not resolve to a POSITION source"; return false; }
        const auto sp = sources.find(vp->second.substr(1));
        if (sp == sources.end()) { err = "POSITION source '" + vp->second + "' not found"; return false; }
        const XmlNode* fa = sp->second->child("float_array");
        if (!fa) { err = "POSITION source without <float_array>"; return false; }
        size_t pstride = 3;
        if (const XmlNode* tc = sp->second->child("technique_common")) if (const XmlNode* acc = tc->child("accessor")) if (!acc->attr("stride").empty()) pstride = (size_t)strtoul(acc->attr("stride").c_str(), nullptr, 10);
        if (pstride < 3) { err = "POSITION accessor stride < 3"; return false; }
        std::vector<double> pos;
        if (!parse_doubles(fa->text, pos)) { err = "malformed <float_array>"; return false; }
        const uint32_t base = (uint32_t)(out.V.size() / 3), nverts = (uint32_t)(pos.size() / pstride);
        for (uint32_t i = 0; i < nverts; i++) for (int k = 0; k < 3; k++) out.V.push_back((float)pos[(size_t)i * pstride + k]);
        auto emit_polygon = [&](const std::vector<double>& idx, size_t first, size_t count) -> bool {
            for (size_t j = 0; j < count; j++) { const double v = idx[(first + j) * stride + voff]; if (v < 0 || v >= nverts) { err = "COLLADA vertex index out of range"; return false; } }
            for (size_t j = 2; j < count; j++) {
                out.F.push_back(base + (uint32_t)idx[first * stride + voff]); out.F.push_back(base + (uint32_t)idx[(first + j - 1) * stride + voff]);
                out.F.push_back(base + (uint32_t)idx[(first + j) * stride + voff]);
            }
            return true;
        };
        if (pgons) {
            for (const XmlNode& pe : prim.kids) if (pe.name == "p") {
                std::vector<double> idx;
                if (!parse_doubles(pe.text, idx) || idx.size() % stride) { err = "malformed <p>"; return false; }
                if (!emit_polygon(idx, 0, idx.size() / stride)) return false;
            }
            continue;
        }
        const XmlNode* pe = prim.child("p");
        if (!pe) continue;                                    // count="0"
        std::vector<double> idx;
        if (!parse_doubles(pe->text, idx) || idx.size() % stride) { err = "malformed <p>"; return false; }
        const size_t ncorner = idx.size() / stride;
        if (tri) { if (ncorner % 3) { err = "<triangles> index count not a multiple of 3"; return false; } for (size_t f = 0; f < ncorner; f += 3) if (!emit_polygon(idx, f, 3)) return false; }
        else {
            const XmlNode* vc = prim.child("vcount");
            std::vector<double> cnt;
            if (!vc || !parse_doubles(vc->text, cnt)) { err = "<polylist> without <vcount>"; return false; }
            size_t first = 0;
            for (double c : cnt) { if (c < 0 || first + (size_t)c > ncorner) { err = "<vcount> exceeds <p>"; return false; } if (!emit_polygon(idx, first, (size_t)c)) return false; first += (size_t)c; }
        }
    }
    return true;
}

void dae_instantiate(const XmlNode& node, const M4& parent, const std::map<std::string, DaeGeom>& geoms, const std::map<std::string, const XmlNode*>& lib_nodes,
                     std::vector<float>& V, std::vector<uint32_t>& F, std::map<std::string, bool>& used, std::string& err, int depth)
{
    if (depth > 64) return;
    M4 L; if (!node_transform(node, L, err)) return;
    const M4 T = m4_mul(parent, L);
    for (const XmlNode& k : node.kids) {
        if (k.name == "instance_geometry" && k.attr("url").size() > 1) {
            const auto g = geoms.find(k.attr("url").substr(1));
            if (g == geoms.end()) continue;
            used[g->first] = true;
            const uint32_t base = (uint32_t)(V.size() / 3);
            for (size_t i = 0; i + 2 < g->second.V.size(); i += 3) {
                const double x = g->second.V[i], y = g->second.V[i + 1], z = g->second.V[i + 2];
                for (int r = 0; r < 3; r++) V.push_back((float)(T.m[r * 4] * x + T.m[r * 4 + 1] * y + T.m[r * 4 + 2] * z + T.m[r * 4 + 3]));
            }
            for (uint32_t f : g->second.F) F.push_back(base + f);
        } else if (k.name == "node") dae_instantiate(k, T, geoms, lib_nodes, V, F, used, err, depth + 1);
        else if (k.name == "instance_node" && k.attr("url").size() > 1) {
            const auto n = lib_nodes.find(k.attr("url").substr(1));
            if (n != lib_nodes.end()) dae_instantiate(*n->second, T, geoms, lib_nodes, V, F, used, err, depth + 1);
        }
    }
}

void collect_nodes(const XmlNode& n, std::map<std::string, const XmlNode*>& out) { for (const XmlNode& k : n.kids) if (k.name == "node") { if (!k.attr("id").empty()) out[k.attr("id")] = &k; collect_nodes(k, out); } }

int load_dae(const char* path, std::vector<float>& V, std::vector<uint32_t>& F, std::string& err)
{
    Reader r; r.fp = fopen(path, "rb");
    if (!r.fp) { err = "cannot open file"; return -1; }
    fseek(r.fp, 0, SEEK_END);
    const long sz = ftell(r.fp);
    fseek(r.fp, 0, SEEK_SET);
    if (sz <= 0) { err = "empty file"; return -1; }
    std::string txt((size_t)sz, '\0');
    if (fread(&txt[0], 1, (size_t)sz, r.fp) != (size_t)sz) { err = "short read"; return -1; }
    XmlNode root;
    XmlParser xp{txt.c_str(), txt.c_str() + txt.size(), err};
    if (!xp.skip_misc() || !xp.parse_element(root)) { if (err.empty()) err = "malformed XML"; return -1; }
    if (root.name != "COLLADA") { err = "not a COLLADA document (root element <" + root.name + ">)"; return -1; }
    std::map<std::string, DaeGeom> geoms; std::vector<std::string> order;
    std::map<std::string, const XmlNode*> lib_nodes;
    for (const XmlNode& lib : root.kids) {
        if (lib.name == "library_geometries")
            for (const XmlNode& g : lib.kids) if (g.name == "geometry") { DaeGeom dg; if (!dae_geometry(g, dg, err)) return -1; if (!dg.F.empty()) { geoms[g.attr("id")] = std::move(dg); order.push_back(g.attr("id")); } }
        if (lib.name == "library_nodes") collect_nodes(lib, lib_nodes);
    }
    M4 rootT = m4_identity();
    if (const char* e = getenv("B2_DAE_APPLY_UP_AXIS")) if (atoi(e)) {
        // assimp's default: bring the document's up axis to +Y (ColladaLoader: UP_X and UP_Z root rotations)
        std::string up = "Y_UP";
        if (const XmlNode* as = root.child("asset")) if (const XmlNode* ua = as->child("up_axis")) up = ua->text;
        if (up.find("Z_UP") != std::string::npos) { const double m[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, -1, 0, 0, 0, 0, 0, 1}; memcpy(rootT.m, m, sizeof(m)); }
        else if (up.find("X_UP") != std::string::npos) { const double m[16] = {0, -1, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; memcpy(rootT.m, m, sizeof(m)); }
    }
    std::map<std::string, bool> used;
    // the instantiated scene (<scene><instance_visual_scene>) or, lacking that, every visual scene
    std::string want;
    if (const XmlNode* sc = root.child("scene")) if (const XmlNode* iv = sc->child("instance_visual_scene")) if (iv->attr("url").size() > 1) want = iv->attr("url").substr(1);
    for (const XmlNode& lib : root.kids) if (lib.name == "library_visual_scenes")
        for (const XmlNode& vs : lib.kids) if (vs.name == "visual_scene" && (want.empty() || vs.attr("id") == want)) {
            XmlNode holder = vs; holder.name = "node";
            // visual_scene itself carries no transform elements; its <node> children do
            dae_instantiate(holder, rootT, geoms, lib_nodes, V, F, used, err, 0);
            if (!err.empty()) return -1;
        }
    if (used.empty()) for (const std::string& id : order) {      // no scene graph at all: take the geometries as they are
        const DaeGeom& g = geoms[id];
        const uint32_t base = (uint32_t)(V.size() / 3);
        V.insert(V.end(), g.V.begin(), g.V.end());
        for (uint32_t f : g.F) F.push_back(base + f);
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
    try {
    if (ext == "ply") rc = load_ply(path, V, F, err);
    else if (ext == "obj") rc = load_obj(path, V, F, err);
    else if (ext == "dae") rc = load_dae(path, V, F, err);
    else { err = "unsupported mesh format '." + ext + "' (supported: .ply, .obj, .dae)"; rc = -2; }
    if (rc == 0 && (V.empty() || F.empty())) { err = "mesh file holds no triangles"; rc = -3; }
    if (rc == 0) for (uint32_t i : F) if (i >= V.size() / 3) { err = "face index out of range"; rc = -1; break; }
    } catch (const std::bad_alloc&) { err = "out of host memory while reading the mesh"; rc = -4; V.clear(); F.clear(); }
    catch (const std::exception& e) { err = std::string("mesh import failed: ") + e.what(); rc = -1; V.clear(); F.clear(); }
    *err_out = err.c_str();
    return rc;
}
