// api.cu -- C ABI (include/rmcl_b200.h) over the sm_100a kernels.  No CPU fallback anywhere: every entry point either runs
// the CUDA path or fails with B2_ERR_CUDA.
#include "kernels.cuh"
#include "lbvh.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

// ---------------------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { (void)cudaGetLastError(); } if (e_ != cudaSuccess) return fail(B2_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define LAUNCHED() do { g_launches.fetch_add(1, std::memory_order_relaxed); cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return fail(B2_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define NOTNULL(p) do { if (!(p)) return fail(B2_ERR_INVALID, "%s: null argument '%s'", __func__, #p); } while (0)

extern "C" const char* b2_last_error(void) { return g_err; }
extern "C" int b2_version(void) { return 100; }
extern "C" uint64_t b2_kernel_launch_count(void) { return g_launches.load(); }
// pending (not yet consumed) CUDA runtime error of the calling thread, "" if none; does not clear it (test hygiene: no entry point may leave one behind)
extern "C" const char* b2_peek_cuda_error(void) { const cudaError_t e = cudaPeekAtLastError(); return e == cudaSuccess ? "" : cudaGetErrorString(e); }
extern "C" int b2_device_count(int* n) { NOTNULL(n); CU(cudaGetDeviceCount(n)); return B2_OK; }

// ---------------------------------------------------------------------------------------------------------------------
struct b2_mesh {
    std::atomic<int> refs{1};           // the creator's reference + one per b2_rcc / b2_pf handle: the BVH stays resident until the last user is gone
    int device = 0; int build_mode = 0;
    B2Node8* d_nodes = nullptr; B2Tri* d_tris = nullptr;
    uint32_t n_nodes = 0, n_tris = 0, n_faces = 0, n_verts = 0, max_depth = 0;
    float build_ms = 0.f, sah = 0.f, abs_max[3] = {0.f, 0.f, 0.f};
    std::vector<uint32_t> level_begin;          // device-built maps: node index ranges per tree level (for b2_mesh_refit)
    uint32_t* d_faces = nullptr;                // device-built maps keep the face list for b2_mesh_refit (12 B per face)
    BvhView view() const
    {
        BvhView v; v.nodes = reinterpret_cast<const float4*>(d_nodes); v.tris = reinterpret_cast<const float4*>(d_tris);
        v.bx = abs_max[0]; v.by = abs_max[1]; v.bz = abs_max[2]; return v;
    }
};

template <typename T> struct DevBuf {
    T* p = nullptr; size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return B2_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, sizeof(T) * n);
        if (e != cudaSuccess) return fail(B2_ERR_OOM, "cudaMalloc(%zu bytes) failed: %s", sizeof(T) * n, cudaGetErrorString(e));
        cap = n; return B2_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
#define RES(call) do { int r_ = (call); if (r_ != B2_OK) return r_; } while (0)

extern "C" int b2_mesh_create(const float* verts, uint32_t nv, const uint32_t* faces, uint32_t nf, int device, int build_mode, b2_mesh** out)
{
    NOTNULL(out); *out = nullptr;
    if (nf == 0 || nv == 0) return fail(B2_ERR_NO_MAP, "EMPTY MAP: %u vertices, %u faces", nv, nf);
    NOTNULL(verts); NOTNULL(faces);
    if (build_mode != B2_BUILD_HOST_SAH && build_mode != B2_BUILD_DEVICE_LBVH) return fail(B2_ERR_UNSUPPORTED, "unknown build_mode %d", build_mode);
    int ndev = 0; CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(B2_ERR_INVALID, "device %d out of range (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    if (build_mode == B2_BUILD_DEVICE_LBVH) {
        for (size_t i = 0; i < 3 * (size_t)nf; i++) if (faces[i] >= nv) return fail(B2_ERR_INVALID, "BVH build failed: face index out of range");
        for (size_t i = 0; i < 3 * (size_t)nv; i++) if (!std::isfinite(verts[i])) return fail(B2_ERR_INVALID, "BVH build failed: non-finite vertex");
        const auto t0 = std::chrono::steady_clock::now();
        float* d_v = nullptr; uint32_t* d_f = nullptr;
        cudaError_t e = cudaMalloc((void**)&d_v, sizeof(float) * 3 * (size_t)nv);
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_f, sizeof(uint32_t) * 3 * (size_t)nf);
        if (e == cudaSuccess) e = cudaMemcpy(d_v, verts, sizeof(float) * 3 * (size_t)nv, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_f, faces, sizeof(uint32_t) * 3 * (size_t)nf, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(d_v); cudaFree(d_f); return fail(B2_ERR_CUDA, "mesh upload failed: %s", cudaGetErrorString(e)); }
        b2_mesh* m = new (std::nothrow) b2_mesh();
        if (!m) { cudaFree(d_v); cudaFree(d_f); return fail(B2_ERR_OOM, "out of host memory"); }
        const char* err = "";
        const int rc = lbvh_build_device(d_v, nv, d_f, nf, &m->d_nodes, &m->n_nodes, &m->d_tris, &m->n_tris, &m->max_depth, m->abs_max, &err, &m->level_begin);
        g_launches.fetch_add(5 + m->max_depth);
        cudaFree(d_v);
        if (rc == 0) m->d_faces = d_f; else cudaFree(d_f);
        if (rc != 0) { delete m; return fail(rc == -5 ? B2_ERR_INVALID : B2_ERR_CUDA, "BVH build failed: %s (%s)", err, cudaGetErrorString(cudaGetLastError())); }
        m->device = device; m->build_mode = build_mode; m->n_faces = nf; m->n_verts = nv; m->sah = 0.f;
        m->build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        *out = m;
        return B2_OK;
    }
    const auto t0 = std::chrono::steady_clock::now();
    B2BvhHost hb; const char* err = "";
    const int rc = b2_build_bvh8_host(verts, nv, faces, nf, &hb, &err);
    if (rc != 0) return fail(rc == -3 ? B2_ERR_NO_MAP : (rc == -4 ? B2_ERR_OOM : B2_ERR_INVALID), "BVH build failed: %s", err);
    b2_mesh* m = new (std::nothrow) b2_mesh();
    if (!m) { b2_free_bvh8_host(&hb); return fail(B2_ERR_OOM, "out of host memory"); }
    m->device = device; m->build_mode = build_mode; m->n_nodes = hb.n_nodes; m->n_tris = hb.n_tris; m->n_faces = nf; m->n_verts = nv;
    m->max_depth = hb.max_depth; m->sah = hb.sah_cost;
    for (int k = 0; k < 3; k++) m->abs_max[k] = hb.abs_max[k];
    cudaError_t e = cudaMalloc((void**)&m->d_nodes, sizeof(B2Node8) * (size_t)hb.n_nodes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_tris, sizeof(B2Tri) * (size_t)std::max(hb.n_tris, 1u));
    if (e == cudaSuccess) e = cudaMemcpy(m->d_nodes, hb.nodes, sizeof(B2Node8) * (size_t)hb.n_nodes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(m->d_tris, hb.tris, sizeof(B2Tri) * (size_t)hb.n_tris, cudaMemcpyHostToDevice);
    b2_free_bvh8_host(&hb);
    if (e != cudaSuccess) { if (m->d_nodes) cudaFree(m->d_nodes); if (m->d_tris) cudaFree(m->d_tris); delete m; return fail(B2_ERR_CUDA, "map upload failed: %s", cudaGetErrorString(e)); }
    m->build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = m;
    return B2_OK;
}

int b2_load_mesh_file(const char* path, std::vector<float>& V, std::vector<uint32_t>& F, const char** err_out);   // mesh_io.cpp

extern "C" int b2_mesh_file_load(const char* path, float** verts, uint32_t* nv, uint32_t** faces, uint32_t* nf)
{
    NOTNULL(path); NOTNULL(verts); NOTNULL(nv); NOTNULL(faces); NOTNULL(nf);
    *verts = nullptr; *faces = nullptr; *nv = 0; *nf = 0;
    std::vector<float> V; std::vector<uint32_t> F; const char* err = "";
    const int rc = b2_load_mesh_file(path, V, F, &err);
    if (rc != 0) return fail(rc == -3 ? B2_ERR_NO_MAP : B2_ERR_INVALID, "mesh import of '%s' failed: %s", path, err);
    float* v = (float*)malloc(sizeof(float) * V.size()); uint32_t* f = (uint32_t*)malloc(sizeof(uint32_t) * F.size());
    if (!v || !f) { free(v); free(f); return fail(B2_ERR_OOM, "out of host memory"); }
    memcpy(v, V.data(), sizeof(float) * V.size()); memcpy(f, F.data(), sizeof(uint32_t) * F.size());
    *verts = v; *faces = f; *nv = (uint32_t)(V.size() / 3); *nf = (uint32_t)(F.size() / 3);
    return B2_OK;
}
extern "C" void b2_mesh_file_free(float* verts, uint32_t* faces) { free(verts); free(faces); }

extern "C" int b2_mesh_create_from_file(const char* path, int device, int build_mode, b2_mesh** out)
{
    NOTNULL(out); *out = nullptr; NOTNULL(path);
    std::vector<float> V; std::vector<uint32_t> F; const char* err = "";
    const int rc = b2_load_mesh_file(path, V, F, &err);
    if (rc != 0) return fail(rc == -3 ? B2_ERR_NO_MAP : B2_ERR_INVALID, "mesh import of '%s' failed: %s", path, err);
    return b2_mesh_create(V.data(), (uint32_t)(V.size() / 3), F.data(), (uint32_t)(F.size() / 3), device, build_mode, out);
}

// Embree / OptiX scene re-commit after the vertices moved (the reference flags dependants with `outdated`,
// Correspondences.hpp:26-31): refit of the resident tree, no rebuild.  Only for device-built maps (they keep their level ranges and faces).
extern "C" int b2_mesh_refit(b2_mesh* m, const float* verts, uint32_t nv, int src_is_device)
{
    NOTNULL(m); NOTNULL(verts);
    if (m->level_begin.size() < 2 || !m->d_faces) return fail(B2_ERR_UNSUPPORTED, "refit needs a map built with B2_BUILD_DEVICE_LBVH");
    if (nv != m->n_verts) return fail(B2_ERR_INVALID, "refit keeps the topology: %u vertices given, the map has %u", nv, m->n_verts);
    CU(cudaSetDevice(m->device));
    const auto t0 = std::chrono::steady_clock::now();
    if (!src_is_device) for (size_t i = 0; i < 3 * (size_t)nv; i++) if (!std::isfinite(verts[i])) return fail(B2_ERR_INVALID, "refit: non-finite vertex");
    DevBuf<float> d_v; DevBuf<unsigned int> d_bits;
    const float* vp = verts;
    if (!src_is_device) {
        RES(d_v.reserve(3 * (size_t)nv));
        cudaError_t e = cudaMemcpy(d_v.p, verts, sizeof(float) * 3 * (size_t)nv, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { d_v.release(); return fail(B2_ERR_CUDA, "refit upload failed: %s", cudaGetErrorString(e)); }
        vp = d_v.p;
    }
    int rc = d_bits.reserve(3);
    if (rc == B2_OK && cudaMemset(d_bits.p, 0, 3 * sizeof(unsigned int)) != cudaSuccess) rc = fail(B2_ERR_CUDA, "refit: memset failed");
    if (rc == B2_OK) {
        cudaDeviceSynchronize();                                            // no trace of any handle may be in flight on the old boxes
        for (size_t l = m->level_begin.size() - 1; l-- > 0;) {
            const uint32_t b = m->level_begin[l], e = m->level_begin[l + 1];
            if (e > b) { k_bvh8_refit_level<<<(e - b + 127) / 128, 128>>>(b, e, m->d_nodes, m->d_tris, vp, m->d_faces); g_launches.fetch_add(1); }
        }
        k_abs_max<<<296, 256>>>(vp, nv, d_bits.p); g_launches.fetch_add(1);
        unsigned int bits[3] = {0, 0, 0};
        cudaError_t e = cudaMemcpy(bits, d_bits.p, sizeof(bits), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(B2_ERR_CUDA, "refit failed: %s", cudaGetErrorString(e));
        else for (int k = 0; k < 3; k++) memcpy(&m->abs_max[k], &bits[k], 4);
    }
    d_v.release(); d_bits.release();
    if (rc == B2_OK) m->build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return rc;
}

// ---- BVH blob: build once, ship to the other ranks / to disk (b2_mesh_bvh_blob for broadcast) ---------------------------
struct B2BlobHeader {
    char     magic[8];                 // "B2BVH8F\0"
    uint32_t version, node_bytes, tri_bytes, n_nodes, n_tris, n_faces, n_verts, max_depth;
    int32_t  build_mode;
    float    abs_max[3], sah;
    uint32_t pad;                      // 64 bytes: the node array that follows stays 16-byte aligned inside an aligned buffer
};
static_assert(sizeof(B2BlobHeader) == 64, "blob header must be 64 bytes");
static const char kBlobMagic[8] = {'B', '2', 'B', 'V', 'H', '8', 'F', 0};

extern "C" int b2_mesh_blob_size(const b2_mesh* m, uint64_t* bytes)
{
    NOTNULL(m); NOTNULL(bytes);
    *bytes = sizeof(B2BlobHeader) + (uint64_t)m->n_nodes * sizeof(B2Node8) + (uint64_t)m->n_tris * sizeof(B2Tri);
    return B2_OK;
}

extern "C" int b2_mesh_export_blob(const b2_mesh* m, void* dst_host, uint64_t capacity)
{
    NOTNULL(m); NOTNULL(dst_host);
    uint64_t need = 0; b2_mesh_blob_size(m, &need);
    if (capacity < need) return fail(B2_ERR_INVALID, "blob buffer too small: %llu < %llu bytes", (unsigned long long)capacity, (unsigned long long)need);
    CU(cudaSetDevice(m->device));
    B2BlobHeader hd; memset(&hd, 0, sizeof(hd));
    memcpy(hd.magic, kBlobMagic, 8);
    hd.version = 1; hd.node_bytes = sizeof(B2Node8); hd.tri_bytes = sizeof(B2Tri); hd.n_nodes = m->n_nodes; hd.n_tris = m->n_tris; hd.n_faces = m->n_faces;
    hd.n_verts = m->n_verts; hd.max_depth = m->max_depth; hd.build_mode = m->build_mode; hd.sah = m->sah;
    for (int k = 0; k < 3; k++) hd.abs_max[k] = m->abs_max[k];
    char* p = static_cast<char*>(dst_host);
    memcpy(p, &hd, sizeof(hd));
    CU(cudaMemcpy(p + sizeof(hd), m->d_nodes, (size_t)m->n_nodes * sizeof(B2Node8), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(p + sizeof(hd) + (size_t)m->n_nodes * sizeof(B2Node8), m->d_tris, (size_t)m->n_tris * sizeof(B2Tri), cudaMemcpyDeviceToHost));
    return B2_OK;
}

extern "C" int b2_mesh_create_from_blob(const void* blob_host, uint64_t bytes, int device, b2_mesh** out)
{
    NOTNULL(out); *out = nullptr; NOTNULL(blob_host);
    if (bytes < sizeof(B2BlobHeader)) return fail(B2_ERR_INVALID, "BVH blob truncated (%llu bytes)", (unsigned long long)bytes);
    B2BlobHeader hd; memcpy(&hd, blob_host, sizeof(hd));
    if (memcmp(hd.magic, kBlobMagic, 8) != 0 || hd.version != 1 || hd.node_bytes != sizeof(B2Node8) || hd.tri_bytes != sizeof(B2Tri))
        return fail(B2_ERR_INVALID, "not a BVH blob of this library version");
    const uint64_t need = sizeof(hd) + (uint64_t)hd.n_nodes * sizeof(B2Node8) + (uint64_t)hd.n_tris * sizeof(B2Tri);
    if (hd.n_nodes == 0 || hd.n_tris == 0) return fail(B2_ERR_NO_MAP, "EMPTY MAP in BVH blob");
    if (bytes < need || hd.max_depth > B2_TRAVERSAL_STACK - 4) return fail(B2_ERR_INVALID, "BVH blob inconsistent (%llu of %llu bytes, depth %u)", (unsigned long long)bytes, (unsigned long long)need, hd.max_depth);
    // structural check of the indices the traversal follows (a corrupt blob must not turn into out-of-bounds device reads)
    const B2Node8* nodes = reinterpret_cast<const B2Node8*>(static_cast<const char*>(blob_host) + sizeof(hd));
    for (uint32_t i = 0; i < hd.n_nodes; i++) {
        uint32_t n_inner = 0, tri_end = 0;
        for (int sl = 0; sl < 8; sl++) {
            const uint8_t meta = nodes[i].meta[sl];
            if ((nodes[i].imask >> sl) & 1u) n_inner++;
            else if (meta) { const uint32_t cnt = (meta >> 5) == 7 ? 3 : ((meta >> 5) == 3 ? 2 : 1); tri_end = std::max(tri_end, (uint32_t)(meta & 0x1fu) + cnt); }
        }
        if ((n_inner && (uint64_t)nodes[i].child_base + n_inner > hd.n_nodes) || (tri_end && (uint64_t)nodes[i].tri_base + tri_end > hd.n_tris))
            return fail(B2_ERR_INVALID, "BVH blob corrupt: node %u points outside the arrays", i);
    }
    int ndev = 0; CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(B2_ERR_INVALID, "device %d out of range (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    const auto t0 = std::chrono::steady_clock::now();
    b2_mesh* m = new (std::nothrow) b2_mesh();
    if (!m) return fail(B2_ERR_OOM, "out of host memory");
    m->device = device; m->build_mode = hd.build_mode; m->n_nodes = hd.n_nodes; m->n_tris = hd.n_tris; m->n_faces = hd.n_faces; m->n_verts = hd.n_verts;
    m->max_depth = hd.max_depth; m->sah = hd.sah;
    for (int k = 0; k < 3; k++) m->abs_max[k] = hd.abs_max[k];
    cudaError_t e = cudaMalloc((void**)&m->d_nodes, sizeof(B2Node8) * (size_t)hd.n_nodes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_tris, sizeof(B2Tri) * (size_t)hd.n_tris);
    if (e == cudaSuccess) e = cudaMemcpy(m->d_nodes, nodes, sizeof(B2Node8) * (size_t)hd.n_nodes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(m->d_tris, reinterpret_cast<const char*>(nodes) + sizeof(B2Node8) * (size_t)hd.n_nodes, sizeof(B2Tri) * (size_t)hd.n_tris, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { if (m->d_nodes) cudaFree(m->d_nodes); if (m->d_tris) cudaFree(m->d_tris); delete m; (void)cudaGetLastError(); return fail(B2_ERR_CUDA, "BVH blob upload failed: %s", cudaGetErrorString(e)); }
    m->build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = m;
    return B2_OK;
}

static void mesh_unref(b2_mesh* m)
{
    if (!m || m->refs.fetch_sub(1) != 1) return;
    cudaSetDevice(m->device);
    if (m->d_nodes) cudaFree(m->d_nodes);
    if (m->d_tris) cudaFree(m->d_tris);
    if (m->d_faces) cudaFree(m->d_faces);
    delete m;
    (void)cudaGetLastError();
}
// Drops the creator's reference.  Handles created on the map keep it alive (rm::EmbreeMapPtr is a shared_ptr in the reference as well,
// micp_localization.cpp:545), so the order in which a garbage collector destroys map and handles does not matter.
extern "C" int b2_mesh_destroy(b2_mesh* m) { mesh_unref(m); return B2_OK; }

extern "C" int b2_mesh_get_info(const b2_mesh* m, b2_mesh_info* info)
{
    NOTNULL(m); NOTNULL(info);
    info->n_faces = m->n_faces; info->n_vertices = m->n_verts; info->n_nodes = m->n_nodes; info->n_leaf_tris = m->n_tris; info->max_depth = m->max_depth;
    info->bvh_bytes = (uint64_t)m->n_nodes * sizeof(B2Node8) + (uint64_t)m->n_tris * sizeof(B2Tri);
    info->build_ms = m->build_ms; info->device = m->device; info->build_mode = m->build_mode; info->sah_cost = m->sah;
    return B2_OK;
}

static int intersect_impl(const b2_mesh* m, const float* origs, const float* dirs, uint32_t n, float tfar,
                          float* t_out, uint32_t* face_out, float* ng_out, uint8_t* hit_out, double* mean_nodes, double* mean_tris)
{
    NOTNULL(m);
    if (n == 0) return B2_OK;
    NOTNULL(origs); NOTNULL(dirs);
    CU(cudaSetDevice(m->device));
    DevBuf<float> d_o, d_d, d_t, d_ng; DevBuf<uint32_t> d_f; DevBuf<uint8_t> d_h; DevBuf<unsigned long long> d_c;
    int rc = B2_OK;
    auto cleanup = [&]() { d_o.release(); d_d.release(); d_t.release(); d_ng.release(); d_f.release(); d_h.release(); d_c.release(); };
    if ((rc = d_o.reserve(3 * (size_t)n)) || (rc = d_d.reserve(3 * (size_t)n)) || (rc = d_t.reserve(n)) || (rc = d_ng.reserve(3 * (size_t)n)) ||
        (rc = d_f.reserve(n)) || (rc = d_h.reserve(n)) || (rc = d_c.reserve(2))) { cleanup(); return rc; }
    cudaError_t e = cudaMemcpy(d_o.p, origs, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_d.p, dirs, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(d_c.p, 0, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) {
        const uint32_t grid = (n + 127) / 128;
        if (mean_nodes || mean_tris) k_intersect<true><<<grid, 128>>>(m->view(), d_o.p, d_d.p, n, tfar, d_t.p, d_f.p, d_ng.p, d_h.p, d_c.p);
        else k_intersect<false><<<grid, 128>>>(m->view(), d_o.p, d_d.p, n, tfar, d_t.p, d_f.p, d_ng.p, d_h.p, d_c.p);
        g_launches.fetch_add(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess && t_out) e = cudaMemcpy(t_out, d_t.p, sizeof(float) * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && face_out) e = cudaMemcpy(face_out, d_f.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && ng_out) e = cudaMemcpy(ng_out, d_ng.p, sizeof(float) * 3 * (size_t)n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && hit_out) e = cudaMemcpy(hit_out, d_h.p, n, cudaMemcpyDeviceToHost);
    unsigned long long c[2] = {0, 0};
    if (e == cudaSuccess && (mean_nodes || mean_tris)) e = cudaMemcpy(c, d_c.p, sizeof(c), cudaMemcpyDeviceToHost);
    cleanup();
    if (e != cudaSuccess) return fail(B2_ERR_CUDA, "b2_mesh_intersect: %s", cudaGetErrorString(e));
    if (mean_nodes) *mean_nodes = (double)c[0] / (double)n;
    if (mean_tris) *mean_tris = (double)c[1] / (double)n;
    return B2_OK;
}

extern "C" int b2_mesh_intersect(const b2_mesh* m, const float* origs, const float* dirs, uint32_t n, float tfar,
                                 float* t_out, uint32_t* face_out, float* ng_out, uint8_t* hit_out)
{
    return intersect_impl(m, origs, dirs, n, tfar, t_out, face_out, ng_out, hit_out, nullptr, nullptr);
}
extern "C" int b2_mesh_intersect_stats(const b2_mesh* m, const float* origs, const float* dirs, uint32_t n, float tfar, double* mean_nodes, double* mean_tris)
{
    double a = 0, b = 0;
    int rc = intersect_impl(m, origs, dirs, n, tfar, nullptr, nullptr, nullptr, nullptr, &a, &b);
    if (rc == B2_OK) { if (mean_nodes) *mean_nodes = a; if (mean_tris) *mean_tris = b; }
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------------
// RCC handle
// ---------------------------------------------------------------------------------------------------------------------
struct HostPin {            // pinned (mapped) staging for small results
    b2_transform T[3]; b2_cross_stats S[2]; IcpState icp; IcpState icp_out; volatile unsigned int flag; unsigned int pad[3];
};

struct b2_rcc {
    b2_mesh* map = nullptr; cudaStream_t stream = 0;
    b2_transform Tsb{};
    bool has_model = false; uint32_t n = 0, width = 0, height = 0, n_origs = 1; float range_min = 0.f, range_max = 0.f;
    float max_dist = 1.0f, adaptive_max_dist_min = 0.15f;
    DevBuf<float> d_dirs, d_origs;
    DevBuf<float> d_dpts; DevBuf<uint8_t> d_dmask; uint32_t n_dataset = 0; DevBuf<float> d_ranges_in;
    DevBuf<float> d_mpts, d_mnrm, d_mranges; DevBuf<uint8_t> d_mhits; DevBuf<uint32_t> d_mfaces; uint32_t n_model = 0; bool found = false;
    DevBuf<double> d_partials; DevBuf<unsigned int> d_ticket; DevBuf<b2_cross_stats> d_stats; DevBuf<IcpState> d_icp;
    DevBuf<b2_transform> d_poses, d_tdelta; DevBuf<uint32_t> d_ncorr; DevBuf<b2_cross_stats> d_bstats;
    HostPin* pin = nullptr;
    int red_grid = 0;
    int fused_grid = 0;                 // blocks of k_icp_loop, one per SM (0: a whole-grid barrier is not available on this device)
    bool pdl_next = false, pdl_armed = false;   // the next find is followed by k_icp_loop launched with programmatic stream serialization / the find let it start early
    DevBuf<unsigned int> d_bar; unsigned int bar_base = 0;      // arrival counter + abort word of the software grid barrier; counter value at the next launch
    unsigned int seq = 0;               // completion sequence number written by k_icp_loop into pin->flag
    bool timing = false; cudaEvent_t ev[3] = {nullptr, nullptr, nullptr}; bool timing_valid = false;
    cudaStream_t aux = nullptr; cudaEvent_t ev_aux = nullptr;     // side stream: scan upload + unpack overlap the find kernel
    uint32_t n_ranges_in = 0;           // real ranges resident in d_ranges_in (set_ranges / correct_once_ranges), needed by b2_rcc_segment
    DevBuf<uint32_t> d_seg_counts, d_seg_offsets, d_seg_totals; DevBuf<float> d_seg_scan, d_seg_map; DevBuf<uint8_t> d_seg_labels;
    int corr_type = B2_CORR_RCC;        // B2_CORR_CPC: find() is a closest-point query per dataset point (CPCEmbree), no sensor model needed
    uint32_t work_n() const { return corr_type == B2_CORR_CPC ? n_dataset : n; }   // correspondences per find
};

static b2_transform tf_identity_pod() { b2_transform T; memset(&T, 0, sizeof(T)); T.R.w = 1.0f; return T; }

extern "C" int b2_rcc_destroy(b2_rcc* h);

// device resources of a fresh handle; any failure leaves the handle in a state b2_rcc_destroy can clean up
static int rcc_init(b2_rcc* h)
{
    b2_mesh* map = h->map;
    cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, map->device));
    h->red_grid = prop.multiProcessorCount;
    {
        int coop = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, map->device);
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_icp_loop<true>, B2_ICP_BLOCK, 0) == cudaSuccess && per_sm > 0)
            h->fused_grid = prop.multiProcessorCount;      // one block per SM
        (void)cudaGetLastError();
    }
    RES(h->d_partials.reserve((size_t)(B2_NACC + 1) * h->red_grid)); RES(h->d_ticket.reserve(1)); RES(h->d_stats.reserve(1)); RES(h->d_icp.reserve(1)); RES(h->d_bar.reserve(2));
    CU(cudaMemset(h->d_ticket.p, 0, sizeof(unsigned int)));
    CU(cudaMemset(h->d_bar.p, 0, 2 * sizeof(unsigned int)));
    CU(cudaHostAlloc((void**)&h->pin, sizeof(HostPin), cudaHostAllocMapped));
    memset((void*)h->pin, 0, sizeof(HostPin));
    CU(cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->ev_aux, cudaEventDisableTiming));
    return B2_OK;
}

extern "C" int b2_rcc_create(b2_mesh* map, b2_rcc** out)
{
    NOTNULL(out); *out = nullptr;
    if (!map) return fail(B2_ERR_NO_MAP, "NO MAP");
    CU(cudaSetDevice(map->device));
    b2_rcc* h = new (std::nothrow) b2_rcc();
    if (!h) return fail(B2_ERR_OOM, "out of host memory");
    h->map = map; h->Tsb = tf_identity_pod();
    map->refs.fetch_add(1);             // released in b2_rcc_destroy
    const int rc = rcc_init(h);
    if (rc != B2_OK) { b2_rcc_destroy(h); return rc; }      // the error text of the failing call stays in b2_last_error
    *out = h;
    return B2_OK;
}

extern "C" int b2_rcc_destroy(b2_rcc* h)
{
    if (!h) return B2_OK;
    cudaSetDevice(h->map->device);
    cudaStreamSynchronize(h->stream);
    h->d_dirs.release(); h->d_origs.release(); h->d_dpts.release(); h->d_dmask.release(); h->d_ranges_in.release();
    h->d_mpts.release(); h->d_mnrm.release(); h->d_mranges.release(); h->d_mhits.release(); h->d_mfaces.release();
    h->d_partials.release(); h->d_ticket.release(); h->d_stats.release(); h->d_icp.release(); h->d_bar.release();
    h->d_poses.release(); h->d_tdelta.release(); h->d_ncorr.release(); h->d_bstats.release();
    if (h->pin) cudaFreeHost(h->pin);
    if (h->aux) { cudaStreamSynchronize(h->aux); cudaStreamDestroy(h->aux); }
    if (h->ev_aux) cudaEventDestroy(h->ev_aux);
    for (int i = 0; i < 3; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    h->d_seg_counts.release(); h->d_seg_offsets.release(); h->d_seg_totals.release(); h->d_seg_scan.release(); h->d_seg_map.release(); h->d_seg_labels.release();
    b2_mesh* map = h->map;
    delete h;
    (void)cudaGetLastError();
    mesh_unref(map);
    return B2_OK;
}

extern "C" int b2_rcc_enable_timing(b2_rcc* h, int enable)
{
    NOTNULL(h);
    CU(cudaSetDevice(h->map->device));
    if (enable) for (int i = 0; i < 3; i++) if (!h->ev[i]) CU(cudaEventCreate(&h->ev[i]));
    h->timing = enable != 0; h->timing_valid = false;
    return B2_OK;
}
extern "C" int b2_rcc_last_timing(b2_rcc* h, float* find_ms, float* reduce_ms)
{
    NOTNULL(h);
    if (!h->timing || !h->timing_valid) return fail(B2_ERR_INVALID, "no timing recorded (b2_rcc_enable_timing + a correct_once call first)");
    CU(cudaSetDevice(h->map->device));
    CU(cudaEventSynchronize(h->ev[2]));
    float a = 0.f, b = 0.f;
    CU(cudaEventElapsedTime(&a, h->ev[0], h->ev[1])); CU(cudaEventElapsedTime(&b, h->ev[1], h->ev[2]));
    if (find_ms) *find_ms = a; if (reduce_ms) *reduce_ms = b;
    return B2_OK;
}

// profiling aid (not part of the public header): SM-clock durations of the last reduction's phases
extern "C" __attribute__((visibility("default"))) int b2_rcc_debug_clocks(b2_rcc* h, unsigned long long* out8)
{
    NOTNULL(h); NOTNULL(out8);
    for (int i = 0; i < 8; i++) out8[i] = h->pin->icp.dbg[i];
    return B2_OK;
}

// profiling aid (not part of the public header): per-warp {start, end} %globaltimer stamps of the NEXT k_rcc_find launches.
// buf_dev: device buffer of 2 x ceil(n_rays / 32) u64, or nullptr to switch the stamps off again.
extern "C" __attribute__((visibility("default"))) int b2_rcc_debug_find_warp_times(b2_rcc* h, unsigned long long* buf_dev)
{
    NOTNULL(h);
    CU(cudaSetDevice(h->map->device));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpyToSymbol(g_find_warp_times, &buf_dev, sizeof(buf_dev)));
    return B2_OK;
}

extern "C" int b2_rcc_set_stream(b2_rcc* h, void* s) { NOTNULL(h); h->stream = (cudaStream_t)s; return B2_OK; }
extern "C" int b2_rcc_set_tsb(b2_rcc* h, const b2_transform* Tsb) { NOTNULL(h); NOTNULL(Tsb); h->Tsb = *Tsb; return B2_OK; }
extern "C" int b2_rcc_set_params(b2_rcc* h, float max_dist, float amin) { NOTNULL(h); h->max_dist = max_dist; h->adaptive_max_dist_min = amin; return B2_OK; }

// upload sensor-frame ray tables.  Direction tables are evaluated on the host with libm cosf/sinf exactly like
// rmagine's SphericalModel::getDirection does on the CPU path (witness rmcl_ros/src/util/conversions.cpp:174-188), so the
// rays are bit-identical to the reference's; this runs once per setModel, not per scan.
static int set_model_tables(b2_rcc* h, uint32_t w, uint32_t hgt, const float* origs, uint32_t n_origs, const float* dirs, float rmin, float rmax)
{
    CU(cudaSetDevice(h->map->device));
    const size_t n = (size_t)w * hgt;
    if (n == 0) { h->has_model = true; h->n = 0; h->width = w; h->height = hgt; return B2_OK; }   // zero-size model: find() silently returns (RCCOptix.cpp:30-34)
    if (n > 0xffffffffu / 4) return fail(B2_ERR_INVALID, "model too large");
    RES(h->d_dirs.reserve(3 * n)); RES(h->d_origs.reserve(3 * (size_t)n_origs));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(h->d_dirs.p, dirs, sizeof(float) * 3 * n, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_origs.p, origs, sizeof(float) * 3 * (size_t)n_origs, cudaMemcpyHostToDevice));
    h->has_model = true; h->n = (uint32_t)n; h->width = w; h->height = hgt; h->n_origs = n_origs; h->range_min = rmin; h->range_max = rmax;
    return B2_OK;
}

extern "C" int b2_rcc_set_model_spherical(b2_rcc* h, const b2_spherical_model* m)
{
    NOTNULL(h); NOTNULL(m);
    const size_t n = (size_t)m->phi_size * m->theta_size;
    std::vector<float> dirs(3 * n);
    for (uint32_t vid = 0; vid < m->phi_size; vid++) {
        const float phi = m->phi_min + (float)vid * m->phi_inc;
        const float cp = cosf(phi), sp = sinf(phi);
        for (uint32_t hid = 0; hid < m->theta_size; hid++) {
            const float theta = m->theta_min + (float)hid * m->theta_inc;
            float* d = &dirs[3 * ((size_t)vid * m->theta_size + hid)];
            d[0] = cp * cosf(theta); d[1] = cp * sinf(theta); d[2] = sp;
        }
    }
    const float o[3] = {0.f, 0.f, 0.f};
    return set_model_tables(h, m->theta_size, m->phi_size, o, 1, dirs.data(), m->range_min, m->range_max);
}

extern "C" int b2_rcc_set_model_pinhole(b2_rcc* h, const b2_pinhole_model* m)
{
    NOTNULL(h); NOTNULL(m);
    const size_t n = (size_t)m->width * m->height;
    std::vector<float> dirs(3 * n);
    for (uint32_t vid = 0; vid < m->height; vid++)
        for (uint32_t hid = 0; hid < m->width; hid++) {
            const float px = ((float)hid - m->cx) / m->fx, py = ((float)vid - m->cy) / m->fy;
            const float nrm = sqrtf(px * px + py * py + 1.0f * 1.0f);
            const float ox = px / nrm, oy = py / nrm, oz = 1.0f / nrm;            // optical frame, normalised
            float* d = &dirs[3 * ((size_t)vid * m->width + hid)];
            d[0] = oz; d[1] = -ox; d[2] = -oy;                                     // x forward, y left, z up
        }
    const float o[3] = {0.f, 0.f, 0.f};
    return set_model_tables(h, m->width, m->height, o, 1, dirs.data(), m->range_min, m->range_max);
}

extern "C" int b2_rcc_set_model_o1dn(b2_rcc* h, uint32_t w, uint32_t hgt, const float orig[3], const float* dirs, float rmin, float rmax)
{
    NOTNULL(h); if ((size_t)w * hgt) { NOTNULL(orig); NOTNULL(dirs); }
    return set_model_tables(h, w, hgt, orig, 1, dirs, rmin, rmax);
}
extern "C" int b2_rcc_set_model_ondn(b2_rcc* h, uint32_t w, uint32_t hgt, const float* origs, const float* dirs, float rmin, float rmax)
{
    NOTNULL(h); if ((size_t)w * hgt) { NOTNULL(origs); NOTNULL(dirs); }
    return set_model_tables(h, w, hgt, origs, (uint32_t)((size_t)w * hgt), dirs, rmin, rmax);
}

extern "C" int b2_rcc_set_dataset(b2_rcc* h, const float* pts, const uint8_t* mask, uint32_t n, int src_is_device)
{
    NOTNULL(h);
    CU(cudaSetDevice(h->map->device));
    if (n) { NOTNULL(pts); }
    RES(h->d_dpts.reserve(3 * (size_t)std::max(n, 1u))); RES(h->d_dmask.reserve(std::max(n, 1u)));
    const cudaMemcpyKind kind = src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (n) {
        CU(cudaMemcpyAsync(h->d_dpts.p, pts, sizeof(float) * 3 * (size_t)n, kind, h->stream));
        if (mask) CU(cudaMemcpyAsync(h->d_dmask.p, mask, n, kind, h->stream));
        else CU(cudaMemsetAsync(h->d_dmask.p, 1, n, h->stream));                  // empty mask == all valid (statistics_p2l semantics)
        if (!src_is_device) CU(cudaStreamSynchronize(h->stream));                 // caller may reuse its buffers
    }
    h->n_dataset = n;
    return B2_OK;
}

static int ranges_to_dataset(b2_rcc* h, const float* ranges, uint32_t n, int src_is_device)
{
    if (!h->has_model) return fail(B2_ERR_INVALID, "set_ranges before setModel");
    if (n != h->n) return fail(B2_ERR_INVALID, "ranges size %u != model size %u", n, h->n);
    if (n == 0) { h->n_dataset = 0; return B2_OK; }
    NOTNULL(ranges);
    RES(h->d_dpts.reserve(3 * (size_t)n)); RES(h->d_dmask.reserve(n)); RES(h->d_ranges_in.reserve(n));
    CU(cudaMemcpyAsync(h->d_ranges_in.p, ranges, sizeof(float) * n, src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    k_dataset_from_ranges<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_ranges_in.p, h->d_dirs.p, h->d_origs.p, h->n_origs, n, h->range_min, h->range_max, h->d_dpts.p, h->d_dmask.p);
    LAUNCHED();
    h->n_dataset = n; h->n_ranges_in = n;
    return B2_OK;
}

extern "C" int b2_rcc_set_ranges(b2_rcc* h, const float* ranges, uint32_t n, int src_is_device)
{
    NOTNULL(h);
    CU(cudaSetDevice(h->map->device));
    RES(ranges_to_dataset(h, ranges, n, src_is_device));
    if (!src_is_device) CU(cudaStreamSynchronize(h->stream));
    return B2_OK;
}

static int reserve_model(b2_rcc* h, size_t n)
{
    // buffers only ever grow (RCCEmbree.cpp:28-33)
    RES(h->d_mpts.reserve(3 * n)); RES(h->d_mnrm.reserve(3 * n)); RES(h->d_mranges.reserve(n)); RES(h->d_mhits.reserve(n)); RES(h->d_mfaces.reserve(n));
    return B2_OK;
}

static RayModel ray_model(const b2_rcc* h)
{
    RayModel m; m.dirs = h->d_dirs.p; m.origs = h->d_origs.p; m.n_origs = h->n_origs; m.n = h->n; m.range_min = h->range_min; m.range_max = h->range_max;
    m.width = h->width; m.height = h->height; return m;
}
static ModelBuffers model_buffers(const b2_rcc* h)
{
    ModelBuffers b; b.pts = h->d_mpts.p; b.nrm = h->d_mnrm.p; b.hits = h->d_mhits.p; b.faces = h->d_mfaces.p; b.ranges = h->d_mranges.p; return b;
}

static int launch_find(b2_rcc* h, const b2_transform* Tbm_host, const IcpState* icp_dev)
{
    static const int prefetch_mode_cp = [] { const char* e = getenv("B2_FIND_PREFETCH"); return e ? atoi(e) : 1; }();
    if (h->corr_type == B2_CORR_CPC) {
        // CPCEmbree::find (CPCEmbree.cpp:17-43): one closest-point query per dataset point
        const uint32_t n = h->n_dataset;
        if (n == 0) { h->n_model = 0; h->found = true; return B2_OK; }
        RES(reserve_model(h, n));
        k_cpc_find<<<(n + B2_FIND_BLOCK - 1) / B2_FIND_BLOCK, B2_FIND_BLOCK, 0, h->stream>>>(h->map->view(), h->map->n_nodes, h->map->n_tris, prefetch_mode_cp, icp_dev,
                                                                                             Tbm_host ? *Tbm_host : tf_identity_pod(), h->Tsb, h->d_dpts.p, n, h->max_dist, model_buffers(h));
        LAUNCHED();
        h->n_model = n; h->found = true;
        return B2_OK;
    }
    if (!h->has_model) return fail(B2_ERR_INVALID, "find before setModel");
    if (h->n == 0) return B2_OK;
    RES(reserve_model(h, h->n));
    const uint32_t grid = (h->n + B2_FIND_BLOCK - 1) / B2_FIND_BLOCK;
    static const int prefetch_mode = [] { const char* e = getenv("B2_FIND_PREFETCH"); return e ? atoi(e) : 1; }();
    // early_dependents: only when every block of this grid is resident in the first wave (14 blocks per SM), so that an early-resident
    // dependent block can never take an SM slot a find block is still waiting for
    const int early = (h->pdl_next && grid <= 14u * (uint32_t)h->red_grid) ? 1 : 0;
    h->pdl_armed = early != 0;
    k_rcc_find<<<grid, B2_FIND_BLOCK, 0, h->stream>>>(h->map->view(), h->map->n_nodes, h->map->n_tris, prefetch_mode, nullptr, icp_dev, Tbm_host ? *Tbm_host : tf_identity_pod(), h->Tsb, ray_model(h), 1u, model_buffers(h), early);
    LAUNCHED();
    h->n_model = h->n; h->found = true;
    return B2_OK;
}

extern "C" int b2_rcc_set_correspondence_type(b2_rcc* h, int type)
{
    NOTNULL(h);
    if (type != B2_CORR_RCC && type != B2_CORR_CPC) return fail(B2_ERR_INVALID, "unknown correspondence type %d", type);
    h->corr_type = type; h->found = false;
    return B2_OK;
}

extern "C" int b2_rcc_find(b2_rcc* h, const b2_transform* Tbm)
{
    NOTNULL(h); NOTNULL(Tbm);
    CU(cudaSetDevice(h->map->device));
    return launch_find(h, Tbm, nullptr);
}

static int launch_reduce(b2_rcc* h, const b2_transform* Tpre_host, float max_dist, IcpState* icp_dev, b2_cross_stats* out_dev)
{
    const uint32_t n = std::min(h->n_dataset, h->n_model);
    int grid = (int)std::min<uint32_t>((uint32_t)h->red_grid, (n + B2_RED_BLOCK - 1) / B2_RED_BLOCK);
    if (grid < 1) grid = 1;
    k_p2l_reduce<<<grid, B2_RED_BLOCK, 0, h->stream>>>(h->d_dpts.p, h->d_dmask.p, h->d_mpts.p, h->d_mnrm.p, h->d_mhits.p, n,
                                                       Tpre_host ? *Tpre_host : tf_identity_pod(), max_dist, icp_dev, h->d_partials.p, h->d_ticket.p, out_dev);
    LAUNCHED();
    return B2_OK;
}

static float adaptive_max_dist(const b2_rcc* h, double cp)
{
    // CorrespondencesCPU.cpp:21-23
    return (float)(h->max_dist * (1.0 - cp) + h->adaptive_max_dist_min * cp);
}

extern "C" int b2_rcc_cross_statistics(b2_rcc* h, const b2_transform* T, double cp, b2_cross_stats* out)
{
    NOTNULL(h); NOTNULL(T); NOTNULL(out);
    CU(cudaSetDevice(h->map->device));
    if (!h->found) return fail(B2_ERR_INVALID, "computeCrossStatistics before find");
    if (h->n_dataset == 0) return fail(B2_ERR_INVALID, "computeCrossStatistics without a dataset");
    RES(launch_reduce(h, T, adaptive_max_dist(h, cp), nullptr, h->d_stats.p));
    CU(cudaMemcpyAsync(&h->pin->S[0], h->d_stats.p, sizeof(b2_cross_stats), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *out = h->pin->S[0];
    return B2_OK;
}

extern "C" int b2_rcc_segment(b2_rcc* h, float min_dist_outlier_scan, float min_dist_outlier_map, float* outlier_scan, uint32_t cap_scan, uint32_t* n_scan,
                              float* outlier_map, uint32_t cap_map, uint32_t* n_map, uint8_t* labels)
{
    NOTNULL(h); NOTNULL(n_scan); NOTNULL(n_map);
    *n_scan = 0; *n_map = 0;
    CU(cudaSetDevice(h->map->device));
    if (h->corr_type != B2_CORR_RCC || !h->has_model) return fail(B2_ERR_INVALID, "segmentation needs a ray-casting handle with a sensor model");
    if (!h->found || h->n_model != h->n) return fail(B2_ERR_INVALID, "segmentation before find");
    if (h->n_ranges_in != h->n) return fail(B2_ERR_INVALID, "segmentation needs the real ranges (set_ranges), have %u of %u", h->n_ranges_in, h->n);
    const uint32_t n = h->n;
    if (n == 0) return B2_OK;
    const uint32_t blocks = (n + B2_SEG_BLOCK - 1) / B2_SEG_BLOCK;
    RES(h->d_seg_counts.reserve(2 * (size_t)blocks)); RES(h->d_seg_offsets.reserve(2 * (size_t)blocks)); RES(h->d_seg_totals.reserve(2));
    RES(h->d_seg_scan.reserve(3 * (size_t)n)); RES(h->d_seg_map.reserve(3 * (size_t)n)); RES(h->d_seg_labels.reserve(n));
    const RayModel m = ray_model(h);
    k_segment<<<blocks, B2_SEG_BLOCK, 0, h->stream>>>(m, h->d_ranges_in.p, h->d_mranges.p, h->d_mnrm.p, min_dist_outlier_scan, min_dist_outlier_map, h->d_seg_counts.p, nullptr, nullptr, nullptr, nullptr);
    LAUNCHED();
    k_segment_scan<<<1, 1024, 0, h->stream>>>(h->d_seg_counts.p, blocks, h->d_seg_offsets.p, h->d_seg_totals.p);
    LAUNCHED();
    k_segment<<<blocks, B2_SEG_BLOCK, 0, h->stream>>>(m, h->d_ranges_in.p, h->d_mranges.p, h->d_mnrm.p, min_dist_outlier_scan, min_dist_outlier_map, h->d_seg_counts.p, h->d_seg_offsets.p,
                                                      h->d_seg_scan.p, h->d_seg_map.p, h->d_seg_labels.p);
    LAUNCHED();
    uint32_t tot[2] = {0, 0};
    CU(cudaMemcpyAsync(tot, h->d_seg_totals.p, sizeof(tot), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    *n_scan = tot[0]; *n_map = tot[1];
    if (outlier_scan && std::min(tot[0], cap_scan)) CU(cudaMemcpyAsync(outlier_scan, h->d_seg_scan.p, sizeof(float) * 3 * (size_t)std::min(tot[0], cap_scan), cudaMemcpyDeviceToHost, h->stream));
    if (outlier_map && std::min(tot[1], cap_map)) CU(cudaMemcpyAsync(outlier_map, h->d_seg_map.p, sizeof(float) * 3 * (size_t)std::min(tot[1], cap_map), cudaMemcpyDeviceToHost, h->stream));
    if (labels) CU(cudaMemcpyAsync(labels, h->d_seg_labels.p, n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return B2_OK;
}

extern "C" int b2_rcc_model_view(b2_rcc* h, float** p, float** nr, uint8_t** hi, uint32_t** f, float** r, uint32_t* n)
{
    NOTNULL(h);
    if (p) *p = h->d_mpts.p; if (nr) *nr = h->d_mnrm.p; if (hi) *hi = h->d_mhits.p; if (f) *f = h->d_mfaces.p; if (r) *r = h->d_mranges.p; if (n) *n = h->n_model;
    return B2_OK;
}
extern "C" int b2_rcc_dataset_view(b2_rcc* h, float** p, uint8_t** m, uint32_t* n)
{
    NOTNULL(h);
    if (p) *p = h->d_dpts.p; if (m) *m = h->d_dmask.p; if (n) *n = h->n_dataset;
    return B2_OK;
}
extern "C" int b2_rcc_download_model(b2_rcc* h, float* p, float* nr, uint8_t* hi, uint32_t* f, float* r)
{
    NOTNULL(h);
    CU(cudaSetDevice(h->map->device));
    const size_t n = h->n_model;
    CU(cudaStreamSynchronize(h->stream));
    if (n == 0) return B2_OK;
    if (p) CU(cudaMemcpy(p, h->d_mpts.p, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost));
    if (nr) CU(cudaMemcpy(nr, h->d_mnrm.p, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost));
    if (hi) CU(cudaMemcpy(hi, h->d_mhits.p, n, cudaMemcpyDeviceToHost));
    if (f) CU(cudaMemcpy(f, h->d_mfaces.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    if (r) CU(cudaMemcpy(r, h->d_mranges.p, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return B2_OK;
}
extern "C" int b2_rcc_download_dataset(b2_rcc* h, float* p, uint8_t* m)
{
    NOTNULL(h);
    CU(cudaSetDevice(h->map->device));
    const size_t n = h->n_dataset;
    CU(cudaStreamSynchronize(h->stream));
    if (n == 0) return B2_OK;
    if (p) CU(cudaMemcpy(p, h->d_dpts.p, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost));
    if (m) CU(cudaMemcpy(m, h->d_dmask.p, n, cudaMemcpyDeviceToHost));
    return B2_OK;
}

static int correct_once_impl(b2_rcc* h, const b2_transform* Tom, const b2_transform* Tbo, uint32_t iterations, double cp,
                             b2_transform* Tom_new, b2_transform* T_onew_oold, b2_cross_stats* Cmerged, const float* ranges_host = nullptr, uint32_t n_ranges = 0)
{
    bool aux_used = false;
    const bool cpc = h->corr_type == B2_CORR_CPC;
    if (ranges_host && cpc) return fail(B2_ERR_INVALID, "correctOnce(ranges) needs ray-casting correspondences (the handle is in closest-point mode)");
    static const int use_coop = [] { const char* e = getenv("B2_FUSED"); return e ? atoi(e) : 2; }();
    static const int use_zc = [] { const char* e = getenv("B2_ZEROCOPY"); return e ? atoi(e) : 1; }();
    // Zero-copy scan: when the caller's buffer is pinned host memory the ICP-loop kernel reads it directly (and unpacks it) instead of
    // memcpy + unpack kernel + cross-stream event.  Needs the register-cached loop (<= 2 pairs per thread) and the fused path.
    const float* zc_ranges = nullptr;
    if (ranges_host) {
        if (!h->has_model) return fail(B2_ERR_INVALID, "set_ranges before setModel");
        if (n_ranges != h->n) return fail(B2_ERR_INVALID, "ranges size %u != model size %u", n_ranges, h->n);
        if (use_zc && use_coop && h->fused_grid > 0 && iterations > 0 && h->n > 0 && h->n <= 2u * (uint32_t)h->fused_grid * B2_ICP_BLOCK) {
            // asked on every call (about a microsecond): an address can change from pinned to pageable between calls
            cudaPointerAttributes pa;
            if (cudaPointerGetAttributes(&pa, ranges_host) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer) zc_ranges = (const float*)pa.devicePointer;
            (void)cudaGetLastError();
        }
        if (h->n > 0) {
            RES(h->d_dpts.reserve(3 * (size_t)h->n)); RES(h->d_dmask.reserve(h->n)); RES(h->d_ranges_in.reserve(h->n));
            if (!zc_ranges) {
                CU(cudaEventRecord(h->ev_aux, h->stream));                // the side stream starts after whatever the main stream had in flight BEFORE this call
                CU(cudaStreamWaitEvent(h->aux, h->ev_aux, 0));
            }
        }
        h->n_dataset = h->n; h->n_ranges_in = h->n;
    }
    // the find kernel does not read the dataset: the scan is uploaded + unpacked on the side stream WHILE it runs (and the host-side
    // cost of issuing the copy is hidden behind the already launched find)
    auto upload_scan = [&]() -> int {
        if (!ranges_host || h->n == 0 || zc_ranges) return B2_OK;
        CU(cudaMemcpyAsync(h->d_ranges_in.p, ranges_host, sizeof(float) * h->n, cudaMemcpyHostToDevice, h->aux));
        k_dataset_from_ranges<<<(h->n + 255) / 256, 256, 0, h->aux>>>(h->d_ranges_in.p, h->d_dirs.p, h->d_origs.p, h->n_origs, h->n, h->range_min, h->range_max,
                                                                      h->d_dpts.p, h->d_dmask.p);
        LAUNCHED();
        CU(cudaEventRecord(h->ev_aux, h->aux));
        aux_used = true;
        return B2_OK;
    };
    if (!cpc && !h->has_model) return fail(B2_ERR_INVALID, "correctOnce before setModel");
    const uint32_t nw = h->work_n();
    if (h->n_dataset != nw) return fail(B2_ERR_INVALID, "dataset size %u != model size %u", h->n_dataset, h->n);
    IcpState& st = h->pin->icp;
    memset(&st, 0, sizeof(st));
    st.Tom = *Tom; st.Tbo = *Tbo; st.Tsb = h->Tsb; st.max_dist = adaptive_max_dist(h, cp);
    st.T_onew_oold = tf_identity_pod(); st.Tom_new = *Tom;
    // pre-transform of the first reduction, evaluated on the host with the same inline functions the kernels use (individually
    // rounded ops on both sides -> identical bits); saves a launch
    tf_store(&st.T_snew_sold, icp_pretransform(tf_from_pod(*Tbo), tf_from_pod(h->Tsb), tf_identity()));
    {
        const Tf Tos = tf_mul(tf_from_pod(*Tbo), tf_from_pod(h->Tsb));
        tf_store(&st.Tos, Tos); tf_store(&st.Tso, tf_inv(Tos)); quat_to_mat(Tos.R, st.Ros);
    }
    bool barrier_used = false;
    static const int use_spin = [] { const char* e = getenv("B2_SPIN"); return e ? atoi(e) : 1; }();
    bool waited = false;
    if (nw > 0 && use_coop && h->fused_grid > 0 && iterations > 0) {
        // find, then ALL inner iterations in one cooperative kernel; state in by kernel parameter, result out through mapped pinned memory
        if (h->timing) CU(cudaEventRecord(h->ev[0], h->stream));
        b2_transform Tbm_host; memset(&Tbm_host, 0, sizeof(Tbm_host));
        tf_store(&Tbm_host, tf_mul(tf_from_pod(*Tom), tf_from_pod(*Tbo)));        // MICPSensor.hpp:148, same inline ops as the kernels
        static const int use_pdl = [] { const char* e = getenv("B2_PDL"); return e ? atoi(e) : 1; }();
        h->pdl_next = use_coop == 2 && use_pdl && !h->timing && h->corr_type == B2_CORR_RCC && (!ranges_host || zc_ranges);      // event records between the two kernels would serialise them anyway
        const int rc_find = launch_find(h, &Tbm_host, nullptr);
        h->pdl_next = false;
        RES(rc_find);
        if (h->timing) CU(cudaEventRecord(h->ev[1], h->stream));
        RES(upload_scan());
        if (aux_used) CU(cudaStreamWaitEvent(h->stream, h->ev_aux, 0));
        int grid = std::min<int>(h->fused_grid, (int)((nw + B2_ICP_BLOCK - 1) / B2_ICP_BLOCK));
        if (grid < 1) grid = 1;
        RES(h->d_partials.reserve((size_t)2 * grid * (B2_NACC + 1)));
        const float* dp = h->d_dpts.p; const uint8_t* dmk = h->d_dmask.p; const float* mp = h->d_mpts.p; const float* mn = h->d_mnrm.p; const uint8_t* mh = h->d_mhits.p;
        uint32_t nel = nw; IcpState* icp_dev = h->d_icp.p; uint32_t its = iterations; double* parts = h->d_partials.p;
        IcpState* host_out = use_spin ? &h->pin->icp_out : nullptr; volatile unsigned int* host_flag = use_spin ? &h->pin->flag : nullptr;
        unsigned int seq = ++h->seq; if (seq == 0) seq = ++h->seq;
        // B2_FUSED=2 (default): ordinary launch + software grid barrier (one block per SM, all resident) -- measured ~10 us less launch
        // overhead per step than the cooperative launch (B2_FUSED=1), which stays available
        unsigned int* bar = h->d_bar.p; unsigned int bar_base = h->bar_base; unsigned int* bar_abort = h->d_bar.p + 1;
        RayModel zc_model = ray_model(h); float* zc_dpts = h->d_dpts.p; uint8_t* zc_dmask = h->d_dmask.p; float* zc_rin = h->d_ranges_in.p;
        if (use_coop == 1) {
            void* args[] = {&dp, &dmk, &mp, &mn, &mh, &nel, &icp_dev, &its, &parts, &st, &host_out, &host_flag, &seq, &bar, &bar_base, &bar_abort, &zc_ranges, &zc_model, &zc_dpts, &zc_dmask, &zc_rin};
            CU(cudaLaunchCooperativeKernel((const void*)k_icp_loop<true>, dim3(grid), dim3(B2_ICP_BLOCK), args, 0, h->stream));
        } else {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(B2_ICP_BLOCK); cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = (h->pdl_armed && !aux_used) ? 1 : 0;      // with a scan upload in flight the kernel also waits on the side stream's event
            cfg.attrs = attr; cfg.numAttrs = 1;
            CU(cudaLaunchKernelEx(&cfg, k_icp_loop<false>, dp, dmk, mp, mn, mh, nel, icp_dev, its, parts, st, host_out, host_flag, seq, bar, bar_base, bar_abort, zc_ranges, zc_model, zc_dpts, zc_dmask, zc_rin));
            h->bar_base += its * (unsigned int)grid;
            barrier_used = true;
        }
        LAUNCHED();
        if (h->timing) { CU(cudaEventRecord(h->ev[2], h->stream)); h->timing_valid = true; }
        if (use_spin) {
            // spin on the completion flag the kernel writes into mapped host memory (a stream synchronise costs several microseconds more)
            const auto t_start = std::chrono::steady_clock::now();
            unsigned long long spins = 0;
            while (h->pin->flag != seq) {
                if ((++spins & 0xfffffull) == 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > 5.0) break;
            }
            if (h->pin->flag == seq) { memcpy(&st, (const void*)&h->pin->icp_out, sizeof(IcpState)); waited = true; }
        }
    } else if (nw > 0) {
        RES(upload_scan());
        if (aux_used) CU(cudaStreamWaitEvent(h->stream, h->ev_aux, 0));
        CU(cudaMemcpyAsync(h->d_icp.p, &st, sizeof(IcpState), cudaMemcpyHostToDevice, h->stream));
        if (h->timing) CU(cudaEventRecord(h->ev[0], h->stream));
        RES(launch_find(h, nullptr, h->d_icp.p));
        if (h->timing) CU(cudaEventRecord(h->ev[1], h->stream));
        for (uint32_t it = 0; it < iterations; it++) RES(launch_reduce(h, nullptr, 0.f, h->d_icp.p, nullptr));
        if (h->timing) { CU(cudaEventRecord(h->ev[2], h->stream)); h->timing_valid = true; }
    } else {
        CU(cudaMemcpyAsync(h->d_icp.p, &st, sizeof(IcpState), cudaMemcpyHostToDevice, h->stream));
    }
    if (!waited) {
        CU(cudaMemcpyAsync(&st, h->d_icp.p, sizeof(IcpState), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        if (barrier_used) {
            // the completion flag never arrived: either the spin is switched off, or the grid barrier gave up (blocks not co-resident)
            unsigned int bar_state[2] = {0u, 0u};
            CU(cudaMemcpy(bar_state, h->d_bar.p, sizeof(bar_state), cudaMemcpyDeviceToHost));
            if (bar_state[1] != 0u) {
                CU(cudaMemset(h->d_bar.p, 0, 2 * sizeof(unsigned int))); h->bar_base = 0;
                return fail(B2_ERR_CUDA, "correctOnce: the grid barrier of the ICP loop timed out (blocks not co-resident); set B2_FUSED=1 or 0");
            }
        }
    }
    if (Tom_new) *Tom_new = st.Tom_new;
    if (T_onew_oold) *T_onew_oold = st.T_onew_oold;
    if (Cmerged) *Cmerged = st.Cmerged_o;
    return B2_OK;
}

extern "C" int b2_rcc_correct_once(b2_rcc* h, const b2_transform* Tom, const b2_transform* Tbo, uint32_t iterations, double cp,
                                   b2_transform* Tom_new, b2_transform* T_onew_oold, b2_cross_stats* Cmerged)
{
    NOTNULL(h); NOTNULL(Tom); NOTNULL(Tbo);
    CU(cudaSetDevice(h->map->device));
    return correct_once_impl(h, Tom, Tbo, iterations, cp, Tom_new, T_onew_oold, Cmerged);
}

extern "C" int b2_rcc_correct_once_ranges(b2_rcc* h, const float* ranges, uint32_t n, const b2_transform* Tom, const b2_transform* Tbo, uint32_t iterations,
                                          double cp, b2_transform* Tom_new, b2_transform* T_onew_oold, b2_cross_stats* Cmerged)
{
    NOTNULL(h); NOTNULL(Tom); NOTNULL(Tbo);
    CU(cudaSetDevice(h->map->device));
    if (n) NOTNULL(ranges);
    return correct_once_impl(h, Tom, Tbo, iterations, cp, Tom_new, T_onew_oold, Cmerged, ranges, n);
}

extern "C" int b2_rcc_correct_batch(b2_rcc* h, const b2_transform* Tbm, uint32_t n_poses, int poses_on_device,
                                    b2_transform* Tdelta, uint32_t* ncorr, b2_cross_stats* stats_b, int out_on_device)
{
    NOTNULL(h);
    CU(cudaSetDevice(h->map->device));
    if (!h->has_model) return fail(B2_ERR_INVALID, "correct before setModel");
    if (n_poses == 0 || h->n == 0) return B2_OK;
    NOTNULL(Tbm);
    if (h->n_dataset != h->n) return fail(B2_ERR_INVALID, "correct before setInputData (dataset %u != model %u)", h->n_dataset, h->n);
    const b2_transform* poses_dev = Tbm;
    if (!poses_on_device) {
        RES(h->d_poses.reserve(n_poses));
        CU(cudaMemcpyAsync(h->d_poses.p, Tbm, sizeof(b2_transform) * (size_t)n_poses, cudaMemcpyHostToDevice, h->stream));
        poses_dev = h->d_poses.p;
    }
    const uint32_t rays_per_block = B2_FUSED_BLOCK * 8;
    const uint32_t bpp = (h->n + rays_per_block - 1) / rays_per_block;
    const uint64_t grid = (uint64_t)bpp * n_poses;
    if (grid > 0x7fffffffull) return fail(B2_ERR_INVALID, "too many poses");
    RES(h->d_partials.reserve((size_t)std::max<uint64_t>(grid, (uint64_t)h->red_grid) * (B2_NACC + 1)));
    k_rcc_fused_batch<<<(uint32_t)grid, B2_FUSED_BLOCK, 0, h->stream>>>(h->map->view(), poses_dev, h->Tsb, ray_model(h), h->d_dpts.p, h->d_dmask.p, h->max_dist,
                                                                       bpp, rays_per_block, h->d_partials.p);
    LAUNCHED();
    b2_transform* td = Tdelta; uint32_t* nc = ncorr; b2_cross_stats* sb = stats_b;
    if (!out_on_device) {
        RES(h->d_tdelta.reserve(n_poses)); RES(h->d_ncorr.reserve(n_poses)); RES(h->d_bstats.reserve(n_poses));
        td = h->d_tdelta.p; nc = h->d_ncorr.p; sb = h->d_bstats.p;
    }
    k_umeyama_from_partials<<<(n_poses + 63) / 64, 64, 0, h->stream>>>(h->d_partials.p, bpp, n_poses, h->Tsb, td, nc, sb);
    LAUNCHED();
    if (!out_on_device) {
        if (Tdelta) CU(cudaMemcpyAsync(Tdelta, td, sizeof(b2_transform) * (size_t)n_poses, cudaMemcpyDeviceToHost, h->stream));
        if (ncorr) CU(cudaMemcpyAsync(ncorr, nc, sizeof(uint32_t) * (size_t)n_poses, cudaMemcpyDeviceToHost, h->stream));
        if (stats_b) CU(cudaMemcpyAsync(stats_b, sb, sizeof(b2_cross_stats) * (size_t)n_poses, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return B2_OK;
}

extern "C" int b2_umeyama_batch(const b2_cross_stats* stats, uint32_t n, b2_transform* out, int on_device, int device, void* stream_)
{
    if (n == 0) return B2_OK;
    NOTNULL(stats); NOTNULL(out);
    CU(cudaSetDevice(device));
    cudaStream_t stream = (cudaStream_t)stream_;
    if (on_device) {
        k_umeyama_batch<<<(n + 63) / 64, 64, 0, stream>>>(stats, n, out);
        LAUNCHED();
        return B2_OK;
    }
    DevBuf<b2_cross_stats> ds; DevBuf<b2_transform> dt;
    int rc;
    if ((rc = ds.reserve(n)) || (rc = dt.reserve(n))) { ds.release(); dt.release(); return rc; }
    cudaError_t e = cudaMemcpyAsync(ds.p, stats, sizeof(b2_cross_stats) * (size_t)n, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) { k_umeyama_batch<<<(n + 63) / 64, 64, 0, stream>>>(ds.p, n, dt.p); g_launches.fetch_add(1); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dt.p, sizeof(b2_transform) * (size_t)n, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    ds.release(); dt.release();
    if (e != cudaSuccess) return fail(B2_ERR_CUDA, "b2_umeyama_batch: %s", cudaGetErrorString(e));
    return B2_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// particle filter
// ---------------------------------------------------------------------------------------------------------------------
struct b2_pf {
    b2_mesh* map = nullptr; cudaStream_t stream = 0;
    DevBuf<PfBeam> d_beams; PfBeam* h_beams = nullptr; size_t h_beams_cap = 0;
    DevBuf<b2_transform> d_poses; DevBuf<b2_particle_attr> d_attrs;
    DevBuf<double> d_part; DevBuf<unsigned int> d_ticket; DevBuf<float> d_out; float* h_out = nullptr; int n_sm = 0;
    int smem_optin = 0;
};

extern "C" int b2_pf_destroy(b2_pf* h);

static int pf_init(b2_pf* h)
{
    b2_mesh* map = h->map;
    CU(cudaDeviceGetAttribute(&h->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, map->device));
    h->smem_optin -= 1024;              // room for the kernel's static shared memory
    CU(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, map->device));
    RES(h->d_part.reserve(2 * (size_t)h->n_sm * 4)); RES(h->d_ticket.reserve(1)); RES(h->d_out.reserve(2));
    CU(cudaMemset(h->d_ticket.p, 0, sizeof(unsigned int)));
    CU(cudaMallocHost((void**)&h->h_out, 2 * sizeof(float)));
    CU(cudaFuncSetAttribute(k_pf_update<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
    CU(cudaFuncSetAttribute(k_pf_update<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin));
    return B2_OK;
}

extern "C" int b2_pf_create(b2_mesh* map, b2_pf** out)
{
    NOTNULL(out); *out = nullptr;
    if (!map) return fail(B2_ERR_NO_MAP, "NO MAP");
    CU(cudaSetDevice(map->device));
    b2_pf* h = new (std::nothrow) b2_pf();
    if (!h) return fail(B2_ERR_OOM, "out of host memory");
    h->map = map;
    map->refs.fetch_add(1);             // released in b2_pf_destroy
    const int rc = pf_init(h);
    if (rc != B2_OK) { b2_pf_destroy(h); return rc; }
    *out = h;
    return B2_OK;
}
extern "C" int b2_pf_destroy(b2_pf* h)
{
    if (!h) return B2_OK;
    cudaSetDevice(h->map->device);
    cudaStreamSynchronize(h->stream);
    h->d_beams.release(); h->d_poses.release(); h->d_attrs.release(); h->d_part.release(); h->d_ticket.release(); h->d_out.release();
    if (h->h_beams) cudaFreeHost(h->h_beams);
    if (h->h_out) cudaFreeHost(h->h_out);
    b2_mesh* map = h->map;
    delete h;
    (void)cudaGetLastError();
    mesh_unref(map);
    return B2_OK;
}
extern "C" int b2_pf_set_stream(b2_pf* h, void* s) { NOTNULL(h); h->stream = (cudaStream_t)s; return B2_OK; }

static uint32_t dir_sort_key(const b2_range_meas& m)
{
    // Morton code of the direction on a 1024^3 lattice: neighbouring beams end up in the same warp (coherent traversal)
    auto q = [](float v) { int i = (int)((v * 0.5f + 0.5f) * 1023.0f); return (uint32_t)std::min(std::max(i, 0), 1023); };
    auto spread = [](uint32_t x) { x &= 0x3ff; x = (x | (x << 16)) & 0x30000ff; x = (x | (x << 8)) & 0x300f00f; x = (x | (x << 4)) & 0x30c30c3; x = (x | (x << 2)) & 0x9249249; return x; };
    return spread(q(m.dir.x)) | (spread(q(m.dir.y)) << 1) | (spread(q(m.dir.z)) << 2);
}

static int pf_update_impl(b2_pf* h, const b2_transform* poses_dev, b2_particle_attr* attrs_dev, uint32_t n, const b2_transform* Tsb,
                          const b2_range_meas* beams, uint32_t n_beams, const b2_pf_params* prm)
{
    if (n == 0 || n_beams == 0) return B2_OK;
    // beams: host -> compact, direction-sorted device table (merge order preserved through PfBeam::slot)
    if (h->h_beams_cap < n_beams) {
        if (h->h_beams) cudaFreeHost(h->h_beams);
        h->h_beams = nullptr; h->h_beams_cap = 0;
        CU(cudaMallocHost((void**)&h->h_beams, sizeof(PfBeam) * (size_t)n_beams));
        h->h_beams_cap = n_beams;
    }
    RES(h->d_beams.reserve(n_beams));
    CU(cudaStreamSynchronize(h->stream));                 // previous launch may still read the staging buffer's device copy
    std::vector<std::pair<uint32_t, uint32_t>> order(n_beams);
    for (uint32_t i = 0; i < n_beams; i++) order[i] = {dir_sort_key(beams[i]), i};
    std::sort(order.begin(), order.end());
    for (uint32_t j = 0; j < n_beams; j++) {
        const b2_range_meas& m = beams[order[j].second];
        PfBeam& b = h->h_beams[j];
        b.ox = m.orig.x; b.oy = m.orig.y; b.oz = m.orig.z; b.dx = m.dir.x; b.dy = m.dir.y; b.dz = m.dir.z; b.range = m.range; b.slot = order[j].second;
    }
    CU(cudaMemcpyAsync(h->d_beams.p, h->h_beams, sizeof(PfBeam) * (size_t)n_beams, cudaMemcpyHostToDevice, h->stream));
    // particles per block: as many as fit the shared-memory evaluation tile, capped so that a block still has enough rays
    const size_t bytes_per_particle = sizeof(float) * (size_t)n_beams;
    if (bytes_per_particle > (size_t)h->smem_optin) return fail(B2_ERR_UNSUPPORTED, "too many beams per update (%u)", n_beams);
    uint32_t ppb = (uint32_t)std::min<size_t>((size_t)h->smem_optin / bytes_per_particle, 64);
    const uint32_t want = std::max(1u, (B2_PF_BLOCK * 8 + n_beams - 1) / n_beams);       // ~8 rays per thread
    ppb = std::max(1u, std::min(ppb, want));
    ppb = std::min(ppb, (uint32_t)B2_PF_BLOCK);
    const uint32_t grid = (n + ppb - 1) / ppb;
    if (prm->correspondence_type == 1) k_pf_update<1><<<grid, B2_PF_BLOCK, bytes_per_particle * ppb, h->stream>>>(h->map->view(), poses_dev, attrs_dev, n, *Tsb, h->d_beams.p, n_beams, *prm, ppb);
    else                               k_pf_update<0><<<grid, B2_PF_BLOCK, bytes_per_particle * ppb, h->stream>>>(h->map->view(), poses_dev, attrs_dev, n, *Tsb, h->d_beams.p, n_beams, *prm, ppb);
    LAUNCHED();
    return B2_OK;
}

extern "C" int b2_pf_sensor_update(b2_pf* h, const b2_transform* poses_dev, b2_particle_attr* attrs_dev, uint32_t n, const b2_transform* Tsb,
                                   const b2_range_meas* beams, uint32_t n_beams, const b2_pf_params* prm)
{
    NOTNULL(h); NOTNULL(Tsb); NOTNULL(prm);
    if (n) { NOTNULL(poses_dev); NOTNULL(attrs_dev); }
    if (n_beams) NOTNULL(beams);
    CU(cudaSetDevice(h->map->device));
    return pf_update_impl(h, poses_dev, attrs_dev, n, Tsb, beams, n_beams, prm);
}

extern "C" int b2_pf_sensor_update_host(b2_pf* h, const b2_transform* poses, b2_particle_attr* attrs, uint32_t n, const b2_transform* Tsb,
                                        const b2_range_meas* beams, uint32_t n_beams, const b2_pf_params* prm)
{
    NOTNULL(h); NOTNULL(Tsb); NOTNULL(prm);
    if (n == 0) return B2_OK;
    NOTNULL(poses); NOTNULL(attrs);
    if (n_beams) NOTNULL(beams);
    CU(cudaSetDevice(h->map->device));
    RES(h->d_poses.reserve(n)); RES(h->d_attrs.reserve(n));
    CU(cudaMemcpyAsync(h->d_poses.p, poses, sizeof(b2_transform) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_attrs.p, attrs, sizeof(b2_particle_attr) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    RES(pf_update_impl(h, h->d_poses.p, h->d_attrs.p, n, Tsb, beams, n_beams, prm));
    CU(cudaMemcpyAsync(attrs, h->d_attrs.p, sizeof(b2_particle_attr) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return B2_OK;
}

extern "C" int b2_pf_motion_update(b2_pf* h, b2_transform* poses_dev, b2_particle_attr* attrs_dev, uint32_t n, const b2_transform* T, double forget_rate, int check_collision)
{
    NOTNULL(h); NOTNULL(T);
    if (n == 0) return B2_OK;
    NOTNULL(poses_dev); NOTNULL(attrs_dev);
    CU(cudaSetDevice(h->map->device));
    if (check_collision) k_pf_motion<true><<<(n + 127) / 128, 128, 0, h->stream>>>(h->map->view(), poses_dev, attrs_dev, n, *T, forget_rate);
    else                 k_pf_motion<false><<<(n + 127) / 128, 128, 0, h->stream>>>(h->map->view(), poses_dev, attrs_dev, n, *T, forget_rate);
    LAUNCHED();
    return B2_OK;
}

extern "C" int b2_pf_resample_gladiator(b2_pf* h, const b2_transform* poses_dev, const b2_particle_attr* attrs_dev, uint32_t n_all, uint32_t first, uint32_t n_local,
                                        b2_transform* poses_new_dev, b2_particle_attr* attrs_new_dev, const b2_gladiator_config* cfg, uint64_t seed, uint32_t step,
                                        const uint32_t* raw_dev, const float* normals_dev)
{
    NOTNULL(h); NOTNULL(cfg);
    if ((uint64_t)first + n_local > n_all) return fail(B2_ERR_INVALID, "champion range %u+%u exceeds the %u particles", first, n_local, n_all);
    if ((raw_dev == nullptr) != (normals_dev == nullptr)) return fail(B2_ERR_INVALID, "raw_dev and normals_dev must be given together");
    if (n_local == 0) return B2_OK;
    NOTNULL(poses_dev); NOTNULL(attrs_dev); NOTNULL(poses_new_dev); NOTNULL(attrs_new_dev);
    if ((const void*)poses_dev == (const void*)poses_new_dev || (const void*)attrs_dev == (const void*)attrs_new_dev)
        return fail(B2_ERR_INVALID, "resampling is not in place: outputs must not alias the inputs (resampling.cu:112-117 double-buffers)");
    CU(cudaSetDevice(h->map->device));
    k_pf_gladiator<<<(n_local + 255) / 256, 256, 0, h->stream>>>(poses_dev, attrs_dev, n_all, first, n_local, poses_new_dev, attrs_new_dev, *cfg, seed, step, raw_dev, normals_dev);
    LAUNCHED();
    return B2_OK;
}

extern "C" int b2_pf_gladiator_randoms(b2_pf* h, uint64_t seed, uint32_t step, uint32_t first, uint32_t n, uint32_t* raw_dev, float* normals_dev)
{
    NOTNULL(h);
    if (n == 0) return B2_OK;
    NOTNULL(raw_dev); NOTNULL(normals_dev);
    CU(cudaSetDevice(h->map->device));
    k_pf_gladiator_randoms<<<(n + 255) / 256, 256, 0, h->stream>>>(seed, step, first, n, raw_dev, normals_dev);
    LAUNCHED();
    return B2_OK;
}

extern "C" int b2_pf_likelihood_stats(b2_pf* h, const b2_particle_attr* attrs_dev, uint32_t n, float* sum_out, float* max_out)
{
    NOTNULL(h);
    CU(cudaSetDevice(h->map->device));
    float s = 0.f, m = 0.f;
    if (n > 0) {
        NOTNULL(attrs_dev);
        const uint32_t grid = std::min<uint32_t>((uint32_t)h->n_sm * 4u, (n + 255) / 256);
        k_pf_stats<<<grid, 256, 0, h->stream>>>(attrs_dev, n, h->d_part.p, h->d_ticket.p, h->d_out.p);
        LAUNCHED();
        CU(cudaMemcpyAsync(h->h_out, h->d_out.p, 2 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        s = h->h_out[0]; m = h->h_out[1];
    }
    if (sum_out) *sum_out = s;
    if (max_out) *max_out = m;
    return B2_OK;
}
