// lb_host.h — host-side plumbing shared by the translation units of liblumen_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>
#include <vector>
#include <cstdint>
#include <cstdio>
#include "lb_device.cuh"

namespace lb {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
        throw CudaError(buf);
    }
}
#define LB_CUDA(expr) ::lb::cuda_check((expr), #expr, __FILE__, __LINE__)
#define LB_LAUNCH_CHECK() ::lb::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__)

// RAII device array (the reference's MemoryBuffer, LumenPT/src/Framework/MemoryBuffer.cpp:26-101, without the implicit syncs)
template <class T>
struct DevBuf {
    T* p = nullptr; size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    // grow-only allocation; contents are not preserved
    void reserve(size_t count) {
        if (count <= n) return;
        release();
        if (count) LB_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
        n = count;
    }
    void upload(const T* src, size_t count, cudaStream_t s) {
        reserve(count);
        if (count) LB_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void zero(cudaStream_t s) { if (n) LB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    size_t bytes() const { return n * sizeof(T); }
};

// Scratch buffer of a build step: stream-ordered allocation from the device's default memory pool (cudaMallocAsync / cudaFreeAsync on the
// build's stream). The renderer raises the pool's release threshold at creation, so a scene that is re-committed every frame (dynamic
// scenes) re-uses the pool's cached blocks instead of paying ~25 cudaMalloc / cudaFree (each a device-wide synchronisation) per build.
template <class T>
struct StreamBuf {
    T* p = nullptr; size_t n = 0; cudaStream_t s = nullptr;
    StreamBuf() = default;
    StreamBuf(const StreamBuf&) = delete; StreamBuf& operator=(const StreamBuf&) = delete;
    ~StreamBuf() { if (p) cudaFreeAsync(p, s); }
    void reserve(size_t count, cudaStream_t stream) {
        if (count <= n) return;
        if (p) { cudaFreeAsync(p, s); p = nullptr; }
        s = stream; n = count;
        if (count) LB_CUDA(cudaMallocAsync((void**)&p, count * sizeof(T), s));
    }
    void zero() { if (n) LB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

struct DeviceBvh {
    DevBuf<Bvh8Node> nodes; DevBuf<DevTri> tris;
    // what a refit needs (bvh_refit): leaf slot -> index of the source triangle, the first node of every level (nodes are created level by
    // level, so a level is a contiguous range), and a full-precision box per node to hand up to its parent
    DevBuf<uint32_t> perm; DevBuf<float4> box_lo, box_hi; std::vector<uint32_t> level_start;
    uint32_t num_nodes = 0, num_tris = 0 /* leaf entries: triangle references */, src_tris = 0 /* triangles it was built from */, levels = 0, ploc_rounds = 0, refits = 0;
    float split_cell = 0.f;     // grid cell of the early split clipping (0: no triangle was split)
    float build_ms = 0.f, alloc_ms = 0.f, refit_ms = 0.f;
    BvhView view(uint32_t* overflow = nullptr) const { return BvhView{nodes.p, tris.p, num_tris, overflow}; }
    size_t bytes() const { return (size_t)num_nodes * sizeof(Bvh8Node) + 3 * (size_t)num_tris * sizeof(DevTri); }      // three rotated triangle copies
};

// GPU build: 63-bit Morton order -> binary hierarchy by PLOC (SAH-quality, default) or LBVH (Karras 2012, fastest build)
// -> greedy surface-area collapse into compressed 8-wide nodes. Replaces optixAccelBuild (LumenPT/src/Framework/OptixWrapper.cpp:46-131).
// tris_in: world-space triangles in any order; the builder writes its own leaf-ordered copy into out.tris.
enum class BvhBuilder { PLOC, LBVH };
// ploc_radius: neighbour search window of PLOC in Morton order (16, 64 or 128)
void bvh_build(cudaStream_t s, const DevTri* tris_in, uint32_t n, DeviceBvh& out, BvhBuilder builder = BvhBuilder::PLOC, int ploc_radius = 16, float split_fraction = 0.f);

// Same topology, new triangle positions (instances moved): the leaf-ordered triangle copies are re-gathered from `tris_in` (the same source
// order the hierarchy was built from) and every node's child boxes are recomputed bottom-up, level by level — no sort, no clustering, no
// collapse. Hits stay exactly what a rebuild would give (they are a pure function of ray and triangle set); only the boxes' tightness,
// i.e. speed, can degrade when instances move far from where the hierarchy was built. Replaces the reference's IAS rebuild on a transform
// change (PTScene.cpp:74-156, PTMeshInstance.cpp:51-103).
void bvh_refit(cudaStream_t s, const DevTri* tris_in, uint32_t n, DeviceBvh& bvh);

inline int grid_for(size_t n, int block) { return (int)((n + block - 1) / block); }

} // namespace lb
