// lb_multigpu.cpp — frame partitioning across the GPUs of one box INSIDE the library: sample sharding with one ncclReduce of the
// fp32 accumulation buffers, row bands with a ReSTIR halo and one gather, for one rank per process (lb_comm_*) and for one process driving n
// GPUs (lb_group_*, ncclCommInitAll). New capability: the reference is single-GPU (one CUDA context, PT/Framework/WaveFrontRenderer.cpp:70-322).
//
// Built on the public C ABI of the renderer (lb_get_settings, lb_get_stream, lb_accum_buffer, lb_hdr_buffer, lb_resolve_accum,
// lb_render_frames): the collectives are enqueued on the renderers' own streams behind the frames they wait for, so there is no host
// synchronisation between rendering and the exchange. NCCL is loaded at run time (dlopen "libnccl.so.2"): a single-GPU application needs no
// NCCL, and a process that already holds one (torch) shares it. No CPU fallback: without NCCL or CUDA devices every call fails with an error.
#include "../../include/lumen_b200.h"
#include <cuda_runtime.h>
#include <nccl.h>
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace {

thread_local std::string g_mg_err;
int mfail(int code, const std::string& msg) { g_mg_err = msg; return code; }

// ---------------------------------------------------------------- NCCL, resolved at run time
struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
    bool load() {
        if (lib) return true;
        // LB_NCCL_LIB names the library file explicitly. It matters when the process will ALSO load a framework that bundles a newer NCCL under
        // the same soname (torch): whichever libnccl.so.2 is mapped first serves both, so the host should point this at the bundled one
        // (lumenrenderer_b200/__init__.py does) instead of letting the system copy win.
        const char* forced = getenv("LB_NCCL_LIB");
        if (forced && *forced) lib = dlopen(forced, RTLD_NOW | RTLD_GLOBAL);
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) { if (lib) break; lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL); }
        if (!lib) { error = std::string("NCCL not found (libnccl.so.2): ") + (dlerror() ? dlerror() : ""); return false; }
        auto sym = [&](const char* n) { void* p = dlsym(lib, n); if (!p) error = std::string("NCCL symbol missing: ") + n; return p; };
        GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId"); CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
        CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll"); CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        Reduce = (decltype(Reduce))sym("ncclReduce"); Send = (decltype(Send))sym("ncclSend"); Recv = (decltype(Recv))sym("ncclRecv");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart"); GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        if (!error.empty()) { dlclose(lib); lib = nullptr; return false; }
        return true;
    }
};
Nccl& nccl() { static Nccl n; return n; }
std::mutex g_mu;

#define MG_NCCL(expr) do { const ncclResult_t rc_ = (expr); if (rc_ != ncclSuccess) return mfail(LB_ERR_CUDA, std::string(#expr) + ": " + nccl().GetErrorString(rc_)); } while (0)
#define MG_CUDA(expr) do { const cudaError_t rc_ = (expr); if (rc_ != cudaSuccess) return mfail(LB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(rc_)); } while (0)
#define MG_LB(expr) do { const int rc_ = (expr); if (rc_ != LB_OK) return mfail(rc_, std::string(#expr) + ": " + lb_last_error()); } while (0)

// owned rows [y0, y1) and rendered rows [h0, h1) of band `rank` (lumenrenderer_b200/sharding.py band_partition is the same arithmetic)
void band_rows(uint32_t height, uint32_t width, uint32_t rank, uint32_t ranks, uint32_t& y0, uint32_t& y1, uint32_t& h0, uint32_t& h1) {
    auto gcd = [](uint32_t a, uint32_t b) { while (b) { const uint32_t t = a % b; a = b; b = t; } return a; };
    const uint32_t step = width ? 256u / gcd(width, 256u) : 1u;
    y0 = (uint32_t)((uint64_t)height * rank / ranks); y1 = (uint32_t)((uint64_t)height * (rank + 1) / ranks);
    h0 = y0 > LB_RESTIR_HALO ? y0 - LB_RESTIR_HALO : 0u; h0 -= h0 % step;
    h1 = y1 + LB_RESTIR_HALO < height ? y1 + LB_RESTIR_HALO : height;
}

struct Comm { ncclComm_t comm = nullptr; int rank = 0, ranks = 1; };
std::map<LbRenderer, Comm>& comms() { static std::map<LbRenderer, Comm> m; return m; }

struct Member { LbRenderer r = nullptr; int device = 0; cudaStream_t stream = nullptr; ncclComm_t comm = nullptr; LbSettings st{}; uint32_t y0 = 0, y1 = 0; };

// enqueue the band gather of one frame: inside one NCCL group, `self` sends its owned rows, the root posts one receive per other rank
int gather_rows(ncclComm_t comm, int rank, int ranks, int root, cudaStream_t stream, const LbSettings& st, const float* merged, float* full) {
    const uint32_t H = st.band_full_height ? st.band_full_height : st.height, W = st.width;
    uint32_t y0, y1, h0, h1; band_rows(H, W, (uint32_t)rank, (uint32_t)ranks, y0, y1, h0, h1);
    const float* own = merged + (size_t)(y0 - st.band_row0) * W * 4;
    if (rank == root) {
        if (!full) return mfail(LB_ERR_INVALID_ARGUMENT, "the root of a band gather needs the full-frame buffer");
        MG_CUDA(cudaMemcpyAsync(full + (size_t)y0 * W * 4, own, (size_t)(y1 - y0) * W * 16, cudaMemcpyDeviceToDevice, stream));
        for (int p = 0; p < ranks; ++p) {
            if (p == root) continue;
            uint32_t a, b, c, d; band_rows(H, W, (uint32_t)p, (uint32_t)ranks, a, b, c, d);
            MG_NCCL(nccl().Recv(full + (size_t)a * W * 4, (size_t)(b - a) * W * 4, ncclFloat, p, comm, stream));
        }
    } else MG_NCCL(nccl().Send(own, (size_t)(y1 - y0) * W * 4, ncclFloat, root, comm, stream));
    return LB_OK;
}

}

struct LbGroup_t {
    int mode = LB_GROUP_SAMPLES;
    std::vector<Member> members;
    float* full = nullptr;              // bands: the gathered frame on member 0's device
    uint32_t width = 0, height = 0;     // of the complete frame
    uint32_t frames_per_member = 0;     // samples: accumulated since creation
    bool reduced = false;
};

extern "C" {

LB_API const char* lb_multigpu_last_error(void) { return g_mg_err.c_str(); }

LB_API int lb_band_settings(const LbSettings* full, uint32_t rank, uint32_t ranks, LbSettings* out, uint32_t* own_y0, uint32_t* own_y1) {
    if (!full || !out || !ranks || rank >= ranks || !full->width || !full->height) return mfail(LB_ERR_INVALID_ARGUMENT, "bad band request");
    if (full->band_full_height) return mfail(LB_ERR_INVALID_ARGUMENT, "the settings already describe a band");
    if (full->height < ranks) return mfail(LB_ERR_INVALID_ARGUMENT, "fewer rows than ranks");
    uint32_t y0, y1, h0, h1; band_rows(full->height, full->width, rank, ranks, y0, y1, h0, h1);
    *out = *full;
    out->height = h1 - h0; out->band_row0 = h0; out->band_full_height = full->height; out->band_own_row0 = y0; out->band_own_rows = y1 - y0;
    if (own_y0) *own_y0 = y0;
    if (own_y1) *own_y1 = y1;
    return LB_OK;
}

LB_API int lb_shard_settings(const LbSettings* base, uint32_t rank, uint32_t ranks, LbSettings* out) {
    if (!base || !out || !ranks || rank >= ranks) return mfail(LB_ERR_INVALID_ARGUMENT, "bad shard request");
    *out = *base;
    out->blend_output = 1u; out->first_frame_count = 2u * rank; out->frame_count_stride = 2u * ranks;
    return LB_OK;
}

// ---------------------------------------------------------------- one rank per process
LB_API int lb_comm_unique_id(uint8_t* id128) {
    if (!id128) return mfail(LB_ERR_INVALID_ARGUMENT, "null");
    std::lock_guard<std::mutex> lock(g_mu);
    if (!nccl().load()) return mfail(LB_ERR_UNSUPPORTED, nccl().error);
    static_assert(sizeof(ncclUniqueId) == LB_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id; MG_NCCL(nccl().GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return LB_OK;
}
LB_API int lb_comm_init(LbRenderer r, const uint8_t* id128, int rank, int ranks) {
    if (!r || !id128 || ranks < 1 || rank < 0 || rank >= ranks) return mfail(LB_ERR_INVALID_ARGUMENT, "bad communicator request");
    std::lock_guard<std::mutex> lock(g_mu);
    if (!nccl().load()) return mfail(LB_ERR_UNSUPPORTED, nccl().error);
    if (comms().count(r)) return mfail(LB_ERR_STATE, "the renderer already has a communicator");
    LbSettings st; MG_LB(lb_get_settings(r, &st));
    MG_CUDA(cudaSetDevice(st.device));
    ncclUniqueId id; memcpy(&id, id128, sizeof id);
    Comm c; c.rank = rank; c.ranks = ranks;
    MG_NCCL(nccl().CommInitRank(&c.comm, ranks, id, rank));
    comms()[r] = c;
    return LB_OK;
}
LB_API int lb_comm_destroy(LbRenderer r) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = comms().find(r);
    if (it == comms().end()) return mfail(LB_ERR_INVALID_HANDLE, "no communicator on this renderer");
    lb_synchronize(r);
    nccl().CommDestroy(it->second.comm);
    comms().erase(it);
    return LB_OK;
}
LB_API int lb_comm_reduce_accum(LbRenderer r, int root, uint32_t total_frames) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = comms().find(r);
    if (it == comms().end()) return mfail(LB_ERR_INVALID_HANDLE, "no communicator on this renderer (lb_comm_init)");
    const Comm& c = it->second;
    if (root < 0 || root >= c.ranks || !total_frames) return mfail(LB_ERR_INVALID_ARGUMENT, "root / total_frames");
    void* side = nullptr; void* send = nullptr; void* recv = nullptr; size_t bytes = 0; LbSettings st;
    MG_LB(lb_get_settings(r, &st)); MG_CUDA(cudaSetDevice(st.device));
    MG_LB(lb_reduce_begin(r, &side, &send, &recv, &bytes));
    // the one collective of the path, on the renderer's side stream: the frames that follow overlap it
    const ncclResult_t nr = nccl().Reduce(send, c.rank == root ? recv : send, bytes / 4, ncclFloat, ncclSum, root, c.comm, (cudaStream_t)side);
    const int er = lb_reduce_end(r, c.rank == root && nr == ncclSuccess ? 1 : 0, total_frames);      // always closes what lb_reduce_begin opened
    if (nr != ncclSuccess) return mfail(LB_ERR_CUDA, std::string("ncclReduce: ") + nccl().GetErrorString(nr));
    if (er != LB_OK) return mfail(er, std::string("lb_reduce_end: ") + lb_last_error());
    return LB_OK;
}
LB_API int lb_comm_gather_bands(LbRenderer r, int root, void* full_frame_device) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = comms().find(r);
    if (it == comms().end()) return mfail(LB_ERR_INVALID_HANDLE, "no communicator on this renderer (lb_comm_init)");
    const Comm& c = it->second;
    if (root < 0 || root >= c.ranks) return mfail(LB_ERR_INVALID_ARGUMENT, "root");
    void* hdr = nullptr; size_t bytes = 0; void* stream = nullptr; LbSettings st;
    MG_LB(lb_hdr_buffer(r, &hdr, &bytes)); MG_LB(lb_get_stream(r, &stream)); MG_LB(lb_get_settings(r, &st));
    if (c.ranks > 1 && !st.band_full_height) return mfail(LB_ERR_STATE, "the renderer is not a band (lb_band_settings)");
    MG_CUDA(cudaSetDevice(st.device));
    MG_NCCL(nccl().GroupStart());
    const int rc = gather_rows(c.comm, c.rank, c.ranks, root, (cudaStream_t)stream, st, (const float*)hdr, (float*)full_frame_device);
    MG_NCCL(nccl().GroupEnd());
    return rc;
}

// ---------------------------------------------------------------- one process, n GPUs
LB_API int lb_group_create(const int* devices, uint32_t n, const LbSettings* settings, int mode, LbGroup* out) {
    if (!devices || !n || !settings || !out || (mode != LB_GROUP_SAMPLES && mode != LB_GROUP_BANDS)) return mfail(LB_ERR_INVALID_ARGUMENT, "bad group request");
    std::lock_guard<std::mutex> lock(g_mu);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return mfail(LB_ERR_CUDA, "no CUDA device: liblumen_b200 has no CPU fallback");
    if (!nccl().load()) return mfail(LB_ERR_UNSUPPORTED, nccl().error);
    std::unique_ptr<LbGroup_t> g(new LbGroup_t());
    g->mode = mode; g->width = settings->width; g->height = settings->height;
    std::vector<ncclComm_t> cs(n);
    for (uint32_t i = 0; i < n; ++i) {
        Member m; m.device = devices[i];
        int rc = mode == LB_GROUP_SAMPLES ? lb_shard_settings(settings, i, n, &m.st) : lb_band_settings(settings, i, n, &m.st, &m.y0, &m.y1);
        if (rc == LB_OK) { m.st.device = devices[i]; rc = lb_create(&m.st, &m.r); if (rc != LB_OK) mfail(rc, std::string("lb_create: ") + lb_last_error()); }
        if (rc == LB_OK) { void* s = nullptr; rc = lb_get_stream(m.r, &s); m.stream = (cudaStream_t)s; }
        if (rc != LB_OK) { for (Member& q : g->members) lb_destroy(q.r); return rc; }
        g->members.push_back(m);
    }
    const ncclResult_t rc = nccl().CommInitAll(cs.data(), (int)n, devices);
    if (rc != ncclSuccess) { for (Member& q : g->members) lb_destroy(q.r); return mfail(LB_ERR_CUDA, std::string("ncclCommInitAll: ") + nccl().GetErrorString(rc)); }
    for (uint32_t i = 0; i < n; ++i) g->members[i].comm = cs[i];
    if (mode == LB_GROUP_BANDS) {
        MG_CUDA(cudaSetDevice(devices[0]));
        MG_CUDA(cudaMalloc((void**)&g->full, (size_t)g->width * g->height * 16));
        MG_CUDA(cudaMemset(g->full, 0, (size_t)g->width * g->height * 16));
    }
    *out = g.release();
    return LB_OK;
}
LB_API int lb_group_size(LbGroup g, uint32_t* n) { if (!g || !n) return mfail(LB_ERR_INVALID_ARGUMENT, "null"); *n = (uint32_t)g->members.size(); return LB_OK; }
LB_API int lb_group_member(LbGroup g, uint32_t i, LbRenderer* out) {
    if (!g || !out || i >= g->members.size()) return mfail(LB_ERR_INVALID_HANDLE, "member");
    *out = g->members[i].r; return LB_OK;
}
LB_API int lb_group_render(LbGroup g, uint32_t frames) {
    if (!g) return mfail(LB_ERR_INVALID_ARGUMENT, "null group");
    std::lock_guard<std::mutex> lock(g_mu);
    // launches are asynchronous: one host thread keeps every GPU's stream fed, frame by frame (a frame is ~25 launches, the GPUs need
    // milliseconds for it)
    for (uint32_t f = 0; f < frames; ++f) {
        for (Member& m : g->members) MG_LB(lb_render_frames(m.r, 1));
        if (g->mode == LB_GROUP_BANDS) {
            MG_NCCL(nccl().GroupStart());
            int rc = LB_OK;
            for (size_t i = 0; i < g->members.size() && rc == LB_OK; ++i) {
                Member& m = g->members[i];
                void* hdr = nullptr; size_t bytes = 0;
                rc = lb_hdr_buffer(m.r, &hdr, &bytes);
                if (rc == LB_OK && cudaSetDevice(m.device) != cudaSuccess) rc = mfail(LB_ERR_CUDA, "cudaSetDevice");
                if (rc == LB_OK) rc = gather_rows(m.comm, (int)i, (int)g->members.size(), 0, m.stream, m.st, (const float*)hdr, i == 0 ? g->full : nullptr);
            }
            MG_NCCL(nccl().GroupEnd());
            if (rc != LB_OK) return rc;
        }
    }
    if (g->mode == LB_GROUP_SAMPLES) g->frames_per_member += frames;
    return LB_OK;
}
// samples: start a new progressive image (every member's accumulation buffer is cleared; the frameCount streams keep advancing)
LB_API int lb_group_reset(LbGroup g) {
    if (!g) return mfail(LB_ERR_INVALID_ARGUMENT, "null group");
    std::lock_guard<std::mutex> lock(g_mu);
    if (g->mode == LB_GROUP_SAMPLES) for (Member& m : g->members) MG_LB(lb_set_blend_mode(m.r, 1));
    g->frames_per_member = 0; g->reduced = false;
    return LB_OK;
}
LB_API int lb_group_reduce(LbGroup g) {
    if (!g) return mfail(LB_ERR_INVALID_ARGUMENT, "null group");
    if (g->mode != LB_GROUP_SAMPLES) return LB_OK;
    std::lock_guard<std::mutex> lock(g_mu);
    if (!g->frames_per_member) return mfail(LB_ERR_STATE, "nothing rendered yet");
    struct Leg { void* side; void* send; void* recv; size_t bytes; };
    std::vector<Leg> legs(g->members.size());
    for (size_t i = 0; i < g->members.size(); ++i) MG_LB(lb_reduce_begin(g->members[i].r, &legs[i].side, &legs[i].send, &legs[i].recv, &legs[i].bytes));
    MG_NCCL(nccl().GroupStart());
    for (size_t i = 0; i < g->members.size(); ++i) {
        Member& m = g->members[i];
        MG_CUDA(cudaSetDevice(m.device));
        MG_NCCL(nccl().Reduce(legs[i].send, i == 0 ? legs[i].recv : legs[i].send, legs[i].bytes / 4, ncclFloat, ncclSum, 0, m.comm, (cudaStream_t)legs[i].side));
    }
    MG_NCCL(nccl().GroupEnd());
    for (size_t i = 0; i < g->members.size(); ++i) MG_LB(lb_reduce_end(g->members[i].r, i == 0 ? 1 : 0, g->frames_per_member * (uint32_t)g->members.size()));
    return LB_OK;
}
LB_API int lb_group_synchronize(LbGroup g) {
    if (!g) return mfail(LB_ERR_INVALID_ARGUMENT, "null group");
    for (Member& m : g->members) MG_LB(lb_synchronize(m.r));
    return LB_OK;
}
LB_API int lb_group_read_hdr(LbGroup g, float* rgba, size_t cap) {
    if (!g || !rgba) return mfail(LB_ERR_INVALID_ARGUMENT, "null");
    const size_t bytes = (size_t)g->width * g->height * 16;
    if (cap < bytes) return mfail(LB_ERR_INVALID_ARGUMENT, "buffer too small");
    MG_LB(lb_group_synchronize(g));
    if (g->mode == LB_GROUP_BANDS) {
        MG_CUDA(cudaSetDevice(g->members[0].device));
        MG_CUDA(cudaMemcpy(rgba, g->full, bytes, cudaMemcpyDeviceToHost));
        return LB_OK;
    }
    MG_LB(lb_read_hdr(g->members[0].r, rgba, cap));
    return LB_OK;
}
LB_API int lb_group_destroy(LbGroup g) {
    if (!g) return mfail(LB_ERR_INVALID_ARGUMENT, "null group");
    std::lock_guard<std::mutex> lock(g_mu);
    for (Member& m : g->members) { lb_synchronize(m.r); if (m.comm) nccl().CommDestroy(m.comm); }
    for (Member& m : g->members) lb_destroy(m.r);
    if (g->full) { cudaSetDevice(g->members[0].device); cudaFree(g->full); }
    delete g;
    return LB_OK;
}

}
