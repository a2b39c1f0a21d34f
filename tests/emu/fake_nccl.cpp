// TEST INFRASTRUCTURE ONLY (tests/emu): an in-process stand-in for the eight NCCL entry points the library binds with
// dlopen("libnccl.so.2") -- ranks are THREADS of one process (one fg_ctx + fg_comm each), collectives are blocking
// rendezvous over host memory (the emulated "device" memory is host memory and its streams are synchronous). It lets
// the CPU suite run the collective host call (fgh_search_batch_sharded: shared planning, the fused exchange + merge,
// the per-shard answers of deep pages and nested queries) with world_size 2 and 3. Built with the soname libnccl.so.2
// and loaded RTLD_GLOBAL by tests/emu/run_sharded_threads.py before the library looks for NCCL. Never shipped.
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "nccl.h"
typedef struct fgemu_stream* cudaStream_t;

namespace {
struct Group {
    int n = 0, joined = 0;
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    uint64_t gen = 0;
    std::vector<const void*> send;
    void barrier() {
        std::unique_lock<std::mutex> g(mu);
        const uint64_t my = gen;
        if (++arrived == n) {
            arrived = 0;
            gen++;
            cv.notify_all();
        } else {
            cv.wait(g, [&] { return gen != my; });
        }
    }
};
std::mutex g_mu;
std::map<std::string, Group*> g_groups;
int g_next_id = 1;
size_t elem(ncclDataType_t t) { return t == ncclInt8 || t == ncclUint8 ? 1 : (t == ncclInt64 || t == ncclUint64 ? 8 : 4); }
}  // namespace

struct ncclComm {
    Group* g;
    int rank;
};

extern "C" {
ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
    std::lock_guard<std::mutex> g(g_mu);
    memset(id, 0, sizeof(*id));
    const int v = g_next_id++;
    memcpy(id->internal, &v, sizeof(v));
    return ncclSuccess;
}
ncclResult_t ncclCommInitRank(ncclComm_t* out, int n, ncclUniqueId id, int rank) {
    Group* grp;
    {
        std::lock_guard<std::mutex> g(g_mu);
        Group*& slot = g_groups[std::string(id.internal, sizeof(id.internal))];
        if (!slot) {
            slot = new Group();
            slot->n = n;
            slot->send.assign((size_t)n, nullptr);
        }
        grp = slot;
        if (grp->n != n || rank < 0 || rank >= n) return ncclInternalError;
    }
    *out = new ncclComm{grp, rank};
    grp->barrier();  // like the real call: returns once every rank has joined
    return ncclSuccess;
}
ncclResult_t ncclCommDestroy(ncclComm_t c) {
    delete c;
    return ncclSuccess;
}
ncclResult_t ncclAllGather(const void* send, void* recv, size_t count, ncclDataType_t t, ncclComm_t c, cudaStream_t) {
    Group* g = c->g;
    const size_t bytes = count * elem(t);
    g->send[(size_t)c->rank] = send;
    g->barrier();
    for (int r = 0; r < g->n; r++)
        if ((const char*)g->send[(size_t)r] != (char*)recv + bytes * (size_t)r) memcpy((char*)recv + bytes * (size_t)r, g->send[(size_t)r], bytes);
    g->barrier();  // nobody reuses a send buffer before every rank has read it
    return ncclSuccess;
}
ncclResult_t ncclAllReduce(const void* send, void* recv, size_t count, ncclDataType_t t, ncclRedOp_t, ncclComm_t c, cudaStream_t) {
    Group* g = c->g;
    g->send[(size_t)c->rank] = send;
    g->barrier();
    std::vector<char> tmp(count * elem(t), 0);
    for (int r = 0; r < g->n; r++)
        for (size_t i = 0; i < count; i++) {
            if (elem(t) == 8) ((uint64_t*)tmp.data())[i] += ((const uint64_t*)g->send[(size_t)r])[i];
            else if (t == ncclFloat32) ((float*)tmp.data())[i] += ((const float*)g->send[(size_t)r])[i];
            else if (elem(t) == 4) ((uint32_t*)tmp.data())[i] += ((const uint32_t*)g->send[(size_t)r])[i];
            else ((uint8_t*)tmp.data())[i] += ((const uint8_t*)g->send[(size_t)r])[i];
        }
    g->barrier();  // (in place: every rank has read every input before anyone writes)
    memcpy(recv, tmp.data(), tmp.size());
    return ncclSuccess;
}
ncclResult_t ncclGroupStart() { return ncclSuccess; }  // every rank issues the grouped calls in the same order: blocking is fine
ncclResult_t ncclGroupEnd() { return ncclSuccess; }
const char* ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "fake NCCL error"; }
}
