// sg_comm.cu — the library-owned NCCL communicator of a data-parallel replica (SURVEY.md §8b "the library owns only TMA
// descriptors, its workspace handle and the NCCL communicator"; §8e: all-reduce of the flat gradient buckets over
// NVLink / NVSwitch, started on a communication stream as soon as a group of gradients is final).
//
// NCCL is bound at RUN time (dlopen of libnccl.so.2): the process that hosts this library (PyTorch) already carries its
// own copy, and binding to the one that is loaded keeps a single NCCL instance per process. Only the stable core
// of the NCCL API is used; its few types are restated here so that the build needs no NCCL header.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>

#include "sg_comm.cuh"

namespace sg {
namespace {

struct UniqueId {
    char internal[128];
};  // ncclUniqueId
constexpr int kNcclFloat32 = 7, kNcclSum = 0, kNcclAvg = 4;

using GetUniqueIdFn = int (*)(UniqueId*);
using CommInitRankFn = int (*)(void**, int, UniqueId, int);
using CommDestroyFn = int (*)(void*);
using AllReduceFn = int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t);
using GetErrorStringFn = const char* (*)(int);
using GetVersionFn = int (*)(int*);

struct Api {
    void* handle = nullptr;
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    AllReduceFn all_reduce = nullptr;
    GetErrorStringFn get_error_string = nullptr;
    GetVersionFn get_version = nullptr;
    bool tried = false;
};
Api g_api;
thread_local char t_comm_err[512] = "";

int comm_fail(const char* what, int code) {
    snprintf(t_comm_err, sizeof(t_comm_err), "%s: %s (nccl result %d)", what,
             (g_api.get_error_string && code > 0) ? g_api.get_error_string(code) : "failed", code);
    return -1;
}

int load_api() {
    if (g_api.handle) return 0;
    if (g_api.tried) {
        snprintf(t_comm_err, sizeof(t_comm_err), "libnccl.so.2 could not be loaded");
        return -1;
    }
    g_api.tried = true;
    // the copy the host process already mapped (PyTorch's) first, then the system one
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        snprintf(t_comm_err, sizeof(t_comm_err), "dlopen(libnccl.so.2) failed: %s", dlerror());
        return -1;
    }
    g_api.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(h, "ncclGetUniqueId"));
    g_api.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(h, "ncclCommInitRank"));
    g_api.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(h, "ncclCommDestroy"));
    g_api.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(h, "ncclAllReduce"));
    g_api.get_error_string = reinterpret_cast<GetErrorStringFn>(dlsym(h, "ncclGetErrorString"));
    g_api.get_version = reinterpret_cast<GetVersionFn>(dlsym(h, "ncclGetVersion"));
    if (!g_api.get_unique_id || !g_api.comm_init_rank || !g_api.comm_destroy || !g_api.all_reduce) {
        snprintf(t_comm_err, sizeof(t_comm_err), "libnccl.so.2 lacks a required symbol");
        return -1;
    }
    g_api.handle = h;
    return 0;
}

}  // namespace

const char* comm_last_error() { return t_comm_err; }

int comm_version() {
    int v = 0;
    if (load_api() != 0 || !g_api.get_version || g_api.get_version(&v) != 0) return -1;
    return v;
}

int comm_unique_id(void* host_id_out, size_t cap) {
    if (!host_id_out || cap < sizeof(UniqueId)) {
        snprintf(t_comm_err, sizeof(t_comm_err), "comm_unique_id: the buffer must hold %zu bytes", sizeof(UniqueId));
        return -1;
    }
    if (load_api() != 0) return -1;
    UniqueId id;
    const int rc = g_api.get_unique_id(&id);
    if (rc != 0) return comm_fail("ncclGetUniqueId", rc);
    memcpy(host_id_out, &id, sizeof(id));
    return 0;
}

int comm_create(Comm* cm, const void* host_id, size_t id_bytes, int rank, int world) {
    if (!cm || !host_id || id_bytes < sizeof(UniqueId) || world < 1 || rank < 0 || rank >= world) {
        snprintf(t_comm_err, sizeof(t_comm_err), "comm_create: bad argument");
        return -1;
    }
    if (load_api() != 0) return -1;
    UniqueId id;
    memcpy(&id, host_id, sizeof(id));
    const int rc = g_api.comm_init_rank(&cm->nccl, world, id, rank);
    if (rc != 0) return comm_fail("ncclCommInitRank", rc);
    cm->rank = rank;
    cm->world = world;
    if (cudaStreamCreateWithFlags(&cm->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&cm->ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&cm->done, cudaEventDisableTiming) != cudaSuccess) {
        snprintf(t_comm_err, sizeof(t_comm_err), "comm_create: cannot create the communication stream / events");
        return -1;
    }
    return 0;
}

void comm_release(Comm* cm) {
    if (!cm) return;
    if (cm->nccl && g_api.comm_destroy) g_api.comm_destroy(cm->nccl);
    if (cm->stream) cudaStreamDestroy(cm->stream);
    if (cm->ready) cudaEventDestroy(cm->ready);
    if (cm->done) cudaEventDestroy(cm->done);
    *cm = Comm();
}

// buf[0..count) <- mean over the ranks, in place, enqueued on `on`.
static int all_reduce_mean(Comm* cm, float* buf, long long count, cudaStream_t on) {
    const int rc = g_api.all_reduce(buf, buf, static_cast<size_t>(count), kNcclFloat32, kNcclAvg, cm->nccl, on);
    if (rc != 0) return comm_fail("ncclAllReduce", rc);
    return 0;
}

int comm_all_reduce_mean(Comm* cm, float* buf, long long count, cudaStream_t s) {
    if (!cm || !cm->nccl) {
        snprintf(t_comm_err, sizeof(t_comm_err), "all-reduce without a communicator (call sg_comm_init first)");
        return -1;
    }
    if (count <= 0) return 0;
    return all_reduce_mean(cm, buf, count, s);
}

// Everything enqueued on `s` so far produces buf; the reduction runs on the communication stream, so kernels enqueued
// on `s` afterwards (the rest of the backward pass) overlap it. comm_join makes `s` wait for all started reductions.
int comm_all_reduce_mean_start(Comm* cm, float* buf, long long count, cudaStream_t s) {
    if (!cm || !cm->nccl) {
        snprintf(t_comm_err, sizeof(t_comm_err), "all-reduce without a communicator (call sg_comm_init first)");
        return -1;
    }
    if (count <= 0) return 0;
    if (cudaEventRecord(cm->ready, s) != cudaSuccess || cudaStreamWaitEvent(cm->stream, cm->ready, 0) != cudaSuccess) {
        snprintf(t_comm_err, sizeof(t_comm_err), "comm: cannot order the communication stream after the launch stream");
        return -1;
    }
    if (all_reduce_mean(cm, buf, count, cm->stream) != 0) return -1;
    cm->pending = true;
    return 0;
}

int comm_join(Comm* cm, cudaStream_t s) {
    if (!cm || !cm->pending) return 0;
    if (cudaEventRecord(cm->done, cm->stream) != cudaSuccess || cudaStreamWaitEvent(s, cm->done, 0) != cudaSuccess) {
        snprintf(t_comm_err, sizeof(t_comm_err), "comm: cannot join the communication stream");
        return -1;
    }
    cm->pending = false;
    return 0;
}

}  // namespace sg
