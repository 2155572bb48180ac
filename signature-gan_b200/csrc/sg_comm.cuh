// sg_comm.cuh — internal interface of the library-owned NCCL communicator (sg_comm.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace sg {

struct Comm {
    void* nccl = nullptr;           // ncclComm_t
    cudaStream_t stream = nullptr;  // communication stream (reductions overlap the launch stream's kernels)
    cudaEvent_t ready = nullptr;    // launch stream -> communication stream
    cudaEvent_t done = nullptr;     // communication stream -> launch stream
    int rank = 0, world = 1;
    bool pending = false;  // a reduction was started on `stream` and not yet joined
};

const char* comm_last_error();
int comm_version();  // NCCL version code of the bound library, -1 if it cannot be loaded
int comm_unique_id(void* host_id_out, size_t cap);
int comm_create(Comm* cm, const void* host_id, size_t id_bytes, int rank, int world);
void comm_release(Comm* cm);
int comm_all_reduce_mean(Comm* cm, float* buf, long long count, cudaStream_t s);        // in order on `s`
int comm_all_reduce_mean_start(Comm* cm, float* buf, long long count, cudaStream_t s);  // on the communication stream
int comm_join(Comm* cm, cudaStream_t s);

}  // namespace sg
