// exchange.h -- the collectives of the sharded job (DESIGN.md "Multi-GPU"), behind one small
// interface so that the same launch plan runs (a) one rank per process over NCCL / NVLink and
// (b) several virtual ranks inside one process with device-to-device copies (how a single-GPU
// box tests the sharded path).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace fqd {

struct Exchange {
    int rank = 0, world = 1;
    // views of the other ranks' arena slabs (CUDA IPC), kept across jobs while the slabs stay the same
    void *peer_map[64] = {};
    unsigned char peer_handle[64][64] = {};
    bool peer_open[64] = {};
    virtual ~Exchange() {}
    // every rank contributes `bytes` bytes; recv holds world * bytes
    virtual int allgather(const void *send, void *recv, size_t bytes, cudaStream_t s) = 0;
    // rank g contributes bytes[g] at recv + off[g] (send is this rank's part)
    virtual int allgatherv(const void *send, void *recv, const size_t *off, const size_t *bytes,
                           cudaStream_t s) = 0;
    // send_bytes[g] bytes at send + send_off[g] go to rank g; recv_bytes[g] arrive from g
    virtual int alltoallv(const void *send, const size_t *send_off, const size_t *send_bytes,
                          void *recv, const size_t *recv_off, const size_t *recv_bytes,
                          cudaStream_t s) = 0;
    virtual int allreduce_max_u8(void *buf, size_t n, cudaStream_t s) = 0;
};

// NCCL-backed exchange (libnccl.so.2 is opened at run time; no link-time dependency).
int nccl_unique_id(uint8_t id[128]);
int nccl_exchange_create(int rank, int world, const uint8_t id[128], Exchange **out);

}  // namespace fqd
