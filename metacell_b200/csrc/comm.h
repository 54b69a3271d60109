// comm.h -- the rank-side communicator behind mcb_comm (include/mcb200.h) and the collectives the sharded stages use.
// Not part of the ABI.  nccl.h is included for its types only; the functions are resolved at run time (comm.cu).
#pragma once
#include "internal.h"
#include <nccl.h>
#include <vector>

struct mcb_comm {
    mcb::CtxRef ref;
    mcb_ctx* ctx = nullptr;
    void* comm = nullptr;        // ncclComm_t (null when nranks == 1)
    int rank = 0, nranks = 1;
    bool owns = false;           // created by mcb_comm_create (destroyed with the handle) vs. borrowed from a mcb_multi
};

namespace mcb {

int  nccl_version();
void comm_init_all(int ndev, const int* devices, void** comms_out);
void comm_destroy_raw(void* comm);

void comm_allgather_inplace(mcb_comm* c, void* buf, size_t bytes_per_rank);
void comm_allgatherv_inplace(mcb_comm* c, void* buf, const size_t* offs, const size_t* counts, cudaStream_t stream = nullptr);
void comm_alltoallv(mcb_comm* c, const void* send, const size_t* send_off, const size_t* send_cnt, void* recv,
                    const size_t* recv_off, const size_t* recv_cnt);
void comm_gatherv(mcb_comm* c, const void* send, size_t send_cnt, void* recv, const size_t* recv_off, const size_t* recv_cnt,
                  int root);
void comm_reduce_sum_u32(mcb_comm* c, void* buf, size_t count, int root);
void comm_broadcast(mcb_comm* c, void* buf, size_t bytes, int root);
void comm_broadcast_on(mcb_comm* c, void* buf, size_t bytes, int root, cudaStream_t stream);
std::vector<int64_t> comm_allgather_i64(mcb_comm* c, int64_t mine);
std::vector<int64_t> comm_allgather_i64v(mcb_comm* c, const int64_t* mine, int n);
std::vector<int64_t> comm_allgather_dev_i64v(mcb_comm* c, const int64_t* dev_mine, int n);

}  // namespace mcb
