// dist.cuh -- plumbing of the single-frequency multi-GPU path: NCCL (loaded at run time, so that the library has no
// link-time dependency and shares the NCCL a host process such as torch has already loaded) or, for the single-GPU self
// test, device copies between virtual ranks; redistribution of the row-slab assembled matrix into the block-cyclic
// column layout of the distributed LU (lu.cuh).
#pragma once
#include "lu.cuh"
#include <string>
#include <vector>

namespace mfbd {

// NCCL communicator of this process (one rank per process, one GPU per rank)
DistComm* make_nccl_comm(int rank, int nranks, const char unique_id[128], std::string& err);
int nccl_unique_id(char out[128], std::string& err);
// all ranks live in this process on one device: collectives are device copies / adds on the (single) stream
DistComm* make_loopback_comm(int nranks);
// rows [r0, r1) of the columns owned by rank q (block-cyclic, block nb) of the planar n x n matrix -> out[plane][local col][row - r0]
void launch_pack_slab(const double* Are, const double* Aim, long long lda, int n, int nb, int P, int q, int r0, int r1, double* out, cudaStream_t st);
// in[plane][local col][row - r0] -> rows [r0, r1) of the local columns (ncl of them) of Lre/Lim
void launch_unpack_slab(const double* in, int ncl, int r0, int r1, double* Lre, double* Lim, long long lda, cudaStream_t st);
// same as pack + unpack for the rank's own rows (no staging)
void launch_copy_own(const double* Are, const double* Aim, long long lda, int n, int nb, int P, int q, int r0, int r1, double* Lre, double* Lim, cudaStream_t st);
// dst[i] += src[i]
void launch_add_into(double* dst, const double* src, size_t n, cudaStream_t st);

}  // namespace mfbd
