// dist.cu -- collectives and redistribution of the single-frequency multi-GPU path (see dist.cuh, lu.cuh).
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the shared library keeps no link-time dependency on it, and a host
// process that has already loaded an NCCL (e.g. the one bundled with torch) shares that copy instead of getting a second one.
// Only the stable C entry points are used; their prototypes are restated here (nccl.h: ncclGetUniqueId, ncclCommInitRank,
// ncclBroadcast, ncclReduce, ncclAllReduce, ncclSend, ncclRecv, ncclGroupStart/End).
#include "dist.cuh"
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>

namespace mfbd {

namespace {
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
enum { NCCL_CHAR = 0, NCCL_DOUBLE = 8, NCCL_SUM = 0 };
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(nccl_uid*) = nullptr;
  int (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
  int (*CommDestroy)(nccl_comm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
};
NcclApi g_nccl;

bool load_nccl(std::string& err) {
  if (g_nccl.handle) return true;
  const char* names[4] = {getenv("MFB_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
  void* h = nullptr;
  for (int i = 0; i < 3 && !h; i++) if (names[i] && names[i][0]) h = dlopen(names[i], RTLD_NOW | RTLD_LOCAL);
  if (!h) { const char* de = dlerror(); err = std::string("cannot load NCCL (libnccl.so.2): ") + (de ? de : "not found") + "; set MFB_NCCL_LIB"; return false; }
  bool ok = true;
  auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) { ok = false; err = std::string("NCCL symbol missing: ") + n; } return p; };
  g_nccl.GetUniqueId = (int (*)(nccl_uid*))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(nccl_comm*, int, nccl_uid, int))sym("ncclCommInitRank");
  g_nccl.CommDestroy = (int (*)(nccl_comm))sym("ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  g_nccl.Broadcast = (int (*)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t))sym("ncclBroadcast");
  g_nccl.Reduce = (int (*)(const void*, void*, size_t, int, int, int, nccl_comm, cudaStream_t))sym("ncclReduce");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t))sym("ncclAllReduce");
  g_nccl.Send = (int (*)(const void*, size_t, int, int, nccl_comm, cudaStream_t))sym("ncclSend");
  g_nccl.Recv = (int (*)(void*, size_t, int, int, nccl_comm, cudaStream_t))sym("ncclRecv");
  g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
  g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
  if (!ok) { dlclose(h); return false; }
  g_nccl.handle = h;
  return true;
}

struct NcclComm : DistComm {
  nccl_comm comm = nullptr; int rank = 0; std::string err;
  int fail(int r, const char* what) { err = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return r > 0 ? -r : -1; }   // negative: not a cudaError
  ~NcclComm() override { if (comm) g_nccl.CommDestroy(comm); }
  int bcast_bytes(int root, const int*, void* const* bufs, size_t bytes, cudaStream_t const* st, int) override {
    int r = g_nccl.Broadcast(bufs[0], bufs[0], bytes, NCCL_CHAR, root, comm, st[0]);
    return r ? fail(r, "ncclBroadcast") : 0;
  }
  int reduce_sum2(int root, const int*, double* const* a, double* const* b, size_t count, cudaStream_t const* st, int) override {
    int r = g_nccl.GroupStart(); if (r) return fail(r, "ncclGroupStart");
    r = g_nccl.Reduce(a[0], a[0], count, NCCL_DOUBLE, NCCL_SUM, root, comm, st[0]); if (r) return fail(r, "ncclReduce");
    r = g_nccl.Reduce(b[0], b[0], count, NCCL_DOUBLE, NCCL_SUM, root, comm, st[0]); if (r) return fail(r, "ncclReduce");
    r = g_nccl.GroupEnd(); return r ? fail(r, "ncclGroupEnd") : 0;
  }
  int allreduce_sum(const int*, double* const* bufs, size_t count, cudaStream_t const* st, int) override {
    int r = g_nccl.AllReduce(bufs[0], bufs[0], count, NCCL_DOUBLE, NCCL_SUM, comm, st[0]);
    return r ? fail(r, "ncclAllReduce") : 0;
  }
  int exchange(double* const* send, const size_t* ns, double* const* recv, const size_t* nr, cudaStream_t st) override {
    int r = g_nccl.GroupStart(); if (r) return fail(r, "ncclGroupStart");
    for (int q = 0; q < P; q++) {
      if (q == rank) continue;
      if (ns[q] > 0) { r = g_nccl.Send(send[q], ns[q], NCCL_DOUBLE, q, comm, st); if (r) return fail(r, "ncclSend"); }
      if (nr[q] > 0) { r = g_nccl.Recv(recv[q], nr[q], NCCL_DOUBLE, q, comm, st); if (r) return fail(r, "ncclRecv"); }
    }
    r = g_nccl.GroupEnd(); return r ? fail(r, "ncclGroupEnd") : 0;
  }
  const char* last_error() override { return err.c_str(); }
};

__global__ void k_add_into(double* __restrict__ dst, const double* __restrict__ src, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] += src[i];
}
// every rank is local and all streams are the same stream: the collectives are stream-ordered copies / adds
struct LoopbackComm : DistComm {
  int find(const int* ranks, int n_local, int r) { for (int i = 0; i < n_local; i++) if (ranks[i] == r) return i; return -1; }
  static int grid(size_t n) { size_t g = (n + 255) / 256; return (int)(g < 1 ? 1 : (g > 1184 ? 1184 : g)); }
  int bcast_bytes(int root, const int* ranks, void* const* bufs, size_t bytes, cudaStream_t const* st, int n_local) override {
    const int ir = find(ranks, n_local, root); if (ir < 0) return -1;
    for (int i = 0; i < n_local; i++) if (i != ir) cudaMemcpyAsync(bufs[i], bufs[ir], bytes, cudaMemcpyDeviceToDevice, st[i]);
    return 0;
  }
  int reduce_sum2(int root, const int* ranks, double* const* a, double* const* b, size_t count, cudaStream_t const* st, int n_local) override {
    const int ir = find(ranks, n_local, root); if (ir < 0) return -1;
    for (int i = 0; i < n_local; i++) if (i != ir) { k_add_into<<<grid(count), 256, 0, st[ir]>>>(a[ir], a[i], count); k_add_into<<<grid(count), 256, 0, st[ir]>>>(b[ir], b[i], count); }
    return 0;
  }
  int allreduce_sum(const int*, double* const* bufs, size_t count, cudaStream_t const* st, int n_local) override {
    for (int i = 1; i < n_local; i++) k_add_into<<<grid(count), 256, 0, st[0]>>>(bufs[0], bufs[i], count);
    for (int i = 1; i < n_local; i++) cudaMemcpyAsync(bufs[i], bufs[0], count * sizeof(double), cudaMemcpyDeviceToDevice, st[0]);
    return 0;
  }
};
}  // namespace

int nccl_unique_id(char out[128], std::string& err) {
  if (!load_nccl(err)) return -1;
  nccl_uid id; memset(&id, 0, sizeof(id));
  int r = g_nccl.GetUniqueId(&id);
  if (r) { err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return r; }
  memcpy(out, id.internal, 128);
  return 0;
}
DistComm* make_nccl_comm(int rank, int nranks, const char unique_id[128], std::string& err) {
  if (!load_nccl(err)) return nullptr;
  NcclComm* c = new NcclComm(); c->P = nranks; c->rank = rank;
  nccl_uid id; memcpy(id.internal, unique_id, 128);
  int r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
  if (r) { err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); c->comm = nullptr; delete c; return nullptr; }
  return c;
}
void launch_add_into(double* dst, const double* src, size_t n, cudaStream_t st) {
  if (n > 0) k_add_into<<<LoopbackComm::grid(n), 256, 0, st>>>(dst, src, n);
}
DistComm* make_loopback_comm(int nranks) { LoopbackComm* c = new LoopbackComm(); c->P = nranks; return c; }

// ------------------------------------------------------------------------------------------------------------------
// Redistribution: the rank that assembled rows [r0, r1) hands, to every rank q, those rows of the columns q owns.
// Local column lc of rank q <-> global column (lc / nb * P + q) * nb + lc % nb.
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_pack_slab(const double* __restrict__ Are, const double* __restrict__ Aim, long long lda, int nb, int P, int q, int r0, int nr, int ncl,
                            double* __restrict__ out) {
  const int lc = blockIdx.y;
  const long long gc = ((long long)(lc / nb) * P + q) * nb + lc % nb;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += gridDim.x * blockDim.x) {
    out[(size_t)lc * nr + i] = Are[gc * lda + r0 + i];
    out[((size_t)ncl + lc) * nr + i] = Aim[gc * lda + r0 + i];
  }
}
__global__ void k_unpack_slab(const double* __restrict__ in, int ncl, int r0, int nr, double* __restrict__ Lre, double* __restrict__ Lim, long long lda) {
  const int lc = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += gridDim.x * blockDim.x) {
    Lre[(long long)lc * lda + r0 + i] = in[(size_t)lc * nr + i];
    Lim[(long long)lc * lda + r0 + i] = in[((size_t)ncl + lc) * nr + i];
  }
}
__global__ void k_copy_own(const double* __restrict__ Are, const double* __restrict__ Aim, long long lda, int nb, int P, int q, int r0, int nr,
                           double* __restrict__ Lre, double* __restrict__ Lim) {
  const int lc = blockIdx.y;
  const long long gc = ((long long)(lc / nb) * P + q) * nb + lc % nb;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += gridDim.x * blockDim.x) {
    Lre[(long long)lc * lda + r0 + i] = Are[gc * lda + r0 + i];
    Lim[(long long)lc * lda + r0 + i] = Aim[gc * lda + r0 + i];
  }
}
static dim3 slab_grid(int nr, int ncl) { int gx = (nr + 255) / 256; if (gx > 64) gx = 64; if (gx < 1) gx = 1; return dim3(gx, ncl); }
void launch_pack_slab(const double* Are, const double* Aim, long long lda, int n, int nb, int P, int q, int r0, int r1, double* out, cudaStream_t st) {
  const int ncl = dist_ncols_local(n, nb, P, q);
  if (r1 > r0 && ncl > 0) k_pack_slab<<<slab_grid(r1 - r0, ncl), 256, 0, st>>>(Are, Aim, lda, nb, P, q, r0, r1 - r0, ncl, out);
}
void launch_unpack_slab(const double* in, int ncl, int r0, int r1, double* Lre, double* Lim, long long lda, cudaStream_t st) {
  if (r1 > r0 && ncl > 0) k_unpack_slab<<<slab_grid(r1 - r0, ncl), 256, 0, st>>>(in, ncl, r0, r1 - r0, Lre, Lim, lda);
}
void launch_copy_own(const double* Are, const double* Aim, long long lda, int n, int nb, int P, int q, int r0, int r1, double* Lre, double* Lim, cudaStream_t st) {
  const int ncl = dist_ncols_local(n, nb, P, q);
  if (r1 > r0 && ncl > 0) k_copy_own<<<slab_grid(r1 - r0, ncl), 256, 0, st>>>(Are, Aim, lda, nb, P, q, r0, r1 - r0, Lre, Lim);
}

}  // namespace mfbd
