// nccl_glue.cpp -- the only collective on the path: an all-gather.
//
// Chain-sharded runs gather traces once at the end; dataset-sharded global fits
// (walker-make-step's sum over datasets, mcmc-fitting.lisp:1068-1069, split across ranks)
// all-gather per-dataset ll/lp every step and every rank then adds them in dataset order,
// so the accept decision is bit-identical on all ranks (an ncclSum all-reduce would not
// fix the order).  libnccl.so.2 is dlopen'ed on first use: if the process already carries
// an NCCL (torch's), that one is picked up by soname; otherwise the system library.
#include <dlfcn.h>

#include <mutex>
#include <string>

#include "mfit_internal.h"

namespace {

typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;

struct Nccl {
  void* h = nullptr;
  int (*GetUniqueId)(NcclUniqueId*);
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int);
  int (*AllGather)(const void*, void*, size_t, int /*dtype*/, NcclComm, void* /*stream*/);
  int (*CommDestroy)(NcclComm);
  const char* (*GetErrorString)(int);
  std::string err;
};

Nccl* nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "/usr/lib/x86_64-linux-gnu/libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (n.h) break;
    }
    if (!n.h) { n.err = std::string("cannot load libnccl: ") + dlerror(); return; }
#define SYM(f)                                                      \
  *(void**)(&n.f) = dlsym(n.h, "nccl" #f);                          \
  if (!n.f) { n.err = "libnccl lacks nccl" #f; n.h = nullptr; return; }
    SYM(GetUniqueId) SYM(CommInitRank) SYM(AllGather) SYM(CommDestroy) SYM(GetErrorString)
#undef SYM
  });
  return &n;
}

}  // namespace

int mfit_nccl_get_unique_id(void* id128, std::string& err) {
  Nccl* n = nccl();
  if (!n->h) { err = n->err; return MFIT_ERR_NCCL; }
  int rc = n->GetUniqueId((NcclUniqueId*)id128);
  if (rc != 0) { err = std::string("ncclGetUniqueId: ") + n->GetErrorString(rc); return MFIT_ERR_NCCL; }
  return 0;
}

int mfit_nccl_comm_init(void** comm, int n_ranks, const void* id128, int rank, std::string& err) {
  Nccl* n = nccl();
  if (!n->h) { err = n->err; return MFIT_ERR_NCCL; }
  NcclUniqueId id = *(const NcclUniqueId*)id128;
  int rc = n->CommInitRank((NcclComm*)comm, n_ranks, id, rank);
  if (rc != 0) { err = std::string("ncclCommInitRank: ") + n->GetErrorString(rc); return MFIT_ERR_NCCL; }
  return 0;
}

int mfit_nccl_allgather_bytes(void* comm, const void* send, void* recv, size_t bytes_per_rank, void* stream,
                              std::string& err) {
  Nccl* n = nccl();
  if (!n->h) { err = n->err; return MFIT_ERR_NCCL; }
  int rc = n->AllGather(send, recv, bytes_per_rank, 0 /* ncclInt8/ncclChar */, (NcclComm)comm, stream);
  if (rc != 0) { err = std::string("ncclAllGather: ") + n->GetErrorString(rc); return MFIT_ERR_NCCL; }
  return 0;
}

void mfit_nccl_comm_destroy(void* comm) {
  Nccl* n = nccl();
  if (n->h && comm) n->CommDestroy((NcclComm)comm);
}
