// context.cu -- rank context, set-up transports (in-process world) and the NCCL communicator.
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.hpp"
#include "context.hpp"
#include "device_matrix.hpp"

namespace hpccg {

RankContext &ctx() {
  static thread_local RankContext c;
  return c;
}

int ctx_allgather(const void *send, long long nbytes, void *recv) {
  RankContext &c = ctx();
  if (c.size == 1) {
    if (nbytes > 0) std::memcpy(recv, send, (size_t)nbytes);
    return 0;
  }
  if (!c.allgather) return fail(HPCCG_ERR_STATE, "rank context has size %d but no set-up collective (hpccg_ctx_set_allgather)", c.size);
  int rc = c.allgather(c.allgather_user, send, nbytes, recv);
  if (rc != 0) return fail(HPCCG_ERR_COMM, "set-up allgather failed with %d", rc);
  return 0;
}

// ---- in-process world: ranks are host threads ---------------------------------------------------------
struct LocalWorld {
  int size = 1;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  long generation = 0;
  std::vector<const void *> contrib;
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const long gen = generation;
    if (++arrived == size) {
      arrived = 0;
      ++generation;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return generation != gen; });
    }
  }
};

struct LocalBinding {
  LocalWorld *world;
  int rank;
};

static int local_allgather(void *user, const void *send, long long nbytes, void *recv) {
  LocalBinding *b = static_cast<LocalBinding *>(user);
  LocalWorld &w = *b->world;
  w.contrib[b->rank] = send;
  w.barrier();
  for (int r = 0; r < w.size; ++r)
    if (nbytes > 0) std::memcpy(static_cast<char *>(recv) + (size_t)r * nbytes, w.contrib[r], (size_t)nbytes);
  w.barrier();
  return 0;
}

// ---- NCCL through dlopen ----------------------------------------------------------------------------------
struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static ncclComm_t g_comm = nullptr;
static int g_comm_rank = 0, g_comm_size = 1;

static void load_nccl() {
  std::call_once(g_nccl_once, [] {
    // torch bundles its own libnccl.so.2; when it is already mapped dlopen returns that one.
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return;
#define HPCCG_SYM(field, name) \
  *reinterpret_cast<void **>(&g_nccl.field) = dlsym(g_nccl.handle, name); \
  if (!g_nccl.field) return;
    HPCCG_SYM(GetUniqueId, "ncclGetUniqueId")
    HPCCG_SYM(CommInitRank, "ncclCommInitRank")
    HPCCG_SYM(CommDestroy, "ncclCommDestroy")
    HPCCG_SYM(Send, "ncclSend")
    HPCCG_SYM(Recv, "ncclRecv")
    HPCCG_SYM(GroupStart, "ncclGroupStart")
    HPCCG_SYM(GroupEnd, "ncclGroupEnd")
    HPCCG_SYM(AllGather, "ncclAllGather")
    HPCCG_SYM(GetErrorString, "ncclGetErrorString")
#undef HPCCG_SYM
    g_nccl.ok = true;
  });
}

static int fail_nccl(ncclResult_t r, const char *what) {
  return fail(HPCCG_ERR_NCCL, "NCCL error %d (%s) in %s", (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
}

#define HPCCG_NCCL(call)                                  \
  do {                                                    \
    ncclResult_t r_ = (call);                             \
    if (r_ != ncclSuccess) return fail_nccl(r_, #call);   \
  } while (0)

bool nccl_ready() { return g_comm != nullptr; }
int nccl_rank() { return g_comm_rank; }
int nccl_size() { return g_comm_size; }

int nccl_allgather_double(double *buf, cudaStream_t stream) {
  if (!g_comm) return fail(HPCCG_ERR_STATE, "no NCCL communicator");
  HPCCG_NCCL(g_nccl.AllGather(buf + g_comm_rank, buf, 1, ncclDouble, g_comm, stream));
  return 0;
}

int nccl_halo_exchange(const double *send_buffer, const int *send_length, double *recv_base, const int *recv_length,
                       const int *neighbors, int num_neighbors, cudaStream_t stream) {
  if (!g_comm) return fail(HPCCG_ERR_STATE, "no NCCL communicator");
  if (num_neighbors == 0) return 0;
  HPCCG_NCCL(g_nccl.GroupStart());
  const double *sp = send_buffer;
  double *rp = recv_base;
  for (int i = 0; i < num_neighbors; ++i) {
    if (recv_length[i] > 0) HPCCG_NCCL(g_nccl.Recv(rp, (size_t)recv_length[i], ncclDouble, neighbors[i], g_comm, stream));
    if (send_length[i] > 0) HPCCG_NCCL(g_nccl.Send(sp, (size_t)send_length[i], ncclDouble, neighbors[i], g_comm, stream));
    rp += recv_length[i];
    sp += send_length[i];
  }
  HPCCG_NCCL(g_nccl.GroupEnd());
  return 0;
}

int nccl_allgather_host(const void *send, long long nbytes, void *recv) {
  if (!g_comm) return fail(HPCCG_ERR_STATE, "no NCCL communicator");
  char *d = nullptr;
  HPCCG_CUDA(cudaMalloc(&d, (size_t)nbytes * (g_comm_size + 1)));
  cudaError_t e = cudaMemcpy(d, send, (size_t)nbytes, cudaMemcpyHostToDevice);
  int rc = 0;
  if (e != cudaSuccess) rc = fail_cuda(e, "allgather staging", __FILE__, __LINE__);
  if (!rc) {
    ncclResult_t r = g_nccl.AllGather(d, d + nbytes, (size_t)nbytes, ncclChar, g_comm, nullptr);
    if (r != ncclSuccess) rc = fail_nccl(r, "ncclAllGather (bytes)");
  }
  if (!rc) {
    e = cudaMemcpy(recv, d + nbytes, (size_t)nbytes * g_comm_size, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = fail_cuda(e, "allgather readback", __FILE__, __LINE__);
  }
  cudaFree(d);
  return rc;
}

// ---- peer-memory link -----------------------------------------------------------------------------------------
struct PeerCard {  // what every rank publishes about itself
  cudaIpcMemHandle_t mailbox, p;
  int ok;                 // this rank could allocate / export
  int pci[3];             // domain, bus, device of this rank's GPU: two ranks on one GPU must never wait for each other in a kernel
  int n, nnb;
  int neighbors[kMaxPeerNb], recv_length[kMaxPeerNb], send_length[kMaxPeerNb];
};

void peer_link_destroy(hpccg_dev_matrix *m) {
  for (void *p : m->ipc_opened) cudaIpcCloseMemHandle(p);
  m->ipc_opened.clear();
  if (m->peer_link) cudaFree(m->peer_link);
  if (m->mailbox) cudaFree(m->mailbox);
  m->peer_link = nullptr;
  m->mailbox = nullptr;
}

int peer_link_create(hpccg_dev_matrix *m, bool eligible) {
  if (m->peer_tried) return 0;
  m->peer_tried = 1;
  const int R = g_comm_size, me = g_comm_rank;
  PeerCard mine;
  std::memset(&mine, 0, sizeof mine);
  mine.ok = (eligible && R <= kMaxRanks && m->num_neighbors <= kMaxPeerNb && m->p != nullptr) ? 1 : 0;
  if (const char *e = std::getenv("HPCCG_B200_COMM"))
    if (std::string(e) == "nccl") mine.ok = 0;  // A/B switch: NCCL send/recv + gathers
  if (mine.ok) {
    if (cudaMalloc(&m->mailbox, sizeof(Mailbox)) != cudaSuccess || cudaMemset(m->mailbox, 0, sizeof(Mailbox)) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine.mailbox, m->mailbox) != cudaSuccess || cudaIpcGetMemHandle(&mine.p, m->p) != cudaSuccess) {
      cudaGetLastError();
      mine.ok = 0;
    }
  }
  {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess) {
      mine.pci[0] = prop.pciDomainID;
      mine.pci[1] = prop.pciBusID;
      mine.pci[2] = prop.pciDeviceID;
    } else {
      cudaGetLastError();
      mine.ok = 0;
    }
  }
  mine.n = m->n;
  mine.nnb = m->num_neighbors;
  for (int i = 0; i < m->num_neighbors && i < kMaxPeerNb; ++i) {
    mine.neighbors[i] = m->neighbors[i];
    mine.recv_length[i] = m->recv_length[i];
    mine.send_length[i] = m->send_length[i];
  }
  HPCCG_CUDA(cudaDeviceSynchronize());  // the memset above is complete before any peer can write
  std::vector<PeerCard> cards(R);
  HPCCG_TRY(nccl_allgather_host(&mine, sizeof mine, cards.data()));
  bool all_ok = true;
  for (const PeerCard &c : cards) all_ok = all_ok && c.ok;
  // Kernels of different ranks wait for each other through the mailboxes; that is only safe when every rank has its
  // own GPU (two waiting kernels on one GPU are not guaranteed to run concurrently).
  for (int a = 0; a < R && all_ok; ++a)
    for (int b = a + 1; b < R; ++b)
      if (std::memcmp(cards[a].pci, cards[b].pci, sizeof cards[a].pci) == 0) all_ok = false;

  PeerLink h;
  std::memset(&h, 0, sizeof h);
  h.rank = me;
  h.size = R;
  h.nnb = m->num_neighbors;
  std::vector<void *> p_base(R, nullptr);
  if (all_ok) {
    for (int r = 0; r < R && all_ok; ++r) {
      if (r == me) {
        h.box[r] = m->mailbox;
        continue;
      }
      void *ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, cards[r].mailbox, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        all_ok = false;
        break;
      }
      m->ipc_opened.push_back(ptr);
      h.box[r] = static_cast<Mailbox *>(ptr);
    }
    int seg = 0;
    for (int i = 0; i < m->num_neighbors && all_ok; ++i) {
      const int q = m->neighbors[i];
      const PeerCard &c = cards[q];
      int slot = -1;
      long long off = c.n;
      for (int t = 0; t < c.nnb; ++t) {
        if (c.neighbors[t] == me) {
          slot = t;
          break;
        }
        off += c.recv_length[t];
      }
      if (slot < 0 || c.recv_length[slot] != m->send_length[i])
        return fail(HPCCG_ERR_STATE, "halo plans of ranks %d and %d do not match", me, q);
      if (!p_base[q]) {
        if (cudaIpcOpenMemHandle(&p_base[q], c.p, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          all_ok = false;
          break;
        }
        m->ipc_opened.push_back(p_base[q]);
      }
      h.nb_dst[i] = static_cast<double *>(p_base[q]) + off;
      h.nb_flag[i] = &h.box[q]->halo_seq[slot];
      h.seg_start[i] = seg;
      seg += m->send_length[i];
    }
    for (int i = m->num_neighbors; i <= kMaxPeerNb; ++i) h.seg_start[i] = seg;
  }
  // every rank must take the same decision: one more round
  int mine_ok = all_ok ? 1 : 0;
  std::vector<int> oks(R, 0);
  HPCCG_TRY(nccl_allgather_host(&mine_ok, sizeof(int), oks.data()));
  for (int v : oks) all_ok = all_ok && v;
  if (!all_ok) {
    peer_link_destroy(m);
    return 0;
  }
  HPCCG_CUDA(cudaMalloc(&m->peer_link, sizeof(PeerLink)));
  HPCCG_CUDA(cudaMemcpy(m->peer_link, &h, sizeof h, cudaMemcpyHostToDevice));
  // the put folded into the p-producing kernel receives the link and the remote destinations as kernel parameters
  m->put_plan.link = m->peer_link;
  for (int i = 0; i < kMaxPeerNb; ++i) m->put_plan.dst[i] = h.nb_dst[i];
  return 0;
}

}  // namespace hpccg

using namespace hpccg;

extern "C" {

int hpccg_ctx_set(int rank, int size) {
  if (size < 1 || rank < 0 || rank >= size) return fail(HPCCG_ERR_ARG, "hpccg_ctx_set: rank %d of %d", rank, size);
  ctx().rank = rank;
  ctx().size = size;
  return 0;
}

int hpccg_ctx_get(int *rank, int *size) {
  if (rank) *rank = ctx().rank;
  if (size) *size = ctx().size;
  return 0;
}

int hpccg_ctx_set_allgather(hpccg_allgather_fn fn, void *user) {
  ctx().allgather = fn;
  ctx().allgather_user = user;
  return 0;
}

int hpccg_local_world_create(int size, void **world) {
  if (size < 1 || !world) return fail(HPCCG_ERR_ARG, "hpccg_local_world_create: bad argument");
  LocalWorld *w = new LocalWorld();
  w->size = size;
  w->contrib.assign(size, nullptr);
  *world = w;
  return 0;
}

int hpccg_local_world_bind(void *world, int rank) {
  LocalWorld *w = static_cast<LocalWorld *>(world);
  if (!w || rank < 0 || rank >= w->size) return fail(HPCCG_ERR_ARG, "hpccg_local_world_bind: bad argument");
  static thread_local LocalBinding binding;
  binding.world = w;
  binding.rank = rank;
  ctx().rank = rank;
  ctx().size = w->size;
  ctx().allgather = local_allgather;
  ctx().allgather_user = &binding;
  return 0;
}

int hpccg_local_world_destroy(void *world) {
  delete static_cast<LocalWorld *>(world);
  return 0;
}

int hpccg_nccl_available(void) {
  load_nccl();
  return g_nccl.ok ? 1 : 0;
}

int hpccg_nccl_unique_id(void *id128) {
  load_nccl();
  if (!g_nccl.ok) return fail(HPCCG_ERR_NCCL, "libnccl.so.2 could not be loaded");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  HPCCG_NCCL(g_nccl.GetUniqueId(static_cast<ncclUniqueId *>(id128)));
  return 0;
}

int hpccg_nccl_init(const void *id128, int rank, int size) {
  load_nccl();
  if (!g_nccl.ok) return fail(HPCCG_ERR_NCCL, "libnccl.so.2 could not be loaded");
  if (g_comm) return fail(HPCCG_ERR_STATE, "NCCL communicator already initialised");
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof id);
  HPCCG_NCCL(g_nccl.CommInitRank(&g_comm, size, id, rank));
  g_comm_rank = rank;
  g_comm_size = size;
  ctx().rank = rank;
  ctx().size = size;
  return 0;
}

int hpccg_nccl_finalize(void) {
  if (g_comm) {
    g_nccl.CommDestroy(g_comm);
    g_comm = nullptr;
    g_comm_rank = 0;
    g_comm_size = 1;
  }
  return 0;
}

}  // extern "C"
