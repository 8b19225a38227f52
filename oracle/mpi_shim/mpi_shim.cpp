// oracle/mpi_shim/mpi_shim.cpp -- TEST INFRASTRUCTURE, not product code.
// Thread-backed implementation of the MPI subset declared in mpi.h: ranks are
// std::threads of one process, point-to-point messages go through per-rank
// mailboxes, collectives through a generation-counted barrier.
#include "mpi.h"

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace {

struct Message {
  int source;
  int tag;
  std::vector<char> payload;
};

struct Mailbox {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Message> queue;
};

struct World {
  int size = 1;
  std::vector<Mailbox> boxes;
  // barrier
  std::mutex bmu;
  std::condition_variable bcv;
  int arrived = 0;
  long generation = 0;
  // allreduce staging: one slot per rank
  std::vector<const void *> contrib;
};

World *g_world = nullptr;
thread_local int t_rank = 0;

size_t type_size(MPI_Datatype dt) { return dt == MPI_INT ? sizeof(int) : sizeof(double); }

void barrier(World &w) {
  std::unique_lock<std::mutex> lk(w.bmu);
  long gen = w.generation;
  if (++w.arrived == w.size) {
    w.arrived = 0;
    ++w.generation;
    w.bcv.notify_all();
  } else {
    w.bcv.wait(lk, [&] { return w.generation != gen; });
  }
}

template <typename T>
void reduce_in_rank_order(World &w, T *out, int count, MPI_Op op) {
  for (int i = 0; i < count; ++i) {
    T acc = static_cast<const T *>(w.contrib[0])[i];
    for (int r = 1; r < w.size; ++r) {
      T v = static_cast<const T *>(w.contrib[r])[i];
      if (op == MPI_SUM) acc = acc + v;
      else if (op == MPI_MIN) acc = v < acc ? v : acc;
      else acc = v > acc ? v : acc;
    }
    out[i] = acc;
  }
}

}  // namespace

extern "C" {

int MPI_Comm_size(MPI_Comm, int *size) {
  *size = g_world ? g_world->size : 1;
  return MPI_SUCCESS;
}

int MPI_Comm_rank(MPI_Comm, int *rank) {
  *rank = g_world ? t_rank : 0;
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm) {
  if (g_world && g_world->size > 1) barrier(*g_world);
  return MPI_SUCCESS;
}

int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype dt, MPI_Op op, MPI_Comm) {
  if (!g_world || g_world->size == 1) {
    std::memcpy(recvbuf, sendbuf, type_size(dt) * count);
    return MPI_SUCCESS;
  }
  World &w = *g_world;
  w.contrib[t_rank] = sendbuf;
  barrier(w);  // every contribution is visible
  // Each rank reduces into a private temporary first: the reference calls
  // MPI_Allreduce(tmp_neighbors, tmp_buffer, ...) where recvbuf of one rank is
  // never another rank's sendbuf, but in-place safety costs nothing here.
  std::vector<char> tmp(type_size(dt) * count);
  if (dt == MPI_INT) reduce_in_rank_order<int>(w, reinterpret_cast<int *>(tmp.data()), count, op);
  else reduce_in_rank_order<double>(w, reinterpret_cast<double *>(tmp.data()), count, op);
  barrier(w);  // nobody still reads the send buffers
  std::memcpy(recvbuf, tmp.data(), tmp.size());
  return MPI_SUCCESS;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm, MPI_Request *req) {
  req->buf = buf;
  req->count = count;
  req->datatype = dt;
  req->source = source;
  req->tag = tag;
  req->active = 1;
  return MPI_SUCCESS;
}

int MPI_Send(const void *buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm) {
  World &w = *g_world;
  Message m;
  m.source = t_rank;
  m.tag = tag;
  m.payload.assign(static_cast<const char *>(buf), static_cast<const char *>(buf) + type_size(dt) * count);
  Mailbox &box = w.boxes[dest];
  {
    std::lock_guard<std::mutex> lk(box.mu);
    box.queue.push_back(std::move(m));
  }
  box.cv.notify_all();
  return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request *req, MPI_Status *status) {
  if (!req->active) return MPI_SUCCESS;
  World &w = *g_world;
  Mailbox &box = w.boxes[t_rank];
  std::unique_lock<std::mutex> lk(box.mu);
  for (;;) {
    for (auto it = box.queue.begin(); it != box.queue.end(); ++it) {
      if (it->tag == req->tag && (req->source == MPI_ANY_SOURCE || req->source == it->source)) {
        size_t want = type_size(req->datatype) * req->count;
        size_t n = it->payload.size() < want ? it->payload.size() : want;
        if (n) std::memcpy(req->buf, it->payload.data(), n);
        if (status) {
          status->MPI_SOURCE = it->source;
          status->MPI_TAG = it->tag;
          status->MPI_ERROR = MPI_SUCCESS;
        }
        box.queue.erase(it);
        req->active = 0;
        return MPI_SUCCESS;
      }
    }
    box.cv.wait(lk);
  }
}

double MPI_Wtime(void) {
  using clk = std::chrono::steady_clock;
  static const clk::time_point t0 = clk::now();
  return std::chrono::duration<double>(clk::now() - t0).count();
}

int MPI_Finalize(void) { return MPI_SUCCESS; }

void hpccg_shim_run(int size, void (*fn)(int, void *), void *arg) {
  World w;
  w.size = size;
  w.boxes = std::vector<Mailbox>(size);
  w.contrib.assign(size, nullptr);
  g_world = &w;
  std::vector<std::thread> threads;
  for (int r = 0; r < size; ++r)
    threads.emplace_back([=] {
      t_rank = r;
      fn(r, arg);
    });
  for (auto &t : threads) t.join();
  g_world = nullptr;
}

}  // extern "C"
