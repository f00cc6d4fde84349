// Stand-in for MPI + NGSolve's parallel containers, for the multi-rank functions of the reference (hybrid smoother, DCC map,
// hybrid matrix split).  TEST INFRASTRUCTURE ONLY (oracle/).
//
// The R "ranks" of a communicator are R host threads of ONE process; a message is a byte copy through a mailbox keyed by
// (source, destination, tag), FIFO per key like MPI's non-overtaking rule.  Sends are eager (the payload is copied when the send
// is started -- legal, MPI forbids touching a send buffer before completion), receives block in Wait.  Persistent requests
// (Send_init / Recv_init / Startall) are what dcc_map.cpp uses.  Written for this repository; see README.md for what it decides.
#pragma once
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>

#include "ngs_standin.hpp"

namespace ngcore {

class World {
  std::mutex m;
  std::condition_variable cv;
  std::map<std::tuple<int, int, int>, std::deque<std::vector<char>>> box;
  std::map<std::tuple<int, int, int>, std::deque<std::shared_ptr<void>>> obox;
  int bar_count = 0, bar_gen = 0;
  std::vector<double> red;
  bool aborted = false;       // a rank failed: wake everybody who waits for it
  void check() const { if (aborted) throw Exception("MPI stand-in: another rank failed"); }

public:
  const int size;
  explicit World(int n) : red(n, 0.0), size(n) {}
  void put(int src, int dst, int tag, const void *p, size_t bytes) {
    std::vector<char> msg((const char *)p, (const char *)p + bytes);
    { std::lock_guard<std::mutex> lk(m); box[{src, dst, tag}].push_back(std::move(msg)); }
    cv.notify_all();
  }
  void get(int src, int dst, int tag, void *p, size_t bytes) {
    std::unique_lock<std::mutex> lk(m);
    auto &q = box[{src, dst, tag}];
    cv.wait(lk, [&] { return aborted || !q.empty(); });
    check();
    if (q.front().size() != bytes) throw Exception("MPI stand-in: message size mismatch");
    std::copy(q.front().begin(), q.front().end(), (char *)p);
    q.pop_front();
  }
  void put_obj(int src, int dst, int tag, std::shared_ptr<void> o) {
    { std::lock_guard<std::mutex> lk(m); obox[{src, dst, tag}].push_back(o); }
    cv.notify_all();
  }
  std::shared_ptr<void> get_obj(int src, int dst, int tag) {
    std::unique_lock<std::mutex> lk(m);
    auto &q = obox[{src, dst, tag}];
    cv.wait(lk, [&] { return aborted || !q.empty(); });
    check();
    auto o = q.front();
    q.pop_front();
    return o;
  }
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    const int gen = bar_gen;
    if (++bar_count == size) { bar_count = 0; bar_gen++; cv.notify_all(); }
    else cv.wait(lk, [&] { return aborted || gen != bar_gen; });
    check();
  }
  void abort() {
    { std::lock_guard<std::mutex> lk(m); aborted = true; }
    cv.notify_all();
  }
  double allreduce_sum(int rank, double v) {       // summed in rank order on every rank
    { std::lock_guard<std::mutex> lk(m); red[rank] = v; }
    barrier();
    double s = 0;
    for (int r = 0; r < size; r++) s += red[r];
    barrier();
    return s;
  }
};

enum NG_MPI_Op { NG_MPI_SUM };
// element size, plus (for NG_MPI_Type_indexed with unit block lengths) the element displacements of a gather type
struct NG_MPI_Datatype {
  size_t bytes = 0;
  std::shared_ptr<std::vector<int>> displs;
};
template <class T> INLINE NG_MPI_Datatype GetMPIType() { return NG_MPI_Datatype{sizeof(T), nullptr}; }
INLINE void NG_MPI_Type_indexed(size_t count, const int *blocklens, const int *displs, NG_MPI_Datatype old, NG_MPI_Datatype *out) {
  for (size_t i = 0; i < count; i++) if (blocklens[i] != 1) throw Exception("MPI stand-in: indexed types with unit blocks only");
  out->bytes = old.bytes;
  out->displs = std::make_shared<std::vector<int>>(displs, displs + count);
}
INLINE void NG_MPI_Type_commit(NG_MPI_Datatype *) {}
INLINE void NG_MPI_Type_free(NG_MPI_Datatype *) {}
struct NG_MPI_Status {};
#define NG_MPI_STATUS_IGNORE ((ngcore::NG_MPI_Status *) nullptr)
constexpr int NG_MPI_TAG_AMG = 1120;

struct MPIRequest {
  enum Kind { SEND, RECV } kind;
  void *buf;
  size_t bytes;
  int peer, tag, me;
  World *w;
  bool active = false, persistent = false;
};
typedef MPIRequest *NG_MPI_Request;
#define NG_MPI_REQUEST_NULL ((ngcore::NG_MPI_Request) nullptr)

INLINE void start_request(NG_MPI_Request r) {
  if (r->kind == MPIRequest::SEND) r->w->put(r->me, r->peer, r->tag, r->buf, r->bytes);   // eager: complete at once
  else r->active = true;
}
INLINE void wait_request(NG_MPI_Request &r) {
  if (!r) return;
  if (r->kind == MPIRequest::RECV && r->active) { r->w->get(r->peer, r->me, r->tag, r->buf, r->bytes); r->active = false; }
  if (!r->persistent) { delete r; r = nullptr; }
}
INLINE void MyMPI_WaitAll(FlatArray<NG_MPI_Request> reqs) { for (auto &r : reqs) wait_request(r); }

class NgMPI_Comm {
public:
  World *w = nullptr;
  int rank = 0;
  NgMPI_Comm() = default;
  NgMPI_Comm(World *aw, int r) : w(aw), rank(r) {}
  int Rank() const { return rank; }
  int Size() const { return w ? w->size : 1; }
  bool isValid() const { return w != nullptr; }
  template <class T> T AllReduce(T v, NG_MPI_Op) const { return T(w->allreduce_sum(rank, double(v))); }
  template <class T> NG_MPI_Request ISend(FlatArray<T> a, int dest, int tag) const {
    w->put(rank, dest, tag, a.Data(), a.Size() * sizeof(T));
    return NG_MPI_REQUEST_NULL;
  }
  template <class T> NG_MPI_Request IRecv(FlatArray<T> a, int src, int tag) const {
    return new MPIRequest{MPIRequest::RECV, a.Data(), a.Size() * sizeof(T), src, tag, rank, w, true, false};
  }
  // whole objects (the hybrid split ships small sparse matrices): a deep copy travels
  template <class OBJ> NG_MPI_Request ISend(const OBJ &o, int dest, int tag) const {
    w->put_obj(rank, dest, tag, std::make_shared<OBJ>(o));
    return NG_MPI_REQUEST_NULL;
  }
  template <class OBJ> void Recv(std::shared_ptr<OBJ> &o, int src, int tag) const { o = std::static_pointer_cast<OBJ>(w->get_obj(src, rank, tag)); }
  template <class OBJ> void Send(const OBJ &o, int dest, int tag) const { w->put_obj(rank, dest, tag, std::make_shared<OBJ>(o)); }
  template <class T> void Recv(FlatArray<T> a, int src, int tag) const { w->get(src, rank, tag, a.Data(), a.Size() * sizeof(T)); }
};
typedef NgMPI_Comm NgsAMG_Comm;

INLINE void NG_MPI_Send_init(void *buf, size_t count, NG_MPI_Datatype t, int dest, int tag, const NgMPI_Comm &c, NG_MPI_Request *req) {
  *req = new MPIRequest{MPIRequest::SEND, buf, count * t.bytes, dest, tag, c.rank, c.w, false, true};
}
INLINE void NG_MPI_Recv_init(void *buf, size_t count, NG_MPI_Datatype t, int src, int tag, const NgMPI_Comm &c, NG_MPI_Request *req) {
  *req = new MPIRequest{MPIRequest::RECV, buf, count * t.bytes, src, tag, c.rank, c.w, false, true};
}
INLINE void NG_MPI_Startall(size_t n, NG_MPI_Request *reqs) { for (size_t i = 0; i < n; i++) start_request(reqs[i]); }

// blocking / immediate raw sends and receives; a send with an indexed type gathers `count` x (the listed elements) first
INLINE std::vector<char> pack_typed(const void *buf, size_t count, const NG_MPI_Datatype &t) {
  if (!t.displs) return std::vector<char>((const char *)buf, (const char *)buf + count * t.bytes);
  std::vector<char> out;
  out.reserve(count * t.displs->size() * t.bytes);
  for (size_t c = 0; c < count; c++)
    for (int d : *t.displs) out.insert(out.end(), (const char *)buf + size_t(d) * t.bytes, (const char *)buf + size_t(d + 1) * t.bytes);
  return out;
}
INLINE void NG_MPI_Send(const void *buf, size_t count, NG_MPI_Datatype t, int dest, int tag, const NgMPI_Comm &c) {
  auto msg = pack_typed(buf, count, t);
  c.w->put(c.rank, dest, tag, msg.data(), msg.size());
}
INLINE void NG_MPI_Isend(const void *buf, size_t count, NG_MPI_Datatype t, int dest, int tag, const NgMPI_Comm &c, NG_MPI_Request *req) {
  NG_MPI_Send(buf, count, t, dest, tag, c);      // eager
  *req = NG_MPI_REQUEST_NULL;
}
INLINE void NG_MPI_Recv(void *buf, size_t count, NG_MPI_Datatype t, int src, int tag, const NgMPI_Comm &c, NG_MPI_Status *) {
  if (t.displs) throw Exception("MPI stand-in: receive into an indexed type");
  c.w->get(src, c.rank, tag, buf, count * t.bytes);
}
// completes the active request with the LOWEST index (real MPI: whichever arrives first -- the stand-in is deterministic)
INLINE void NG_MPI_Waitany(int n, NG_MPI_Request *reqs, int *index, NG_MPI_Status *) {
  for (int i = 0; i < n; i++)
    if (reqs[i]) { wait_request(reqs[i]); *index = i; return; }
  *index = -1;
}

// ---- tables ---------------------------------------------------------------------------------------------------------
template <class T> class Table {
  std::vector<size_t> first{0};
  std::vector<T> data;

public:
  Table() = default;
  explicit Table(const FlatArray<int> &perrow) {
    first.assign(perrow.Size() + 1, 0);
    for (size_t i = 0; i < perrow.Size(); i++) first[i + 1] = first[i] + size_t(perrow[i]);
    data.resize(first.back());
  }
  size_t Size() const { return first.size() - 1; }
  FlatArray<T> operator[](size_t i) const { return FlatArray<T>(first[i + 1] - first[i], const_cast<T *>(data.data()) + first[i]); }
  struct It {
    const Table *t;
    size_t i;
    FlatArray<T> operator*() const { return (*t)[i]; }
    It &operator++() { ++i; return *this; }
    bool operator!=(const It &o) const { return i != o.i; }
  };
  It begin() const { return It{this, 0}; }
  It end() const { return It{this, Size()}; }
};
template <class T> class TableCreator {   // two passes over the same loop: count, then fill
  int mode = 2;
  size_t nrows;
  Array<int> cnt;
  Table<T> tab;

public:
  explicit TableCreator(size_t n) : nrows(n), cnt(n) { cnt = 0; }
  bool Done() const { return mode > 3; }
  void operator++(int) {
    if (mode == 2) { tab = Table<T>(cnt); cnt = 0; }
    mode++;
  }
  void Add(size_t row, const T &v) {
    if (mode == 2) cnt[row]++;
    else tab[row][cnt[row]++] = v;
  }
  void Add(size_t row, FlatArray<T> vs) { for (auto &v : vs) Add(row, v); }
  Table<T> MoveTable() { return std::move(tab); }
};
}  // namespace ngcore

namespace ngla {
// who shares which dof: per dof the ascending list of the OTHER ranks holding it; per neighbour the ascending list of shared dofs
// (ascending on both sides = the same order on both sides for the partitions used here; the harness checks it)
class ParallelDofs {
  NgMPI_Comm comm;
  size_t ndof;
  int es;
  Array<int> procs;
  Table<int> dof_procs, ex_dofs;

public:
  ParallelDofs(NgMPI_Comm c, size_t n, int entrysize, FlatArray<int> peers, const Table<int> &exdofs) : comm(c), ndof(n), es(entrysize) {
    procs.SetSize(peers.Size());
    for (size_t k = 0; k < peers.Size(); k++) procs[k] = peers[k];
    for (size_t k = 1; k < procs.Size(); k++) if (procs[k - 1] >= procs[k]) throw Exception("ParallelDofs: neighbours must be ascending");
    Array<int> cnt(exdofs.Size());
    for (size_t k = 0; k < exdofs.Size(); k++) cnt[k] = int(exdofs[k].Size());
    ex_dofs = Table<int>(cnt);
    TableCreator<int> c_dp(n);
    for (; !c_dp.Done(); c_dp++)
      for (size_t k = 0; k < exdofs.Size(); k++)
        for (size_t j = 0; j < exdofs[k].Size(); j++) {
          ex_dofs[k][j] = exdofs[k][j];
          if (j && exdofs[k][j - 1] >= exdofs[k][j]) throw Exception("ParallelDofs: exchange dofs must be ascending");
          c_dp.Add(exdofs[k][j], procs[k]);
        }
    dof_procs = c_dp.MoveTable();
  }
  const NgMPI_Comm &GetCommunicator() const { return comm; }
  size_t GetNDofLocal() const { return ndof; }
  int GetEntrySize() const { return es; }
  FlatArray<int> GetDistantProcs() const { return procs; }
  FlatArray<int> GetDistantProcs(size_t dof) const { return dof_procs[dof]; }
  FlatArray<int> GetExchangeDofs(int proc) const {
    for (size_t k = 0; k < procs.Size(); k++) if (procs[k] == proc) return ex_dofs[k];
    return FlatArray<int>();
  }
};
}  // namespace ngla
