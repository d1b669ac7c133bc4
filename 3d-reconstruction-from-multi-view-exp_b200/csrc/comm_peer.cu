// Point-sharded exchange over NVLink peer memory: the two sums an inner
// solve needs from all ranks -- the partial reduced system [sum Y Y^T | rhs row | U | dF] and the
// trial cost -- are formed by this library's own kernels writing straight into the other GPUs'
// memory, not by a host-driven collective.  One process per GPU; every rank maps every other
// rank's exchange window (CUDA IPC), so a solve stays ONE fixed kernel sequence on ONE stream and
// is captured in the same CUDA graph as the single-GPU loop.
//
// Window of a rank (one cudaMalloc, exported with cudaIpcGetMemHandle):
//     [ header: flags, epochs, the small slots | red (the engine's reduce buffer) |
//       stage[world] (one copy of the red layout per source rank) ]
//
// Sum of `red` (deterministic two-shot: reduce-scatter + all-gather):
//   The buffer is cut into segments: row r of P restricted to its lower triangle (columns 0..r;
//   nothing above the diagonal is ever read, which halves the traffic), and rows of n_pad doubles
//   of the U/dF tail.  Segments are dealt to the ranks in blocks of kSegBlock.
//   comm_push_kernel    every rank stores its copy of each segment it does not own into the
//                       owner's stage[rank] (16-byte peer stores over NVLink); the last CTA to
//                       finish publishes flag_push[rank] = epoch on every peer.
//   comm_reduce_kernel  the owner waits for all flag_push, adds the copies in RANK ORDER (own copy
//                       from its red) -- so every rank receives the same bits, and the same bits
//                       as any other run with that world size -- and stores the sum into every
//                       rank's red; the last CTA publishes flag_bcast[rank] = epoch everywhere.
//   comm_wait_kernel    waits for all flag_bcast and advances the epoch.
//   Hazards: a peer overwrites my stage / red only after it has seen my flags of the previous
//   epoch, and I leave an exchange only after every peer has finished reading what I sent.
// Sum of the cost (comm_small_kernel): every rank stores (cost, singular flag) into slot [rank] of
//   every peer, flags it, and adds the slots in rank order.  A singular point block on one rank
//   thereby stops all ranks in the same solve (the reference raises LinAlgError, :128).
//
// Flags are monotone 64-bit epochs written with st.release.sys after a system-scope fence and
// polled with ld.acquire.sys; staged data is read with ld.global.cg (L2 only).  Every spin has a
// wall-clock limit: a missing peer turns into BA_ERR_COMM instead of a hung GPU.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "ba_common.cuh"

namespace ba {

constexpr int kSegBlock = 4;  // consecutive segments dealt to one rank
__device__ unsigned long long g_spin_limit_ns = 30ull * 1000000000ull;  // BA_COMM_SPIN_LIMIT_S overrides

struct CommHeader {
  unsigned long long flag_push[kMaxRanks];      // [src]   src's copies for me have landed
  unsigned long long flag_bcast[kMaxRanks];     // [owner] owner's sums have landed in my red
  unsigned long long flag_small[2][kMaxRanks];  // [parity][src]
  double small[2][kMaxRanks][2];                // [parity][src] (cost, singular flag)
  unsigned long long flag_chol[kMaxRanks];      // [src]   src has finished its share of a trailing update
  unsigned long long epoch_chol;                // completed trailing-update exchanges (local)
  unsigned long long epoch;                     // completed sums of red (local)
  unsigned long long epoch_small;               // completed small sums (local)
  unsigned int count_push, count_reduce;        // last-CTA-done counters (local)
};
constexpr size_t kHeaderBytes = 4096;
static_assert(sizeof(CommHeader) <= kHeaderBytes, "header region too small");

struct CommDev {
  int rank, world;
  int n_pad, n_seg;
  int64_t red_len;
  CommHeader* hdr[kMaxRanks];
  double* red[kMaxRanks];
  double* stage[kMaxRanks];  // stage[o] + src * red_len
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Spin until *flag >= want; false on timeout.
__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long want) {
  if (ld_acquire_sys(flag) >= want) return true;
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(flag) < want) {
    __nanosleep(64);
    if (global_ns() - t0 > g_spin_limit_ns) return false;
  }
  return true;
}

__device__ __forceinline__ void comm_fail(ba_lm_state* ctl) {
  atomicExch(&ctl->status, (int)BA_ERR_COMM);
  atomicExch(&ctl->done, 1);
}

__device__ __forceinline__ void seg_range(int s, int n_pad, int64_t red_len, int64_t& start, int& len) {
  if (s < n_pad) {
    start = (int64_t)s * n_pad;
    const int l = (s + 2) & ~1;  // columns 0..s, rounded up to a 16-byte piece
    len = l < n_pad ? l : n_pad;
  } else {
    start = (int64_t)n_pad * n_pad + (int64_t)(s - n_pad) * n_pad;
    const int64_t rem = red_len - start;
    len = rem < n_pad ? (int)rem : n_pad;
  }
}

// True for exactly one CTA of the grid: the last one to arrive.  All of the grid's global and peer
// stores issued before the call are then visible system-wide (fence cumulativity).
__device__ __forceinline__ bool last_cta_done(unsigned int* counter) {
  __shared__ bool s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int prev = atomicAdd(counter, 1u);
    s_last = prev == gridDim.x - 1;
    if (s_last) {
      *counter = 0;
      __threadfence_system();
    }
  }
  __syncthreads();
  return s_last;
}

__global__ void __launch_bounds__(256)
comm_push_kernel(CommDev cd, ba_lm_state* ctl, int use_ctl) {
  if (use_ctl && ctl->done) return;
  CommHeader* me = cd.hdr[cd.rank];
  const unsigned long long epoch = me->epoch + 1;
  const int s = blockIdx.x;
  const int owner = (s / kSegBlock) % cd.world;
  if (owner != cd.rank) {
    int64_t start;
    int len;
    seg_range(s, cd.n_pad, cd.red_len, start, len);
    const double2* src = reinterpret_cast<const double2*>(cd.red[cd.rank] + start);
    double2* dst = reinterpret_cast<double2*>(cd.stage[owner] + (int64_t)cd.rank * cd.red_len + start);
    for (int q = threadIdx.x; q < len / 2; q += blockDim.x) dst[q] = src[q];
  }
  if (last_cta_done(&me->count_push)) {
    if (threadIdx.x < cd.world && threadIdx.x != cd.rank)
      st_release_sys(&cd.hdr[threadIdx.x]->flag_push[cd.rank], epoch);
  }
}

__global__ void __launch_bounds__(256)
comm_reduce_kernel(CommDev cd, ba_lm_state* ctl, int use_ctl) {
  if (use_ctl && ctl->done) return;
  CommHeader* me = cd.hdr[cd.rank];
  const unsigned long long epoch = me->epoch + 1;
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < cd.world && threadIdx.x != cd.rank)
    if (!spin_until(&me->flag_push[threadIdx.x], epoch)) s_ok = 0;
  __syncthreads();
  if (!s_ok) {
    if (threadIdx.x == 0) comm_fail(ctl);
  } else {
    // k-th segment owned by this rank
    const int k = blockIdx.x;
    const int s = ((k / kSegBlock) * cd.world + cd.rank) * kSegBlock + k % kSegBlock;
    if (s < cd.n_seg) {
      int64_t start;
      int len;
      seg_range(s, cd.n_pad, cd.red_len, start, len);
      const double* stage = cd.stage[cd.rank];
      for (int q = threadIdx.x; q < len / 2; q += blockDim.x) {
        const int64_t idx = start + 2 * q;
        double2 acc = make_double2(0.0, 0.0);
        for (int src = 0; src < cd.world; ++src) {
          double2 v;
          if (src == cd.rank) v = *reinterpret_cast<const double2*>(cd.red[cd.rank] + idx);
          else v = __ldcg(reinterpret_cast<const double2*>(stage + (int64_t)src * cd.red_len + idx));
          if (src == 0) acc = v;
          else { acc.x += v.x; acc.y += v.y; }
        }
        for (int p = 0; p < cd.world; ++p) *reinterpret_cast<double2*>(cd.red[p] + idx) = acc;
      }
    }
  }
  if (last_cta_done(&me->count_reduce)) {
    if (threadIdx.x < cd.world && threadIdx.x != cd.rank)
      st_release_sys(&cd.hdr[threadIdx.x]->flag_bcast[cd.rank], epoch);
  }
}

__global__ void comm_wait_kernel(CommDev cd, ba_lm_state* ctl, int use_ctl) {
  if (use_ctl && ctl->done) return;
  CommHeader* me = cd.hdr[cd.rank];
  const unsigned long long epoch = me->epoch + 1;
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < cd.world && threadIdx.x != cd.rank)
    if (!spin_until(&me->flag_bcast[threadIdx.x], epoch)) s_ok = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!s_ok) comm_fail(ctl);
    me->epoch = epoch;
  }
}

// cost_buf[slot] <- sum over ranks (rank order); singular flag OR-ed into every rank's status.
__global__ void comm_small_kernel(CommDev cd, double* cost_buf, int slot, ba_lm_state* ctl, int use_ctl) {
  if (use_ctl && ctl->done) return;
  CommHeader* me = cd.hdr[cd.rank];
  const unsigned long long epoch = me->epoch_small + 1;
  const int par = (int)(epoch & 1ull);
  const int t = threadIdx.x;
  __shared__ int s_ok;
  if (t == 0) s_ok = 1;
  __syncthreads();
  const double mine = cost_buf[slot];
  const double flag = ctl->status == BA_ERR_SINGULAR ? 1.0 : 0.0;
  if (t < cd.world && t != cd.rank) {
    CommHeader* peer = cd.hdr[t];
    *reinterpret_cast<double2*>(&peer->small[par][cd.rank][0]) = make_double2(mine, flag);
    __threadfence_system();
    st_release_sys(&peer->flag_small[par][cd.rank], epoch);
    if (!spin_until(&me->flag_small[par][t], epoch)) s_ok = 0;
  }
  __syncthreads();
  if (t == 0) {
    if (!s_ok) {
      comm_fail(ctl);
    } else {
      double sum = 0.0, bad = 0.0;
      for (int r = 0; r < cd.world; ++r) {
        double2 v;
        if (r == cd.rank) v = make_double2(mine, flag);
        else v = __ldcg(reinterpret_cast<const double2*>(&me->small[par][r][0]));
        sum = r == 0 ? v.x : sum + v.x;
        bad += v.y;
      }
      cost_buf[slot] = sum;
      if (bad > 0.0 && ctl->status == BA_OK) ctl->status = BA_ERR_SINGULAR;
    }
    me->epoch_small = epoch;
  }
}

// Closes one divided trailing update of the Cholesky (k4_cholesky.cu): everything this rank stored
// into the peers' matrices in the preceding kernel is published, and the peers' shares are awaited.
__global__ void comm_chol_sync_kernel(CommDev cd, ba_lm_state* ctl, int use_ctl) {
  if (use_ctl && ctl->done) return;
  CommHeader* me = cd.hdr[cd.rank];
  const unsigned long long epoch = me->epoch_chol + 1;
  const int t = threadIdx.x;
  __shared__ int s_ok;
  if (t == 0) s_ok = 1;
  __syncthreads();
  if (t < cd.world && t != cd.rank) {
    __threadfence_system();
    st_release_sys(&cd.hdr[t]->flag_chol[cd.rank], epoch);
    if (!spin_until(&me->flag_chol[t], epoch)) s_ok = 0;
  }
  __syncthreads();
  if (t == 0) {
    if (!s_ok) comm_fail(ctl);
    me->epoch_chol = epoch;
  }
}

// ---- host side --------------------------------------------------------------------------------
static inline int64_t round_up_i64(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// Windows outlive engines: allocating, exporting and mapping a window costs milliseconds (and a
// multi-gigabyte cudaFree synchronises the device), more than a whole small adjustment.  A rank
// that creates engine after engine of the same shape -- the usual way the reference's class is
// used -- therefore gets its previous window back, with the peers' mappings still open and the
// epochs simply continuing.  Every rank runs the same sequence of creations, so all ranks hit or
// miss together; the handshake blob carries the epochs and ba_comm_connect refuses windows that
// are out of step.
struct Window {
  int device = 0, rank = 0, world = 0, n_pad = 0;
  int64_t red_len = 0;
  void* base = nullptr;
  size_t bytes = 0;
  cudaIpcMemHandle_t handle{};
  cudaIpcMemHandle_t peer_handle[kMaxRanks]{};
  void* peer_base[kMaxRanks] = {};
  bool mapped = false, in_use = false;
};
static std::mutex g_win_mutex;
static std::vector<Window*> g_windows;

struct Comm {
  CommDev dev{};
  Window* win = nullptr;
  bool connected = false;
};

struct HandshakeBlob {
  cudaIpcMemHandle_t handle;
  unsigned long long epoch, epoch_small;
  unsigned long long fresh;  // 1: this window was allocated by this call (nobody maps it yet)
};
static_assert(sizeof(HandshakeBlob) == BA_COMM_HANDLE_BYTES, "handshake blob size");

int launch_comm_allreduce_red(ba_engine* e, bool conditional, cudaStream_t s) {
  Comm* c = e->comm;
  if (!c || !c->connected) { set_error("exchange window not connected"); return BA_ERR_STATE; }
  ProfScope ps(e, PG_COMM, s);
  const int use_ctl = conditional ? 1 : 0;
  const CommDev& cd = c->dev;
  comm_push_kernel<<<cd.n_seg, 256, 0, s>>>(cd, e->ctl, use_ctl);
  BA_LAUNCH_CHECK();
  const int seg_blocks = (cd.n_seg + kSegBlock - 1) / kSegBlock;
  const int owned = (seg_blocks + cd.world - 1) / cd.world * kSegBlock;
  comm_reduce_kernel<<<owned, 256, 0, s>>>(cd, e->ctl, use_ctl);
  BA_LAUNCH_CHECK();
  comm_wait_kernel<<<1, 32, 0, s>>>(cd, e->ctl, use_ctl);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

int launch_comm_allreduce_cost(ba_engine* e, int slot, bool conditional, cudaStream_t s) {
  Comm* c = e->comm;
  if (!c || !c->connected) { set_error("exchange window not connected"); return BA_ERR_STATE; }
  ProfScope ps(e, PG_COMM, s);
  comm_small_kernel<<<1, 32, 0, s>>>(c->dev, e->cost_buf, slot, e->ctl, conditional ? 1 : 0);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

void comm_chol_split(ba_engine* e, CholSplit* out) {
  *out = CholSplit();
  const Comm* c = e->comm;
  if (!c || !c->connected) return;
  out->rank = c->dev.rank;
  out->world = c->dev.world;
  for (int q = 0; q < c->dev.world; ++q) out->S_peer[q] = c->dev.red[q];
}

int launch_comm_chol_sync(ba_engine* e, bool conditional, cudaStream_t s) {
  Comm* c = e->comm;
  if (!c || !c->connected) { set_error("exchange window not connected"); return BA_ERR_STATE; }
  comm_chol_sync_kernel<<<1, 32, 0, s>>>(c->dev, e->ctl, conditional ? 1 : 0);
  BA_LAUNCH_CHECK();
  return BA_OK;
}

// The engine lets go of its window; the window (and the peers' mappings) stay for the next engine.
void comm_free(ba_engine* e) {
  Comm* c = e->comm;
  if (!c) return;
  cudaDeviceSynchronize();
  {
    std::lock_guard<std::mutex> lock(g_win_mutex);
    if (c->win) c->win->in_use = false;
  }
  delete c;
  e->comm = nullptr;
}

static void fill_rank(Comm* c, int p, void* base) {
  char* b = static_cast<char*>(base);
  const int64_t red_bytes = round_up_i64(c->dev.red_len * (int64_t)sizeof(double), 256);
  c->dev.hdr[p] = reinterpret_cast<CommHeader*>(b);
  c->dev.red[p] = reinterpret_cast<double*>(b + kHeaderBytes);
  c->dev.stage[p] = reinterpret_cast<double*>(b + kHeaderBytes + red_bytes);
}

static int comm_create(ba_engine* e, int rank, int world) {
  if (e->comm) { set_error("exchange window already created"); return BA_ERR_STATE; }
  if (world < 2 || world > kMaxRanks || rank < 0 || rank >= world) {
    set_error("peer exchange needs 2..%d ranks (got rank %d of %d)", kMaxRanks, rank, world);
    return BA_ERR_INVALID;
  }
  BA_CUDA(cudaSetDevice(e->device));
  if (const char* lim = std::getenv("BA_COMM_SPIN_LIMIT_S")) {
    const unsigned long long ns = (unsigned long long)(std::atof(lim) * 1e9);
    if (ns > 0) BA_CUDA(cudaMemcpyToSymbol(g_spin_limit_ns, &ns, sizeof(ns)));
  }
  Comm* c = new (std::nothrow) Comm();
  if (!c) { set_error("out of host memory"); return BA_ERR_CUDA; }
  c->dev.rank = rank;
  c->dev.world = world;
  c->dev.n_pad = e->n_pad;
  c->dev.red_len = e->red_len;
  const int64_t tail = e->red_len - (int64_t)e->n_pad * e->n_pad;
  c->dev.n_seg = e->n_pad + (int)((tail + e->n_pad - 1) / e->n_pad);
  const int64_t red_bytes = round_up_i64(e->red_len * (int64_t)sizeof(double), 256);
  Window* w = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_win_mutex);
    for (Window* cand : g_windows)
      if (!cand->in_use && cand->device == e->device && cand->rank == rank && cand->world == world &&
          cand->red_len == e->red_len && cand->n_pad == e->n_pad) {
        w = cand;
        w->in_use = true;
        break;
      }
  }
  if (!w) {
    w = new (std::nothrow) Window();
    if (!w) { delete c; set_error("out of host memory"); return BA_ERR_CUDA; }
    w->device = e->device; w->rank = rank; w->world = world; w->n_pad = e->n_pad; w->red_len = e->red_len;
    // stage slots keep the red layout (stage[src][idx]); only the owned lower-triangle segments
    // of each slot are ever touched
    w->bytes = kHeaderBytes + (size_t)red_bytes + (size_t)world * (size_t)e->red_len * sizeof(double);
    cudaError_t err = cudaMalloc(&w->base, w->bytes);
    if (err == cudaSuccess) err = cudaMemset(w->base, 0, kHeaderBytes);
    if (err == cudaSuccess) err = cudaIpcGetMemHandle(&w->handle, w->base);
    if (err != cudaSuccess) {
      set_error("exchange window of %zu bytes: %s", w->bytes, cudaGetErrorString(err));
      if (w->base) cudaFree(w->base);
      delete w;
      delete c;
      return BA_ERR_CUDA;
    }
    w->in_use = true;
    std::lock_guard<std::mutex> lock(g_win_mutex);
    g_windows.push_back(w);
  }
  c->win = w;
  fill_rank(c, rank, w->base);
  e->comm = c;
  return BA_OK;
}

}  // namespace ba

using namespace ba;

extern "C" {

int ba_comm_create(ba_engine* e, int rank, int world, void* handle_out) {
  if (!e || !handle_out) { set_error("null argument"); return BA_ERR_INVALID; }
  BA_TRY(comm_create(e, rank, world));
  Comm* c = e->comm;
  // the reduce buffer moves into the window: peers store the sums straight into it
  BA_CUDA(cudaMemcpy(c->dev.red[rank], e->red, (size_t)e->red_len * sizeof(double), cudaMemcpyDeviceToDevice));
  cudaFreeAsync(e->red, (cudaStream_t)0);
  e->red = c->dev.red[rank];
  e->red_in_window = true;
  for (int k = 0; k < 2; ++k)  // graphs captured before hold the old pointer
    if (e->solve_graph[k]) { cudaGraphExecDestroy(e->solve_graph[k]); e->solve_graph[k] = nullptr; }
  HandshakeBlob blob;
  blob.handle = c->win->handle;
  unsigned long long ep[2];
  BA_CUDA(cudaMemcpy(ep, &c->dev.hdr[rank]->epoch, sizeof(ep), cudaMemcpyDeviceToHost));  // synchronises
  blob.epoch = ep[0];
  blob.epoch_small = ep[1];
  blob.fresh = c->win->mapped ? 0ull : 1ull;
  std::memcpy(handle_out, &blob, sizeof(blob));
  return BA_OK;
}

int ba_comm_connect(ba_engine* e, const void* handles) {
  if (!e || !handles || !e->comm) { set_error("ba_comm_create must come first"); return BA_ERR_STATE; }
  Comm* c = e->comm;
  Window* w = c->win;
  BA_CUDA(cudaSetDevice(e->device));
  const HandshakeBlob* hb = static_cast<const HandshakeBlob*>(handles);
  const HandshakeBlob& mine = hb[c->dev.rank];
  for (int p = 0; p < c->dev.world; ++p)
    if (hb[p].epoch != mine.epoch || hb[p].epoch_small != mine.epoch_small) {
      set_error("exchange windows out of step: rank %d is at epoch %llu/%llu, rank %d at %llu/%llu "
                "(all ranks must create and run their engines in the same order)",
                p, hb[p].epoch, hb[p].epoch_small, c->dev.rank, mine.epoch, mine.epoch_small);
      return BA_ERR_STATE;
    }
  for (int p = 0; p < c->dev.world; ++p) {
    if (p == c->dev.rank) continue;
    const bool same = w->mapped && w->peer_base[p] &&
                      std::memcmp(&w->peer_handle[p], &hb[p].handle, sizeof(cudaIpcMemHandle_t)) == 0;
    if (!same) {
      if (w->peer_base[p]) { cudaIpcCloseMemHandle(w->peer_base[p]); w->peer_base[p] = nullptr; }
      void* base = nullptr;
      BA_CUDA(cudaIpcOpenMemHandle(&base, hb[p].handle, cudaIpcMemLazyEnablePeerAccess));
      w->peer_base[p] = base;
      w->peer_handle[p] = hb[p].handle;
    }
    fill_rank(c, p, w->peer_base[p]);
  }
  w->mapped = true;
  c->connected = true;
  return BA_OK;
}

int ba_comm_disconnect(ba_engine* e) {
  if (!e) { set_error("null engine"); return BA_ERR_INVALID; }
  if (!e->comm) return BA_OK;
  BA_CUDA(cudaSetDevice(e->device));
  if (e->own_stream) BA_CUDA(cudaStreamSynchronize(e->own_stream));
  BA_CUDA(cudaDeviceSynchronize());
  e->comm->connected = false;
  return BA_OK;
}

int ba_comm_world(ba_engine* e, int* rank, int* world) {
  if (!e) { set_error("null engine"); return BA_ERR_INVALID; }
  const bool on = e->comm && e->comm->connected;
  if (rank) *rank = on ? e->comm->dev.rank : 0;
  if (world) *world = on ? e->comm->dev.world : 1;
  return BA_OK;
}

}  // extern "C"
