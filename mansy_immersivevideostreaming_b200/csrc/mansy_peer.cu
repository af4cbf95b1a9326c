// mansy_peer.cu -- the one exchange of a rollout as OUR kernel over NVLink peer memory.
//
// Environments never interact (bitrate_selection/envs/mansy_env.py holds no cross-env state), so the only
// inter-GPU traffic of a rollout is the all-gather of fixed-size per-env episode statistics ([N_local][6] float64:
// sum qoe, qoe1, qoe2, qoe3, steps, episodes -- what `_log`, envs/mansy_env.py:271-290, averages per episode).
// At 4 096 envs per GPU that is 196 KB per rank: a collective library call costs more in launch latency than the
// bytes cost on the wire.  Here every rank owns a "mailbox" allocation that its peers map through CUDA IPC
// (one process per GPU); ONE kernel packs the six totals columns out of the simulator's statistics rows and
// stores them straight into every peer's mailbox (16-byte stores over NVLink / NVSwitch), publishes an epoch
// flag per destination with release semantics at system scope, and its last block waits until the flags of all
// peers have arrived -- so when the kernel completes, the gathered [world * N_local][6] array is in local HBM and
// ordinary stream order makes it visible to whatever consumes it.  The same flags give a device-side barrier
// (`mansy_peer_barrier`) that aligns the GPUs of a job to within a few microseconds without a host round trip.
//
// Mailbox layout (per rank):  [0, 4096) flags:  uint64 barrier[16] | uint64 gather[2][16] | uint32 done_blocks
//                             [4096, ...) data: [2 parities][world][slot_bytes]
// Two data parities: a rank can only push epoch e + 2 after every peer has pushed e + 1, which each peer does
// after (in stream order) it consumed epoch e.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include <cuda_runtime.h>

#include "mansy_sim.cuh"

namespace mansy {
int set_error(int code, const std::string &msg);
void count_launch();
const SimDev *sim_dev_of(mansy_handle_t h);     // mansy_sim.cu
int sim_device_of(mansy_handle_t h);            // mansy_sim.cu
}  // namespace mansy

using namespace mansy;

#define MANSY_CUDA(expr)                                                                         \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return mansy::set_error(MANSY_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

namespace {

constexpr int kMaxPeers = 16;
constexpr size_t kHeaderBytes = 4096;
constexpr int kTotCols = 6;      // MANSY_STAT_TOT_SUM_QOE .. MANSY_STAT_TOT_EPISODES

struct PeerDev {
  uint8_t *box[kMaxPeers];       // mapped mailbox of every rank (box[rank] = the local one)
  int32_t world, rank;
  uint64_t slot_bytes;
  int32_t debug;                 // MANSY_PEER_TIMELINE=1: the gather kernel prints its own phase times (globaltimer, ns)
};

__device__ __forceinline__ uint64_t *flag_barrier(uint8_t *box, int src) { return reinterpret_cast<uint64_t *>(box) + src; }
__device__ __forceinline__ uint64_t *flag_gather(uint8_t *box, int parity, int src) {
  return reinterpret_cast<uint64_t *>(box) + 16 + parity * 16 + src;
}
__device__ __forceinline__ uint32_t *done_blocks(uint8_t *box) { return reinterpret_cast<uint32_t *>(box + 16 * 8 * 3); }

__device__ __forceinline__ void st_release_sys(uint64_t *p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ uint32_t *timeout_flag(uint8_t *box) { return reinterpret_cast<uint32_t *>(box + 16 * 8 * 3 + 8); }
// Wait until *f >= epoch; gives up after ~4 s (a peer that died must not hang this GPU) and records it.
__device__ __forceinline__ void wait_flag(const uint64_t *f, uint64_t epoch, uint8_t *mine) {
  uint64_t t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (ld_acquire_sys(f) < epoch) {
    __nanosleep(64);
    uint64_t t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 4000000000ULL) { atomicExch(timeout_flag(mine), 1u); break; }
  }
}

// Device-side barrier over the job's GPUs: lane t publishes `epoch` in rank t's mailbox and waits for rank t's flag here.
__global__ void peer_barrier_kernel(const PeerDev P, uint64_t epoch) {
  const int t = threadIdx.x;
  if (t >= P.world) return;
  st_release_sys(flag_barrier(P.box[t], P.rank), epoch);
  wait_flag(flag_barrier(P.box[P.rank], t), epoch, P.box[P.rank]);
}

// Pack the six totals columns of stats[n][16] into [n][6] rows (plain single-GPU form, no peers).
__global__ void stats_totals_kernel(const double *__restrict__ stats, int n, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per (env, column pair)
  if (i >= n * 3) return;
  const int e = i / 3, c = (i % 3) * 2;
  const double2 v = *reinterpret_cast<const double2 *>(stats + (size_t)e * MANSY_STATS_DOUBLES + MANSY_STAT_TOT_SUM_QOE + c);
  *reinterpret_cast<double2 *>(out + (size_t)e * kTotCols + c) = v;
}

// All-gather of the totals: pack + push to every mailbox + flags + wait (see the file header).
__global__ void __launch_bounds__(256) peer_allgather_stats_kernel(const PeerDev P, const double *__restrict__ stats, int n,
                                                                  uint64_t epoch) {
  // (launched with programmatic stream serialisation: scheduled while the rollout kernel before it drains, ordered here)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  uint64_t ts0 = 0, ts1 = 0, ts2 = 0, ts3 = 0, ts4 = 0;
  if (P.debug) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts0));
  const int parity = (int)(epoch & 1);
  const size_t region = (size_t)P.world * P.slot_bytes;
  const size_t my_off = kHeaderBytes + (size_t)parity * region + (size_t)P.rank * P.slot_bytes;
  const int items = n * 3;                                   // double2 items
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < items; i += gridDim.x * blockDim.x) {
    const int e = i / 3, c = (i % 3) * 2;
    const double2 v = *reinterpret_cast<const double2 *>(stats + (size_t)e * MANSY_STATS_DOUBLES + MANSY_STAT_TOT_SUM_QOE + c);
    for (int d = 0; d < P.world; ++d) {
      const int dst = (P.rank + d) % P.world;                // own copy first, then the peers in ring order
      *reinterpret_cast<double2 *>(P.box[dst] + my_off + ((size_t)e * kTotCols + c) * sizeof(double)) = v;
    }
  }
  if (P.world == 1) return;                                  // one GPU: a local pack, ordered by the stream
  // Release pattern with ONE system-scope fence per block (a fence in each of the 256 threads was most of the first
  // version's 13 us; each one waits for the acknowledgements of the remote stores, ~3 us over NVLink): the block's stores
  // happen before the barrier, thread 0's fence orders them -- cumulatively -- before its counter increment.  The last block
  // to increment has every block's stores performed at their destinations; its thread 0 fences once more (the acquire side
  // of the counter, the release side of the flags) and posts the flags itself as relaxed system-scope stores -- fence and
  // stores in one thread.  (The first cut ran fence, fence, fence + st.release in a row there: 10 of the kernel's 12 us.)
  __syncthreads();
  if (P.debug) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts1));
  __shared__ bool last;
  uint8_t *mine = P.box[P.rank];
  if (threadIdx.x == 0) {
    fence_acq_rel_sys();
    last = atomicAdd(done_blocks(mine), 1u) == gridDim.x - 1;
    if (last) {
      if (P.debug) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts2));
      fence_acq_rel_sys();
      for (int d = 0; d < P.world; ++d)
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(flag_gather(P.box[d], parity, P.rank)), "l"(epoch) : "memory");
      if (P.debug) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts3));
      *done_blocks(mine) = 0u;
    }
  }
  __syncthreads();
  if (!last) return;
  const int t = threadIdx.x;
  if (t < P.world) {
    wait_flag(flag_gather(mine, parity, t), epoch, mine);
    if (P.debug) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts4));
  }
  if (P.debug && t == 0)
    printf("gather rank %d epoch %llu: start->pushed %llu ns, ->last block %llu, ->flags posted %llu, ->peer 0's flag seen %llu\n", P.rank,
           (unsigned long long)epoch, (unsigned long long)(ts1 - ts0), (unsigned long long)(ts2 - ts1), (unsigned long long)(ts3 - ts2),
           (unsigned long long)(ts4 - ts3));
}

}  // namespace

struct mansy_peer {
  PeerDev dev;
  int device = 0;
  bool connected = false;
  void *opened[kMaxPeers] = {nullptr};
  uint64_t barrier_epoch = 0, gather_epoch = 0;
};

extern "C" {

int mansy_peer_create(int32_t world, int32_t rank, int64_t slot_bytes, int device, mansy_peer_t *out) {
  if (!out) return set_error(MANSY_E_INVALID, "out is NULL");
  *out = nullptr;
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return set_error(MANSY_E_INVALID, "bad world / rank (world <= 16)");
  if (slot_bytes < 16 || (slot_bytes & 15)) return set_error(MANSY_E_INVALID, "slot_bytes must be a positive multiple of 16");
  DeviceScope dscope(device);
  MANSY_CUDA(dscope.err);
  mansy_peer *p = new (std::nothrow) mansy_peer();
  if (!p) return set_error(MANSY_E_NOMEM, "out of host memory");
  p->device = device;
  memset(&p->dev, 0, sizeof(p->dev));
  p->dev.world = world; p->dev.rank = rank; p->dev.slot_bytes = (uint64_t)slot_bytes;
  p->dev.debug = (getenv("MANSY_PEER_TIMELINE") && getenv("MANSY_PEER_TIMELINE")[0] == '1') ? 1 : 0;
  const size_t bytes = kHeaderBytes + 2 * (size_t)world * (size_t)slot_bytes;
  void *d = nullptr;
  if (cudaMalloc(&d, bytes) != cudaSuccess) { delete p; return set_error(MANSY_E_NOMEM, "cudaMalloc failed (peer mailbox)"); }
  if (cudaMemset(d, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    cudaFree(d); delete p;
    return set_error(MANSY_E_CUDA, "cudaMemset failed (peer mailbox)");
  }
  p->dev.box[rank] = static_cast<uint8_t *>(d);
  p->connected = world == 1;
  *out = p;
  return MANSY_OK;
}

int mansy_peer_export(mansy_peer_t p, void *handle_out) {
  if (!p || !handle_out) return set_error(MANSY_E_INVALID, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == MANSY_PEER_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  DeviceScope dscope(p->device);
  MANSY_CUDA(dscope.err);
  MANSY_CUDA(cudaIpcGetMemHandle(&h, p->dev.box[p->dev.rank]));
  memcpy(handle_out, &h, sizeof(h));
  return MANSY_OK;
}

int mansy_peer_connect(mansy_peer_t p, const void *all_handles) {
  if (!p || !all_handles) return set_error(MANSY_E_INVALID, "NULL argument");
  if (p->connected) return MANSY_OK;
  DeviceScope dscope(p->device);
  MANSY_CUDA(dscope.err);
  for (int r = 0; r < p->dev.world; ++r) {
    if (r == p->dev.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const uint8_t *>(all_handles) + (size_t)r * sizeof(h), sizeof(h));
    void *ptr = nullptr;
    MANSY_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->opened[r] = ptr;
    p->dev.box[r] = static_cast<uint8_t *>(ptr);
  }
  p->connected = true;
  return MANSY_OK;
}

int mansy_peer_destroy(mansy_peer_t p) {
  if (!p) return MANSY_OK;
  DeviceScope dscope(p->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < kMaxPeers; ++r)
    if (p->opened[r]) cudaIpcCloseMemHandle(p->opened[r]);
  if (p->dev.box[p->dev.rank]) cudaFree(p->dev.box[p->dev.rank]);
  delete p;
  return MANSY_OK;
}

int mansy_peer_barrier(mansy_peer_t p, void *stream) {
  if (!p) return set_error(MANSY_E_INVALID, "NULL argument");
  if (!p->connected) return set_error(MANSY_E_STATE, "peer group is not connected");
  if (p->dev.world == 1) return MANSY_OK;
  DeviceScope dscope(p->device);
  MANSY_CUDA(dscope.err);
  peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(p->dev, ++p->barrier_epoch);
  count_launch();
  MANSY_CUDA(cudaGetLastError());
  return MANSY_OK;
}

int mansy_peer_allgather_stats(mansy_peer_t p, mansy_handle_t h, void *stream, const double **gathered_dev) {
  if (!p || !h) return set_error(MANSY_E_INVALID, "NULL argument");
  if (!p->connected) return set_error(MANSY_E_STATE, "peer group is not connected");
  DeviceScope dscope(p->device);
  MANSY_CUDA(dscope.err);
  const SimDev *S = sim_dev_of(h);
  const size_t need = (size_t)S->n_envs * kTotCols * sizeof(double);
  if (need > p->dev.slot_bytes) return set_error(MANSY_E_INVALID, "mailbox slot smaller than n_envs x 6 doubles");
  const uint64_t epoch = ++p->gather_epoch;
  const int items = S->n_envs * 3;
  int grid = (items + 255) / 256;
  if (grid > 64) grid = 64;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;    // its launch overlaps the tail of the rollout kernel before it
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.stream = static_cast<cudaStream_t>(stream);
  cfg.attrs = attr;
  static const bool pdl = !(getenv("MANSY_PEER_GATHER_PDL") && getenv("MANSY_PEER_GATHER_PDL")[0] == '0');
  cfg.numAttrs = pdl ? 1 : 0;
  MANSY_CUDA(cudaLaunchKernelEx(&cfg, peer_allgather_stats_kernel, p->dev, static_cast<const double *>(S->stats), (int)S->n_envs, epoch));
  count_launch();
  MANSY_CUDA(cudaGetLastError());
  if (gathered_dev)
    *gathered_dev = reinterpret_cast<const double *>(p->dev.box[p->dev.rank] + kHeaderBytes +
                                                     (size_t)(epoch & 1) * p->dev.world * p->dev.slot_bytes);
  return MANSY_OK;
}

int mansy_peer_timed_out(mansy_peer_t p, int32_t *flag_host) {
  if (!p || !flag_host) return set_error(MANSY_E_INVALID, "NULL argument");
  uint32_t v = 0;
  MANSY_CUDA(cudaMemcpy(&v, p->dev.box[p->dev.rank] + 16 * 8 * 3 + 8, sizeof(v), cudaMemcpyDeviceToHost));
  *flag_host = (int32_t)v;
  return MANSY_OK;
}

int mansy_episode_totals(mansy_handle_t h, double *totals_dev, void *stream) {
  if (!h || !totals_dev) return set_error(MANSY_E_INVALID, "NULL argument");
  if (reinterpret_cast<uintptr_t>(totals_dev) & 15) return set_error(MANSY_E_INVALID, "totals must be 16-byte aligned");
  DeviceScope dscope(sim_device_of(h));
  MANSY_CUDA(dscope.err);
  const SimDev *S = sim_dev_of(h);
  const int items = S->n_envs * 3;
  stats_totals_kernel<<<(items + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(S->stats, S->n_envs, totals_dev);
  count_launch();
  MANSY_CUDA(cudaGetLastError());
  return MANSY_OK;
}

}  // extern "C"
