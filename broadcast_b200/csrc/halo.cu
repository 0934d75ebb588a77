// Halo exchange of the i-slabs by PEER STORES over NVLink (the exchange the reference would do with
// MPI between blocks, cylinder.py:499-527 for the periodic cut), one process per GPU, no NCCL call and no host work on the data path.
//
// Every rank owns a MAILBOX in its own HBM (cudaMalloc, exported with cudaIpcGetMemHandle, opened by its two neighbours):
//     recv[parity 0/1][side 0/1][rows][gh]   packed halo columns, side = which of MY halos the data fills (0 left, 1 right)
//     arrived[side]                          step number of the newest complete delivery into recv[step & 1][side]
//     step, cta_done[2], error               local bookkeeping
// One exchange = two launches on the caller's stream (both capturable in a CUDA graph, the step number lives on the device):
//   k_halo_push    reads the gh first / last owned columns of w and stores them straight into the neighbours' mailboxes
//                  (st.global over NVLink), __threadfence_system, last CTA releases arrived[] = step in the neighbours' memory;
//   k_halo_unpack  spins on its OWN arrived[] flags (local HBM, acquire at system scope), copies recv[step & 1] into the halo
//                  columns of w, last CTA publishes step.
// Two recv buffers (step parity) make the protocol safe without a "consumed" flag: a neighbour's push of step s+2 into the
// buffer of step s follows its unpack of step s+1, which waited for my push s+1, which follows my unpack s in stream order.
// A spin that sees no delivery for ~2 s sets `error` and gives up (a dead neighbour must not hang the GPU).
#include <cstdint>
#include <cstring>
#include <new>
#include "../../include/broadcast_b200.h"
#include "kernels.cuh"

namespace bcast {
void count_launches(int n);
}
using namespace bcast;

namespace {

struct MailHdr {
  unsigned long long arrived[2];
  unsigned long long step;
  unsigned int cta_done[2];
  unsigned int error;
  unsigned int pad[9];
};
static_assert(sizeof(MailHdr) == 72 || sizeof(MailHdr) % 8 == 0, "header");
constexpr size_t HDR_BYTES = 128;

struct Halo {
  int device = 0;
  long long side_doubles = 0;     // rows * gh
  unsigned char* mail = nullptr;  // my mailbox
  unsigned char* peer[2] = {nullptr, nullptr};   // neighbours' mailboxes (IPC mappings): 0 = left neighbour, 1 = right neighbour
  bool ipc[2] = {false, false};
  int launches = 0;
};

__host__ __device__ inline MailHdr* hdr(unsigned char* m) { return reinterpret_cast<MailHdr*>(m); }
__host__ __device__ inline double* recv_buf(unsigned char* m, long long side_doubles, int parity, int side) {
  return reinterpret_cast<double*>(m + HDR_BYTES) + ((long long)(parity * 2 + side)) * side_doubles;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// w: (planes * nj) rows of ni doubles.  Columns [gh, 2gh) go to the left neighbour's RIGHT halo (its side 1), columns
// [im, im + gh) to the right neighbour's LEFT halo (its side 0).
__global__ void __launch_bounds__(256) k_halo_push(const double* __restrict__ w, long long rows, int ni, int gh, int im, unsigned char* mine,
                                                   unsigned char* left, unsigned char* right, long long side_doubles) {
  MailHdr* h = hdr(mine);
  const unsigned long long s = *reinterpret_cast<volatile unsigned long long*>(&h->step) + 1;
  const int par = (int)(s & 1);
  const long long n = rows * gh;
  double* dl = left ? recv_buf(left, side_doubles, par, 1) : nullptr;
  double* dr = right ? recv_buf(right, side_doubles, par, 0) : nullptr;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / gh;
    const int k = (int)(t - r * gh);
    const double* row = w + r * ni;
    if (dl) dl[t] = row[gh + k];
    if (dr) dr[t] = row[im + k];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(&h->cta_done[0], 1u);
    if (done == gridDim.x - 1) {
      h->cta_done[0] = 0;
      __threadfence_system();
      if (left) st_release_sys(&hdr(left)->arrived[1], s);
      if (right) st_release_sys(&hdr(right)->arrived[0], s);
    }
  }
}

__global__ void __launch_bounds__(256) k_halo_unpack(double* __restrict__ w, long long rows, int ni, int gh, int im, unsigned char* mine,
                                                     int has_left, int has_right, long long side_doubles, long long spin_limit) {
  MailHdr* h = hdr(mine);
  const unsigned long long s = *reinterpret_cast<volatile unsigned long long*>(&h->step) + 1;
  const int par = (int)(s & 1);
  __shared__ int ok;
  if (threadIdx.x == 0) {
    int good = 1;
    const long long t0 = clock64();
    for (int side = 0; side < 2; ++side) {
      if (!(side == 0 ? has_left : has_right)) continue;
      while (ld_acquire_sys(&h->arrived[side]) < s) {
        if (clock64() - t0 > spin_limit) { good = 0; break; }
        __nanosleep(64);
      }
    }
    if (!good) atomicExch(&h->error, 1u);
    ok = good;
  }
  __syncthreads();
  if (ok) {
    const long long n = rows * gh;
    const double* sl = recv_buf(mine, side_doubles, par, 0);
    const double* sr = recv_buf(mine, side_doubles, par, 1);
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
      const long long r = t / gh;
      const int k = (int)(t - r * gh);
      double* row = w + r * ni;
      if (has_left) row[k] = __ldcv(sl + t);
      if (has_right) row[im + gh + k] = __ldcv(sr + t);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(&h->cta_done[1], 1u);
    if (done == gridDim.x - 1) {
      h->cta_done[1] = 0;
      *reinterpret_cast<volatile unsigned long long*>(&h->step) = s;
    }
  }
}

}  // namespace

extern "C" {

int bcd_halo_create(void** handle, int gh, long long rows, unsigned char* ipc_handle_out64) {
  if (!handle || gh < 1 || rows < 1) return BC_ERR_ARG;
  Halo* H = new (std::nothrow) Halo;
  if (!H) return BC_ERR_ALLOC;
  cudaGetDevice(&H->device);
  H->side_doubles = rows * gh;
  const size_t bytes = HDR_BYTES + sizeof(double) * 4 * (size_t)H->side_doubles;
  if (cudaMalloc(&H->mail, bytes) != cudaSuccess) { delete H; return BC_ERR_ALLOC; }
  cudaMemset(H->mail, 0, bytes);
  cudaDeviceSynchronize();
  if (ipc_handle_out64) {
    cudaIpcMemHandle_t ih;
    cudaError_t e = cudaIpcGetMemHandle(&ih, H->mail);
    if (e != cudaSuccess) { cudaFree(H->mail); delete H; return (int)e; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    memcpy(ipc_handle_out64, &ih, 64);
  }
  *handle = H;
  return BC_OK;
}

// side: 0 = my left neighbour, 1 = my right neighbour.  `ipc_handle64` = the neighbour's handle from bcd_halo_create (another
// process); `same_process_mailbox` != NULL instead connects two halos of ONE process (several devices driven by one host thread,
// tests): the pointer returned by bcd_halo_mailbox of the neighbour, peer access enabled here.
int bcd_halo_connect(void* handle, int side, const unsigned char* ipc_handle64, void* same_process_mailbox, int peer_device) {
  Halo* H = static_cast<Halo*>(handle);
  if (!H || side < 0 || side > 1) return BC_ERR_ARG;
  if (same_process_mailbox) {
    if (peer_device != H->device) {
      cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return (int)e;
      cudaGetLastError();
    }
    H->peer[side] = static_cast<unsigned char*>(same_process_mailbox);
    H->ipc[side] = false;
    return BC_OK;
  }
  if (!ipc_handle64) return BC_ERR_ARG;
  cudaIpcMemHandle_t ih;
  memcpy(&ih, ipc_handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return (int)e;
  H->peer[side] = static_cast<unsigned char*>(p);
  H->ipc[side] = true;
  return BC_OK;
}

void* bcd_halo_mailbox(void* handle) { return handle ? static_cast<Halo*>(handle)->mail : nullptr; }
// the mapping of the neighbour's mailbox on `side` (NULL when not connected): a neighbour that sits on BOTH sides (two slabs of a
// periodic block) is opened once and connected to the second side through this pointer (same_process_mailbox)
void* bcd_halo_peer(void* handle, int side) { return handle && side >= 0 && side <= 1 ? static_cast<Halo*>(handle)->peer[side] : nullptr; }

// one exchange of the gh halo columns of w ((planes * nj) rows of ni = im + 2 gh doubles) with the connected neighbours
int bcd_halo_exchange(void* handle, double* w, long long rows, int ni, int gh, void* stream) {
  Halo* H = static_cast<Halo*>(handle);
  if (!H || !w || rows * gh != H->side_doubles || ni < 3 * gh) return BC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int im = ni - 2 * gh;
  const long long n = rows * gh;
  int ctas = (int)((n + 255) / 256);
  if (ctas > 64) ctas = 64;
  if (!H->peer[0] && !H->peer[1]) return BC_OK;
  k_halo_push<<<ctas, 256, 0, st>>>(w, rows, ni, gh, im, H->mail, H->peer[0], H->peer[1], H->side_doubles);
  k_halo_unpack<<<ctas, 256, 0, st>>>(w, rows, ni, gh, im, H->mail, H->peer[0] != nullptr, H->peer[1] != nullptr, H->side_doubles,
                                      4000000000LL);
  count_launches(2);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? BC_OK : (int)e;
}

// 0 = fine; 1 = an unpack gave up waiting for a neighbour (results of that step are invalid)
int bcd_halo_error(void* handle) {
  Halo* H = static_cast<Halo*>(handle);
  if (!H) return BC_ERR_ARG;
  MailHdr h;
  if (cudaMemcpy(&h, H->mail, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int)h.error;
}

int bcd_halo_destroy(void* handle) {
  Halo* H = static_cast<Halo*>(handle);
  if (!H) return BC_OK;
  cudaDeviceSynchronize();
  for (int s = 0; s < 2; ++s)
    if (H->peer[s] && H->ipc[s]) cudaIpcCloseMemHandle(H->peer[s]);
  cudaFree(H->mail);
  delete H;
  return BC_OK;
}

// ---- CUDA-graph capture of a sequence of bcd_* calls on `stream` (the per-step sequence exchange + fills + residual: launch
// latency, not device time, bounds the step at 8 GPUs) -------------------------------------------------------------------------
struct GraphExec {
  cudaGraphExec_t exec;
  long long launches;
};
static thread_local long long g_cap_launches0 = 0;
long long bc_launch_count(void);

int bcd_graph_begin(void* stream) {
  g_cap_launches0 = bc_launch_count();
  cudaError_t e = cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeThreadLocal);
  return e == cudaSuccess ? BC_OK : (int)e;
}
int bcd_graph_end(void* stream, void** exec_out) {
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture((cudaStream_t)stream, &g);
  if (e != cudaSuccess) return (int)e;
  GraphExec* G = new (std::nothrow) GraphExec{nullptr, bc_launch_count() - g_cap_launches0};
  if (!G) { cudaGraphDestroy(g); return BC_ERR_ALLOC; }
  e = cudaGraphInstantiate(&G->exec, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) { delete G; return (int)e; }
  *exec_out = G;
  return BC_OK;
}
int bcd_graph_launch(void* exec, void* stream) {
  GraphExec* G = static_cast<GraphExec*>(exec);
  if (!G) return BC_ERR_ARG;
  cudaError_t e = cudaGraphLaunch(G->exec, (cudaStream_t)stream);
  count_launches((int)G->launches);
  return e == cudaSuccess ? BC_OK : (int)e;
}
int bcd_graph_destroy(void* exec) {
  GraphExec* G = static_cast<GraphExec*>(exec);
  if (G) {
    cudaGraphExecDestroy(G->exec);
    delete G;
  }
  return BC_OK;
}

}  // extern "C"
