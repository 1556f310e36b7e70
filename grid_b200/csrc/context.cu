// context.cu -- device context, streams, timers, NCCL communicator bootstrap.
// Replaces the reference's Grid_init/acceleratorInit (ref: Grid/threads/Accelerator.cc:19-110) and the
// CartesianCommunicator plumbing (ref: Grid/communicator/Communicator_mpi3.cc:226-307) with one process per
// GPU + NCCL.  NCCL is dlopen'ed lazily so that single-GPU use has no dependency on it and so that inside a
// python process the already loaded torch-bundled libnccl.so.2 is reused.
#include "internal.hpp"
#include "comm.hpp"
#include <dlfcn.h>
#include <cstring>
#include <mutex>

#include <set>
namespace gb {
std::set<const gb_context *> &live_contexts() { static std::set<const gb_context *> s; return s; }
bool context_alive(const gb_context *ctx) { return ctx && live_contexts().count(ctx) != 0; }
static thread_local std::string g_last_error;
void set_last_error(const std::string &m) { g_last_error = m; }

void check_launch(gb_context *ctx, const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw Error(GB_ERR_CUDA, std::string("kernel launch failed: ") + what + ": " + cudaGetErrorString(e));
}

// ------------------------------------------------------------------ NCCL via dlopen
NcclApi g_nccl;
static std::once_flag g_nccl_once;
static void load_nccl() {
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return;
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
  g_nccl.Send = (decltype(g_nccl.Send))dlsym(h, "ncclSend");
  g_nccl.Recv = (decltype(g_nccl.Recv))dlsym(h, "ncclRecv");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(h, "ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(h, "ncclGroupEnd");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.Send && g_nccl.Recv && g_nccl.GroupStart && g_nccl.GroupEnd;
}
NcclApi &nccl() {
  std::call_once(g_nccl_once, load_nccl);
  if (!g_nccl.ok) throw Error(GB_ERR_COMM, "libnccl.so.2 could not be loaded");
  return g_nccl;
}
void nccl_check(int r, const char *what) {
  if (r != 0) {
    const char *s = g_nccl.GetErrorString ? g_nccl.GetErrorString((ncclResult_t)r) : "?";
    throw Error(GB_ERR_COMM, std::string(what) + ": NCCL error " + std::to_string(r) + " " + s);
  }
}

void global_sum(gb_context *ctx, double *v, int n) {
  if (ctx->nranks == 1) return;
  GB_REQUIRE(n <= 8, "global_sum: at most 8 values");
  // host scalars -> device -> all-reduce -> host (ref: GlobalSum = MPI_Allreduce of host scalars)
  std::memcpy(ctx->h_result, v, n * sizeof(double));
  GB_CUDA(cudaMemcpyAsync(ctx->d_result, ctx->h_result, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  nccl_check(nccl().AllReduce(ctx->d_result, ctx->d_result, n, NCCL_DOUBLE, NCCL_SUM, ctx->nccl, ctx->stream), "ncclAllReduce");
  GB_CUDA(cudaMemcpyAsync(ctx->h_result, ctx->d_result, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  std::memcpy(v, ctx->h_result, n * sizeof(double));
}
// in-stream all-reduce of device-resident doubles (no host round trip); no-op on a single rank
void device_global_sum(gb_context *ctx, double *d_vals, int n) {
  if (ctx->nranks == 1) return;
  nccl_check(nccl().AllReduce(d_vals, d_vals, n, NCCL_DOUBLE, NCCL_SUM, ctx->nccl, ctx->stream), "ncclAllReduce");
}
} // namespace gb

namespace gb { void host_pipe_release(gb_context *ctx); }   // dhop_host.cu
using namespace gb;

__global__ void gb_l2_flush_kernel(float4 *p, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

extern "C" {

const char *gb_last_error(void) { return g_last_error.c_str(); }

int gb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int gb_context_create(int device, gb_context **out) {
  GB_API_BEGIN
  GB_REQUIRE(out != nullptr, "null output pointer");
  int n = gb_device_count();
  if (n <= 0) throw Error(GB_ERR_NO_DEVICE, "no CUDA device visible: libgridb200 has no CPU fallback");
  GB_REQUIRE(device >= 0 && device < n, "device index out of range");
  GB_CUDA(cudaSetDevice(device));
  gb_context *c = new gb_context();
  c->device = device;
  cudaDeviceProp prop;
  GB_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  GB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    // the halo stream outranks the compute stream so that exchange kernels are scheduled ahead of queued interior CTAs
    int lo = 0, hi = 0;
    GB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    GB_CUDA(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
  }
  GB_CUDA(cudaEventCreate(&c->ev_start));
  GB_CUDA(cudaEventCreate(&c->ev_stop));
  GB_CUDA(cudaEventCreateWithFlags(&c->ev_comm, cudaEventDisableTiming));
  GB_CUDA(cudaEventCreateWithFlags(&c->ev_comp, cudaEventDisableTiming));
  c->max_partials = 4096;
  GB_CUDA(cudaMalloc(&c->d_partials, sizeof(double) * 4 * c->max_partials));
  GB_CUDA(cudaMalloc(&c->d_result, sizeof(double) * 8));
  GB_CUDA(cudaMallocHost(&c->h_result, sizeof(double) * 8));
  GB_CUDA(cudaMalloc(&c->d_scalars, sizeof(double) * 8));
  GB_CUDA(cudaEventCreateWithFlags(&c->ev_scalar, cudaEventDisableTiming));
  gb::live_contexts().insert(c);
  *out = c;
  GB_API_END
}

int gb_context_destroy(gb_context *c) {
  GB_API_BEGIN
  if (!c) return GB_OK;
  gb::live_contexts().erase(c);
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  gb::host_pipe_release(c);
  if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
  cudaFree(c->d_partials); cudaFree(c->d_result); cudaFreeHost(c->h_result); cudaFree(c->d_scalars); cudaEventDestroy(c->ev_scalar);
  if (c->l2_scratch) cudaFree(c->l2_scratch);
  if (c->staging) cudaFree(c->staging);
  for (auto &b : c->field_pool) cudaFree(b.second);
  cudaEventDestroy(c->ev_start); cudaEventDestroy(c->ev_stop); cudaEventDestroy(c->ev_comm); cudaEventDestroy(c->ev_comp);
  cudaStreamDestroy(c->stream); cudaStreamDestroy(c->comm_stream);
  delete c;
  GB_API_END
}

int gb_synchronize(gb_context *ctx) {
  GB_API_BEGIN
  GB_CUDA(cudaStreamSynchronize(ctx->comm_stream));
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  GB_API_END
}

int gb_timer_start(gb_context *ctx) {
  GB_API_BEGIN
  GB_CUDA(cudaEventRecord(ctx->ev_start, ctx->stream));
  GB_API_END
}
int gb_timer_stop(gb_context *ctx, double *elapsed_ms) {
  GB_API_BEGIN
  GB_CUDA(cudaEventRecord(ctx->ev_stop, ctx->stream));
  GB_CUDA(cudaEventSynchronize(ctx->ev_stop));
  float ms = 0;
  GB_CUDA(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
  *elapsed_ms = ms;
  GB_API_END
}
int64_t gb_launch_count(gb_context *ctx) { return ctx ? ctx->launches : 0; }

int gb_flush_l2(gb_context *ctx) {
  GB_API_BEGIN
  if (!ctx->l2_scratch) {
    ctx->l2_scratch_bytes = (size_t)256 << 20; // 2x the 126 MB L2
    GB_CUDA(cudaMalloc(&ctx->l2_scratch, ctx->l2_scratch_bytes));
  }
  gb_l2_flush_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((float4 *)ctx->l2_scratch, ctx->l2_scratch_bytes / 16);
  check_launch(ctx, "l2_flush");
  GB_API_END
}

// ------------------------------------------------------------------ communicator
int gb_comm_unique_id(void *id_out) {
  GB_API_BEGIN
  static_assert(sizeof(ncclUniqueId) == GB_UNIQUE_ID_BYTES, "unique id size");
  ncclUniqueId id;
  nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(id_out, &id, sizeof(id));
  GB_API_END
}
int gb_comm_init(gb_context *ctx, int rank, int nranks, const void *id_bytes) {
  GB_API_BEGIN
  GB_REQUIRE(ctx->nccl == nullptr, "communicator already initialised");
  GB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank");
  if (nranks > 1) {
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, sizeof(id));
    GB_CUDA(cudaSetDevice(ctx->device));
    nccl_check(nccl().CommInitRank(&ctx->nccl, nranks, id, rank), "ncclCommInitRank");
  }
  ctx->rank = rank;
  ctx->nranks = nranks;
  GB_API_END
}
int gb_comm_rank(gb_context *ctx, int *rank, int *nranks) {
  if (rank) *rank = ctx->rank;
  if (nranks) *nranks = ctx->nranks;
  return GB_OK;
}
int gb_comm_global_sum(gb_context *ctx, double *vals, int n) {
  GB_API_BEGIN
  global_sum(ctx, vals, n);
  GB_API_END
}
int gb_comm_barrier(gb_context *ctx) {
  GB_API_BEGIN
  double v = 0;
  GB_CUDA(cudaStreamSynchronize(ctx->comm_stream));
  GB_CUDA(cudaStreamSynchronize(ctx->stream));
  global_sum(ctx, &v, 1);
  GB_API_END
}
}
