/*
 * plg_comm.cu - the cross-PROCESS half of site sharding: one rank per GPU, every rank owns a
 * contiguous pattern slice of every CLV, and the only values that couple the ranks are the sums
 * of reference src/core_likelihood_avx.c:1259 (`logl +=`) and src/core_derivatives_avx2.c:756-765
 * (d_f, dd_f).  They are combined by a 1- or 2-double ncclAllReduce on the partition's own stream,
 * between the reduction kernel and the hand-over of the result to the host - so a plain C caller
 * of pll_compute_edge_loglikelihood / _root_loglikelihood / pll_compute_likelihood_derivatives
 * running under mpirun / torchrun gets the global value from the unchanged call.
 *
 * NCCL is loaded at run time (dlopen): the library has no link-time dependency on it and a
 * process that never calls pll_gpu_comm_init never touches it.  The caller carries the 128-byte
 * unique id from rank 0 to the other ranks with whatever it has (MPI_Bcast, torch.distributed,
 * a file): pll_gpu_comm_unique_id on rank 0, pll_gpu_comm_init(id, nranks, rank) everywhere.
 */
#include <dlfcn.h>

#include "plg_internal.cuh"

typedef struct { char internal[128]; } nccl_unique_id;
typedef void * nccl_comm;

static struct
{
  void * handle;
  int (*get_unique_id)(nccl_unique_id *);
  int (*comm_init_rank)(nccl_comm *, int, nccl_unique_id, int);
  int (*all_reduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t);
  int (*comm_destroy)(nccl_comm);
  const char * (*get_error_string)(int);
  nccl_comm comm;
  int nranks, rank, device;
} g_nccl;

static int load_nccl()
{
  if (g_nccl.handle) return PLG_OK;
  const char * names[] = {getenv("PLL_GPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so", NULL};
  for (int i = 0; i < 3 && !g_nccl.handle; ++i)
    if (names[i] && *names[i]) g_nccl.handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl.handle)
  {
    plg_set_error("pll_gpu_comm: NCCL not found (libnccl.so.2; set PLL_GPU_NCCL_LIB): %s", dlerror());
    return PLG_E_UNSUPPORTED;
  }
  *(void **)&g_nccl.get_unique_id = dlsym(g_nccl.handle, "ncclGetUniqueId");
  *(void **)&g_nccl.comm_init_rank = dlsym(g_nccl.handle, "ncclCommInitRank");
  *(void **)&g_nccl.all_reduce = dlsym(g_nccl.handle, "ncclAllReduce");
  *(void **)&g_nccl.comm_destroy = dlsym(g_nccl.handle, "ncclCommDestroy");
  *(void **)&g_nccl.get_error_string = dlsym(g_nccl.handle, "ncclGetErrorString");
  if (!g_nccl.get_unique_id || !g_nccl.comm_init_rank || !g_nccl.all_reduce || !g_nccl.comm_destroy)
  {
    plg_set_error("pll_gpu_comm: the NCCL library lacks an entry point");
    dlclose(g_nccl.handle);
    g_nccl.handle = NULL;
    return PLG_E_UNSUPPORTED;
  }
  return PLG_OK;
}

static int nccl_fail(const char * what, int rc)
{
  plg_set_error("%s failed: %s", what, g_nccl.get_error_string ? g_nccl.get_error_string(rc) : "NCCL error");
  return PLG_E_CUDA;
}

extern "C" int plg_comm_unique_id(unsigned char * id)
{
  int rc = load_nccl();
  if (rc) return rc;
  nccl_unique_id u;
  int n = g_nccl.get_unique_id(&u);
  if (n) return nccl_fail("ncclGetUniqueId", n);
  memcpy(id, u.internal, sizeof(u.internal));
  return PLG_OK;
}

extern "C" int plg_comm_init(const unsigned char * id, int nranks, int rank, int device)
{
  if (g_nccl.comm)
  {
    plg_set_error("pll_gpu_comm_init: a communicator exists already");
    return PLG_E_INVALID;
  }
  if (!id || nranks < 1 || rank < 0 || rank >= nranks)
  {
    plg_set_error("pll_gpu_comm_init: bad arguments");
    return PLG_E_INVALID;
  }
  int rc = load_nccl();
  if (rc) return rc;
  if (device < 0) PLG_CUDA(cudaGetDevice(&device));
  PLG_CUDA(cudaSetDevice(device));
  nccl_unique_id u;
  memcpy(u.internal, id, sizeof(u.internal));
  nccl_comm comm = NULL;
  int n = g_nccl.comm_init_rank(&comm, nranks, u, rank);
  if (n) return nccl_fail("ncclCommInitRank", n);
  g_nccl.comm = comm;
  g_nccl.nranks = nranks;
  g_nccl.rank = rank;
  g_nccl.device = device;
  return PLG_OK;
}

extern "C" int plg_comm_finalize(void)
{
  if (g_nccl.comm)
  {
    g_nccl.comm_destroy(g_nccl.comm);
    g_nccl.comm = NULL;
  }
  g_nccl.nranks = 0;
  return PLG_OK;
}

extern "C" int plg_comm_size(void) { return g_nccl.comm ? g_nccl.nranks : 0; }

/* A context created while a communicator exists on its device reduces across the ranks. */
bool plg_comm_covers(int device) { return g_nccl.comm != NULL && g_nccl.nranks > 1 && g_nccl.device == device; }

__global__ void k_comm_publish(const double * __restrict__ reduced, double * result, unsigned long long seq)
{
  result[0] = reduced[0];
  result[1] = reduced[1];
  __threadfence_system();
  *reinterpret_cast<volatile unsigned long long *>(result + 4) = seq;
}

/* Enqueued right after a reduction kernel whose sink was the context's device scratch: sums the
 * two doubles over the ranks in place and hands them to the host words with the sequence flag. */
int plg_comm_allreduce_publish(plg_context * ctx, unsigned long long seq)
{
  int n = g_nccl.all_reduce(ctx->comm_buf, ctx->comm_buf, 2, 8 /* ncclFloat64 */, 0 /* ncclSum */, g_nccl.comm,
                            ctx->stream);
  if (n) return nccl_fail("ncclAllReduce", n);
  k_comm_publish<<<1, 1, 0, ctx->stream>>>(ctx->comm_buf, ctx->result_dev, seq);
  PLG_LAUNCH_CHECK(ctx);
  ctx->stats.collectives++;
  return PLG_OK;
}
