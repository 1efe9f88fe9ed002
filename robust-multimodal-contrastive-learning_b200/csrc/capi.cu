// C-ABI plumbing: per-thread error text, device queries, and the host-buffer step used for the
// end-to-end measurement.  No kernels here.
#include <stdarg.h>

#include <mutex>

#include "common.cuh"

namespace rmcl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static std::mutex mu;
  static int cached[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
    return -1;
  }
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 0 || dev >= 64) return -1;
  if (cached[dev] == 0) {
    int n = 0, major = 0;
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) {
      set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
      return -1;
    }
    if (major != 10) {
      set_error("rmcl_b200 is built for sm_100a only; device %d has compute capability %d.x", dev, major);
      return -1;
    }
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace rmcl

extern "C" const char* rmcl_last_error(void) { return rmcl::g_err; }
extern "C" int rmcl_version(void) { return 1000 * 0 + 1; }
extern "C" int rmcl_sm_count(void) {
  const int n = rmcl::sm_count();
  return n > 0 ? n : RMCL_E_CUDA;
}

// Side stream + events of rmcl_step_host, one set per calling thread and device (created lazily,
// never destroyed: they live as long as the process, like the thread that uses them).
struct HostStepCtx {
  int device = -1;
  cudaStream_t copy = nullptr;
  cudaEvent_t in_ready = nullptr, loss_ready = nullptr;
};
static thread_local HostStepCtx g_hs;

static int host_step_ctx(HostStepCtx** out) {
  int dev = 0;
  RMCL_CUDA_OK(cudaGetDevice(&dev));
  if (g_hs.device != dev) {
    g_hs = HostStepCtx();
    RMCL_CUDA_OK(cudaStreamCreateWithFlags(&g_hs.copy, cudaStreamNonBlocking));
    RMCL_CUDA_OK(cudaEventCreateWithFlags(&g_hs.in_ready, cudaEventDisableTiming));
    RMCL_CUDA_OK(cudaEventCreateWithFlags(&g_hs.loss_ready, cudaEventDisableTiming));
    g_hs.device = dev;
  }
  *out = &g_hs;
  return RMCL_OK;
}

extern "C" int rmcl_step_host(const rmcl_ema_chunk* chunks_dev, int64_t n_chunks, double m,
                              rmcl_dtype param_dtype, const void* q_host,
                              const void* k_host, rmcl_dtype qk_dtype, void* q_dev, void* k_dev, void* queue,
                              rmcl_dtype queue_dtype, int64_t* ptr_dev, int B, int C, int64_t K, float tau, int path,
                              float* loss_dev, float* dq_dev, float* k_hat_dev, float* loss_host, float* dq_host,
                              void* workspace, size_t workspace_bytes, void* stream) {
  RMCL_CHECK_ARG(q_host && k_host && q_dev && k_dev && loss_dev && dq_dev && k_hat_dev && loss_host && dq_host,
                 "rmcl_step_host: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  HostStepCtx* cx = nullptr;
  int rc = host_step_ctx(&cx);
  if (rc != RMCL_OK) return rc;
  // The input copies ride a side stream under the EMA (which needs neither q nor k); the result
  // copies ride it under the enqueue.  The previous call ended with both streams drained, so the
  // device buffers are free to overwrite.
  const size_t qk_bytes = (size_t)B * C * rmcl::dtype_size(qk_dtype);
  RMCL_CUDA_OK(cudaMemcpyAsync(q_dev, q_host, qk_bytes, cudaMemcpyHostToDevice, cx->copy));
  RMCL_CUDA_OK(cudaMemcpyAsync(k_dev, k_host, qk_bytes, cudaMemcpyHostToDevice, cx->copy));
  RMCL_CUDA_OK(cudaEventRecord(cx->in_ready, cx->copy));
  rc = rmcl_ema_multi(chunks_dev, n_chunks, m, param_dtype, stream);
  if (rc != RMCL_OK) return rc;
  RMCL_CUDA_OK(cudaStreamWaitEvent(s, cx->in_ready, 0));
  rc = rmcl_infonce_fwd_bwd(q_dev, qk_dtype, k_dev, qk_dtype, queue, queue_dtype, B, C, K, K, tau, 1.0f,
                            RMCL_INFONCE_NORMALIZE_K, path, loss_dev, nullptr, nullptr, nullptr, nullptr, dq_dev,
                            nullptr, k_hat_dev, workspace, workspace_bytes, stream);
  if (rc != RMCL_OK) return rc;
  RMCL_CUDA_OK(cudaEventRecord(cx->loss_ready, s));
  RMCL_CUDA_OK(cudaStreamWaitEvent(cx->copy, cx->loss_ready, 0));
  RMCL_CUDA_OK(cudaMemcpyAsync(loss_host, loss_dev, sizeof(float), cudaMemcpyDeviceToHost, cx->copy));
  RMCL_CUDA_OK(cudaMemcpyAsync(dq_host, dq_dev, (size_t)B * C * sizeof(float), cudaMemcpyDeviceToHost, cx->copy));
  rc = rmcl_enqueue(queue, queue_dtype, k_hat_dev, RMCL_F32, ptr_dev, B, C, K, K, stream);
  if (rc != RMCL_OK) return rc;
  RMCL_CUDA_OK(cudaStreamSynchronize(cx->copy));
  RMCL_CUDA_OK(cudaStreamSynchronize(s));
  return RMCL_OK;
}
