// libctr_b200: version, error plumbing, device gate (no CPU fallback, sm_100 only).
#include <mutex>

#include "common.cuh"

namespace ctr {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }

int fail_arg(const char* fn, const char* what) {
  set_error(std::string(fn) + ": " + what);
  return CTR_ERR_ARG;
}

int check_cuda(cudaError_t e, const char* fn) {
  if (e == cudaSuccess) return CTR_OK;
  set_error(std::string(fn) + ": CUDA error: " + cudaGetErrorString(e));
  return CTR_ERR_CUDA;
}

static std::mutex g_mu;
static int g_major[64];
static int g_sms[64];
static bool g_seen[64];

static int query_device(int* major, int* sms) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) {
    set_error(std::string("no usable CUDA device: ") + cudaGetErrorString(e));
    cudaGetLastError();
    return CTR_ERR_ARCH;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_seen[dev]) {
    int mj = 0, sm = 0;
    cudaDeviceGetAttribute(&mj, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    g_major[dev] = mj;
    g_sms[dev] = sm;
    g_seen[dev] = true;
  }
  *major = g_major[dev];
  *sms = g_sms[dev];
  return CTR_OK;
}

int ensure_arch() {
  int mj = 0, sm = 0;
  int r = query_device(&mj, &sm);
  if (r != CTR_OK) return r;
  if (mj != 10) {
    set_error("libctr_b200 is built for sm_100a only; current device has compute capability " +
              std::to_string(mj) + ".x (no fallback path exists)");
    return CTR_ERR_ARCH;
  }
  return CTR_OK;
}

static cudaStream_t g_aux[64];
static cudaEvent_t g_fork[64], g_join[64];

bool aux_stream(cudaStream_t* stream, cudaEvent_t* fork, cudaEvent_t* join) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_aux[dev] == nullptr) {
    if (cudaStreamCreateWithFlags(&g_aux[dev], cudaStreamNonBlocking) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&g_fork[dev], cudaEventDisableTiming) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&g_join[dev], cudaEventDisableTiming) != cudaSuccess) return false;
  }
  *stream = g_aux[dev];
  *fork = g_fork[dev];
  *join = g_join[dev];
  return true;
}

static std::string g_opt_names[32];
static int g_opt_values[32];
static int g_opt_n = 0;

int option_get(const char* name, int dflt) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = 0; i < g_opt_n; ++i)
    if (g_opt_names[i] == name) return g_opt_values[i];
  return dflt;
}

int sm_count() {
  int mj = 0, sm = 0;
  if (query_device(&mj, &sm) != CTR_OK) return 148;
  return sm > 0 ? sm : 148;
}

}  // namespace ctr

extern "C" {

int ctr_version(void) { return 100; }

const char* ctr_last_error(void) { return ctr::g_last_error.c_str(); }

int ctr_device_check(void) { return ctr::ensure_arch(); }

int ctr_set_option(const char* name, int value) {
  if (name == nullptr) return ctr::fail_arg("ctr_set_option", "null name");
  const std::string n(name);
  static const char* const known[] = {"bwd_aggregate", "adam_rows_inflight", "adam_rows_bf",
                                      "tcg_dw_stages", "tcg_dw_splits", "fwd_prefetch_record",
                                      "mid_coop", "eb_debug", "tower_dw_splits"};
  bool ok = false;
  for (const char* k : known) ok = ok || n == k;
  if (!ok) return ctr::fail_arg("ctr_set_option", "unknown option");
  std::lock_guard<std::mutex> lk(ctr::g_mu);
  for (int i = 0; i < ctr::g_opt_n; ++i)
    if (ctr::g_opt_names[i] == n) {
      ctr::g_opt_values[i] = value;
      return CTR_OK;
    }
  ctr::g_opt_names[ctr::g_opt_n] = n;
  ctr::g_opt_values[ctr::g_opt_n++] = value;
  return CTR_OK;
}

}  // extern "C"
