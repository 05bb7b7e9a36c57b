// Library-level pieces of the C ABI: error text, device check, pipeline diagnostics.
#include <atomic>
#include <mutex>
#include <string.h>

#include <cstddef>
#include "common.cuh"
#include "../../include/iiseg.h"

namespace iiseg {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static int32_t* g_diag_host = nullptr;
static int32_t* g_diag_dev = nullptr;
static std::once_flag g_diag_once;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  return dev;
}

static std::atomic<int> g_sm_reserve{0};

int num_sms() {           // per device ordinal: a process may drive several GPUs
  static int sms[kMaxDevices] = {0};
  const int dev = current_device();
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev] = n;
  }
  const int r = g_sm_reserve.load(std::memory_order_relaxed);
  return (r > 0 && r < sms[dev] - 16) ? sms[dev] - r : sms[dev];
}

// 64 pinned, device-mapped words.  A kernel whose mbarrier wait times out writes
// {magic, code, block, aux} here before trapping; host memory survives the dead context.
int32_t* diag_device_ptr() {
  std::call_once(g_diag_once, [] {
    void* h = nullptr;
    if (cudaHostAlloc(&h, 64 * sizeof(int32_t), cudaHostAllocMapped) == cudaSuccess) {
      memset(h, 0, 64 * sizeof(int32_t));
      void* d = nullptr;
      if (cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
        g_diag_host = static_cast<int32_t*>(h);
        g_diag_dev = static_cast<int32_t*>(d);
      }
    }
  });
  return g_diag_dev;
}

}  // namespace iiseg

extern "C" int iiseg_abi_version(void) { return IISEG_ABI_VERSION; }
extern "C" int iiseg_reserve_sms(int n) {
  const int old = iiseg::g_sm_reserve.exchange(n < 0 ? 0 : n, std::memory_order_relaxed);
  return old;
}
extern "C" int iiseg_conv_desc_size(void) { return (int)sizeof(iiseg_conv_desc); }
extern "C" int iiseg_conv_desc_last_offset(void) { return (int)offsetof(iiseg_conv_desc, upd_cpad); }
extern "C" int iiseg_deconv_desc_size(void) { return (int)sizeof(iiseg_deconv_desc); }

extern "C" const char* iiseg_last_error(void) { return iiseg::g_err; }

extern "C" int iiseg_device_check(int dev) {
  using namespace iiseg;
  int major = 0, minor = 0;
  IISEG_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  IISEG_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  IISEG_CHECK(major == 10, "device %d is sm_%d%d; libiiseg is built for sm_100a only", dev, major, minor);
  return 0;
}

extern "C" int iiseg_read_diag(int32_t* out, int n) {
  if (iiseg::g_diag_host == nullptr || out == nullptr) return 0;
  if (n > 64) n = 64;
  for (int i = 0; i < n; ++i) out[i] = iiseg::g_diag_host[i];
  return n;
}

extern "C" int64_t iiseg_launch_count(void) { return iiseg::g_launches.load(std::memory_order_relaxed); }
