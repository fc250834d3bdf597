// Library plumbing: error string, version, launch counter.
#include <stdarg.h>
#include <stdlib.h>
#include "common.cuh"

namespace davf {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
std::atomic<int64_t> g_launch_kind[kNumKinds];
std::atomic<int> g_pdl{[] { const char* e = getenv("DAVF_PDL"); return e ? atoi(e) : 0; }()};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace davf

extern "C" {
const char* davf_last_error(void) { return davf::g_err; }
int davf_version(void) { return 1; }
int davf_device_sm(void) {
  int dev = 0, major = 0, minor = 0;
  DAVF_CUDA(cudaGetDevice(&dev));
  DAVF_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DAVF_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}
int64_t davf_launch_count(void) { return davf::g_launches.load(); }
int davf_set_pdl(int on) { davf::g_pdl.store(on ? 1 : 0); return DAVF_OK; }
int64_t davf_launch_count_kind(int kind) {
  if (kind <= 0 || kind >= davf::kNumKinds) return davf::g_launches.load();
  return davf::g_launch_kind[kind].load();
}
}
