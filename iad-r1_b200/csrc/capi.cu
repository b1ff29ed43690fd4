// extern "C" surface of libiadr1_b200.so (declared in include/iadr1_b200.h).
#include "runtime.h"
#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace iadr1 {
static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};
int set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static std::atomic<bool> g_pdl{false};
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed); }
void pdl_set(bool on) { g_pdl.store(on); }
}  // namespace iadr1

extern "C" {
const char* iadr1_last_error(void) { return iadr1::g_err; }
int iadr1_version(void) { return 100; }
long long iadr1_launch_count(void) { return iadr1::g_launches.load(); }
void iadr1_reset_launch_count(void) { iadr1::g_launches.store(0); }
int iadr1_gemm_bf16(const iadr1_gemm_t* d, void* stream) {
  if (!d) return iadr1::set_error("null gemm descriptor");
  return iadr1::launch_gemm(*d, static_cast<cudaStream_t>(stream));
}
int iadr1_gemm_pick_block_n(int N, int b_mn) { return iadr1::pick_block_n_public(N, b_mn); }
int iadr1_set_pdl(int on) {
  iadr1::pdl_set(on != 0);
  return 0;
}
int iadr1_gemm_profile_enable(int on) {
  iadr1::gemm_profile_enable(on);
  return 0;
}
int iadr1_gemm_profile_collect(double* total_ms, double* total_flops, double* max_launch_ms, long long* launches,
                               const char* csv_path) {
  const long long n = iadr1::gemm_profile_collect(total_ms, total_flops, max_launch_ms, csv_path);
  if (launches) *launches = n;
  return 0;
}
}
