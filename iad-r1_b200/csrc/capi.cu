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
int iadr1_trace(int op, unsigned long long* out, int max_words) {
  static unsigned long long* buf = nullptr;
  const size_t words = 2 + 2 * 8190;
  if (op == 1) {          // install + reset
    if (!buf && cudaMalloc(&buf, words * 8) != cudaSuccess) return iadr1::set_error("trace: cudaMalloc failed");
    cudaMemset(buf, 0, words * 8);
    iadr1::trace_install_decode(buf);
    iadr1::trace_install_gemm(buf);
    iadr1::trace_install_rowops(buf);
    iadr1::trace_install_chain(buf);
  } else if (op == 0) {   // remove
    iadr1::trace_install_decode(nullptr);
    iadr1::trace_install_gemm(nullptr);
    iadr1::trace_install_rowops(nullptr);
    iadr1::trace_install_chain(nullptr);
  } else if (op == 2) {   // collect (device must be idle)
    if (!buf || !out) return iadr1::set_error("trace: nothing to collect");
    const size_t n = (size_t)max_words < words ? (size_t)max_words : words;
    if (cudaMemcpy(out, buf, n * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return iadr1::set_error("trace: copy failed");
  }
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
