// Library-level entry points: version, error string, launch accounting, run-time tunables.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "common.cuh"

namespace diga {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Index of the current device clamped to [0, 64): key of the per-device caches (cudaFuncSetAttribute is per device).
int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 63;
  return dev;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// Tunables let the GPU-side sweep (tools/tune.py) pick launch shapes without a rebuild.  Defaults are
// the values chosen from those sweeps; DIGA_TUNE_<NAME> in the environment overrides at first use.
static std::mutex g_tune_mu;
static std::map<std::string, int> g_tune;

int tunable(const char* name, int dflt) {
  std::lock_guard<std::mutex> lk(g_tune_mu);
  auto it = g_tune.find(name);
  if (it != g_tune.end()) return it->second;
  std::string env = std::string("DIGA_TUNE_") + name;
  for (auto& ch : env) ch = (char)toupper((unsigned char)ch);
  int v = dflt;
  if (const char* e = getenv(env.c_str())) v = atoi(e);
  g_tune[name] = v;
  return v;
}

}  // namespace diga

extern "C" {

int diga_version(void) { return 100; }  // 0.1.0
const char* diga_last_error_string(void) { return diga::g_err; }
int64_t diga_launch_count(void) { return diga::g_launches.load(std::memory_order_relaxed); }

// Not part of the reference-facing ABI: used by the tuning sweep and the tests.
int diga_set_tunable(const char* name, int value) {
  std::lock_guard<std::mutex> lk(diga::g_tune_mu);
  diga::g_tune[name] = value;
  return 0;
}

}  // extern "C"
