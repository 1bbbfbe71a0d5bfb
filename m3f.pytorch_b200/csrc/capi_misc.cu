// ABI version + launch accounting.
#include "../../include/m3t_b200.h"
#include "common.cuh"
#include <cstdlib>

namespace m3t {
long long g_launch_count = 0;
int g_sm_reserve = 0;
int usable_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms = n;
  }
  const int r = __atomic_load_n(&g_sm_reserve, __ATOMIC_RELAXED);
  return sms - r > 8 ? sms - r : 8;
}
int g_pdl = -1;
bool pdl_enabled() {
  int v = __atomic_load_n(&g_pdl, __ATOMIC_RELAXED);
  if (v < 0) {
    const char* e = getenv("M3T_PDL");
    v = (e && e[0] == '1') ? 1 : 0;     // off unless asked for: see engine.py capture() for where it pays
    __atomic_store_n(&g_pdl, v, __ATOMIC_RELAXED);
  }
  return v != 0;
}
}

extern "C" int m3t_abi_version(void) { return 1; }
extern "C" long long m3t_launch_count(void) { return __atomic_load_n(&m3t::g_launch_count, __ATOMIC_RELAXED); }
extern "C" int m3t_set_pdl(int on) {
  const int prev = m3t::pdl_enabled() ? 1 : 0;
  if (on >= 0) __atomic_store_n(&m3t::g_pdl, on ? 1 : 0, __ATOMIC_RELAXED);
  return prev;
}
extern "C" int m3t_set_sm_reserve(int n) {
  const int prev = __atomic_load_n(&m3t::g_sm_reserve, __ATOMIC_RELAXED);
  if (n >= 0) __atomic_store_n(&m3t::g_sm_reserve, n, __ATOMIC_RELAXED);
  return prev;
}
