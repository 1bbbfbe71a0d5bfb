// ABI version + launch accounting.
#include "../../include/m3t_b200.h"
#include "common.cuh"

namespace m3t {
long long g_launch_count = 0;
}

extern "C" int m3t_abi_version(void) { return 1; }
extern "C" long long m3t_launch_count(void) { return __atomic_load_n(&m3t::g_launch_count, __ATOMIC_RELAXED); }
