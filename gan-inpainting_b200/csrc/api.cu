// Error reporting and library identity for the C-ABI (see include/b200_inpaint.h).
#include "common.cuh"

#include <string.h>

static thread_local char g_err[512] = "";

void b200_set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

static long long g_launches = 0;
void b200_count_launch(void) { ++g_launches; }

extern "C" {
const char* b200_last_error(void) { return g_err; }
// number of kernels this library has launched since load (bench.py's gpu_launches evidence)
long long b200_launch_count(void) { return g_launches; }
int b200_abi_version(void) { return 1; }
}
