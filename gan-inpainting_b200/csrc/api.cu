// Error reporting and library identity for the C-ABI (see include/b200_inpaint.h).
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

static thread_local char g_err[512] = "";

void b200_set_error(const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
}

static long long g_launches = 0;
void b200_count_launch(void) { ++g_launches; }

// Programmatic dependent launch: B200_PDL=1 every kernel, B200_PDL=2 every kernel EXCEPT the persistent tcgen05 GEMMs (a GEMM needs
// all 227 KB of shared memory of every SM, so letting it become resident early buys nothing and its spinning CTAs delay the
// predecessor's tail), unset / 0: off.  Measured on B200 with the step replayed as a CUDA graph: see profiles/experiments/README.md.
static int pdl_mode(void) {
  static const int mode = [] {
    const char* e = getenv("B200_PDL");
    return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 0;
  }();
  return mode;
}
bool b200_pdl_enabled(void) { return pdl_mode() != 0; }
bool b200_pdl_gemm_enabled(void) { return pdl_mode() == 1; }

static const uint32_t* g_seed_counter = nullptr;
const uint32_t* b200_seed_counter_ptr(void) { return g_seed_counter; }

static int g_sm_limit = 148;
int b200_sm_limit(void) { return g_sm_limit; }

extern "C" {
// SMs the persistent GEMM kernels may occupy (one CTA per SM, static tile schedule).  A data-parallel trainer lowers it while
// gradient all-reduces are in flight: the collective's CTAs then never take an SM a persistent CTA was counting on (a CTA that
// starts late stretches the whole launch).  n <= 0 or n > 148 restores all SMs.
void b200_set_sm_limit(int n) { g_sm_limit = (n <= 0 || n > 148) ? 148 : n; }
// device uint32 added (x odd constant) to the seed of every dropout mask drawn by kernels enqueued afterwards; NULL = off
void b200_set_seed_counter(const unsigned* dev_counter) { g_seed_counter = dev_counter; }
// replaying a captured graph launches kernels this library enqueued at capture time: account for them
void b200_launch_count_add(long long n) { g_launches += n; }
const char* b200_last_error(void) { return g_err; }
// number of kernels this library has launched since load (bench.py's gpu_launches evidence)
long long b200_launch_count(void) { return g_launches; }
int b200_abi_version(void) { return 1; }
}
