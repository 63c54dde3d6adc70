// core.cu -- error reporting and version of libdgs_b200.so
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dgs_internal.h"

namespace dgs {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

unsigned long long g_kernel_launches = 0;
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("DGS_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}
bool g_prof_on = false;

namespace {
struct ProfSpan { cudaEvent_t a, b; int fam; };
std::vector<ProfSpan> g_spans;       // recorded, not yet read
std::vector<cudaEvent_t> g_free;     // recycled events
cudaEvent_t g_open[PROF_N];
cudaEvent_t get_event() {
  if (!g_free.empty()) { cudaEvent_t e = g_free.back(); g_free.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void prof_begin(cudaStream_t st, int family) {
  g_open[family] = get_event();
  cudaEventRecord(g_open[family], st);
}
void prof_end(cudaStream_t st, int family) {
  cudaEvent_t b = get_event();
  cudaEventRecord(b, st);
  g_spans.push_back({g_open[family], b, family});
}
}  // namespace dgs

extern "C" {
int dgs_version(void) { return DGS_VERSION; }
const char* dgs_last_error(void) { return dgs::get_error(); }
unsigned long long dgs_kernel_launch_count(void) { return dgs::g_kernel_launches; }

int dgs_profile_enable(int on) {
  dgs::g_prof_on = on != 0;
  return DGS_OK;
}

int dgs_profile_read(float* ms_sum, int* span_count, int n_families) {
  // synchronises on every recorded span; accumulates elapsed ms per family, then clears the record
  for (int i = 0; i < n_families; i++) { if (ms_sum) ms_sum[i] = 0.f; if (span_count) span_count[i] = 0; }
  for (auto& sp : dgs::g_spans) {
    cudaEventSynchronize(sp.b);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess && sp.fam < n_families) {
      if (ms_sum) ms_sum[sp.fam] += ms;
      if (span_count) span_count[sp.fam]++;
    }
    dgs::g_free.push_back(sp.a);
    dgs::g_free.push_back(sp.b);
  }
  dgs::g_spans.clear();
  return dgs::PROF_N;
}
}
