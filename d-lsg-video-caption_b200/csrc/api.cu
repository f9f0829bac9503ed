// C-ABI glue: error reporting, version, GEMM dispatch (see include/dlsg.h).
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include "common.cuh"

namespace dlsg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return (int)e;
  }
  return 0;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DLSG_PDL"); v = (e && e[0] == '0') ? 0 : 1; }   // on by default; DLSG_PDL=0 disables
  return v == 1;
}

int gemm_tc_dispatch(const dlsg_gemm_t* g, cudaStream_t st);
int gemm_simt_dispatch(const dlsg_gemm_t* g, cudaStream_t st);
void gemm_tc_set_trace(void* p);

}  // namespace dlsg

extern "C" {

int dlsg_version(void) { return 100; }
int dlsg_sm_arch(void) { return 100; }
const char* dlsg_last_error(void) { return dlsg::g_err; }
void dlsg_debug_gemm_trace(void* dev_buf) { dlsg::gemm_tc_set_trace(dev_buf); }

int dlsg_gemm(const dlsg_gemm_t* p, void* stream) {
  if (!p) { dlsg::set_error("dlsg_gemm: null params"); return -1; }
  if (p->impl == DLSG_GEMM_TC) return dlsg::gemm_tc_dispatch(p, (cudaStream_t)stream);
  if (p->flags & DLSG_EPI_ATOMIC) {
    dlsg_gemm_t q = *p;
    q.flags = (q.flags & ~DLSG_EPI_ATOMIC) | DLSG_EPI_ACCUM;
    return dlsg::gemm_simt_dispatch(&q, (cudaStream_t)stream);
  }
  return dlsg::gemm_simt_dispatch(p, (cudaStream_t)stream);
}

}  // extern "C"
