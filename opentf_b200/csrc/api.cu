// api.cu -- handle, error string, version; no kernels.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";
unsigned long long g_ntf_launches = 0;

void ntf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int ntf_version(void) { return NTF_ABI_VERSION; }

extern "C" int ntf_last_error(char* buf, size_t n) {
  if (!buf || n == 0) return NTF_ERR_BAD_ARG;
  strncpy(buf, g_err, n - 1);
  buf[n - 1] = 0;
  return NTF_OK;
}

extern "C" int ntf_create(int device, ntf_ctx** out) {
  NTF_REQUIRE(out != nullptr, NTF_ERR_BAD_ARG, "ntf_create: out is NULL");
  *out = nullptr;
  int count = 0;
  NTF_CUDA(cudaGetDeviceCount(&count));
  NTF_REQUIRE(device >= 0 && device < count, NTF_ERR_BAD_ARG, "ntf_create: device %d of %d", device, count);
  cudaDeviceProp prop;
  NTF_CUDA(cudaGetDeviceProperties(&prop, device));
  // no CPU fallback and no other architecture: the kernels are compiled for sm_100a only
  NTF_REQUIRE(prop.major == 10, NTF_ERR_DEVICE, "ntf_create: device %d is sm_%d%d, libntf_b200 needs sm_100 (B200)",
              device, prop.major, prop.minor);
  ntf_ctx* c = new ntf_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  c->encode_tiled = nullptr;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) c->encode_tiled = fn;
  (void)cudaGetLastError();
  *out = c;
  return NTF_OK;
}

extern "C" int ntf_destroy(ntf_ctx* ctx) {
  delete ctx;
  return NTF_OK;
}

extern "C" int ntf_sm_count(const ntf_ctx* ctx) { return ctx ? ctx->sm_count : NTF_ERR_BAD_ARG; }

extern "C" unsigned long long ntf_launch_count(int reset) {
  const unsigned long long v = g_ntf_launches;
  if (reset) g_ntf_launches = 0;
  return v;
}
