// api.cu -- ABI bookkeeping (version, thread-local last-error string).
#include "common.cuh"
#include <string.h>

namespace usip {
static thread_local char g_last_error[256] = "";
void set_last_error(const char* what) {
  strncpy(g_last_error, what ? what : "", sizeof(g_last_error) - 1);
  g_last_error[sizeof(g_last_error) - 1] = 0;
}
}  // namespace usip

extern "C" int usip_abi_version(void) { return USIP_B200_ABI_VERSION; }
extern "C" const char* usip_last_error(void) { return usip::g_last_error; }
