// Library-wide state of the C ABI (error string, launch counter).
#include <stdlib.h>
#include "common.cuh"
#include "../../include/q3tts_b200.h"
namespace q3t {
thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
static int env_pdl() { const char* e = getenv("Q3T_PDL"); return (e && e[0] == '0') ? 0 : 1; }
int g_use_pdl = env_pdl();
}
extern "C" int q3t_abi_version(void) { return Q3T_ABI_VERSION; }
extern "C" const char* q3t_last_error(void) { return q3t::g_err; }
extern "C" unsigned long long q3t_launch_count(void) { return q3t::g_launches; }
