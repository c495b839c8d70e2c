// C-ABI plumbing shared by every entry point of libdad3d.so: thread-local error text, launch counter, version.
#include <atomic>
#include <string>

#include "../../include/dad3d.h"
#include "common.h"

namespace dad3d {
static thread_local std::string t_last_error;
std::atomic<unsigned long long> g_launches{0};
void set_error(const std::string& msg) { t_last_error = msg; }
}  // namespace dad3d

extern "C" {
const char* dad3d_last_error(void) { return dad3d::t_last_error.c_str(); }
int dad3d_version(void) { return 100; }
unsigned long long dad3d_launch_count(void) { return dad3d::g_launches.load(std::memory_order_relaxed); }
}
