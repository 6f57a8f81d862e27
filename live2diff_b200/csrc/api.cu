// Error plumbing + misc exports of the C ABI (include/l2d_b200.h).
#include <atomic>
#include <cstdlib>

#include "common.cuh"

namespace l2d {

static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("L2D_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}

bool pdl_family(int bit) {
  static const int mask = [] {
    const char* e = getenv("L2D_PDL_MASK");
    return e ? atoi(e) : ~0;
  }();
  return (mask >> bit) & 1;
}

}  // namespace l2d

#ifndef L2D_BUILD_HASH_STR
#define L2D_BUILD_HASH_STR "unstamped"
#endif
// sha1 of the sources this library was compiled from (csrc/build.py); build() compares it with the tree
// and _lib.py refuses a library whose hash differs from the sources lying next to it
extern "C" const char l2d_build_hash_marker[] = "L2D_BUILD_HASH=" L2D_BUILD_HASH_STR;
extern "C" const char* l2d_build_hash(void) { return l2d_build_hash_marker + 15; }
extern "C" int l2d_abi_version(void) { return L2D_ABI_VERSION; }
extern "C" const char* l2d_last_error(void) { return l2d::g_last_error.c_str(); }
extern "C" int64_t l2d_launch_count(void) { return l2d::g_launches.load(std::memory_order_relaxed); }
