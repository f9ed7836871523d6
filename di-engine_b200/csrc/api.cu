// Library-level entry points of the C ABI declared in include/b200rl.h.
#include "../../include/b200rl.h"
#include "common.cuh"

extern "C" int b200rl_version(void) { return 100; }
extern "C" int b200rl_built_for_sm(void) { return 100; }
extern "C" size_t b200rl_workspace_bytes(void) { return (size_t)WS_MIN_BYTES; }
