// profc.h -- shadows the reference's src/profc.h (found first through -I ref_shim).
// TEST INFRASTRUCTURE.  The reference's scoped profiler is not part of the hot path's
// arithmetic: it would time every relax call with std::chrono + a mutex, print a table into the
// test output at process exit, and that exit-time report reads function-local static
// ProfileNode objects from the destructor of a singleton that outlives some of them (static
// destruction order), which is undefined behaviour inside a long-lived host process.
#ifndef SMG_REF_SHIM_PROFC
#define SMG_REF_SHIM_PROFC
#define PROFC_NODE(name)
#endif
