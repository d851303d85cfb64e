/* oracle/ref_shim/lemon/config.h — stands in for the file the reference's CMake generates from
 * third_party/lemon/lemon/config.h.in (no LP / MIP solver back-ends: openMVG's default build has none).
 * Test infrastructure only: lets oracle/Makefile compile the reference's own main_ComputeMatches where it lies. */
#define LEMON_VERSION "1.3"
#define LEMON_HAVE_LONG_LONG 1
#define LEMON_USE_PTHREAD 1
