// Compatibility shim force-included (-include) into every translation unit of the
// headless build of the REFERENCE renderer (see build_ref.sh).  TEST INFRASTRUCTURE ONLY.
//
// It supplies what the reference's MSVC/Win32 dialect expects and nothing else; no arithmetic of
// the reference is touched.  Std headers are pulled in FIRST so that the `private -> public`
// define below (which lets ref_driver.cpp walk GeometrySet::geometries, Tunnel::grid and
// Tunnel::root to number geometries / cells / nodes) never reaches a system header.
#ifndef RTB_REF_COMPAT_H
#define RTB_REF_COMPAT_H

#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <math.h>
#include <float.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <map>
#include <queue>
#include <string>
#include <unordered_map>
#include <vector>
#include <mutex>
#include <chrono>
#include <omp.h>

#define __forceinline inline
#ifndef _countof
#define _countof(a) (sizeof(a) / sizeof((a)[0]))
#endif

static inline int fopen_s(FILE **fp, const char *name, const char *mode)
{
    *fp = fopen(name, mode);
    return *fp ? 0 : 1;
}
#define vsprintf_s(buf, size, fmt, args) vsnprintf((buf), (size), (fmt), (args))

// Path of the binary STL fixture used by preset 3 (Scripts.cpp:151 names a cwd-relative file).
const char *ref_stl_path();

// Observation hooks (sed-inserted by build_ref.sh next to existing statements; REF_HOOKS=0
// compiles them out for the timing build).
#if REF_HOOKS
void ref_hook_cell(int linearIndex);
void ref_hook_node(const void *node);
void ref_hook_tri();
void ref_hook_ray();
#define REF_HOOK_CELL(idx) ref_hook_cell(idx)
#define REF_HOOK_NODE(n) ref_hook_node((const void *)(n))
#define REF_HOOK_TRI() ref_hook_tri()
#define REF_HOOK_RAY() ref_hook_ray()
#else
#define REF_HOOK_CELL(idx) ((void)0)
#define REF_HOOK_NODE(n) ((void)0)
#define REF_HOOK_TRI() ((void)0)
#define REF_HOOK_RAY() ((void)0)
#endif

#define private public
#define protected public

#endif
