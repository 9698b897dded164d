#!/usr/bin/env bash
# Headless build of the UNMODIFIED reference renderer (windy32/win32-ray-tracing-demo,
# src/RayTracingOpt) into oracle/_ref/libref.so (+ libref_timing.so without observation hooks).
#
# TEST INFRASTRUCTURE ONLY.  The sources are compiled from where they lie under
# $RTB_REFERENCE (default /root/reference); a scratch copy with the mechanical MSVC->g++ patches
# listed below is made under a temp dir and deleted afterwards.  Nothing but the two .so files
# (and the STL fixture path note) is written into the repo, and oracle/_ref/ is git-ignored.
#
# Patches (none touches arithmetic):
#   P1 RayContext.h:10      `struct RayContext() :` -> `RayContext() :`          (MSVC-only syntax)
#   P2 Camera.h:21, .cpp:5  `Vector &front` -> `Vector front`                     (rvalue -> non-const ref)
#   P3 Tunnel.cpp:46        name the temporary Ray                                (same)
#   P4 MainWindow.cpp:69,145 trace/radiance take `Ray r` by value                 (same)
#   P5 Color.h:12           `Color()` zero-initialises                            (removes UB at MainWindow.cpp:92-94)
#   P6 Scripts.cpp:151      "ball.stl" -> ref_stl_path()                          (fixture location)
#   P7 observation hooks    REF_HOOK_* macros after Tunnel.cpp:866,1206,1266, Triangle.cpp:40,
#                           and the scene.intersect() call of trace/radiance       (compiled out when REF_HOOKS=0)
#   Utils.cpp (Win32) is replaced by oracle/ref/ref_utils.cpp; MainWindow.cpp is not compiled:
#   only its lines 69-249 (trace, radiance), 251-303 (Render's pixel loop) and 305-312 (the 8-bit
#   output stage; SetPixel / RGB are stand-ins in ref_driver.cpp) are lifted into a
#   generated include used by oracle/ref/ref_driver.cpp.
set -euo pipefail

HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${RTB_REFERENCE:-/root/reference}"
SRC="$REF/src/RayTracingOpt"
OUT="$HERE/_ref"

if [ ! -d "$SRC" ]; then
    echo "build_ref.sh: reference sources not found at $SRC (nothing built)" >&2
    exit 3
fi

TMP="$(mktemp -d /tmp/rtb_ref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT"

cp "$SRC"/*.h "$SRC"/*.cpp "$TMP"/
rm -f "$TMP/MainWindow.cpp" "$TMP/Utils.cpp" "$TMP/resource.h"

expect() { # expect <file> <line> <substring>: guard every line-addressed patch
    sed -n "${2}p" "$SRC/$1" | grep -qF -- "$3" || { echo "build_ref.sh: $1:$2 does not contain '$3'" >&2; exit 4; }
}

expect RayContext.h 10 'struct RayContext() :'
sed -i '10s/struct RayContext() :/RayContext() :/' "$TMP/RayContext.h"                       # P1
expect Camera.h 21 'Vector &front'
sed -i '21s/Vector &front/Vector front/' "$TMP/Camera.h"                                      # P2
expect Camera.cpp 5 'Vector &front'
sed -i '5s/Vector &front/Vector front/' "$TMP/Camera.cpp"
expect Tunnel.cpp 46 'intersectWithPolygonAtOrigin(Ray(newOrigin, newDir), distance)'
sed -i '46s/.*/    Ray refTmpRay(newOrigin, newDir); bool intersect = intersectWithPolygonAtOrigin(refTmpRay, distance);/' "$TMP/Tunnel.cpp"  # P3
expect Color.h 12 'Color()'
sed -i '12s/Color()/Color() : r(0), g(0), b(0)/' "$TMP/Color.h"                              # P5
expect Scripts.cpp 151 '"ball.stl"'
sed -i '151s/"ball.stl"/ref_stl_path()/' "$TMP/Scripts.cpp"                                   # P6
expect Tunnel.cpp 866 'grid.get(cur_i, cur_j, cur_k);'
sed -i '866s/$/ REF_HOOK_CELL((cur_i * grid.yLength + cur_j) * grid.zLength + cur_k);/' "$TMP/Tunnel.cpp"   # P7
expect Tunnel.cpp 1206 'float splitVal = currNode->splitPlane;'
sed -i '1206s/$/ REF_HOOK_NODE(currNode);/' "$TMP/Tunnel.cpp"
expect Tunnel.cpp 1266 'float minDistance = FLT_MAX;'
sed -i '1266s/$/ REF_HOOK_NODE(currNode);/' "$TMP/Tunnel.cpp"
expect Triangle.cpp 40 'IntersectResult result(false);'
sed -i '40s/$/ REF_HOOK_TRI();/' "$TMP/Triangle.cpp"

# Lift trace() / radiance() / Render()'s loop out of MainWindow.cpp (never copied into the repo).
expect MainWindow.cpp 69 'Color trace(GeometrySet &scene, Ray &r, int depth'
expect MainWindow.cpp 145 'Color radiance(GeometrySet &scene, Ray &r, int depth'
expect MainWindow.cpp 251 'int Render(GeometrySet &scene, PerspectiveCamera &camera'
expect MainWindow.cpp 303 'int t2 = Utils::GetTickCount();'
expect MainWindow.cpp 311 'SetPixel(hdcBuffer, i / height, i % height, RGB(r, g, b));'
{
    sed -n '69,249p' "$SRC/MainWindow.cpp" \
        | sed -e 's/Ray &r, int depth/Ray r, int depth/' \
              -e 's/IntersectResult result = scene.intersect(r);/IntersectResult result = scene.intersect(r); REF_HOOK_RAY();/'   # P4, P7
    sed -n '251,303p' "$SRC/MainWindow.cpp"
    echo '    ref_render_epilogue(colors);   // float framebuffer, before the output stage saturates it in place'
    sed -n '305,312p' "$SRC/MainWindow.cpp"   # the output stage: saturate, (int)(c*255), SetPixel
    echo '    delete []colors;'
    echo '    return t2 - t1;'
    echo '}'
} > "$TMP/ref_lifted_render.inc"

CXXFLAGS="-O2 -std=c++14 -fopenmp -fpermissive -w -Wno-narrowing -ffp-contract=off -fPIC -include $HERE/ref/compat.h -I$TMP -I$HERE"
SOURCES="Vector.cpp Point.cpp Matrix.cpp Camera.cpp Geometry.cpp GeometrySet.cpp Plane.cpp Sphere.cpp
Triangle.cpp Grid.cpp Polygon.cpp Tunnel.cpp TunnelGenerator.cpp Material.cpp SolidColorMaterial.cpp
CheckerMaterial.cpp RadianceCheckerMaterial.cpp PhongMaterial.cpp GlassMaterial.cpp RandomColorMaterial.cpp
Scripts.cpp"

build_variant() { # <hooks 0|1> <output>
    local objs=()
    mkdir -p "$TMP/obj$1"
    for f in $SOURCES; do
        g++ $CXXFLAGS -DREF_HOOKS=$1 -c "$TMP/$f" -o "$TMP/obj$1/${f%.cpp}.o" &
        objs+=("$TMP/obj$1/${f%.cpp}.o")
    done
    g++ $CXXFLAGS -DREF_HOOKS=$1 -c "$HERE/ref/ref_utils.cpp" -o "$TMP/obj$1/ref_utils.o" &
    g++ $CXXFLAGS -DREF_HOOKS=$1 -c "$HERE/ref/ref_driver.cpp" -o "$TMP/obj$1/ref_driver.o" &
    wait
    g++ -shared -fopenmp -o "$2" "${objs[@]}" "$TMP/obj$1/ref_utils.o" "$TMP/obj$1/ref_driver.o"
}

build_variant 1 "$OUT/libref.so"
build_variant 0 "$OUT/libref_timing.so"
echo "built $OUT/libref.so $OUT/libref_timing.so"

# ---- libref_sat.so: the reference with its EXACT grid binning switched on (patch P8) -- the checker of the exact-binning
# option of the product's grid builders.  Tunnel.cpp:435 `#if 0` guards the branch that calls Triangle::intersectWithGrid
# (Triangle.cpp:152-199); it names a variable `size` that only exists inside the regular-grid block above (lines 377-391), so
# the branch does not compile as it stands: P8 flips the guard and declares `float size = grid.cellSizeX;` in front of the
# triangle loop (the regular grid's cells are cubes of that edge; for the flat grid the reference's branch would be wrong,
# so only regular grids are compared with this build).  Nothing else differs from libref.so.
expect Tunnel.cpp 435 '#if 0'
expect Tunnel.cpp 411 '// For each triangle'
cp "$TMP/Tunnel.cpp" "$TMP/Tunnel.cpp.plain"
sed -i -e '435s/#if 0/#if 1/' -e '411s/$/ float size = grid.cellSizeX;/' "$TMP/Tunnel.cpp"
g++ $CXXFLAGS -DREF_HOOKS=1 -c "$TMP/Tunnel.cpp" -o "$TMP/obj1/Tunnel_sat.o"
mv "$TMP/Tunnel.cpp.plain" "$TMP/Tunnel.cpp"
sobjs=()
for f in $SOURCES; do if [ "$f" = Tunnel.cpp ]; then sobjs+=("$TMP/obj1/Tunnel_sat.o"); else sobjs+=("$TMP/obj1/${f%.cpp}.o"); fi; done
g++ -shared -fopenmp -o "$OUT/libref_sat.so" "${sobjs[@]}" "$TMP/obj1/ref_utils.o" "$TMP/obj1/ref_driver.o"
echo "built $OUT/libref_sat.so"

# ---- Route A of INTEGRATION.md, compiled: the reference's own objects (hooks off) + ref/route_a.cpp (SceneFlattener over the
# reference's classes, CudaRender as the RenderProc) + the product's librtb200.so -> libroute_a.so (tests/test_route_a.py)
RTB_PKG="$(cd "$HERE/.." && pwd)/win32-ray-tracing-demo_b200"
if [ -f "$RTB_PKG/librtb200.so" ]; then
    g++ $CXXFLAGS -DREF_HOOKS=0 -c "$HERE/ref/route_a.cpp" -o "$TMP/obj0/route_a.o"
    robjs=()
    for f in $SOURCES; do robjs+=("$TMP/obj0/${f%.cpp}.o"); done
    g++ -shared -fopenmp -o "$OUT/libroute_a.so" "${robjs[@]}" "$TMP/obj0/ref_utils.o" "$TMP/obj0/route_a.o" \
        -L"$RTB_PKG" -lrtb200 -Wl,-rpath,'$ORIGIN/../../win32-ray-tracing-demo_b200'
    echo "built $OUT/libroute_a.so"
else
    echo "build_ref.sh: $RTB_PKG/librtb200.so not built yet -- libroute_a.so skipped" >&2
fi

# ---- PerformanceTest (src/PerformanceTest): its own sources + ref/ref_pt_driver.cpp -> libref_pt.so ----------
# main.cpp (console front end) and Utils.cpp (Win32) are replaced by the driver; two mechanical patches, both the
# MSVC-dialect kind already applied to RayTracingOpt above (P1, P3), none touching arithmetic.
PT="$REF/src/PerformanceTest"
if [ -d "$PT" ]; then
    mkdir -p "$TMP/pt"
    cp "$PT"/*.h "$PT"/*.cpp "$TMP/pt/"
    rm -f "$TMP/pt/main.cpp" "$TMP/pt/Utils.cpp"
    expect_pt() { sed -n "${2}p" "$PT/$1" | grep -qF -- "$3" || { echo "build_ref.sh: PerformanceTest/$1:$2 does not contain '$3'" >&2; exit 4; }; }
    expect_pt RayContext.h 10 'struct RayContext() :'
    sed -i '10s/struct RayContext() :/RayContext() :/' "$TMP/pt/RayContext.h"
    expect_pt ConvexAcc.cpp 23 'intersectWithPolygonAtOrigin(Ray(newOrigin, newDir), distance)'
    sed -i '23s/.*/    Ray refTmpRay(newOrigin, newDir); bool intersect = intersectWithPolygonAtOrigin(refTmpRay, distance);/' "$TMP/pt/ConvexAcc.cpp"
    PTFLAGS="-O2 -std=c++14 -fopenmp -fpermissive -w -Wno-narrowing -ffp-contract=off -fPIC -include $HERE/ref/compat.h -I$TMP/pt -I$HERE"
    ptobjs=()
    for f in "$TMP"/pt/*.cpp; do
        g++ $PTFLAGS -c "$f" -o "${f%.cpp}.o" &
        ptobjs+=("${f%.cpp}.o")
    done
    g++ $PTFLAGS -c "$HERE/ref/ref_pt_driver.cpp" -o "$TMP/pt/ref_pt_driver.o" &
    wait
    g++ -shared -fopenmp -o "$OUT/libref_pt.so" "${ptobjs[@]}" "$TMP/pt/ref_pt_driver.o"
    echo "built $OUT/libref_pt.so"
fi
