#!/bin/bash
# Drop-in demonstration: configure and build the reference's vendored ggml with -DGGML_TMAC=ON against THIS
# library's CMake package, with ggml/src/ggml-tmac.cpp replaced by t-mac_b200/ggml/ggml-tmac.cpp.  The reference tree is
# read-only: it is copied to a scratch directory first.  Needs /root/reference (build container only).
#   -DGGML_TMAC_TVM_THREADPOOL=ON selects the branch of ggml.c that calls the hook once per mat-vec from thread 0
#   (ref:ggml.c:12610-12630) -- the right shape for a GPU backend -- and links the package target `t_mac`.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${REF:-/root/reference/3rdparty/llama.cpp}
W=${W:-/tmp/ggml_tmac_build}
rm -rf "$W"; mkdir -p "$W"
# 1. install the package (library built by ./build.sh is reused: configure-only install of headers + config + kcfg)
PREFIX="$W/prefix"
mkdir -p "$PREFIX/lib/cmake/TMAC" "$PREFIX/include"
cp "$ROOT/t-mac_b200/libtmac_b200.so" "$PREFIX/lib/"
cp "$ROOT/include/tmac_b200.h" "$PREFIX/include/"
cp -r "$ROOT/t-mac_b200/include/t-mac" "$PREFIX/include/"
python "$ROOT/tools/make_kcfg.py" --preset llama-2-7b-2bit --out "$PREFIX/lib/kcfg.ini" >/dev/null
sed -e "s|@PACKAGE_INIT@|macro(set_and_check v p)\n  set(\${v} \"\${p}\")\nendmacro()\nmacro(check_required_components n)\nendmacro()|" \
    -e "s|@PACKAGE_TMAC_LIB_INSTALL_DIR@|$PREFIX/lib|" -e "s|@PACKAGE_TMAC_INCLUDE_INSTALL_DIR@|$PREFIX/include|" \
    "$ROOT/cmake/TMACConfig.cmake.in" > "$PREFIX/lib/cmake/TMAC/TMACConfig.cmake"
# 2. scratch copy of the vendored ggml with the hook file swapped
cp -r "$REF/ggml" "$W/ggml"
cp "$ROOT/t-mac_b200/ggml/ggml-tmac.cpp" "$W/ggml/src/ggml-tmac.cpp"
# 3. configure + build libggml (through a two-line parent project, as llama.cpp's own top level does: ggml is not standalone)
printf 'cmake_minimum_required(VERSION 3.14)\nproject(ggml_tmac_demo C CXX)\nadd_subdirectory(ggml)\n' > "$W/CMakeLists.txt"
cmake -S "$W" -B "$W/build" -DGGML_TMAC=ON -DGGML_TMAC_TVM_THREADPOOL=ON -DCMAKE_PREFIX_PATH="$PREFIX" \
      -DGGML_NATIVE=OFF -DBUILD_SHARED_LIBS=ON -DCMAKE_BUILD_TYPE=Release > "$W/configure.log" 2>&1 || { tail -20 "$W/configure.log"; exit 1; }
grep -E "TMAC found|TMAC not found" "$W/configure.log"
cmake --build "$W/build" -j 8 --target ggml > "$W/build.log" 2>&1 || { grep -E "error|Error" "$W/build.log" | head -20; exit 1; }
LIB=$(find "$W/build" -name "libggml.so*" | head -1)
echo "built $LIB"
nm -D --defined-only "$LIB" | grep -c " T ggml_tmac_" | xargs echo "ggml_tmac_* symbols defined in libggml:"
nm -D --undefined-only "$LIB" | grep -E "ggml_tmac_|tmac_b200" | awk '{print $2}' | tr '\n' ' '; echo "<- resolved by libtmac_b200.so"
ldd "$LIB" | grep tmac_b200 || true
