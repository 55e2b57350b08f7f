#!/bin/sh
# Compiles the reference's libfwk camera/frustum math from the sources where they lie
# (never copied into this repository) into oracle/_ref/ref_camera.
#   usage: build_ref.sh /root/reference oracle/_ref/ref_camera
# The error/IO paths of libfwk (asserts, XML load/save, text formatting) are not linked: they are
# never reached by the calls ref_camera.cpp makes, so unresolved symbols are left unresolved.
set -e
REF="$1"; OUT="$2"; FWK="$REF/libfwk"
HERE="$(dirname "$0")"
mkdir -p "$(dirname "$OUT")"
CXX=/usr/bin/g++; [ -x "$CXX" ] || CXX=g++
$CXX -std=c++20 -O1 -DNDEBUG -DFWK_DWARF_DISABLED -I "$FWK/include" "$HERE/ref_camera.cpp" \
  "$FWK/src/math/matrix4.cpp" "$FWK/src/math/matrix3.cpp" "$FWK/src/math/frustum.cpp" \
  "$FWK/src/math/plane.cpp" "$FWK/src/math/ray.cpp" "$FWK/src/math/rotation.cpp" "$FWK/src/math/base.cpp" \
  "$FWK/src/gfx/camera.cpp" "$FWK/src/gfx/orbiting_camera.cpp" \
  -Wl,--unresolved-symbols=ignore-all -o "$OUT"
