#!/bin/bash
# integration/build.sh -- builds the reference-side binding where a JDK exists: libgsjni.so (the JNI shim over
# libgenestrip_b200.so) and the Java classes under integration/java against the reference's jar.
#
#   GENESTRIP_JAR=/path/to/genestrip-core.jar[:more.jar] JAVA_HOME=/path/to/jdk bash integration/build.sh
#
# This repository's image has no JDK: the script says so and exits 0 (the C side is covered by tests/test_capi_symbols.py:
# integration/c/call_order.c runs the call order, gs_jni.cpp is type-checked against tests/jni_stub/jni.h, and every native
# method of GsNative.java is matched against the shim's exports).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
OUT="${OUT:-$HERE/_build}"
JAVAC="$(command -v javac || true)"
[ -n "$JAVA_HOME" ] && [ -x "$JAVA_HOME/bin/javac" ] && JAVAC="$JAVA_HOME/bin/javac"
if [ -z "$JAVAC" ]; then
  echo "integration/build.sh: no javac on PATH and no JAVA_HOME -- nothing built (expected in the build image)"; exit 0
fi
[ -z "$JAVA_HOME" ] && JAVA_HOME="$(dirname "$(dirname "$(readlink -f "$JAVAC")")")"
if [ ! -f "$JAVA_HOME/include/jni.h" ]; then
  echo "integration/build.sh: $JAVA_HOME/include/jni.h not found -- nothing built"; exit 0
fi
mkdir -p "$OUT/classes"
[ -f "$ROOT/genestrip_b200/_lib/libgenestrip_b200.so" ] || python "$ROOT/genestrip_b200/build.py"
g++ -O2 -std=c++17 -fPIC -shared -I"$JAVA_HOME/include" -I"$JAVA_HOME/include/linux" -I"$ROOT/include" "$HERE/jni/gs_jni.cpp" \
    -L"$ROOT/genestrip_b200/_lib" -lgenestrip_b200 -Wl,-rpath,"$ROOT/genestrip_b200/_lib" -o "$OUT/libgsjni.so"
echo "built $OUT/libgsjni.so"
if [ -z "$GENESTRIP_JAR" ]; then
  echo "integration/build.sh: GENESTRIP_JAR not set -- GsNative.java only (the other classes extend the reference's)"
  "$JAVAC" -d "$OUT/classes" "$HERE/java/org/metagene/genestrip/gpu/GsNative.java"
else
  "$JAVAC" -cp "$GENESTRIP_JAR" -d "$OUT/classes" $(find "$HERE/java" -name '*.java')
fi
echo "classes in $OUT/classes; run with -Djava.library.path=$OUT"
