#!/usr/bin/env bash
# Compiles the reference's OWN C++ operators, from the sources where they lie
# under /root/reference/vinum_cpp/src, into oracle/_ref/ref_vinum_lib*.so.
# The reference's CMake build is not used (it needs network for gtest and pins
# Arrow 3.0); this is a direct g++ recipe.
# Outputs ONLY under oracle/_ref/ (git-ignored; travels to the GPU box).
# No reference source is copied into the repo.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${VINUM_REF_SRC:-/root/reference/vinum_cpp/src}"
OUT="$HERE/_ref"
if [ ! -d "$SRC" ]; then
  echo "build_ref.sh: $SRC absent (GPU box?) - keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
PY="${PYTHON:-python}"
# The reference's pure-Python layers (parser AST, binder, planner, executor, operators) as ONE zip archive
# next to the compiled operators: `bench.py --impl reference` and tests/test_gpu_dropin.py drive the
# reference's own QueryPlanner + RecursiveExecutor on the GPU box, where /root/reference does not exist.
# A build artefact like the .so (git-ignored, never edited, never imported by vinum_b200/).
PYREF="$OUT/vinum_pyref.zip"
REFPY="$(dirname "$(dirname "$SRC")")/vinum"
if [ -d "$REFPY" ] && { [ ! -f "$PYREF" ] || [ -n "$(find "$REFPY" -name '*.py' -newer "$PYREF" | head -1)" ]; }; then
  $PY - "$REFPY" "$PYREF" <<'PYEOF'
import sys, zipfile, pathlib
src, dst = pathlib.Path(sys.argv[1]), sys.argv[2]
with zipfile.ZipFile(dst, "w", zipfile.ZIP_DEFLATED) as z:
    for f in sorted(src.rglob("*.py")):
        rel = f.relative_to(src.parent)
        if "tests" in rel.parts:
            continue
        z.write(f, str(rel))
PYEOF
  echo "build_ref.sh: packed $PYREF"
fi
PA=$($PY -c "import pyarrow as pa;print(pa.get_library_dirs()[0])")
PAINC=$($PY -c "import pyarrow as pa;print(pa.get_include())")
PYINC=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
PBINC=$($PY -c "import pybind11;print(pybind11.get_include())")
NPINC=$($PY -c "import numpy;print(numpy.get_include())")
EXT=$($PY -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
TARGET="$OUT/ref_vinum_lib$EXT"
# GenericHashAggregate (string / any-type keys): generic_hash_aggregate.h:37 calls
# Scalar::Equals(shared_ptr<Scalar>), an overload Arrow dropped after 3.0.  The header and its
# .cpp are copied into the (git-ignored) build directory with that ONE token patched
# (`*iter_two` -> `**iter_two`), compiled from there and deleted again; nothing else changes.
PATCHED="$OUT/obj/patched"
mkdir -p "$PATCHED"
sed 's/Equals(\*iter_two)/Equals(**iter_two)/' "$SRC/operators/aggregate/generic_hash_aggregate.h" > "$PATCHED/generic_hash_aggregate.h"
cp "$SRC/operators/aggregate/generic_hash_aggregate.cpp" "$PATCHED/generic_hash_aggregate.cpp"
FLAGS="-O2 -std=c++20 -fPIC -fvisibility=hidden -w -include $HERE/compat.h \
  -I$PATCHED -I$PAINC -I$SRC -I$SRC/operators -I$SRC/operators/aggregate -I$SRC/operators/sort \
  -I$PYINC -I$PBINC -I$NPINC"
FILES="$HERE/ref_wrapper.cpp \
  $SRC/common/huge_int.cpp $SRC/common/array_iterators.cpp \
  $SRC/operators/aggregate/agg_func_factory.cpp $SRC/operators/aggregate/base_aggregate.cpp \
  $SRC/operators/aggregate/one_group_aggregate.cpp \
  $SRC/operators/aggregate/single_numerical_hash_aggregate.cpp \
  $SRC/operators/aggregate/multi_numerical_hash_aggregate.cpp \
  $PATCHED/generic_hash_aggregate.cpp \
  $SRC/operators/sort/sort.cpp $SRC/operators/table_batch_reader.cpp"
newest_src=$(stat -c %Y $HERE/ref_wrapper.cpp $SRC/common/*.cpp $SRC/operators/aggregate/*.cpp $SRC/operators/aggregate/*.h \
  $SRC/operators/sort/sort.cpp "$HERE/compat.h" "$HERE/build_ref.sh" | sort -n | tail -1)
if [ -f "$TARGET" ] && [ "$(stat -c %Y "$TARGET")" -ge "$newest_src" ]; then
  rm -rf "$OUT/obj"; echo "build_ref.sh: $TARGET up to date"; exit 0
fi
pids=()
objs=()
for f in $FILES; do
  o="$OUT/obj/$(basename "${f%.cpp}").o"
  objs+=("$o")
  g++ $FLAGS -c "$f" -o "$o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
g++ -shared "${objs[@]}" -L"$PA" -l:libarrow_python.so.2400 -l:libarrow_compute.so.2400 \
    -l:libarrow.so.2400 -Wl,-rpath,"$PA" -o "$TARGET"
rm -rf "$OUT/obj"
echo "build_ref.sh: built $TARGET"
