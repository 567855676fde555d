#!/usr/bin/env bash
# Compile the reference's OWN Python extension module (python/src/binding.cpp, pybind11), from the source where it lies
# under /root/reference, against THIS repo's include/ and libcuda_zstd_b200.so.  Nothing of the reference's
# codec is linked.  binding.cpp itself calls ZSTD_getFrameContentSize (header parse in HybridEngine.decompress,
# binding.cpp:464-473): the module -- test infrastructure, not the product library -- links the system libzstd.so.1 for
# that one symbol through the declaration-only shim oracle/shim/zstd.h.
# Outputs go to oracle/_ref/python/ (git-ignored, travels to the GPU box): cuda_zstd/_core*.so plus a copy of the
# reference's pure-Python package files and its test_basic.py, which tests/test_gpu_reference_suite.py runs.
set -eu
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${REF:-/root/reference}"
OUT="$ROOT/oracle/_ref/python"
LIBDIR="$ROOT/custom-nvcomp-with-zstd_b200"
[ -f "$REF/python/src/binding.cpp" ] || { echo "no reference tree at $REF: nothing to build"; exit 0; }
[ -f "$LIBDIR/libcuda_zstd_b200.so" ] || { echo "build the library first"; exit 1; }
PY="${PYTHON:-python}"
EXT="$($PY -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
PYINC="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
PBINC="$($PY -c 'import pybind11; print(pybind11.get_include())')"
mkdir -p "$OUT/cuda_zstd" "$OUT/tests"
SO="$OUT/cuda_zstd/_core$EXT"
if [ ! -f "$SO" ] || [ "$REF/python/src/binding.cpp" -nt "$SO" ] || [ "$LIBDIR/libcuda_zstd_b200.so" -nt "$SO" ]; then
  g++ -std=c++17 -O2 -shared -fPIC -fvisibility=hidden -I"$ROOT/include" -I"$ROOT/oracle/shim" -I/usr/local/cuda/include -I"$PYINC" -I"$PBINC" \
      "$REF/python/src/binding.cpp" -o "$SO" -L"$LIBDIR" -lcuda_zstd_b200 -L/usr/local/cuda/lib64 -lcudart -l:libzstd.so.1 \
      -Wl,-rpath,'$ORIGIN/../../../../custom-nvcomp-with-zstd_b200' > "$OUT/build.log" 2>&1 || { tail -30 "$OUT/build.log"; exit 1; }
fi
cp "$REF"/python/cuda_zstd/__init__.py "$REF"/python/cuda_zstd/_core.pyi "$REF"/python/cuda_zstd/py.typed "$OUT/cuda_zstd/"
cp "$REF"/python/tests/__init__.py "$REF"/python/tests/conftest.py "$REF"/python/tests/test_basic.py "$OUT/tests/"
echo "reference Python module built against the drop-in: $SO"
