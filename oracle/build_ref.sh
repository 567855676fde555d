#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds the UNMODIFIED reference from the sources where they lie under
# /root/reference into oracle/_ref/ (git-ignored, travels to the GPU box via gpurun):
#   oracle/_ref/libcuda_zstd_ref.a      all 29 translation units of /root/reference/src
#   oracle/_ref/libref_hybrid.so        oracle/ref_hybrid_driver.cpp + that archive + libzstd.so.1
# The reference's own CMake is NOT run; this is the short recipe SURVEY.md section 8(c) probed.
# No reference source is copied into the repo.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF/src" ] || { echo "build_ref: $REF absent; keeping prebuilt $OUT"; exit 0; }
mkdir -p "$OUT/obj"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr --expt-extended-lambda -Xcompiler -fPIC -w -I$REF/include -I$REF/src -I$HERE/shim"
JOBS=${JOBS:-6}
ls "$REF"/src/*.cu "$REF"/src/*.cpp | xargs -P "$JOBS" -I{} bash -c '
  f="{}"; o="'"$OUT"'/obj/$(basename "$f").o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ]; then '"$NVCC $FLAGS"' -x cu -c "$f" -o "$o"; fi'
rm -f "$OUT/libcuda_zstd_ref.a"
ar rcs "$OUT/libcuda_zstd_ref.a" "$OUT"/obj/*.o
$NVCC $FLAGS -shared -o "$OUT/libref_hybrid.so" "$HERE/ref_hybrid_driver.cpp" \
  -Xlinker --whole-archive "$OUT/libcuda_zstd_ref.a" -Xlinker --no-whole-archive \
  -Xlinker --allow-multiple-definition -l:libzstd.so.1 -lcudart
echo "build_ref: OK -> $OUT/libref_hybrid.so"
