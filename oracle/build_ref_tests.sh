#!/usr/bin/env bash
# Compile the reference's own batch-path test programs, from the sources where they lie under /root/reference/tests,
# against THIS repo's include/ and libcuda_zstd_b200.so.  Nothing of the reference is linked: the programs exercise the
# drop-in boundary only (ZstdBatchManager, NvcompV5BatchManager, the extern "C" calls, the inference API).
# Outputs go to oracle/_ref/tests/ (git-ignored, travels to the GPU box); tests/test_gpu_reference_suite.py runs them.
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${REF:-/root/reference}"
OUT="$ROOT/oracle/_ref/tests"
LIBDIR="$ROOT/custom-nvcomp-with-zstd_b200"
[ -d "$REF/tests" ] || { echo "no reference tree at $REF: nothing to build"; exit 0; }
[ -f "$LIBDIR/libcuda_zstd_b200.so" ] || { echo "build the library first"; exit 1; }
mkdir -p "$OUT"
# every reference test that includes only the public manager / nvcomp / types / safe_alloc headers
TESTS="test_c_api.cpp test_c_api_edge_cases.cu test_compressible_data.cu test_concurrency_repro.cu test_correctness.cu
test_extended_validation.cu test_gpu_bitstream.cu test_inference_api.cu test_lz77_comprehensive.cu test_metadata_roundtrip.cu
test_nvcomp_batch.cu test_nvcomp_interface.cu test_parallel_compression.cu test_rfc8878_compliance.cu test_roundtrip.cu
test_scale_repro.cu test_two_phase_unit.cu test_pipeline_integration.cu"
# ... and the reference's benchmark programs of the same path (run by hand: tools/run_ref_benchmarks.sh)
BENCHES="benchmark_batch_throughput.cu benchmark_nvcomp_interface.cu benchmark_c_api.cu benchmark_block_size.cu"
ok=0; bad=0
build_bench() {
  local f="$1" t="${1%.*}"
  if [ "$OUT/$t" -nt "$REF/benchmarks/$f" ] && [ "$OUT/$t" -nt "$LIBDIR/libcuda_zstd_b200.so" ]; then return 0; fi
  nvcc -std=c++17 -x cu -O2 -Xcompiler -fopenmp -gencode arch=compute_100a,code=sm_100a -I"$ROOT/include" -I"$REF/benchmarks" -I"$REF/tests" \
       "$REF/benchmarks/$f" -o "$OUT/$t" -L"$LIBDIR" -lcuda_zstd_b200 -lgomp \
       -Xlinker -rpath -Xlinker '$ORIGIN/../../../custom-nvcomp-with-zstd_b200' > "$OUT/$t.build.log" 2>&1
}
build_one() {
  local f="$1" t="${1%.*}"
  if [ "$OUT/$t" -nt "$REF/tests/$f" ] && [ "$OUT/$t" -nt "$LIBDIR/libcuda_zstd_b200.so" ]; then return 0; fi
  nvcc -std=c++17 -x cu -O1 -gencode arch=compute_100a,code=sm_100a -I"$ROOT/include" -I"$REF/tests" "$REF/tests/$f" -o "$OUT/$t" \
       -L"$LIBDIR" -lcuda_zstd_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../../custom-nvcomp-with-zstd_b200' > "$OUT/$t.build.log" 2>&1
}
pids=()
for f in $TESTS; do build_one "$f" & pids+=($!); if [ ${#pids[@]} -ge 6 ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi; done
for f in $BENCHES; do build_bench "$f" & done
wait
for f in $TESTS $BENCHES; do t="${f%.*}"; if [ -x "$OUT/$t" ]; then ok=$((ok+1)); else bad=$((bad+1)); echo "FAILED to build $t (see $OUT/$t.build.log)"; fi; done
echo "reference test programs built against the drop-in: $ok ok, $bad failed"
[ "$bad" -eq 0 ]
