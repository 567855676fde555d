#!/usr/bin/env bash
# Run the reference's own benchmark programs (built against this library by oracle/build_ref_tests.sh) on the GPU box.
# usage (on the box): tools/run_ref_benchmarks.sh > gpurun_out/ref_benchmarks.txt
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
BIN="$ROOT/oracle/_ref/tests"
export LD_LIBRARY_PATH="$ROOT/custom-nvcomp-with-zstd_b200:$LD_LIBRARY_PATH"
for b in benchmark_batch_throughput benchmark_nvcomp_interface benchmark_c_api benchmark_block_size; do
  echo "===== $b (reference source, unchanged; linked against libcuda_zstd_b200.so) ====="
  if [ -x "$BIN/$b" ]; then (cd "$BIN" && timeout 300 "./$b" 2>&1 | tail -60); echo "exit: $?"; else echo "not built"; fi
done
