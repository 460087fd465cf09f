#!/bin/bash
# CPU-side sanitizer pass over the native host code (no GPU): libmrt_host.so under ASan + UBSan through the loader / export / builder
# tests, the SAH builder's native checker under ASan + UBSan and under TSan, and the oracle under TSan. The in-tree libraries are put back afterwards.
set -e
cd "$(dirname "$0")/.."
OUT=gpurun_out/sanitize; mkdir -p $OUT
g++ -O1 -g -std=c++17 -ffp-contract=off -fno-fast-math -fPIC -shared -fsanitize=address,undefined -fno-omit-frame-pointer -pthread -I include \
    -o $OUT/libmrt_host.so mass_raytrace_b200/host/mrt_host.cpp
cp mass_raytrace_b200/libmrt_host.so $OUT/libmrt_host.so.orig
trap 'cp $OUT/libmrt_host.so.orig mass_raytrace_b200/libmrt_host.so; touch mass_raytrace_b200/libmrt_host.so' EXIT
cp $OUT/libmrt_host.so mass_raytrace_b200/libmrt_host.so
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
    python -m pytest tests/test_image_export.py tests/test_obj_loader.py tests/test_ply_loader.py tests/test_host_vs_oracle.py tests/test_oracle_kat.py tests/test_oracle_render.py \
    -q -p no:cacheprovider 2>&1 | tee $OUT/host_asan.log | tail -2
if grep -q "runtime error\|AddressSanitizer" $OUT/host_asan.log; then echo "sanitizer reports in $OUT/host_asan.log"; exit 1; fi
for san in address,undefined thread; do
    g++ -O1 -g -std=c++17 -pthread -fsanitize=$san -fno-omit-frame-pointer -I mass_raytrace_b200/csrc -o $OUT/check_bvh_$$ \
        tests/native/check_bvh_build.cpp mass_raytrace_b200/csrc/mrt_bvh_build.cpp
    $OUT/check_bvh_$$ > $OUT/bvh_$san.log 2>&1 || { echo "builder under -fsanitize=$san failed: $OUT/bvh_$san.log"; exit 1; }
    if grep -q "runtime error\|Sanitizer" $OUT/bvh_$san.log; then echo "sanitizer reports in $OUT/bvh_$san.log"; exit 1; fi
    rm -f $OUT/check_bvh_$$
done
# the oracle (test infrastructure, but it decides parity): its threaded tree build and both render modes under TSan
g++ -O1 -g -std=c++17 -ffp-contract=off -fno-fast-math -fPIC -pthread -fsanitize=thread -shared -o $OUT/liboracle_tsan.so oracle/oracle.cpp
cp $OUT/libmrt_host.so.orig mass_raytrace_b200/libmrt_host.so; touch mass_raytrace_b200/libmrt_host.so
cp oracle/liboracle.so $OUT/liboracle.so.orig
trap 'cp $OUT/libmrt_host.so.orig mass_raytrace_b200/libmrt_host.so; touch mass_raytrace_b200/libmrt_host.so; cp $OUT/liboracle.so.orig oracle/liboracle.so; touch oracle/liboracle.so' EXIT
cp $OUT/liboracle_tsan.so oracle/liboracle.so
{ LD_PRELOAD="$(gcc -print-file-name=libtsan.so)" TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" python -m pytest tests/test_oracle_render.py tests/test_oracle_kat.py -q -p no:cacheprovider
  LD_PRELOAD="$(gcc -print-file-name=libtsan.so)" TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" python tools/oracle_tsan_big.py; } > $OUT/oracle_tsan.log 2>&1 || { echo "oracle under TSan failed: $OUT/oracle_tsan.log"; exit 1; }
if grep -q "ThreadSanitizer" $OUT/oracle_tsan.log; then echo "sanitizer reports in $OUT/oracle_tsan.log"; exit 1; fi
echo "host code clean under ASan, UBSan and TSan; oracle clean under TSan"
