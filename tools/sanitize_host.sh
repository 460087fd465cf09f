#!/bin/bash
# CPU-side sanitizer pass over the native host code (no GPU): libmrt_host.so under ASan + UBSan through the loader / export / builder
# tests, and the SAH builder's native checker under ASan + UBSan and under TSan. The in-tree library is put back afterwards.
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
echo "host code clean under ASan, UBSan and TSan"
