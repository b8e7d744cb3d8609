#!/bin/bash
# Build a variant of liblpsr_b200.so out of tree (A/B measurements): tools/build_variant.sh NAME "EXTRA NVCC FLAGS" -> ab/lib_NAME.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
V=/tmp/variants/$1
rm -rf $V; mkdir -p $V
cp -r "$ROOT/license-plate-detection-and-recognition-with-image-enhancement_b200" $V/pkg
cp -r "$ROOT/include" $V/include
rm -rf $V/pkg/build $V/pkg/liblpsr_b200.so
(cd $V/pkg && LPSR_NVCC_EXTRA="$2" python build_ext.py --force > $V/build.log 2>&1)
mkdir -p "$ROOT/ab"; cp $V/pkg/liblpsr_b200.so "$ROOT/ab/lib_$1.so"
echo "built ab/lib_$1.so"
