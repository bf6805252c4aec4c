#!/bin/bash
# Builds a variant of the library with extra compile definitions, for same-box A/B runs (tools/ab_gpu.sh):
#   tools/build_variant.sh <name> [-DFOO=1 ...]   ->  variants/lib<name>.so   (variants/ is git-ignored)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d /tmp/ft8b_variant_XXXX)
mkdir -p $tmp/ft8_demodulator_b200 $tmp/include $root/variants
cp -r $root/ft8_demodulator_b200/csrc $tmp/ft8_demodulator_b200/
cp $root/include/ft8b200.h $tmp/include/
rm -f $tmp/ft8_demodulator_b200/csrc/*.o
make -C $tmp/ft8_demodulator_b200/csrc -j8 LIB=$root/variants/lib$name.so \
  NVCCFLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v $*" > $tmp/build.log 2>&1 || { tail -30 $tmp/build.log; exit 1; }
grep -h -A2 "stft1920_kernelIfLi3ELb1\|sync_scores_pair_kernelILi2ELb1\|13decode_kernel" $tmp/ft8_demodulator_b200/csrc/*.ptxas.log | grep -v "^--" | cut -c1-160 || true
rm -rf $tmp
echo "built variants/lib$name.so"
