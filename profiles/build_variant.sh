#!/bin/bash
# usage: profiles/build_variant.sh <name> "<extra nvcc flags>"   ->  elimaloc_b200/lib_<name>.so (git-ignored, travels with gpurun)
# then, under gpurun:  bash profiles/quick_bench.sh "" "@elimaloc_b200/lib_<name>.so "   (A/B against the default build)
# and for parity:      ELIMALOC_B200_LIB=elimaloc_b200/lib_<name>.so python -m pytest tests -m gpu -x -q
# e.g. profiles/build_variant.sh greedy "-DELM_GREEDY_ITEMS"
set -e
n="$1"; shift
make -C "$(dirname "$0")/../elimaloc_b200/csrc" -j8 -s OUT=../lib_$n.so OBJDIR=../../build/obj_$n EXTRA_NVFLAGS="$*"
ls -la "$(dirname "$0")/../elimaloc_b200/lib_$n.so"
