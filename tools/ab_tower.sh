#!/bin/bash
# Builds variant libraries of the tower kernel for A/B timing on the GPU box:
#   tools/ab_tower.sh name1 "-DFLAG=1" name2 "-DOTHER=2" ...   ->  rl-selfplay-mnk_b200/build/variants/lib_<name>.so
# then on the box:  MNK_LIB=rl-selfplay-mnk_b200/build/variants/lib_<name>.so python tools/time_tower.py
set -e
cd "$(dirname "$0")/.."
P=rl-selfplay-mnk_b200
VAR=${VAR_DIR:-$P/build/variants}; mkdir -p $VAR $P/build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --std=c++17 -Xcompiler -fPIC -Iinclude $flags \
       -c $P/csrc/${SRC:-mnk_resnet}.cu -o $P/build/variants/${SRC:-mnk_resnet}_$name.o
  objs=$(ls $P/build/*.o | grep -v "/${SRC:-mnk_resnet}.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $VAR/lib_$name.so $objs $P/build/variants/${SRC:-mnk_resnet}_$name.o
  echo "built lib_$name.so [$flags]"
done
