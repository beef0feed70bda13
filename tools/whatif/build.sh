#!/bin/bash
# build.sh NAME [-D...]: one what-if binary tools/whatif/whatif_NAME; prints registers / spills
HERE="$(cd "$(dirname "$0")" && pwd)"
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xptxas -v -DW_NAME="\"$name\"" "$@" \
    -o "$HERE/whatif_$name" "$HERE/whatif.cu" 2>&1 | grep -E "Used|spill" | tr '\n' ' ' | sed "s/^/$name: /"
echo
