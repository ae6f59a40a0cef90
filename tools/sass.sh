#!/bin/bash
# tools/sass.sh FILE.cu KERNEL_SUBSTRING [extra nvcc flags] -> /tmp/t/k.sass (first matching kernel), prints registers
cd /root/repo/dimo_b200/csrc || exit 1
mkdir -p /tmp/t
F=$1; K=$2; shift 2
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I ../../include --expt-relaxed-constexpr "$@" -Xptxas -v -c $F -o /tmp/t/k.o 2> /tmp/t/k.log || { cat /tmp/t/k.log | grep -v "^ptxas info" | head -30; exit 1; }
grep -E "Compiling|Used|spill" /tmp/t/k.log | paste - - - | sed 's/ptxas info    : //g' | grep "$K" | sed -E "s/Compiling entry function '([^']*)' for 'sm_100a'/\1/" | cut -c1-260
SYM=$(cuobjdump -elf /tmp/t/k.o 2>/dev/null | grep -o "_Z[A-Za-z0-9_]*$K[A-Za-z0-9_]*" | sort -u | head -${3:-1} | tail -1)
SYM=$(grep -o "_Z[A-Za-z0-9_]*" /tmp/t/k.log | grep "$K" | sort -u | head -1)
cuobjdump -sass -fun "$SYM" /tmp/t/k.o | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##' > /tmp/t/k.sass
echo "$SYM: $(wc -l < /tmp/t/k.sass) SASS lines -> /tmp/t/k.sass"
