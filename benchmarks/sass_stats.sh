#!/bin/bash
# compile one csrc file for sm_100a, print ptxas register/spill lines for kernels matching $2 and the
# SASS opcode histogram of the first matching kernel:   benchmarks/sass_stats.sh qgemm_sm100.cu 'Li256ELi7E'
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/numpy_quant_b200/csrc/$1
PAT=${2:-.}
OBJ=/tmp/sass_$$.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -Xptxas -v -c "$SRC" -o $OBJ 2> /tmp/sass_$$.log || { grep -E "error" -A3 /tmp/sass_$$.log | head -40; exit 1; }
grep -E "warning" -A2 /tmp/sass_$$.log | head -20 || true
grep -E "Compiling entry function '[^']*$PAT" -A3 /tmp/sass_$$.log | grep -E "Compiling|registers|spill" | sed 's/ptxas info    ://' | paste - - - | cut -c1-260
FN=$(grep -oE "Compiling entry function '[^']*$PAT[^']*'" /tmp/sass_$$.log | head -1 | sed "s/.*'\(.*\)'/\1/")
cuobjdump -sass -fun "$FN" $OBJ > /tmp/last.sass
echo "SASS lines: $(wc -l < /tmp/last.sass)  ($FN)"
grep -oE "^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T]+ +)?[A-Z0-9_.]+" /tmp/last.sass | awk '{print $NF}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -${3:-24} | paste - - - -
rm -f $OBJ
