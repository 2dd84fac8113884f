#!/bin/bash
# developer tool: rebuild plg_derivatives.cu with a few -D variants on the GPU box and time them
cd "$(dirname "$0")/.."
for v in "-DPLG_DERDNA_U=1 -DPLG_DERDNA_MINB=4" "-DPLG_DERDNA_U=2 -DPLG_DERDNA_MINB=1" "-DPLG_DERDNA_U=2 -DPLG_DERDNA_MINB=2" "-DPLG_DERDNA_U=2 -DPLG_DERDNA_MINB=3" "-DPLG_DERDNA_U=4 -DPLG_DERDNA_MINB=1" "-DPLG_DERDNA_U=1 -DPLG_DERDNA_MINB=2" "-DPLG_DERDNA_U=3 -DPLG_DERDNA_MINB=2"; do
  rm -f libpll_b200/csrc/build/plg_derivatives.cu.o
  make -C libpll_b200/csrc EXTRA="$v" > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  echo "== $v: $(grep -A2 'k_derivatives_dnaILi4E' libpll_b200/csrc/build/plg_derivatives.ptxas.txt | grep -E 'Used' | head -1)"
  python tools/newton_bench.py 2>&1 | grep -E "derivative pass ii|newton ii"
done
