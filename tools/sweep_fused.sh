#!/bin/bash
# developer tool: rebuild plg_traverse.cu with -D variants on the GPU box and time the fused traversal
cd "$(dirname "$0")/.."
run() {
  rm -f libpll_b200/csrc/build/plg_traverse.cu.o
  make -C libpll_b200/csrc EXTRA="$1" > /dev/null 2>&1 || { echo "build failed: $1"; return; }
  echo "== $1 slots=${2:-4}: $(grep -A2 'k_traverse_dnaILi4' libpll_b200/csrc/build/plg_traverse.ptxas.txt | grep -E 'Used' | head -1 | sed 's/ptxas info    : //')"
  PLL_GPU_FUSED_SLOTS=${2:-4} timeout 300 python tools/quick_bench.py --tips 1000 --sites 400000 --fast-tips 2>&1 | tail -1 | cut -c1-110
}
for v in "$@"; do
  IFS='|' read -r flags slots <<< "$v"
  run "$flags" "$slots"
done
